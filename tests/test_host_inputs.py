"""Small host-side inputs of `bayesTyper genotype`: the samples file (Sample.cpp:38-70) and chromosome ploidy
(ChromosomePloidy.cpp:40-185)."""
import pytest

from bayestyper_b200 import ploidy


def test_samples_file(tmp_path):
    p = tmp_path / "samples.tsv"
    p.write_text("NA12878\tF\t/data/NA12878\nNA12891\tMale\tkmc/NA12891\nNA12892\tFemale\tx\n")
    assert ploidy.read_samples(p) == [("NA12878", "F", "/data/NA12878"), ("NA12891", "M", "kmc/NA12891"), ("NA12892", "F", "x")]
    p.write_text("a\tF\n")
    with pytest.raises(ValueError, match="three tab-separated"):
        ploidy.read_samples(p)
    p.write_text("a\tfemale\tz\n")
    with pytest.raises(ValueError, match="Gender|gender"):
        ploidy.read_samples(p)


def test_default_ploidy_rules():
    cp = ploidy.ChromosomePloidy(["chr1", "chrX", "Y", "chrY", "x", "chrUn_decoy"], ["F", "M", "M"], decoys=["chrUn_decoy"])
    assert cp.gender_ploidy("chr1") == (2, 2) and cp.sample_ploidy("chr1") == [2, 2, 2]
    assert cp.gender_ploidy("chrX") == (2, 1) and cp.sample_ploidy("chrX") == [2, 1, 1]
    assert cp.gender_ploidy("x") == (2, 1)
    assert cp.gender_ploidy("Y") == (0, 1) and cp.sample_ploidy("chrY") == [0, 1, 1]
    with pytest.raises(KeyError):
        cp.gender_ploidy("chrUn_decoy")


def test_ploidy_file(tmp_path):
    p = tmp_path / "ploidy.tsv"
    p.write_text("chr1\t2\t2\nchrX\t1\t1\nchrM\t1\t1\nchrW\t0\t2\n")
    cp = ploidy.ChromosomePloidy(["chr1", "chrX", "chrM"], ["F", "M"], p)
    assert cp.sample_ploidy("chrX") == [1, 1] and cp.gender_ploidy("chrM") == (1, 1)
    with pytest.raises(ValueError, match="does not appear"):
        ploidy.ChromosomePloidy(["chr2"], ["F"], p)
    p.write_text("chr1\t2\t3\n")
    with pytest.raises(ValueError, match="between zero and two"):
        ploidy.ChromosomePloidy(["chr1"], ["F"], p)
    p.write_text("chr1\t2\t2\nchr1\t2\t2\n")
    with pytest.raises(ValueError, match="multiple times"):
        ploidy.ChromosomePloidy(["chr1"], ["F"], p)
    p.write_text("chr1 2 2\n")
    with pytest.raises(ValueError, match="three tab-separated"):
        ploidy.ChromosomePloidy(["chr1"], ["F"], p)


def test_cluster_data_files_of_the_reference(tmp_path):
    """`<out>_cluster_data/intercluster_regions.txt.gz` and `parameter_kmers.fa.gz` as the reference's cluster stage wrote them
    (tests/golden/make_cluster_data_fixture.py, oracle-R) read back; the regions equal graph_builder's; every parameter k-mer is a
    canonical k-mer of an intercluster region; writing reproduces the files' content."""
    import gzip
    import shutil

    import numpy as np

    from bayestyper_b200 import cluster_data, graph_builder, synth
    from tests._fixtures import GOLD
    from tests.golden.make_fixtures import PIPE_WORKLOADS
    w = PIPE_WORKLOADS["pipe_mixed_3s"]()
    for name in ("intercluster_regions.txt.gz", "parameter_kmers.fa.gz"):
        shutil.copy(GOLD / f"cluster_data_mixed_3s.{name}", tmp_path / name)
    regions = cluster_data.read_intercluster_regions(tmp_path / "intercluster_regions")
    mine = graph_builder.intercluster_regions(w.chrom, w.reference, w.variants)
    assert sorted((a, b) for _, _, a, b in regions) == sorted(mine)
    assert all(c == w.chrom and d is False for c, d, _, _ in regions)
    lengths = [b - a for _, _, a, b in regions]
    assert lengths == sorted(lengths, reverse=True)
    cluster_data.write_intercluster_regions(tmp_path / "again", [(w.chrom, False, a, b) for a, b in mine])
    again = cluster_data.read_intercluster_regions(tmp_path / "again")
    assert [b - a for _, _, a, b in again] == lengths and sorted(again) == sorted(regions)

    km = cluster_data.read_parameter_kmers(tmp_path / "parameter_kmers")
    assert km.shape == (400, 2)
    pool = np.concatenate([synth.canonical_kmers(w.reference[a:b + 1]) for a, b in mine])
    as_void = lambda x: np.ascontiguousarray(x).view([("a", np.uint64), ("b", np.uint64)]).ravel()
    assert np.isin(as_void(km), as_void(pool)).all()
    cluster_data.write_parameter_kmers(tmp_path / "pk_again", km)
    assert gzip.open(tmp_path / "pk_again.fa.gz").read() == gzip.open(tmp_path / "parameter_kmers.fa.gz").read()
    assert cluster_data.write_parameter_kmers(tmp_path / "capped", km, max_kmers=10) == 10
    assert len(cluster_data.read_parameter_kmers(tmp_path / "capped")) == 10
    with pytest.raises(ValueError, match="ACGT"):
        cluster_data.strings_to_kmers(np.frombuffer(b"N" * 55, np.uint8))
