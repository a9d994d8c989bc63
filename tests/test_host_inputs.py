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
