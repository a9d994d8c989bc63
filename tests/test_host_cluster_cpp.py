"""host/btcluster (include/btgpu_cluster.hpp, the native cluster / group / graph builder) against the unit the REFERENCE built
(oracle-R fixtures: adversarial single-contig candidate sets, a genome with several contigs and decoys) and against the Python
builder on the same files."""
import gzip
import subprocess
from pathlib import Path

import numpy as np
import pytest

from bayestyper_b200 import btd, graph_builder, synth, vcfio
from tests._fixtures import GOLD

ROOT = Path(__file__).resolve().parent.parent
KEYS = ("group_cluster_off", "group_nvar", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst", "cluster_idx",
        "cl_vertex_off", "cl_var_off", "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_nested", "v_refvar_off", "v_in_off", "v_in_src",
        "var_pos", "var_dep", "var_nalt", "var_alt_off", "alt_reflen", "alt_seq_off", "alt_seq")


def _exe():
    exe = ROOT / "host" / "btcluster"
    src = [ROOT / "host" / "btcluster.cpp", ROOT / "include" / "btgpu_cluster.hpp", ROOT / "host" / "btd.hpp"]
    if not exe.exists() or any(s.stat().st_mtime > exe.stat().st_mtime for s in src):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), "-I", str(ROOT / "host"), str(src[0]), "-lz", "-o", str(exe)])
    return exe


def _write_fasta(path, contigs):
    with open(path, "wb") as f:
        for n, s in contigs.items():
            f.write(b">" + n.encode() + b" description\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + b"\n")


def _write_vcf(path, cand, gz=False):
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
    for n, var in cand.items():
        for i, v in enumerate(var):
            info = "ACO=" + ",".join(v.aco) if getattr(v, "aco", None) else "."
            lines.append(f"{n}\t{v.pos + 1}\t{getattr(v, 'id', None) or f'{n}_{i}'}\t{v.ref.decode()}\t{','.join(a.decode() for a in v.alts)}\t.\t.\tDP=3;{info}")
    data = ("\n".join(lines) + "\n").encode()
    Path(path).write_bytes(gzip.compress(data) if gz else data)


def _run(tmp_path, genome, cand, decoys=None, gz=False):
    _write_fasta(tmp_path / "genome.fa", genome)
    vcf = tmp_path / ("c.vcf.gz" if gz else "c.vcf")
    _write_vcf(vcf, cand, gz)
    cmd = [str(_exe()), str(tmp_path / "genome.fa"), str(vcf), str(tmp_path / "out.btd")]
    if decoys:
        _write_fasta(tmp_path / "decoy.fa", decoys)
        cmd += ["--decoy", str(tmp_path / "decoy.fa")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return btd.read(tmp_path / "out.btd")


@pytest.mark.parametrize("golden", ["graphs_adversarial", "graphs_deep"])
def test_adversarial_candidate_sets_identical_to_reference(tmp_path, golden):
    d = btd.read(GOLD / f"{golden}.btd")
    for c in range(int(d["meta.n_cases"][0])):
        ref = bytes(d[f"c{c}.reference"])
        alleles = bytes(d[f"c{c}.alleles"]).split(b"\n")
        var = [synth.Variant(int(p), al.split(b",")[0], al.split(b",")[1:]) for p, al in zip(d[f"c{c}.var_pos"].tolist(), alleles)]
        b = _run(tmp_path, {"chrF": ref}, {"chrF": var}, gz=bool(c % 2))
        for k in KEYS:
            g = d[f"c{c}.g.{k}"]
            assert len(b[k]) == len(g) and (np.asarray(b[k]) == g).all(), (c, k)
        gv, off = d[f"c{c}.g.v_refvar"], d[f"c{c}.g.v_refvar_off"]
        for v in range(len(off) - 1):
            assert set(b["v_refvar"][int(off[v]):int(off[v + 1])].tolist()) == set(gv[int(off[v]):int(off[v + 1])].tolist())
        assert sorted((int(x), int(y)) for _, _, x, y in b["regions"]) == sorted((int(x), int(y)) for x, y in d[f"c{c}.regions"])


def test_genome_identical_to_reference_and_to_the_python_builder(tmp_path):
    d = btd.read(GOLD / "graphs_genome.btd")
    names = bytes(d["meta.contigs"]).decode().split("\n")
    n_decoys = int(d["meta.n_decoys"][0])
    seqs = {n: bytes(d[f"seq.{n}"]) for n in names}
    genome = {n: seqs[n] for n in names[:-n_decoys]}
    decoys = {n: seqs[n] for n in names[-n_decoys:]}
    cand = {}
    for n in bytes(d["meta.cand_contigs"]).decode().split("\n"):
        alleles = bytes(d[f"cand.{n}.alleles"]).split(b"\n")
        cand[n] = [synth.Variant(int(p), al.split(b",")[0], al.split(b",")[1:]) for p, al in zip(d[f"cand.{n}.pos"].tolist(), alleles)]
    b = _run(tmp_path, genome, cand, decoys)
    for k in KEYS:
        assert len(b[k]) == len(d["g." + k]) and (np.asarray(b[k]) == d["g." + k]).all(), k
    got_names = bytes(b["contig_names"]).decode().split("\n")
    chroms = bytes(d["g.chroms"])
    ref_names = [chroms[int(a):int(c)].decode() for a, c in zip(d["g.chrom_off"][:-1], d["g.chrom_off"][1:])]
    assert [got_names[i] for i in b["group_contig"]] == ref_names
    want = sorted(tuple(ln.split("\t")) for ln in bytes(d["regions"]).decode().split("\n"))
    assert sorted((got_names[c], str(int(f)), str(int(x)), str(int(y))) for c, f, x, y in b["regions"]) == want
    # and key for key against the Python builder reading the same files
    py = graph_builder.build_genome_graphs({**vcfio.read_fasta(tmp_path / "genome.fa"), **vcfio.read_fasta(tmp_path / "decoy.fa")},
                                           vcfio.read_candidates(tmp_path / "c.vcf"), decoys=list(decoys))
    for k in KEYS + ("v_refvar", "group_start", "group_end", "group_contig", "var_contig", "var_input_idx"):
        assert (np.asarray(b[k]) == np.asarray(py[k])).all(), k
    assert (b["regions"] == py["regions"]).all()
    ids = bytes(b["var_ids"]).decode()
    assert [ids[int(a):int(c)] for a, c in zip(b["var_id_off"][:-1], b["var_id_off"][1:])] == \
        [vcfio.read_candidates(tmp_path / "c.vcf")[got_names[c]][int(i)].id for c, i in zip(py["var_contig"], py["var_input_idx"])]


def test_origins_ids_and_errors(tmp_path):
    from tests.golden.make_vcf_fixtures import VCF_WORKLOADS
    w = VCF_WORKLOADS["vcf_nested_2s"]()
    b = _run(tmp_path, {w.chrom: w.reference}, {w.chrom: w.variants})
    py = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
    aco = bytes(b["alt_aco"]).decode()
    assert [aco[int(a):int(c)] for a, c in zip(b["alt_aco_off"][:-1], b["alt_aco_off"][1:])] == py["alt_aco"] and any(py["alt_aco"])
    for k in KEYS:
        assert (np.asarray(b[k]) == np.asarray(py[k])).all(), k
    _write_vcf(tmp_path / "bad.vcf", {"chrGone": [synth.Variant(80, b"A", [b"C"])]})
    r = subprocess.run([str(_exe()), str(tmp_path / "genome.fa"), str(tmp_path / "bad.vcf"), str(tmp_path / "o.btd")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not hold" in r.stderr
    r = subprocess.run([str(_exe()), str(tmp_path / "genome.fa"), str(tmp_path / "genome.fa"), str(tmp_path / "o.btd")], capture_output=True, text=True)
    assert r.returncode == 1 and ".vcf" in r.stderr
    # malformed records are named, not reported as a bare std::stoul failure
    (tmp_path / "pos.vcf").write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n" + w.chrom + "\tabc\t.\tA\tC\t.\t.\t.\n")
    r = subprocess.run([str(_exe()), str(tmp_path / "genome.fa"), str(tmp_path / "pos.vcf"), str(tmp_path / "o.btd")], capture_output=True, text=True)
    assert r.returncode == 1 and "not a positive integer" in r.stderr
    (tmp_path / "nohdr.vcf").write_text(w.chrom + "\t5\t.\tA\tC\t.\t.\t.\n")
    r = subprocess.run([str(_exe()), str(tmp_path / "genome.fa"), str(tmp_path / "nohdr.vcf"), str(tmp_path / "o.btd")], capture_output=True, text=True)
    assert r.returncode == 1 and "#CHROM" in r.stderr


def test_native_builder_behind_the_python_interface():
    from tests.golden.make_vcf_genome_fixture import genome_workload
    parts, empty, decoys = genome_workload()
    genome = {n: w.reference for n, w in parts.items()}
    genome[empty[0]] = empty[1]
    full = {**genome, **decoys}
    cand = {n: w.variants for n, w in parts.items()}
    a = graph_builder.build_genome_graphs(full, cand, decoys=list(decoys))
    b = graph_builder.build_genome_graphs_native(full, cand, decoys=list(decoys))
    for k, v in a.items():
        if isinstance(v, list):
            assert b[k] == v, k
        else:
            assert np.asarray(b[k]).dtype == np.asarray(v).dtype and (np.asarray(b[k]) == np.asarray(v)).all(), k
    with pytest.raises(ValueError, match="sorted by position"):
        graph_builder.build_genome_graphs_native({"c": b"A" * 300}, {"c": [synth.Variant(100, b"A", [b"C"]), synth.Variant(90, b"A", [b"C"])]})


def test_inference_units_identical_to_reference(tmp_path):
    d = btd.read(GOLD / "graphs_units.btd")
    names = bytes(d["meta.contigs"]).decode().split("\n")
    n_decoys = int(d["meta.n_decoys"][0])
    seqs = {n: bytes(d[f"seq.{n}"]) for n in names}
    cand = {}
    for n in bytes(d["meta.cand_contigs"]).decode().split("\n"):
        alleles = bytes(d[f"cand.{n}.alleles"]).split(b"\n")
        cand[n] = [synth.Variant(int(p), al.split(b",")[0], al.split(b",")[1:]) for p, al in zip(d[f"cand.{n}.pos"].tolist(), alleles)]
    _write_fasta(tmp_path / "genome.fa", {n: seqs[n] for n in names[:-n_decoys]})
    _write_fasta(tmp_path / "decoy.fa", {n: seqs[n] for n in names[-n_decoys:]})
    _write_vcf(tmp_path / "c.vcf", cand)
    r = subprocess.run([str(_exe()), str(tmp_path / "genome.fa"), str(tmp_path / "c.vcf"), str(tmp_path / "out.btd"), "--decoy", str(tmp_path / "decoy.fa"),
                        "--min-unit-variants", str(int(d["meta.min_unit_variants"][0]))], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n_units = int(d["meta.n_units"][0])
    assert not (tmp_path / f"out_unit_{n_units + 1}.btd").exists()
    py_units, py_regions = graph_builder.build_genome_units(seqs, cand, decoys=names[-n_decoys:], min_unit_variants=int(d["meta.min_unit_variants"][0]))
    for u in range(n_units):
        b = btd.read(tmp_path / f"out_unit_{u + 1}.btd")
        for k in ("var_pos", "cluster_idx", "group_nvar", "group_cluster_off", "seq", "v_in_src"):
            assert len(b[k]) == len(d[f"u{u}.{k}"]) and (np.asarray(b[k]) == d[f"u{u}.{k}"]).all(), (u, k)
        for k in KEYS + ("group_contig", "var_contig", "var_input_idx"):
            assert (np.asarray(b[k]) == np.asarray(py_units[u][k])).all(), (u, k)
        assert (b["regions"] == py_regions).all()
    r = subprocess.run([str(_exe()), str(tmp_path / "genome.fa"), str(tmp_path / "c.vcf"), str(tmp_path / "o2.btd"), "--decoy", str(tmp_path / "decoy.fa"),
                        "--min-unit-variants", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "no usable variant" in r.stderr


def test_intercluster_regions_file_equals_the_reference_file(tmp_path):
    """btcluster --regions-prefix against <out>_cluster_data/intercluster_regions.txt.gz of the reference's cluster stage
    (tests/golden/make_cluster_data_fixture.py): same lines in the same order, regions of equal length included."""
    from tests.golden.make_fixtures import PIPE_WORKLOADS
    w = PIPE_WORKLOADS["pipe_mixed_3s"]()
    _write_fasta(tmp_path / "genome.fa", {w.chrom: w.reference})
    _write_vcf(tmp_path / "c.vcf", {w.chrom: w.variants})
    r = subprocess.run([str(_exe()), str(tmp_path / "genome.fa"), str(tmp_path / "c.vcf"), str(tmp_path / "out.btd"), "--regions-prefix", str(tmp_path / "intercluster_regions")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = gzip.open(GOLD / "cluster_data_mixed_3s.intercluster_regions.txt.gz").read()
    got = gzip.open(tmp_path / "intercluster_regions.txt.gz").read()
    lengths = [int(l.split(b"\t")[3]) - int(l.split(b"\t")[2]) for l in want.splitlines()]
    assert len(set(lengths)) < len(lengths)          # the fixture does hold ties
    assert got == want


def test_driver_inputs_with_the_native_builder():
    from bayestyper_b200 import driver
    from tests.golden.make_fixtures import E2E_WORKLOADS
    w = E2E_WORKLOADS["e2e_mixed_3s"]()
    a = driver.Inputs(w.chrom, w.reference, w.variants, list(w.genders), spectra=None).prepare()
    b = driver.Inputs(w.chrom, w.reference, w.variants, list(w.genders), spectra=None, native_builder=True).prepare()
    assert a.regions == b.regions
    for k, v in a.graphs.items():
        if isinstance(v, list):
            assert b.graphs[k] == v, k
        else:
            assert (np.asarray(b.graphs[k]) == np.asarray(v)).all(), k
