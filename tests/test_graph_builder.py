"""The host-side cluster/graph builder (bayestyper_b200/graph_builder.py) against graphs the REFERENCE built
(VariantFileParser + VariantClusterGraph constructor, dumped by oracle-R into the pipe_* fixtures)."""
import numpy as np
import pytest

from bayestyper_b200 import btd, graph_builder
from tests._fixtures import GOLD
from tests.golden.make_fixtures import PIPE_WORKLOADS


@pytest.mark.parametrize("name", list(PIPE_WORKLOADS))
def test_graphs_identical_to_reference(name):
    d = btd.read(GOLD / f"{name}.btd")
    g = {k[2:]: v for k, v in d.items() if k.startswith("g.")}
    w = PIPE_WORKLOADS[name]()
    b = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
    for k in ("group_cluster_off", "cl_vertex_off", "cl_var_off", "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_nested",
              "v_in_off", "v_in_src", "var_pos", "var_dep", "var_nalt", "alt_reflen", "alt_seq", "v_refvar_off", "cluster_idx",
              "group_nvar", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst"):
        assert len(b[k]) == len(g[k]), k
        assert (b[k] == g[k]).all(), k
    # reference_variant_indices come out of an unordered_set: compare as sets per vertex
    for v in range(len(g["v_flags"])):
        a0, a1 = int(g["v_refvar_off"][v]), int(g["v_refvar_off"][v + 1])
        assert set(b["v_refvar"][a0:a1].tolist()) == set(g["v_refvar"][a0:a1].tolist())
    regs = graph_builder.intercluster_regions(w.chrom, w.reference, w.variants)
    ref_regs = sorted((int(a), int(bb)) for dec, a, bb in d["regions"])
    assert sorted(regs) == ref_regs


def test_unordered_container_order_matches_the_toolchain(tmp_path):
    """stdhash_order.UnorderedUInt against std::unordered_set<unsigned> of this toolchain's libstdc++ (the containers whose
    iteration order the reference's cluster order, merge survivor and dependency-edge order come from)."""
    import random
    import shutil
    import subprocess
    from bayestyper_b200.stdhash_order import UnorderedUInt
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = tmp_path / "probe.cpp"
    src.write_text("""
#include <unordered_set>
#include <cstdio>
int main() { std::unordered_set<unsigned> s; char op; unsigned k;
  while (std::scanf(" %c", &op) == 1) {
    if (op == 'i') { if (std::scanf("%u", &k) == 1) s.insert(k); }
    else if (op == 'e') { if (std::scanf("%u", &k) == 1) s.erase(k); }
    else if (op == 'n') { s = std::unordered_set<unsigned>(); }
    else if (op == 'p') { for (auto x : s) std::printf("%u ", x); std::printf("\\n"); } } }
""")
    subprocess.check_call(["g++", "-O1", "-o", str(tmp_path / "probe"), str(src)])
    rnd = random.Random(5)
    ops, want = [], []
    for _ in range(200):
        ops.append("n")
        u, keys = UnorderedUInt(), []
        for _ in range(rnd.randint(1, 150)):
            if keys and rnd.random() < 0.25:
                k = rnd.choice(keys); keys.remove(k); u.erase(k); ops.append(f"e {k}")
            else:
                k = rnd.randint(0, rnd.choice([20, 200, 5000]))
                if u.insert(k):
                    keys.append(k)
                ops.append(f"i {k}")
            if rnd.random() < 0.2:
                ops.append("p"); want.append(" ".join(map(str, u)))
        ops.append("p"); want.append(" ".join(map(str, u)))
    got = subprocess.run([str(tmp_path / "probe")], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
    assert [g.strip() for g in got] == want


@pytest.mark.parametrize("golden, min_cases, min_multi, min_largest", [("graphs_adversarial", 20, 20, 2), ("graphs_deep", 10, 10, 14)])
def test_adversarial_candidate_sets_identical_to_reference(golden, min_cases, min_multi, min_largest):
    """Nested / overlapping / bridging deletions, multi-allelic variants with several reference spans, '*' alleles, copy-number
    insertions in front of tandem repeats, N runs, excluded variants: graphs, groups, dependency edges and intercluster regions
    the REFERENCE built (tools/fuzz_graph_builder.py --write-golden, oracle-R) against graph_builder on the stored cases.
    graphs_deep: deletions nested four levels deep, groups of up to 20 clusters (the reference's unordered_map of clusters rehashes past
    13 and 29 entries: stdhash_order has to follow)."""
    from bayestyper_b200 import synth
    d = btd.read(GOLD / f"{golden}.btd")
    n = int(d["meta.n_cases"][0])
    assert n >= min_cases
    n_multi = largest = 0
    for c in range(n):
        ref = bytes(d[f"c{c}.reference"])
        alleles = bytes(d[f"c{c}.alleles"]).split(b"\n")
        var = []
        for p, al in zip(d[f"c{c}.var_pos"].tolist(), alleles):
            t = al.split(b",")
            var.append(synth.Variant(int(p), t[0], t[1:]))
        b = graph_builder.build_unit_graphs("chrF", ref, var)
        g = {k[len(f"c{c}.g."):]: v for k, v in d.items() if k.startswith(f"c{c}.g.")}
        for k in g:
            if k == "v_refvar":
                continue
            assert len(b[k]) == len(g[k]) and (np.asarray(b[k]) == g[k]).all(), (c, k)
        for v in range(len(g["v_flags"])):
            a0, a1 = int(g["v_refvar_off"][v]), int(g["v_refvar_off"][v + 1])
            assert set(b["v_refvar"][a0:a1].tolist()) == set(g["v_refvar"][a0:a1].tolist())
        assert sorted((int(x), int(y)) for x, y in b["regions"]) == sorted((int(x), int(y)) for x, y in d[f"c{c}.regions"])
        n_multi += int((np.diff(g["group_cluster_off"].astype(np.int64)) > 1).sum())
        largest = max(largest, int(np.diff(g["group_cluster_off"].astype(np.int64)).max()))
    assert n_multi >= min_multi and largest >= min_largest          # the stored cases do exercise groups of several clusters


def test_candidate_vcf_and_fasta_readers(tmp_path):
    """The files `bayesTyper cluster` takes (VariantFileParser.cpp:67-167, Chromosomes.cpp:72-117) read back into the same candidate set:
    ids, ACO origins, '.vcf.gz', and the graphs built from the files equal the graphs built from the in-memory workload."""
    import gzip
    from bayestyper_b200 import synth, vcfio
    from tests.golden.make_vcf_fixtures import VCF_WORKLOADS
    w = VCF_WORKLOADS["vcf_nested_2s"]()
    km = synth.unique_kmers(synth.canonical_kmers(w.reference[:300]))[0]
    wd = synth.write_workdir(w, tmp_path, spectra=[(km, np.ones(len(km), np.uint8))] * 2)
    genome = vcfio.read_fasta(wd / "genome.fa")
    assert list(genome) == [w.chrom] and genome[w.chrom] == w.reference
    (tmp_path / "variants.vcf.gz").write_bytes(gzip.compress((wd / "variants.vcf").read_bytes()))
    for name in ("variants.vcf", "variants.vcf.gz"):
        cand = vcfio.read_candidates(tmp_path / name)
        assert list(cand) == [w.chrom] and len(cand[w.chrom]) == len(w.variants)
        for i, (c, v) in enumerate(zip(cand[w.chrom], w.variants)):
            assert (c.pos, c.ref, c.alts, c.id, c.aco) == (v.pos, v.ref, v.alts, v.id or f"v{i}", v.aco)
    a = graph_builder.build_unit_graphs(w.chrom, genome[w.chrom], cand[w.chrom])
    b = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
    for k in b:
        assert (np.asarray(a[k]) == np.asarray(b[k])).all() if not isinstance(b[k], list) else a[k] == b[k], k
    assert any(a["alt_aco"]) and (np.diff(a["group_cluster_off"].astype(np.int64)) > 1).any()
    with pytest.raises(ValueError, match="vcf"):
        vcfio.read_candidates(tmp_path / "genome.fa")


def test_unit_order_of_the_end_to_end_fixtures():
    """The variants of the unit, in unit order, as the reference's run listed them (e2e fixtures, staged nested case included):
    what the GPU end-to-end test asserts first, checked here without a GPU."""
    from tests.golden.make_fixtures import E2E_NEXT_WORKLOADS, E2E_WORKLOADS
    for name, mk in {**E2E_WORKLOADS, **E2E_NEXT_WORKLOADS}.items():
        d = btd.read(GOLD / f"{name}.btd")
        w = mk()
        g = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
        assert (g["var_pos"] == d["ref.var_pos"]).all(), name


def test_genome_of_several_contigs_identical_to_reference():
    """build_genome_graphs on a genome with several contigs, a contig without variants, decoy contigs (one with variants, which are
    dropped) against the unit the REFERENCE built (tools/fuzz_genome_builder.py --write-golden, oracle-R with --decoy-file): groups
    of all contigs sorted together, contig per group, intercluster regions with the decoy flag."""
    from bayestyper_b200 import synth
    d = btd.read(GOLD / "graphs_genome.btd")
    names = bytes(d["meta.contigs"]).decode().split("\n")
    n_decoys = int(d["meta.n_decoys"][0])
    genome = {n: bytes(d[f"seq.{n}"]) for n in names}
    cand = {}
    for n in bytes(d["meta.cand_contigs"]).decode().split("\n"):
        alleles = bytes(d[f"cand.{n}.alleles"]).split(b"\n")
        cand[n] = [synth.Variant(int(p), al.split(b",")[0], al.split(b",")[1:]) for p, al in zip(d[f"cand.{n}.pos"].tolist(), alleles)]
    b = graph_builder.build_genome_graphs(genome, cand, decoys=names[-n_decoys:])
    for k, v in d.items():
        if k.startswith("g.") and k[2:] not in ("v_refvar", "chroms", "chrom_off"):
            assert len(b[k[2:]]) == len(v) and (np.asarray(b[k[2:]]) == v).all(), k
    chroms = bytes(d["g.chroms"])
    ref_names = [chroms[int(a):int(c)].decode() for a, c in zip(d["g.chrom_off"][:-1], d["g.chrom_off"][1:])]
    assert ref_names == [b["contig_names"][i] for i in b["group_contig"]] and len(set(ref_names)) >= 2
    want = sorted(tuple(ln.split("\t")) for ln in bytes(d["regions"]).decode().split("\n"))
    got = sorted((b["contig_names"][c], str(int(f)), str(int(x)), str(int(y))) for c, f, x, y in b["regions"])
    assert got == want and any(r[1] == "1" for r in want)
    assert len(b["var_contig"]) == len(b["var_pos"])
    with pytest.raises(ValueError, match="does not hold"):
        graph_builder.build_genome_graphs(genome, {"chrGone": [synth.Variant(80, b"A", [b"C"])]})


def test_inference_units_identical_to_reference():
    """The split of a candidate set into inference units (main.cpp:219,233-247; VariantFileParser.cpp:286-290) against the units the
    REFERENCE's parser formed for the same --min-number-of-unit-variants (tools/fuzz_genome_builder.py --units --write-golden; oracle-R's
    `btref units` runs the reference's own unit loop)."""
    from bayestyper_b200 import synth
    d = btd.read(GOLD / "graphs_units.btd")
    names = bytes(d["meta.contigs"]).decode().split("\n")
    n_decoys = int(d["meta.n_decoys"][0])
    genome = {n: bytes(d[f"seq.{n}"]) for n in names}
    cand = {}
    for n in bytes(d["meta.cand_contigs"]).decode().split("\n"):
        alleles = bytes(d[f"cand.{n}.alleles"]).split(b"\n")
        cand[n] = [synth.Variant(int(p), al.split(b",")[0], al.split(b",")[1:]) for p, al in zip(d[f"cand.{n}.pos"].tolist(), alleles)]
    units, regions = graph_builder.build_genome_units(genome, cand, decoys=names[-n_decoys:], min_unit_variants=int(d["meta.min_unit_variants"][0]))
    assert len(units) == int(d["meta.n_units"][0]) >= 3
    for u, b in enumerate(units):
        for k in ("var_pos", "cluster_idx", "group_nvar", "group_cluster_off", "seq", "v_in_src"):
            assert len(b[k]) == len(d[f"u{u}.{k}"]) and (np.asarray(b[k]) == d[f"u{u}.{k}"]).all(), (u, k)
        chroms = bytes(d[f"u{u}.chroms"])
        ref_names = [chroms[int(a):int(c)].decode() for a, c in zip(d[f"u{u}.chrom_off"][:-1], d[f"u{u}.chrom_off"][1:])]
        assert ref_names == [b["contig_names"][i] for i in b["group_contig"]]
    one, regions_one = graph_builder.build_genome_units(genome, cand, decoys=names[-n_decoys:], min_unit_variants=10**9)
    whole = graph_builder.build_genome_graphs(genome, cand, decoys=names[-n_decoys:])
    assert len(one) == 1 and (one[0]["var_pos"] == whole["var_pos"]).all() and (regions_one == whole["regions"]).all() and (regions == regions_one).all()
    assert sum(len(b["var_pos"]) for b in units) == len(whole["var_pos"])
    with pytest.raises(ValueError, match="no usable variant"):
        graph_builder.build_genome_units(genome, cand, decoys=names[-n_decoys:], min_unit_variants=1)
