"""Differential check of bayestyper_b200/graph_builder.py against the REFERENCE's VariantFileParser + VariantClusterGraph +
VariantClusterGroup (oracle-R: oracle/_ref/btref run --cluster-only --dump-graphs) on adversarial candidate sets:
nested and overlapping deletions, deletions bridging clusters (merges), multi-allelic variants with different reference
spans, '*' alleles, copy-number insertions in front of tandem repeats, N runs, reference mismatches, variants at the contig ends.

Runs only where /root/reference was compiled (this container).  `--write-golden` stores the cases and the reference's arrays as
tests/golden/graphs_adversarial.btd for tests/test_graph_builder.py.

    python tools/fuzz_graph_builder.py [--cases 40] [--seed 1] [--write-golden]
"""
from __future__ import annotations

import argparse
import gzip
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import btd, graph_builder, synth  # noqa: E402

BTREF = ROOT / "oracle" / "_ref" / "btref"
K = 55
ACGT = b"ACGT"
KEYS = ("group_cluster_off", "group_nvar", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst", "cluster_idx",
        "cl_vertex_off", "cl_var_off", "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_nested", "v_refvar_off", "v_in_off", "v_in_src",
        "var_pos", "var_dep", "var_nalt", "var_alt_off", "alt_reflen", "alt_seq_off", "alt_seq")


def _snv(rng, r, p):
    i = ACGT.find(r[p])             # on an N the variant is dropped by put()
    return synth.Variant(p, bytes(r[p:p + 1]), [bytes([ACGT[(max(i, 0) + int(rng.integers(1, 4))) % 4]])])


def adversarial_case(seed: int, length: int = 6000):
    """One contig and a position-sorted candidate set mixing the situations listed in the module docstring."""
    rng = np.random.default_rng(seed)
    r = bytearray(synth.random_reference(length, seed + 1000))
    for _ in range(int(rng.integers(0, 3))):                                    # N runs
        p = int(rng.integers(200, length - 200))
        r[p:p + int(rng.integers(1, 30))] = b"N" * int(rng.integers(1, 30))
    r = bytearray(r[:length])
    length = len(r)
    cnv_sites = []
    for _ in range(int(rng.integers(0, 3))):                                    # tandem repeats: unit x copies
        unit = int(rng.integers(60, 140)); copies = int(rng.integers(2, 5))
        p = int(rng.integers(300, length - 300 - unit * copies))
        for c in range(1, copies):
            r[p + c * unit:p + (c + 1) * unit] = r[p:p + unit]
        cnv_sites.append((p, unit, copies))
    r = bytes(r)
    var = {}

    def put(v):
        if v.pos not in var and all(c in ACGT for c in v.ref) and v.pos + len(v.ref) < length:
            var[v.pos] = v

    for p, unit, copies in cnv_sites:                                           # insertion / deletion of one repeat unit in front
        if rng.random() < 0.5:
            put(synth.Variant(p - 1, r[p - 1:p], [r[p - 1:p] + r[p:p + unit]]))
        else:
            put(synth.Variant(p - 1, r[p - 1:p + unit], [r[p - 1:p]]))
        if rng.random() < 0.5:
            q = p + unit * copies + int(rng.integers(5, 90))
            put(_snv(rng, r, q))
    for _ in range(int(rng.integers(2, 7))):                                    # large deletions with inner variants, some nested twice
        ln = int(rng.integers(60, 700))
        p = int(rng.integers(60, length - ln - 120))
        put(synth.Variant(p, r[p:p + 1 + ln], [r[p:p + 1]]))
        for q in (p + 1 + np.flatnonzero(rng.random(ln) < 0.02)).tolist():
            kind = rng.random()
            if kind < 0.6:
                put(_snv(rng, r, q))
            elif kind < 0.8:
                d = int(rng.integers(1, 40))
                put(synth.Variant(q, r[q:q + 1 + d], [r[q:q + 1]]))
            elif kind < 0.9 and ln > 250:
                d = int(rng.integers(70, min(ln - 100, 300)))
                put(synth.Variant(q, r[q:q + 1 + d], [r[q:q + 1]]))         # deletion inside the deletion (may stick out)
            else:
                ins = bytes(ACGT[i] for i in rng.integers(0, 4, int(rng.integers(1, 70))))
                put(synth.Variant(q, r[q:q + 1], [r[q:q + 1] + ins]))
    for _ in range(int(rng.integers(5, 40))):                                   # background: SNVs, indels, multi-allelics
        p = int(rng.integers(0, length - 80))
        kind = rng.random()
        if kind < 0.5:
            put(_snv(rng, r, p))
        elif kind < 0.7:
            d1 = int(rng.integers(2, 30)); d2 = int(rng.integers(1, d1))
            alts = [r[p:p + 1], r[p:p + 1] + r[p + 1 + d2:p + 1 + d1]]      # two deletions of different length: two reference spans
            alts = list(dict.fromkeys(x for x in alts if x != r[p:p + 1 + d1]))
            put(synth.Variant(p, r[p:p + 1 + d1], alts))
        elif kind < 0.85:
            ins = bytes(ACGT[i] for i in rng.integers(0, 4, int(rng.integers(1, 20))))
            other = _snv(rng, r, p).alts[0]
            put(synth.Variant(p, r[p:p + 1], [r[p:p + 1] + ins, other]))
        elif kind < 0.92:
            put(synth.Variant(p, bytes([ACGT[(ACGT.index(r[p]) + 1) % 4]]) if r[p] in ACGT else b"A", [b"G" if r[p:p + 1] != b"G" else b"T"]))   # REF mismatch: excluded
        else:
            q = int(rng.integers(0, 54)) if rng.random() < 0.5 else length - 1 - int(rng.integers(0, 60))                                  # contig ends
            if 0 <= q < length and r[q] in ACGT:
                put(_snv(rng, r, q))
    out = [var[p] for p in sorted(var)]
    # '*' alleles: an SNV under an upstream deletion may carry one
    ends = []
    for v in out:
        open_dep = any(e >= v.pos for e in ends)
        if open_dep and len(v.ref) == 1 and len(v.alts) == 1 and rng.random() < 0.3:
            v.alts.append(b"*")
        if bytes(r[v.pos:v.pos + len(v.ref)]).upper() == v.ref.upper() and v.pos >= K - 1:
            for a in v.alts:
                if a != b"*":
                    rr, _ = graph_builder._right_trim(v.ref, a)
                    if v.pos + len(rr) - 1 + K <= length:
                        ends.append(v.pos + len(rr) - 1)
    if seed % 3 == 0:                                                           # soft-masked stretches and lower-case alleles
        rb = bytearray(r)
        for _ in range(int(rng.integers(1, 6))):
            p = int(rng.integers(0, length - 300)); q = p + int(rng.integers(20, 300))
            rb[p:q] = bytes(rb[p:q]).lower()
        r = bytes(rb)
        for v in out:
            if rng.random() < 0.3:
                v.ref = v.ref.lower()
            if rng.random() < 0.3:
                v.alts = [a.lower() for a in v.alts]
    return "chrF", r, out


def deep_case(seed: int, length: int = 12000):
    """Deletions inside deletions inside deletions (up to four levels) with SNVs at the bottom: groups of 10+ clusters, which takes the
    reference's unordered_map of clusters past its first rehash, and variants that overlap three or more clusters at once."""
    rng = np.random.default_rng(seed)
    r = synth.random_reference(length, seed + 100)
    var = {}

    def put(p, ref, alts):
        if p not in var and 60 < p and p + len(ref) < length - 60:
            var[p] = synth.Variant(p, ref, alts)

    def fill(lo, hi, depth):
        if hi - lo < 250 or depth > 3:
            for p in rng.integers(lo + 60, max(lo + 61, hi - 60), size=int(rng.integers(0, 3))).tolist():
                put(int(p), r[p:p + 1], [ACGT[(ACGT.index(r[p]) + 1) % 4:][:1]])
            return
        cuts = np.sort(rng.integers(lo + 70, hi - 70, size=2 * int(rng.integers(1, 4))))
        for a, b in zip(cuts[0::2].tolist(), cuts[1::2].tolist()):
            if b - a > 120:
                put(a, r[a:a + 1 + b - a], [r[a:a + 1]])
                fill(a, b, depth + 1)
            else:
                put(a, r[a:a + 1], [ACGT[(ACGT.index(r[a]) + 1) % 4:][:1]])

    for _ in range(3):
        a = int(rng.integers(200, length - 4000)); ln = int(rng.integers(1500, 3500))
        put(a, r[a:a + 1 + ln], [r[a:a + 1]])
        fill(a, a + ln, 1)
    return "chrF", r, [var[p] for p in sorted(var)]


def reference_graphs(chrom, reference, variants):
    w = synth.Workload("fuzz", chrom, reference, variants, np.zeros((1, len(variants), 2), np.int64), ["F"])
    km = synth.unique_kmers(synth.canonical_kmers(reference[:400].replace(b"N", b"A")))[0]
    with tempfile.TemporaryDirectory() as td:
        wd = synth.write_workdir(w, td, spectra=[(km, np.full(len(km), 10, np.uint8))])
        r = subprocess.run([str(BTREF), "run", "--workdir", str(wd), "--threads", "2", "--seed", "1", "--dump-graphs", "--cluster-only"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            return None, None, (r.stderr or r.stdout)[-400:]
        g = btd.read(Path(wd) / "ref_out" / "graphs.btd")
        reg = Path(wd) / "ref_out" / "bayestyper_cluster_data" / "intercluster_regions.txt.gz"
        raw = reg.read_bytes()
        txt = gzip.decompress(raw).decode() if raw[:2] == b"\x1f\x8b" else raw.decode()
        regions = sorted((int(t[2]), int(t[3])) for t in (ln.split("\t") for ln in txt.splitlines()))
    return g, regions, None


def compare(g, regions, b):
    for k in KEYS:
        if len(b[k]) != len(g[k]) or not (np.asarray(b[k]) == np.asarray(g[k])).all():
            return k
    for v in range(len(g["v_flags"])):
        a0, a1 = int(g["v_refvar_off"][v]), int(g["v_refvar_off"][v + 1])
        if set(b["v_refvar"][a0:a1].tolist()) != set(g["v_refvar"][a0:a1].tolist()):
            return "v_refvar"
    if sorted((int(x), int(y)) for x, y in b["regions"]) != regions:
        return "regions"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--write-golden", action="store_true")
    ap.add_argument("--deep", action="store_true", help="deep-nesting cases (deep_case) instead of the mixed adversarial ones; a mismatch in "
                    "cluster_idx / group_src there is the allocator-dependent overlap-set order of the reference (DESIGN.md §7)")
    ap.add_argument("--golden-name", default="graphs_adversarial")
    ap.add_argument("--native", action="store_true", help="also run host/btcluster (include/btgpu_cluster.hpp) on every case")
    a = ap.parse_args()
    golden = {}
    n_bad = n_ref_abort = 0
    stats = dict(groups=0, nested_groups=0, clusters=0, max_group=0, deps=0)
    for c in range(a.cases):
        seed = a.seed * 1000 + c
        chrom, ref, var = (deep_case if a.deep else adversarial_case)(seed)
        g, regions, err = reference_graphs(chrom, ref, var)
        try:
            b = graph_builder.build_unit_graphs(chrom, ref, var)
            mine_err = None
        except (ValueError, AssertionError) as e:
            b, mine_err = None, str(e)
        if g is None:
            n_ref_abort += 1
            print(f"case {seed}: reference aborted ({err.strip().splitlines()[-1] if err.strip() else '?'}); builder: {mine_err or 'built'}")
            continue
        if b is None:
            n_bad += 1
            print(f"case {seed}: builder raised '{mine_err}' where the reference built {len(g['cluster_idx'])} clusters")
            continue
        bad = compare(g, regions, b)
        if not bad and a.native:
            nb = graph_builder.build_genome_graphs_native({chrom: ref}, {chrom: var})
            nb["regions"] = nb["regions"][:, 2:]
            bad = compare(g, regions, nb)
            bad = bad and "native:" + bad
        gsz = np.diff(g["group_cluster_off"].astype(np.int64))
        stats["groups"] += len(gsz); stats["nested_groups"] += int((gsz > 1).sum()); stats["clusters"] += int(gsz.sum())
        stats["max_group"] = max(stats["max_group"], int(gsz.max())); stats["deps"] += len(g["group_edge_src"])
        if bad:
            n_bad += 1
            print(f"case {seed}: MISMATCH in {bad}")
        elif a.write_golden:
            i = len([k for k in golden if k.endswith(".reference")])
            golden[f"c{i}.reference"] = np.frombuffer(ref, np.uint8)
            golden[f"c{i}.var_pos"] = np.array([v.pos for v in var], np.int64)
            alleles = [b",".join([v.ref] + v.alts) for v in var]
            golden[f"c{i}.alleles"] = np.frombuffer(b"\n".join(alleles), np.uint8)
            golden[f"c{i}.regions"] = np.array(regions, np.int64).reshape(-1, 2)
            for k in KEYS + ("v_refvar",):
                golden[f"c{i}.g.{k}"] = np.asarray(g[k]) if not isinstance(g[k], (bytes, str)) else np.frombuffer(g[k], np.uint8)
    print(f"{a.cases} cases: {n_bad} mismatches, {n_ref_abort} reference aborts; reference built {stats}")
    if a.write_golden:
        golden["meta.n_cases"] = np.array([len([k for k in golden if k.endswith('.reference')])], np.uint32)
        btd.write(ROOT / "tests" / "golden" / f"{a.golden_name}.btd", golden)
        print(f"wrote tests/golden/{a.golden_name}.btd", golden["meta.n_cases"][0], "cases")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
