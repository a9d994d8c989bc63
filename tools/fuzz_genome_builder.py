"""Differential check of graph_builder.build_genome_graphs (several contigs, a contig without variants, decoy contigs with and without variants) against the REFERENCE (oracle-R: btref run --cluster-only --dump-graphs
--decoy-file).  `--write-golden` stores one case as tests/golden/graphs_genome.btd.

    python tools/fuzz_genome_builder.py [--cases 10] [--seed 1] [--write-golden]
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
sys.path.insert(0, str(ROOT / "tools"))
from bayestyper_b200 import btd, graph_builder, synth  # noqa: E402
import fuzz_graph_builder as F  # noqa: E402


def genome_case(seed):
    rng = np.random.default_rng(seed)
    genome, cand = {}, {}
    names = ["chr1", "chr2", "chrX", "chr10"][:int(rng.integers(2, 5))]
    for i, n in enumerate(names):
        _, ref, var = F.adversarial_case(seed * 10 + i, length=int(rng.integers(2500, 5000)))
        genome[n] = ref
        cand[n] = var
    genome["chrEmpty"] = synth.random_reference(int(rng.integers(30, 400)), seed + 77)         # no variants; may be shorter than k
    decoys = {"decoyA": synth.random_reference(700, seed + 78), "decoyB": synth.random_reference(300, seed + 79)}
    da = decoys["decoyA"]
    cand["decoyA"] = [synth.Variant(p, da[p:p + 1], [b"A" if da[p:p + 1] != b"A" else b"C"]) for p in (100, 130, 400)]
    order = list(names) + ["decoyA"]
    rng.shuffle(order)
    cand = {n: cand[n] for n in order}
    return genome, decoys, cand


def reference_run(genome, decoys, cand):
    with tempfile.TemporaryDirectory() as td:
        wd = Path(td)
        with open(wd / "genome.fa", "wb") as f:
            for n, s in genome.items():
                f.write(b">" + n.encode() + b" some description\n")
                for i in range(0, len(s), 60):
                    f.write(s[i:i + 60] + b"\n")
        with open(wd / "decoy.fa", "wb") as f:
            for n, s in decoys.items():
                f.write(b">" + n.encode() + b"\n" + s + b"\n")
        with open(wd / "variants.vcf", "w") as f:
            f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
            for n, var in cand.items():
                for i, v in enumerate(var):
                    f.write(f"{n}\t{v.pos + 1}\t{n}_{i}\t{v.ref.decode()}\t{','.join(a.decode() for a in v.alts)}\t.\t.\t.\n")
        first = next(iter(genome.values()))
        km = synth.unique_kmers(synth.canonical_kmers(first[:400].replace(b"N", b"A")))[0]
        synth.write_kmer_file(str(wd / "S1.kmers.bin"), km, np.full(len(km), 10, np.uint8))
        (wd / "samples.tsv").write_text(f"S1\tF\t{wd / 'S1'}\n")
        r = subprocess.run([str(F.BTREF), "run", "--workdir", str(wd), "--threads", "2", "--seed", "1", "--dump-graphs", "--cluster-only",
                            "--decoy-file", str(wd / "decoy.fa")], capture_output=True, text=True)
        if r.returncode != 0:
            return None, None, (r.stderr or r.stdout)[-400:]
        g = btd.read(wd / "ref_out" / "graphs.btd")
        raw = (wd / "ref_out" / "bayestyper_cluster_data" / "intercluster_regions.txt.gz").read_bytes()
        txt = gzip.decompress(raw).decode() if raw[:2] == b"\x1f\x8b" else raw.decode()
        regions = sorted((t[0], int(t[1]), int(t[2]), int(t[3])) for t in (ln.split("\t") for ln in txt.splitlines()))
    return g, regions, None


def reference_units(genome, decoys, cand, min_unit_variants):
    """btref units: the groups of every inference unit + the regions of the whole run."""
    with tempfile.TemporaryDirectory() as td:
        wd = Path(td)
        with open(wd / "genome.fa", "wb") as f:
            for n, s in genome.items():
                f.write(b">" + n.encode() + b"\n" + s + b"\n")
        with open(wd / "decoy.fa", "wb") as f:
            for n, s in decoys.items():
                f.write(b">" + n.encode() + b"\n" + s + b"\n")
        with open(wd / "variants.vcf", "w") as f:
            f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
            for n, var in cand.items():
                for i, v in enumerate(var):
                    f.write(f"{n}\t{v.pos + 1}\t{n}_{i}\t{v.ref.decode()}\t{','.join(a.decode() for a in v.alts)}\t.\t.\t.\n")
        (wd / "samples.tsv").write_text("S1\tF\tnone\n")
        r = subprocess.run([str(F.BTREF), "units", "--workdir", str(wd), "--threads", "2", "--min-unit-variants", str(min_unit_variants),
                            "--decoy-file", str(wd / "decoy.fa")], capture_output=True, text=True)
        if r.returncode != 0:
            return None, None, (r.stderr or r.stdout)[-400:]
        units = []
        i = 1
        while (wd / "ref_out" / f"graphs_unit_{i}.btd").exists():
            units.append(btd.read(wd / "ref_out" / f"graphs_unit_{i}.btd"))
            i += 1
        raw = (wd / "ref_out" / "intercluster_regions.txt.gz").read_bytes()
        txt = gzip.decompress(raw).decode() if raw[:2] == b"\x1f\x8b" else raw.decode()
        regions = sorted((t[0], int(t[1]), int(t[2]), int(t[3])) for t in (ln.split("\t") for ln in txt.splitlines()))
    return units, regions, None


def compare(g, regions, b):
    for k in F.KEYS:
        if len(b[k]) != len(g[k]) or not (np.asarray(b[k]) == np.asarray(g[k])).all():
            return k
    chroms = bytes(g["chroms"]) if not isinstance(g["chroms"], (bytes, str)) else g["chroms"]
    chroms = chroms.encode() if isinstance(chroms, str) else chroms
    ref_names = [chroms[int(a):int(c)].decode() for a, c in zip(g["chrom_off"][:-1], g["chrom_off"][1:])]
    if ref_names != [b["contig_names"][i] for i in b["group_contig"]]:
        return "group_contig"
    mine = sorted((b["contig_names"][c], int(d), int(x), int(y)) for c, d, x, y in b["regions"])
    if mine != regions:
        return "regions"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--write-golden", action="store_true")
    ap.add_argument("--units", action="store_true", help="compare the split into inference units (btref units) instead")
    a = ap.parse_args()
    n_bad = 0
    if a.units:
        for c in range(a.cases):
            seed = a.seed * 100 + c
            genome, decoys, cand = genome_case(seed)
            total = sum(len(v) for v in cand.values())
            for min_unit in (max(1, total // 7), max(1, total // 3), total + 5):
                ref_units, regions, err = reference_units(genome, decoys, cand, min_unit)
                if ref_units is None:
                    try:
                        graph_builder.build_genome_units({**genome, **decoys}, cand, decoys=list(decoys), min_unit_variants=min_unit)
                        verdict = "builder built it: MISMATCH"
                        n_bad += 1
                    except ValueError as e:
                        verdict = f"builder refuses too ({e})"
                    print(f"case {seed} min_unit {min_unit}: reference aborted ({err.strip().splitlines()[-1][-90:] if err.strip() else '?'}); {verdict}")
                    continue
                mine, mine_regions = graph_builder.build_genome_units({**genome, **decoys}, cand, decoys=list(decoys), min_unit_variants=min_unit)
                bad = None if len(mine) == len(ref_units) else f"number of units {len(mine)} vs {len(ref_units)}"
                for u, (g, b) in enumerate(zip(ref_units, mine)):
                    b = dict(b); b["regions"] = mine_regions
                    bad = bad or compare(g, regions, b)
                print(f"case {seed} min_unit {min_unit}: {'MISMATCH: ' + str(bad) if bad else 'identical'} ({len(ref_units)} units, "
                      f"{[len(g['cluster_idx']) for g in ref_units]} clusters)")
                n_bad += bool(bad)
                if a.write_golden and not bad and c == 0 and len(ref_units) >= 3:
                    pack = {"meta.contigs": np.frombuffer("\n".join(list(genome) + list(decoys)).encode(), np.uint8), "meta.n_decoys": np.array([len(decoys)], np.uint32),
                            "meta.cand_contigs": np.frombuffer("\n".join(cand).encode(), np.uint8), "meta.min_unit_variants": np.array([min_unit], np.uint32),
                            "meta.n_units": np.array([len(ref_units)], np.uint32)}
                    for n, s_ in {**genome, **decoys}.items():
                        pack[f"seq.{n}"] = np.frombuffer(s_, np.uint8)
                    for n, var in cand.items():
                        pack[f"cand.{n}.pos"] = np.array([v.pos for v in var], np.int64)
                        pack[f"cand.{n}.alleles"] = np.frombuffer(b"\n".join(b",".join([v.ref] + v.alts) for v in var), np.uint8)
                    for u, g in enumerate(ref_units):
                        for k_ in ("var_pos", "cluster_idx", "group_nvar", "group_cluster_off", "seq", "v_in_src", "chrom_off"):
                            pack[f"u{u}.{k_}"] = np.asarray(g[k_])
                        ch = g["chroms"]
                        pack[f"u{u}.chroms"] = np.frombuffer(ch.encode() if isinstance(ch, str) else bytes(ch), np.uint8)
                    btd.write(ROOT / "tests" / "golden" / "graphs_units.btd", pack)
                    print("wrote tests/golden/graphs_units.btd")
                    a.write_golden = False
        print(f"{a.cases} cases, {n_bad} bad")
        return 1 if n_bad else 0
    for c in range(a.cases):
        seed = a.seed * 100 + c
        genome, decoys, cand = genome_case(seed)
        g, regions, err = reference_run(genome, decoys, cand)
        if g is None:
            print(f"case {seed}: reference aborted: {err.strip().splitlines()[-1] if err.strip() else '?'}")
            n_bad += 1
            continue
        b = graph_builder.build_genome_graphs({**genome, **decoys}, cand, decoys=list(decoys))
        bad = compare(g, regions, b)
        print(f"case {seed}: {'MISMATCH in ' + bad if bad else 'identical'} ({len(g['cluster_idx'])} clusters, {len(regions)} regions, contigs {list(cand)})")
        n_bad += bool(bad)
        if a.write_golden and not bad and c == 0:
            pack = {"meta.contigs": np.frombuffer("\n".join(list(genome) + list(decoys)).encode(), np.uint8), "meta.n_decoys": np.array([len(decoys)], np.uint32),
                    "meta.cand_contigs": np.frombuffer("\n".join(cand).encode(), np.uint8)}
            for n, s in {**genome, **decoys}.items():
                pack[f"seq.{n}"] = np.frombuffer(s, np.uint8)
            for n, var in cand.items():
                pack[f"cand.{n}.pos"] = np.array([v.pos for v in var], np.int64)
                pack[f"cand.{n}.alleles"] = np.frombuffer(b"\n".join(b",".join([v.ref] + v.alts) for v in var), np.uint8)
            for k in F.KEYS + ("v_refvar", "chrom_off"):
                pack[f"g.{k}"] = np.asarray(g[k])
            ch = g["chroms"]
            pack["g.chroms"] = np.frombuffer(ch.encode() if isinstance(ch, str) else bytes(ch), np.uint8)
            pack["regions"] = np.frombuffer("\n".join("\t".join(map(str, r)) for r in regions).encode(), np.uint8)
            btd.write(ROOT / "tests" / "golden" / "graphs_genome.btd", pack)
            print("wrote tests/golden/graphs_genome.btd")
    print(f"{a.cases} cases, {n_bad} bad")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
