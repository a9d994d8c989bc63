"""Differential check of the VCF writer (include/btgpu_vcf.hpp through host/btvcf) against the REFERENCE's own GenotypeWriter on fresh
seeded workloads: for every case oracle-R (oracle/_ref/btref) runs cluster + genotype and writes its VCF; the writer is then given the
reference's numbers as flat arrays plus the variant description (tests/test_vcf_writer.py::_arrays_from_vcf, both the full and the
right-trimmed allele description) and must reproduce the file byte for byte (QUAL's last printed digit tolerated, as in the test).

Runs only where /root/reference was compiled (this container).   python tools/fuzz_vcf_writer.py [--cases 24] [--seed 1]
"""
from __future__ import annotations

import argparse
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import btd, synth  # noqa: E402
from tests.test_vcf_writer import _arrays_from_vcf, _build_btvcf, _qual_tolerant_equal  # noqa: E402

BTREF = ROOT / "oracle" / "_ref" / "btref"


def workload(i: int, seed: int):
    kind = i % 6
    if kind == 0:
        return "mixed_1s", synth.small_mixed(120, 12_000, 1, seed=seed)
    if kind == 1:
        return "mixed_3s", synth.small_mixed(100, 10_000, 3, seed=seed, frac_indel=0.3)
    if kind == 2:
        return "chrX_2s", synth.small_mixed(80, 9_000, 2, seed=seed, chrom="chrX")
    if kind == 3:
        w = synth.nested_sv(5, 14_000, 2, seed=seed, n_background=40, sv_len=(150, 500), repeat_frac=0.5)
        for j, v in enumerate(w.variants):
            if j % 4 == 0:
                v.id = f"rs{seed}_{j}"
                v.aco = [("gatk:platypus", "manta", "gatk")[(j + a) % 3] for a in range(len(v.alts))]
        return "nested_2s", w
    if kind == 4:
        return "deep_2s", synth.deep_nested(2, 12_000, 2, seed=seed, n_background=60)
    return "indel_4s", synth.small_mixed(90, 9_000, 4, seed=seed, frac_indel=0.5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=24)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    exe = _build_btvcf()
    n_bad = n_lines = n_qual = 0
    for i in range(a.cases):
        seed = a.seed * 1000 + i
        name, w = workload(i, seed)
        with tempfile.TemporaryDirectory() as td:
            wd = synth.write_workdir(w, td, n_errors=1500)
            r = subprocess.run([str(BTREF), "run", "--workdir", str(wd), "--threads", "4", "--seed", str(20190401 + i)], capture_output=True, text=True)
            if r.returncode != 0:
                print(f"case {seed} ({name}): reference aborted: {(r.stderr.strip().splitlines() or ['?'])[-1]}")
                continue
            want = (Path(wd) / "ref_out" / "bayestyper.vcf").read_text()
            for trim in (False, True):
                btd.write(Path(td) / "in.btd", _arrays_from_vcf(want, w.reference, trim))
                r = subprocess.run([str(exe), str(Path(td) / "in.btd"), str(Path(td) / "out.vcf")], capture_output=True, text=True)
                if r.returncode != 0:
                    n_bad += 1
                    print(f"case {seed} ({name}, trim={trim}): writer failed: {r.stderr.strip()[:200]}")
                    continue
                gl, wl = (Path(td) / "out.vcf").read_text().splitlines(), want.splitlines()
                bad = len(gl) != len(wl)
                for x, y in zip(gl, wl):
                    n_lines += 1
                    if x != y:
                        if not y.startswith("#") and _qual_tolerant_equal(x, y):
                            n_qual += 1
                        else:
                            bad = True
                            print(f"case {seed} ({name}, trim={trim}):\n  got : {x[:240]}\n  want: {y[:240]}")
                            break
                n_bad += bad
    print(f"{a.cases} cases x 2 descriptions: {n_bad} mismatches, {n_lines} lines compared, {n_qual} QUAL last-digit differences")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
