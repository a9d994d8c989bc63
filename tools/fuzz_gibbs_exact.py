"""Live check that oracle-P in mt19937 mode reproduces the REFERENCE exactly on workloads other than the committed fixtures: for every seeded
case oracle-R (oracle/_ref/btref) runs cluster + genotype (default mode or --noise-genotyping), the fixture machinery of
tests/golden/make_fixtures.py packs its haplotype descriptors, tables, noise trace and VCF numbers, and oracle-P must return the same diplotype
tallies (GPP / APP equal to the last printed digit), GT / GQ / SAF and — in the lock-step modes — the same noise-rate trace.  The assertions are
those of tests/test_ref_parity_exact.py.  Runs only where /root/reference was compiled.

    python tools/fuzz_gibbs_exact.py [--cases 12] [--seed 1]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import synth  # noqa: E402
from tests import _oracle as O  # noqa: E402
from tests._fixtures import GOLD, GibbsFixture  # noqa: E402
from tests.golden import make_fixtures as mf  # noqa: E402
from tests.test_ref_parity_exact import _print_equal  # noqa: E402


def workload(i, seed):
    kind = i % 4
    if kind == 0:
        return "mixed", synth.small_mixed(90, 9_000, 1 + i % 3, seed=seed, frac_indel=0.3)
    if kind == 1:
        return "chrX", synth.small_mixed(80, 9_000, 2, seed=seed, chrom="chrX")
    if kind == 2:
        return "nested", synth.nested_sv(5, 16_000, 2, seed=seed, n_background=50, sv_len=(150, 500), repeat_frac=0.5)
    return "deep", synth.deep_nested(2, 12_000, 2, seed=seed, n_background=60)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    n_bad = 0
    for i in range(a.cases):
        seed = a.seed * 1000 + i
        name, w = workload(i, seed)
        joint = i % 2 == 1
        tmp = f"_fuzz_exact_{seed}"
        try:
            mf.make(tmp, w, 10**6, seed=20190401 + i, n_errors=1500, extra_args=("--noise-genotyping",) if joint else (), store_tables=False)
            fx = GibbsFixture(tmp)
            cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
            with O.reference_streams(fx.groups):
                if joint:
                    res, trace = O.oracle_estimate_noise_and_genotypes(fx.unit, cd, fx.opts())
                else:
                    trace = O.oracle_estimate_noise(fx.unit, cd, fx.opts())
                    res = O.oracle_estimate_genotypes(fx.unit, cd, fx.opts())
            assert trace.shape == fx.noise_trace.shape and (trace[:, :2] == fx.noise_trace[:, :2]).all(), "trace rows"
            _print_equal(trace[:, 2:], fx.noise_trace[:, 2:])
            assert (res["gpp"] == fx.ref["gpp"]).all(), "gpp"
            assert (res["app"] == fx.ref["app"]).all(), "app"
            for k in ("gt", "gq", "saf"):
                assert (res[k] == fx.ref[k]).all(), k
            _print_equal(res["nak"], fx.ref["nak"])
            _print_equal(res["mac"], fx.ref["mac"])
            print(f"case {seed} ({name}, {'joint' if joint else 'default'}, S={fx.S}, {fx.unit.Cn} clusters): exact")
        except AssertionError as e:
            n_bad += 1
            print(f"case {seed} ({name}, {'joint' if joint else 'default'}): MISMATCH {e}")
        finally:
            (GOLD / f"{tmp}.btd").unlink(missing_ok=True)
    print(f"{a.cases} cases, {n_bad} mismatches")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
