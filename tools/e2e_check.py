"""End-to-end check of a staged fixture on a GPU box: tests/golden/make_fixtures.py E2E_NEXT_WORKLOADS -> driver.run -> the
reference's calls (same statistical bars as tests/test_gpu_e2e.py).  A case that is green here moves to E2E_WORKLOADS.

    gpurun --timeout 600 -- 'python tools/e2e_check.py e2e_nested_2s'
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import btd, driver, synth  # noqa: E402
from tests.golden.make_fixtures import E2E_NEXT_WORKLOADS  # noqa: E402


def main(name):
    d = btd.read(ROOT / "tests" / "golden" / f"{name}.btd")
    w = E2E_NEXT_WORKLOADS[name]()
    spectra = synth.sample_spectra(w, 4, int(d["meta.n_errors"][0]))
    graphs, unit, res, info = driver.run(w.chrom, w.reference, w.variants, spectra, w.genders, driver.Options(random_seed=int(d["meta.seed"][0])))
    S = len(spectra)
    ok = True

    def check(cond, what):
        nonlocal ok
        print(("ok   " if cond else "FAIL ") + what)
        ok = ok and bool(cond)

    check((graphs["var_pos"] == d["ref.var_pos"]).all(), "same variants in the same unit order")
    gsz = np.diff(graphs["group_cluster_off"].astype(np.int64))
    print(f"groups {len(gsz)}, of which {int((gsz > 1).sum())} hold several clusters (max {int(gsz.max())})")
    nb_p, nb_size = info["nb"]
    mean = nb_size * (1 - nb_p) / nb_p
    ref_mean = d["tab.nb_p_size"][:, 1] * (1 - d["tab.nb_p_size"][:, 0]) / d["tab.nb_p_size"][:, 0]
    check(np.abs(mean / ref_mean - 1).max() < 0.03, f"NB mean within 3 % ({mean} vs {ref_mean})")
    r = info["noise_rates"] / d["tab.noise_rates"]
    check((r > 0.3).all() and (r < 3).all(), f"noise rates same order of magnitude (ratio {r})")
    gt_o, gt_r = res["gt"].reshape(-1, S, 2), d["ref.gt"].reshape(-1, S, 2)
    same = (gt_o == gt_r).all(axis=2)
    check(same.mean() > 0.98, f"GT agreement {same.mean():.4f}")
    called_both = (gt_o[..., 0] != 0xFFFF) & (gt_r[..., 0] != 0xFFFF)
    check((~same & called_both).sum() <= max(2, int(0.003 * same.size)), f"hard disagreements {int((~same & called_both).sum())}")
    dg = np.abs(res["gpp"] - d["ref.gpp"])
    check(dg.mean() < 3e-3, f"mean |dGPP| {dg.mean():.2e} (max {dg.max():.3f})")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "e2e_nested_2s"))
