"""End-to-end check of a composition on a GPU box: tests/golden/make_fixtures.py E2E_NEXT_WORKLOADS -> driver.run -> the
reference's calls (same statistical bars as tests/test_gpu_e2e.py).  Both checks below are also `-m gpu` tests since round 2
(tests/test_gpu_e2e.py imports main / main_genome from here).

    gpurun --timeout 600 -- 'python tools/e2e_check.py e2e_nested_2s; python tools/e2e_check.py genome'

`genome` runs the staged whole-genome composition (bayestyper_b200/driver_genome.py) on the multi-contig fixture.
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


def main_genome():
    """driver_genome.genotype_genome on the multi-contig fixture against the calls in the reference's VCF (tests/golden/vcf_genome_2s.vcf.gz)."""
    import gzip

    from bayestyper_b200 import driver_genome, vcfio
    from tests.golden.make_vcf_genome_fixture import GENDERS, genome_workload, sample_spectra
    parts, empty, decoys = genome_workload()
    genome = {n: w.reference for n, w in parts.items()}
    genome[empty[0]] = empty[1]
    cand = {n: w.variants for n, w in parts.items()}
    inp = driver_genome.GenomeInputs({**genome, **decoys}, cand, list(GENDERS), sample_spectra(parts, decoys), decoys=tuple(decoys))
    out_vcf = ROOT / "gpurun_out" / "genome_check.vcf"
    out_vcf.parent.mkdir(exist_ok=True)
    graphs, res, info = driver_genome.genotype_genome(inp, driver.Options(random_seed=20190401), vcf_out=out_vcf, sample_names=["S1", "S2"])
    tmp = ROOT / "gpurun_out" / "genome_ref.vcf"
    tmp.write_bytes(gzip.open(ROOT / "tests" / "golden" / "vcf_genome_2s.vcf.gz").read())
    _, ref_rows = vcfio.read_vcf(tmp)
    _, got_rows = vcfio.read_vcf(out_vcf)
    ok = True

    def check(cond, what):
        nonlocal ok
        print(("ok   " if cond else "FAIL ") + what)
        ok = ok and bool(cond)

    check([(r["chrom"], r["pos"], r["id"], r["ref"], r["alt"]) for r in got_rows] == [(r["chrom"], r["pos"], r["id"], r["ref"], r["alt"]) for r in ref_rows],
          f"same records in the same order ({len(ref_rows)})")
    n = same = hard = 0
    dgpp = []
    for a, b in zip(got_rows, ref_rows):
        check_info = all(a["info"].get(k) == b["info"].get(k) for k in ("VCS", "VCR", "VCGS", "VCGR"))
        ok = ok and check_info
        for sa, sb in zip(a["samples"], b["samples"]):
            n += 1
            same += sa["GT"] == sb["GT"]
            hard += sa["GT"] != sb["GT"] and "." not in sa["GT"] and "." not in sb["GT"]
            if "GPP" in sa and "GPP" in sb and len(sa["GPP"]) == len(sb["GPP"]):
                dgpp.append(np.abs(np.array(sa["GPP"]) - np.array(sb["GPP"])).max())
    check(same / n > 0.98, f"GT agreement {same / n:.4f} (haploid chrX male calls included)")
    check(hard <= max(2, int(0.003 * n)), f"hard disagreements {hard}")
    check(np.mean(dgpp) < 3e-3, f"mean max|dGPP| {np.mean(dgpp):.2e}")
    print("nb", info["nb"], "noise", info["noise_rates"])
    return 0 if ok else 1


if __name__ == "__main__":
    from bayestyper_b200 import capi
    _lib = capi.load()
    capi.check(_lib.btg_init(0), _lib)
    name = sys.argv[1] if len(sys.argv) > 1 else "e2e_nested_2s"
    sys.exit(main_genome() if name == "genome" else main(name))
