"""The reference's own code (oracle-R) on the FULL BASELINE.json configs[1] — the same seeded candidate set, reference sequence and genotypes as
bench.py's build_batch (synth.config_b), sample spectrum synthesised on the CPU (synth.sample_spectra: same NB(15, 25) model, 500,000 error
k-mers) — every stage timed by btref.  One step takes the reference minutes, so this is not part of `bench.py --impl reference` (which projects
the full-config rate from a bounded sample); it is the measurement that the projection is checked against.
Runs only where /root/reference was compiled.      python tools/reference_full_config.py [--threads N] [--out profiles/<name>.json]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r2_reference_full_configB.json"))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--config", default="B", choices=["B", "D"], help="B: configs[1]; D: the configs[3] shape of bench.py --config D (30 samples, --noise-genotyping)")
    a = ap.parse_args()
    if a.config == "D" and a.out.endswith("r2_reference_full_configB.json"):
        a.out = a.out.replace("configB", "configD")
    t0 = time.time()
    if a.config == "B":
        w = synth.config_b(n_variants=int(300_000 * a.scale), length=int(40_800_000 * a.scale) + int(10_000_000 * a.scale), n_prefix=int(10_000_000 * a.scale))
        n_errors = int(500_000 * a.scale)
    else:           # bench.py build_batch_d: same generators and seeds
        import numpy as np
        S, n_var = 30, max(60, int(20_000 * a.scale))
        ref = synth.random_reference(n_var * 136, 11)
        var = synth.make_variants(ref, n_var, 12, 0.075, 0.075)
        af = np.random.default_rng(31).beta(0.2, 0.8, size=len(var))
        w = synth.Workload("D", "chr1", ref, var, synth.make_genotypes(len(var), S, 32, allele_freq=af), ["F" if i % 2 == 0 else "M" for i in range(S)])
        n_errors = int(50_000 * a.scale)
    with tempfile.TemporaryDirectory() as td:
        synth.write_workdir(w, td, n_errors=n_errors)
        setup_s = time.time() - t0
        print(f"workdir ready after {setup_s:.0f} s: {len(w.variants)} variants", flush=True)
        t1 = time.time()
        subprocess.check_call([str(ROOT / "oracle" / "_ref" / "btref"), "run", "--workdir", td, "--threads", str(a.threads), "--seed", "20190401"] + (["--noise-genotyping"] if a.config == "D" else []), stdout=subprocess.DEVNULL)
        wall = time.time() - t1
        tj = json.loads((Path(td) / "ref_out" / "timings.json").read_text())
    kmer = sum(tj.get(k, 0.0) for k in ("findVariantClusterPaths", "countPathMultigroupKmers", "countPathKmers", "countInterclusterKmers", "parseSampleKmers", "classifyPathKmers"))
    noise, geno = tj.get("estimateNoise", 0.0), tj.get("estimateGenotypes", 0.0) + tj.get("estimateNoiseAndGenotypes", 0.0)
    step = kmer + noise + geno
    out = {"what": "the reference's own translation units (oracle-R) on the FULL " + ("configs[1] workload" if a.config == "B" else "configs[3]-shaped workload of bench.py --config D (30 samples, --noise-genotyping)") + ", one step", "threads": a.threads, "host": f"{os.cpu_count()} cores (build container)",
           "variants": len(w.variants), "clusters": tj["num_clusters"], "clusters_genotyped": tj["clusters_genotyped"],
           "kmer_stages_s": kmer, "estimateNoise_s": noise, "estimateGenotypes_s": geno, "step_s": step, "clusters_per_s": tj["clusters_genotyped"] / step,
           "wall_s_incl_parsing_and_cluster_construction": wall, "scale": a.scale, "timings": tj}
    Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps({k: v for k, v in out.items() if k != "timings"}))


if __name__ == "__main__":
    main()
