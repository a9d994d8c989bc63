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
    a = ap.parse_args()
    t0 = time.time()
    w = synth.config_b(n_variants=int(300_000 * a.scale), length=int(40_800_000 * a.scale) + int(10_000_000 * a.scale), n_prefix=int(10_000_000 * a.scale))
    with tempfile.TemporaryDirectory() as td:
        synth.write_workdir(w, td, n_errors=int(500_000 * a.scale))
        setup_s = time.time() - t0
        print(f"workdir ready after {setup_s:.0f} s: {len(w.variants)} variants", flush=True)
        t1 = time.time()
        subprocess.check_call([str(ROOT / "oracle" / "_ref" / "btref"), "run", "--workdir", td, "--threads", str(a.threads), "--seed", "20190401"], stdout=subprocess.DEVNULL)
        wall = time.time() - t1
        tj = json.loads((Path(td) / "ref_out" / "timings.json").read_text())
    kmer = sum(tj.get(k, 0.0) for k in ("findVariantClusterPaths", "countPathMultigroupKmers", "countPathKmers", "countInterclusterKmers", "parseSampleKmers", "classifyPathKmers"))
    noise, geno = tj.get("estimateNoise", 0.0), tj.get("estimateGenotypes", 0.0)
    step = kmer + noise + geno
    out = {"what": "the reference's own translation units (oracle-R) on the FULL configs[1] workload, one step", "threads": a.threads, "host": f"{os.cpu_count()} cores (build container)",
           "variants": len(w.variants), "clusters": tj["num_clusters"], "clusters_genotyped": tj["clusters_genotyped"],
           "kmer_stages_s": kmer, "estimateNoise_s": noise, "estimateGenotypes_s": geno, "step_s": step, "clusters_per_s": tj["clusters_genotyped"] / step,
           "wall_s_incl_parsing_and_cluster_construction": wall, "scale": a.scale, "timings": tj}
    Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps({k: v for k, v in out.items() if k != "timings"}))


if __name__ == "__main__":
    main()
