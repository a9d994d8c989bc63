"""configs[3]-shaped timing (30 samples, --noise-genotyping joint mode) on a synthetic unit: descriptors synthesised directly
(synth_unit), population allele frequencies ~ Beta(0.2, 0.8), Hardy-Weinberg genotypes.  python tools/prof_joint.py [n_variants] [S]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bayestyper_b200 import capi, engine, synth, synth_unit, unit as U

n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lib = capi.load()
capi.check(lib.btg_init(0), lib)
rng = np.random.default_rng(31)
ref = synth.random_reference(n_var * 136, 11)
var = synth.make_variants(ref, n_var, 12, 0.075, 0.075)
af = rng.beta(0.2, 0.8, size=len(var))
g = synth.make_genotypes(len(var), S, 32, allele_freq=af)
w = synth.Workload("D", "chr1", ref, var, g, ["F" if i % 2 == 0 else "M" for i in range(S)])
t = time.time()
unit = synth_unit.build_unit(w, seed=14)
print("unit: clusters", unit.Cn, "variants", unit.n_variants, "samples", S, "H hist", np.bincount(unit.a["cl_nhap"])[:12], "build %.1f s" % (time.time() - t), flush=True)
nb_p, nb_size = [0.6] * S, [22.5] * S
opts = U.default_opts(min_frac=U.min_fraction_observed(nb_p, nb_size))
eng = engine.InferenceEngine(unit)
for mode in ("default", "joint"):
    cd = engine.CountDistribution(nb_p, nb_size)
    torch.cuda.synchronize(); t = time.perf_counter()
    if mode == "default":
        eng.estimate_noise(cd, opts, want_trace=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        eng.estimate_genotypes(cd, opts)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("default mode: estimateNoise %.2f s (%.0f us/iteration), estimateGenotypes %.2f s -> %.0f clusters/s" % (t1 - t, (t1 - t) / 7000 * 1e6, t2 - t1, unit.Cn / (t2 - t)), flush=True)
    else:
        eng.estimate_noise_and_genotypes(cd, opts, want_trace=False)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("joint mode (--noise-genotyping): %.2f s (%.0f us/iteration) -> %.0f clusters/s" % (t2 - t, (t2 - t) / 7000 * 1e6, unit.Cn / (t2 - t)), flush=True)
    cd.close()
