"""Short driver for ncu: one estimateNoise chain on a synthetic unit."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bayestyper_b200 import capi, engine, synth, synth_unit, unit as U

lib = capi.load()
capi.check(lib.btg_init(0))
n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 350
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ref = synth.random_reference(n_var * 136, 11)
var = synth.make_variants(ref, n_var, 12, 0.075, 0.075)
g = synth.make_genotypes(len(var), 1, 13)
w = synth.Workload("B", "chr22", ref, var, g, ["F"])
unit = synth_unit.build_unit(w, seed=14)
print("clusters", unit.Cn)
cd = engine.CountDistribution([0.6], [22.5])
opts = U.default_opts(min_frac=[0.5], chains=chains, burn=iters // 2, samples=iters - iters // 2)
eng = engine.InferenceEngine(unit)
for rep in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    eng.estimate_noise(cd, opts, want_trace=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("estimate_noise s", dt, "us per iteration", dt / (iters * chains) * 1e6, flush=True)
