"""Short driver for ncu: the Gibbs kernel on a small synthetic unit."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bayestyper_b200 import capi, engine, synth, synth_unit, unit as U

lib = capi.load()
capi.check(lib.btg_init(0))
n_var = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
ref = synth.random_reference(n_var * 136, 11)
var = synth.make_variants(ref, n_var, 12, 0.075, 0.075)
g = synth.make_genotypes(len(var), 1, 13)
w = synth.Workload("B", "chr22", ref, var, g, ["F"])
unit = synth_unit.build_unit(w, seed=14)
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 1
unit = synth_unit.tile_unit(unit, tile)
print("clusters", unit.Cn, "H hist", np.bincount(unit.a["cl_nhap"])[:10])
cd = engine.CountDistribution([0.6], [22.5]); cd.set_noise_rates([0.02])
opts = U.default_opts(min_frac=[0.5])
eng = engine.InferenceEngine(unit)
s = torch.cuda.ExternalStream(lib.btg_get_stream())
for it in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    capi.check(lib.btg_estimate_genotypes_async(eng.h, cd.h, C.addressof(opts), None))
    e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1)
    print("gibbs ms", ms, "clusters/s", unit.Cn / ms * 1e3, flush=True)
