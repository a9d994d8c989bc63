"""Short driver for ncu: Bloom probes against a DRAM-sized filter (n=1e9 -> 1.8 GB)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bayestyper_b200 import capi

lib = capi.load()
capi.check(lib.btg_init(0))
K = 55
n_filter = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
nq = 32_000_000
g = torch.Generator(device="cuda").manual_seed(1)
b = capi.check(lib.btg_bloom_create(n_filter, 1e-3, K))
chunk = 100_000_000
first = None
for i in range(n_filter // chunk):
    kf = torch.empty((chunk, 2), dtype=torch.int64, device="cuda").random_(generator=g)
    kf[:, 1] &= (1 << 46) - 1
    torch.cuda.synchronize()
    capi.check(lib.btg_bloom_insert_dev(b, kf.data_ptr(), chunk, None))
    torch.cuda.synchronize()
    if first is None:
        first = kf[: nq // 2].clone()
q = torch.empty((nq, 2), dtype=torch.int64, device="cuda").random_(generator=g)
q[:, 1] &= (1 << 46) - 1
q[: nq // 2] = first
q = q[torch.randperm(nq, device="cuda")].contiguous()
hit = torch.zeros(nq, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
import os
modes = [int(x) for x in os.environ.get("MODES", "0").split(",")]
s = torch.cuda.ExternalStream(lib.btg_get_stream())
for mode in modes:
    lib.btg_debug_set_probe_mode(mode)
    for _ in range(3):
        capi.check(lib.btg_bloom_lookup_dev(b, q.data_ptr(), nq, hit.data_ptr(), None))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5):
        capi.check(lib.btg_bloom_lookup_dev(b, q.data_ptr(), nq, hit.data_ptr(), None))
    e1.record(s)
    s.synchronize()
    print("mode", mode, "L2_FETCH", os.environ.get("BTG_L2_FETCH"), "ms per lookup launch", e0.elapsed_time(e1) / 5, "hit rate", float(hit.float().mean()), flush=True)
