"""Exploratory microbenchmark of the Bloom probe kernel (not the contract bench)."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bayestyper_b200 import capi

lib = capi.load()
capi.check(lib.btg_init(0))
K = 55
res = []
for n_filter, frac_present in [(100_000_000, 0.0), (100_000_000, 1.0), (100_000_000, 0.5), (1_000_000_000, 0.5)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    nq = 64_000_000
    # random packed k-mers on device
    kf = torch.empty((n_filter if n_filter <= 200_000_000 else 200_000_000, 2), dtype=torch.int64, device="cuda").random_(generator=g)
    kf[:, 1] &= (1 << 46) - 1
    b = capi.check(lib.btg_bloom_create(n_filter, 1e-3, K))
    torch.cuda.synchronize()
    # insert (n_filter k-mers; for 1e9 insert the 200M five times with a perturbation)
    t0 = time.time()
    reps = max(1, n_filter // len(kf))
    for r in range(reps):
        if r:
            kf[:, 0] += 0x1E3779B97F4A7C15
        capi.check(lib.btg_bloom_insert_dev(b, kf.data_ptr(), len(kf), None))
    nb = np.zeros(1, np.uint8)
    m = C.c_uint64(); nk = C.c_uint64(); nh = C.c_uint32()
    lib.btg_bloom_info(b, C.byref(nk), C.byref(m), C.byref(nh))
    lib.btg_shutdown.restype = None
    torch.cuda.synchronize()
    # queries
    npres = int(nq * frac_present)
    q = torch.empty((nq, 2), dtype=torch.int64, device="cuda").random_(generator=g)
    q[:, 1] &= (1 << 46) - 1
    if npres:
        q[:npres] = kf[:npres]
        q = q[torch.randperm(nq, device="cuda")].contiguous()
    hit = torch.zeros(nq, dtype=torch.uint8, device="cuda")
    pr = torch.zeros(nq, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        capi.check(lib.btg_bloom_lookup_probes_dev(b, q.data_ptr(), nq, hit.data_ptr(), pr.data_ptr(), s.cuda_stream))
        s.synchronize()
        probes = int(pr.sum(dtype=torch.int64))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            capi.check(lib.btg_bloom_lookup_dev(b, q.data_ptr(), nq, hit.data_ptr(), s.cuda_stream))
        e0.record(s)
        it = 10
        for _ in range(it):
            capi.check(lib.btg_bloom_lookup_dev(b, q.data_ptr(), nq, hit.data_ptr(), s.cuda_stream))
        e1.record(s)
        s.synchronize()
    ms = e0.elapsed_time(e1) / it
    alg = 16 * nq + 32 * probes + nq
    res.append(dict(n_filter=n_filter, filter_MB=m.value / 8e6, nh=nh.value, frac_present=frac_present, nq=nq,
                    probes_per_kmer=probes / nq, ms=ms, gkmers_s=nq / ms / 1e6, alg_GBs=alg / ms / 1e6,
                    hit_rate=float(hit.float().mean())))
    print(json.dumps(res[-1]), flush=True)
    lib.btg_bloom_free(b)
    del kf, q, hit, pr
    torch.cuda.empty_cache()
json.dump(res, open("gpurun_out/bench_bloom.json", "w"), indent=1)
