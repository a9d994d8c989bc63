"""Short driver for ncu / timing: the sample k-mer stream (k_table_add_sample) against an exact table of the
chr22-like shape: N_KEYS distinct path k-mers, N_REC KMC-ordered records of which ~40 % hit."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bayestyper_b200 import capi

n_keys = int(float(sys.argv[1])) if len(sys.argv) > 1 else 21_000_000
n_rec = int(float(sys.argv[2])) if len(sys.argv) > 2 else 47_000_000
order = sys.argv[3] if len(sys.argv) > 3 else "kmc"
lib = capi.load()
capi.check(lib.btg_init(0), lib)
dev = torch.device("cuda", 0)
s = torch.cuda.ExternalStream(lib.btg_get_stream())
g = torch.Generator(device=dev).manual_seed(1)


def rand_kmers(n):
    k = torch.empty((n, 2), dtype=torch.int64, device=dev).random_(generator=g)
    k[:, 1] &= (1 << 46) - 1
    return k


def keys_of(km):
    lo = torch.empty(len(km), dtype=torch.int64, device=dev); hi = torch.empty_like(lo)
    torch.cuda.synchronize()
    capi.check(lib.btg_table_keys_from_kmers_dev(km.data_ptr(), len(km), lo.data_ptr(), hi.data_ptr(), None), lib)
    s.synchronize()
    return lo, hi


def kmc_sort(km):
    lo, hi = keys_of(km)
    o = torch.sort(lo, stable=True).indices
    o = o[torch.sort(hi[o], stable=True).indices]
    return km[o].contiguous(), lo[o].contiguous(), hi[o].contiguous()


table_km, kw0, kw1 = kmc_sort(rand_kmers(n_keys))
n_hit = int(0.4 * n_rec)
rec = torch.cat([table_km[torch.randint(0, n_keys, (n_hit,), device=dev, generator=g)].unique(dim=0), rand_kmers(n_rec - n_hit)])
if order == "kmc":
    rec, _, _ = kmc_sort(rec)
else:
    rec = rec[torch.randperm(len(rec), device=dev, generator=g)].contiguous()
cts = torch.ones(len(rec), dtype=torch.uint8, device=dev)
bits = int(min(24, max(8, np.ceil(np.log2(n_keys)))))
lut = torch.zeros((1 << bits) + 1, dtype=torch.int64, device=dev)
torch.cumsum(torch.bincount(kw1 >> (46 - bits), minlength=1 << bits), 0, out=lut[1:])
capi.check(lib.btg_table_set_index_dev(lut.data_ptr(), bits), lib)
counts = torch.zeros(n_keys, dtype=torch.uint8, device=dev); has = torch.zeros(n_keys, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
print("keys", n_keys, "records", len(rec), "order", order, "lut_bits", bits, flush=True)
for rep in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    capi.check(lib.btg_table_add_sample_kmers_dev(kw0.data_ptr(), kw1.data_ptr(), n_keys, rec.data_ptr(), cts.data_ptr(), len(rec), 1, 0, counts.data_ptr(),
                                                  has.data_ptr(), None), lib)
    e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1)
    hits = int(has.sum())
    alg = len(rec) * 17 + n_keys * 16 + hits
    print("stream ms %.3f  algorithmic GB/s %.1f  hits %d" % (ms, alg / ms / 1e6, hits), flush=True)
