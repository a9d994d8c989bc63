"""One tiny pass of the hot path on cuda:0, checked against the oracle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def run() -> None:
    lib = capi.load()
    capi.check(lib.btg_init(0), lib)
    from tests import _oracle as O   # the checker
    oracle = O.load()
    K = 55
    # k-mer match: scan a sequence, build a filter from half its k-mers, look all up
    seq = O.random_seq(20_000, 123, 0.002)
    kmers = np.zeros((len(seq), 2), np.uint64)
    valid = np.zeros(len(seq), np.uint8)
    capi.check(lib.btg_scan_sequence(seq, len(seq), capi.ptr(kmers), capi.ptr(valid)), lib)
    ref_k, ref_p = O.scan(seq)
    assert (kmers[valid == 1] == ref_k).all(), "scan mismatch"
    half = np.ascontiguousarray(ref_k[::2])
    b = capi.check(lib.btg_bloom_create(len(half), 1e-3, K), lib)
    capi.check(lib.btg_bloom_insert(b, capi.ptr(half), len(half)), lib)
    hit = np.zeros(len(ref_k), np.uint8)
    rk = np.ascontiguousarray(ref_k)
    capi.check(lib.btg_bloom_lookup(b, capi.ptr(rk), len(rk), capi.ptr(hit)), lib)
    m = oracle.bto_bloom_num_bits(len(half), 1e-3)
    nh = oracle.bto_bloom_num_hashes(m, len(half))
    bits = O.bloom_build(half, m, nh)
    assert (hit == O.bloom_lookup(bits, m, nh, rk)).all(), "bloom mismatch"
    lib.btg_bloom_free(b)
    from . import smoke_ext
    smoke_ext.run(lib)
    print(f"smoke ok: {len(ref_k)} k-mers scanned+probed bit-exact; launches={lib.btg_launch_count()}")
