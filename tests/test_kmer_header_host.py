"""Compiles bayestyper_b200/csrc/kmer.cuh with plain g++ and checks its integer
arithmetic (the code the kernels run) against the oracle on the CPU, so that
GPU minutes are not spent finding arithmetic slips."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import _oracle as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def hk(tmp_path_factory):
    so = tmp_path_factory.mktemp("hk") / "hk.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so),
                           str(ROOT / "tests" / "host_kmer_check.cpp")])
    L = C.CDLL(str(so))
    L.hk_hash.restype = C.c_uint64
    L.hk_hash.argtypes = [C.c_uint64, C.c_uint64]
    L.hk_canonical.argtypes = [C.c_uint64, C.c_uint64, O.u64p]
    L.hk_roundtrip.argtypes = [C.c_uint64, C.c_uint64, O.u64p]
    L.hk_mod.restype = C.c_uint64
    L.hk_mod.argtypes = [C.c_uint64, C.c_uint64]
    L.hk_contains.argtypes = [O.u8p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint)]
    L.hk_locs.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, O.u64p]
    L.hk_root.argtypes = [C.c_uint64, C.c_uint64]
    L.hk_scan.restype = C.c_size_t
    L.hk_scan.argtypes = [C.c_char_p, C.c_size_t, O.u64p, O.u64p, O.u32p, C.c_size_t]
    return L


def test_hash_canonical_roundtrip(hk, oracle):
    out = np.zeros(2, np.uint64)
    ref = np.zeros(2, np.uint64)
    for km in O.random_kmers(3000, 11):
        w0, w1 = int(km[0]), int(km[1])
        assert hk.hk_hash(w0, w1) == oracle.bto_ntp64(km, 55)
        hk.hk_roundtrip(w0, w1, out)
        assert (out == km).all()
        hk.hk_canonical(w0, w1, out)
        oracle.bto_canonical(km, 55, ref)
        assert (out == ref).all()
        assert hk.hk_root(w0, w1) == oracle.bto_threaded_bloom_root(km, 55)


def test_palindrome_tie_is_forward(hk, oracle):
    s = O.random_seq(27, 5).decode()
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    # odd k cannot be its own reverse complement, but equal prefixes exercise the compare depth
    seq = s + "A" + "".join(comp[c] for c in reversed(s))
    km = O.pack(seq)
    out = np.zeros(2, np.uint64)
    ref = np.zeros(2, np.uint64)
    hk.hk_canonical(int(km[0]), int(km[1]), out)
    oracle.bto_canonical(km, 55, ref)
    assert (out == ref).all()


def test_fast_modulo(hk):
    rng = np.random.default_rng(7)
    ms = [1, 2, 3, 15, 20, 14378, 14377589, 43132762072, 57510352271, 2**36 - 5, 2**63 + 12345, 2**64 - 1]
    hs = [0, 1, 2**64 - 1, 2**63] + [int(x) for x in rng.integers(0, 2**64, 500, dtype=np.uint64)]
    for m in ms:
        for h in hs:
            assert hk.hk_mod(h, m) == h % m


def test_probe_locations_and_contains(hk, oracle):
    for n, fpr in [(1, 1e-3), (1000, 1e-3), (5000, 1e-4), (77, 0.3)]:
        m = oracle.bto_bloom_num_bits(n, fpr)
        nh = oracle.bto_bloom_num_hashes(m, n)
        kmers = O.random_kmers(n, n)
        bits = O.bloom_build(kmers, m, nh)
        probe = np.concatenate([kmers[:200], O.random_kmers(400, n + 1)])
        hit, probes = O.bloom_lookup(bits, m, nh, probe, want_probes=True)
        locs = np.zeros(nh, np.uint64)
        ref = np.zeros(nh, np.uint64)
        for km, h, p in zip(probe, hit, probes):
            np_ = C.c_uint(0)
            assert hk.hk_contains(bits, m, nh, int(km[0]), int(km[1]), C.byref(np_)) == h
            assert np_.value == p
            hk.hk_locs(m, nh, int(km[0]), int(km[1]), locs)
            oracle.bto_bloom_locs(km, 55, m, nh, ref)
            assert (locs == ref).all()


@pytest.mark.parametrize("n_frac", [0.0, 0.01, 0.2])
def test_rolling_scan(hk, oracle, n_frac):
    seq = O.random_seq(20000, 21, n_frac)
    ref_k, ref_p = O.scan(seq)
    cap = len(seq)
    out = np.zeros((cap, 2), np.uint64)
    hs = np.zeros(cap, np.uint64)
    pos = np.zeros(cap, np.uint32)
    n = hk.hk_scan(seq, len(seq), out.reshape(-1), hs, pos, cap)
    assert n == len(ref_k)
    assert (out[:n] == ref_k).all() and (pos[:n] == ref_p).all()
    for km, h in zip(ref_k[:: max(1, n // 500)], hs[:n][:: max(1, n // 500)]):
        assert oracle.bto_ntp64(km, 55) == h   # rolled hash == from-scratch NTP64 of the canonical k-mer
