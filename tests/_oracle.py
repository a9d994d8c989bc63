"""ctypes view of oracle/libbtoracle.so (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
_lib = None

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def load():
    global _lib
    if _lib is not None:
        return _lib
    so = ROOT / "oracle" / "libbtoracle.so"
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "libbtoracle.so"])
    L = C.CDLL(str(so))
    L.bto_pack_kmer.argtypes = [C.c_char_p, C.c_uint, u64p]
    L.bto_unpack_kmer.argtypes = [u64p, C.c_uint, C.c_char_p]
    L.bto_canonical.argtypes = [u64p, C.c_uint, u64p]
    L.bto_ntp64.restype = C.c_uint64
    L.bto_ntp64.argtypes = [u64p, C.c_uint]
    L.bto_ntp64_seeded.restype = C.c_uint64
    L.bto_ntp64_seeded.argtypes = [u64p, C.c_uint, C.c_uint]
    L.bto_nt_rhval.restype = C.c_uint64
    L.bto_nt_rhval.argtypes = [u64p, C.c_uint]
    L.bto_bloom_num_bits.restype = C.c_uint64
    L.bto_bloom_num_bits.argtypes = [C.c_uint64, C.c_float]
    L.bto_bloom_num_hashes.restype = C.c_uint
    L.bto_bloom_num_hashes.argtypes = [C.c_uint64, C.c_uint64]
    L.bto_threaded_bloom_sub_kmers.restype = C.c_uint64
    L.bto_threaded_bloom_sub_kmers.argtypes = [C.c_uint64]
    L.bto_threaded_bloom_root.restype = C.c_uint
    L.bto_threaded_bloom_root.argtypes = [u64p, C.c_uint]
    L.bto_bloom_locs.argtypes = [u64p, C.c_uint, C.c_uint64, C.c_uint, u64p]
    L.bto_bloom_insert.argtypes = [u8p, C.c_uint64, C.c_uint, C.c_uint, u64p, C.c_size_t]
    L.bto_bloom_lookup.argtypes = [u8p, C.c_uint64, C.c_uint, C.c_uint, u64p, C.c_size_t, u8p, C.c_void_p]
    L.bto_scan_sequence.restype = C.c_size_t
    L.bto_scan_sequence.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, u64p, u32p, C.c_size_t]
    _lib = L
    return L


# ---- helpers shared by tests -------------------------------------------------
K = 55


def pack(seq: str) -> np.ndarray:
    out = np.zeros(2, np.uint64)
    assert load().bto_pack_kmer(seq.encode(), K, out) == 0
    return out


def unpack(km: np.ndarray) -> str:
    buf = C.create_string_buffer(K + 1)
    load().bto_unpack_kmer(np.ascontiguousarray(km, np.uint64), K, buf)
    return buf.value.decode()


def random_kmers(n: int, seed: int) -> np.ndarray:
    """n random packed 55-mers, (n,2) uint64, boundary layout."""
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 2**64, size=(n, 2), dtype=np.uint64)
    w[:, 1] &= np.uint64((1 << (2 * K - 64)) - 1)
    return w


def random_seq(n: int, seed: int, n_frac: float = 0.0) -> bytes:
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
    if n_frac > 0:
        s = s.copy()
        s[rng.random(n) < n_frac] = ord("N")
    return s.tobytes()


def scan(seq: bytes):
    L = load()
    cap = max(len(seq), 1)
    out = np.zeros((cap, 2), np.uint64)
    pos = np.zeros(cap, np.uint32)
    n = L.bto_scan_sequence(seq, len(seq), K, out.reshape(-1), pos, cap)
    return out[:n], pos[:n]


def bloom_build(kmers: np.ndarray, m: int, nh: int) -> np.ndarray:
    bits = np.zeros((m + 7) // 8, np.uint8)
    load().bto_bloom_insert(bits, m, nh, K, np.ascontiguousarray(kmers).reshape(-1), len(kmers))
    return bits


def bloom_lookup(bits: np.ndarray, m: int, nh: int, kmers: np.ndarray, want_probes=False):
    hit = np.zeros(len(kmers), np.uint8)
    probes = np.zeros(len(kmers), np.uint8)
    load().bto_bloom_lookup(bits, m, nh, K, np.ascontiguousarray(kmers).reshape(-1), len(kmers), hit,
                            probes.ctypes.data_as(C.c_void_p))
    return (hit, probes) if want_probes else hit


# ---- oracle-P (gibbs_oracle.cpp) ------------------------------------------------------------
def _bind_gibbs(L):
    if getattr(L, "_gibbs_bound", False):
        return
    dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    L.bto_count_dist_create.restype = C.c_void_p
    L.bto_count_dist_create.argtypes = [C.c_uint32, dp, dp, C.c_float, C.c_float]
    L.bto_count_dist_set_noise_rates.argtypes = [C.c_void_p, dp]
    L.bto_count_dist_get_noise_rates.argtypes = [C.c_void_p, dp]
    L.bto_count_dist_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bto_count_dist_free.argtypes = [C.c_void_p]
    L.bto_nb_moments_to_parameters.argtypes = [C.c_double, C.c_double, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.bto_estimate_genotypes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bto_estimate_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bto_estimate_noise_and_genotypes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bto_set_rng_mode.argtypes = [C.c_int]
    L.bto_set_group_indices.argtypes = [C.c_void_p]
    L._gibbs_bound = True


class reference_streams:
    """`with reference_streams(group_indices):` — oracle-P draws from std::mt19937 + libstdc++ distributions with the reference's
    seeds (gibbs_oracle.cpp mode 1); group_indices = index of every group of the unit in the reference's full unit."""

    def __init__(self, group_indices=None):
        self.idx = None if group_indices is None else np.ascontiguousarray(group_indices, np.uint64)

    def __enter__(self):
        L = load(); _bind_gibbs(L)
        L.bto_set_rng_mode(1)
        L.bto_set_group_indices(self.idx.ctypes.data if self.idx is not None else None)
        return self

    def __exit__(self, *a):
        L = load()
        L.bto_set_rng_mode(0)
        L.bto_set_group_indices(None)


class OracleCountDist:
    def __init__(self, nb_p, nb_size, prior=(1.0, 0.01)):
        self.L = load()
        _bind_gibbs(self.L)
        self.S = len(nb_p)
        self.h = self.L.bto_count_dist_create(self.S, np.ascontiguousarray(nb_p, np.float64), np.ascontiguousarray(nb_size, np.float64), prior[0], prior[1])

    def set_noise_rates(self, rates):
        self.L.bto_count_dist_set_noise_rates(self.h, np.ascontiguousarray(rates, np.float64))

    def noise_rates(self):
        out = np.zeros(self.S)
        self.L.bto_count_dist_get_noise_rates(self.h, out)
        return out

    def tables(self):
        g = np.zeros((self.S, 256, 256)); n = np.zeros((self.S, 256))
        self.L.bto_count_dist_tables(self.h, g.ctypes.data, n.ctypes.data)
        return g, n

    def __del__(self):
        try:
            self.L.bto_count_dist_free(self.h)
        except Exception:
            pass


def oracle_estimate_genotypes(unit, cd: OracleCountDist, opts, want_tally=False):
    L = load()
    _bind_gibbs(L)
    res, arrays = unit.alloc_result()
    desc = unit.desc()
    toff = unit.tally_offsets()
    tally = np.zeros(int(toff[-1]), np.uint32) if want_tally else None
    rc = L.bto_estimate_genotypes(C.addressof(desc), cd.h, C.addressof(opts), C.addressof(res),
                                  tally.ctypes.data if want_tally else None, toff.ctypes.data)
    assert rc == 0, rc
    return (arrays, tally) if want_tally else arrays


def oracle_estimate_noise(unit, cd: OracleCountDist, opts):
    L = load()
    _bind_gibbs(L)
    desc = unit.desc()
    rows = opts.n_chains * (opts.gibbs_burn_in + opts.gibbs_samples + 1) + 1
    trace = np.zeros((rows, 2 + unit.S))
    rc = L.bto_estimate_noise(C.addressof(desc), cd.h, C.addressof(opts), trace.ctypes.data)
    assert rc == 0, rc
    return trace


def oracle_estimate_noise_and_genotypes(unit, cd: OracleCountDist, opts):
    L = load()
    _bind_gibbs(L)
    res, arrays = unit.alloc_result()
    desc = unit.desc()
    rows = opts.n_chains * (opts.gibbs_burn_in + opts.gibbs_samples + 1)
    trace = np.zeros((rows, 2 + unit.S))
    rc = L.bto_estimate_noise_and_genotypes(C.addressof(desc), cd.h, C.addressof(opts), C.addressof(res), trace.ctypes.data)
    assert rc == 0, rc
    return arrays, trace
