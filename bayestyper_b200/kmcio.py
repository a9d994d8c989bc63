"""Writers of KMC databases (<prefix>.kmc_pre / .kmc_suf) in the two layouts the reference's vendored KMC API 2.3.0 reads
(external/kmc_api/kmc_file.cpp:177-292): KMC1 (kmc_version 0) and KMC2 (0x200, what KMC 2/3 write).  Used to give the synthetic
samples real databases (the reference's `genotype` and `makeBloom` read KMC files) and to pin include/btgpu_kmc.hpp against the
reference's own reader (tests/test_kmc.py).  Packed k-mers are in the C ABI's layout (nucleotide i at bits [2i, 2i+1])."""
from __future__ import annotations

import struct

import numpy as np

from . import synth

K = synth.K


def _codes(kmers: np.ndarray, k: int) -> np.ndarray:
    """(n, k) uint8 nucleotide codes of packed k-mers."""
    n = len(kmers)
    out = np.zeros((n, k), np.uint8)
    for i in range(k):
        w = kmers[:, i >> 5]
        out[:, i] = ((w >> np.uint64(2 * (i & 31))) & np.uint64(3)).astype(np.uint8)
    return out


def _records(codes: np.ndarray, counts: np.ndarray, p: int, counter_size: int) -> bytes:
    """suffix bytes (4 nt per byte, MSB first) + little-endian counter per record."""
    n, k = codes.shape
    sb = (k - p) // 4
    suf = codes[:, p:].reshape(n, sb, 4).astype(np.uint8)
    packed = (suf[:, :, 0] << 6) | (suf[:, :, 1] << 4) | (suf[:, :, 2] << 2) | suf[:, :, 3]
    c = np.asarray(counts, np.uint64)
    cb = np.stack([((c >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8) for b in range(counter_size)], 1)
    return np.concatenate([packed.astype(np.uint8), cb], 1).tobytes()


def _prefix_values(codes: np.ndarray, p: int) -> np.ndarray:
    v = np.zeros(len(codes), np.int64)
    for i in range(p):
        v = (v << 2) | codes[:, i]
    return v


def write_kmc1(prefix, kmers: np.ndarray, counts: np.ndarray, k: int = K, lut_prefix_length: int = 7, counter_size: int = 1, min_count: int = 1,
               max_count: int = 255, both_strands: bool = True) -> None:
    """KMC1 layout: records in lexicographic order; LUT[4^p] of first-record indices; 5-word header."""
    assert (k - lut_prefix_length) % 4 == 0
    kmers = np.ascontiguousarray(kmers, np.uint64).reshape(-1, 2)
    o = synth.kmc_order(kmers)
    kmers, counts = kmers[o], np.asarray(counts)[o]
    codes = _codes(kmers, k)
    pv = _prefix_values(codes, lut_prefix_length)
    lut = np.searchsorted(pv, np.arange(4 ** lut_prefix_length), side="left").astype(np.uint64)
    header = struct.pack("<5Q", k | (0 << 32), counter_size | (lut_prefix_length << 32), min_count | ((max_count & 0xFFFFFFFF) << 32), len(kmers),
                         (0 if both_strands else 1))
    with open(str(prefix) + ".kmc_pre", "wb") as f:
        f.write(b"KMCP" + lut.tobytes() + header + struct.pack("<I", len(header)) + b"KMCP")
    with open(str(prefix) + ".kmc_suf", "wb") as f:
        f.write(b"KMCS" + _records(codes, counts, lut_prefix_length, counter_size) + b"KMCS")


def write_kmc2(prefix, kmers: np.ndarray, counts: np.ndarray, k: int = K, lut_prefix_length: int = 7, counter_size: int = 1, min_count: int = 1,
               max_count: int = 255, both_strands: bool = True, signature_len: int = 5, n_bins: int = 8, seed: int = 0) -> None:
    """KMC2 layout: k-mers are dealt to n_bins bins (by a stand-in for KMC's minimiser signatures: a seeded hash of the k-mer; the
    listing API never consults the signature map), each bin sorted, LUT[bin][4^p], one extra LUT word, signature map, 40-byte header."""
    assert (k - lut_prefix_length) % 4 == 0
    kmers = np.ascontiguousarray(kmers, np.uint64).reshape(-1, 2)
    counts = np.asarray(counts)
    rng = np.random.default_rng(seed)
    salt = rng.integers(1, 2**63, dtype=np.uint64) | np.uint64(1)
    bins = (((kmers[:, 0] ^ (kmers[:, 1] << np.uint64(7))) * salt) >> np.uint64(40)).astype(np.int64) % n_bins
    hi, lo = synth.lexicographic_words(kmers)
    o = np.lexsort((lo, hi, bins))                                 # bin by bin, each bin in lexicographic order
    kmers, counts, bins = kmers[o], counts[o], bins[o]
    codes = _codes(kmers, k)
    pv = _prefix_values(codes, lut_prefix_length)
    key = bins * (4 ** lut_prefix_length) + pv
    lut = np.searchsorted(key, np.arange(n_bins * 4 ** lut_prefix_length), side="left").astype(np.uint64)
    sig_map = rng.integers(0, n_bins, size=4 ** signature_len + 1, dtype=np.uint32)
    fields = struct.pack("<7IQB", k, 0, counter_size, lut_prefix_length, signature_len, min_count, max_count, len(kmers), 0 if both_strands else 1)
    header = fields + b"\0" * (40 - len(fields)) + struct.pack("<I", 0x200)       # 37 bytes of fields, padding, and the version word that closes the header
    with open(str(prefix) + ".kmc_pre", "wb") as f:
        f.write(b"KMCP" + lut.tobytes() + struct.pack("<Q", len(kmers)) + sig_map.tobytes() + header + struct.pack("<I", len(header)) + b"KMCP")
    with open(str(prefix) + ".kmc_suf", "wb") as f:
        f.write(b"KMCS" + _records(codes, counts, lut_prefix_length, counter_size) + b"KMCS")
    return kmers, counts                                           # in listing order


def read_kmc(prefix):
    """(kmers (n, 2) uint64 in the C ABI's layout, counts (n,) uint32, info dict) of a KMC1 / KMC2 database in listing order, with the
    database's [min_count, max_count] window applied — the numpy twin of include/btgpu_kmc.hpp (CKMCFile::OpenForListing / ReadNextKmer,
    external/kmc_api/kmc_file.cpp:66-99,177-292,428-515), for the Python mirror's sample loading."""
    pre = open(str(prefix) + ".kmc_pre", "rb").read()
    if len(pre) < 8 or pre[:4] != b"KMCP" or pre[-4:] != b"KMCP":
        raise ValueError(f"{prefix}.kmc_pre: not a KMC file (marker)")
    end = len(pre) - 4
    version = struct.unpack_from("<I", pre, end - 8)[0]
    header_offset = pre[end - 4]
    h = end - 4 - header_offset
    if version == 0x200:
        k, mode, counter_size, p, sig_len, min_count, max_count, total, flag = struct.unpack_from("<7IQB", pre, h)
        lut_bytes = h - ((4 ** sig_len + 1) * 4) - 4 - 8
    elif version == 0:
        d0, d1, d2, d3, d4 = struct.unpack_from("<5Q", pre, h)
        k, mode, counter_size, p = d0 & 0xFFFFFFFF, d0 >> 32, d1 & 0xFFFFFFFF, d1 >> 32
        min_count, max_count, total, flag = d2 & 0xFFFFFFFF, (d2 >> 32) + (d4 & 0xFFFFFFFF00000000), d3, int((d4 & 0xF) == 1)
        lut_bytes = h - 4
    else:
        raise ValueError("unsupported KMC database version")
    if mode != 0:
        raise ValueError("KMC databases with quality-aware counters (mode 1) are not supported")
    lut = np.frombuffer(pre, np.uint64, lut_bytes // 8, 4).astype(np.int64)
    suf = np.fromfile(str(prefix) + ".kmc_suf", np.uint8)
    if len(suf) < 8 or bytes(suf[:4]) != b"KMCS" or bytes(suf[-4:]) != b"KMCS":
        raise ValueError(f"{prefix}.kmc_suf: not a KMC file (marker)")
    sb = (k - p) // 4
    rec = suf[4:4 + total * (sb + counter_size)].reshape(total, sb + counter_size)
    counts = np.zeros(total, np.uint64)
    for b in range(counter_size):
        counts |= rec[:, sb + b].astype(np.uint64) << np.uint64(8 * b)
    # prefix of record r = (index of the last LUT entry <= r) mod 4^p  (LUT entries are starts; empty prefixes repeat a start)
    starts = np.append(lut, total + 1)
    prefix_index = np.searchsorted(starts, np.arange(total), side="right") - 1
    pv = prefix_index % (4 ** p)
    kmers = np.zeros((total, 2), np.uint64)
    nt = 0
    for i in range(p):
        code = ((pv >> (2 * (p - 1 - i))) & 3).astype(np.uint64)
        kmers[:, nt >> 5] |= code << np.uint64(2 * (nt & 31)); nt += 1
    for b in range(sb):
        for sft in (6, 4, 2, 0):
            code = ((rec[:, b] >> sft) & 3).astype(np.uint64)
            kmers[:, nt >> 5] |= code << np.uint64(2 * (nt & 31)); nt += 1
    keep = (counts >= min_count) & (counts <= max_count)
    info = {"kmer_length": int(k), "counter_size": int(counter_size), "lut_prefix_length": int(p), "min_count": int(min_count), "max_count": int(max_count),
            "total_kmers": int(total), "both_strands": not flag, "kmc_version": int(version)}
    return np.ascontiguousarray(kmers[keep]), counts[keep].astype(np.uint32), info
