"""The text files of `<out>_cluster_data/` that `bayesTyper cluster` leaves for `bayesTyper genotype`:

* `intercluster_regions.txt.gz` — gzip text, `contig \\t is_decoy(0/1) \\t start \\t end` (0-based, inclusive), longest region first
  (written VariantFileParser.cpp:59-65,1186-1212, read KmerCounter.cpp:346-372);
* `parameter_kmers.fa.gz` — gzip text, `>k55` then up to `max_parameter_kmers` canonical 55-mers, one per line
  (written KmerHash.cpp:137-178, read main.cpp:543-581).

The third file of that directory, `multigroup_kmers.bloomMeta/.bloomData`, is a KmerBloom file (btg_bloom_save / btg_bloom_load).
K-mers cross this module in the boundary layout of include/btgpu.h ((n, 2) uint64 words of bitset<110>, nucleotide i in bits
2i, 2i+1 with A=0 C=1 G=2 T=3).
"""
from __future__ import annotations

import gzip

import numpy as np

K = 55
_ACGT = np.frombuffer(b"ACGT", np.uint8)
_CODE = np.full(256, 255, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i


def write_intercluster_regions(prefix, regions) -> None:
    """regions: (contig, is_decoy, start, end).  Sorted by length, longest first (std::sort in the reference: the order of regions of
    equal length is unspecified there; input order is kept here)."""
    rows = sorted(regions, key=lambda r: -(r[3] - r[2]))
    with gzip.GzipFile(str(prefix) + ".txt.gz", "wb", mtime=0) as f:
        for contig, decoy, start, end in rows:
            if start > end:
                raise ValueError("intercluster region with start > end")
            f.write(f"{contig}\t{int(bool(decoy))}\t{int(start)}\t{int(end)}\n".encode())


def read_intercluster_regions(prefix) -> list:
    """[(contig, is_decoy, start, end)] in file order."""
    out = []
    with gzip.open(str(prefix) + ".txt.gz", "rt") as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            if len(t) != 4:
                raise ValueError(f"intercluster region line with {len(t)} columns")
            out.append((t[0], bool(int(t[1])), int(t[2]), int(t[3])))
    return out


def kmers_to_strings(kmers: np.ndarray) -> np.ndarray:
    """(n, 2) uint64 boundary words -> (n, 55) uint8 nucleotide characters (Nucleotide::bitToNt)."""
    kmers = np.ascontiguousarray(kmers, np.uint64).reshape(-1, 2)
    i = np.arange(K)
    word = np.where(i < 32, 0, 1)
    shift = (2 * (i % 32)).astype(np.uint64)
    codes = (kmers[:, word] >> shift) & np.uint64(3)
    return _ACGT[codes.astype(np.int64)]


def strings_to_kmers(chars: np.ndarray) -> np.ndarray:
    """(n, 55) uint8 nucleotide characters -> (n, 2) uint64 boundary words (Nucleotide::ntToBit); non-ACGT raises."""
    chars = np.ascontiguousarray(chars, np.uint8).reshape(-1, K)
    codes = _CODE[chars]
    if (codes == 255).any():
        raise ValueError("parameter k-mer with a character outside ACGT")
    codes = codes.astype(np.uint64)
    out = np.zeros((len(chars), 2), np.uint64)
    for i in range(K):
        out[:, i // 32] |= codes[:, i] << np.uint64(2 * (i % 32))
    return out


def write_parameter_kmers(prefix, kmers: np.ndarray, max_kmers: int = 1_000_000) -> int:
    """The first max_kmers k-mers, in the order given (the reference writes its shuffled hash in bucket order)."""
    chars = kmers_to_strings(np.asarray(kmers)[:max_kmers])
    lines = np.concatenate([chars, np.full((len(chars), 1), ord("\n"), np.uint8)], axis=1)
    with gzip.GzipFile(str(prefix) + ".fa.gz", "wb", mtime=0) as f:
        f.write(f">k{K}\n".encode())
        f.write(lines.tobytes())
    return len(chars)


def read_parameter_kmers(prefix) -> np.ndarray:
    raw = open(str(prefix) + ".fa.gz", "rb").read()
    if raw[:2] == b"\x1f\x8b":           # gzip as the reference writes it; a plain-text file (oracle-R's iostreams shim does not compress) is accepted too
        raw = gzip.decompress(raw)
    nl = raw.find(b"\n")
    head, body = raw[:nl], raw[nl + 1:]
    if head != f">k{K}".encode():
        raise ValueError(f"parameter k-mer file starts with {head!r}, expected >k{K}")
    if len(body) % (K + 1):
        raise ValueError("parameter k-mer file holds a line that is not 55 nucleotides long")
    lines = np.frombuffer(body, np.uint8).reshape(-1, K + 1)
    if len(lines) and (lines[:, K] != ord("\n")).any():
        raise ValueError("parameter k-mer file holds a line that is not 55 nucleotides long")
    return strings_to_kmers(lines[:, :K])
