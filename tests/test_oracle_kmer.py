"""Pins oracle/kmer_oracle.c to the known-answer vectors of SURVEY.md §4.

Those vectors were produced from the UNMODIFIED reference sources (g++ 13.3,
-DBT_KMER_SIZE=55: Kmer.tpp, nthash.hpp, BloomFilter.hpp, KmerBloom.cpp); the
reference itself ships no test vectors.  tests/test_ref_parity.py
additionally checks the oracle against oracle/_ref when that is built.
"""
import json
from pathlib import Path

import numpy as np

from tests import _oracle as O

GOLD = json.loads((Path(__file__).parent / "golden" / "kmer_kat.json").read_text())


def test_canonical_and_hash_vectors(oracle):
    seq = GOLD["seq60"]
    kmers, pos = O.scan(seq.encode())
    assert len(kmers) == len(GOLD["windows"]) == 6
    for w, km, p in zip(GOLD["windows"], kmers, pos):
        assert p == w["w"] + 54
        assert O.unpack(km) == w["canonical"]
        fwd = seq[w["w"]:w["w"] + 55]
        assert (O.unpack(km) == fwd) == bool(w["fwd"])
        assert oracle.bto_ntp64(km, 55) == int(w["ntp64"], 16)
        assert oracle.bto_ntp64_seeded(km, 55, 1029283129) % 65536 == w["root"]
        assert oracle.bto_threaded_bloom_root(km, 55) == w["root"]
        # rolling check: F = hash of forward window, R = hash of its reverse complement
        f = O.pack(fwd)
        assert oracle.bto_ntp64(f, 55) == int(w["F"], 16)
        assert oracle.bto_nt_rhval(f, 55) == int(w["R"], 16)


def test_extra_hash_vector(oracle):
    km = O.pack("AAAC" + "A" * 50 + "T")
    assert oracle.bto_ntp64(km, 55) == int(GOLD["extra"]["ntp64"], 16)


def test_bloom_sizing_vectors(oracle):
    for row in GOLD["sizing"]:
        bits = oracle.bto_bloom_num_bits(row["n"], row["fpr"])
        assert bits == row["bits"], row
        assert oracle.bto_bloom_num_hashes(bits, row["n"]) == row["h"], row


def test_probe_locations_vector(oracle):
    km = O.pack(GOLD["windows"][0]["canonical"])
    locs = np.zeros(10, np.uint64)
    oracle.bto_bloom_locs(km, 55, GOLD["probe"]["m"], 10, locs)
    assert locs.tolist() == GOLD["probe"]["locs"]


def test_insert_lookup_roundtrip(oracle):
    kmers = O.random_kmers(5000, 1)
    m = oracle.bto_bloom_num_bits(5000, 1e-3)
    nh = oracle.bto_bloom_num_hashes(m, 5000)
    bits = O.bloom_build(kmers, m, nh)
    assert O.bloom_lookup(bits, m, nh, kmers).all()
    other = O.random_kmers(20000, 2)
    hit, probes = O.bloom_lookup(bits, m, nh, other, want_probes=True)
    assert hit.mean() < 5e-3                      # fpr 1e-3 design
    assert 1 <= probes.min() and probes.max() <= nh
    assert (probes[hit == 1] == nh).all()


def test_scan_resets_on_non_acgt(oracle):
    seq = O.random_seq(300, 3)
    seq = seq[:100] + b"N" + seq[101:]
    kmers, pos = O.scan(seq)
    # windows overlapping position 100 are skipped
    assert set(pos.tolist()) == set(range(54, 100)) | set(range(155, 300))
    assert len(O.scan(b"")[0]) == 0
    assert len(O.scan(b"ACGT" * 13)[0]) == 0     # 52 nt < k


def test_canonical_is_min_of_strand_strings(oracle):
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    for km in O.random_kmers(200, 4):
        s = O.unpack(km)
        rc = "".join(comp[c] for c in reversed(s))
        out = np.zeros(2, np.uint64)
        oracle.bto_canonical(km, 55, out)
        assert O.unpack(out) == min(s, rc)


def _perturbed(kmers, seed):
    """One random bit of the 110 flipped per k-mer (MakeBloom::testbloom, MakeBloom.cpp:333-334)."""
    rng = np.random.default_rng(seed)
    bit = rng.integers(0, 2 * 55, len(kmers))
    out = kmers.copy()
    lo = bit < 64
    out[lo, 0] ^= np.uint64(1) << bit[lo].astype(np.uint64)
    out[~lo, 1] ^= np.uint64(1) << (bit[~lo] - 64).astype(np.uint64)
    return out


def test_bloom_self_test_rates(oracle):
    """The reference's own (disabled) Bloom self-test, MakeBloom::testbloom (src/bayesTyperTools/MakeBloom.cpp:311-375): every inserted k-mer is
    found; k-mers one bit away from an inserted one and random k-mers that are not in the set are found at the design false-positive rate."""
    n, fpr = 100_000, 1e-3
    kmers = O.random_kmers(n, 11)
    m = oracle.bto_bloom_num_bits(n, fpr)
    nh = oracle.bto_bloom_num_hashes(m, n)
    bits = O.bloom_build(kmers, m, nh)
    assert O.bloom_lookup(bits, m, nh, kmers).all()                                  # original: all found
    present = {tuple(k) for k in kmers.tolist()}
    for probe in (_perturbed(kmers, 12), O.random_kmers(n, 13)):
        absent = np.array([tuple(k) not in present for k in probe.tolist()])
        assert absent.mean() > 0.999                                                 # a 110-bit k-mer space: collisions with the set are not expected
        fp = O.bloom_lookup(bits, m, nh, probe[absent]).mean()
        assert fpr / 2 < fp < fpr * 2, fp                                            # 100 expected hits: +-100 % is > 6 sigma
