"""GPU parity (bit-exact) of the k-mer match primitives against the oracle,
through the C ABI.  Covers SURVEY.md §8 rows a1-a6 and the scan loop of a9/a10."""
import ctypes as C

import numpy as np
import pytest

from bayestyper_b200 import capi
from tests import _oracle as O

pytestmark = pytest.mark.gpu
K = 55


def _info(btg, b):
    n, m, h = C.c_uint64(), C.c_uint64(), C.c_uint32()
    capi.check(btg.btg_bloom_info(b, C.byref(n), C.byref(m), C.byref(h)))
    return n.value, m.value, h.value


def test_hash_and_canonical_bit_exact(btg, oracle):
    kmers = O.random_kmers(200_000, 1)
    h = np.zeros(len(kmers), np.uint64)
    capi.check(btg.btg_kmer_hash(capi.ptr(kmers), len(kmers), capi.ptr(h)))
    ref = np.array([oracle.bto_ntp64(km, K) for km in kmers[:5000]], np.uint64)
    assert (h[:5000] == ref).all()
    canon = np.zeros_like(kmers)
    capi.check(btg.btg_kmer_canonical(capi.ptr(kmers), len(kmers), capi.ptr(canon)))
    out = np.zeros(2, np.uint64)
    for km, c in zip(kmers[:5000], canon[:5000]):
        oracle.bto_canonical(km, K, out)
        assert (out == c).all()
    # idempotence at full size
    canon2 = np.zeros_like(kmers)
    capi.check(btg.btg_kmer_canonical(capi.ptr(canon), len(canon), capi.ptr(canon2)))
    assert (canon == canon2).all()


def test_golden_vectors_on_device(btg):
    import json
    from pathlib import Path
    gold = json.loads((Path(__file__).parent / "golden" / "kmer_kat.json").read_text())
    seq = gold["seq60"].encode()
    kmers = np.zeros((len(seq), 2), np.uint64)
    valid = np.zeros(len(seq), np.uint8)
    capi.check(btg.btg_scan_sequence(seq, len(seq), capi.ptr(kmers), capi.ptr(valid)))
    assert valid.tolist() == [0] * 54 + [1] * 6
    got = kmers[54:]
    for w, km in zip(gold["windows"], got):
        assert O.unpack(km) == w["canonical"]
    h = np.zeros(6, np.uint64)
    got = np.ascontiguousarray(got)
    capi.check(btg.btg_kmer_hash(capi.ptr(got), 6, capi.ptr(h)))
    assert [f"{int(x):016x}" for x in h] == [w["ntp64"] for w in gold["windows"]]


@pytest.mark.parametrize("n,fpr", [(1, 1e-3), (1000, 1e-3), (200_000, 1e-3), (50_000, 1e-4)])
def test_bloom_insert_bits_identical_and_lookup(btg, oracle, n, fpr):
    kmers = O.random_kmers(n, n)
    b = capi.check(btg.btg_bloom_create(n, fpr, K))
    nk, m, nh = _info(btg, b)
    assert m == oracle.bto_bloom_num_bits(n, fpr) and nh == oracle.bto_bloom_num_hashes(m, n)
    capi.check(btg.btg_bloom_insert(b, capi.ptr(kmers), n))
    dev_bits = np.zeros((m + 7) // 8, np.uint8)
    capi.check(btg.btg_bloom_download(b, capi.ptr(dev_bits), dev_bits.size))
    ref_bits = O.bloom_build(kmers, m, nh)
    assert (dev_bits == ref_bits).all()          # byte-identical .bloomData
    probe = np.concatenate([kmers[: min(n, 50_000)], O.random_kmers(100_000, n + 7)])
    hit = np.zeros(len(probe), np.uint8)
    capi.check(btg.btg_bloom_lookup(b, capi.ptr(probe), len(probe), capi.ptr(hit)))
    assert (hit == O.bloom_lookup(ref_bits, m, nh, probe)).all()
    btg.btg_bloom_free(b)


def test_bloom_file_roundtrip_and_reference_format(btg, oracle, tmp_path):
    n = 30_000
    kmers = O.random_kmers(n, 3)
    m = oracle.bto_bloom_num_bits(n, 1e-3)
    nh = oracle.bto_bloom_num_hashes(m, n)
    ref_bits = O.bloom_build(kmers, m, nh)
    # a filter written in the reference's on-disk format (KmerBloom::save)
    prefix = str(tmp_path / "s1")
    open(prefix + ".bloomMeta", "w").write(f"{n}\t{m}\t55\n")
    ref_bits.tofile(prefix + ".bloomData")
    b = capi.check(btg.btg_bloom_load(prefix.encode(), K))
    assert _info(btg, b) == (n, m, nh)
    probe = np.concatenate([kmers[:1000], O.random_kmers(5000, 9)])
    hit = np.zeros(len(probe), np.uint8)
    capi.check(btg.btg_bloom_lookup(b, capi.ptr(probe), len(probe), capi.ptr(hit)))
    assert (hit == O.bloom_lookup(ref_bits, m, nh, probe)).all()
    out = str(tmp_path / "out")
    capi.check(btg.btg_bloom_save(b, out.encode()))
    assert open(out + ".bloomMeta").read() == f"{n}\t{m}\t55\n"
    assert (np.fromfile(out + ".bloomData", np.uint8) == ref_bits).all()
    btg.btg_bloom_free(b)
    assert btg.btg_bloom_load(str(tmp_path / "missing").encode(), K) is None
    assert b"Unable to open" in btg.btg_last_error()
    assert btg.btg_bloom_load(prefix.encode(), 31) is None


def test_probe_counts_match_reference_early_exit(btg, oracle):
    import torch
    n = 100_000
    kmers = O.random_kmers(n, 5)
    b = capi.check(btg.btg_bloom_create(n, 1e-3, K))
    capi.check(btg.btg_bloom_insert(b, capi.ptr(kmers), n))
    _, m, nh = _info(btg, b)
    probe = np.concatenate([kmers[:20_000], O.random_kmers(80_000, 6)])
    d_k = torch.from_numpy(probe.view(np.int64)).cuda()
    d_hit = torch.zeros(len(probe), dtype=torch.uint8, device="cuda")
    d_pr = torch.zeros(len(probe), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    capi.check(btg.btg_bloom_lookup_probes_dev(b, d_k.data_ptr(), len(probe), d_hit.data_ptr(), d_pr.data_ptr(), None))
    import ctypes
    torch.cuda.synchronize()
    capi.check(btg.btg_bloom_lookup_dev(b, d_k.data_ptr(), 0, d_hit.data_ptr(), None))  # empty input is a no-op
    # library stream is non-blocking: order against it explicitly
    dev_bits = np.zeros((m + 7) // 8, np.uint8)
    capi.check(btg.btg_bloom_download(b, capi.ptr(dev_bits), dev_bits.size))  # syncs the library stream
    hit, probes = O.bloom_lookup(dev_bits, m, nh, probe, want_probes=True)
    assert (d_hit.cpu().numpy() == hit).all()
    assert (d_pr.cpu().numpy() == probes).all()
    btg.btg_bloom_free(b)


def test_threaded_bloom_matches_subfilter_model(btg, oracle):
    n = 300_000
    kmers = O.random_kmers(n, 8)
    t = capi.check(btg.btg_tbloom_create(n, 1e-4, K))
    sk, sb, nh = C.c_uint64(), C.c_uint64(), C.c_uint32()
    capi.check(btg.btg_tbloom_info(t, C.byref(sk), C.byref(sb), C.byref(nh)))
    assert sk.value == oracle.bto_threaded_bloom_sub_kmers(n)
    assert sb.value == oracle.bto_bloom_num_bits(sk.value, 1e-4)
    assert nh.value == oracle.bto_bloom_num_hashes(sb.value, sk.value)
    capi.check(btg.btg_tbloom_insert(t, capi.ptr(kmers), n))
    sub_bytes = (sb.value + 7) // 8
    bits = np.zeros((65536, sub_bytes), np.uint8)
    capi.check(btg.btg_tbloom_download(t, capi.ptr(bits), bits.size))
    # oracle model: route each k-mer to sub-filter rootIndex, then KmerBloom insert
    ref = np.zeros_like(bits)
    sel = kmers[:20_000]
    roots = np.array([oracle.bto_threaded_bloom_root(km, K) for km in sel])
    for r in np.unique(roots):
        ref[r] = O.bloom_build(sel[roots == r], sb.value, nh.value)
    # every bit the oracle sets for the subset must be set on the device
    assert ((bits & ref) == ref).all()
    t2 = capi.check(btg.btg_tbloom_create(n, 1e-4, K))
    capi.check(btg.btg_tbloom_insert(t2, capi.ptr(sel), len(sel)))
    bits2 = np.zeros_like(bits)
    capi.check(btg.btg_tbloom_download(t2, capi.ptr(bits2), bits2.size))
    assert (bits2 == ref).all()                  # exact equality on the same input set
    probe = np.concatenate([sel[:5000], O.random_kmers(20_000, 10)])
    hit = np.zeros(len(probe), np.uint8)
    capi.check(btg.btg_tbloom_lookup(t2, capi.ptr(probe), len(probe), capi.ptr(hit)))
    proots = np.array([oracle.bto_threaded_bloom_root(km, K) for km in probe])
    exp = np.array([O.bloom_lookup(ref[r], sb.value, nh.value, km[None, :])[0] for km, r in zip(probe, proots)])
    assert (hit == exp).all()
    btg.btg_tbloom_free(t)
    btg.btg_tbloom_free(t2)


@pytest.mark.parametrize("n_frac,length", [(0.0, 100_000), (0.02, 100_000), (0.0, 54), (0.0, 55), (1.0, 500)])
def test_scan_sequence_bit_exact(btg, oracle, n_frac, length):
    seq = O.random_seq(length, 31, n_frac)
    kmers = np.zeros((length, 2), np.uint64)
    valid = np.zeros(length, np.uint8)
    capi.check(btg.btg_scan_sequence(seq, length, capi.ptr(kmers), capi.ptr(valid)))
    ref_k, ref_p = O.scan(seq)
    assert valid.sum() == len(ref_k)
    assert (np.nonzero(valid)[0] == ref_p).all()
    assert (kmers[valid == 1] == ref_k).all()
    assert (kmers[valid == 0] == 0).all()


def test_scan_lookup_fused_equals_scan_then_lookup(btg, oracle):
    import torch
    seq = O.random_seq(400_000, 41, 0.001)
    ref_k, ref_p = O.scan(seq)
    n = 50_000
    present = ref_k[:: len(ref_k) // n][:n]
    b = capi.check(btg.btg_bloom_create(n, 1e-3, K))
    capi.check(btg.btg_bloom_insert(b, capi.ptr(np.ascontiguousarray(present)), len(present)))
    _, m, nh = _info(btg, b)
    bits = np.zeros((m + 7) // 8, np.uint8)
    capi.check(btg.btg_bloom_download(b, capi.ptr(bits), bits.size))
    d_seq = torch.frombuffer(bytearray(seq), dtype=torch.uint8).cuda()
    d_hit = torch.zeros(len(seq), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    capi.check(btg.btg_scan_sequence_lookup_dev(b, d_seq.data_ptr(), len(seq), d_hit.data_ptr(), None))
    capi.check(btg.btg_bloom_download(b, capi.ptr(bits), bits.size))   # sync
    exp = np.zeros(len(seq), np.uint8)
    exp[ref_p] = O.bloom_lookup(bits, m, nh, ref_k)
    assert (d_hit.cpu().numpy() == exp).all()
    btg.btg_bloom_free(b)


def test_full_size_properties(btg):
    """BASELINE-size checks through size-independent properties: inserted k-mers are
    always found; canonical(k) == canonical(revcomp(k)) via scan of both strands."""
    n = 8_000_000
    kmers = O.random_kmers(n, 99)
    b = capi.check(btg.btg_bloom_create(n, 1e-3, K))
    capi.check(btg.btg_bloom_insert(b, capi.ptr(kmers), n))
    hit = np.zeros(n, np.uint8)
    capi.check(btg.btg_bloom_lookup(b, capi.ptr(kmers), n, capi.ptr(hit)))
    assert hit.all()
    other = O.random_kmers(2_000_000, 100)
    hit2 = np.zeros(len(other), np.uint8)
    capi.check(btg.btg_bloom_lookup(b, capi.ptr(other), len(other), capi.ptr(hit2)))
    assert 2e-4 < hit2.mean() < 3e-3             # designed fpr 1e-3
    btg.btg_bloom_free(b)
    seq = O.random_seq(2_000_000, 5)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rc = seq.translate(comp)[::-1]
    k1 = np.zeros((len(seq), 2), np.uint64); v1 = np.zeros(len(seq), np.uint8)
    k2 = np.zeros((len(seq), 2), np.uint64); v2 = np.zeros(len(seq), np.uint8)
    capi.check(btg.btg_scan_sequence(seq, len(seq), capi.ptr(k1), capi.ptr(v1)))
    capi.check(btg.btg_scan_sequence(rc, len(rc), capi.ptr(k2), capi.ptr(v2)))
    assert (k1[54:] == k2[54:][::-1]).all()


def test_bloom_self_test_rates_on_device(btg, oracle):
    """The reference's own (disabled) Bloom self-test, MakeBloom::testbloom (src/bayesTyperTools/MakeBloom.cpp:311-375), through btg_bloom_*:
    every inserted k-mer is found; k-mers one bit away from an inserted one, and random k-mers, are found at the design false-positive rate —
    and k-mer for k-mer where the oracle's filter finds them."""
    from tests.test_oracle_kmer import _perturbed
    n, fpr = 100_000, 1e-3
    kmers = O.random_kmers(n, 11)
    b = capi.check(btg.btg_bloom_create(n, fpr, K))
    nk, m, nh = _info(btg, b)
    capi.check(btg.btg_bloom_insert(b, capi.ptr(kmers), n))
    ref_bits = O.bloom_build(kmers, m, nh)
    for probe, lo, hi in ((kmers, 1.0, 1.0), (_perturbed(kmers, 12), fpr / 2, fpr * 2), (O.random_kmers(n, 13), fpr / 2, fpr * 2)):
        probe = np.ascontiguousarray(probe)
        hit = np.zeros(len(probe), np.uint8)
        capi.check(btg.btg_bloom_lookup(b, capi.ptr(probe), len(probe), capi.ptr(hit)))
        assert (hit == O.bloom_lookup(ref_bits, m, nh, probe)).all()
        assert lo <= hit.mean() <= hi, hit.mean()
    btg.btg_bloom_free(b)
