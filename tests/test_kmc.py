"""KMC databases: include/btgpu_kmc.hpp (the reader the host programs use) and
bayestyper_b200/kmcio.py (KMC1 / KMC2 writers for the synthetic samples) against the REFERENCE's vendored KMC API 2.3.0
(CKMCFile::OpenForListing / ReadNextKmer, driven by `btref kmc-list`): the API must list exactly the k-mers and counts that were
written, and btkmc must list exactly what the API lists, in the same order.  CPU only."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from bayestyper_b200 import kmcio, synth

ROOT = Path(__file__).resolve().parent.parent
BTREF = ROOT / "oracle" / "_ref" / "btref"
K = 55


def _build_btkmc():
    from bayestyper_b200 import build
    build.build_host()
    return ROOT / "host" / "btkmc"


def _strings(kmers):
    codes = kmcio._codes(np.ascontiguousarray(kmers, np.uint64).reshape(-1, 2), K)
    return ["".join("ACGT"[c] for c in row) for row in codes]


def _spectrum(n, seed):
    rng = np.random.default_rng(seed)
    seq = bytes(rng.choice(list(b"ACGT"), n + K - 1).astype(np.uint8))
    km = np.unique(synth.canonical_kmers(seq), axis=0)
    counts = rng.integers(1, 256, size=len(km)).astype(np.uint32)
    counts[:5] = [1, 255, 2, 254, 128]
    return km, counts


def _listing(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return [l for l in r.stdout.splitlines() if not l.startswith("#")]


@pytest.mark.parametrize("layout,p,counter_size", [("kmc1", 7, 1), ("kmc1", 3, 2), ("kmc2", 7, 1), ("kmc2", 11, 4)])
def test_reader_and_writers_agree_with_the_reference_api(tmp_path, layout, p, counter_size):
    exe = _build_btkmc()
    km, counts = _spectrum(6000, 5)
    if counter_size > 1:
        counts = counts * np.uint32(37)                      # counts beyond one byte
    db = tmp_path / "sample"
    maxc = int(counts.max())
    if layout == "kmc1":
        kmcio.write_kmc1(db, km, counts, lut_prefix_length=p, counter_size=counter_size, max_count=maxc)
        o = synth.kmc_order(km)
        want = [f"{s}\t{c}" for s, c in zip(_strings(km[o]), counts[o])]   # a KMC1 database lists in lexicographic order
    else:
        lk, lc = kmcio.write_kmc2(db, km, counts, lut_prefix_length=p, counter_size=counter_size, max_count=maxc, n_bins=8 if p < 11 else 2, seed=3)
        want = [f"{s}\t{c}" for s, c in zip(_strings(lk), lc)]             # KMC2: bin by bin, each bin sorted
    mine = _listing([str(exe), "list", str(db)])
    assert mine == want
    if BTREF.exists():                                       # the reference's own reader (build container; absent on the GPU box)
        ref = _listing([str(BTREF), "kmc-list", "--db", str(db)])
        assert ref == want, "the reference's KMC API does not list what kmcio wrote"
        assert mine == ref
    rk, rc, rinfo = kmcio.read_kmc(db)                       # the numpy twin used by the Python mirror
    assert [f"{s}\t{c}" for s, c in zip(_strings(rk), rc)] == want and rinfo["total_kmers"] == len(km)
    info = subprocess.run([str(exe), "info", str(db)], capture_output=True, text=True).stdout
    assert f"kmer_length {K}" in info and f"total_kmers {len(km)}" in info and f"lut_prefix_length {p}" in info


def test_count_window_and_errors(tmp_path):
    """ReadNextKmer skips records outside [min_count, max_count] (kmc_file.cpp:509-513); broken files fail loudly."""
    exe = _build_btkmc()
    km, counts = _spectrum(2000, 9)
    db = tmp_path / "w"
    kmcio.write_kmc1(db, km, counts, min_count=3, max_count=200)
    o = synth.kmc_order(km)
    keep = (counts[o] >= 3) & (counts[o] <= 200)
    want = [f"{s}\t{c}" for s, c in zip(np.array(_strings(km[o]))[keep], counts[o][keep])]
    assert _listing([str(exe), "list", str(db)]) == want
    rk, rc, _ = kmcio.read_kmc(db)
    assert [f"{s}\t{c}" for s, c in zip(_strings(rk), rc)] == want
    if BTREF.exists():
        assert _listing([str(BTREF), "kmc-list", "--db", str(db)]) == want
    (tmp_path / "bad.kmc_pre").write_bytes(b"KMCPxxxxKMCX")
    (tmp_path / "bad.kmc_suf").write_bytes(b"KMCSKMCS")
    r = subprocess.run([str(exe), "list", str(tmp_path / "bad")], capture_output=True, text=True)
    assert r.returncode == 1 and "not a KMC file" in r.stderr


def test_driver_inputs_from_kmc_files(tmp_path, oracle):
    """driver.inputs_from_kmc: the per-sample files `bayesTyper genotype` reads (<samples>.tsv rows: id, gender, KMC prefix; the KMC
    database; the .bloomMeta / .bloomData pair of makeBloom) become the Python mirror's Inputs without touching the GPU."""
    from bayestyper_b200 import driver
    from tests import _oracle as O
    km, counts = _spectrum(3000, 21)
    counts = np.minimum(counts * 3, 700).astype(np.uint32)            # some counts beyond 255: addSampleCount saturates
    kmcio.write_kmc2(tmp_path / "S1", km, counts, counter_size=2, max_count=700, n_bins=4, seed=1)
    kmcio.write_kmc1(tmp_path / "S2", km[::2], counts[::2], counter_size=2, max_count=700)
    m = oracle.bto_bloom_num_bits(len(km), 1e-3)
    nh = oracle.bto_bloom_num_hashes(m, len(km))
    bits = O.bloom_build(np.ascontiguousarray(km), m, nh)
    (tmp_path / "S1.bloomMeta").write_text(f"{len(km)}\t{m}\t{K}\n")
    bits.tofile(tmp_path / "S1.bloomData")
    ref = b"ACGT" * 100
    inp = driver.inputs_from_kmc("chr1", ref, [], [("S1", "Female", tmp_path / "S1"), ("S2", "M", tmp_path / "S2")])
    assert inp.genders == ["F", "M"] and inp.blooms is None            # S2 has no filter files: filters are built on the device
    (k1, c1), (k2, c2) = inp.spectra
    want = {tuple(k): min(int(c), 255) for k, c in zip(km.tolist(), counts.tolist())}
    assert len(k1) == len(km) and all(want[tuple(k)] == int(c) for k, c in zip(k1.tolist(), c1.tolist()))
    assert len(k2) == len(km[::2]) and c1.dtype == np.uint8 and int(c1.max()) == 255
    inp1 = driver.inputs_from_kmc("chr1", ref, [], [("S1", "F", tmp_path / "S1")])
    (b, n, nbits), = inp1.blooms
    assert n == len(km) and nbits == m and (b == bits).all()
