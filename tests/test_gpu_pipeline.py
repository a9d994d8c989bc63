"""GPU parity of the genotype-side k-mer stages against the reference's own
VariantClusterHaplotypes, dumped by oracle-R after countPathKmers / countInterclusterKmers / the sample k-mer
feed / classifyPathKmers / getHaplotypeCandidates.  Inputs are the reference's graphs, best paths, intercluster
regions and multigroup Bloom; everything the Gibbs sampler consumes must come out identical:
row order (k-mer by k-mer), multiplicity matrix, counts, inter-cluster multiplicities, unique/multicluster lists,
haplotype->allele tables, and the per-variant coverage bitmaps (as sets)."""
import ctypes as C
import hashlib

import numpy as np
import pytest
import torch

from bayestyper_b200 import btd, capi, kmer_pipeline, synth
from tests._fixtures import GOLD
from tests.golden.make_fixtures import PIPE_WORKLOADS

pytestmark = pytest.mark.gpu
K = 55


def _load(name):
    d = btd.read(GOLD / f"{name}.btd")
    g = {k[2:]: v for k, v in d.items() if k.startswith("g.")}
    h = {k[2:]: v for k, v in d.items() if k.startswith("h.")}
    w = PIPE_WORKLOADS[name]()
    spectra = synth.sample_spectra(w, 4, int(d["meta.n_errors"][0]))
    sha = b"".join(hashlib.sha256(k.tobytes() + c.tobytes()).digest() for k, c in spectra)
    assert sha == d["meta.kmer_sha"].tobytes()
    return d, g, h, w, spectra


def _vh_sets(a, rows):
    out = []
    for r in range(rows):
        s = set()
        for e in range(int(a["kmer_vh_off"][r]), int(a["kmer_vh_off"][r + 1])):
            bits = a["vh_bits"][int(a["vh_bits_off"][e]):int(a["vh_bits_off"][e + 1])]
            s.add((int(a["vh_var"][e]), bits.tobytes()))
        out.append(s)
    return out


@pytest.mark.parametrize("name", list(PIPE_WORKLOADS))
def test_haplotype_candidates_identical_to_reference(btg, name):
    d, g, h, w, spectra = _load(name)
    S = len(spectra)
    V = np.diff(g["cl_vertex_off"]).astype(np.int64)
    n_paths = np.diff(g["cl_path_off"]).astype(np.int64) // V
    pipe = kmer_pipeline.KmerPipeline(g, n_paths, g["path_bits"], S, w.genders)
    pipe.enumerate_path_kmers()
    male_ploidy = 1 if w.chrom.lower() in ("x", "chrx") else 2
    regs = d["regions"]
    pipe.scan_regions(w.reference, [(int(a), int(b)) for dec, a, b in regs if dec == 0], 2, male_ploidy, False)
    for s, (km, ct) in enumerate(spectra):
        kd = torch.from_numpy(km.view(np.int64)).cuda()
        cdv = torch.from_numpy(ct).cuda()
        pipe.add_sample(s, kd, cdv)
    mg = capi.check(btg.btg_bloom_from_bytes(capi.ptr(d["mg.data"]), int(d["mg.meta"][0]), int(d["mg.meta"][1]), K), btg)
    u = pipe.build_unit(multigroup_bloom=mg)
    btg.btg_bloom_free(mg)
    a = u.a
    assert (a["cl_nhap"] == np.diff(h["cl_hap_off"])).all()
    assert (a["cl_kmer_off"] == h["cl_kmer_off"]).all(), "row counts per cluster differ"
    assert (u.kmer_words.reshape(-1) == h["kmer_words"]).all(), "row order / k-mer identity differs"
    for k in ("mult", "k_has_counts", "k_counts", "k_ic", "cl_uniq_off", "uniq_idx", "cl_multi_off", "multi_idx", "hap_alleles", "cl_mult_off", "cl_hapvar_off"):
        assert (a[k] == h[k]).all(), k
    rows = len(a["k_has_counts"])
    assert (a["kmer_vh_off"] == h["kmer_vh_off"]).all()
    assert _vh_sets(a, rows) == _vh_sets(h, rows)
    # flags of the table entries that have records
    fl = h["k_flags"]
    assert ((fl & 2) != 0).sum() == len(a["multi_idx"])
    # nested clusters: per-haplotype nested cluster lists and the dependency map (VariantClusterGraph.cpp:1006-1010, 1112-1133)
    for k in ("hap_nested_off", "hap_nested", "cl_dep_off", "dep_cluster", "dep_var_off", "dep_var"):
        assert (a[k] == h[k]).all(), k
    # multicluster rows: one shared record per k-mer, exactly on the rows the reference flags
    sh = a["k_shared"] != 0xFFFFFFFF
    assert (sh == ((fl & 2) != 0)).all()
    if sh.any():
        words = u.kmer_words[sh]
        ids = a["k_shared"][sh]
        by_id = {}
        for wd, i in zip(map(bytes, words), ids.tolist()):
            assert by_id.setdefault(i, wd) == wd
        assert len(set(by_id.values())) == len(by_id)


def _sync(btg):
    capi.check(btg.btg_kmer_hash(capi.ptr(np.zeros((1, 2), np.uint64)), 1, capi.ptr(np.zeros(1, np.uint64))), btg)  # syncs the library stream


def _random_kmers(rng, n):
    km = rng.integers(0, 2**64, size=(n, 2), dtype=np.uint64)
    km[:, 1] &= np.uint64((1 << 46) - 1)
    return km


def test_table_lookup_and_saturation(btg):
    """btg_table_keys_*, btg_table_lookup_dev / add_sample: key order = lexicographic k-mer order (KMC record order), hits,
    misses, duplicates saturate at 255 (KmerCounts.cpp:178-189); with and without a prefix index, sorted and shuffled streams."""
    from bayestyper_b200 import synth
    capi.check(btg.btg_table_set_index_dev(None, 0), btg)
    rng = np.random.default_rng(3)
    km = _random_kmers(rng, 5000)
    km = km[synth.kmc_order(km)]                                            # the order a KMC database lists them in
    kmd = torch.from_numpy(km.view(np.int64)).cuda()
    kw0 = torch.empty(len(km), dtype=torch.int64, device="cuda"); kw1 = torch.empty_like(kw0)
    torch.cuda.synchronize()
    capi.check(btg.btg_table_keys_from_kmers_dev(kmd.data_ptr(), len(km), kw0.data_ptr(), kw1.data_ptr(), None), btg)
    _sync(btg)
    lo, hi = kw0.cpu().numpy(), kw1.cpu().numpy()
    assert (np.lexsort((lo, hi)) == np.arange(len(km))).all(), "table key order is not the lexicographic k-mer order"
    ehi, elo = synth.lexicographic_words(km)
    assert (hi.view(np.uint64) == ehi).all() and ((lo.view(np.uint64) ^ np.uint64(1 << 63)) == elo).all()
    back = torch.empty_like(kmd)
    capi.check(btg.btg_table_keys_to_kmers_dev(kw0.data_ptr(), kw1.data_ptr(), len(km), back.data_ptr(), None), btg)
    _sync(btg)
    assert (back.cpu().numpy().view(np.uint64) == km).all()
    q = np.concatenate([km[::7], _random_kmers(rng, 300)])
    qd = torch.from_numpy(q.view(np.int64)).cuda()
    idx = torch.zeros(len(q), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    capi.check(btg.btg_table_lookup_dev(kw0.data_ptr(), kw1.data_ptr(), len(km), qd.data_ptr(), len(q), idx.data_ptr(), None), btg)
    exp = np.concatenate([np.arange(0, len(km), 7), -np.ones(300, np.int64)])
    _sync(btg)
    assert (idx.cpu().numpy() == exp).all()

    def stream(records, cts, n_samples, sample, use_lut_bits):
        counts = torch.zeros((len(km), n_samples), dtype=torch.uint8, device="cuda")
        rec = torch.zeros(len(km), dtype=torch.uint8, device="cuda")
        rd = torch.from_numpy(records.view(np.int64)).cuda(); cd_ = torch.from_numpy(cts).cuda()
        lut_d = None
        if use_lut_bits:
            lut = np.concatenate([[0], np.cumsum(np.bincount(hi >> (46 - use_lut_bits), minlength=1 << use_lut_bits))]).astype(np.int64)
            lut_d = torch.from_numpy(lut).cuda()
        torch.cuda.synchronize()
        capi.check(btg.btg_table_set_index_dev(lut_d.data_ptr() if use_lut_bits else None, use_lut_bits), btg)
        capi.check(btg.btg_table_add_sample_kmers_dev(kw0.data_ptr(), kw1.data_ptr(), len(km), rd.data_ptr(), cd_.data_ptr(), len(cts), n_samples, sample,
                                                      counts.data_ptr(), rec.data_ptr(), None), btg)
        _sync(btg)
        capi.check(btg.btg_table_set_index_dev(None, 0), btg)
        return counts.cpu().numpy(), rec.cpu().numpy()

    dup = np.concatenate([km[:10]] * 3)                                      # each of 10 k-mers three times
    for bits in (0, 4, 10, 16):                                              # no index / long buckets (binary search) / short buckets (linear)
        c, r = stream(dup, np.full(30, 100, np.uint8), 3, 1, bits)
        assert (c[:10, 1] == 255).all() and c[:10, [0, 2]].sum() == 0 and c[10:].sum() == 0
        assert r[:10].all() and not r[10:].any()
    # a whole "database": every third key present + misses, KMC-ordered and shuffled, ragged length (not a multiple of the batch)
    recs = np.concatenate([km[::3], _random_kmers(rng, 2001)])
    cts = rng.integers(1, 255, size=len(recs)).astype(np.uint8)
    expc = np.zeros(len(km), np.uint8); expc[::3] = cts[:len(km[::3])]
    o = synth.kmc_order(recs)
    for order in (o, rng.permutation(len(recs))):
        for bits in (0, 12):
            c, r = stream(np.ascontiguousarray(recs[order]), np.ascontiguousarray(cts[order]), 1, 0, bits)
            assert (c[:, 0] == expc).all() and (r == (expc > 0)).all()
    # through a prefix index: lookups
    bits = 10
    lut = np.concatenate([[0], np.cumsum(np.bincount(hi >> (46 - bits), minlength=1 << bits))]).astype(np.int64)
    lut_d = torch.from_numpy(lut).cuda(); idx2 = torch.zeros(len(q), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    capi.check(btg.btg_table_set_index_dev(lut_d.data_ptr(), bits), btg)
    capi.check(btg.btg_table_lookup_dev(kw0.data_ptr(), kw1.data_ptr(), len(km), qd.data_ptr(), len(q), idx2.data_ptr(), None), btg)
    _sync(btg)
    capi.check(btg.btg_table_set_index_dev(None, 0), btg)
    assert (idx2.cpu().numpy() == exp).all()


def test_stream_full_size_properties(btg):
    """The sample-stream kernel at a size the tiled path dominates (2 M keys, 6 M records), through size-independent properties: the
    table after a KMC-ordered stream equals the table after the same records shuffled (order independence), the number of keys
    with a record equals the number of distinct present records, and the sum of the counts equals the sum over present records."""
    from bayestyper_b200 import synth
    rng = np.random.default_rng(11)
    n_keys, n_hit, n_miss = 2_000_000, 2_400_000, 3_600_000
    km = _random_kmers(rng, n_keys)
    km = km[synth.kmc_order(km)]
    kmd = torch.from_numpy(km.view(np.int64)).cuda()
    kw0 = torch.empty(n_keys, dtype=torch.int64, device="cuda"); kw1 = torch.empty_like(kw0)
    torch.cuda.synchronize()
    capi.check(btg.btg_table_keys_from_kmers_dev(kmd.data_ptr(), n_keys, kw0.data_ptr(), kw1.data_ptr(), None), btg)
    _sync(btg)
    hi = kw1.cpu().numpy()
    bits = 20
    lut = np.concatenate([[0], np.cumsum(np.bincount(hi >> (46 - bits), minlength=1 << bits))]).astype(np.int64)
    lut_d = torch.from_numpy(lut).cuda()
    pick = rng.integers(0, n_keys, size=n_hit)                       # present records, with repeats (saturating adds)
    recs = np.concatenate([km[pick], _random_kmers(rng, n_miss)])
    cts = rng.integers(1, 40, size=len(recs)).astype(np.uint8)
    expect = np.minimum(np.bincount(pick, weights=cts[:n_hit].astype(np.float64), minlength=n_keys), 255).astype(np.uint8)
    tables = []
    for order in (synth.kmc_order(recs), rng.permutation(len(recs))):
        rd = torch.from_numpy(np.ascontiguousarray(recs[order]).view(np.int64)).cuda(); cd_ = torch.from_numpy(np.ascontiguousarray(cts[order])).cuda()
        counts = torch.zeros(n_keys, dtype=torch.uint8, device="cuda"); rec = torch.zeros(n_keys, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        capi.check(btg.btg_table_set_index_dev(lut_d.data_ptr(), bits), btg)
        capi.check(btg.btg_table_add_sample_kmers_dev(kw0.data_ptr(), kw1.data_ptr(), n_keys, rd.data_ptr(), cd_.data_ptr(), len(recs), 1, 0, counts.data_ptr(),
                                                      rec.data_ptr(), None), btg)
        _sync(btg)
        capi.check(btg.btg_table_set_index_dev(None, 0), btg)
        tables.append((counts.cpu().numpy(), rec.cpu().numpy()))
    (c0, r0), (c1, r1) = tables
    assert (c0 == c1).all() and (r0 == r1).all(), "the table depends on the order of the stream"
    assert int(r0.sum()) == len(np.unique(pick)) and int(c0.astype(np.int64).sum()) == int(expect.astype(np.int64).sum())
    assert (c0 == expect).all()
