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


def test_table_lookup_and_saturation(btg):
    """btg_table_lookup_dev / add_sample: hits, misses, duplicates saturate at 255 (KmerCounts.cpp:178-189)."""
    capi.check(btg.btg_table_set_index_dev(None, 0), btg)
    rng = np.random.default_rng(3)
    keys = rng.integers(-2**63, 2**63 - 1, size=(5000, 2), dtype=np.int64)
    keys[:, 1] &= (1 << 46) - 1
    o = np.lexsort((keys[:, 0], keys[:, 1]))
    keys = keys[o]
    kw0 = torch.from_numpy(keys[:, 0].copy()).cuda(); kw1 = torch.from_numpy(keys[:, 1].copy()).cuda()
    q = np.concatenate([keys[::7], rng.integers(-2**63, 2**63 - 1, size=(300, 2), dtype=np.int64)])
    q[-300:, 1] &= (1 << 46) - 1
    qd = torch.from_numpy(q).cuda()
    idx = torch.zeros(len(q), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    capi.check(btg.btg_table_lookup_dev(kw0.data_ptr(), kw1.data_ptr(), len(keys), qd.data_ptr(), len(q), idx.data_ptr(), None), btg)
    exp = np.concatenate([np.arange(0, len(keys), 7), -np.ones(300, np.int64)])
    capi.check(btg.btg_kmer_hash(capi.ptr(np.zeros((1, 2), np.uint64)), 1, capi.ptr(np.zeros(1, np.uint64))), btg)  # syncs the library stream
    assert (idx.cpu().numpy() == exp).all()
    counts = torch.zeros((len(keys), 3), dtype=torch.uint8, device="cuda")
    rec = torch.zeros(len(keys), dtype=torch.uint8, device="cuda")
    dup = torch.from_numpy(np.concatenate([keys[:10]] * 3)).cuda()          # each of 10 k-mers three times
    cts = torch.full((30,), 100, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    capi.check(btg.btg_table_add_sample_kmers_dev(kw0.data_ptr(), kw1.data_ptr(), len(keys), dup.data_ptr(), cts.data_ptr(), 30, 3, 1, counts.data_ptr(), rec.data_ptr(), None), btg)
    capi.check(btg.btg_kmer_hash(capi.ptr(np.zeros((1, 2), np.uint64)), 1, capi.ptr(np.zeros(1, np.uint64))), btg)
    c = counts.cpu().numpy()
    # same probes through a prefix index
    bits = 10
    lut = np.concatenate([[0], np.cumsum(np.bincount(keys[:, 1] >> (46 - bits), minlength=1 << bits))]).astype(np.int64)
    lut_d = torch.from_numpy(lut).cuda(); idx2 = torch.zeros(len(q), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    capi.check(btg.btg_table_set_index_dev(lut_d.data_ptr(), bits), btg)
    capi.check(btg.btg_table_lookup_dev(kw0.data_ptr(), kw1.data_ptr(), len(keys), qd.data_ptr(), len(q), idx2.data_ptr(), None), btg)
    capi.check(btg.btg_kmer_hash(capi.ptr(np.zeros((1, 2), np.uint64)), 1, capi.ptr(np.zeros(1, np.uint64))), btg)
    capi.check(btg.btg_table_set_index_dev(None, 0), btg)
    assert (idx2.cpu().numpy() == exp).all()
    assert (c[:10, 1] == 255).all() and c[:10, [0, 2]].sum() == 0 and c[10:].sum() == 0
    assert rec.cpu().numpy()[:10].all() and not rec.cpu().numpy()[10:].any()
