"""GPU parity (bit-exact) of the candidate-path search against the REFERENCE's own
findVariantClusterPaths output (fixtures dumped by oracle-R, tests/golden/make_fixtures.py): same graphs, same
sample Bloom filters, same --random-seed -> identical best_paths_indices, row for row."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from bayestyper_b200 import btd, capi, synth
from tests._fixtures import GOLD
from tests.golden.make_fixtures import PATH_WORKLOADS

pytestmark = pytest.mark.gpu
K = 55


class GraphsDesc(C.Structure):
    _fields_ = [("n_clusters", C.c_uint32)] + [(n, C.c_void_p) for n in
                ("cl_vertex_off", "v_seq_off", "seq", "v_flags", "v_in_off", "v_in_src", "cl_group", "cl_idx")]


def _load(name):
    d = btd.read(GOLD / f"{name}.btd")
    g = {k[2:]: v for k, v in d.items() if k.startswith("g.")}
    w = PATH_WORKLOADS[name]()
    spectra = synth.sample_spectra(w, 4, int(d["meta.n_errors"][0]))
    sha = b"".join(hashlib.sha256(k.tobytes() + c.tobytes()).digest() for k, c in spectra)
    assert sha == d["meta.kmer_sha"].tobytes(), "seeded generator no longer reproduces the fixture's k-mer sets"
    return g, spectra, int(d["meta.seed"][0]), int(d["meta.max_hap"][0])


def _desc(g):
    C_ = len(g["cl_vertex_off"]) - 1
    gco = g["group_cluster_off"]
    cl_group = np.repeat(np.arange(len(gco) - 1, dtype=np.uint32), np.diff(gco).astype(np.int64))
    keep = {"cl_vertex_off": np.ascontiguousarray(g["cl_vertex_off"], np.uint64), "v_seq_off": np.ascontiguousarray(g["v_seq_off"], np.uint64),
            "seq": np.ascontiguousarray(g["seq"], np.uint8), "v_flags": np.ascontiguousarray(g["v_flags"], np.uint8),
            "v_in_off": np.ascontiguousarray(g["v_in_off"], np.uint64), "v_in_src": np.ascontiguousarray(g["v_in_src"], np.uint32),
            "cl_group": cl_group, "cl_idx": np.ascontiguousarray(g["cluster_idx"], np.uint32)}
    d = GraphsDesc()
    d.n_clusters = C_
    for k, v in keep.items():
        setattr(d, k, v.ctypes.data)
    return d, keep


@pytest.mark.parametrize("mode", ["per_sample", "batch", "first_then_batch"])
@pytest.mark.parametrize("name", list(PATH_WORKLOADS))
def test_best_paths_identical_to_reference(btg, name, mode):
    """per_sample: one launch per sample (KmerCounter::findVariantClusterPaths as the reference calls it); batch: all samples in one launch
    (btg_find_sample_paths_batch: the (cluster, sample) searches side by side, merged per cluster in sample order); first_then_batch: sample 0
    alone, the others as a batch that starts at sample 1.  All three must leave the reference's best_paths_indices."""
    g, spectra, seed, max_hap = _load(name)
    desc, keep = _desc(g)
    S = len(spectra)
    gr = capi.check(btg.btg_graphs_upload(C.addressof(desc), S, max_hap), btg)
    blooms = []
    for km, _ in spectra:
        b = capi.check(btg.btg_bloom_create(len(km), 1e-3, K), btg)       # makeBloom: KmerBloom(n, 0.001)
        capi.check(btg.btg_bloom_insert(b, capi.ptr(km), len(km)), btg)
        blooms.append(b)
    arr = (C.c_void_p * S)(*blooms)
    if mode == "per_sample":
        for s, b in enumerate(blooms):
            capi.check(btg.btg_find_sample_paths(gr, b, s, seed, max_hap), btg)
    elif mode == "batch":
        capi.check(btg.btg_find_sample_paths_batch(gr, arr, 0, S, seed, max_hap), btg)
    else:
        capi.check(btg.btg_find_sample_paths(gr, blooms[0], 0, seed, max_hap), btg)
        if S > 1:
            rest = (C.c_void_p * (S - 1))(*blooms[1:])
            capi.check(btg.btg_find_sample_paths_batch(gr, rest, 1, S - 1, seed, max_hap), btg)
    Cn = desc.n_clusters
    n_paths = np.zeros(Cn, np.uint32)
    off = np.zeros(Cn + 1, np.uint64)
    capi.check(btg.btg_get_best_paths(gr, capi.ptr(n_paths), capi.ptr(off), None, 0), btg)
    mem = np.zeros(int(off[-1]), np.uint8)
    capi.check(btg.btg_get_best_paths(gr, capi.ptr(n_paths), capi.ptr(off), capi.ptr(mem), mem.size), btg)
    for b in blooms:
        btg.btg_bloom_free(b)
    V = np.diff(g["cl_vertex_off"]).astype(np.int64)
    ref_n = (np.diff(g["cl_path_off"]).astype(np.int64) // V)
    assert (n_paths == ref_n).all(), f"{(n_paths != ref_n).sum()} clusters differ in the number of best paths"
    assert (off == g["cl_path_off"]).all()
    assert (mem == g["path_bits"]).all()
    btg.btg_graphs_free(gr)


def test_batch_with_global_scratch_equals_per_sample(btg, monkeypatch):
    """Working sets that do not fit the shared-memory budget take a per-warp slice of a global arena in the batch kernel (a per-cluster slice
    in the one-sample kernel): the dense 3-sample fixture searched with a pool of 32 haplotypes per sample, both ways, a second pass after
    btg_graphs_reset included (the arena and the turn counters are reused)."""
    g, spectra, seed, _ = _load("paths_mixed_3s")
    desc, keep = _desc(g)
    S = len(spectra)
    blooms = []
    for km, _ in spectra:
        b = capi.check(btg.btg_bloom_create(len(km), 1e-3, K), btg)
        capi.check(btg.btg_bloom_insert(b, capi.ptr(km), len(km)), btg)
        blooms.append(b)
    arr = (C.c_void_p * S)(*blooms)
    got = []
    for mode in ("per_sample", "batch", "batch"):
        if not got:
            gr = capi.check(btg.btg_graphs_upload(C.addressof(desc), S, 32), btg)
        else:
            capi.check(btg.btg_graphs_reset(gr), btg)
        if mode == "per_sample":
            for s, b in enumerate(blooms):
                capi.check(btg.btg_find_sample_paths(gr, b, s, seed, 32), btg)
        else:
            capi.check(btg.btg_find_sample_paths_batch(gr, arr, 0, S, seed, 32), btg)
        n_paths = np.zeros(desc.n_clusters, np.uint32)
        off = np.zeros(desc.n_clusters + 1, np.uint64)
        capi.check(btg.btg_get_best_paths(gr, capi.ptr(n_paths), capi.ptr(off), None, 0), btg)
        mem = np.zeros(int(off[-1]), np.uint8)
        capi.check(btg.btg_get_best_paths(gr, capi.ptr(n_paths), capi.ptr(off), capi.ptr(mem), mem.size), btg)
        got.append((n_paths, mem))
    btg.btg_graphs_free(gr)
    for b in blooms:
        btg.btg_bloom_free(b)
    for n_paths, mem in got[1:]:
        assert (n_paths == got[0][0]).all() and (mem == got[0][1]).all()


def test_rejects_bad_arguments(btg):
    g, spectra, seed, max_hap = _load("paths_snv_1s")
    desc, keep = _desc(g)
    assert btg.btg_graphs_upload(C.addressof(desc), 1, 64) is None            # > 32 haplotypes per sample
    gr = capi.check(btg.btg_graphs_upload(C.addressof(desc), 1, 32), btg)
    km = spectra[0][0]
    b = capi.check(btg.btg_bloom_create(len(km), 1e-3, K), btg)
    assert btg.btg_find_sample_paths(gr, b, 1, seed, 32) < 0                  # sample index beyond capacity
    btg.btg_bloom_free(b)
    btg.btg_graphs_free(gr)
