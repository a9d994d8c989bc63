"""The genotype-side k-mer stages through the btg_counter handle of the C ABI (csrc/counter.cu: device sorts / run-length encodings /
scans between the kernels, no torch) against the reference's own VariantClusterHaplotypes dumped by oracle-R — the same fixtures and the same
checks as tests/test_gpu_pipeline.py runs on the torch-glue mirror: row order (k-mer by k-mer), multiplicity matrix, counts, inter-cluster
multiplicities, unique / multicluster lists, haplotype -> allele tables, coverage bitmaps (as sets), nested tables, shared records."""
import numpy as np
import pytest
import torch

from bayestyper_b200 import capi, counter, engine, unit as U
from tests import _oracle as O
from tests.golden.make_fixtures import PIPE_WORKLOADS
from tests.test_gpu_pipeline import K, _load, _vh_sets

pytestmark = pytest.mark.gpu


def _run(btg, name):
    d, g, h, w, spectra = _load(name)
    S = len(spectra)
    V = np.diff(g["cl_vertex_off"]).astype(np.int64)
    n_paths = np.diff(g["cl_path_off"]).astype(np.int64) // V
    kc = counter.KmerCounter(g, n_paths, g["path_bits"], S, w.genders)
    n_keys = kc.count_path_kmers()
    male_ploidy = 1 if w.chrom.lower() in ("x", "chrx") else 2
    seq = np.frombuffer(w.reference, np.uint8)
    parts = []
    for dec, a, b in d["regions"]:
        if dec == 0:
            parts += [seq[int(a):int(b) + 1], np.frombuffer(b"N", np.uint8)]
    buf = torch.from_numpy(np.concatenate(parts)).cuda() if parts else torch.zeros(0, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    kc.count_intercluster_kmers(buf.data_ptr(), buf.numel(), 2, male_ploidy, False)
    keep = []
    for s, (km, ct) in enumerate(spectra):
        kd = torch.from_numpy(km.view(np.int64)).cuda(); cdv = torch.from_numpy(ct).cuda()
        torch.cuda.synchronize()
        kc.parse_sample_kmers(s, kd.data_ptr(), cdv.data_ptr(), cdv.numel())
        keep.append((kd, cdv))
    mg = capi.check(btg.btg_bloom_from_bytes(capi.ptr(d["mg.data"]), int(d["mg.meta"][0]), int(d["mg.meta"][1]), K), btg)
    handle = kc.build_unit(multigroup_bloom=mg)
    btg.btg_bloom_free(mg)
    return d, g, h, w, kc, handle, n_keys


@pytest.mark.parametrize("name", list(PIPE_WORKLOADS))
def test_counter_handle_identical_to_reference(btg, name):
    d, g, h, w, kc, handle, n_keys = _run(btg, name)
    u = kc.unit_arrays()
    a = u.a
    assert (a["cl_nhap"] == np.diff(h["cl_hap_off"])).all()
    assert (a["cl_kmer_off"] == h["cl_kmer_off"]).all(), "row counts per cluster differ"
    for k in ("mult", "k_has_counts", "k_counts", "k_ic", "cl_uniq_off", "uniq_idx", "cl_multi_off", "multi_idx", "hap_alleles", "cl_mult_off", "cl_hapvar_off"):
        assert (a[k] == h[k]).all(), k
    rows = len(a["k_has_counts"])
    assert (a["kmer_vh_off"] == h["kmer_vh_off"]).all()
    assert _vh_sets(a, rows) == _vh_sets(h, rows)
    for k in ("hap_nested_off", "hap_nested", "cl_dep_off", "dep_cluster", "dep_var_off", "dep_var"):
        assert (a[k] == h[k]).all(), k
    fl = h["k_flags"]
    sh = a["k_shared"] != 0xFFFFFFFF
    assert (sh == ((fl & 2) != 0)).all()
    # the table: distinct path k-mers in lexicographic order
    lo, hi = kc.array("key_lo", np.int64), kc.array("key_hi", np.int64)
    assert len(lo) == n_keys and (np.lexsort((lo, hi)) == np.arange(n_keys)).all()
    btg.btg_unit_free(handle)
    kc.close()


def test_unit_from_the_counter_runs_the_sampler(btg):
    """The handle's unit goes straight into the Gibbs stage (no host copy of the row-level arrays): same tallies as oracle-P on the
    descriptor the handle reports."""
    d, g, h, w, kc, handle, _ = _run(btg, "pipe_mixed_3s")
    u = kc.unit_arrays()
    nb_p, nb_size = d["t.nb_p_size"][:, 0].copy(), d["t.nb_p_size"][:, 1].copy()
    opts = U.default_opts(min_frac=U.min_fraction_observed(nb_p, nb_size), chains=3, burn=10, samples=20)
    rates = [0.01] * u.S
    ocd = O.OracleCountDist(nb_p, nb_size); ocd.set_noise_rates(rates)
    gcd = engine.CountDistribution(nb_p, nb_size); gcd.set_noise_rates(rates)
    ores = O.oracle_estimate_genotypes(u, ocd, opts)
    eng = engine.InferenceEngine.from_handle(u, handle)
    gres = eng.estimate_genotypes(gcd, opts)
    assert np.abs(gres["gpp"] - ores["gpp"]).max() <= 1e-4
    for k in ("gt", "gq", "saf"):
        assert (gres[k] == ores[k]).all(), k
    eng.close(); kc.close()
