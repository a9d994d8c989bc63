"""GPU parity of the Gibbs path against oracle-P through the C ABI:
same Philox streams -> identical diplotype tallies, GPP/APP within 1e-4 (north-star tolerance),
identical GT/GQ/SAF; count tables within 5e-11 relative (f64 libm differences only)."""
import numpy as np
import pytest

from bayestyper_b200 import engine, unit as U
from tests import _oracle as O
from tests._fixtures import GIBBS_FIXTURES, GibbsFixture

pytestmark = pytest.mark.gpu
GPP_TOL = 1e-4   # BASELINE.json north_star: "posteriors within 1e-4 of reference under fixed seed"


def _both(fx, opts, noise_rates=None):
    rates = fx.tab["noise_rates"] if noise_rates is None else noise_rates
    ocd = O.OracleCountDist(fx.nb_p, fx.nb_size); ocd.set_noise_rates(rates)
    gcd = engine.CountDistribution(fx.nb_p, fx.nb_size); gcd.set_noise_rates(rates)
    return ocd, gcd


@pytest.mark.parametrize("name", GIBBS_FIXTURES)
def test_count_tables(btg, name):
    fx = GibbsFixture(name)
    ocd, gcd = _both(fx, None)
    og, on = ocd.tables()
    gg, gn = gcd.tables()
    assert (np.isfinite(og) == np.isfinite(gg)).all()
    fin = np.isfinite(og)
    # f64 lgamma(obs + size*m) - lgamma(size*m) cancels ~1e4-sized terms: device and glibc libm agree to a few ulp of those
    rel = np.abs(gg[fin] - og[fin]) / np.maximum(1.0, np.abs(og[fin]))
    assert rel.max() < 5e-11
    assert (np.abs(gn - on) / np.maximum(1.0, np.abs(on))).max() < 5e-11
    # and against the reference's own tables (fixture)
    ref = fx.tab["genomic_log_pmf"]
    assert (np.abs(gg[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))).max() < 5e-11


@pytest.mark.parametrize("name", GIBBS_FIXTURES)
def test_estimate_genotypes_matches_oracle(btg, name):
    fx = GibbsFixture(name)
    opts = fx.opts()
    ocd, gcd = _both(fx, opts)
    ores, otally = O.oracle_estimate_genotypes(fx.unit, ocd, opts, want_tally=True)
    eng = engine.InferenceEngine(fx.unit)
    gres = eng.estimate_genotypes(gcd, opts)
    toff = fx.unit.tally_offsets()
    n_bad = 0
    for c in range(fx.unit.Cn):
        t = eng.cluster_tally(c).reshape(-1)
        n_bad += not (t == otally[int(toff[c]):int(toff[c + 1])]).all()
    assert n_bad == 0, f"{n_bad} clusters with different diplotype tallies"
    assert np.abs(gres["gpp"] - ores["gpp"]).max() <= GPP_TOL
    assert np.abs(gres["app"] - ores["app"]).max() <= GPP_TOL
    for k in ("gt", "gq", "saf", "an", "ac", "anc", "hc", "ploidy"):
        assert (gres[k] == ores[k]).all(), k
    for k in ("nak", "fak", "mac", "af", "acp"):
        assert np.abs(gres[k] - ores[k]).max() <= 1e-4, k
    # and statistically against the reference's VCF values held by the fixture
    d = np.abs(gres["gpp"] - fx.ref["gpp"])
    assert d.mean() < 2e-3 and np.quantile(d, 0.99) <= 0.1 + 1e-6
    eng.close()


@pytest.mark.parametrize("name", ["gibbs_mixed_3s", "gibbs_chrx_2s"])
@pytest.mark.parametrize("split_cost", ["0", "4000000000"])
def test_chain_split_is_invisible(btg, name, split_cost, monkeypatch):
    """Default mode: the chains of a cluster are independent, so a cluster whose chains run on 20 threads with private state
    (every cluster with BTG_SPLIT_COST=0, none with a huge threshold) gives the tallies of the sequential oracle."""
    monkeypatch.setenv("BTG_SPLIT_COST", split_cost)
    fx = GibbsFixture(name)
    opts = fx.opts(chains=23, burn=5, samples=12)          # more chains than virtual threads: thread v runs chains v and v + 20
    ocd, gcd = _both(fx, opts)
    ores, otally = O.oracle_estimate_genotypes(fx.unit, ocd, opts, want_tally=True)
    eng = engine.InferenceEngine(fx.unit)
    gres = eng.estimate_genotypes(gcd, opts)
    toff = fx.unit.tally_offsets()
    for c in range(fx.unit.Cn):
        assert (eng.cluster_tally(c).reshape(-1) == otally[int(toff[c]):int(toff[c + 1])]).all(), c
    assert np.abs(gres["gpp"] - ores["gpp"]).max() <= GPP_TOL
    for k in ("gt", "gq", "saf", "an", "ac"):
        assert (gres[k] == ores[k]).all(), k
    for k in ("nak", "fak", "mac"):
        assert np.abs(gres[k] - ores[k]).max() <= 1e-4, k
    eng.close()


def test_seed_and_shard_invariance(btg):
    """Per-group streams: a group's result does not depend on which other groups share the launch
    (the reference's determinism contract, README.md:9), and changes with the seed."""
    fx = GibbsFixture("gibbs_mixed_3s")
    opts = fx.opts(chains=3, burn=20, samples=40)
    _, gcd = _both(fx, opts)
    eng = engine.InferenceEngine(fx.unit)
    full = eng.estimate_genotypes(gcd, opts)
    half = np.arange(fx.unit.G // 2, fx.unit.G)
    sub = fx.unit.subset_groups(half)
    eng2 = engine.InferenceEngine(sub)
    o2 = fx.opts(chains=3, burn=20, samples=40, group_base=int(half[0]))
    part = eng2.estimate_genotypes(gcd, o2)
    v0 = int(fx.unit.a["cl_var_off"][int(fx.unit.a["group_cluster_off"][half[0]])])
    g0 = int(full["geno_off"][v0])
    assert (part["gpp"] == full["gpp"][g0:]).all()
    assert (part["gt"] == full["gt"][v0 * fx.S * 2:]).all()
    o3 = fx.opts(chains=3, burn=20, samples=40, seed=7)
    other = eng.estimate_genotypes(gcd, o3)
    assert (other["nak"] != full["nak"]).any()      # different k-mer subsamples
    eng.close(); eng2.close()


def test_estimate_noise_matches_oracle(btg):
    fx = GibbsFixture("gibbs_snv_1s")
    opts = fx.opts(chains=3, burn=30, samples=60)
    ocd, gcd = _both(fx, opts)
    otrace = O.oracle_estimate_noise(fx.unit, ocd, opts)
    eng = engine.InferenceEngine(fx.unit)
    gtrace = eng.estimate_noise(gcd, opts)
    assert gtrace.shape == otrace.shape
    assert (gtrace[:, :2] == otrace[:, :2]).all()
    rel = np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]
    assert rel.max() < 1e-9, (rel.max(), int(rel.argmax()), gtrace[int(rel.argmax()) // fx.S], otrace[int(rel.argmax()) // fx.S])
    assert np.abs(gcd.noise_rates() - ocd.noise_rates()).max() / ocd.noise_rates().max() < 1e-9
    eng.close()


@pytest.mark.parametrize("name,parts", [("gibbs_snv_1s", 2), ("gibbs_chrx_2s", 3), ("gibbs_mixed_3s", 8)])
def test_estimate_noise_by_chains_equals_the_whole_run(btg, name, parts):
    """btg_estimate_noise_chains: the chains of estimateNoise dealt to `parts` ranks (here one after the other on one GPU), their per-chain
    sums added up and finished by btg_count_dist_finish_noise = btg_estimate_noise, bit for bit (rates and every trace row)."""
    fx = GibbsFixture(name)
    opts = fx.opts(chains=5, burn=10, samples=25)
    _, whole_cd = _both(fx, opts)
    eng = engine.InferenceEngine(fx.unit)
    want = eng.estimate_noise(whole_cd, opts)
    _, cd = _both(fx, opts)
    total = np.zeros((opts.n_chains, fx.S))
    trace = np.zeros_like(want)
    for r in range(parts):
        sums, tr = eng.estimate_noise_chains(cd, opts, r, parts, want_trace=True)
        mine = np.arange(opts.n_chains) % parts == r
        assert (sums[~mine] == 0).all() and (sums[mine] > 0).all()
        total += sums
        trace += tr
    cd.finish_noise(total, opts.gibbs_samples)
    assert (cd.noise_rates() == whole_cd.noise_rates()).all()
    assert (trace[:-1] == want[:-1]).all()                      # the last row ("0 0", final rates) is written by the finish of a whole run only
    assert (want[-1, 2:] == cd.noise_rates()).all()
    g1, n1 = cd.tables(); g2, n2 = whole_cd.tables()
    assert (n1 == n2).all()
    eng.close()


def test_multi_sample_noise(btg):
    fx = GibbsFixture("gibbs_chrx_2s")
    opts = fx.opts(chains=2, burn=10, samples=20)
    ocd, gcd = _both(fx, opts)
    otrace = O.oracle_estimate_noise(fx.unit, ocd, opts)
    eng = engine.InferenceEngine(fx.unit)
    gtrace = eng.estimate_noise(gcd, opts)
    assert (np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]).max() < 1e-9
    eng.close()


@pytest.mark.parametrize("name", ["gibbs_snv_1s", "gibbs_chrx_2s"])
def test_noise_genotyping_joint_mode_matches_oracle(btg, name):
    """--noise-genotyping: InferenceEngine::estimateNoiseAndGenotypes (all groups in lock-step)."""
    fx = GibbsFixture(name)
    opts = fx.opts(chains=3, burn=20, samples=40)
    ocd, gcd = _both(fx, opts)
    ores, otrace = O.oracle_estimate_noise_and_genotypes(fx.unit, ocd, opts)
    eng = engine.InferenceEngine(fx.unit)
    gres, gtrace = eng.estimate_noise_and_genotypes(gcd, opts)
    assert gtrace.shape == otrace.shape and (gtrace[:, :2] == otrace[:, :2]).all()
    assert (np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]).max() < 1e-9
    assert np.abs(gres["gpp"] - ores["gpp"]).max() <= GPP_TOL
    for k in ("gt", "gq", "saf", "an", "ac"):
        assert (gres[k] == ores[k]).all(), k
    for k in ("nak", "fak", "mac", "app", "acp"):
        assert np.abs(gres[k] - ores[k]).max() <= 1e-4, k
    eng.close()


def test_malformed_group_forest_rejected_loudly(btg):
    """A multi-cluster group needs sources + edges forming a forest over its clusters (VariantClusterGroup.cpp:47-107)."""
    fx = GibbsFixture("gibbs_snv_1s")
    a = dict(fx.unit.a)
    gco = a["group_cluster_off"].copy()
    a["group_cluster_off"] = np.concatenate([gco[:1], gco[2:]])      # merge the first two groups, drop one source
    a["group_ploidy"] = a["group_ploidy"][fx.S:]
    a["group_src_off"] = np.concatenate([a["group_src_off"][:1], a["group_src_off"][2:] - 1])
    a["group_src"] = a["group_src"][1:]
    a["group_edge_off"] = a["group_edge_off"][:-1]
    bad = U.Unit(a, fx.S)
    with pytest.raises(Exception, match="forest"):
        engine.InferenceEngine(bad)


def test_nested_unit_noise_estimation_uses_single_cluster_groups_only(btg):
    """InferenceEngine::estimateNoise draws from groups of ONE cluster (InferenceEngine.cpp:144-151)."""
    fx = GibbsFixture("gibbs_nested_2s")
    opts = fx.opts(chains=2, burn=20, samples=30)
    ocd, gcd = _both(fx, opts)
    otrace = O.oracle_estimate_noise(fx.unit, ocd, opts)
    eng = engine.InferenceEngine(fx.unit)
    gtrace = eng.estimate_noise(gcd, opts)
    assert np.abs(gtrace - otrace).max() <= 1e-9 * max(1.0, np.abs(otrace).max())
    eng.close()


def test_empty_unit(btg):
    fx = GibbsFixture("gibbs_snv_1s")
    empty = fx.unit.subset_groups(np.zeros(0, np.int64))
    eng = engine.InferenceEngine(empty)
    _, gcd = _both(fx, None)
    res = eng.estimate_genotypes(gcd, fx.opts())
    assert res["gpp"].size == 0
    eng.close()


def test_large_clusters_take_the_cooperative_paths_and_match_oracle(btg):
    """A synthetic unit dense enough for clusters with up to 9 haplotype candidates and 600 k-mers (6 samples): in the lock-step
    chain these take the warp paths (dense k-mer tiles, grid-wide fill tasks, term-parallel fills for >= 16 k-mers per entry,
    warp-cooperative construct), in the default mode the chain split.  Noise trace and diplotype tallies must equal the
    sequential CPU restatement."""
    from bayestyper_b200 import synth, synth_unit
    w = synth.small_mixed(260, 20000, 6, 77, 0.15)
    unit = synth_unit.build_unit(w, seed=3, max_cluster_variants=6)
    H = unit.a["cl_nhap"].astype(np.int64)
    nu = np.diff(unit.a["cl_uniq_off"].astype(np.int64))
    per_sample_cost = (H * (H + 1) // 2) * (nu // 10 + 1)
    assert ((per_sample_cost > 128) & (nu // 10 >= 16)).sum() >= 20 and H.max() >= 8, "workload no longer exercises the large-cluster paths"
    nb_p, nb_size = [0.6] * unit.S, [22.5] * unit.S
    opts = U.default_opts(min_frac=U.min_fraction_observed(nb_p, nb_size), chains=3, burn=5, samples=10)
    ocd = O.OracleCountDist(nb_p, nb_size)
    gcd = engine.CountDistribution(nb_p, nb_size)
    eng = engine.InferenceEngine(unit)
    otrace = O.oracle_estimate_noise(unit, ocd, opts)
    gtrace = eng.estimate_noise(gcd, opts)
    assert (gtrace[:, :2] == otrace[:, :2]).all()
    assert (np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]).max() < 1e-9
    gcd.set_noise_rates(ocd.noise_rates())
    ores, otally = O.oracle_estimate_genotypes(unit, ocd, opts, want_tally=True)
    gres = eng.estimate_genotypes(gcd, opts)
    toff = unit.tally_offsets()
    bad = [c for c in range(unit.Cn) if not (eng.cluster_tally(c).reshape(-1) == otally[int(toff[c]):int(toff[c + 1])]).all()]
    assert not bad, f"clusters with different tallies: {bad[:10]}"
    assert np.abs(gres["gpp"] - ores["gpp"]).max() <= GPP_TOL and (gres["gt"] == ores["gt"]).all()
    # the joint mode on the same unit (all groups have one cluster)
    ocd2 = O.OracleCountDist(nb_p, nb_size); gcd2 = engine.CountDistribution(nb_p, nb_size)
    jres_o, jtrace_o = O.oracle_estimate_noise_and_genotypes(unit, ocd2, opts)
    jres_g, jtrace_g = eng.estimate_noise_and_genotypes(gcd2, opts)
    assert (np.abs(jtrace_g[:, 2:] - jtrace_o[:, 2:]) / jtrace_o[:, 2:]).max() < 1e-9
    assert np.abs(jres_g["gpp"] - jres_o["gpp"]).max() <= GPP_TOL and (jres_g["gt"] == jres_o["gt"]).all()
    eng.close(); gcd.close(); gcd2.close()
