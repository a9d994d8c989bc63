"""GPU parity of the warp-per-cluster Gibbs path (csrc/gibbs_wide.cu: one warp per cluster or nested group, lane = sample)
against oracle-P through the C ABI — same Philox streams, identical diplotype tallies.

The wide arena layout is chosen at upload for units with >= 8 samples (BASELINE configs[3]: 30 samples, --noise-genotyping);
BTG_WIDE=1 forces it for the small fixtures too, so every stored case (1-3 samples, chrX ploidies, nested groups) also runs
through the lane = sample kernels.  The S = 30 fixtures were produced by the reference itself (tests/golden/make_fixtures.py
exact) and oracle-P reproduces them bit for bit in its mt19937 mode (tests/test_ref_parity_exact.py)."""
import numpy as np
import pytest

from bayestyper_b200 import engine, unit as U
from tests import _oracle as O
from tests._fixtures import GibbsFixture

pytestmark = pytest.mark.gpu
GPP_TOL = 1e-4


def _both(fx, rates=None):
    ocd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    gcd = engine.CountDistribution(fx.nb_p, fx.nb_size)
    if rates is not None:
        ocd.set_noise_rates(rates); gcd.set_noise_rates(rates)
    return ocd, gcd


def _same_results(gres, ores):
    assert np.abs(gres["gpp"] - ores["gpp"]).max() <= GPP_TOL
    assert np.abs(gres["app"] - ores["app"]).max() <= GPP_TOL
    for k in ("gt", "gq", "saf", "an", "ac", "anc", "hc", "ploidy"):
        assert (gres[k] == ores[k]).all(), k
    for k in ("nak", "fak", "mac", "af", "acp"):
        assert np.abs(gres[k] - ores[k]).max() <= 1e-4, k


def _same_tallies(eng, unit, otally):
    toff = unit.tally_offsets()
    bad = [c for c in range(unit.Cn) if not (eng.cluster_tally(c).reshape(-1) == otally[int(toff[c]):int(toff[c + 1])]).all()]
    assert not bad, f"{len(bad)} clusters with different diplotype tallies: {bad[:10]}"


@pytest.mark.parametrize("name", ["gibbs_snv_1s", "gibbs_mixed_3s", "gibbs_chrx_2s", "gibbs_nested_2s", "gibbs_joint_30s"])
def test_default_mode_wide(btg, name, monkeypatch):
    monkeypatch.setenv("BTG_WIDE", "1")
    fx = GibbsFixture(name)
    opts = fx.opts(chains=4, burn=30, samples=60)
    ocd, gcd = _both(fx, fx.tab["noise_rates"])
    ores, otally = O.oracle_estimate_genotypes(fx.unit, ocd, opts, want_tally=True)
    eng = engine.InferenceEngine(fx.unit)
    gres = eng.estimate_genotypes(gcd, opts)
    _same_tallies(eng, fx.unit, otally)
    _same_results(gres, ores)
    eng.close()


@pytest.mark.parametrize("split_cost", ["0", "4000000000"])
def test_chain_split_wide(btg, split_cost, monkeypatch):
    monkeypatch.setenv("BTG_WIDE", "1")
    monkeypatch.setenv("BTG_SPLIT_COST", split_cost)
    fx = GibbsFixture("gibbs_mixed_3s")
    opts = fx.opts(chains=23, burn=5, samples=12)
    ocd, gcd = _both(fx, fx.tab["noise_rates"])
    ores, otally = O.oracle_estimate_genotypes(fx.unit, ocd, opts, want_tally=True)
    eng = engine.InferenceEngine(fx.unit)
    gres = eng.estimate_genotypes(gcd, opts)
    _same_tallies(eng, fx.unit, otally)
    _same_results(gres, ores)
    eng.close()


@pytest.mark.parametrize("name", ["gibbs_snv_1s", "gibbs_chrx_2s", "gibbs_nested_2s", "gibbs_joint_30s"])
def test_estimate_noise_wide(btg, name, monkeypatch):
    monkeypatch.setenv("BTG_WIDE", "1")
    fx = GibbsFixture(name)
    opts = fx.opts(chains=3, burn=20, samples=40)
    ocd, gcd = _both(fx)
    otrace = O.oracle_estimate_noise(fx.unit, ocd, opts)
    eng = engine.InferenceEngine(fx.unit)
    gtrace = eng.estimate_noise(gcd, opts)
    assert gtrace.shape == otrace.shape and (gtrace[:, :2] == otrace[:, :2]).all()
    assert (np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]).max() < 1e-9
    assert np.abs(gcd.noise_rates() / ocd.noise_rates() - 1).max() < 1e-9
    eng.close()


@pytest.mark.parametrize("name,wide", [("gibbs_joint_30s", None), ("gibbs_joint_2s", "1"), ("gibbs_chrx_2s", "1"),
                                       ("gibbs_joint_nested_2s", "1"), ("gibbs_joint_nested_2s", "0"), ("gibbs_nested_2s", "0")])
def test_joint_mode_wide(btg, name, wide, monkeypatch):
    """--noise-genotyping (InferenceEngine::estimateNoiseAndGenotypes, InferenceEngine.cpp:384-472) with 30 samples, and with nested
    groups (which take the warp-per-group kernel on either arena layout)."""
    if wide is not None:
        monkeypatch.setenv("BTG_WIDE", wide)
    fx = GibbsFixture(name)
    opts = fx.opts(chains=3, burn=20, samples=40)
    ocd, gcd = _both(fx)
    ores, otrace = O.oracle_estimate_noise_and_genotypes(fx.unit, ocd, opts)
    eng = engine.InferenceEngine(fx.unit)
    gres, gtrace = eng.estimate_noise_and_genotypes(gcd, opts)
    assert gtrace.shape == otrace.shape and (gtrace[:, :2] == otrace[:, :2]).all()
    assert (np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]).max() < 1e-9
    _same_results(gres, ores)
    eng.close()


def test_joint_30_samples_full_length_matches_reference_calls(btg):
    """The S = 30 fixture at the reference's own chain lengths (20 x (100 + 250)): the GPU equals oracle-P (Philox), and both
    agree with the calls the reference printed for this run to within Monte-Carlo error of independent streams."""
    fx = GibbsFixture("gibbs_joint_30s")
    opts = fx.opts()
    ocd, gcd = _both(fx)
    ores, otrace = O.oracle_estimate_noise_and_genotypes(fx.unit, ocd, opts)
    eng = engine.InferenceEngine(fx.unit)
    gres, gtrace = eng.estimate_noise_and_genotypes(gcd, opts)
    assert (np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]).max() < 1e-9
    _same_results(gres, ores)
    S = fx.S
    same = (gres["gt"].reshape(-1, S, 2) == fx.ref["gt"].reshape(-1, S, 2)).all(axis=2)
    assert same.mean() > 0.97
    assert np.abs(gres["gpp"] - fx.ref["gpp"]).mean() < 4e-3
    eng.close()


def test_clusters_without_dense_caches(btg, monkeypatch):
    """Wide layout, clusters whose S x (H+1)(H+2)/2 cache would exceed the cap keep no cache and find the outcome by a second walk
    of the enumeration: results must not depend on it (12 samples, up to 9+ haplotype candidates; the cap is lowered through
    BTG_WIDE_CACHE_CAP so that most clusters of this unit take the uncached path)."""
    from bayestyper_b200 import synth, synth_unit
    monkeypatch.setenv("BTG_WIDE_CACHE_CAP", "64")
    w = synth.small_mixed(200, 16000, 12, 78, 0.15)
    unit = synth_unit.build_unit(w, seed=5, max_cluster_variants=5)
    nb_p, nb_size = [0.6] * unit.S, [22.5] * unit.S
    opts = U.default_opts(min_frac=U.min_fraction_observed(nb_p, nb_size), chains=2, burn=5, samples=10)
    ocd = O.OracleCountDist(nb_p, nb_size); gcd = engine.CountDistribution(nb_p, nb_size)
    eng = engine.InferenceEngine(unit)
    jres_o, jtrace_o = O.oracle_estimate_noise_and_genotypes(unit, ocd, opts)
    jres_g, jtrace_g = eng.estimate_noise_and_genotypes(gcd, opts)
    assert (np.abs(jtrace_g[:, 2:] - jtrace_o[:, 2:]) / jtrace_o[:, 2:]).max() < 1e-9
    _same_results(jres_g, jres_o)
    ocd.set_noise_rates([0.01] * unit.S); gcd.set_noise_rates([0.01] * unit.S)
    ores, otally = O.oracle_estimate_genotypes(unit, ocd, opts, want_tally=True)
    gres = eng.estimate_genotypes(gcd, opts)
    _same_tallies(eng, unit, otally)
    _same_results(gres, ores)
    eng.close()
