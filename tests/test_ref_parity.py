"""Pins oracle-P (oracle/gibbs_oracle.cpp) to the REFERENCE's behaviour, using fixtures that the
reference's own code produced (tests/golden/make_fixtures.py -> oracle-R):

  * log-pmf tables: bit-for-bit the reference's CountDistribution caches;
  * posteriors: the reference draws from mt19937/libstdc++, oracle-P from Philox (SURVEY.md §7 hard
    part 1), so agreement is statistical: identical hard calls on confidently called sites, GPP
    within chain-level Monte-Carlo error, NAK/MAC within subsampling error.
"""
import numpy as np
import pytest

from tests import _oracle as O
from tests._fixtures import GIBBS_FIXTURES, GibbsFixture


@pytest.mark.parametrize("name", GIBBS_FIXTURES)
def test_count_tables_bit_exact_vs_reference(name):
    fx = GibbsFixture(name)
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    cd.set_noise_rates(fx.tab["noise_rates"])
    g, n = cd.tables()
    ref_g, ref_n = fx.tab["genomic_log_pmf"], fx.tab["noise_log_pmf"]
    assert (np.isfinite(g) == np.isfinite(ref_g)).all()
    fin = np.isfinite(ref_g)
    assert (g[fin] == ref_g[fin]).all()
    assert (n == ref_n).all()


@pytest.mark.parametrize("name", GIBBS_FIXTURES)
def test_posteriors_statistically_equal_to_reference(name):
    fx = GibbsFixture(name)
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    cd.set_noise_rates(fx.tab["noise_rates"])
    # per-group seeds follow the group's index in the full unit: run group by group is not needed for a
    # statistical comparison, the fixture's local indices give equally valid independent streams
    res = O.oracle_estimate_genotypes(fx.unit, cd, fx.opts())
    S = fx.S
    ref = fx.ref
    d_gpp = np.abs(res["gpp"] - ref["gpp"])
    # 20 chains x 250 samples: one chain landing elsewhere moves GPP by 0.05
    assert d_gpp.mean() < 2e-3
    assert np.quantile(d_gpp, 0.99) <= 0.1 + 1e-6
    # hard calls: identical wherever both sides are far from the 0.99 threshold
    gt_o, gt_r = res["gt"].reshape(-1, S, 2), ref["gt"].reshape(-1, S, 2)
    same = (gt_o == gt_r).all(axis=2)
    assert same.mean() > 0.985
    # disagreements must be threshold cases (one side uncalled) rather than different alleles
    diff = ~same
    called_both = (gt_o[..., 0] != 0xFFFF) & (gt_r[..., 0] != 0xFFFF)
    assert (diff & called_both).sum() <= max(1, int(0.002 * same.size))
    # k-mer statistics: NAK is the mean subsampled k-mer count, MAC the mean count per copy
    m = (ref["nak"] >= 0) & (res["nak"] >= 0)
    assert ((ref["nak"] >= 0) == (res["nak"] >= 0)).mean() > 0.99
    assert np.abs(res["nak"][m] - ref["nak"][m]).mean() < 1.5
    mm = (ref["mac"] >= 0) & (res["mac"] >= 0)
    assert np.abs(res["mac"][mm] - ref["mac"][mm]).mean() < 1.0


def test_noise_estimation_matches_reference_trace_statistically():
    fx = GibbsFixture("gibbs_snv_1s")
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    opts = fx.opts(chains=4, burn=50, samples=100)
    trace = O.oracle_estimate_noise(fx.unit, cd, opts)
    assert trace.shape == (4 * 151 + 1, 3)
    assert (trace[-1, :2] == 0).all()
    post = trace[(trace[:, 1] > 50)][:, 2]
    # the fixture's groups are a subset of the reference run's, so compare the scale only
    ref_post = fx.noise_trace[fx.noise_trace[:, 1] > 100][:, 2]
    assert 0.2 < post.mean() / ref_post.mean() < 5
    assert abs(cd.noise_rates()[0] - post.mean()) < 1e-12


def test_nb_moments():
    import ctypes as C
    L = O.load(); O._bind_gibbs(L)
    p, size = C.c_double(), C.c_double()
    L.bto_nb_moments_to_parameters(15.0, 25.0, 1, C.byref(p), C.byref(size))
    assert abs(p.value - 0.6) < 1e-15 and abs(size.value - 22.5) < 1e-12    # NB(15,25) -> p=0.6 size=22.5
    fx = GibbsFixture("gibbs_snv_1s")
    cd = O.OracleCountDist([0.6], [22.5])
    g, _ = cd.tables()
    for (c, m), want in {(0, 1): -11.493576534734791, (15, 1): -2.5354124696085716, (30, 2): -2.8784684276944397}.items():
        assert abs(g[0, m, c] - want) < 1e-12                                   # SURVEY §4 logPmf vectors
    L.bto_nb_moments_to_parameters(10.0, 5.0, 2, C.byref(p), C.byref(size))   # var < mean: p capped at 0.99
    assert abs(p.value - 0.99) < 1e-15


def test_joint_noise_genotyping_statistically_equal_to_reference():
    """--noise-genotyping (InferenceEngine::estimateNoiseAndGenotypes): the fixture holds every group of the reference run, so
    the noise-rate trace is comparable in distribution, not only the calls."""
    fx = GibbsFixture("gibbs_joint_2s")
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    res, trace = O.oracle_estimate_noise_and_genotypes(fx.unit, cd, fx.opts())
    S = fx.S
    gt_o, gt_r = res["gt"].reshape(-1, S, 2), fx.ref["gt"].reshape(-1, S, 2)
    same = (gt_o == gt_r).all(axis=2)
    assert same.mean() > 0.97
    called_both = (gt_o[..., 0] != 0xFFFF) & (gt_r[..., 0] != 0xFFFF)
    assert (~same & called_both).sum() <= max(1, int(0.004 * same.size))
    assert np.abs(res["gpp"] - fx.ref["gpp"]).mean() < 4e-3
    ref_tr = fx.noise_trace
    assert trace.shape == ref_tr.shape
    post_o = trace[trace[:, 1] > 100][:, 2:]
    post_r = ref_tr[ref_tr[:, 1] > 100][:, 2:]
    # posterior means of the per-sample noise rates agree within a few percent (7000 draws each)
    assert np.abs(post_o.mean(0) / post_r.mean(0) - 1).max() < 0.1
