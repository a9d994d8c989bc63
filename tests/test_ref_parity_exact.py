"""Pins oracle-P (oracle/gibbs_oracle.cpp) to the REFERENCE bit for bit.

oracle-P has two draw sources that share every other line (bto_set_rng_mode): the kernels' Philox streams, and — here —
std::mt19937 + the libstdc++ distributions with the reference's seeds, running-stream semantics and unordered-container walks
(SURVEY.md appendix C).  In that mode it must reproduce what the reference itself computed for the committed fixtures
(tests/golden/make_fixtures.py -> oracle-R): the same diplotype tallies, hence GPP / APP equal to the last printed digit, the
same GT / GQ / SAF, and the same noise-rate trace.  Together with tests/test_gpu_gibbs.py (GPU == oracle-P in Philox mode,
identical tallies) this closes the parity chain without a statistical link.

The reference prints floats with 6 significant digits (GenotypeWriter.cpp:291-312, InferenceEngine.cpp:205,229): posteriors are
multiples of 1/5000 and survive exactly; NAK / MAC / noise rates are compared to print precision.
"""
import numpy as np
import pytest

from tests import _oracle as O
from tests._fixtures import GIBBS_FIXTURES, GibbsFixture

PRINT_REL = 6e-6   # half a unit in the 6th significant digit


def _print_equal(got, ref):
    m = np.abs(ref) > 0
    assert (np.abs(got[m] - ref[m]) / np.abs(ref[m])).max(initial=0) <= PRINT_REL
    assert np.abs(got[~m]).max(initial=0) <= 5e-7


@pytest.mark.parametrize("name", GIBBS_FIXTURES)
def test_default_mode_equals_reference(name):
    """InferenceEngine::estimateGenotypes: the fixture's groups keep their index in the reference's unit (seed derivation)."""
    fx = GibbsFixture(name)
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    cd.set_noise_rates(fx.tab["noise_rates"])
    with O.reference_streams(fx.groups):
        res = O.oracle_estimate_genotypes(fx.unit, cd, fx.opts())
    assert (res["gpp"] == fx.ref["gpp"]).all()
    assert (res["app"] == fx.ref["app"]).all()
    for k in ("gt", "gq", "saf"):
        assert (res[k] == fx.ref[k]).all(), k
    assert (res["fak"] == fx.ref["fak"]).all()
    _print_equal(res["nak"], fx.ref["nak"])
    _print_equal(res["mac"], fx.ref["mac"])


@pytest.mark.parametrize("name", ["gibbs_joint_2s", "gibbs_joint_30s", "gibbs_joint_nested_2s", "gibbs_joint_deep_2s"])
def test_joint_mode_equals_reference(name):
    """--noise-genotyping (InferenceEngine::estimateNoiseAndGenotypes): the fixture holds every group of the reference run, so
    the lock-step noise-rate trace is reproduced row for row as well."""
    fx = GibbsFixture(name)
    assert fx.unit.G == len(fx.groups) and (fx.groups == np.arange(fx.unit.G)).all()
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    with O.reference_streams(fx.groups):
        res, trace = O.oracle_estimate_noise_and_genotypes(fx.unit, cd, fx.opts())
    assert trace.shape == fx.noise_trace.shape and (trace[:, :2] == fx.noise_trace[:, :2]).all()
    _print_equal(trace[:, 2:], fx.noise_trace[:, 2:])
    assert (res["gpp"] == fx.ref["gpp"]).all() and (res["app"] == fx.ref["app"]).all()
    for k in ("gt", "gq", "saf"):
        assert (res[k] == fx.ref[k]).all(), k
    _print_equal(res["nak"], fx.ref["nak"])
    _print_equal(res["mac"], fx.ref["mac"])


@pytest.mark.parametrize("name", ["gibbs_full_2s", "gibbs_deep_2s"])
def test_noise_estimation_equals_reference(name):
    """InferenceEngine::estimateNoise on a fixture that holds the reference's whole unit: group selection (the engine's own
    mt19937 + std::shuffle), per-chain genotyper seeds and the running CountDistribution stream give the reference's trace and
    final rates.  gibbs_deep_2s: a unit whose multi-cluster groups are deeply nested (they are skipped by the selection, and
    genotyped exactly afterwards)."""
    fx = GibbsFixture(name)
    assert (fx.groups == np.arange(fx.unit.G)).all()
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    with O.reference_streams(fx.groups):
        trace = O.oracle_estimate_noise(fx.unit, cd, fx.opts())
        rates = cd.noise_rates()
        res = O.oracle_estimate_genotypes(fx.unit, cd, fx.opts())
    assert trace.shape == fx.noise_trace.shape and (trace[:, :2] == fx.noise_trace[:, :2]).all()
    _print_equal(trace[:, 2:], fx.noise_trace[:, 2:])
    assert np.abs(rates / fx.tab["noise_rates"] - 1).max() < 1e-12      # tables.btd holds the rates at full precision
    assert (res["gpp"] == fx.ref["gpp"]).all()
    for k in ("gt", "gq", "saf"):
        assert (res[k] == fx.ref[k]).all(), k


def test_modes_do_not_leak():
    """The draw source is process state of the checker: leaving the context restores the Philox streams."""
    fx = GibbsFixture("gibbs_snv_1s")
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    cd.set_noise_rates(fx.tab["noise_rates"])
    opts = fx.opts(chains=2, burn=5, samples=10)
    a = O.oracle_estimate_genotypes(fx.unit, cd, opts)
    with O.reference_streams(fx.groups):
        b = O.oracle_estimate_genotypes(fx.unit, cd, opts)
    c = O.oracle_estimate_genotypes(fx.unit, cd, opts)
    assert (a["gpp"] == c["gpp"]).all() and (a["nak"] == c["nak"]).all()
    assert (a["nak"] != b["nak"]).any()
