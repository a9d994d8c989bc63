"""The negative-binomial fit from the parameter k-mers of the cluster stage: given the
<out>_cluster_data/parameter_kmers.fa.gz the REFERENCE's `cluster` wrote (tests/golden/nbfit_mixed_2s.btd, make_fixtures.py nbfit),
the genotype-side fit — genome scan for the inter-cluster multiplicities, sample k-mer stream for the counts, modal multiplicity,
moments -> (p, size) — must give the reference's parameters (KmerHash.cpp:257-347, CountDistribution.cpp:66-141).  The reference
accumulates the moments with Welford updates in hash order, this path with two passes in f64: equal to rounding (1e-12)."""
from pathlib import Path

import numpy as np
import pytest

from bayestyper_b200 import btd, driver, synth
from tests.golden.make_fixtures import NBFIT_WORKLOADS

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def test_nb_fit_from_the_reference_parameter_kmers(btg):
    d = btd.read(GOLD / "nbfit_mixed_2s.btd")
    w = NBFIT_WORKLOADS["nbfit_mixed_2s"]()
    spectra = synth.sample_spectra(w, 4, int(d["meta.n_errors"][0]))
    inp = driver.Inputs(w.chrom, w.reference, w.variants, list(w.genders), spectra, parameter_kmers=d["parameter_kmers"])
    _, _, _, info = driver.genotype(inp, driver.Options(random_seed=int(d["meta.seed"][0]), n_chains=1, gibbs_burn_in=2, gibbs_samples=2))
    nb_p, nb_size = info["nb"]
    ref = d["tab.nb_p_size"]
    assert np.abs(nb_p / ref[:, 0] - 1).max() < 1e-12, (nb_p, ref[:, 0])
    assert np.abs(nb_size / ref[:, 1] - 1).max() < 1e-12, (nb_size, ref[:, 1])
    # every k-mer on the list was found outside the clusters, at the multiplicity the fit used
    assert all(n > 1000 for _, n, _, _ in info["nb_fit"])


def test_nb_fit_refuses_a_sample_without_a_modal_class(btg):
    """A sample with ploidy 0 on the contig (female on chrY) has no parameter k-mer with multiplicity >= 1: the reference asserts
    (CountDistribution.cpp:113-119); the driver raises instead of carrying NaN tables into the Gibbs stage."""
    w = synth.small_mixed(60, 8000, 1, seed=5, chrom="chrY")
    spectra = synth.sample_spectra(w, 4, 500)
    inp = driver.Inputs(w.chrom, w.reference, w.variants, ["F"], spectra)
    with pytest.raises(Exception, match="negative binomial cannot be fitted|modal multiplicity"):
        driver.genotype(inp, driver.Options(n_chains=1, gibbs_burn_in=1, gibbs_samples=1))
