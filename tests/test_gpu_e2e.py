"""End to end on the GPU: reference sequence + candidate variants + sample k-mer spectra -> graphs (host) ->
path search -> k-mer table -> haplotype candidates -> NB fit -> noise estimation -> Gibbs genotyping, entirely
through libbtgpu (bayestyper_b200/driver.py), compared with what the REFERENCE's cluster + genotype wrote for the
same inputs (oracle-R fixtures).  Different random streams (Philox vs mt19937): the comparison is statistical."""
import numpy as np
import pytest

from bayestyper_b200 import btd, driver, synth
from tests._fixtures import GOLD
from tests.golden.make_fixtures import E2E_WORKLOADS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(E2E_WORKLOADS))
def test_pipeline_matches_reference_calls(btg, name):
    d = btd.read(GOLD / f"{name}.btd")
    w = E2E_WORKLOADS[name]()
    spectra = synth.sample_spectra(w, 4, int(d["meta.n_errors"][0]))
    graphs, unit, res, info = driver.run(w.chrom, w.reference, w.variants, spectra, w.genders, driver.Options(random_seed=int(d["meta.seed"][0])))
    S = len(spectra)
    assert (graphs["var_pos"] == d["ref.var_pos"]).all()                      # same clusters, same order
    # negative-binomial fit from parameter k-mers (different subsample of the same population)
    nb_p, nb_size = info["nb"]
    mean = nb_size * (1 - nb_p) / nb_p
    ref_mean = d["tab.nb_p_size"][:, 1] * (1 - d["tab.nb_p_size"][:, 0]) / d["tab.nb_p_size"][:, 0]
    assert np.abs(mean / ref_mean - 1).max() < 0.03
    assert np.abs(nb_p / d["tab.nb_p_size"][:, 0] - 1).max() < 0.1
    # noise rates: same order of magnitude (20 x 250 draws around a posterior whose width depends on the k-mer subsample)
    r = info["noise_rates"] / d["tab.noise_rates"]
    assert (r > 0.3).all() and (r < 3).all()
    # calls
    gt_o, gt_r = res["gt"].reshape(-1, S, 2), d["ref.gt"].reshape(-1, S, 2)
    same = (gt_o == gt_r).all(axis=2)
    assert same.mean() > 0.98
    called_both = (gt_o[..., 0] != 0xFFFF) & (gt_r[..., 0] != 0xFFFF)
    assert (~same & called_both).sum() <= max(2, int(0.003 * same.size))     # disagreements are threshold (./.) cases
    dg = np.abs(res["gpp"] - d["ref.gpp"])
    assert dg.mean() < 3e-3
    # and against the truth the spectra were drawn from
    truth = w.genotypes.transpose(1, 0, 2).sum(axis=2)                        # (variants, S) alt allele count
    order = np.argsort([v.pos for v in w.variants])
    pos_to_truth = {w.variants[i].pos + 1: truth[i] for i in order}
    t = np.array([pos_to_truth[int(p)] for p in graphs["var_pos"]])
    called = gt_o[..., 0] != 0xFFFF
    alt_count = np.where(called, gt_o.astype(np.int64).sum(axis=2), -1)
    assert (alt_count[called] == t[called]).mean() > 0.995
    assert called.mean() > 0.93


def _e2e_tool():
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("e2e_check", Path(__file__).resolve().parent.parent / "tools" / "e2e_check.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_nested_composition_matches_reference_calls(btg, capsys):
    """Deletions that contain other variants, end to end: host cluster construction -> nested path search -> nested tables -> group-per-thread
    Gibbs, against the reference's calls for the same candidate set (fixture e2e_nested_2s; the checks are tools/e2e_check.py's, the same
    statistical bars as above)."""
    rc = _e2e_tool().main("e2e_nested_2s")
    out = capsys.readouterr().out
    assert rc == 0, out
    assert "FAIL" not in out


def test_genome_composition_matches_reference_vcf(btg, capsys):
    """driver_genome.genotype_genome — several contigs, a contig without variants, a decoy, haploid chrX calls of the male sample — against
    the VCF the reference wrote for the same genome (tests/golden/vcf_genome_2s.vcf.gz): same records in the same order, same VCS / VCR /
    VCGS / VCGR, GT agreement and posterior distance at the bars of the single-contig cases."""
    rc = _e2e_tool().main_genome()
    out = capsys.readouterr().out
    assert rc == 0, out
    assert "FAIL" not in out
