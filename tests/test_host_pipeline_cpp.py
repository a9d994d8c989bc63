"""host/btpipeline — a C++ program, no Python or torch in its process — drives BOTH hot paths through the C ABI for one unit: sample Bloom filters,
findVariantClusterPaths, the KmerCounter stages (btg_counter), the NB fit, estimateNoise and estimateGenotypes.  Its results must equal, bit for bit,
what the Python mirror (driver.genotype, same library, same seeds) returns for the same inputs."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from bayestyper_b200 import btd, driver, synth

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,seed", [(1, 71), (3, 72)])
def test_cpp_host_runs_both_paths(btg, tmp_path, S, seed):
    import make_pipeline_bundle as mpb
    exe = ROOT / "host" / "btpipeline"
    if not exe.exists():
        pytest.skip("host/btpipeline not built")
    w = synth.small_mixed(300, 30_000, S, seed=seed)
    spectra = synth.sample_spectra(w, 4, 1500)
    btd.write(tmp_path / "bundle.btd", mpb.bundle(w, spectra))
    args = ["--random-seed", "20190401", "--gibbs-burn-in", "20", "--gibbs-samples", "40", "--number-of-gibbs-chains", "3"]
    out = subprocess.run([str(exe), str(tmp_path / "bundle.btd"), str(tmp_path / "out.btd"), *args], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = btd.read(tmp_path / "out.btd")
    inp = driver.Inputs(w.chrom, w.reference, w.variants, list(w.genders), spectra)
    _, _, res, info = driver.genotype(inp, driver.Options(random_seed=20190401, gibbs_burn_in=20, gibbs_samples=40, n_chains=3))
    assert (got["nb_p"] == info["nb"][0]).all() and (got["nb_size"] == info["nb"][1]).all()
    assert (got["noise_rates"] == info["noise_rates"]).all()
    for k in ("gt", "gq", "gpp", "app", "nak", "fak", "mac", "saf", "an", "ac", "acp"):
        assert (got[k] == res[k]).all(), k
    assert int(got["meta"][0]) == info["n_path_kmers"] and int(got["meta"][2]) > 10      # kernels were launched by the C++ process itself
