"""The C++ host over the C ABI (include/btgpu.hpp, host/btgenotype.cpp): builds with g++ against libbtgpu.so, refuses to
run without a GPU (no CPU fallback), and on a GPU produces exactly what the Python mirror produces for the same unit."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from bayestyper_b200 import btd, build

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def test_host_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = build.build_host()
    assert exe.exists()
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe), str(GOLD / "gibbs_snv_1s.btd"), str(tmp_path / "o.btd")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,joint", [("gibbs_mixed_3s", False), ("gibbs_chrx_2s", True)])
def test_cpp_host_equals_python_mirror(btg, tmp_path, name, joint):
    from bayestyper_b200 import engine
    from tests._fixtures import GibbsFixture
    exe = build.build_host()
    out = tmp_path / "res.btd"
    args = [str(exe), str(GOLD / f"{name}.btd"), str(out), "--number-of-gibbs-chains", "2", "--gibbs-burn-in", "10", "--gibbs-samples", "20"]
    fx = GibbsFixture(name)
    args += ["--random-seed", str(fx.seed)] + (["--noise-genotyping"] if joint else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = btd.read(out)
    opts = fx.opts(chains=2, burn=10, samples=20)
    cd = engine.CountDistribution(fx.nb_p, fx.nb_size)
    eng = engine.InferenceEngine(fx.unit)
    if joint:
        want, trace = eng.estimate_noise_and_genotypes(cd, opts)
    else:
        trace = eng.estimate_noise(cd, opts)
        want = eng.estimate_genotypes(cd, opts)
    assert (got["noise_trace"] == trace).all() and (got["noise_rates"] == cd.noise_rates()).all()
    for k in ("gt", "gq", "gpp", "app", "nak", "fak", "mac", "saf", "ploidy", "an", "ac", "af", "acp", "anc", "hc"):
        assert (got[k] == want[k]).all(), k
    eng.close(); cd.close()
