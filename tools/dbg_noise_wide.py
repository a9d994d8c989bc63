import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["BTG_WIDE"] = sys.argv[2] if len(sys.argv) > 2 else "1"
from bayestyper_b200 import capi, engine
from tests import _oracle as O
from tests._fixtures import GibbsFixture
lib = capi.load(); capi.check(lib.btg_init(0), lib)
fx = GibbsFixture(sys.argv[1] if len(sys.argv) > 1 else "gibbs_snv_1s")
opts = fx.opts(chains=3, burn=20, samples=40)
ocd = O.OracleCountDist(fx.nb_p, fx.nb_size); gcd = engine.CountDistribution(fx.nb_p, fx.nb_size)
otrace = O.oracle_estimate_noise(fx.unit, ocd, opts)
eng = engine.InferenceEngine(fx.unit)
gtrace = eng.estimate_noise(gcd, opts)
rel = np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]
bad = np.flatnonzero(rel.max(axis=1) > 1e-13)
print("rows differing:", len(bad), "of", len(rel))
for r in bad[:20]:
    print(int(r), gtrace[r, :2], rel[r].max(), gtrace[r, 2:5], otrace[r, 2:5])
