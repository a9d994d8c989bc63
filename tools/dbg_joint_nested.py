import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["BTG_WIDE"] = sys.argv[2] if len(sys.argv) > 2 else "0"
from bayestyper_b200 import capi, engine
from tests import _oracle as O
from tests._fixtures import GibbsFixture
lib = capi.load(); capi.check(lib.btg_init(0), lib)
fx = GibbsFixture(sys.argv[1] if len(sys.argv) > 1 else "gibbs_joint_nested_2s")
u = fx.unit
gsz = np.diff(u.a["group_cluster_off"].astype(np.int64))
print("groups", u.G, "nested groups", (gsz > 1).sum(), "S", u.S)
def run(unit, tag):
    opts = fx.opts(chains=1, burn=2, samples=3)
    ocd = O.OracleCountDist(fx.nb_p, fx.nb_size); gcd = engine.CountDistribution(fx.nb_p, fx.nb_size)
    ores, otrace = O.oracle_estimate_noise_and_genotypes(unit, ocd, opts)
    eng = engine.InferenceEngine(unit)
    gres, gtrace = eng.estimate_noise_and_genotypes(gcd, opts)
    rel = np.abs(gtrace[:, 2:] - otrace[:, 2:]) / otrace[:, 2:]
    print(tag, "max rel", rel.max(), "rows", [(int(r), float(rel[r].max())) for r in range(len(rel))])
    eng.close()
run(u, "all groups")
run(u.subset_groups(np.flatnonzero(gsz == 1)), "single-cluster groups only")
run(u.subset_groups(np.flatnonzero(gsz > 1)), "nested groups only")
for g in np.flatnonzero(gsz > 1)[:6]:
    run(u.subset_groups(np.array([g])), f"group {g} ({gsz[g]} clusters)")
