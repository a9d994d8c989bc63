"""Loads the committed Gibbs fixtures (tests/golden/*.btd, made by make_fixtures.py from the reference)."""
from pathlib import Path

import numpy as np

from bayestyper_b200 import btd, unit as U

GOLD = Path(__file__).parent / "golden"
GIBBS_FIXTURES = ["gibbs_snv_1s", "gibbs_mixed_3s", "gibbs_chrx_2s", "gibbs_nested_2s", "gibbs_deep_2s"]


class GibbsFixture:
    def __init__(self, name):
        d = btd.read(GOLD / f"{name}.btd")
        self.name = name
        self.S = int(d["meta.n_samples"][0])
        self.unit = U.Unit({k[5:]: v for k, v in d.items() if k.startswith("unit.")}, self.S)
        self.tab = {k[4:]: v for k, v in d.items() if k.startswith("tab.")}
        self.ref = {k[4:]: v for k, v in d.items() if k.startswith("ref.")}
        self.groups = d["meta.groups"]
        self.seed = int(d["meta.seed"][0])
        self.noise_trace = d["meta.noise_trace"]
        self.nb_p = self.tab["nb_p_size"][:, 0].copy()
        self.nb_size = self.tab["nb_p_size"][:, 1].copy()

    def opts(self, **kw):
        kw.setdefault("seed", self.seed)
        kw.setdefault("min_frac", U.min_fraction_observed(self.nb_p, self.nb_size))
        return U.default_opts(**kw)


def gt_strings(gt, S):
    """(n_variants, S) list of VCF GT strings from the flat gt array."""
    g = gt.reshape(-1, S, 2)
    out = []
    for v in range(g.shape[0]):
        row = []
        for s in range(S):
            a, b = int(g[v, s, 0]), int(g[v, s, 1])
            fa = "." if a == 0xFFFF else str(a)
            if b == 0xFFFE:
                row.append(fa)
            else:
                row.append(fa + "/" + ("." if b == 0xFFFF else str(b)))
        out.append(row)
    return out
