"""Host-side logic of bench.py that needs no GPU: the correctness-at-size check printed in the bench line and the sizing of the
reference arm's bounded sample."""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bench"] = mod
    spec.loader.exec_module(mod)
    return mod


class _Inp:
    pass


def _inputs(n_variants=300, S=2, seed=5):
    from bayestyper_b200 import graph_builder, synth
    w = synth.small_mixed(n_variants, 30_000, S, seed=seed)
    inp = _Inp()
    inp.variants = w.variants
    inp.truth = w.genotypes
    inp.graphs = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
    return w, inp


def test_truth_agreement_of_perfect_calls_and_of_wrong_calls():
    b = _bench()
    S = 2
    w, inp = _inputs(S=S)
    pos = {v.pos + 1: i for i, v in enumerate(w.variants)}
    idx = np.array([pos[int(p)] for p in inp.graphs["var_pos"]])
    gt = np.ascontiguousarray(w.genotypes.transpose(1, 0, 2)[idx]).astype(np.uint16)      # (variants in unit order, S, 2): the truth itself
    out = b.truth_agreement(inp, {"gt": gt.reshape(-1)}, S)
    assert out["gt_matches_truth_frac"] == 1.0 and out["called_frac"] == 1.0 and out["variant_sample_pairs"] == gt.shape[0] * S
    # an uncalled genotype is left out; a wrong one is counted
    gt2 = gt.copy()
    gt2[0, 0] = 0xFFFF
    gt2[1, 1] = (gt2[1, 1] + 1) % 2 if gt2[1, 1].sum() != 1 else np.array([0, 0], np.uint16)
    out = b.truth_agreement(inp, {"gt": gt2.reshape(-1)}, S)
    n = gt.shape[0] * S
    assert abs(out["called_frac"] - (n - 1) / n) < 1e-12
    assert abs(out["gt_matches_truth_frac"] - (n - 2) / (n - 1)) < 1e-12
    # rows that do not line up with the candidate set are reported, not mis-scored
    assert "error" in b.truth_agreement(inp, {"gt": gt.reshape(-1)[:-2 * S]}, S)
    assert b.truth_agreement(_Inp(), {"gt": gt.reshape(-1)}, S) is None                  # no truth attached (inputs from files)


def test_reference_arm_sample_follows_the_step_count():
    b = _bench()
    assert b.reference_sample_variants("B", 1, 0) == 12_000
    assert b.reference_sample_variants("B", 3, 3) == 12_000
    assert b.reference_sample_variants("B", 20, 5) == 2_857            # the driver's K = 20, W = 5: 21 runs
    assert b.reference_sample_variants("B", 1000, 5) == 2_000
    assert b.reference_sample_variants("D", 1, 0) == 120 and b.reference_sample_variants("D", 20, 5) == 40
    # the projection of the reference arm uses the full config's own counts
    assert b.FULL_B["variants"] == 299_455 and b.FULL_B["clusters"] == 200_730
