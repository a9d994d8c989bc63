"""The host-side cluster/graph builder (bayestyper_b200/graph_builder.py) against graphs the REFERENCE built
(VariantFileParser + VariantClusterGraph constructor, dumped by oracle-R into the pipe_* fixtures)."""
import numpy as np
import pytest

from bayestyper_b200 import btd, graph_builder
from tests._fixtures import GOLD
from tests.golden.make_fixtures import PIPE_WORKLOADS


@pytest.mark.parametrize("name", list(PIPE_WORKLOADS))
def test_graphs_identical_to_reference(name):
    d = btd.read(GOLD / f"{name}.btd")
    g = {k[2:]: v for k, v in d.items() if k.startswith("g.")}
    w = PIPE_WORKLOADS[name]()
    b = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
    for k in ("group_cluster_off", "cl_vertex_off", "cl_var_off", "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_nested",
              "v_in_off", "v_in_src", "var_pos", "var_dep", "var_nalt", "alt_reflen", "alt_seq", "v_refvar_off", "cluster_idx",
              "group_nvar", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst"):
        assert len(b[k]) == len(g[k]), k
        assert (b[k] == g[k]).all(), k
    # reference_variant_indices come out of an unordered_set: compare as sets per vertex
    for v in range(len(g["v_flags"])):
        a0, a1 = int(g["v_refvar_off"][v]), int(g["v_refvar_off"][v + 1])
        assert set(b["v_refvar"][a0:a1].tolist()) == set(g["v_refvar"][a0:a1].tolist())
    regs = graph_builder.intercluster_regions(w.chrom, w.reference, w.variants)
    ref_regs = sorted((int(a), int(bb)) for dec, a, bb in d["regions"])
    assert sorted(regs) == ref_regs
