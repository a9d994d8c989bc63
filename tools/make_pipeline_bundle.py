"""Writes the input bundle of host/btpipeline (C++ host of both hot paths over the C ABI) for a synthetic workload:
python tools/make_pipeline_bundle.py <out.btd> [n_variants] [length] [n_samples] [seed]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import btd, driver, graph_builder, ploidy as ploidy_rules, synth  # noqa: E402


def bundle(w, spectra, parameter_kmers=None) -> dict:
    g = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
    regions = [(int(a), int(b)) for a, b in g["regions"]]
    pf, pm = ploidy_rules.ChromosomePloidy([w.chrom], list(w.genders), None).gender_ploidy(w.chrom)
    out = {"g." + k: np.ascontiguousarray(v) for k, v in g.items() if isinstance(v, np.ndarray) and k in (
        "group_cluster_off", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst", "cluster_idx", "cl_vertex_off", "v_seq_off",
        "seq", "v_flags", "v_in_off", "v_in_src", "v_var", "v_allele", "v_refvar_off", "v_refvar", "v_nested", "cl_var_off", "var_dep")}
    out["g.var_nalleles"] = (1 + np.asarray(g["var_dep"], np.uint16) + np.asarray(g["var_nalt"], np.uint16)).astype(np.uint16)
    for k, dt in (("g.group_cluster_off", np.uint64), ("g.group_src_off", np.uint64), ("g.group_src", np.uint32), ("g.group_edge_off", np.uint64), ("g.group_edge_src", np.uint32),
                  ("g.group_edge_dst", np.uint32), ("g.cluster_idx", np.uint32), ("g.cl_vertex_off", np.uint64), ("g.v_seq_off", np.uint64), ("g.seq", np.uint8), ("g.v_flags", np.uint8),
                  ("g.v_in_off", np.uint64), ("g.v_in_src", np.uint32), ("g.v_var", np.uint16), ("g.v_allele", np.uint16), ("g.v_refvar_off", np.uint64), ("g.v_refvar", np.uint16),
                  ("g.v_nested", np.uint32), ("g.cl_var_off", np.uint64), ("g.var_dep", np.uint8)):
        if k in out:
            out[k] = np.ascontiguousarray(out[k], dt)
    out["regions"] = driver._region_buffer(w.reference, regions)
    for s, (km, ct) in enumerate(spectra):
        out[f"s{s}.kmers"] = np.ascontiguousarray(km, np.uint64).reshape(-1, 2)
        out[f"s{s}.counts"] = np.ascontiguousarray(ct, np.uint8)
    out["meta.genders"] = np.array([0 if x in ("F", 0) else 1 for x in w.genders], np.uint8)
    out["meta.ploidy"] = np.array([pf, pm], np.uint32)
    if parameter_kmers is not None:
        out["parameter_kmers"] = np.ascontiguousarray(parameter_kmers, np.uint64).reshape(-1, 2)
    return out


if __name__ == "__main__":
    n_var = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    length = int(sys.argv[3]) if len(sys.argv) > 3 else 40_000
    S = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 71
    w = synth.small_mixed(n_var, length, S, seed=seed)
    btd.write(sys.argv[1], bundle(w, synth.sample_spectra(w, 4, 2000)))
