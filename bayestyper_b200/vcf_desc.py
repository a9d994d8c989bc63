"""The "vcf.*" description that the VCF writer (include/btgpu_vcf.hpp, host/btvcf, host/btgenotype --vcf) needs next to the result
arrays: per variant of the unit, in unit order, what the reference keeps in VariantInfo and in the cluster fields of Genotypes
(VCS / VCR / VCGS / VCGR).  Built from graph_builder's output (groups of several clusters included)."""
from __future__ import annotations

import numpy as np


def _strs(items):
    b = [x.encode() if isinstance(x, str) else bytes(x) for x in items]
    off = np.concatenate([[0], np.cumsum([len(x) for x in b])]).astype(np.uint64)
    return (np.frombuffer(b"".join(b), np.uint8).copy() if b else np.zeros(0, np.uint8)), off


def unit_variant_order(graphs: dict) -> np.ndarray:
    """Index into the caller's (position-sorted) variant list for every variant of the unit, in unit order."""
    return np.asarray(graphs["var_input_idx"], np.int64)


def describe(chrom: str, reference: bytes, variants, graphs: dict, sample_names, genome_filename: str = "", graph_options_header: str = "",
             genotype_options_header: str = "", ids=None) -> dict:
    """BTD1 arrays "vcf.*" (see host/btvcf.cpp) for the unit of ONE contig (graph_builder.build_unit_graphs).
    ids: per input variant (default: the candidates' own ids, else v<index>, the ids synth.write_workdir writes)."""
    vorder = unit_variant_order(graphs)
    if ids is None:
        ids = [getattr(v, "id", None) or f"v{i}" for i, v in enumerate(variants)]
    ids = list(ids)
    n_var = len(graphs["var_pos"])
    return _describe([chrom], [reference], [0], np.zeros(n_var, np.uint32), np.zeros(len(graphs["group_cluster_off"]) - 1, np.uint32),
                     [ids[i] for i in vorder], graphs, sample_names, genome_filename, graph_options_header, genotype_options_header)


def describe_genome(genome: dict, candidates: dict, graphs: dict, sample_names, decoys=(), genome_filename: str = "", graph_options_header: str = "",
                    genotype_options_header: str = "") -> dict:
    """The same for the unit of a whole genome (graph_builder.build_genome_graphs): contigs in genome order — the order of the
    `##contig` lines and of the records (GenotypeWriter.cpp:460-481,523-531) — decoy contigs flagged so that they stay out of the header."""
    names = list(graphs["contig_names"])
    decoys = set(decoys)
    ids = []
    for c, i in zip(np.asarray(graphs["var_contig"], np.int64), unit_variant_order(graphs)):
        v = candidates[names[c]][int(i)]
        ids.append(getattr(v, "id", None) or f"{names[c]}_{int(i)}")
    return _describe(names, [genome[n] for n in names], [int(n in decoys) for n in names], np.asarray(graphs["var_contig"], np.uint32),
                     np.asarray(graphs["group_contig"], np.uint32), ids, graphs, sample_names, genome_filename, graph_options_header, genotype_options_header)


def _describe(contig_names, contig_seqs, contig_decoy, var_contig, group_contig, unit_ids, graphs, sample_names, genome_filename, graph_options_header,
              genotype_options_header) -> dict:
    cvo = np.asarray(graphs["cl_var_off"], np.int64)
    pos = np.asarray(graphs["var_pos"], np.int64)                       # 1-based
    vao = np.asarray(graphs["var_alt_off"], np.int64)
    reflen = np.asarray(graphs["alt_reflen"], np.int64)
    aso = np.asarray(graphs["alt_seq_off"], np.int64)
    alt_bytes = bytes(np.asarray(graphs["alt_seq"], np.uint8))
    n_var = len(pos)
    vcs = np.zeros(n_var, np.uint32)
    vcgs = np.zeros(n_var, np.uint32)
    vcr = [None] * n_var
    vcgr = [None] * n_var
    gco = np.asarray(graphs["group_cluster_off"], np.int64)
    for g in range(len(gco) - 1):                                                                    # VariantClusterGroup::region, number of clusters
        v0, v1 = cvo[gco[g]], cvo[gco[g + 1]]
        vcgs[v0:v1] = gco[g + 1] - gco[g]
        vcgr[v0:v1] = [f"{contig_names[group_contig[g]]}:{int(graphs['group_start'][g])}-{int(graphs['group_end'][g])}"] * (v1 - v0)
    for c in range(len(cvo) - 1):
        v0, v1 = cvo[c], cvo[c + 1]
        end = max(pos[v] - 1 + reflen[vao[v]:vao[v + 1]].max() - 1 for v in range(v0, v1)) + 1       # 1-based end of the cluster's reference span
        region = f"{contig_names[var_contig[v0]]}:{pos[v0]}-{end}"                                   # VariantCluster region (first variant .. last reference base)
        for v in range(v0, v1):
            vcs[v] = v1 - v0
            vcr[v] = region
    a = {}
    a["vcf.sample_names"], a["vcf.sample_names_off"] = _strs(sample_names)
    a["vcf.contig_names"], a["vcf.contig_names_off"] = _strs(contig_names)
    a["vcf.contig_seq"], a["vcf.contig_seq_off"] = _strs(contig_seqs)
    a["vcf.contig_decoy"] = np.asarray(contig_decoy, np.uint8)
    a["vcf.genome_filename"] = np.frombuffer(genome_filename.encode(), np.uint8).copy()
    a["vcf.graph_options_header"] = np.frombuffer(graph_options_header.encode(), np.uint8).copy()
    a["vcf.genotype_options_header"] = np.frombuffer(genotype_options_header.encode(), np.uint8).copy()
    a["vcf.ids"], a["vcf.ids_off"] = _strs(unit_ids)
    a["vcf.vcr"], a["vcf.vcr_off"] = _strs(vcr)
    a["vcf.vcgr"], a["vcf.vcgr_off"] = _strs(vcgr)
    a["vcf.alt_seq"], a["vcf.alt_seq_off"] = _strs([alt_bytes[aso[i]:aso[i + 1]] for i in range(len(reflen))])
    a["vcf.alt_aco"], a["vcf.alt_aco_off"] = _strs(graphs.get("alt_aco") or [""] * len(reflen))
    a["vcf.alt_ref_length"] = reflen.astype(np.uint32)
    a["vcf.alt_off"] = vao.astype(np.uint64)
    a["vcf.contig"] = np.asarray(var_contig, np.uint32)
    a["vcf.position"] = pos.astype(np.uint32)
    a["vcf.has_dependency"] = np.asarray(graphs["var_dep"], np.uint8)
    a["vcf.vcs"] = vcs
    a["vcf.vcgs"] = vcgs
    return a


def write_vcf(out_path, result: dict, description: dict, n_samples: int) -> None:
    """Write `<out>.vcf` from the result arrays of InferenceEngine.estimate_genotypes (unit order) and describe()'s arrays through the
    C++ writer (host/btvcf, include/btgpu_vcf.hpp — the restatement of the reference's GenotypeWriter)."""
    import subprocess
    import tempfile
    from pathlib import Path

    from . import btd, build
    build.build_host()
    exe = Path(build.ROOT) / "host" / "btvcf"
    keys = ("gt", "gq", "gpp", "app", "nak", "fak", "mac", "saf", "ploidy", "an", "ac", "af", "acp", "anc", "hc")
    arrays = dict(description)
    arrays["meta.n_samples"] = np.array([n_samples], np.uint32)
    arrays.update({k: np.ascontiguousarray(result[k]) for k in keys})
    with tempfile.TemporaryDirectory() as td:
        btd.write(Path(td) / "in.btd", arrays)
        r = subprocess.run([str(exe), str(Path(td) / "in.btd"), str(out_path)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("btvcf failed: " + r.stderr.strip())
