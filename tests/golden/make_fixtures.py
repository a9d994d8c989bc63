"""Generates the committed Gibbs fixtures from the REFERENCE ITSELF (oracle-R = the
reference's translation units compiled in place, oracle/ref_build).  Run in the build
container (needs /root/reference):  python tests/golden/make_fixtures.py

Each fixture <name>.btd holds, for a seeded synthetic workload (bayestyper_b200/synth.py):
  unit.*   the flat haplotype-candidate descriptors dumped from VariantClusterGraph::
           getHaplotypeCandidates (a13/a14 output) for a subset of groups
  tab.*    CountDistribution: NB parameters, noise rates after estimateNoise, both log-pmf caches
  ref.*    what the reference wrote to its VCF for those variants (GT, GQ, GPP, APP, NAK, FAK, MAC, SAF)
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import btd, synth, unit as U, vcfio  # noqa: E402

BTREF = ROOT / "oracle" / "_ref" / "btref"


def make(name, workload, n_groups, seed=20190401, n_errors=20000, extra_args=(), store_tables=True):
    with tempfile.TemporaryDirectory() as td:
        wd = synth.write_workdir(workload, td, n_errors=n_errors)
        subprocess.check_call([str(BTREF), "run", "--workdir", str(wd), "--threads", "8", "--seed", str(seed), "--dump-graphs", "--dump-haps", *extra_args],
                              stdout=subprocess.DEVNULL)
        out = Path(wd) / "ref_out"
        g = btd.read(out / "graphs.btd"); h = btd.read(out / "haps.btd"); t = btd.read(out / "tables.btd")
        S = int(h["meta"][0])
        # ploidy per group/sample as ChromosomePloidy assigns it (default human rules, ChromosomePloidy.cpp:40-105)
        G = len(g["group_cluster_off"]) - 1
        chrom = workload.chrom.lower()
        pl = np.full((G, S), 2, np.uint8)
        for s, gd in enumerate(workload.genders):
            if chrom in ("x", "chrx") and gd == "M":
                pl[:, s] = 1
            if chrom in ("y", "chry"):
                pl[:, s] = 1 if gd == "M" else 0
        un = U.from_ref_dumps(h, g, workload.genders, pl.reshape(-1))
        # keep a spread of groups: the largest ones and a stride through the rest
        rng = np.random.default_rng(7)
        nested_groups = np.flatnonzero(np.diff(g["group_cluster_off"]) > 1)     # always keep every multi-cluster group
        keep = np.unique(np.concatenate([np.arange(min(20, G)), nested_groups, rng.choice(G, size=min(n_groups, G), replace=False)]))
        sub = un.subset_groups(keep)
        _, rows = vcfio.read_vcf(out / "bayestyper.vcf")
        byid = {r["id"]: r for r in rows}
        ids = [bytes(g["var_ids"][g["var_id_off"][i]:g["var_id_off"][i + 1]]).decode() for i in range(len(g["var_pos"]))]
        gco, cvo = un.a["group_cluster_off"], un.a["cl_var_off"]
        var_sel = np.concatenate([np.arange(cvo[c], cvo[c + 1]) for gi in keep for c in range(gco[gi], gco[gi + 1])]).astype(np.int64)
        res, arr = sub.alloc_result()
        ref = {k: arr[k].copy() for k in ("gt", "gq", "gpp", "app", "nak", "fak", "mac", "saf")}
        nAs = sub.a["var_nalleles"]
        for j, v in enumerate(var_sel):
            r = byid[ids[v]]
            nA = int(nAs[j]); nG = nA * (nA + 1) // 2
            for s in range(S):
                sr = r["samples"][s]
                gt = sr["GT"].replace("|", "/").split("/")
                o = (j * S + s) * 2
                ref["gt"][o] = 0xFFFF if gt[0] == "." else int(gt[0])
                ref["gt"][o + 1] = (0xFFFF if gt[1] == "." else int(gt[1])) if len(gt) > 1 else 0xFFFE
                ref["gq"][j * S + s] = sr.get("GQ", 0)
                if "GPP" in sr:
                    gp = np.array(sr["GPP"], np.float32)
                    ref["gpp"][int(arr["geno_off"][j]) + s * nG: int(arr["geno_off"][j]) + s * nG + len(gp)] = gp
                a0 = int(arr["allele_off"][j]) + s * nA
                for k in ("app", "nak", "fak", "mac", "saf"):
                    if k.upper() in sr:
                        ref[k][a0:a0 + nA] = np.array(sr[k.upper()], ref[k].dtype)
        pack = {"unit." + k: v for k, v in sub.a.items()}
        # the two log-pmf caches are 0.5 MB per sample: fixtures with many samples keep the parameters only (the caches are pinned by the others)
        pack.update({"tab." + k: v for k, v in t.items() if store_tables or k not in ("genomic_log_pmf", "noise_log_pmf")})
        pack.update({"ref." + k: v for k, v in ref.items()})
        pack["meta.n_samples"] = np.array([S], np.uint32)
        pack["meta.groups"] = keep.astype(np.uint32)            # indices in the full unit (seed derivation)
        pack["meta.seed"] = np.array([seed], np.uint32)
        pack["meta.noise_trace"] = np.loadtxt(out / "bayestyper_noise_parameters.txt", skiprows=1, ndmin=2)
        btd.write(Path(__file__).parent / f"{name}.btd", pack)
        print(name, "groups", len(keep), "clusters", sub.Cn, "variants", sub.n_variants, "H max", int(sub.a["cl_nhap"].max()))


def make_paths(name, workload, seed=20190401, n_errors=5000, max_hap=32):
    """Path-search fixture: the reference's graphs (VariantClusterGraph ctor) and the best_paths_indices its
    findVariantClusterPaths left after all samples.  The sample k-mer sets are NOT stored: the test regenerates
    them with the same seeded generator and checks their checksum."""
    import hashlib
    with tempfile.TemporaryDirectory() as td:
        spectra = synth.sample_spectra(workload, 4, n_errors)
        wd = synth.write_workdir(workload, td, spectra=spectra)
        subprocess.check_call([str(BTREF), "run", "--workdir", str(wd), "--threads", "4", "--seed", str(seed), "--dump-graphs",
                               "--cluster-only", "--max-sample-haplotypes", str(max_hap)], stdout=subprocess.DEVNULL)
        g = btd.read(Path(wd) / "ref_out" / "graphs.btd")
    keep = ["group_cluster_off", "cluster_idx", "cl_vertex_off", "v_seq_off", "seq", "v_flags", "v_in_off", "v_in_src", "cl_path_off", "path_bits"]
    pack = {"g." + k: g[k] for k in keep}
    pack["meta.seed"] = np.array([seed], np.uint32)
    pack["meta.max_hap"] = np.array([max_hap], np.uint32)
    pack["meta.n_errors"] = np.array([n_errors], np.uint32)
    pack["meta.kmer_sha"] = np.frombuffer(b"".join(hashlib.sha256(k.tobytes() + c.tobytes()).digest() for k, c in spectra), np.uint8)
    btd.write(Path(__file__).parent / f"{name}.btd", pack)
    nv = np.diff(g["cl_vertex_off"]); npaths = np.diff(g["cl_path_off"]) // np.maximum(nv, 1)
    print(name, "clusters", len(nv), "max V", int(nv.max()), "paths hist", np.bincount(npaths.astype(int))[:12])


def make_pipeline(name, workload, seed=20190401, n_errors=3000):
    """k-mer pipeline fixture: reference graphs + best paths + intercluster regions + multigroup Bloom in,
    the reference's VariantClusterHaplotypes (classifyPathKmers + getHaplotypeCandidates output) out."""
    import hashlib
    with tempfile.TemporaryDirectory() as td:
        spectra = synth.sample_spectra(workload, 4, n_errors)
        wd = synth.write_workdir(workload, td, spectra=spectra)
        subprocess.check_call([str(BTREF), "run", "--workdir", str(wd), "--threads", "1", "--seed", str(seed), "--dump-graphs", "--dump-haps",
                               "--skip-genotype"], stdout=subprocess.DEVNULL)
        out = Path(wd) / "ref_out"
        g = btd.read(out / "graphs.btd"); h = btd.read(out / "haps.btd"); t = btd.read(out / "tables.btd")
        regions = [ln.split("\t") for ln in (out / "bayestyper_cluster_data" / "intercluster_regions.txt.gz").read_text().splitlines()]
        mg_meta = (out / "bayestyper_cluster_data" / "multigroup_kmers.bloomMeta").read_text().split()
        mg_data = np.fromfile(out / "bayestyper_cluster_data" / "multigroup_kmers.bloomData", np.uint8)
    pack = {"g." + k: v for k, v in g.items()}
    pack.update({"h." + k: v for k, v in h.items()})
    pack["t.nb_p_size"] = t["nb_p_size"]
    pack["regions"] = np.array([[int(r[1]), int(r[2]), int(r[3])] for r in regions], np.int64).reshape(-1, 3)
    pack["mg.meta"] = np.array([int(mg_meta[0]), int(mg_meta[1])], np.uint64)
    pack["mg.data"] = mg_data
    pack["meta.seed"] = np.array([seed], np.uint32)
    pack["meta.n_errors"] = np.array([n_errors], np.uint32)
    pack["meta.kmer_sha"] = np.frombuffer(b"".join(hashlib.sha256(k.tobytes() + c.tobytes()).digest() for k, c in spectra), np.uint8)
    btd.write(Path(__file__).parent / f"{name}.btd", pack)
    print(name, "clusters", len(g["cl_vertex_off"]) - 1, "rows", len(h["k_has_counts"]), "regions", len(regions), "multigroup n", mg_meta[0])


def make_e2e(name, workload, seed=20190401, n_errors=20000):
    """End-to-end fixture: what the reference's cluster + genotype wrote for a workload (VCF fields, NB fit, noise rates)."""
    with tempfile.TemporaryDirectory() as td:
        spectra = synth.sample_spectra(workload, 4, n_errors)
        wd = synth.write_workdir(workload, td, spectra=spectra)
        subprocess.check_call([str(BTREF), "run", "--workdir", str(wd), "--threads", "8", "--seed", str(seed), "--dump-graphs"], stdout=subprocess.DEVNULL)
        out = Path(wd) / "ref_out"
        g = btd.read(out / "graphs.btd"); t = btd.read(out / "tables.btd")
        _, rows = vcfio.read_vcf(out / "bayestyper.vcf")
    S = len(spectra)
    byid = {r["id"]: r for r in rows}
    ids = [bytes(g["var_ids"][g["var_id_off"][i]:g["var_id_off"][i + 1]]).decode() for i in range(len(g["var_pos"]))]
    nA = (1 + g["var_dep"].astype(np.int64) + g["var_nalt"]).astype(np.int64)
    gt = np.zeros((len(ids), S, 2), np.uint16); gq = np.zeros((len(ids), S), np.uint32)
    gpp_off = np.concatenate([[0], np.cumsum(S * nA * (nA + 1) // 2)]); gpp = np.zeros(int(gpp_off[-1]), np.float32)
    for j, i in enumerate(ids):
        nG = int(nA[j] * (nA[j] + 1) // 2)
        for s_ in range(S):
            sr = byid[i]["samples"][s_]
            a = sr["GT"].replace("|", "/").split("/")
            gt[j, s_, 0] = 0xFFFF if a[0] == "." else int(a[0])
            gt[j, s_, 1] = (0xFFFF if a[1] == "." else int(a[1])) if len(a) > 1 else 0xFFFE
            gq[j, s_] = sr.get("GQ", 0)
            if "GPP" in sr:
                gp = np.array(sr["GPP"], np.float32)
                gpp[int(gpp_off[j]) + s_ * nG:int(gpp_off[j]) + s_ * nG + len(gp)] = gp
    pack = {"ref.gt": gt.reshape(-1), "ref.gq": gq.reshape(-1), "ref.gpp": gpp, "ref.var_pos": g["var_pos"], "tab.nb_p_size": t["nb_p_size"],
            "tab.noise_rates": t["noise_rates"], "meta.seed": np.array([seed], np.uint32), "meta.n_errors": np.array([n_errors], np.uint32)}
    btd.write(Path(__file__).parent / f"{name}.btd", pack)
    print(name, "variants", len(ids), "nb", t["nb_p_size"].tolist(), "noise", t["noise_rates"].tolist())


def make_nbfit(name, workload, seed=20190401, n_errors=8000):
    """NB-fit fixture: the parameter k-mers the reference's cluster stage wrote (<out>_cluster_data/parameter_kmers.fa.gz) and the negative
    binomial (p, size) its genotype stage fitted from them (CountDistribution::setGenomicCountDistributions)."""
    from bayestyper_b200 import cluster_data
    with tempfile.TemporaryDirectory() as td:
        spectra = synth.sample_spectra(workload, 4, n_errors)
        wd = synth.write_workdir(workload, td, spectra=spectra)
        subprocess.check_call([str(BTREF), "run", "--workdir", str(wd), "--threads", "4", "--seed", str(seed), "--skip-genotype"], stdout=subprocess.DEVNULL)
        out = Path(wd) / "ref_out"
        t = btd.read(out / "tables.btd")
        pk = cluster_data.read_parameter_kmers(out / "bayestyper_cluster_data" / "parameter_kmers")
    btd.write(Path(__file__).parent / f"{name}.btd", {"parameter_kmers": np.ascontiguousarray(pk, np.uint64), "tab.nb_p_size": t["nb_p_size"],
                                                         "meta.seed": np.array([seed], np.uint32), "meta.n_errors": np.array([n_errors], np.uint32)})
    print(name, "parameter k-mers", len(pk), "nb", t["nb_p_size"].tolist())


NBFIT_WORKLOADS = {"nbfit_mixed_2s": lambda: synth.small_mixed(250, 30_000, 2, seed=77)}

E2E_WORKLOADS = {
    "e2e_snv_1s": lambda: synth.config_a(n_variants=1500, length=150_000),
    "e2e_mixed_3s": lambda: synth.small_mixed(800, 80_000, 3, seed=71),
    "e2e_configA_1s": lambda: synth.config_a(),          # BASELINE.json configs[0] at FULL size: 1 sample, 10k SNVs on 1 Mb (70 s of the reference here)
}
E2E_N_ERRORS = {"e2e_configA_1s": 100_000}              # sequencing-error k-mers added to the spectra (default 20,000)

# the nested end-to-end case: run by tools/e2e_check.py and, since round 2, by tests/test_gpu_e2e.py::test_nested_composition_matches_reference_calls
E2E_NEXT_WORKLOADS = {
    "e2e_nested_2s": lambda: synth.nested_sv(25, 100_000, 2, seed=31, n_background=400, sv_len=(150, 600), repeat_frac=0.4),
}

DEEP_WORKLOAD = lambda: synth.deep_nested(4, 22_000, 2, seed=7, n_background=400)   # noqa: E731

PIPE_WORKLOADS = {
    "pipe_snv_1s": lambda: synth.config_a(n_variants=500, length=50_000),
    "pipe_mixed_3s": lambda: synth.small_mixed(350, 25_000, 3, seed=52),
    "pipe_chrx_2s": lambda: synth.small_mixed(250, 20_000, 2, seed=61, chrom="chrX"),
    "pipe_nested_2s": lambda: synth.nested_sv(12, 50_000, 2, seed=9, n_background=120, sv_len=(150, 600), repeat_frac=0.6),
    "pipe_deep_2s": DEEP_WORKLOAD,
}

PATH_WORKLOADS = {
    "paths_snv_1s": lambda: synth.config_a(n_variants=1500, length=150_000),
    "paths_mixed_3s": lambda: synth.small_mixed(900, 60_000, 3, seed=21),
    "paths_dense_2s": lambda: synth.small_mixed(1500, 40_000, 2, seed=44, frac_indel=0.3),
    "paths_nested_2s": lambda: synth.nested_sv(20, 80_000, 2, seed=13, n_background=200, sv_len=(150, 600), repeat_frac=0.5),
    "paths_deep_2s": DEEP_WORKLOAD,
}

if __name__ == "__main__":
    import sys as _sys
    if len(_sys.argv) > 1 and _sys.argv[1] == "pipe":
        for nm, fn in PIPE_WORKLOADS.items():
            make_pipeline(nm, fn())
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "e2e-next":
        for nm, fn in E2E_NEXT_WORKLOADS.items():
            make_e2e(nm, fn())
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "joint":
        make("gibbs_joint_2s", synth.small_mixed(260, 24_000, 2, seed=83), 400, extra_args=("--noise-genotyping",))
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "nbfit":
        for nm, fn in NBFIT_WORKLOADS.items():
            make_nbfit(nm, fn())
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "exact":
        # fixtures that hold EVERY group of the reference run: the lock-step modes can be reproduced row for row (tests/test_ref_parity_exact.py)
        make("gibbs_full_2s", synth.small_mixed(150, 15_000, 2, seed=91), 10**6, n_errors=4000)
        make("gibbs_joint_30s", synth.small_mixed(110, 11_000, 30, seed=97), 10**6, n_errors=3000, extra_args=("--noise-genotyping",), store_tables=False)
        make("gibbs_joint_nested_2s", synth.nested_sv(8, 30_000, 2, seed=15, n_background=60, sv_len=(150, 500), repeat_frac=0.6), 10**6, n_errors=4000,
             extra_args=("--noise-genotyping",))
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "deep":
        # deletions inside deletions inside deletions: groups of ten and more clusters, dependency forests several levels deep; the Gibbs fixture holds
        # every group of the reference run
        w = DEEP_WORKLOAD()
        make("gibbs_deep_2s", w, 10**6, n_errors=4000)
        make("gibbs_joint_deep_2s", w, 10**6, n_errors=4000, extra_args=("--noise-genotyping",))     # CPU parity chain only (oracle-P == reference)
        make_pipeline("pipe_deep_2s", w)
        make_paths("paths_deep_2s", w)
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "nested-kmer":
        make_pipeline("pipe_nested_2s", PIPE_WORKLOADS["pipe_nested_2s"]())
        make_paths("paths_nested_2s", PATH_WORKLOADS["paths_nested_2s"]())
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "nested":
        # deletions spanning SNVs (nested clusters, has_dependency) with duplicated segments (multicluster k-mers)
        make("gibbs_nested_2s", synth.nested_sv(30, 120_000, 2, seed=5, n_background=300, sv_len=(150, 600), repeat_frac=0.6), 40)
        _sys.exit(0)
    if len(_sys.argv) > 1 and _sys.argv[1] == "e2e":
        for nm, fn in E2E_WORKLOADS.items():
            make_e2e(nm, fn(), n_errors=E2E_N_ERRORS.get(nm, 20000))
        _sys.exit(0)
    for nm, fn in PATH_WORKLOADS.items():
        make_paths(nm, fn(), max_hap=8 if nm == "paths_dense_2s" else 32)
    make("gibbs_snv_1s", synth.config_a(n_variants=1200, length=120_000), 160)
    make("gibbs_mixed_3s", synth.small_mixed(700, 60_000, 3, seed=21), 120)
    make("gibbs_chrx_2s", synth.small_mixed(400, 50_000, 2, seed=33, chrom="chrX"), 80)
