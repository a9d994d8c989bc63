"""Host mirror of the stage order of `bayesTyper cluster` + `bayesTyper genotype`
(src/bayesTyper/main.cpp:233-252 and :594-643) over the C ABI, for non-nested candidate sets.

Every per-k-mer / per-cluster computation runs in libbtgpu kernels; this module only sequences
them (like KmerCounter / InferenceEngine do in the reference) and moves arrays across the boundary.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
import time

import numpy as np
import torch

from . import capi, engine, graph_builder, kmer_pipeline, ploidy as ploidy_rules, unit as U

K = 55


class GraphsDesc(C.Structure):
    _fields_ = [("n_clusters", C.c_uint32)] + [(n, C.c_void_p) for n in
                ("cl_vertex_off", "v_seq_off", "seq", "v_flags", "v_in_off", "v_in_src", "cl_group", "cl_idx")]


@dataclasses.dataclass
class Options:
    """Defaults of src/bayesTyper/main.cpp:125-137,378-403."""
    random_seed: int = 20190401
    max_sample_haplotypes: int = 32
    gibbs_burn_in: int = 100
    gibbs_samples: int = 250
    n_chains: int = 20
    kmer_subsampling_rate: float = 0.1
    max_haplotype_variant_kmers: int = 500
    noise_rate_prior: tuple = (1.0, 0.01)
    min_genotype_posterior: float = 0.99
    min_number_of_kmers: float = 1.0
    disable_observed_kmers: bool = False
    bloom_fpr: float = 0.001          # bayesTyperTools makeBloom default (src/bayesTyperTools/main.cpp:127)
    max_parameter_kmers: int = 1_000_000
    chromosome_ploidy_file: str = None   # --chromosome-ploidy-file (ChromosomePloidy.cpp:96-180); default: human X / Y rules by name
    noise_genotyping: bool = False       # --noise-genotyping: InferenceEngine::estimateNoiseAndGenotypes instead of estimateNoise + estimateGenotypes
    noise_split: str = "chains"          # a sharded unit's estimateNoise: "chains" = every rank runs its share of the (independent) chains on the whole unit,
                                         # no exchange while they run; "groups" = every rank runs all chains on its own groups and the ranks add up their
                                         # noise counts inside the chain kernel after every iteration (NVLink mailboxes; what the joint mode always does)
    paths_batch: bool = dataclasses.field(default_factory=lambda: os.environ.get("BTG_PATHS_BATCH", "1") != "0")
                                         # findVariantClusterPaths of all samples in one launch (btg_find_sample_paths_batch; BTG_PATHS_BATCH=0: one launch per sample)
    kmer_stages: str = "abi"             # "abi": KmerCounter's stages through the btg_counter handle of the C ABI (csrc/counter.cu, what a C++ host calls);
                                         # "torch": the torch-glue mirror (kmer_pipeline.py) — also what a sharded unit uses (it subsets the unit on the device)


@dataclasses.dataclass
class Shard:
    """One inference unit over several GPUs (one rank per process and GPU).  Rank r takes the groups r, r + world,
    r + 2 world ... of the size-sorted unit (balanced shards; btg_gibbs_opts.group_index_base / group_index_stride keep every group's
    streams).  `allgather(obj) -> [obj of every rank]` is the host transport (torch.distributed.all_gather_object over NCCL or gloo) for the
    one real exchange of the k-mer path — every rank needs the best paths of ALL clusters to tell which path k-mers occur in several
    groups — and for the mailbox handles of `comm` (shard.Comm), through which the lock-step Gibbs modes add up their noise counts
    inside the chain kernel."""
    world: int
    rank: int
    allgather: object
    comm: object = None

    def my_groups(self, n_groups: int) -> np.ndarray:
        return np.arange(self.rank, n_groups, self.world, dtype=np.int64)


def _take_csr(off, idx):
    off = np.asarray(off).astype(np.int64)
    lens = off[idx + 1] - off[idx]
    new_off = np.concatenate([[0], np.cumsum(lens)])
    return new_off, np.repeat(off[idx] - new_off[:-1], lens) + np.arange(int(new_off[-1]))


def subset_graphs_for_paths(graphs: dict, groups: np.ndarray):
    """The arrays findVariantClusterPaths reads, for the clusters of `groups` only (the clusters keep the index of their group in the
    whole unit: it seeds their path search, KmerCounter.cpp:65).  Returns (graphs subset, cluster indices in the whole unit)."""
    gco = np.asarray(graphs["group_cluster_off"], np.int64)
    new_gco, clusters = _take_csr(gco, groups)
    cvo, verts = _take_csr(graphs["cl_vertex_off"], clusters)
    vso, nts = _take_csr(graphs["v_seq_off"], verts)
    vio, ins = _take_csr(graphs["v_in_off"], verts)
    sub = {"group_cluster_off": new_gco.astype(np.uint64), "cl_vertex_off": cvo.astype(np.uint64), "v_seq_off": vso.astype(np.uint64),
           "seq": np.asarray(graphs["seq"])[nts], "v_flags": np.asarray(graphs["v_flags"])[verts], "v_in_off": vio.astype(np.uint64),
           "v_in_src": np.asarray(graphs["v_in_src"])[ins], "cluster_idx": np.asarray(graphs["cluster_idx"])[clusters],
           "cl_group_global": np.repeat(groups.astype(np.uint32), np.diff(new_gco))}
    return sub, clusters


def merge_best_paths(graphs: dict, parts, cluster_lists):
    """Best paths of every cluster of the unit from the per-rank results [(n_paths, membership bytes)] (cluster_lists[r] = the
    whole-unit cluster indices rank r searched, in its order)."""
    V = np.diff(np.asarray(graphs["cl_vertex_off"], np.int64))
    n_paths = np.zeros(len(V), np.int64)
    for (np_r, _), cl in zip(parts, cluster_lists):
        n_paths[cl] = np_r
    off = np.concatenate([[0], np.cumsum(n_paths * V)])
    mem = np.zeros(int(off[-1]), np.uint8)
    for (np_r, mem_r), cl in zip(parts, cluster_lists):
        lens = (n_paths * V)[cl]
        loc = np.concatenate([[0], np.cumsum(lens)])
        mem[np.repeat(off[cl] - loc[:-1], lens) + np.arange(int(loc[-1]))] = mem_r
    return n_paths, mem


def find_variant_cluster_paths(lib, graphs: dict, sample_blooms, opt: Options, cache: dict | None = None):
    """KmerCounter::findVariantClusterPaths (KmerCounter.cpp:70-103): samples in order, one Bloom at a time."""
    gco = graphs["group_cluster_off"]
    keep = {
        "cl_vertex_off": np.ascontiguousarray(graphs["cl_vertex_off"], np.uint64), "v_seq_off": np.ascontiguousarray(graphs["v_seq_off"], np.uint64),
        "seq": np.ascontiguousarray(graphs["seq"], np.uint8), "v_flags": np.ascontiguousarray(graphs["v_flags"], np.uint8),
        "v_in_off": np.ascontiguousarray(graphs["v_in_off"], np.uint64), "v_in_src": np.ascontiguousarray(graphs["v_in_src"], np.uint32),
        "cl_group": np.ascontiguousarray(graphs["cl_group_global"], np.uint32) if "cl_group_global" in graphs
                    else np.repeat(np.arange(len(gco) - 1, dtype=np.uint32), np.diff(gco).astype(np.int64)),
        "cl_idx": np.ascontiguousarray(graphs["cluster_idx"], np.uint32),
    }
    d = GraphsDesc()
    d.n_clusters = len(keep["cl_vertex_off"]) - 1
    for k, v in keep.items():
        setattr(d, k, v.ctypes.data)
    # the graphs of a unit (and the path-search scratch) stay in HBM between passes when the caller gives a cache (Inputs.resident_cache)
    key = ("graphs", id(graphs), len(sample_blooms), opt.max_sample_haplotypes)
    gr = cache.get(key) if cache is not None else None
    if gr is None:
        gr = capi.check(lib.btg_graphs_upload(C.addressof(d), len(sample_blooms), opt.max_sample_haplotypes), lib)
        if cache is not None:
            cache[key] = gr
    else:
        capi.check(lib.btg_graphs_reset(gr), lib)
    try:
        if len(sample_blooms) > 1 and opt.paths_batch:
            # all samples of the unit in one launch (their filters are resident anyway): same best paths, the slowest cluster's latency paid once
            arr = (C.c_void_p * len(sample_blooms))(*sample_blooms)
            capi.check(lib.btg_find_sample_paths_batch(gr, arr, 0, len(sample_blooms), opt.random_seed, opt.max_sample_haplotypes), lib)
        else:
            for s, b in enumerate(sample_blooms):
                capi.check(lib.btg_find_sample_paths(gr, b, s, opt.random_seed, opt.max_sample_haplotypes), lib)
        n_paths = np.zeros(d.n_clusters, np.uint32)
        off = np.zeros(d.n_clusters + 1, np.uint64)
        capi.check(lib.btg_get_best_paths(gr, capi.ptr(n_paths), capi.ptr(off), None, 0), lib)
        mem = np.zeros(int(off[-1]), np.uint8)
        capi.check(lib.btg_get_best_paths(gr, capi.ptr(n_paths), capi.ptr(off), capi.ptr(mem), mem.size), lib)
    finally:
        if cache is None:
            lib.btg_graphs_free(gr)
    return n_paths.astype(np.int64), mem


def estimate_nb_parameters(pipe: kmer_pipeline.KmerPipeline, region_buf, spectra_dev, genders, opt: Options, ploidy=(2, 2), parameter_kmers=None):
    """Parameter k-mers -> per-sample negative-binomial (p, size): the genotype-side half of
    countInterclusterParameterKmers + calculateKmerStats + setGenomicCountDistributions
    (KmerCounter.cpp:171-250, KmerHash.cpp:257-347, CountDistribution.cpp:66-141).
    Parameter k-mers = inter-cluster reference k-mers that are not path k-mers, Bernoulli-subsampled to at most
    3 x max_parameter_kmers and capped at max_parameter_kmers (the reference's draw order comes from mt19937 +
    hash iteration order; here it is a seeded device permutation — same population, different sample).
    parameter_kmers: (n, 2) uint64 packed k-mers of <out>_cluster_data/parameter_kmers.fa.gz as the `cluster` stage wrote them
    (cluster_data.read_parameter_kmers; main.cpp:543-584 reads the same file): the fit then uses exactly those k-mers, as
    `bayesTyper genotype` does, and the negative-binomial parameters equal the reference's for the same file."""
    lib, dev = pipe.lib, pipe.dev
    with torch.cuda.stream(pipe.ext):
        buf = region_buf
        n = buf.numel()
        km = torch.empty((n, 2), dtype=torch.int64, device=dev)
        valid = torch.empty(n, dtype=torch.uint8, device=dev)
        capi.check(lib.btg_scan_sequence_dev(buf.data_ptr(), n, km.data_ptr(), valid.data_ptr(), None), lib)
        km = km[valid.to(torch.bool)]
        # distinct k-mers + genomic multiplicity (as table keys: lexicographic order)
        n_valid = km.shape[0]
        k_lo = torch.empty(n_valid, dtype=torch.int64, device=dev)
        k_hi = torch.empty(n_valid, dtype=torch.int64, device=dev)
        capi.check(lib.btg_table_keys_from_kmers_dev(km.data_ptr(), n_valid, k_lo.data_ptr(), k_hi.data_ptr(), None), lib)
        o = torch.sort(k_lo, stable=True).indices
        o = o[torch.sort(k_hi[o], stable=True).indices]
        s_lo, s_hi = k_lo[o], k_hi[o]
        new = torch.ones(n_valid, dtype=torch.bool, device=dev)
        new[1:] = (s_lo[1:] != s_lo[:-1]) | (s_hi[1:] != s_hi[:-1])
        keys = km[o][new].contiguous()                              # packed k-mers of the distinct keys, key order
        kw0, kw1 = s_lo[new].contiguous(), s_hi[new].contiguous()
        occ = torch.diff(torch.cat([torch.nonzero(new).squeeze(1), torch.tensor([n_valid], device=dev)]))
        idx = torch.empty(len(keys), dtype=torch.int64, device=dev)
        pipe.use_index()
        capi.check(lib.btg_table_lookup_dev(pipe.kw0.data_ptr(), pipe.kw1.data_ptr(), pipe.n_keys, keys.data_ptr(), len(keys), idx.data_ptr(), None), lib)
        not_path = idx < 0
        kw0, kw1, occ = kw0[not_path], kw1[not_path], occ[not_path]
        if parameter_kmers is not None:
            # the k-mers the cluster stage chose: keep the inter-cluster k-mers that are on the list (their genomic multiplicity is the
            # occurrence count of the scan, KmerCounts::addInterclusterMultiplicity)
            pk = _to_dev(np.ascontiguousarray(parameter_kmers, np.uint64), np.int64, dev).reshape(-1, 2).contiguous()
            p_lo = torch.empty(len(pk), dtype=torch.int64, device=dev); p_hi = torch.empty_like(p_lo)
            capi.check(lib.btg_table_keys_from_kmers_dev(pk.data_ptr(), len(pk), p_lo.data_ptr(), p_hi.data_ptr(), None), lib)
            po = torch.sort(p_lo, stable=True).indices
            po = po[torch.sort(p_hi[po], stable=True).indices]
            p_lo, p_hi = p_lo[po].contiguous(), p_hi[po].contiguous()
            cand = torch.empty((len(occ), 2), dtype=torch.int64, device=dev)
            kw0c, kw1c = kw0.contiguous(), kw1.contiguous()
            capi.check(lib.btg_table_keys_to_kmers_dev(kw0c.data_ptr(), kw1c.data_ptr(), len(occ), cand.data_ptr(), None), lib)
            hit = torch.empty(len(occ), dtype=torch.int64, device=dev)
            pipe.use_index(False)
            capi.check(lib.btg_table_lookup_dev(p_lo.data_ptr(), p_hi.data_ptr(), len(pk), cand.data_ptr(), len(occ), hit.data_ptr(), None), lib)
            sel = hit >= 0
            kw0, kw1, occ = kw0[sel], kw1[sel], occ[sel]
        else:
            total = int(km.shape[0])
            frac = min(1.0, 3.0 * opt.max_parameter_kmers / max(total, 1))
            g = torch.Generator(device=dev).manual_seed(opt.random_seed)
            sel = torch.rand(len(occ), device=dev, generator=g) < frac
            kw0, kw1, occ = kw0[sel], kw1[sel], occ[sel]
            if len(occ) > opt.max_parameter_kmers:
                perm = torch.randperm(len(occ), device=dev, generator=g)[:opt.max_parameter_kmers].sort().values
                kw0, kw1, occ = kw0[perm], kw1[perm], occ[perm]
        kw0, kw1 = kw0.contiguous(), kw1.contiguous()
        S = len(spectra_dev)
        counts = torch.zeros((len(occ), S), dtype=torch.uint8, device=dev)
        rec = torch.zeros(len(occ), dtype=torch.uint8, device=dev)
        pipe.use_index(False)                       # the parameter k-mers are a second, un-indexed table
        for si, (kd, cd_) in enumerate(spectra_dev):
            capi.check(lib.btg_table_add_sample_kmers_dev(kw0.data_ptr(), kw1.data_ptr(), len(occ), kd.data_ptr(), cd_.data_ptr(), cd_.numel(), S, si,
                                                          counts.data_ptr(), rec.data_ptr(), None), lib)
        nb_p, nb_size, used = [], [], []
        for si in range(S):
            mult = torch.clamp(occ * ploidy[0 if genders[si] in ("F", 0) else 1], max=255)
            hist = torch.bincount(mult, minlength=256)[1:33]            # max_nb_kmer_multiplicity = 32 (CountDistribution.cpp:42)
            # CountDistribution::setGenomicCountDistributions asserts a modal class and valid moments (CountDistribution.cpp:113-119): a sample
            # whose ploidy on this contig is 0, or a unit without parameter k-mers, must not reach the Gibbs stage with NaN tables
            if int(hist.max()) == 0:
                raise capi.BtgError(f"sample {si}: no parameter k-mer with genomic multiplicity 1..32 (ploidy {ploidy[0 if genders[si] in ('F', 0) else 1]} on this contig, "
                                    f"{len(occ)} parameter k-mers): the negative binomial cannot be fitted")
            m = int(torch.argmax(hist)) + 1                             # first maximum, as the reference's strict '>' scan
            c = counts[mult == m, si].to(torch.float64)
            if c.numel() < 2:
                raise capi.BtgError(f"sample {si}: {c.numel()} parameter k-mer(s) at the modal multiplicity {m}: mean and variance undefined")
            mean = float(c.mean()); var = float(c.var(unbiased=True))
            p, size = C.c_double(), C.c_double()
            lib.btg_nb_moments_to_parameters(mean, var, m, C.byref(p), C.byref(size))
            nb_p.append(p.value); nb_size.append(size.value); used.append((m, int(hist[m - 1]), mean, var))
    pipe.ext.synchronize()
    return np.array(nb_p), np.array(nb_size), used


@dataclasses.dataclass
class Inputs:
    """Everything `cluster` + `genotype` read for one contig: host-side (the reference's files) ..."""
    chrom: str
    reference: bytes
    variants: list
    genders: list
    spectra: list                      # per sample (kmers (n,2) uint64, counts (n,) uint8) numpy / pinned torch
    blooms: list = None                # per sample (bytes uint8, num_kmers, num_bits) = <prefix>.bloomData/.bloomMeta; None = build
    graphs: dict = None                # filled by prepare()
    regions: list = None
    # ... and optionally already resident in HBM
    spectra_dev: list = None
    blooms_dev: list = None
    region_buf_dev: object = None
    native_builder: bool = False       # cluster construction through host/btcluster (same arrays, ~15x faster) instead of graph_builder.py
    parameter_kmers: object = None     # (n, 2) uint64: <out>_cluster_data/parameter_kmers.fa.gz of the cluster stage (None: chosen here)
    resident_cache: dict = dataclasses.field(default_factory=dict)   # device handles that outlive a pass (graphs of the unit + path-search scratch)

    def prepare(self):
        if self.graphs is None:
            if self.native_builder:
                self.graphs = graph_builder.build_genome_graphs_native({self.chrom: self.reference}, {self.chrom: self.variants})
                self.graphs["regions"] = self.graphs["regions"][:, 2:]
            else:
                self.graphs = graph_builder.build_unit_graphs(self.chrom, self.reference, self.variants)
            self.regions = [(int(a), int(b)) for a, b in self.graphs["regions"]]
        return self

    def make_resident(self, lib, opt: "Options"):
        dev = torch.device("cuda", torch.cuda.current_device())
        if self.spectra_dev is None:
            self.spectra_dev = [(_to_dev(k, np.int64, dev), _to_dev(c, np.uint8, dev)) for k, c in self.spectra]
        torch.cuda.synchronize()
        if self.blooms_dev is None:
            self.blooms_dev = []
            for kd, _ in self.spectra_dev:                             # bayesTyperTools makeBloom
                b = capi.check(lib.btg_bloom_create(kd.shape[0], opt.bloom_fpr, K), lib)
                capi.check(lib.btg_bloom_insert_dev(b, kd.data_ptr(), kd.shape[0], None), lib)
                self.blooms_dev.append(b)
        if self.region_buf_dev is None:
            self.region_buf_dev = torch.from_numpy(_region_buffer(self.reference, self.regions)).to(dev)
        torch.cuda.synchronize()
        return self

    def free(self, lib):
        for b in self.blooms_dev or []:
            lib.btg_bloom_free(b)
        self.blooms_dev = None
        for k, h in list(self.resident_cache.items()):
            if k[0] == "graphs":
                lib.btg_graphs_free(h)
        self.resident_cache.clear()


def _to_dev(a, dtype, dev):
    if isinstance(a, torch.Tensor):
        return a.to(dev, non_blocking=True)
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(dtype) if a.dtype != dtype else a).to(dev)


def _region_buffer(reference: bytes, regions) -> np.ndarray:
    seq = np.frombuffer(reference, np.uint8)
    parts = []
    sep = np.frombuffer(b"N", np.uint8)
    for a, b in regions:
        parts.append(seq[a:b + 1]); parts.append(sep)
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def inputs_from_kmc(chrom: str, reference: bytes, variants, samples) -> Inputs:
    """Inputs for one contig from the files `bayesTyper genotype` reads per sample: samples = [(id, gender, kmc_prefix)] as in
    <samples>.tsv (Sample.cpp:38-70); the KMC databases (KMC1 or KMC2) are read with kmcio.read_kmc, and a <kmc_prefix>.bloomMeta /
    .bloomData pair written by makeBloom is used when present (otherwise the filter is built on the device)."""
    from pathlib import Path

    from . import kmcio
    spectra, blooms, genders = [], [], []
    for _, gender, prefix in samples:
        km, ct, _info = kmcio.read_kmc(prefix)
        spectra.append((km, np.minimum(ct, 255).astype(np.uint8)))          # addSampleCount saturates at 255 (KmerCounts.cpp:178-189)
        genders.append("F" if str(gender).upper().startswith("F") else "M")
        meta = Path(str(prefix) + ".bloomMeta")
        if meta.exists():
            n, bits, k_file = (int(x) for x in meta.read_text().split())
            data = np.fromfile(str(prefix) + ".bloomData", np.uint8)
            if k_file != K:                                            # KmerBloom(prefix) asserts the k-mer size (KmerBloom.cpp:83)
                raise capi.BtgError(f"{meta}: filter built for k = {k_file}, this library is built for k = {K}")
            if data.size != (bits + 7) // 8:
                raise capi.BtgError(f"{prefix}.bloomData holds {data.size} bytes, {meta.name} declares {bits} bits ({(bits + 7) // 8} bytes)")
            blooms.append((data, n, bits))
    return Inputs(chrom, reference, variants, genders, spectra, blooms=blooms if len(blooms) == len(samples) else None)


class _StageClock:
    """BTG_STAGE_TIMES=1: wall time of every stage of genotype() with a device synchronisation after each (diagnostics; off by default)."""

    def __init__(self):
        self.on = os.environ.get("BTG_STAGE_TIMES") == "1"
        self.out = {}
        if self.on:
            torch.cuda.synchronize()
        self.t = time.perf_counter()

    def mark(self, name):
        if not self.on:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        self.out[name] = self.out.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now


def genotype(inp: Inputs, opt: Options | None = None, nb_params=None, noise_rates=None, resident: bool = False, want_unit: bool = False,
             vcf_out=None, sample_names=None, shard: Shard | None = None):
    """One pass of both hot paths: path search -> k-mer table -> haplotype candidates -> NB fit -> noise -> Gibbs.
    resident=True uses the device copies made by Inputs.make_resident (the `value` leg of bench.py); otherwise every
    input crosses the boundary from host memory inside this call (the `e2e` leg).
    shard: this process is one rank of a unit sharded over several GPUs (see Shard): it searches the paths of its own groups, exchanges
    the best paths, runs the (cheap) k-mer table stages on the whole unit and the Gibbs stages on its own groups; the returned result
    arrays and info["n_clusters"] cover this rank's groups only (info["n_clusters_total"]: the unit)."""
    opt = opt or Options()
    lib = capi.load()
    inp.prepare()
    S = len(inp.spectra) if inp.spectra is not None else len(inp.spectra_dev)
    dev = torch.device("cuda", torch.cuda.current_device())
    info = {}
    clk = _StageClock()
    info["stage_wall_ms"] = clk.out
    own_blooms = False
    if resident:
        spectra_dev, blooms, region_buf = inp.spectra_dev, inp.blooms_dev, inp.region_buf_dev
    else:
        spectra_dev = [(_to_dev(k, np.int64, dev), _to_dev(c, np.uint8, dev)) for k, c in inp.spectra]
        torch.cuda.synchronize()
        own_blooms = True
        if inp.blooms is not None:                                      # KmerBloom(prefix): load the sample filters
            blooms = [capi.check(lib.btg_bloom_from_bytes(capi.ptr(b), nk, nb, K), lib) for b, nk, nb in inp.blooms]
        else:
            blooms = []
            for kd, _ in spectra_dev:
                b = capi.check(lib.btg_bloom_create(kd.shape[0], opt.bloom_fpr, K), lib)
                capi.check(lib.btg_bloom_insert_dev(b, kd.data_ptr(), kd.shape[0], None), lib)
                blooms.append(b)
        region_buf = torch.from_numpy(_region_buffer(inp.reference, inp.regions)).to(dev)
        torch.cuda.synchronize()
    female_ploidy, male_ploidy = ploidy_rules.ChromosomePloidy([inp.chrom], inp.genders, opt.chromosome_ploidy_file).gender_ploidy(inp.chrom)
    sharded = shard is not None and shard.world > 1
    G = len(inp.graphs["group_cluster_off"]) - 1
    if sharded:
        mine = shard.my_groups(G)
        sub, my_clusters = subset_graphs_for_paths(inp.graphs, mine)
        if resident:                                   # the shard's sub-graphs are part of the resident state too
            sub = inp.resident_cache.setdefault(("subgraphs", shard.rank, shard.world), sub)
        part = find_variant_cluster_paths(lib, sub, blooms, opt, inp.resident_cache if resident else None)
        parts = shard.allgather((part[0], part[1], my_clusters))
        n_paths, mem = merge_best_paths(inp.graphs, [(p[0], p[1]) for p in parts], [p[2] for p in parts])
    else:
        n_paths, mem = find_variant_cluster_paths(lib, inp.graphs, blooms, opt, inp.resident_cache if resident else None)
    clk.mark("inputs+paths")
    if own_blooms:
        for b in blooms:
            lib.btg_bloom_free(b)
    ploidy = np.tile(np.array([female_ploidy if g in ("F", 0) else male_ploidy for g in inp.genders], np.uint8), G)
    if opt.kmer_stages == "abi" and not sharded and not want_unit:
        return _genotype_abi(lib, inp, opt, n_paths, mem, S, spectra_dev, region_buf, (female_ploidy, male_ploidy), ploidy, nb_params, noise_rates, info, vcf_out, sample_names, clk)
    pipe = kmer_pipeline.KmerPipeline(inp.graphs, n_paths, mem, S, inp.genders)
    info["n_path_kmers"] = pipe.enumerate_path_kmers()
    pipe.scan_buffer(region_buf, female_ploidy, male_ploidy, False)
    for s, (kd, cdv) in enumerate(spectra_dev):
        pipe.add_sample(s, kd, cdv)
    ploidy = np.tile(np.array([female_ploidy if g in ("F", 0) else male_ploidy for g in inp.genders], np.uint8), G)
    # row-level unit arrays stay in HBM unless the caller wants the unit on the host (tests, fixtures)
    unit = pipe.build_unit(multigroup_bloom=None, ploidy=ploidy, device_resident=not want_unit)
    if nb_params is None:
        nb_p, nb_size, used = estimate_nb_parameters(pipe, region_buf, spectra_dev, inp.genders, opt, (female_ploidy, male_ploidy), inp.parameter_kmers)
        info["nb_fit"] = used
    else:
        nb_p, nb_size = nb_params
    cd = engine.CountDistribution(nb_p, nb_size, opt.noise_rate_prior)
    info["n_clusters_total"] = unit.Cn
    sdesc = keep = None
    min_frac = None if opt.disable_observed_kmers else U.min_fraction_observed(nb_p, nb_size)
    if sharded and not opt.noise_genotyping and noise_rates is None and opt.noise_split == "chains":
        # estimateNoise by chains: this rank runs the chains rank, rank + world, ... on the WHOLE unit; the ranks' per-chain sums are added up on the host
        whole = engine.InferenceEngine(unit)
        wopts = U.default_opts(seed=opt.random_seed, burn=opt.gibbs_burn_in, samples=opt.gibbs_samples, chains=opt.n_chains, rate=opt.kmer_subsampling_rate,
                               max_hv=opt.max_haplotype_variant_kmers, min_gpp=opt.min_genotype_posterior, min_kmers=opt.min_number_of_kmers, min_frac=min_frac)
        sums, _ = whole.estimate_noise_chains(cd, wopts, shard.rank, shard.world)
        whole.close()
        total = np.zeros_like(sums)
        for part in shard.allgather(sums):           # rank order; every row is non-zero on exactly one rank, so the order does not matter
            total += part
        cd.finish_noise(total, opt.gibbs_samples)
        noise_rates = cd.noise_rates()
        info["noise_split"] = "chains"
    if sharded:   # the Gibbs stages on this rank's groups; the lock-step modes see the whole unit through the shard descriptor
        from . import shard as shard_mod
        sdesc, keep = shard_mod.shard_desc(unit, shard.comm)
        unit = unit.subset_groups(mine)
    eng = engine.InferenceEngine(unit)
    gopts = U.default_opts(seed=opt.random_seed, burn=opt.gibbs_burn_in, samples=opt.gibbs_samples, chains=opt.n_chains, rate=opt.kmer_subsampling_rate,
                           max_hv=opt.max_haplotype_variant_kmers, min_gpp=opt.min_genotype_posterior, min_kmers=opt.min_number_of_kmers,
                           min_frac=None if opt.disable_observed_kmers else U.min_fraction_observed(nb_p, nb_size),
                           group_base=shard.rank if sharded else 0, group_stride=shard.world if sharded else 1)
    if opt.noise_genotyping:
        res, info["noise_trace"] = eng.estimate_noise_and_genotypes(cd, gopts, want_trace=False, shard=sdesc)
        info["noise_rates"] = cd.noise_rates()
    else:
        if noise_rates is None:
            info["noise_trace"] = eng.estimate_noise(cd, gopts, want_trace=False, shard=sdesc)
        else:
            cd.set_noise_rates(noise_rates)
        info["noise_rates"] = cd.noise_rates()
        res = eng.estimate_genotypes(cd, gopts)
    info["nb"] = (nb_p, nb_size)
    info["n_clusters"] = unit.Cn
    nh = np.asarray(unit.a["cl_nhap"], np.int64)
    info["haplotype_candidates"] = {"max": int(nh.max()) if len(nh) else 0, "q50": float(np.quantile(nh, 0.5)) if len(nh) else 0, "q99": float(np.quantile(nh, 0.99)) if len(nh) else 0,
                                    "clusters_over_32": int((nh > 32).sum())}
    eng.close(); cd.close()
    if vcf_out is not None:   # GenotypeWriter (include/btgpu_vcf.hpp through host/btvcf): the result arrays are in unit order, so is the description
        from . import vcf_desc
        names = list(sample_names) if sample_names is not None else [f"S{i + 1}" for i in range(S)]
        vcf_desc.write_vcf(vcf_out, res, vcf_desc.describe(inp.chrom, inp.reference, inp.variants, inp.graphs, names), S)
    return (inp.graphs, unit if want_unit else None, res, info)


def _genotype_abi(lib, inp, opt, n_paths, mem, S, spectra_dev, region_buf, gender_ploidy, ploidy, nb_params, noise_rates, info, vcf_out, sample_names, clk=None):
    """The stages after the path search through handles of the C ABI only (what host/btpipeline.cpp does in C++): btg_counter (countPathKmers,
    countInterclusterKmers, parseSampleKmers, classifyPathKmers + getHaplotypeCandidates, NB fit) -> btg_unit -> btg_count_dist -> Gibbs."""
    from . import counter
    clk = clk or _StageClock()
    torch.cuda.synchronize()
    kc = counter.KmerCounter(inp.graphs, n_paths, mem, S, inp.genders)
    try:
        clk.mark("counter create")
        info["n_path_kmers"] = kc.count_path_kmers()
        kc.count_intercluster_kmers(region_buf.data_ptr(), region_buf.numel(), gender_ploidy[0], gender_ploidy[1], False)
        for s, (kd, cdv) in enumerate(spectra_dev):
            kc.parse_sample_kmers(s, kd.data_ptr(), cdv.data_ptr(), cdv.numel())
        clk.mark("k-mer table stages")
        handle = kc.build_unit(ploidy)
        clk.mark("build unit")
        if nb_params is None:
            nb_p, nb_size, used = kc.fit_nb(region_buf.data_ptr(), region_buf.numel(), spectra_dev, gender_ploidy, inp.parameter_kmers, opt.random_seed, opt.max_parameter_kmers)
            info["nb_fit"] = [(m, n, None, None) for m, n in used]
        else:
            nb_p, nb_size = nb_params
        sizes = U.Unit({**{k: np.zeros(0, dt) for k, dt in U._DESC_FIELDS}, "group_cluster_off": inp.graphs["group_cluster_off"], "cl_kmer_off": np.zeros(kc.Cn + 1, np.uint64),
                        "cl_var_off": inp.graphs["cl_var_off"], "var_nalleles": kc._keep["var_nalleles"], "cl_nhap": np.asarray(n_paths, np.uint32)}, S)   # sizes the result arrays only
    finally:
        kc_done = kc
    clk.mark("NB fit")
    cd = engine.CountDistribution(nb_p, nb_size, opt.noise_rate_prior)
    eng = engine.InferenceEngine.from_handle(sizes, handle)
    kc_done.close()
    clk.mark("count distribution + counter free")
    gopts = U.default_opts(seed=opt.random_seed, burn=opt.gibbs_burn_in, samples=opt.gibbs_samples, chains=opt.n_chains, rate=opt.kmer_subsampling_rate,
                           max_hv=opt.max_haplotype_variant_kmers, min_gpp=opt.min_genotype_posterior, min_kmers=opt.min_number_of_kmers,
                           min_frac=None if opt.disable_observed_kmers else U.min_fraction_observed(nb_p, nb_size))
    if opt.noise_genotyping:
        res, info["noise_trace"] = eng.estimate_noise_and_genotypes(cd, gopts, want_trace=False)
    else:
        if noise_rates is None:
            info["noise_trace"] = eng.estimate_noise(cd, gopts, want_trace=False)
        else:
            cd.set_noise_rates(noise_rates)
        clk.mark("estimateNoise")
        res = eng.estimate_genotypes(cd, gopts)
    clk.mark("estimateNoiseAndGenotypes" if opt.noise_genotyping else "estimateGenotypes")
    info["noise_rates"] = cd.noise_rates()
    info["nb"] = (nb_p, nb_size)
    info["n_clusters"] = info["n_clusters_total"] = sizes.Cn
    nh = np.asarray(n_paths, np.int64)
    info["haplotype_candidates"] = {"max": int(nh.max()) if len(nh) else 0, "q50": float(np.quantile(nh, 0.5)) if len(nh) else 0, "q99": float(np.quantile(nh, 0.99)) if len(nh) else 0,
                                    "clusters_over_32": int((nh > 32).sum())}
    eng.close(); cd.close()
    clk.mark("free")
    if vcf_out is not None:
        from . import vcf_desc
        names = list(sample_names) if sample_names is not None else [f"S{i + 1}" for i in range(S)]
        vcf_desc.write_vcf(vcf_out, res, vcf_desc.describe(inp.chrom, inp.reference, inp.variants, inp.graphs, names), S)
    return (inp.graphs, None, res, info)


def run(chrom: str, reference: bytes, variants, spectra, genders, opt: Options | None = None, nb_params=None, noise_rates=None):
    """cluster + genotype for one contig and S samples from host inputs; returns (graphs, unit, result arrays, info)."""
    inp = Inputs(chrom, reference, variants, list(genders), spectra)
    return genotype(inp, opt, nb_params, noise_rates, resident=False, want_unit=True)
