"""cluster + genotype for a whole genome (several contigs, decoys, per-contig ploidy) through libbtgpu.

STATUS: staged for round 2 — written when the round's GPU budget was spent, so this composition has NOT run on a GPU yet
(tools/e2e_check.py genome runs it against the reference's multi-contig VCF, tests/golden/vcf_genome_2s.vcf.gz).  Everything it
calls is covered on its own: the unit's graphs (graph_builder.build_genome_graphs, identical to the reference), the device stages
(driver.genotype, the one-contig composition the GPU tests and bench.py run), the VCF description (vcf_desc.describe_genome).

What changes against driver.genotype (one contig):
* intercluster regions are scanned per ploidy class — (female, male) ploidy of the contig, decoy or not — because a k-mer's genomic
  multiplicity is the ploidy-weighted sum over its occurrences (KmerCounter.cpp:297-338, KmerCounts.cpp:95-113);
* the ploidy of a (group, sample) pair follows the group's contig (ChromosomePloidy::getSamplePloidy);
* the negative-binomial fit weighs every occurrence of a parameter k-mer with its contig's ploidy and leaves out k-mers that also
  occur in decoy sequence (main.cpp:306-340: decoy k-mers are entered with `false` and never written).
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np
import torch

from . import capi, engine, graph_builder, kmer_pipeline, ploidy as ploidy_rules, unit as U
from .driver import K, Options, _to_dev, find_variant_cluster_paths


@dataclasses.dataclass
class GenomeInputs:
    genome: dict                       # contig -> sequence bytes, FASTA order, decoy contigs included
    candidates: dict                   # contig -> position-sorted candidates (vcfio.read_candidates)
    genders: list
    spectra: list                      # per sample (kmers (n,2) uint64, counts (n,) uint8)
    decoys: tuple = ()
    blooms: list = None                # per sample (bytes, num_kmers, num_bits) or None = build on the device
    graphs: dict = None

    def prepare(self):
        if self.graphs is None:
            self.graphs = graph_builder.build_genome_graphs(self.genome, self.candidates, self.decoys)
        return self


def region_classes(genome: dict, graphs: dict, chrom_ploidy: ploidy_rules.ChromosomePloidy) -> dict:
    """(female ploidy, male ploidy, is_decoy) -> 'N'-separated buffer of the intercluster regions of that class."""
    names = graphs["contig_names"]
    parts: dict = {}
    sep = np.frombuffer(b"N", np.uint8)
    for c, decoy, a, b in np.asarray(graphs["regions"], np.int64).tolist():
        key = (2, 2, True) if decoy else (*chrom_ploidy.gender_ploidy(names[c]), False)      # decoy sequence carries no ploidy
        parts.setdefault(key, []).extend([np.frombuffer(genome[names[c]], np.uint8)[a:b + 1], sep])
    return {k: np.concatenate(v) for k, v in parts.items()}


def group_ploidy(graphs: dict, chrom_ploidy: ploidy_rules.ChromosomePloidy) -> np.ndarray:
    """(groups x samples) ploidy, row-major: what Unit.group_ploidy holds."""
    names = graphs["contig_names"]
    per_contig = {i: np.array(chrom_ploidy.sample_ploidy(n), np.uint8) for i, n in enumerate(names) if i in set(np.asarray(graphs["group_contig"]).tolist())}
    return np.concatenate([per_contig[int(c)] for c in graphs["group_contig"]]) if len(graphs["group_contig"]) else np.zeros(0, np.uint8)


def _scan_kmers(lib, buf):
    n = buf.numel()
    km = torch.empty((n, 2), dtype=torch.int64, device=buf.device)
    valid = torch.empty(n, dtype=torch.uint8, device=buf.device)
    capi.check(lib.btg_scan_sequence_dev(buf.data_ptr(), n, km.data_ptr(), valid.data_ptr(), None), lib)
    return km[valid.to(torch.bool)]


def estimate_nb_parameters_classes(pipe, class_bufs: dict, spectra_dev, genders, opt: Options):
    """driver.estimate_nb_parameters for several ploidy classes: class_bufs = region_classes() moved to the device."""
    lib, dev = pipe.lib, pipe.dev
    with torch.cuda.stream(pipe.ext):
        kms, wf, wm = [], [], []
        decoy_km = None
        total = 0
        for (pf, pm, decoy), buf in class_bufs.items():
            km = _scan_kmers(lib, buf)
            if decoy:
                decoy_km = km
                continue
            total += int(km.shape[0])
            kms.append(km)
            wf.append(torch.full((km.shape[0],), pf, dtype=torch.int64, device=dev))
            wm.append(torch.full((km.shape[0],), pm, dtype=torch.int64, device=dev))
        km = torch.cat(kms).contiguous()
        wf, wm = torch.cat(wf), torch.cat(wm)
        n_valid = km.shape[0]
        k_lo = torch.empty(n_valid, dtype=torch.int64, device=dev)
        k_hi = torch.empty(n_valid, dtype=torch.int64, device=dev)
        capi.check(lib.btg_table_keys_from_kmers_dev(km.data_ptr(), n_valid, k_lo.data_ptr(), k_hi.data_ptr(), None), lib)
        o = torch.sort(k_lo, stable=True).indices
        o = o[torch.sort(k_hi[o], stable=True).indices]
        s_lo, s_hi = k_lo[o], k_hi[o]
        new = torch.ones(n_valid, dtype=torch.bool, device=dev)
        new[1:] = (s_lo[1:] != s_lo[:-1]) | (s_hi[1:] != s_hi[:-1])
        seg = torch.cumsum(new.to(torch.int64), 0) - 1
        n_keys = int(seg[-1]) + 1 if n_valid else 0
        occ_f = torch.zeros(n_keys, dtype=torch.int64, device=dev).index_add_(0, seg, wf[o])          # ploidy-weighted multiplicity, females
        occ_m = torch.zeros(n_keys, dtype=torch.int64, device=dev).index_add_(0, seg, wm[o])          # ... males
        keys = km[o][new].contiguous()
        kw0, kw1 = s_lo[new].contiguous(), s_hi[new].contiguous()
        idx = torch.empty(n_keys, dtype=torch.int64, device=dev)
        pipe.use_index()
        capi.check(lib.btg_table_lookup_dev(pipe.kw0.data_ptr(), pipe.kw1.data_ptr(), pipe.n_keys, keys.data_ptr(), n_keys, idx.data_ptr(), None), lib)
        keep = idx < 0                                                                              # not a path k-mer
        pipe.use_index(False)
        if decoy_km is not None and decoy_km.shape[0]:                                              # nor a decoy k-mer
            hit = torch.empty(decoy_km.shape[0], dtype=torch.int64, device=dev)
            dk = decoy_km.contiguous()
            capi.check(lib.btg_table_lookup_dev(kw0.data_ptr(), kw1.data_ptr(), n_keys, dk.data_ptr(), dk.shape[0], hit.data_ptr(), None), lib)
            hit = hit[hit >= 0]
            keep[hit] = False
        kw0, kw1, occ_f, occ_m = kw0[keep], kw1[keep], occ_f[keep], occ_m[keep]
        frac = min(1.0, 3.0 * opt.max_parameter_kmers / max(total, 1))
        g = torch.Generator(device=dev).manual_seed(opt.random_seed)
        sel = torch.rand(len(occ_f), device=dev, generator=g) < frac
        kw0, kw1, occ_f, occ_m = kw0[sel], kw1[sel], occ_f[sel], occ_m[sel]
        if len(occ_f) > opt.max_parameter_kmers:
            perm = torch.randperm(len(occ_f), device=dev, generator=g)[:opt.max_parameter_kmers].sort().values
            kw0, kw1, occ_f, occ_m = kw0[perm], kw1[perm], occ_f[perm], occ_m[perm]
        kw0, kw1 = kw0.contiguous(), kw1.contiguous()
        S = len(spectra_dev)
        counts = torch.zeros((len(occ_f), S), dtype=torch.uint8, device=dev)
        rec = torch.zeros(len(occ_f), dtype=torch.uint8, device=dev)
        for si, (kd, cd_) in enumerate(spectra_dev):
            capi.check(lib.btg_table_add_sample_kmers_dev(kw0.data_ptr(), kw1.data_ptr(), len(occ_f), kd.data_ptr(), cd_.data_ptr(), cd_.numel(), S, si,
                                                          counts.data_ptr(), rec.data_ptr(), None), lib)
        nb_p, nb_size, used = [], [], []
        for si in range(S):
            mult = torch.clamp(occ_f if genders[si] in ("F", 0) else occ_m, max=255)
            hist = torch.bincount(mult, minlength=256)[1:33]
            m = int(torch.argmax(hist)) + 1
            c = counts[mult == m, si].to(torch.float64)
            mean = float(c.mean()); var = float(c.var(unbiased=True))
            p, size = C.c_double(), C.c_double()
            lib.btg_nb_moments_to_parameters(mean, var, m, C.byref(p), C.byref(size))
            nb_p.append(p.value); nb_size.append(size.value); used.append((m, int(hist[m - 1]), mean, var))
    pipe.ext.synchronize()
    return np.array(nb_p), np.array(nb_size), used


def genotype_genome(inp: GenomeInputs, opt: Options | None = None, nb_params=None, noise_rates=None, vcf_out=None, sample_names=None):
    """driver.genotype for a genome; returns (graphs, result arrays in unit order, info)."""
    opt = opt or Options()
    lib = capi.load()
    inp.prepare()
    graphs = inp.graphs
    S = len(inp.spectra)
    dev = torch.device("cuda", torch.cuda.current_device())
    info = {}
    non_decoy = [n for n in graphs["contig_names"] if n not in set(inp.decoys)]
    chrom_ploidy = ploidy_rules.ChromosomePloidy(non_decoy, inp.genders, opt.chromosome_ploidy_file)
    spectra_dev = [(_to_dev(k, np.int64, dev), _to_dev(c, np.uint8, dev)) for k, c in inp.spectra]
    class_bufs = {k: torch.from_numpy(v).to(dev) for k, v in region_classes(inp.genome, graphs, chrom_ploidy).items()}
    torch.cuda.synchronize()
    if inp.blooms is not None:
        blooms = [capi.check(lib.btg_bloom_from_bytes(capi.ptr(b), nk, nb, K), lib) for b, nk, nb in inp.blooms]
    else:
        blooms = []
        for kd, _ in spectra_dev:
            b = capi.check(lib.btg_bloom_create(kd.shape[0], opt.bloom_fpr, K), lib)
            capi.check(lib.btg_bloom_insert_dev(b, kd.data_ptr(), kd.shape[0], None), lib)
            blooms.append(b)
    n_paths, mem = find_variant_cluster_paths(lib, graphs, blooms, opt)
    for b in blooms:
        lib.btg_bloom_free(b)
    pipe = kmer_pipeline.KmerPipeline(graphs, n_paths, mem, S, inp.genders)
    info["n_path_kmers"] = pipe.enumerate_path_kmers()
    for (pf, pm, decoy), buf in class_bufs.items():
        pipe.scan_buffer(buf, pf, pm, decoy)
    for s, (kd, cdv) in enumerate(spectra_dev):
        pipe.add_sample(s, kd, cdv)
    unit = pipe.build_unit(multigroup_bloom=None, ploidy=group_ploidy(graphs, chrom_ploidy), device_resident=True)
    if nb_params is None:
        nb_p, nb_size, info["nb_fit"] = estimate_nb_parameters_classes(pipe, class_bufs, spectra_dev, inp.genders, opt)
    else:
        nb_p, nb_size = nb_params
    cd = engine.CountDistribution(nb_p, nb_size, opt.noise_rate_prior)
    eng = engine.InferenceEngine(unit)
    gopts = U.default_opts(seed=opt.random_seed, burn=opt.gibbs_burn_in, samples=opt.gibbs_samples, chains=opt.n_chains, rate=opt.kmer_subsampling_rate,
                           max_hv=opt.max_haplotype_variant_kmers, min_gpp=opt.min_genotype_posterior, min_kmers=opt.min_number_of_kmers,
                           min_frac=None if opt.disable_observed_kmers else U.min_fraction_observed(nb_p, nb_size))
    if noise_rates is None:
        info["noise_trace"] = eng.estimate_noise(cd, gopts, want_trace=False)
    else:
        cd.set_noise_rates(noise_rates)
    info["noise_rates"] = cd.noise_rates()
    info["nb"] = (nb_p, nb_size)
    res = eng.estimate_genotypes(cd, gopts)
    info["n_clusters"] = unit.Cn
    eng.close(); cd.close()
    if vcf_out is not None:
        from . import vcf_desc
        names = list(sample_names) if sample_names is not None else [f"S{i + 1}" for i in range(S)]
        vcf_desc.write_vcf(vcf_out, res, vcf_desc.describe_genome(inp.genome, inp.candidates, graphs, names, inp.decoys), S)
    return graphs, res, info
