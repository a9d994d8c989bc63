"""Host-side mirror of the C-ABI structs of include/btgpu.h (btg_unit_desc,
btg_gibbs_opts, btg_genotype_result) as ctypes Structures backed by numpy arrays."""
from __future__ import annotations

import ctypes as C

import numpy as np

MAX_SAMPLES = 30
_P = C.c_void_p

_DESC_FIELDS = [
    ("sample_gender", np.uint8), ("group_ploidy", np.uint8), ("group_cluster_off", np.uint64),
    ("group_src_off", np.uint64), ("group_src", np.uint32), ("group_edge_off", np.uint64),
    ("group_edge_src", np.uint32), ("group_edge_dst", np.uint32), ("cluster_idx", np.uint32),
    ("cl_nhap", np.uint32), ("cl_kmer_off", np.uint64), ("cl_var_off", np.uint64), ("cl_mult_off", np.uint64),
    ("mult", np.uint8), ("k_has_counts", np.uint8), ("k_counts", np.uint8), ("k_ic", np.uint8), ("k_shared", np.uint32),
    ("cl_uniq_off", np.uint64), ("uniq_idx", np.uint32), ("cl_multi_off", np.uint64), ("multi_idx", np.uint32),
    ("kmer_vh_off", np.uint64), ("vh_var", np.uint16), ("vh_bits_off", np.uint64), ("vh_bits", np.uint8),
    ("cl_hapvar_off", np.uint64), ("hap_alleles", np.uint16), ("var_nalleles", np.uint16), ("var_dep", np.uint8),
    ("hap_nested_off", np.uint64), ("hap_nested", np.uint32), ("cl_dep_off", np.uint64), ("dep_cluster", np.uint32),
    ("dep_var_off", np.uint64), ("dep_var", np.uint16),
]


class UnitDesc(C.Structure):
    _fields_ = [("n_samples", C.c_uint32), ("n_groups", C.c_uint32), ("n_clusters", C.c_uint32)] + [(n, _P) for n, _ in _DESC_FIELDS]


class GibbsOpts(C.Structure):
    _fields_ = [
        ("random_seed", C.c_uint32), ("gibbs_burn_in", C.c_uint16), ("gibbs_samples", C.c_uint16),
        ("n_chains", C.c_uint16), ("group_index_stride", C.c_uint16), ("kmer_subsampling_rate", C.c_float),
        ("max_haplotype_variant_kmers", C.c_uint32), ("min_genotype_posterior", C.c_float),
        ("min_number_of_kmers", C.c_float), ("min_fraction_observed_kmers", C.c_float * MAX_SAMPLES),
        ("group_index_base", C.c_uint64),
    ]


_RES_FIELDS = [
    ("allele_off", np.uint64), ("geno_off", np.uint64), ("gt", np.uint16), ("gq", np.uint32), ("gpp", np.float32),
    ("app", np.float32), ("nak", np.float32), ("fak", np.float32), ("mac", np.float32), ("saf", np.uint16),
    ("ploidy", np.uint8), ("an", np.uint32), ("valt_off", np.uint64), ("ac", np.uint32), ("af", np.float32),
    ("acp", np.float32), ("anc", np.uint8), ("hc", np.uint16),
]


class GenotypeResult(C.Structure):
    _fields_ = [("n_variants", C.c_uint64)] + [(n, _P) for n, _ in _RES_FIELDS]


def default_opts(seed: int = 20190401, burn: int = 100, samples: int = 250, chains: int = 20, rate: float = 0.1,
                 max_hv: int = 500, min_gpp: float = 0.99, min_kmers: float = 1.0, min_frac=None, group_base: int = 0, group_stride: int = 1) -> GibbsOpts:
    o = GibbsOpts()
    o.random_seed, o.gibbs_burn_in, o.gibbs_samples, o.n_chains = seed, burn, samples, chains
    o.kmer_subsampling_rate, o.max_haplotype_variant_kmers = rate, max_hv
    o.min_genotype_posterior, o.min_number_of_kmers = min_gpp, min_kmers
    for i in range(MAX_SAMPLES):
        o.min_fraction_observed_kmers[i] = 0.0 if min_frac is None or i >= len(min_frac) else float(min_frac[i])
    o.group_index_base = group_base
    o.group_index_stride = group_stride
    return o


def min_fraction_observed(nb_p, nb_size, beta: float = 0.275):
    """Filters ctor (src/bayesTyper/Filters.cpp:42-53): 1 - exp(-0.275 * NB mean) in float."""
    mean = np.asarray(nb_size, np.float64) * (1 - np.asarray(nb_p, np.float64)) / np.asarray(nb_p, np.float64)
    return (1 - np.exp(-(np.float32(beta) * mean))).astype(np.float32)


class Unit:
    """Flat haplotype-candidate descriptors of an inference unit (numpy arrays + ctypes view)."""

    # row-level arrays that btg_unit_upload_dev accepts as device pointers
    DEVICE_FIELDS = ("mult", "k_has_counts", "k_counts", "k_ic", "k_shared", "uniq_idx", "kmer_vh_off", "vh_var", "vh_bits_off", "vh_bits", "hap_alleles")

    def __init__(self, arrays: dict, n_samples: int, dev: dict | None = None):
        """dev: optional {field: torch tensor on the device} for DEVICE_FIELDS (same bit patterns as the numpy dtypes); those
        fields may then be absent from `arrays` and are materialised on the host only when `host()` is called."""
        self.a = {}
        self.dev = dict(dev) if dev else None
        for name, dt in _DESC_FIELDS:
            if self.dev is not None and name in self.dev and name not in arrays:
                continue
            self.a[name] = np.ascontiguousarray(arrays[name], dt)
        self.S = n_samples
        self.G = len(self.a["group_cluster_off"]) - 1
        self.Cn = len(self.a["cl_kmer_off"]) - 1
        self.n_variants = int(self.a["cl_var_off"][-1])

    def desc(self) -> UnitDesc:
        d = UnitDesc()
        d.n_samples, d.n_groups, d.n_clusters = self.S, self.G, self.Cn
        for name, _ in _DESC_FIELDS:
            setattr(d, name, self.a[name].ctypes.data if name in self.a else None)
        return d

    def dev_desc(self):
        """(UnitDesc of device pointers, n_vh, n_vh_bits) for btg_unit_upload_dev, or None for a host-only unit."""
        if not self.dev:
            return None
        d = UnitDesc()
        for name in self.DEVICE_FIELDS:
            if name in self.dev:
                setattr(d, name, self.dev[name].data_ptr())
        return d, int(self.dev["vh_var"].numel()), int(self.dev["vh_bits"].numel())

    def host(self) -> "Unit":
        """Materialise device-resident fields on the host (tests, fixtures, sharding)."""
        if self.dev:
            dt = dict(_DESC_FIELDS)
            for name, t in self.dev.items():
                if name not in self.a:
                    self.a[name] = np.ascontiguousarray(t.cpu().numpy().view(dt[name]).reshape(-1))
        return self

    # ---- sizes of the result arrays -----------------------------------------------------------
    def alloc_result(self):
        nA = self.a["var_nalleles"].astype(np.uint64)
        S = np.uint64(self.S)
        r = {
            "allele_off": np.concatenate([[0], np.cumsum(S * nA)]).astype(np.uint64),
            "geno_off": np.concatenate([[0], np.cumsum(S * nA * (nA + np.uint64(1)) // np.uint64(2))]).astype(np.uint64),
            "valt_off": np.concatenate([[0], np.cumsum(nA)]).astype(np.uint64),
        }
        nv = self.n_variants
        nall, ngen, nalt = int(r["allele_off"][-1]), int(r["geno_off"][-1]), int(r["valt_off"][-1])
        sizes = {"gt": nv * self.S * 2, "gq": nv * self.S, "gpp": ngen, "app": nall, "nak": nall, "fak": nall, "mac": nall,
                 "saf": nall, "ploidy": nv * self.S, "an": nv, "ac": nalt, "af": nalt, "acp": nalt, "anc": nalt, "hc": nv}
        for name, dt in _RES_FIELDS:
            if name not in r:
                r[name] = np.zeros(sizes[name], dt)
        res = GenotypeResult()
        res.n_variants = nv
        for name, _ in _RES_FIELDS:
            setattr(res, name, r[name].ctypes.data)
        return res, r

    def tally_offsets(self):
        H = self.a["cl_nhap"].astype(np.uint64)
        n = (H + np.uint64(1)) * (H + np.uint64(2)) // np.uint64(2) * np.uint64(self.S)
        return np.concatenate([[0], np.cumsum(n)]).astype(np.uint64)

    def subset_groups(self, groups) -> "Unit":
        """A new Unit holding only the given groups (used for shards and bounded CPU samples).  Row-level arrays that live on the
        device (self.dev) are gathered there: nothing crosses the bus."""
        groups = np.asarray(groups, np.int64)
        a = self.a
        dev = self.dev or {}

        def take_csr(off, idx):
            """(new offsets, selected element indices) of rows `idx` of a CSR offset array — numpy or torch (device) alike."""
            if isinstance(off, np.ndarray):
                off = off.astype(np.int64)
                lens = off[idx + 1] - off[idx]
                new_off = np.concatenate([[0], np.cumsum(lens)])
                sel = np.repeat(off[idx] - new_off[:-1], lens) + np.arange(int(new_off[-1]))
                return new_off.astype(np.uint64), sel
            import torch
            idx_t = idx if isinstance(idx, torch.Tensor) else torch.from_numpy(idx).to(off.device)
            lens = off[idx_t + 1] - off[idx_t]
            new_off = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=off.device)
            torch.cumsum(lens, 0, out=new_off[1:])
            total = int(new_off[-1])
            sel = torch.repeat_interleave(off[idx_t] - new_off[:-1], lens, output_size=total) + torch.arange(total, device=off.device)
            return new_off, sel

        def field(name):
            return dev[name] if name in dev and name not in a else a[name]

        def gather(name, sel, width=1):
            x = field(name)
            if isinstance(x, np.ndarray):
                sel_h = sel if isinstance(sel, np.ndarray) else sel.cpu().numpy()
                return x.reshape(-1, width)[sel_h].reshape(-1) if width > 1 else x[sel_h]
            import torch
            sel_t = sel if isinstance(sel, torch.Tensor) else torch.from_numpy(sel).to(x.device)
            return (x.reshape(-1, width)[sel_t].reshape(-1) if width > 1 else x[sel_t]).contiguous()

        S = self.S
        _, clusters = take_csr(a["group_cluster_off"], groups)
        new = {}
        new["sample_gender"] = a["sample_gender"]
        new["group_ploidy"] = a["group_ploidy"].reshape(-1, S)[groups].reshape(-1)
        new["group_cluster_off"], _ = take_csr(a["group_cluster_off"], groups)
        new["group_src_off"], sel = take_csr(a["group_src_off"], groups); new["group_src"] = a["group_src"][sel]
        new["group_edge_off"], sel = take_csr(a["group_edge_off"], groups)
        new["group_edge_src"] = a["group_edge_src"][sel]; new["group_edge_dst"] = a["group_edge_dst"][sel]
        new["cluster_idx"] = a["cluster_idx"][clusters]
        new["cl_nhap"] = a["cl_nhap"][clusters]
        new["cl_kmer_off"], rows = take_csr(a["cl_kmer_off"], clusters)
        new["cl_var_off"], vars_ = take_csr(a["cl_var_off"], clusters)
        new["cl_mult_off"], sel = take_csr(a["cl_mult_off"], clusters); new["mult"] = gather("mult", sel)
        new["k_has_counts"] = gather("k_has_counts", rows)
        new["k_counts"] = gather("k_counts", rows, S)
        new["k_ic"] = gather("k_ic", rows, 2)
        new["k_shared"] = gather("k_shared", rows)
        new["cl_uniq_off"], sel = take_csr(a["cl_uniq_off"], clusters); new["uniq_idx"] = gather("uniq_idx", sel)
        new["cl_multi_off"], sel = take_csr(a["cl_multi_off"], clusters); new["multi_idx"] = a["multi_idx"][sel]
        vh_off = field("kmer_vh_off")
        new["kmer_vh_off"], vh = take_csr(vh_off, rows if isinstance(vh_off, np.ndarray) else rows)
        new["vh_var"] = gather("vh_var", vh)
        new["vh_bits_off"], sel = take_csr(field("vh_bits_off"), vh); new["vh_bits"] = gather("vh_bits", sel)
        new["cl_hapvar_off"], sel = take_csr(a["cl_hapvar_off"], clusters); new["hap_alleles"] = gather("hap_alleles", sel)
        new["var_nalleles"] = a["var_nalleles"][vars_]; new["var_dep"] = a["var_dep"][vars_]
        # haplotype-indexed arrays
        hap_off = np.concatenate([[0], np.cumsum(a["cl_nhap"].astype(np.int64))])
        _, haps = take_csr(hap_off, clusters)
        new["hap_nested_off"], sel = take_csr(a["hap_nested_off"], haps); new["hap_nested"] = a["hap_nested"][sel]
        new["cl_dep_off"], deps = take_csr(a["cl_dep_off"], clusters); new["dep_cluster"] = a["dep_cluster"][deps]
        new["dep_var_off"], sel = take_csr(a["dep_var_off"], deps); new["dep_var"] = a["dep_var"][sel]
        host = {k: v for k, v in new.items() if isinstance(v, np.ndarray)}
        on_dev = {k: v for k, v in new.items() if not isinstance(v, np.ndarray)}
        return Unit(host, S, on_dev or None)


def from_ref_dumps(haps: dict, graphs: dict, genders, ploidy=None) -> Unit:
    """Unit from oracle-R's haps.btd + graphs.btd dumps (tests: same input for all three arms)."""
    S = int(haps["meta"][0])
    G = len(graphs["group_cluster_off"]) - 1
    a = dict(haps)
    a["sample_gender"] = np.array([0 if g in ("F", 0) else 1 for g in genders], np.uint8)
    a["group_ploidy"] = np.full(G * S, 2, np.uint8) if ploidy is None else np.asarray(ploidy, np.uint8)
    for k in ("group_cluster_off", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst", "cluster_idx", "cl_var_off"):
        a[k] = graphs[k]
    a["cl_nhap"] = np.diff(haps["cl_hap_off"]).astype(np.uint32)
    a["var_nalleles"] = (1 + graphs["var_dep"].astype(np.uint16) + graphs["var_nalt"]).astype(np.uint16)
    a["var_dep"] = graphs["var_dep"]
    # multicluster k-mers (KmerCounts::has_multicluster_occ, bit 1 of k_flags): rows of different clusters holding the
    # same k-mer share one KmerCounts record, hence one per-sample multiplicity (KmerCounts.cpp:205-224)
    shared = np.full(len(haps["k_has_counts"]), 0xFFFFFFFF, np.uint32)
    multi = np.flatnonzero((haps["k_flags"] & 2) != 0)
    if len(multi):
        words = haps["kmer_words"].reshape(-1, 2)[multi]
        _, ids = np.unique(words, axis=0, return_inverse=True)
        shared[multi] = ids.reshape(-1).astype(np.uint32)
    a["k_shared"] = shared
    return Unit(a, S)
