"""Genotype-side k-mer stages on the device: the host mirror of KmerCounter's
countPathKmers / countInterclusterKmers / parseSampleKmers / classifyPathKmers and of
VariantClusterGraph::getHaplotypeCandidates (include/bayesTyper/KmerCounter.hpp:61-67,
src/bayesTyper/VariantClusterGraph.cpp:800-1135).

The CUDA kernels (csrc/table.cu) do the per-nucleotide / per-record work: rolling canonical
k-mers + ntHash along every best path, the 17 B/record sample stream and the genome scan probing
the exact path-k-mer table.  The relational glue between them (sort, run detection, prefix sums,
scatter into CSR) is torch tensor ops on the same device — plumbing, no per-k-mer Python.
There is no CPU path: every tensor lives in HBM.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import capi
from .unit import Unit

K = 55


class PathWalkDesc(C.Structure):
    _fields_ = [("n_clusters", C.c_uint32), ("n_paths", C.c_uint64)] + [(n, C.c_void_p) for n in (
        "cl_vertex_off", "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_refvar_off", "v_refvar",
        "cl_path_off", "path_mem_off", "path_mem", "path_cluster")]


def _dev(a, dtype, device):
    return torch.from_numpy(np.ascontiguousarray(a).view(dtype) if np.asarray(a).dtype != dtype else np.ascontiguousarray(a)).to(device)


def _i64(a, device):
    """uint64 numpy -> int64 torch (bit pattern preserved)."""
    return torch.from_numpy(np.ascontiguousarray(a, np.uint64).view(np.int64)).to(device)


def _excl_cumsum(x):
    out = torch.zeros(x.numel() + 1, dtype=torch.int64, device=x.device)
    torch.cumsum(x.to(torch.int64), 0, out=out[1:])
    return out


def _on_library_stream(fn):
    """Run a method with the library's stream as torch's current stream (kernels and tensor glue must be ordered
    on ONE stream; the library stream is non-blocking, so the legacy default stream would race with it)."""
    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        torch.cuda.current_stream(self.dev).synchronize()      # inputs prepared by the caller
        with torch.cuda.stream(self.ext):
            out = fn(self, *a, **kw)
        self.ext.synchronize()
        return out
    return wrapper


class KmerPipeline:
    """One inference unit: graphs + best paths in, flat haplotype-candidate descriptors (Unit) out."""

    def __init__(self, graphs: dict, n_paths: np.ndarray, path_mem: np.ndarray, n_samples: int, genders, device="cuda"):
        self.lib = capi.load()
        self.dev = torch.device(device)
        self.ext = torch.cuda.ExternalStream(self.lib.btg_get_stream(), device=self.dev)
        self.S = n_samples
        self.genders = list(genders)
        g = graphs
        d = self.dev
        self.g = g
        self.C = len(g["cl_vertex_off"]) - 1
        self.t = {
            "cl_vertex_off": _i64(g["cl_vertex_off"], d), "v_seq_off": _i64(g["v_seq_off"], d),
            "seq": torch.from_numpy(np.ascontiguousarray(g["seq"], np.uint8)).to(d),
            "v_flags": torch.from_numpy(np.ascontiguousarray(g["v_flags"], np.uint8)).to(d),
            "v_var": torch.from_numpy(np.ascontiguousarray(g["v_var"], np.uint16).view(np.int16)).to(d),
            "v_allele": torch.from_numpy(np.ascontiguousarray(g["v_allele"], np.uint16).view(np.int16)).to(d),
            "v_refvar_off": _i64(g["v_refvar_off"], d),
            "v_refvar": torch.from_numpy(np.ascontiguousarray(g["v_refvar"], np.uint16).view(np.int16)).to(d) if len(g["v_refvar"]) else torch.zeros(1, dtype=torch.int16, device=d),
        }
        n_paths = np.ascontiguousarray(n_paths, np.int64)
        V = np.diff(g["cl_vertex_off"]).astype(np.int64)
        self.n_paths = n_paths
        self.P = int(n_paths.sum())
        cl_path_off = np.concatenate([[0], np.cumsum(n_paths)]).astype(np.int64)
        path_mem_off = np.concatenate([[0], np.cumsum(n_paths * V)]).astype(np.int64)
        assert path_mem_off[-1] == len(path_mem)
        self.cl_path_off = cl_path_off
        self.path_mem_host, self.path_mem_off_host = np.asarray(path_mem, np.uint8), path_mem_off
        self.t["cl_path_off"] = torch.from_numpy(cl_path_off).to(d)
        self.t["path_mem_off"] = torch.from_numpy(path_mem_off).to(d)
        self.t["path_mem"] = torch.from_numpy(np.ascontiguousarray(path_mem, np.uint8)).to(d)
        path_cluster = np.repeat(np.arange(self.C, dtype=np.int32), n_paths)
        self.t["path_cluster"] = torch.from_numpy(path_cluster).to(d)
        gco = np.asarray(g["group_cluster_off"], np.int64)
        self.cl_group = torch.from_numpy(np.repeat(np.arange(len(gco) - 1, dtype=np.int64), np.diff(gco))).to(d)
        self.desc = PathWalkDesc()
        self.desc.n_clusters = self.C
        self.desc.n_paths = self.P
        for name in ("cl_vertex_off", "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_refvar_off", "v_refvar", "cl_path_off", "path_mem_off", "path_mem", "path_cluster"):
            setattr(self.desc, name, self.t[name].data_ptr())
        self.stream = None      # NULL = the library stream (== self.ext)
        torch.cuda.synchronize(d)

    # ---- countPathKmers: enumerate, sort, deduplicate -> table keys -------------------------------------------
    @_on_library_stream
    def enumerate_path_kmers(self):
        lib, d, P = self.lib, self.dev, self.P
        n_occ = torch.zeros(P, dtype=torch.int32, device=d)
        n_cov = torch.zeros(P, dtype=torch.int32, device=d)
        status = torch.zeros(self.C, dtype=torch.int32, device=d)
        capi.check(lib.btg_walk_paths_dev(C.addressof(self.desc), 0, n_occ.data_ptr(), n_cov.data_ptr(), None, None, None, None, None, None, None, None,
                                          status.data_ptr(), self.stream), lib)
        occ_off = _excl_cumsum(n_occ)
        cov_off = _excl_cumsum(n_cov)
        N, NC = int(occ_off[-1]), int(cov_off[-1])
        self.w0 = torch.empty(N, dtype=torch.int64, device=d)
        self.w1 = torch.empty(N, dtype=torch.int64, device=d)
        self.occ_path = torch.empty(N, dtype=torch.int32, device=d)
        self.occ_nt = torch.empty(N, dtype=torch.int32, device=d)
        self.cov_occ = torch.empty(max(NC, 1), dtype=torch.int64, device=d)
        self.cov_var = torch.empty(max(NC, 1), dtype=torch.int16, device=d)
        capi.check(lib.btg_walk_paths_dev(C.addressof(self.desc), 1, None, None, occ_off.data_ptr(), cov_off.data_ptr(), self.w0.data_ptr(), self.w1.data_ptr(),
                                          self.occ_path.data_ptr(), self.occ_nt.data_ptr(), self.cov_occ.data_ptr(), self.cov_var.data_ptr(), status.data_ptr(),
                                          self.stream), lib)
        if int(status.max()) != 0:
            raise capi.BtgError("path walk: running-variant capacity exceeded in cluster %d" % int(torch.nonzero(status)[0]))
        self.NC = NC
        # distinct keys, ascending (key_hi, key_lo) = lexicographic k-mer order (the order of a KMC database)
        o = torch.sort(self.w0, stable=True).indices
        o = o[torch.sort(self.w1[o], stable=True).indices]
        s0, s1 = self.w0[o], self.w1[o]
        new = torch.ones(N, dtype=torch.bool, device=d)
        if N > 1:
            new[1:] = (s0[1:] != s0[:-1]) | (s1[1:] != s1[:-1])
        key_of_sorted = torch.cumsum(new.to(torch.int64), 0) - 1
        self.kw0, self.kw1 = s0[new].contiguous(), s1[new].contiguous()
        self.n_keys = int(self.kw0.numel())
        self.occ_key = torch.empty(N, dtype=torch.int64, device=d)
        self.occ_key[o] = key_of_sorted
        # prefix index over key_hi (46 bits = the first 23 nucleotides): ~1 key per bucket
        self.lut_bits = int(min(24, max(8, np.ceil(np.log2(max(self.n_keys, 2))))))
        buckets = self.kw1 >> (2 * K - 64 - self.lut_bits)
        cnt = torch.bincount(buckets, minlength=1 << self.lut_bits)
        self.lut = _excl_cumsum(cnt)
        S = self.S
        self.counts = torch.zeros((self.n_keys, S), dtype=torch.uint8, device=d)
        self.ic = torch.zeros((self.n_keys, 2), dtype=torch.uint8, device=d)
        self.max_mult = torch.zeros(self.n_keys + 4, dtype=torch.uint8, device=d)
        self.decoy = torch.zeros(self.n_keys, dtype=torch.uint8, device=d)
        self.has_record = torch.zeros(self.n_keys, dtype=torch.uint8, device=d)
        return self.n_keys

    def use_index(self, on: bool = True):
        """(Re)install this table's prefix index in the library (the index is per-process state)."""
        capi.check(self.lib.btg_table_set_index_dev(self.lut.data_ptr() if on else None, self.lut_bits if on else 0), self.lib)

    # ---- countInterclusterKmers -------------------------------------------------------------------------------
    @_on_library_stream
    def scan_regions(self, sequence: bytes, regions, ploidy_female: int = 2, ploidy_male: int = 2, is_decoy: bool = False):
        """regions: iterable of (start, end) inclusive on `sequence` (intercluster_regions.txt.gz rows of one contig)."""
        seq = np.frombuffer(sequence, np.uint8)
        parts = []
        for a, b in regions:
            parts.append(seq[a:b + 1]); parts.append(np.frombuffer(b"N", np.uint8))
        if not parts:
            return
        buf = torch.from_numpy(np.concatenate(parts)).to(self.dev)
        self._scan(buf, ploidy_female, ploidy_male, is_decoy)

    @_on_library_stream
    def scan_buffer(self, buf: torch.Tensor, ploidy_female: int = 2, ploidy_male: int = 2, is_decoy: bool = False):
        """Same, on a resident buffer of regions separated by 'N'."""
        if buf.numel():
            self._scan(buf, ploidy_female, ploidy_male, is_decoy)

    def _scan(self, buf, ploidy_female, ploidy_male, is_decoy):
        self.use_index()
        capi.check(self.lib.btg_table_scan_region_dev(self.kw0.data_ptr(), self.kw1.data_ptr(), self.n_keys, buf.data_ptr(), buf.numel(), int(is_decoy),
                                                      ploidy_female, ploidy_male, self.ic.data_ptr(), self.max_mult.data_ptr(), self.decoy.data_ptr(),
                                                      self.has_record.data_ptr(), self.stream), self.lib)

    # ---- parseSampleKmers -------------------------------------------------------------------------------------
    @_on_library_stream
    def add_sample(self, sample_idx: int, kmers_dev: torch.Tensor, counts_dev: torch.Tensor):
        self.use_index()
        capi.check(self.lib.btg_table_add_sample_kmers_dev(self.kw0.data_ptr(), self.kw1.data_ptr(), self.n_keys, kmers_dev.data_ptr(), counts_dev.data_ptr(),
                                                           counts_dev.numel(), self.S, sample_idx, self.counts.data_ptr(), self.has_record.data_ptr(),
                                                           self.stream), self.lib)

    def _nested_tables(self):
        """HaplotypeInfo::nested_variant_cluster_indices and nested_variant_cluster_dependency
        (VariantClusterGraph.cpp:1006-1010, 1112-1133).  Only clusters holding a vertex that stands for a nested
        cluster contribute, so this is a short host loop over those clusters."""
        g = self.g
        v_nested = np.asarray(g["v_nested"], np.uint32) if "v_nested" in g else np.zeros(0, np.uint32)
        cvo = np.asarray(g["cl_vertex_off"], np.int64)
        hap_nested_off = np.zeros(self.P + 1, np.uint64)
        cl_dep_off = np.zeros(self.C + 1, np.uint64)
        nested_parts, dep_cluster, dep_var_off, dep_var = [], [], [0], []
        marked = np.flatnonzero(v_nested != 0xFFFFFFFF)
        if len(marked):
            path_start = np.concatenate([[0], np.cumsum(self.n_paths.astype(np.int64))])
            hap_len = np.zeros(self.P, np.int64)
            per_hap = {}
            for c in np.unique(np.searchsorted(cvo, marked, side="right") - 1):
                v0, v1 = cvo[c], cvo[c + 1]
                nv = v1 - v0
                bits = self.path_mem_host[self.path_mem_off_host[c]:self.path_mem_off_host[c + 1]].reshape(-1, nv)
                nest = v_nested[v0:v1]
                has = nest != 0xFFFFFFFF
                for p in range(bits.shape[0]):
                    lst = np.sort(nest[has & (bits[p] != 0)])
                    per_hap[path_start[c] + p] = lst
                    hap_len[path_start[c] + p] = len(lst)
                deps = {}
                for v in np.flatnonzero(has):
                    vars_ = []
                    if g["v_var"][v0 + v] != 0xFFFF:
                        vars_.append(int(g["v_var"][v0 + v]))
                    vars_ += [int(x) for x in g["v_refvar"][g["v_refvar_off"][v0 + v]:g["v_refvar_off"][v0 + v + 1]]]
                    deps[int(nest[v])] = sorted(vars_, reverse=True)
                for key in sorted(deps):
                    dep_cluster.append(key); dep_var += deps[key]; dep_var_off.append(len(dep_var))
                cl_dep_off[c + 1] = len(deps)
            hap_nested_off[1:] = np.cumsum(hap_len)
            nested_parts = [per_hap[h] for h in sorted(per_hap)]
        cl_dep_off = np.cumsum(cl_dep_off).astype(np.uint64)
        hap_nested = np.concatenate(nested_parts).astype(np.uint32) if nested_parts else np.zeros(0, np.uint32)
        return (hap_nested_off, hap_nested, cl_dep_off, np.asarray(dep_cluster, np.uint32), np.asarray(dep_var_off, np.uint64),
                np.asarray(dep_var, np.uint16))

    # ---- classifyPathKmers + getHaplotypeCandidates -----------------------------------------------------------
    @_on_library_stream
    def build_unit(self, multigroup_bloom=None, var_nalleles=None, var_dep=None, ploidy=None, device_resident: bool = False) -> Unit:
        d, g = self.dev, self.g
        N = self.occ_key.numel()
        occ_cluster = self.t["path_cluster"].to(torch.int64)[self.occ_path.to(torch.int64)]
        occ_local_path = self.occ_path.to(torch.int64) - self.t["cl_path_off"][occ_cluster]
        # rows = distinct (cluster, key), multiplicities = occurrences per (row, path)
        pair = occ_cluster * self.n_keys + self.occ_key
        pair_u, pair_inv = torch.unique(pair, return_inverse=True)                 # sorted by (cluster, key)
        R = pair_u.numel()
        row_cluster = pair_u // self.n_keys
        row_key = pair_u % self.n_keys
        maxH = int(self.n_paths.max()) if len(self.n_paths) else 1
        trip = pair_inv * maxH + occ_local_path
        trip_u, trip_cnt = torch.unique(trip, return_counts=True)
        t_row, t_path = trip_u // maxH, trip_u % maxH
        t_cnt = torch.clamp(trip_cnt, max=255)                                       # uchar saturation (VariantClusterGraph.cpp:889-893)
        row_max = torch.zeros(R, dtype=torch.int64, device=d).scatter_reduce(0, t_row, t_cnt, "amax")
        first_occ = torch.full((R,), N, dtype=torch.int64, device=d).scatter_reduce(0, pair_inv, torch.arange(N, device=d), "amin")
        # per key: records, cluster / group occurrence, multiplicity, exclusion (KmerCounts.cpp:93-159)
        nk = self.n_keys
        rec = self.has_record.to(torch.bool) | (torch.zeros(nk, dtype=torch.int64, device=d).scatter_reduce(0, row_key, row_max, "amax") > 127)
        n_cl = torch.zeros(nk, dtype=torch.int64, device=d).scatter_add(0, row_key, torch.ones(R, dtype=torch.int64, device=d))
        sum_max = torch.zeros(nk, dtype=torch.int64, device=d).scatter_add(0, row_key, row_max)
        max_hap = torch.clamp(self.max_mult[:nk].to(torch.int64) + sum_max, max=255)
        multicluster = rec & (n_cl >= 2)
        if multigroup_bloom is not None:
            keys = self.key_kmers()
            hit = torch.zeros(nk, dtype=torch.uint8, device=d)
            capi.check(self.lib.btg_bloom_lookup_dev(multigroup_bloom, keys.data_ptr(), nk, hit.data_ptr(), self.stream), self.lib)
            multigroup = hit.to(torch.bool)
        else:   # exact: the k-mer occurs in more than one group
            gk = torch.unique(self.cl_group[row_cluster] * nk + row_key)
            n_gr = torch.zeros(nk, dtype=torch.int64, device=d).scatter_add(0, gk % nk, torch.ones_like(gk))
            multigroup = n_gr >= 2
        excluded = rec & (self.decoy.to(torch.bool) | (max_hap > 127) | multigroup)
        self.flags = {"rec": rec, "multicluster": multicluster, "multigroup": multigroup & rec, "excluded": excluded, "max_hap": max_hap}
        # kept rows in first-seen order within their cluster (kmer_row_indices, VariantClusterGraph.cpp:1056)
        keep = ~excluded[row_key]
        order = torch.argsort(row_cluster[keep] * (N + 1) + first_occ[keep])
        kept = torch.nonzero(keep).squeeze(1)[order]
        Rk = kept.numel()
        new_row = torch.full((R,), -1, dtype=torch.int64, device=d)
        new_row[kept] = torch.arange(Rk, device=d)
        k_cluster, k_key = row_cluster[kept], row_key[kept]
        cl_rows = torch.zeros(self.C, dtype=torch.int64, device=d).scatter_add(0, k_cluster, torch.ones(Rk, dtype=torch.int64, device=d))
        cl_kmer_off = _excl_cumsum(cl_rows)
        H = torch.from_numpy(self.n_paths).to(d)
        cl_mult_off = _excl_cumsum(cl_rows * H)
        local_row = torch.arange(Rk, device=d) - cl_kmer_off[k_cluster]
        mult = torch.zeros(int(cl_mult_off[-1]), dtype=torch.uint8, device=d)
        tk = new_row[t_row]
        sel = tk >= 0
        tkr = tk[sel]
        mult[cl_mult_off[k_cluster[tkr]] + local_row[tkr] * H[k_cluster[tkr]] + t_path[sel]] = t_cnt[sel].to(torch.uint8)
        is_multi = multicluster[k_key]
        ar = torch.arange(Rk, device=d)
        uniq_rows, multi_rows = ar[~is_multi], ar[is_multi]
        cl_uniq = torch.zeros(self.C, dtype=torch.int64, device=d).scatter_add(0, k_cluster[uniq_rows], torch.ones_like(uniq_rows))
        cl_multi = torch.zeros(self.C, dtype=torch.int64, device=d).scatter_add(0, k_cluster[multi_rows], torch.ones_like(multi_rows))
        # coverage bitmaps: (row, variant) -> haplotypes (variant_haplotype_indices)
        if self.NC:
            c_occ = self.cov_occ[:self.NC]
            c_row = new_row[pair_inv[c_occ]]
            ok = c_row >= 0
            c_row, c_var, c_path = c_row[ok], self.cov_var[:self.NC][ok].to(torch.int64) & 0xFFFF, occ_local_path[c_occ[ok]]
            ev = c_row * 65536 + c_var
            ev_u, ev_inv = torch.unique(ev, return_inverse=True)
            e_row, e_var = ev_u // 65536, ev_u % 65536
            e_H = H[k_cluster[e_row]]
            vh_bits_off = _excl_cumsum(e_H)
            vh_bits = torch.zeros(int(vh_bits_off[-1]), dtype=torch.uint8, device=d)
            vh_bits[vh_bits_off[ev_inv] + c_path] = 1
            kmer_vh = torch.zeros(Rk, dtype=torch.int64, device=d).scatter_add(0, e_row, torch.ones_like(e_row))
        else:
            e_var = torch.zeros(0, dtype=torch.int64, device=d); vh_bits_off = torch.zeros(1, dtype=torch.int64, device=d)
            vh_bits = torch.zeros(0, dtype=torch.uint8, device=d); kmer_vh = torch.zeros(Rk, dtype=torch.int64, device=d)
        # haplotype -> allele
        cl_var_off = _i64(g["cl_var_off"], d)
        nvar = cl_var_off[1:] - cl_var_off[:-1]
        if var_nalleles is None:
            var_nalleles = (1 + np.asarray(g["var_dep"], np.uint16) + np.asarray(g["var_nalt"], np.uint16)).astype(np.uint16)
            var_dep = np.asarray(g["var_dep"], np.uint8)
        vna = torch.from_numpy(np.ascontiguousarray(var_nalleles, np.uint16).view(np.int16)).to(d)
        hapvar_off = _excl_cumsum(H * nvar)
        hap_alleles = torch.zeros(int(hapvar_off[-1]) + 1, dtype=torch.int16, device=d)
        capi.check(self.lib.btg_path_alleles_dev(C.addressof(self.desc), cl_var_off.data_ptr(), vna.data_ptr(), hapvar_off.data_ptr(), hap_alleles.data_ptr(),
                                                 self.stream), self.lib)
        self.ext.synchronize()
        # multicluster k-mers of one group share a KmerCounts record: one id per such key (KmerCounts.cpp:205-224)
        shared_id = torch.cumsum(multicluster.to(torch.int64), 0) - 1
        k_shared = torch.where(is_multi, shared_id[k_key], torch.full_like(k_key, 0xFFFFFFFF))
        hap_nested_off, hap_nested, cl_dep_off, dep_cluster, dep_var_off, dep_var = self._nested_tables()
        G = len(g["group_cluster_off"]) - 1
        cpu = lambda t, dt: t.cpu().numpy().astype(dt) if t.dtype != torch.int16 else t.cpu().numpy().view(np.uint16)
        n_multi = int(multi_rows.numel())
        dev = None
        if device_resident and n_multi == 0:
            # the row-level arrays stay in HBM (btg_unit_upload_dev copies them device to device); bit patterns = the ABI's dtypes
            dev = {
                "mult": mult.contiguous(), "k_has_counts": rec[k_key].to(torch.uint8).contiguous(),
                "k_counts": self.counts[k_key].reshape(-1).contiguous(), "k_ic": self.ic[k_key].reshape(-1).contiguous(),
                "k_shared": k_shared.to(torch.int32).contiguous(), "uniq_idx": local_row[uniq_rows].to(torch.int32).contiguous(),
                "kmer_vh_off": _excl_cumsum(kmer_vh).contiguous(), "vh_var": e_var.to(torch.int16).contiguous(),
                "vh_bits_off": vh_bits_off.contiguous(), "vh_bits": vh_bits.contiguous(), "hap_alleles": hap_alleles[:-1].contiguous(),
            }
            self.ext.synchronize()
        a = {
            "sample_gender": np.array([0 if x in ("F", 0) else 1 for x in self.genders], np.uint8),
            "group_ploidy": np.full(G * self.S, 2, np.uint8) if ploidy is None else np.asarray(ploidy, np.uint8),
            "group_cluster_off": g["group_cluster_off"], "group_src_off": g["group_src_off"], "group_src": g["group_src"],
            "group_edge_off": g["group_edge_off"], "group_edge_src": g["group_edge_src"], "group_edge_dst": g["group_edge_dst"],
            "cluster_idx": g["cluster_idx"], "cl_nhap": self.n_paths.astype(np.uint32),
            "cl_kmer_off": cpu(cl_kmer_off, np.uint64), "cl_var_off": g["cl_var_off"], "cl_mult_off": cpu(cl_mult_off, np.uint64),
            "cl_uniq_off": cpu(_excl_cumsum(cl_uniq), np.uint64),
            "cl_multi_off": cpu(_excl_cumsum(cl_multi), np.uint64), "multi_idx": cpu(local_row[multi_rows], np.uint32),
            "cl_hapvar_off": cpu(hapvar_off, np.uint64),
            "var_nalleles": np.asarray(var_nalleles, np.uint16), "var_dep": np.asarray(var_dep, np.uint8),
            "hap_nested_off": hap_nested_off, "hap_nested": hap_nested,
            "cl_dep_off": cl_dep_off, "dep_cluster": dep_cluster, "dep_var_off": dep_var_off, "dep_var": dep_var,
        }
        if dev is None:
            a.update({
                "mult": cpu(mult, np.uint8), "k_has_counts": cpu(rec[k_key], np.uint8),
                "k_counts": cpu(self.counts[k_key].reshape(-1), np.uint8), "k_ic": cpu(self.ic[k_key].reshape(-1), np.uint8),
                "k_shared": cpu(k_shared, np.uint32), "uniq_idx": cpu(local_row[uniq_rows], np.uint32),
                "kmer_vh_off": cpu(_excl_cumsum(kmer_vh), np.uint64), "vh_var": cpu(e_var, np.uint16),
                "vh_bits_off": cpu(vh_bits_off, np.uint64), "vh_bits": cpu(vh_bits, np.uint8),
                "hap_alleles": cpu(hap_alleles[:-1], np.uint16),
            })
        u = Unit(a, self.S, dev)
        if dev is None:
            u.kmer_words = self.key_kmers(k_key).cpu().numpy().view(np.uint64)
        return u

    def key_kmers(self, idx=None):
        """Table keys (all, or those at `idx`) as packed k-mers in the ABI's boundary layout, (n, 2) int64 on the device."""
        lo, hi = (self.kw0, self.kw1) if idx is None else (self.kw0[idx].contiguous(), self.kw1[idx].contiguous())
        out = torch.empty((lo.numel(), 2), dtype=torch.int64, device=self.dev)
        capi.check(self.lib.btg_table_keys_to_kmers_dev(lo.data_ptr(), hi.data_ptr(), lo.numel(), out.data_ptr(), self.stream), self.lib)
        return out
