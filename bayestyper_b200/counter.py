"""Host mirror of KmerCounter's genotype-side stage methods (include/bayesTyper/KmerCounter.hpp:61-67) over the btg_counter handle of
the C ABI (csrc/counter.cu): every method is one libbtgpu call.  This is the path a C++ host takes (host/btpipeline.cpp); kmer_pipeline.py
is the older torch-glue implementation of the same stages, kept as the second implementation the tests compare this one with."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .unit import Unit, _DESC_FIELDS

_P = C.c_void_p


class CounterDesc(C.Structure):
    _fields_ = [("n_samples", C.c_uint32), ("n_groups", C.c_uint32), ("n_clusters", C.c_uint32)] + [(n, _P) for n in (
        "sample_gender", "group_cluster_off", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst", "cluster_idx",
        "cl_vertex_off", "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_refvar_off", "v_refvar", "v_nested", "n_paths", "path_mem",
        "cl_var_off", "var_nalleles", "var_dep")]


_DT = {"sample_gender": np.uint8, "group_cluster_off": np.uint64, "group_src_off": np.uint64, "group_src": np.uint32, "group_edge_off": np.uint64,
       "group_edge_src": np.uint32, "group_edge_dst": np.uint32, "cluster_idx": np.uint32, "cl_vertex_off": np.uint64, "v_seq_off": np.uint64, "seq": np.uint8,
       "v_flags": np.uint8, "v_var": np.uint16, "v_allele": np.uint16, "v_refvar_off": np.uint64, "v_refvar": np.uint16, "v_nested": np.uint32,
       "n_paths": np.uint32, "path_mem": np.uint8, "cl_var_off": np.uint64, "var_nalleles": np.uint16, "var_dep": np.uint8}


class KmerCounter:
    """One inference unit: graphs + best paths in, the unit resident in HBM out."""

    def __init__(self, graphs: dict, n_paths, path_mem, n_samples: int, genders):
        self.lib = capi.load()
        self._bind()
        g = graphs
        self.S = n_samples
        a = {k: g[k] for k in ("group_cluster_off", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst", "cluster_idx", "cl_vertex_off",
                               "v_seq_off", "seq", "v_flags", "v_var", "v_allele", "v_refvar_off", "v_refvar", "cl_var_off", "var_dep")}
        a["sample_gender"] = np.array([0 if x in ("F", 0) else 1 for x in genders], np.uint8)
        a["n_paths"] = np.asarray(n_paths)
        a["path_mem"] = np.asarray(path_mem)
        a["var_nalleles"] = (1 + np.asarray(g["var_dep"], np.uint16) + np.asarray(g["var_nalt"], np.uint16)).astype(np.uint16)
        if "v_nested" in g:
            a["v_nested"] = g["v_nested"]
        self._keep = {k: np.ascontiguousarray(v, _DT[k]) for k, v in a.items()}
        d = CounterDesc()
        d.n_samples, d.n_groups, d.n_clusters = n_samples, len(g["group_cluster_off"]) - 1, len(g["cl_vertex_off"]) - 1
        for k, v in self._keep.items():
            setattr(d, k, v.ctypes.data)
        self.G, self.Cn = d.n_groups, d.n_clusters
        self.h = capi.check(self.lib.btg_counter_create(C.addressof(d)), self.lib)

    def _bind(self):
        L = self.lib
        if getattr(L, "_counter_bound", False):
            return
        L.btg_counter_create.restype = _P; L.btg_counter_create.argtypes = [_P]
        L.btg_counter_free.argtypes = [_P]
        L.btg_counter_count_path_kmers.argtypes = [_P, C.POINTER(C.c_uint64)]
        L.btg_counter_count_intercluster_kmers.argtypes = [_P, _P, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32]
        L.btg_counter_parse_sample_kmers.argtypes = [_P, C.c_uint32, _P, _P, C.c_size_t]
        L.btg_counter_build_unit.restype = _P; L.btg_counter_build_unit.argtypes = [_P, _P, _P]
        L.btg_counter_array.restype = C.c_int64; L.btg_counter_array.argtypes = [_P, C.c_char_p, _P, C.c_uint64]
        L.btg_counter_fit_nb.argtypes = [_P, _P, C.c_size_t, C.c_uint32, C.c_uint32, _P, _P, _P, _P, C.c_size_t, C.c_uint32, C.c_uint64, _P, _P, _P, _P]
        L._counter_bound = True

    def count_path_kmers(self) -> int:
        n = C.c_uint64()
        capi.check(self.lib.btg_counter_count_path_kmers(self.h, C.byref(n)), self.lib)
        return n.value

    def count_intercluster_kmers(self, buf_dev_ptr: int, length: int, ploidy_female: int = 2, ploidy_male: int = 2, is_decoy: bool = False):
        capi.check(self.lib.btg_counter_count_intercluster_kmers(self.h, buf_dev_ptr, length, int(is_decoy), ploidy_female, ploidy_male), self.lib)

    def parse_sample_kmers(self, sample_idx: int, kmers_dev_ptr: int, counts_dev_ptr: int, n: int):
        capi.check(self.lib.btg_counter_parse_sample_kmers(self.h, sample_idx, kmers_dev_ptr, counts_dev_ptr, n), self.lib)

    def build_unit(self, group_ploidy=None, multigroup_bloom=None):
        """-> btg_unit handle (resident in HBM); engine.InferenceEngine.from_handle wraps it."""
        pl = np.ascontiguousarray(np.full(self.G * self.S, 2, np.uint8) if group_ploidy is None else group_ploidy, np.uint8)
        self._ploidy = pl
        return capi.check(self.lib.btg_counter_build_unit(self.h, multigroup_bloom, pl.ctypes.data), self.lib)

    def fit_nb(self, buf_dev_ptr: int, length: int, spectra_dev, ploidy=(2, 2), parameter_kmers=None, random_seed: int = 0, max_parameter_kmers: int = 1_000_000):
        """btg_counter_fit_nb: per-sample negative binomial (p, size) from the parameter k-mers; spectra_dev = [(kmers tensor, counts tensor)] in HBM.
        Returns (nb_p, nb_size, [(modal multiplicity, k-mers in the class)])."""
        S = self.S
        kp = (C.c_void_p * S)(*[int(k.data_ptr()) for k, _ in spectra_dev])
        cp = (C.c_void_p * S)(*[int(c.data_ptr()) for _, c in spectra_dev])
        nn = (C.c_size_t * S)(*[int(c.numel()) for _, c in spectra_dev])
        pk = None if parameter_kmers is None else np.ascontiguousarray(parameter_kmers, np.uint64).reshape(-1, 2)
        nb_p, nb_size = np.zeros(S), np.zeros(S)
        modal, n_modal = np.zeros(S, np.uint32), np.zeros(S, np.uint64)
        capi.check(self.lib.btg_counter_fit_nb(self.h, buf_dev_ptr, length, ploidy[0], ploidy[1], kp, cp, nn, None if pk is None else pk.ctypes.data,
                                               0 if pk is None else len(pk), random_seed, max_parameter_kmers, nb_p.ctypes.data, nb_size.ctypes.data,
                                               modal.ctypes.data, n_modal.ctypes.data), self.lib)
        return nb_p, nb_size, [(int(m), int(n)) for m, n in zip(modal, n_modal)]

    def array(self, field: str, dtype) -> np.ndarray:
        n = self.lib.btg_counter_array(self.h, field.encode(), None, 0)
        if n < 0:
            raise capi.BtgError(self.lib.btg_last_error().decode())
        out = np.zeros(int(n), dtype)
        if n:
            capi.check(self.lib.btg_counter_array(self.h, field.encode(), out.ctypes.data, out.nbytes), self.lib)
        return out

    def unit_arrays(self) -> Unit:
        """The descriptor of the last build_unit as a host Unit (tests, fixtures)."""
        dt = dict(_DESC_FIELDS)
        a = {k: self._keep[k] for k in ("sample_gender", "group_cluster_off", "group_src_off", "group_src", "group_edge_off", "group_edge_src", "group_edge_dst", "cluster_idx",
                                        "cl_var_off", "var_nalleles", "var_dep")}
        a["group_ploidy"] = self._ploidy
        a["cl_nhap"] = self._keep["n_paths"].astype(np.uint32)
        for f in ("cl_kmer_off", "cl_mult_off", "mult", "k_has_counts", "k_counts", "k_ic", "k_shared", "cl_uniq_off", "uniq_idx", "cl_multi_off", "multi_idx", "kmer_vh_off",
                  "vh_var", "vh_bits_off", "vh_bits", "cl_hapvar_off", "hap_alleles", "hap_nested_off", "hap_nested", "cl_dep_off", "dep_cluster", "dep_var_off", "dep_var"):
            a[f] = self.array(f, dt[f])
        return Unit(a, self.S)

    def close(self):
        if self.h:
            self.lib.btg_counter_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
