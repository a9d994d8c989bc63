"""ctypes prototypes for the graph / table / Gibbs entry points of include/btgpu.h."""
from __future__ import annotations

import ctypes as C

vp = C.c_void_p


def bind(L):
    L.btg_count_dist_create.restype = vp
    L.btg_count_dist_create.argtypes = [C.c_uint32, vp, vp, C.c_float, C.c_float]
    L.btg_nb_moments_to_parameters.restype = None
    L.btg_nb_moments_to_parameters.argtypes = [C.c_double, C.c_double, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.btg_count_dist_set_noise_rates.argtypes = [vp, vp]
    L.btg_count_dist_get_noise_rates.argtypes = [vp, vp]
    L.btg_count_dist_tables.argtypes = [vp, vp, vp]
    L.btg_count_dist_free.argtypes = [vp]
    L.btg_unit_upload.restype = vp
    L.btg_unit_upload.argtypes = [vp]
    L.btg_unit_upload_dev.restype = vp
    L.btg_unit_upload_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64]
    L.btg_unit_free.argtypes = [vp]
    L.btg_estimate_genotypes.argtypes = [vp, vp, vp, vp]
    L.btg_estimate_noise.argtypes = [vp, vp, vp, vp]
    L.btg_unit_cluster_tally.argtypes = [vp, C.c_uint32, vp, C.c_uint64]
    L.btg_estimate_genotypes_async.argtypes = [vp, vp, vp, vp]
    L.btg_unit_download_result.argtypes = [vp, vp, vp]
    L.btg_get_stream.restype = vp
    L.btg_graphs_upload.restype = vp
    L.btg_graphs_upload.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.btg_graphs_free.argtypes = [vp]
    L.btg_graphs_reset.argtypes = [vp]
    L.btg_graphs_path_stats.argtypes = [vp, vp]
    L.btg_find_sample_paths.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]
    L.btg_find_sample_paths_batch.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.btg_get_best_paths.argtypes = [vp, vp, vp, vp, C.c_uint64]
    L.btg_walk_paths_dev.argtypes = [vp, C.c_int] + [vp] * 11 + [vp]
    L.btg_path_alleles_dev.argtypes = [vp, vp, vp, vp, vp, vp]
    L.btg_table_lookup_dev.argtypes = [vp, vp, C.c_int64, vp, C.c_size_t, vp, vp]
    L.btg_table_add_sample_kmers_dev.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.btg_table_scan_region_dev.argtypes = [vp, vp, C.c_int64, vp, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]
    L.btg_table_set_index_dev.argtypes = [vp, C.c_int]
    L.btg_table_keys_from_kmers_dev.argtypes = [vp, C.c_size_t, vp, vp, vp]
    L.btg_table_keys_to_kmers_dev.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.btg_estimate_noise_and_genotypes.argtypes = [vp, vp, vp, vp, vp]
    L.btg_comm_create.restype = vp
    L.btg_comm_create.argtypes = [C.c_uint32, C.c_uint32, vp]
    L.btg_comm_connect.argtypes = [vp, vp]
    L.btg_comm_free.argtypes = [vp]
    L.btg_comm_free.restype = None
    L.btg_estimate_noise_sharded.argtypes = [vp, vp, vp, vp, vp]
    L.btg_estimate_noise_and_genotypes_sharded.argtypes = [vp, vp, vp, vp, vp, vp]
    L.btg_estimate_noise_chains.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp]
    L.btg_count_dist_finish_noise.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
