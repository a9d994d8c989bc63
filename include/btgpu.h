/*
 * btgpu.h — C ABI of libbtgpu.so: B200 (sm_100a) implementation of BayesTyper's
 * two hot paths (k-mer match, per-cluster Gibbs sampler).
 *
 * The reference (bioinformatics-centre/BayesTyper, /root/reference) has no FFI;
 * its seams are C++ class interfaces.  Each entry point below names the
 * reference interface it replaces (file:line relative to /root/reference).
 * INTEGRATION.md shows the reference-side binding for each.
 *
 * Conventions
 *   - return 0 on success, a negative BTG_E* code on failure; the message is
 *     available from btg_last_error() (thread-local).  Nothing here calls
 *     exit() (the reference does: src/kmerBloom/KmerBloom.cpp:67-71).
 *   - plain pointers and sizes only.  Pointers named *_dev are device pointers
 *     on the library's current device; all others are host pointers owned by
 *     the caller.  Opaque handles own their device memory.
 *   - `stream` arguments are cudaStream_t passed as void* (NULL = the library's
 *     own stream); *_dev entry points are asynchronous on that stream, host
 *     entry points return after their results are in the caller's buffers.
 *   - packed k-mers are 2 x uint64_t each, in the in-memory layout of the
 *     reference's std::bitset<2k> (include/bayesTyper/Kmer.hpp:40-50):
 *     nucleotide i in bits [2i,2i+1] (A=0 C=1 G=2 T=3, include/bayesTyper/
 *     Nucleotide.hpp:39-70); word 0 = nucleotides 0..31.  A bitset<110> can be
 *     memcpy'd into this layout.
 *   - k is fixed at compile time in the reference (-DBT_KMER_SIZE, default 55,
 *     CMakeLists.txt:13); this library is built for BTG_KMER_SIZE = 55 and
 *     rejects any other k.
 */
#ifndef BTGPU_H
#define BTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define BTG_KMER_SIZE 55
#define BTG_MAX_SAMPLES 30 /* src/bayesTyper/main.cpp:72 */

#define BTG_OK 0
#define BTG_EINVAL (-1)  /* bad argument */
#define BTG_ECUDA (-2)   /* CUDA runtime error */
#define BTG_EIO (-3)     /* file error */
#define BTG_ENOMEM (-4)  /* allocation failed */
#define BTG_ESTATE (-5)  /* call sequence error */

/* ---- library ---------------------------------------------------------- */
int btg_init(int device);            /* select device, create the library stream */
void btg_shutdown(void);
const char *btg_last_error(void);
int btg_version(void);
int btg_device_sm_count(void);
/* pinned host memory for callers that want overlapped staging */
void *btg_host_alloc(size_t bytes);
void btg_host_free(void *p);
/* device memory for callers that do not link the CUDA runtime (the *_dev entry points read device pointers); copies are synchronous */
void *btg_device_alloc(size_t bytes);
void btg_device_free(void *p);
int btg_copy_to_device(void *dst_dev, const void *src_host, size_t bytes);
int btg_copy_to_host(void *dst_host, const void *src_dev, size_t bytes);
/* number of kernel launches issued by this library since btg_init / last reset */
/* the library's cudaStream_t (host entry points run on it), for callers that time with CUDA events */
void *btg_get_stream(void);
uint64_t btg_launch_count(void);
void btg_launch_count_reset(void);

/* ---- KmerBloom  (include/kmerBloom/KmerBloom.hpp:48-77) ----------------- */
typedef struct btg_bloom btg_bloom;
typedef struct btg_unit btg_unit;   /* an inference unit resident in HBM (defined with the Gibbs entry points below) */

/* KmerBloom(num_kmers, fpr)            src/kmerBloom/KmerBloom.cpp:53-60 */
btg_bloom *btg_bloom_create(uint64_t num_kmers, float fpr, int k);
/* KmerBloom(prefix): <prefix>.bloomMeta / .bloomData    KmerBloom.cpp:62-89 */
btg_bloom *btg_bloom_load(const char *prefix, int k);
/* same, from memory (the .bloomMeta fields + the raw .bloomData bytes, (num_bits+7)/8 of them; KmerBloom.cpp:62-89) */
btg_bloom *btg_bloom_from_bytes(const uint8_t *data, uint64_t num_kmers, uint64_t num_bits, int k);
/* KmerBloom::save(prefix)              KmerBloom.cpp:148-164 */
int btg_bloom_save(const btg_bloom *b, const char *prefix);
int btg_bloom_info(const btg_bloom *b, uint64_t *num_kmers, uint64_t *num_bits, uint32_t *num_hashes);
/* raw filter bytes, BloomFilter::storeFilter layout (external/ntHash/BloomFilter.hpp:260-264) */
int btg_bloom_download(const btg_bloom *b, uint8_t *out, uint64_t nbytes);
/* KmerBloom::addKmer(bitset)           KmerBloom.cpp:178-182 -> BloomFilter::insertF */
int btg_bloom_insert(btg_bloom *b, const uint64_t *kmers, size_t n);
/* KmerBloom::lookup(bitset) const      KmerBloom.cpp:196-200 -> BloomFilter::containsF */
int btg_bloom_lookup(const btg_bloom *b, const uint64_t *kmers, size_t n, uint8_t *hit);
/* device-resident variants (inputs/outputs already in HBM) */
int btg_bloom_insert_dev(btg_bloom *b, const uint64_t *kmers_dev, size_t n, void *stream);
int btg_bloom_lookup_dev(const btg_bloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev, void *stream);
/* same, additionally writing the number of probes the reference's early-exit
 * loop would have executed per k-mer (BloomFilter.hpp:149-161) — the unit of
 * the algorithmic-bytes model.  probes_dev may be NULL. */
int btg_bloom_lookup_probes_dev(const btg_bloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev,
                                uint8_t *probes_dev, void *stream);
void btg_bloom_free(btg_bloom *b);

/* ---- ThreadedKmerBloom (KmerBloom.hpp:79-108; KmerBloom.cpp:203-286) ----- *
 * 65,536 independent sub-filters selected by NTP64(kmer,k,1029283129)%65536.
 * The reference's per-sub-filter mutex (getKmerLock) exists only to make
 * test-then-set atomic (KmerCounter.cpp:124-139).                            */
typedef struct btg_tbloom btg_tbloom;
btg_tbloom *btg_tbloom_create(uint64_t num_kmers, float fpr, int k);
int btg_tbloom_info(const btg_tbloom *b, uint64_t *sub_kmers, uint64_t *sub_bits, uint32_t *num_hashes);
int btg_tbloom_insert(btg_tbloom *b, const uint64_t *kmers, size_t n);
int btg_tbloom_lookup(const btg_tbloom *b, const uint64_t *kmers, size_t n, uint8_t *hit);
int btg_tbloom_insert_dev(btg_tbloom *b, const uint64_t *kmers_dev, size_t n, void *stream);
int btg_tbloom_lookup_dev(const btg_tbloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev, void *stream);
int btg_tbloom_download(const btg_tbloom *b, uint8_t *out, uint64_t nbytes); /* 65536 * ((sub_bits+7)/8) bytes */
void btg_tbloom_free(btg_tbloom *b);

/* ---- k-mer primitives (Kmer.tpp:182-255; nthash.hpp:262-282) ------------- */
/* NTP64 of each packed k-mer as given (no canonicalisation) */
int btg_kmer_hash(const uint64_t *kmers, size_t n, uint64_t *hash_out);
/* KmerPair::getLexicographicalLowestKmer of each packed k-mer */
int btg_kmer_canonical(const uint64_t *kmers, size_t n, uint64_t *canon_out);
/* Rolling enumeration over an ASCII nucleotide string exactly as KmerPair::move
 * drives it (non-ACGT resets): writes the canonical k-mer ending at every
 * position p >= k-1 whose window is all-ACGT; valid_out[p]=1 there, else 0 and
 * the k-mer slot is zeroed.  kmers_out holds 2*len words.  Follows the scan
 * loops of KmerCounter.cpp:291-334 and VariantClusterGraph.cpp:800-846. */
int btg_scan_sequence(const char *seq, size_t len, uint64_t *kmers_out, uint8_t *valid_out);
int btg_scan_sequence_dev(const char *seq_dev, size_t len, uint64_t *kmers_out_dev, uint8_t *valid_out_dev, void *stream);
/* fused scan + Bloom probe (the a10 genome scan against a filter): hit_out[p]=1
 * iff the window ending at p is valid and its canonical k-mer is in b */
int btg_scan_sequence_lookup_dev(const btg_bloom *b, const char *seq_dev, size_t len, uint8_t *hit_out_dev, void *stream);


/* ======================= candidate-path search (k-mer match, hot loop A) ===================== *
 * Replaces KmerCounter::findVariantClusterPaths (include/bayesTyper/KmerCounter.hpp:57,
 * src/bayesTyper/KmerCounter.cpp:59-103) -> VariantClusterGroup::findSamplePaths ->
 * VariantClusterGraph::findSamplePaths (src/bayesTyper/VariantClusterGraph.cpp:389-798).
 * The graphs are the reference's boost::adjacency_list per cluster, CSR-flattened (vertices are
 * already in topological = index order, VariantClusterGraph.cpp:433).                           */
typedef struct btg_graphs_desc {
    uint32_t n_clusters;
    const uint64_t *cl_vertex_off;   /* [C+1] vertices per cluster */
    const uint64_t *v_seq_off;       /* [V+1] -> seq */
    const uint8_t *seq;              /* VariantClusterGraphVertex::sequence as one code (0..3) per nucleotide */
    const uint8_t *v_flags;          /* [V] bit0 is_first_nucleotides_redundant, bit1 is_disconnected */
    const uint64_t *v_in_off;        /* [V+1] -> v_in_src */
    const uint32_t *v_in_src;        /* in-edge sources (cluster-local vertex ids), in boost::in_edges order */
    const uint32_t *cl_group;        /* [C] index of the cluster's group in the unit (after the size sort) */
    const uint32_t *cl_idx;          /* [C] variant_cluster_idx within its group */
} btg_graphs_desc;

typedef struct btg_graphs btg_graphs;
btg_graphs *btg_graphs_upload(const btg_graphs_desc *desc, uint32_t max_samples, uint32_t max_sample_haplotypes);
/* work done by the path searches since upload / reset: out3 = {k-mer lookups, Bloom probes executed under the reference's early-exit order
 * (BloomFilter.hpp:149-161), nucleotides walked} — the units of the algorithmic-bytes model of the k-mer-match kernel */
int btg_graphs_path_stats(const btg_graphs *g, uint64_t *out3);
/* forget the best paths found so far (a new pass over the samples of the same unit: the graphs and the scratch stay in HBM) */
int btg_graphs_reset(btg_graphs *g);
void btg_graphs_free(btg_graphs *g);
/* one sample's pass (samples must be submitted in order 0..S-1: addPathIndices merges order-dependently,
 * VariantClusterGraph.cpp:726-798).  seed of cluster c: random_seed + (group+1)*(sample_idx+1) + cluster_idx.
 * Asynchronous on the library stream (the passes of a unit queue up behind each other; the Bloom filter must stay alive until
 * btg_get_best_paths or a synchronisation); a scratch overflow is reported by btg_get_best_paths. */
int btg_find_sample_paths(btg_graphs *g, const btg_bloom *sample_bloom, uint32_t sample_idx, uint32_t random_seed,
                          uint32_t max_sample_haplotypes);
/* KmerCounter::findVariantClusterPaths (KmerCounter.cpp:70-103) for the samples sample_first .. sample_first + n_samples - 1 in ONE launch — the
 * reference loops over the samples with one filter resident at a time (main.cpp:219-247); here their filters are resident together: the
 * (cluster, sample) searches run side by side and every cluster merges its samples' paths in sample order, so the best paths equal those of
 * n_samples btg_find_sample_paths calls bit for bit, while the latency of the unit's slowest cluster is paid once instead of once per sample.
 * Batches must be submitted in sample order like single samples; falls back to one launch per sample when n_samples == 1 or the per-warp
 * scratch of the unit's largest clusters does not fit. */
int btg_find_sample_paths_batch(btg_graphs *g, const btg_bloom *const *sample_blooms, uint32_t sample_first, uint32_t n_samples,
                                uint32_t random_seed, uint32_t max_sample_haplotypes);
/* VariantClusterGraph::best_paths_indices (VariantClusterGraph.hpp:98; filled by addPathIndices, VariantClusterGraph.cpp:726-798): n_paths_out[C]; path_off_out[C+1] (prefix sums of n_paths*V, optional); membership_out
 * (optional) one byte per (path, vertex), path-major                                             */
int btg_get_best_paths(const btg_graphs *g, uint32_t *n_paths_out, uint64_t *path_off_out, uint8_t *membership_out,
                       uint64_t membership_bytes);

/* ======================= path k-mers and the exact count table (k-mer match, hot loop B) ===== *
 * Device-pointer entry points (suffix _dev): the host glue keeps graphs, best paths, sample k-mer
 * streams and the table columns resident in HBM and composes these kernels (bayestyper_b200/
 * kmer_pipeline.py mirrors KmerCounter's genotype-side stages: countPathKmers, countInterclusterKmers,
 * parseSampleKmers, classifyPathKmers, include/bayesTyper/KmerCounter.hpp:61-67).
 * Table keys: the distinct path k-mers in lexicographic order of their strings (A<C<G<T from nucleotide 0) — the
 * order of the records of a KMC1 database (external/kmc_api/kmc_file.cpp:428-515; KMC2: of every signature bin), so a
 * sample's stream walks the table front to back.  A key is the k-mer as the 110-bit integer V = sum_i code(nt_i) << 2*(54-i), split into two
 * signed 64-bit columns for sorting: key_hi = V >> 64 (46 bits) and key_lo = (V & (2^64-1)) ^ 2^63; keys ascend
 * by (key_hi, key_lo).  btg_table_keys_{from,to}_kmers_dev convert from / to the ABI's packed k-mers.        */
typedef struct btg_pathwalk_desc {          /* every pointer is a device pointer */
    uint32_t n_clusters;
    uint64_t n_paths;                        /* best paths of all clusters, cluster-major */
    const uint64_t *cl_vertex_off;           /* as btg_graphs_desc */
    const uint64_t *v_seq_off;
    const uint8_t *seq;
    const uint8_t *v_flags;
    const uint16_t *v_var;                   /* [V] VariantClusterGraphVertex::variant_allele_idx.first (0xFFFF none) */
    const uint16_t *v_allele;                /* [V] .second */
    const uint64_t *v_refvar_off;            /* [V+1] -> v_refvar: reference_variant_indices */
    const uint16_t *v_refvar;
    const uint64_t *cl_path_off;             /* [C+1] first path of each cluster */
    const uint64_t *path_mem_off;            /* [C+1] byte offset of the cluster's rows in path_mem */
    const uint8_t *path_mem;                 /* best_paths_indices: one byte per (path, vertex) */
    const uint32_t *path_cluster;            /* [n_paths] */
} btg_pathwalk_desc;

/* Walks every best path (KmerPair reset at disconnected vertices) — the loop shared by countPathKmers,
 * classifyPathKmers and getHaplotypeCandidates (VariantClusterGraph.cpp:800-1135).
 * emit = 0: n_occ[p] / n_cov[p] = number of k-mer windows / (window, covered variant) pairs of path p.
 * emit = 1: with occ_off / cov_off = exclusive prefix sums of those, writes per window the canonical k-mer
 *           as a table key (key_lo, key_hi), its path and nucleotide index, and per pair (window id, variant): the
 *           running_variants coverage of updateVariantPathIndices (:1137-1184).
 * status[c] = 2 if a cluster exceeds the running-variant capacity.                             */
int btg_walk_paths_dev(const btg_pathwalk_desc *d, int emit, uint32_t *n_occ, uint32_t *n_cov, const uint64_t *occ_off,
                       const uint64_t *cov_off, int64_t *key_lo, int64_t *key_hi, uint32_t *occ_path, uint32_t *occ_nt,
                       int64_t *cov_occ, uint16_t *cov_var, uint32_t *status, void *stream);
/* packed k-mers (2 x uint64 each) <-> table key columns */
int btg_table_keys_from_kmers_dev(const uint64_t *kmers, size_t n, int64_t *key_lo, int64_t *key_hi, void *stream);
int btg_table_keys_to_kmers_dev(const int64_t *key_lo, const int64_t *key_hi, size_t n, uint64_t *kmers, void *stream);
/* HaplotypeInfo::variant_allele_indices of every best path (VariantClusterGraph.cpp:983-992,1091-1098) */
int btg_path_alleles_dev(const btg_pathwalk_desc *d, const uint64_t *cl_var_off, const uint16_t *var_nalleles,
                         const uint64_t *hapvar_off, uint16_t *hap_alleles, void *stream);
/* Optional prefix index over the sorted keys (cf. the prefix LUT of a KMC database, external/kmc_api/kmc_file.cpp:236-290):
 * lut[b] = index of the first key whose key_hi top lut_bits bits (of 46; the first lut_bits/2 nucleotides) are >= b,
 * b in [0, 2^lut_bits]; NULL clears.
 * Applies to the three table probes below, on the CALLING HOST THREAD, until changed (thread-local: a probe of another key table
 * must clear or replace it first; a probe without an index is correct, only slower).  btg_counter handles install their own. */
int btg_table_set_index_dev(const int64_t *lut, int lut_bits);
/* KmerCountsHash::findKmer (KmerHash.hpp:82, ObservedKmerCountsHash: KmerHash.cpp:230-243) on a batch: index into the key arrays or -1 */
int btg_table_lookup_dev(const int64_t *key_lo, const int64_t *key_hi, int64_t n_keys, const uint64_t *kmers, size_t n,
                         int64_t *idx_out, void *stream);
/* KmerCounter::parseSampleKmers for one batch of one sample (KmerCounter.cpp:388-429): for every (k-mer, count)
 * record present in the table, counts[idx][sample] saturating-adds count (KmerCounts.cpp:178-189).  Any record order
 * is accepted; the order of a KMC database (= the table's) is the fast one.                                 */
int btg_table_add_sample_kmers_dev(const int64_t *key_lo, const int64_t *key_hi, int64_t n_keys, const uint64_t *kmers,
                                   const uint8_t *counts, size_t n, uint32_t n_samples, uint32_t sample_idx,
                                   uint8_t *table_counts, uint8_t *has_record, void *stream);
/* KmerCounter::countInterclusterKmers for one region (KmerCounter.cpp:291-334): rolling scan, probe,
 * KmerCounts::addInterclusterMultiplicity (KmerCounts.cpp:98-118). ic = [n_keys][2] (female, male)        */
int btg_table_scan_region_dev(const int64_t *key_lo, const int64_t *key_hi, int64_t n_keys, const char *seq, size_t len,
                              int is_decoy, uint32_t ploidy_female, uint32_t ploidy_male, uint8_t *ic, uint8_t *max_mult,
                              uint8_t *decoy, uint8_t *has_record, void *stream);

/* ======================= KmerCounter: the genotype-side k-mer stages behind one handle ======= *
 * Replaces the stage methods KmerCounter::{countPathKmers, countInterclusterKmers, parseSampleKmers, classifyPathKmers}
 * (include/bayesTyper/KmerCounter.hpp:61-67; called from src/bayesTyper/main.cpp:594-603), the count table they share
 * (KmerCountsHash / KmerCounts, include/bayesTyper/KmerHash.hpp:73-108, src/bayesTyper/KmerCounts.cpp:93-189) and
 * VariantClusterGraph::{countPathKmers, classifyPathKmers, getHaplotypeCandidates} (src/bayesTyper/VariantClusterGraph.cpp:
 * 800-1184) for one inference unit.  The handle owns the table (sorted 128-bit keys + count / multiplicity / flag columns)
 * and every intermediate in HBM; the relational steps between the kernels are device sorts / run-length encodings / scans.
 * All arrays of the descriptor are host memory (copied by btg_counter_create).                                            */
typedef struct btg_counter_desc {
    uint32_t n_samples, n_groups, n_clusters;
    const uint8_t *sample_gender;        /* [S] 0 = Female, 1 = Male */
    const uint64_t *group_cluster_off;   /* [G+1] as btg_unit_desc (passed through to the unit) */
    const uint64_t *group_src_off;       /* [G+1] */
    const uint32_t *group_src;
    const uint64_t *group_edge_off;      /* [G+1] */
    const uint32_t *group_edge_src, *group_edge_dst;
    const uint32_t *cluster_idx;         /* [C] */
    const uint64_t *cl_vertex_off;       /* [C+1] graphs as btg_pathwalk_desc, host side */
    const uint64_t *v_seq_off;           /* [V+1] */
    const uint8_t *seq;
    const uint8_t *v_flags;              /* [V] */
    const uint16_t *v_var, *v_allele;    /* [V] variant_allele_idx (0xFFFF none) */
    const uint64_t *v_refvar_off;        /* [V+1] */
    const uint16_t *v_refvar;
    const uint32_t *v_nested;            /* [V] nested_variant_cluster_idx of the vertex or 0xFFFFFFFF; NULL: no nested clusters */
    const uint32_t *n_paths;             /* [C] best paths per cluster (btg_get_best_paths) */
    const uint8_t *path_mem;             /* best_paths_indices, one byte per (path, vertex), cluster-major */
    const uint64_t *cl_var_off;          /* [C+1] */
    const uint16_t *var_nalleles;        /* [n_variants] */
    const uint8_t *var_dep;              /* [n_variants] */
} btg_counter_desc;
typedef struct btg_counter btg_counter;
btg_counter *btg_counter_create(const btg_counter_desc *desc);
void btg_counter_free(btg_counter *k);
/* KmerCounter::countPathKmers (KmerCounter.cpp:252-289): the distinct canonical k-mers of every best path become the table keys */
int btg_counter_count_path_kmers(btg_counter *k, uint64_t *n_path_kmers_out);
/* KmerCounter::countInterclusterKmers (KmerCounter.cpp:291-386) for a device buffer of inter-cluster regions separated by 'N' */
int btg_counter_count_intercluster_kmers(btg_counter *k, const char *seq_dev, size_t len, int is_decoy, uint32_t ploidy_female, uint32_t ploidy_male);
/* KmerCounter::parseSampleKmers (KmerCounter.cpp:431-524) for one sample's (k-mer, count) records resident in HBM */
int btg_counter_parse_sample_kmers(btg_counter *k, uint32_t sample_idx, const uint64_t *kmers_dev, const uint8_t *counts_dev, size_t n);
/* KmerCounter::classifyPathKmers (KmerCounter.cpp:526-600) + getHaplotypeCandidates of every cluster: the inference unit, resident in HBM.
 * multigroup = the cluster stage's multigroup_kmers filter (main.cpp:345-351) or NULL (exact: k-mers that occur in several groups);
 * group_ploidy [G*S] as btg_unit_desc.                                                                                        */
btg_unit *btg_counter_build_unit(btg_counter *k, const btg_bloom *multigroup, const uint8_t *group_ploidy);
/* The negative-binomial fit of every sample from the parameter k-mers: the genotype-side half of countInterclusterParameterKmers +
 * calculateKmerStats + setGenomicCountDistributions (KmerCounter.cpp:171-250, KmerHash.cpp:257-347, CountDistribution.cpp:66-141).
 * seq_dev: the inter-cluster regions ('N'-separated) on the device; sample_*: every sample's (k-mer, count) records in HBM.
 * parameter_kmers (host, n x 2 packed words) = <out>_cluster_data/parameter_kmers.fa.gz of the cluster stage: the fit then uses exactly
 * those k-mers, as `bayesTyper genotype` does (main.cpp:543-584); NULL: a seeded Bernoulli subsample of the inter-cluster k-mers that are
 * not path k-mers, capped at max_parameter_kmers.  Outputs per sample: NB (p, size) of one haploid copy, the modal multiplicity used and
 * the number of k-mers in that class (both optional).  Fails (BTG_ESTATE) when a sample has no modal class (ploidy 0 on the contig).   */
int btg_counter_fit_nb(btg_counter *k, const char *seq_dev, size_t len, uint32_t ploidy_female, uint32_t ploidy_male, const uint64_t *const *sample_kmers_dev,
                       const uint8_t *const *sample_counts_dev, const size_t *sample_n, const uint64_t *parameter_kmers, size_t n_parameter_kmers, uint32_t random_seed,
                       uint64_t max_parameter_kmers, double *nb_p_out, double *nb_size_out, uint32_t *modal_multiplicity_out, uint64_t *n_modal_kmers_out);
/* number of elements (out = NULL) or a host copy of an array of the last btg_counter_build_unit, by its btg_unit_desc field name; also
 * "key_lo" / "key_hi" (table keys) and "key_flags" (bit0 record, bit1 multicluster, bit2 multigroup, bit3 excluded)               */
int64_t btg_counter_array(btg_counter *k, const char *field, void *out, uint64_t out_bytes);

/* ======================= per-cluster Gibbs sampler =========================== *
 * Replaces InferenceEngine::{estimateNoise,estimateGenotypes,estimateNoiseAndGenotypes}
 * (include/bayesTyper/InferenceEngine.hpp:62-64, src/bayesTyper/InferenceEngine.cpp:135-472)
 * and everything below it: VariantClusterGroup::estimateGenotypes, VariantClusterGenotyper,
 * VariantClusterHaplotypes, CountDistribution, (Sparse)FrequencyDistribution,
 * DiscreteSampler, SparsityEstimator, CountAllocation, KmerStats.                */

/* ---- CountDistribution (include/bayesTyper/CountDistribution.hpp:44-90) ------ */
typedef struct btg_count_dist btg_count_dist;
/* CountDistribution ctor + setGenomicCountDistributions' table build
 * (src/bayesTyper/CountDistribution.cpp:51-64,215-238,267-312): per sample the
 * negative-binomial (p, size) of one haploid copy; builds the [S][256][256]
 * genomic log-pmf cache on the device.  prior = --noise-rate-prior (shape, scale). */
btg_count_dist *btg_count_dist_create(uint32_t n_samples, const double *nb_p, const double *nb_size,
                                      float noise_prior_shape, float noise_prior_scale);
/* NegativeBinomialDistribution::momentsToParameters (NegativeBinomialDistribution.cpp:68-79)
 * followed by the division of size by the modal multiplicity (CountDistribution.cpp:115-116) */
void btg_nb_moments_to_parameters(double mean, double var, uint32_t multiplicity, double *p_out, double *size_out);
/* CountDistribution::setNoiseRates (CountDistribution.cpp:153-161): rebuilds the [S][256] Poisson cache */
int btg_count_dist_set_noise_rates(btg_count_dist *cd, const double *noise_rates);
int btg_count_dist_get_noise_rates(const btg_count_dist *cd, double *noise_rates_out);
/* copies of the device tables (tests / reporting): genomic [S][256][256], noise [S][256] */
int btg_count_dist_tables(const btg_count_dist *cd, double *genomic_out, double *noise_out);
void btg_count_dist_free(btg_count_dist *cd);

/* ---- inference unit: flat haplotype-candidate descriptors ---------------------
 * The state a VariantClusterGenotyper holds after its constructor
 * (src/bayesTyper/VariantClusterGenotyper.cpp:59-106): VariantClusterHaplotypes
 * (include/bayesTyper/VariantClusterHaplotypes.hpp:45-112) of every cluster of every
 * VariantClusterGroup, CSR-flattened.  All arrays are host memory owned by the caller
 * and are copied by btg_unit_upload.  Offsets are uint64 prefix sums.            */
typedef struct btg_unit_desc {
    uint32_t n_samples;               /* S <= 30 */
    uint32_t n_groups;                /* G, in unit order (after the size sort, src/bayesTyper/main.cpp:247) */
    uint32_t n_clusters;              /* C: group vertices concatenated in group order */
    const uint8_t *sample_gender;     /* [S] 0 = Female, 1 = Male (Utils::Gender) */
    const uint8_t *group_ploidy;      /* [G*S] Utils::Ploidy (0 Null, 1 Haploid, 2 Diploid) on the group's chromosome */
    const uint64_t *group_cluster_off;/* [G+1] */
    const uint64_t *group_src_off;    /* [G+1] -> group_src: source vertices (local cluster index), VariantClusterGroup.hpp:86 */
    const uint32_t *group_src;
    const uint64_t *group_edge_off;   /* [G+1] -> nested-cluster tree edges (local indices), in out_edges order */
    const uint32_t *group_edge_src;
    const uint32_t *group_edge_dst;
    const uint32_t *cluster_idx;      /* [C] variant_cluster_idx inside the group (seeds; nested lookups) */
    const uint32_t *cl_nhap;          /* [C] number of haplotype candidates H */
    const uint64_t *cl_kmer_off;      /* [C+1] rows = k-mers (KmerInfo) */
    const uint64_t *cl_var_off;       /* [C+1] variants */
    const uint64_t *cl_mult_off;      /* [C+1] -> mult */
    const uint8_t *mult;              /* haplotype_kmer_multiplicities, row-major K x H per cluster */
    const uint8_t *k_has_counts;      /* [rows] KmerInfo::counts != nullptr */
    const uint8_t *k_counts;          /* [rows*S] KmerCounts::getSampleCount */
    const uint8_t *k_ic;              /* [rows*2] getInterclusterMultiplicity(Female), (Male) */
    const uint32_t *k_shared;         /* [rows] group-shared multiplicity record of a multicluster k-mer, else 0xFFFFFFFF */
    const uint64_t *cl_uniq_off;      /* [C+1] -> uniq_idx  (unique_kmer_indices, local row ids) */
    const uint32_t *uniq_idx;
    const uint64_t *cl_multi_off;     /* [C+1] -> multi_idx (multicluster_kmer_indices) */
    const uint32_t *multi_idx;
    const uint64_t *kmer_vh_off;      /* [rows+1] -> vh_var: KmerInfo::variant_haplotype_indices */
    const uint16_t *vh_var;           /* variant index (local) */
    const uint64_t *vh_bits_off;      /* [n_vh+1] -> vh_bits: one byte per haplotype (H of them) */
    const uint8_t *vh_bits;
    const uint64_t *cl_hapvar_off;    /* [C+1] -> hap_alleles */
    const uint16_t *hap_alleles;      /* HaplotypeInfo::variant_allele_indices, H x nvar per cluster */
    const uint16_t *var_nalleles;     /* [n_variants] VariantInfo::numberOfAlleles() (include/bayesTyper/VariantInfo.hpp:77-80) */
    const uint8_t *var_dep;           /* [n_variants] VariantInfo::has_dependency */
    const uint64_t *hap_nested_off;   /* [n_haplotypes+1] -> hap_nested: sorted nested_variant_cluster_indices */
    const uint32_t *hap_nested;
    const uint64_t *cl_dep_off;       /* [C+1] -> dep_cluster: nested_variant_cluster_dependency keys (ascending) */
    const uint32_t *dep_cluster;
    const uint64_t *dep_var_off;      /* [n_dep+1] -> dep_var (descending variant indices) */
    const uint16_t *dep_var;
} btg_unit_desc;

btg_unit *btg_unit_upload(const btg_unit_desc *desc);
/* The same for callers that built the row-level arrays on the device (the k-mer stages do): every non-NULL pointer in `dev`
 * is a DEVICE pointer that replaces the host array of the same name in `desc` (which may then be NULL) — for these fields only:
 * mult, k_has_counts, k_counts, k_ic, k_shared, uniq_idx, kmer_vh_off, vh_var, vh_bits_off, vh_bits, hap_alleles.
 * n_vh / n_vh_bits are the lengths of vh_var / vh_bits (the host cannot read kmer_vh_off[rows] / vh_bits_off[n_vh] then).
 * Units with multicluster k-mers keep k_shared and k_has_counts on the host (they are validated there).  dev = NULL: as above. */
btg_unit *btg_unit_upload_dev(const btg_unit_desc *desc, const btg_unit_desc *dev, uint64_t n_vh, uint64_t n_vh_bits);
void btg_unit_free(btg_unit *u);

/* the options InferenceEngine / Filters read (src/bayesTyper/main.cpp:389-403) */
typedef struct btg_gibbs_opts {
    uint32_t random_seed;                 /* --random-seed */
    uint16_t gibbs_burn_in;               /* 100 */
    uint16_t gibbs_samples;               /* 250 */
    uint16_t n_chains;                    /* 20  */
    uint16_t group_index_stride;          /* group j of this btg_unit is group group_index_base + j * stride of the whole unit (0 = 1): a rank of a
                                             sharded unit takes every world-th group of the size-sorted unit, which balances the shards */
    float kmer_subsampling_rate;          /* 0.1 */
    uint32_t max_haplotype_variant_kmers; /* 500 */
    float min_genotype_posterior;         /* 0.99 (Filters) */
    float min_number_of_kmers;            /* 1 */
    float min_fraction_observed_kmers[BTG_MAX_SAMPLES]; /* Filters.cpp:42-53; 0 when --disable-observed-kmers */
    uint64_t group_index_base;            /* index of this shard's first group in the whole unit (multi-GPU shards keep the reference's per-group seeds) */
} btg_gibbs_opts;

/* per-variant results: the fields of `Genotypes` (include/bayesTyper/Genotypes.hpp:46-99)
 * as flat arrays owned by the caller.  nA = var_nalleles[v]; nG = nA*(nA+1)/2.
 * Blocks are laid out variant-major, then sample: index (off[v] + s*nX + i).      */
typedef struct btg_genotype_result {
    uint64_t n_variants;
    const uint64_t *allele_off;  /* [n_variants+1] prefix sums of S*nA  (app, nak, fak, mac, saf) */
    const uint64_t *geno_off;    /* [n_variants+1] prefix sums of S*nG  (gpp) */
    uint16_t *gt;                /* [n_variants*S*2] genotype_estimate, 0xFFFF = '.', second 0xFFFE = haploid/absent */
    uint32_t *gq;                /* [n_variants*S] */
    float *gpp;                  /* genotype_posteriors (first nA entries used when haploid) */
    float *app;                  /* allele_posteriors */
    float *nak, *fak, *mac;      /* AlleleKmerStats means, -1 when empty */
    uint16_t *saf;               /* allele_filters */
    uint8_t *ploidy;             /* [n_variants*S] ploidy the sample was genotyped with */
    uint32_t *an;                /* [n_variants] VariantStats::total_count */
    const uint64_t *valt_off;    /* [n_variants+1] prefix sums of nA (acp, anc) ; ac/af use entries 1.. */
    uint32_t *ac;                /* [sum nA] alt_allele_counts at index a (a>=1) */
    float *af;                   /* [sum nA] alt_allele_frequency */
    float *acp;                  /* [sum nA] allele_call_probabilities */
    uint8_t *anc;                /* [sum nA] 1 = allele not covered by any haplotype candidate */
    uint16_t *hc;                /* [n_variants] num_candidates */
} btg_genotype_result;

/* InferenceEngine::estimateGenotypes (InferenceEngine.cpp:278-382): default mode, fixed noise rates */
int btg_estimate_genotypes(btg_unit *u, const btg_count_dist *cd, const btg_gibbs_opts *opts, btg_genotype_result *out);
/* the same, split for callers that keep results on the device: launch on `stream` (NULL = library
 * stream) without synchronising, then copy the result arrays out when needed */
int btg_estimate_genotypes_async(btg_unit *u, const btg_count_dist *cd, const btg_gibbs_opts *opts, void *stream);
int btg_unit_download_result(btg_unit *u, btg_genotype_result *out, void *stream);
/* InferenceEngine::estimateNoise (InferenceEngine.cpp:135-276): updates cd's noise rates; trace_out (optional)
 * receives the <prefix>_noise_parameters.txt rows: [n_chains*(iters+1)+1][2+S] doubles (chain, iteration, rates..) */
int btg_estimate_noise(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, double *trace_out);
/* InferenceEngine::estimateNoiseAndGenotypes (InferenceEngine.cpp:384-472, --noise-genotyping): all groups in lock-step, noise
 * rates redrawn after every iteration, genotypers persisting across chains; trace rows: [n_chains*(iters+1)][2+S] */
int btg_estimate_noise_and_genotypes(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, btg_genotype_result *out,
                                     double *trace_out);
/* ---- one inference unit sharded over several GPUs (one rank per process and GPU) ----------
 * Groups are independent, so a rank uploads the groups group_index_base + j * group_index_stride (j < n_groups) of the unit as
 * its btg_unit and keeps the reference's per-group seeds through btg_gibbs_opts.group_index_base / group_index_stride.  btg_estimate_genotypes
 * needs nothing else.  The lock-step modes (estimateNoise, estimateNoiseAndGenotypes) merge every thread's
 * CountAllocation once per iteration (InferenceEngine.cpp:226-229,445-448); across ranks that merge is the per-sample
 * (n_obs, sum) pair written into every peer's mailbox over NVLink from inside the chain kernel, after which all ranks
 * draw the same noise rates from the shared CountDistribution stream.  Results equal the single-GPU run bit for bit.
 * A communicator owns this rank's mailbox; the 64-byte handles are exchanged by the host (any transport).          */
#define BTG_COMM_HANDLE_BYTES 64
typedef struct btg_comm btg_comm;
btg_comm *btg_comm_create(uint32_t world, uint32_t rank, uint8_t *handle_out /* [BTG_COMM_HANDLE_BYTES] */);
int btg_comm_connect(btg_comm *c, const uint8_t *all_handles /* [world][BTG_COMM_HANDLE_BYTES], rank order */);
void btg_comm_free(btg_comm *c);
typedef struct btg_shard_desc {
    btg_comm *comm;                    /* NULL = this rank holds the whole unit */
    uint64_t n_groups_total;           /* groups of the whole unit */
    const uint32_t *group_n_clusters;  /* [n_groups_total] clusters per group (estimateNoise uses single-cluster groups, InferenceEngine.cpp:144-151) */
    const uint32_t *group_n_variants;  /* [n_groups_total] variants per group (the 100,000-variant batch, InferenceEngine.cpp:50,183-189) */
} btg_shard_desc;
int btg_estimate_noise_sharded(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *shard, double *trace_out);
int btg_estimate_noise_and_genotypes_sharded(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *shard,
                                             btg_genotype_result *out, double *trace_out);
/* estimateNoise of ONE unit spread over several GPUs by CHAINS.  Under this library's stream contract the chains of estimateNoise
 * are independent (own noise stream per chain, reset frequencies, fresh k-mer subsample) and the result is the
 * mean of their post-burn-in rates (InferenceEngine.cpp:259-264), so rank r of N can run the chains c = r (mod N) of the WHOLE unit
 * with no exchange while they run — the per-iteration merge of the reference (InferenceEngine.cpp:226-229) stays inside one GPU.
 * chain_sums_out [n_chains][S] receives the post-burn-in rate sums of the chains this call ran (other rows 0); cd is not
 * updated.  The host adds the ranks' arrays up (every row is non-zero on exactly one rank) and calls
 * btg_count_dist_finish_noise on every rank: same summation order as btg_estimate_noise, same rates bit for bit.
 * trace_out (optional): as btg_estimate_noise, rows of the other ranks' chains and the final row 0.                              */
int btg_estimate_noise_chains(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, uint32_t chain_first, uint32_t chain_stride,
                              double *chain_sums_out, double *trace_out);
int btg_count_dist_finish_noise(btg_count_dist *cd, const double *chain_sums /* [n_chains][S] */, uint32_t n_chains, uint32_t gibbs_samples);

/* raw diplotype tallies of one cluster (tests): [(H+1)(H+2)/2][S] uint32, pair (h1<=h2), index h2*(h2+1)/2+h1, H = "missing" */
int btg_unit_cluster_tally(const btg_unit *u, uint32_t cluster, uint32_t *tally_out, uint64_t n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BTGPU_H */
