// counter.cu — the genotype-side k-mer stages of KmerCounter behind ONE handle of the C ABI (sm_100a).
//
// Replaces, for one inference unit:
//   KmerCounter::countPathKmers / countInterclusterKmers / parseSampleKmers   include/bayesTyper/KmerCounter.hpp:61-67,
//                                                                             src/bayesTyper/KmerCounter.cpp:252-524
//   VariantClusterGraph::countPathKmers / classifyPathKmers / getHaplotypeCandidates (+ updateVariantPathIndices)
//                                                                             src/bayesTyper/VariantClusterGraph.cpp:800-1184
//   KmerCountsHash / KmerCounts (the count table and its flags)               include/bayesTyper/KmerHash.hpp:73-108,
//                                                                             src/bayesTyper/KmerCounts.cpp:93-189
//
// The reference keeps an unordered hash of heap records and walks every cluster's paths three times through it.  Here the table is a
// sorted array of 128-bit keys with flat columns, and the relational steps between the per-nucleotide / per-record kernels of
// table.cu (distinct keys, distinct (cluster, key) rows, multiplicities per (row, path), coverage bitmaps per (row, variant)) are
// device radix sorts, run-length encodings and scans (CUB) plus scatter kernels — no hash, no host pass over k-mers.  Round 1 did
// this glue with torch tensor ops in the Python mirror (bayestyper_b200/kmer_pipeline.py), which a C++ host could not call; that
// mirror stays as the second implementation the tests compare with (both are pinned to the reference's dumps).
#include <algorithm>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"
#include "kmer.cuh"

using namespace btg;

namespace {

// ---- device buffers ------------------------------------------------------------------------------------------------------
// Stream-ordered allocations from the device's default memory pool (cudaMallocAsync on the library stream): the stages allocate ~100
// intermediates of up to a few hundred MB per unit; with cudaMalloc / cudaFree each of them is a driver call that synchronises the device
// (0.4 s per stage on configs[1]); the pool keeps the memory between stages and units (release threshold = never).
inline void pool_setup() {
    static bool done = false;
    if (done) return;
    cudaMemPool_t pool;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done = true;
}
template <class T> struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DBuf() { release(); }
    void release() { if (p) cudaFreeAsync(p, ctx().stream); p = nullptr; n = 0; }
    bool alloc(size_t count, bool zero = false, cudaStream_t s = nullptr) {
        release();
        pool_setup();
        n = count;
        if (cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), ctx().stream) != cudaSuccess) { p = nullptr; n = 0; cudaGetLastError(); return false; }
        if (zero) cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T), ctx().stream);
        (void)s;
        return true;
    }
    bool upload(const T *h, size_t count, cudaStream_t s) {
        if (!alloc(count)) return false;
        return count == 0 || cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s) == cudaSuccess;
    }
    std::vector<T> download(cudaStream_t s) const {
        std::vector<T> h(n);
        if (n) { cudaMemcpyAsync(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
        return h;
    }
};

struct Temp {   // CUB temporary storage, grown on demand
    void *p = nullptr;
    size_t cap = 0;
    ~Temp() { if (p) cudaFreeAsync(p, ctx().stream); }
    bool need(size_t bytes) {
        if (bytes <= cap) return true;
        if (p) cudaFreeAsync(p, ctx().stream);
        cap = bytes + bytes / 4 + 256;
        pool_setup();
        return cudaMallocAsync(&p, cap, ctx().stream) == cudaSuccess;
    }
};

#define CK(call) do { if ((call) != cudaSuccess) return false; } while (0)

// stable sort of (key, value) pairs by key; in/out in place (ping-pong buffers handled here)
template <class K, class V> bool sort_pairs(Temp &tmp, DBuf<K> &keys, DBuf<V> &vals, size_t n, cudaStream_t s, int end_bit = sizeof(K) * 8) {
    if (n == 0) return true;
    DBuf<K> k2; DBuf<V> v2;
    if (!k2.alloc(n) || !v2.alloc(n)) return false;
    size_t bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, k2.p, vals.p, v2.p, n, 0, end_bit, s));
    if (!tmp.need(bytes)) return false;
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, k2.p, vals.p, v2.p, n, 0, end_bit, s));
    std::swap(keys.p, k2.p); std::swap(vals.p, v2.p);
    return true;
}
template <class T> bool exclusive_sum(Temp &tmp, const T *in, T *out, size_t n, cudaStream_t s) {
    if (n == 0) return true;
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
    if (!tmp.need(bytes)) return false;
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, s));
    return true;
}
template <class T> bool inclusive_sum(Temp &tmp, const T *in, T *out, size_t n, cudaStream_t s) {
    if (n == 0) return true;
    size_t bytes = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, s));
    if (!tmp.need(bytes)) return false;
    CK(cub::DeviceScan::InclusiveSum(tmp.p, bytes, in, out, n, s));
    return true;
}
// distinct values of a SORTED array with their run lengths; returns the number of runs through n_runs_host
template <class K> bool run_length(Temp &tmp, const K *sorted, size_t n, DBuf<K> &uniq, DBuf<uint32_t> &counts, size_t &n_runs_host, cudaStream_t s) {
    n_runs_host = 0;
    if (!uniq.alloc(n) || !counts.alloc(n)) return false;
    if (n == 0) return true;
    DBuf<uint64_t> d_n;
    if (!d_n.alloc(1)) return false;
    size_t bytes = 0;
    CK(cub::DeviceRunLengthEncode::Encode(nullptr, bytes, sorted, uniq.p, counts.p, d_n.p, (int64_t)n, s));
    if (!tmp.need(bytes)) return false;
    CK(cub::DeviceRunLengthEncode::Encode(tmp.p, bytes, sorted, uniq.p, counts.p, d_n.p, (int64_t)n, s));
    uint64_t h = 0;
    CK(cudaMemcpyAsync(&h, d_n.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    n_runs_host = (size_t)h;
    uniq.n = counts.n = n_runs_host;
    return true;
}

constexpr unsigned kB = 256;
inline unsigned grid_of(size_t n) { return (unsigned)std::min<size_t>((n + kB - 1) / kB, 1u << 30); }

// ---- small kernels ---------------------------------------------------------------------------------------------------------
__global__ void k_iota(uint32_t *a, size_t n) { const size_t i = blockIdx.x * (size_t)kB + threadIdx.x; if (i < n) a[i] = (uint32_t)i; }
__global__ void k_gather_i64(const int64_t *src, const uint32_t *idx, int64_t *dst, size_t n) { const size_t i = blockIdx.x * (size_t)kB + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }
// flags[i] = 1 when sorted key i starts a new run of (hi, lo)
__global__ void k_key_flags(const int64_t *lo, const int64_t *hi, uint32_t *flag, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n) flag[i] = i == 0 || lo[i] != lo[i - 1] || hi[i] != hi[i - 1];
}
// distinct keys and, per occurrence (original order), the index of its key
__global__ void k_emit_keys(const int64_t *lo, const int64_t *hi, const uint32_t *flag, const uint32_t *rank /* inclusive sum of flag */, const uint32_t *order,
                            int64_t *kw0, int64_t *kw1, uint32_t *occ_key, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = rank[i] - 1;
    if (flag[i]) { kw0[k] = lo[i]; kw1[k] = hi[i]; }
    occ_key[order[i]] = k;
}
// prefix index over key_hi: lut[b] = first key whose top lut_bits of the 46 are >= b
__global__ void k_build_lut(const int64_t *kw1, size_t n_keys, int shift, int64_t *lut, uint32_t n_buckets) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i > n_keys) return;
    const int64_t b_prev = i == 0 ? -1 : (kw1[i - 1] >> shift);
    const int64_t b_cur = i == n_keys ? (int64_t)n_buckets : (kw1[i] >> shift);
    for (int64_t b = b_prev + 1; b <= b_cur; b++) lut[b] = (int64_t)i;
}
// occurrence -> (cluster * n_keys + key), local path
__global__ void k_occ_pairs(const uint32_t *occ_path, const uint32_t *occ_key, const uint32_t *path_cluster, const uint64_t *cl_path_off, uint64_t n_keys,
                            uint64_t *pair, uint32_t *local_path, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = occ_path[i], c = path_cluster[p];
    pair[i] = (uint64_t)c * n_keys + occ_key[i];
    local_path[i] = p - (uint32_t)cl_path_off[c];
}
__global__ void k_pair_flags(const uint64_t *sorted, uint32_t *flag, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n) flag[i] = i == 0 || sorted[i] != sorted[i - 1];
}
// rows: per sorted occurrence its row = rank - 1; scatter back to the original occurrence order, first occurrence per row
__global__ void k_rows_of_occ(const uint64_t *sorted_pair, const uint32_t *flag, const uint32_t *rank, const uint32_t *order, uint64_t n_keys,
                              uint32_t *occ_row, uint32_t *row_cluster, uint32_t *row_key, uint32_t *first_occ, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = rank[i] - 1, o = order[i];
    occ_row[o] = r;
    if (flag[i]) { row_cluster[r] = (uint32_t)(sorted_pair[i] / n_keys); row_key[r] = (uint32_t)(sorted_pair[i] % n_keys); }
    atomicMin(first_occ + r, o);
}
__global__ void k_trips(const uint32_t *occ_row, const uint32_t *local_path, uint64_t max_h, uint64_t *trip, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n) trip[i] = (uint64_t)occ_row[i] * max_h + local_path[i];
}
// per distinct (row, path): saturated count; per row: max count
__global__ void k_trip_stats(const uint64_t *trip_u, const uint32_t *trip_cnt, uint64_t max_h, uint32_t *t_row, uint32_t *t_path, uint8_t *t_cnt, uint32_t *row_max, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = (uint32_t)(trip_u[i] / max_h), c = trip_cnt[i] > 255u ? 255u : trip_cnt[i];   // uchar saturation (VariantClusterGraph.cpp:889-893)
    t_row[i] = r; t_path[i] = (uint32_t)(trip_u[i] % max_h); t_cnt[i] = (uint8_t)c;
    atomicMax(row_max + r, c);
}
// per key over its rows: number of clusters, sum and max of the row maxima (KmerCounts::addClusterMultiplicity, KmerCounts.cpp:137-159)
__global__ void k_key_stats(const uint32_t *row_key, const uint32_t *row_max, uint32_t *n_cl, uint32_t *sum_max, uint32_t *max_max, size_t n_rows) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n_rows) return;
    const uint32_t k = row_key[i];
    atomicAdd(n_cl + k, 1u); atomicAdd(sum_max + k, row_max[i]); atomicMax(max_max + k, row_max[i]);
}
__global__ void k_group_key(const uint32_t *row_cluster, const uint32_t *row_key, const uint32_t *cl_group, uint64_t n_keys, uint64_t *gk, size_t n_rows) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n_rows) gk[i] = (uint64_t)cl_group[row_cluster[i]] * n_keys + row_key[i];
}
__global__ void k_count_groups(const uint64_t *gk_u, uint64_t n_keys, uint32_t *n_gr, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n) atomicAdd(n_gr + (uint32_t)(gk_u[i] % n_keys), 1u);
}
// flags of a key (KmerCounts.cpp:93-159): bit0 record, bit1 multicluster, bit2 multigroup, bit3 excluded
__global__ void k_key_flags_final(const uint8_t *has_record, const uint8_t *decoy, const uint8_t *max_mult, const uint32_t *n_cl, const uint32_t *sum_max,
                                  const uint32_t *max_max, const uint32_t *n_gr, const uint8_t *mg_hit, uint8_t *flags, size_t n_keys) {
    const size_t k = blockIdx.x * (size_t)kB + threadIdx.x;
    if (k >= n_keys) return;
    const bool rec = has_record[k] || max_max[k] > 127u;
    uint32_t mh = (uint32_t)max_mult[k] + sum_max[k];
    if (mh > 255u) mh = 255u;
    const bool multigroup = mg_hit ? mg_hit[k] != 0 : n_gr[k] >= 2u;
    const bool excluded = rec && (decoy[k] || mh > 127u || multigroup);
    flags[k] = (uint8_t)((rec ? 1 : 0) | (rec && n_cl[k] >= 2u ? 2 : 0) | (rec && multigroup ? 4 : 0) | (excluded ? 8 : 0));
}
// kept rows: sort key = cluster * (N + 1) + first occurrence (kmer_row_indices order, VariantClusterGraph.cpp:1056)
__global__ void k_keep_rows(const uint32_t *row_cluster, const uint32_t *row_key, const uint32_t *first_occ, const uint8_t *flags, uint64_t n_occ, uint64_t *sort_key, uint32_t *keep,
                            size_t n_rows) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n_rows) return;
    const bool k = !(flags[row_key[i]] & 8);
    keep[i] = k;
    sort_key[i] = k ? (uint64_t)row_cluster[i] * (n_occ + 1) + first_occ[i] : ~0ull;
}
__global__ void k_new_rows(const uint32_t *kept /* old row ids in new order */, const uint32_t *row_cluster, const uint32_t *row_key, uint32_t *new_row, uint32_t *k_cluster, uint32_t *k_key,
                           uint64_t *cl_rows, size_t n_kept) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n_kept) return;
    const uint32_t r = kept[i];
    new_row[r] = (uint32_t)i; k_cluster[i] = row_cluster[r]; k_key[i] = row_key[r];
    atomicAdd((unsigned long long *)cl_rows + row_cluster[r], 1ull);
}
__global__ void k_mul_u64(const uint64_t *a, const uint32_t *b, uint64_t *out, size_t n) { const size_t i = blockIdx.x * (size_t)kB + threadIdx.x; if (i < n) out[i] = a[i] * b[i]; }
__global__ void k_fill_mult(const uint32_t *t_row, const uint32_t *t_path, const uint8_t *t_cnt, const uint32_t *new_row, const uint32_t *k_cluster, const uint64_t *cl_kmer_off,
                            const uint64_t *cl_mult_off, const uint32_t *n_paths, uint8_t *mult, size_t n_trips) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n_trips) return;
    const uint32_t nr = new_row[t_row[i]];
    if (nr == 0xFFFFFFFFu) return;
    const uint32_t c = k_cluster[nr];
    mult[cl_mult_off[c] + (uint64_t)(nr - cl_kmer_off[c]) * n_paths[c] + t_path[i]] = t_cnt[i];
}
// per kept row: the columns the sampler reads + unique / multicluster membership
__global__ void k_row_columns(const uint32_t *k_cluster, const uint32_t *k_key, const uint8_t *flags, const uint8_t *counts, const uint8_t *ic, const uint32_t *shared_id, uint32_t S,
                              uint8_t *k_has_counts, uint8_t *k_counts, uint8_t *k_ic, uint32_t *k_shared, uint32_t *is_uniq, uint32_t *is_multi, uint64_t *cl_uniq, uint64_t *cl_multi,
                              size_t n_kept) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n_kept) return;
    const uint32_t k = k_key[i], c = k_cluster[i];
    const uint8_t f = flags[k];
    k_has_counts[i] = f & 1;
    for (uint32_t s = 0; s < S; s++) k_counts[i * S + s] = counts[(size_t)k * S + s];
    k_ic[i * 2] = ic[(size_t)k * 2]; k_ic[i * 2 + 1] = ic[(size_t)k * 2 + 1];
    const bool multi = f & 2;
    k_shared[i] = multi ? shared_id[k] : 0xFFFFFFFFu;
    is_uniq[i] = !multi; is_multi[i] = multi;
    atomicAdd((unsigned long long *)(multi ? cl_multi : cl_uniq) + c, 1ull);
}
__global__ void k_list_rows(const uint32_t *is_sel, const uint32_t *rank /* exclusive sum of is_sel */, const uint32_t *k_cluster, const uint64_t *cl_kmer_off, uint32_t *out, size_t n_kept) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n_kept && is_sel[i]) out[rank[i]] = (uint32_t)(i - cl_kmer_off[k_cluster[i]]);
}
__global__ void k_flag_bit(const uint8_t *flags, uint8_t bit, uint32_t *out, size_t n) { const size_t i = blockIdx.x * (size_t)kB + threadIdx.x; if (i < n) out[i] = (flags[i] & bit) ? 1u : 0u; }
// coverage events (occurrence, variant) -> (new row << 16 | variant), ~0 when the row was dropped
__global__ void k_cov_events(const int64_t *cov_occ, const uint16_t *cov_var, const uint32_t *occ_row, const uint32_t *new_row, uint64_t *ev, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n) return;
    const uint32_t nr = new_row[occ_row[cov_occ[i]]];
    ev[i] = nr == 0xFFFFFFFFu ? ~0ull : ((uint64_t)nr << 16) | cov_var[i];
}
__global__ void k_ev_heads(const uint64_t *ev_u, const uint32_t *k_cluster, const uint32_t *n_paths, uint16_t *vh_var, uint64_t *e_h, uint64_t *kmer_vh, size_t n_ev) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n_ev) return;
    const uint32_t row = (uint32_t)(ev_u[i] >> 16);
    vh_var[i] = (uint16_t)(ev_u[i] & 0xFFFFu);
    e_h[i] = n_paths[k_cluster[row]];
    atomicAdd((unsigned long long *)kmer_vh + row, 1ull);
}
// the sorted coverage events: entry j belongs to distinct event rank[j] - 1; set the bit of its path
__global__ void k_vh_bits(const uint32_t *rank, const uint32_t *ev_order, const int64_t *cov_occ, const uint32_t *local_path, const uint64_t *vh_bits_off, uint8_t *vh_bits, size_t n_valid) {
    const size_t j = blockIdx.x * (size_t)kB + threadIdx.x;
    if (j >= n_valid) return;
    vh_bits[vh_bits_off[rank[j] - 1] + local_path[cov_occ[ev_order[j]]]] = 1;
}

// ---- NB fit (parameter k-mers) ------------------------------------------------------------------------------------------------
__global__ void k_valid_flags(const uint8_t *valid, uint32_t *flag, size_t n) { const size_t i = blockIdx.x * (size_t)kB + threadIdx.x; if (i < n) flag[i] = valid[i] != 0; }
__global__ void k_compact_kmers(const ulonglong2 *km, const uint32_t *flag, const uint32_t *rank /* exclusive */, ulonglong2 *out, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n && flag[i]) out[rank[i]] = km[i];
}
// distinct sorted keys: start index of every run
__global__ void k_run_starts(const uint32_t *flag, const uint32_t *rank /* inclusive */, const int64_t *lo, const int64_t *hi, int64_t *u_lo, int64_t *u_hi, uint32_t *start, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n && flag[i]) { const uint32_t r = rank[i] - 1; u_lo[r] = lo[i]; u_hi[r] = hi[i]; start[r] = (uint32_t)i; }
}
// splitmix64 of (seed, key): the Bernoulli of the parameter k-mer subsample (a property of the k-mer, so the choice does not depend on the scan order)
__device__ __forceinline__ uint64_t mix64(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
// select[i] = the inter-cluster k-mer is not a path k-mer and is a parameter k-mer (on the list when one is given, else a Bernoulli draw); occ = run length
__global__ void k_param_select(const uint32_t *start, size_t n_distinct, size_t n_total, const int64_t *path_idx, const int64_t *list_idx, const int64_t *u_lo, const int64_t *u_hi,
                               uint64_t seed, double frac, uint32_t *sel, uint32_t *occ) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i >= n_distinct) return;
    occ[i] = (uint32_t)((i + 1 < n_distinct ? start[i + 1] : n_total) - start[i]);
    bool ok = path_idx[i] < 0;
    if (ok) {
        if (list_idx) ok = list_idx[i] >= 0;
        else ok = (double)(mix64(mix64(seed ^ (uint64_t)u_lo[i]) ^ (uint64_t)u_hi[i]) >> 11) * (1.0 / 9007199254740992.0) < frac;
    }
    sel[i] = ok;
}
__global__ void k_compact_params(const uint32_t *sel, const uint32_t *rank /* exclusive */, const int64_t *u_lo, const int64_t *u_hi, const uint32_t *occ, int64_t *p_lo, int64_t *p_hi,
                                 uint32_t *p_occ, size_t n) {
    const size_t i = blockIdx.x * (size_t)kB + threadIdx.x;
    if (i < n && sel[i]) { const uint32_t r = rank[i]; p_lo[r] = u_lo[i]; p_hi[r] = u_hi[i]; p_occ[r] = occ[i]; }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------
struct btg_counter {
    uint32_t S = 0, G = 0, C = 0;
    uint64_t P = 0, V = 0, n_var = 0;
    cudaStream_t s = nullptr;
    Temp tmp;
    // host copies of the small descriptors (they end up in the unit descriptor)
    std::vector<uint8_t> h_gender, h_var_dep;
    std::vector<uint64_t> h_group_cluster_off, h_group_src_off, h_group_edge_off, h_cl_vertex_off, h_cl_var_off, h_v_refvar_off, h_cl_path_off, h_path_mem_off;
    std::vector<uint32_t> h_group_src, h_group_edge_src, h_group_edge_dst, h_cluster_idx, h_n_paths, h_v_nested, h_cl_group;
    std::vector<uint16_t> h_var_nalleles, h_v_var, h_v_refvar;
    std::vector<uint8_t> h_path_mem;
    // graphs on the device (btg_pathwalk_desc)
    DBuf<uint64_t> cl_vertex_off, v_seq_off, v_refvar_off, cl_path_off, path_mem_off, cl_var_off;
    DBuf<uint8_t> seq, v_flags, path_mem;
    DBuf<uint16_t> v_var, v_allele, v_refvar, var_nalleles;
    DBuf<uint32_t> path_cluster, n_paths, cl_group;
    btg_pathwalk_desc walk{};
    // occurrences
    uint64_t N = 0, NC = 0;
    DBuf<uint32_t> occ_path, occ_nt, occ_key;
    DBuf<int64_t> cov_occ;
    DBuf<uint16_t> cov_var;
    // table
    uint64_t n_keys = 0;
    int lut_bits = 0;
    DBuf<int64_t> kw0, kw1, lut;
    DBuf<uint8_t> counts, ic, max_mult, decoy, has_record, key_flags;
    bool counted = false;
    // unit arrays (device, after build_unit) + host offsets
    std::vector<uint64_t> h_cl_kmer_off, h_cl_mult_off, h_cl_uniq_off, h_cl_multi_off, h_cl_hapvar_off;
    std::vector<uint32_t> h_multi_idx;
    DBuf<uint8_t> u_mult, u_has_counts, u_counts, u_ic, u_vh_bits;
    DBuf<uint32_t> u_shared, u_uniq_idx, u_multi_idx;
    DBuf<uint64_t> u_kmer_vh_off, u_vh_bits_off;
    DBuf<uint16_t> u_vh_var, u_hap_alleles;
    uint64_t n_vh = 0, n_vh_bits = 0, n_rows = 0;
    std::vector<uint64_t> h_hap_nested_off, h_dep_var_off, h_cl_dep_off;
    std::vector<uint32_t> h_hap_nested, h_dep_cluster;
    std::vector<uint16_t> h_dep_var;

    void use_index() { btg_table_set_index_dev(lut.p, lut_bits); }
};

namespace {

template <class T> void copy_vec(std::vector<T> &dst, const T *src, size_t n) { dst.assign(src, src + n); }

// sorted distinct table keys of a set of packed k-mers on the device: (lo, hi) ascending, optionally the start of every run
bool distinct_keys(btg_counter &k, const uint64_t *kmers_dev, size_t n, DBuf<int64_t> &u_lo, DBuf<int64_t> &u_hi, DBuf<uint32_t> *starts, size_t &n_distinct) {
    cudaStream_t s = k.s;
    n_distinct = 0;
    DBuf<int64_t> lo, hi, key;
    DBuf<uint32_t> order, flag, rank;
    if (!lo.alloc(n) || !hi.alloc(n) || !key.alloc(n) || !order.alloc(n) || !flag.alloc(n) || !rank.alloc(n)) return false;
    if (n == 0) { u_lo.alloc(0); u_hi.alloc(0); if (starts) starts->alloc(0); return true; }
    if (btg_table_keys_from_kmers_dev(kmers_dev, n, lo.p, hi.p, s) != BTG_OK) return false;
    k_iota<<<grid_of(n), kB, 0, s>>>(order.p, n);
    CK(cudaMemcpyAsync(key.p, lo.p, n * 8, cudaMemcpyDeviceToDevice, s));
    if (!sort_pairs(k.tmp, key, order, n, s)) return false;
    k_gather_i64<<<grid_of(n), kB, 0, s>>>(hi.p, order.p, key.p, n);
    if (!sort_pairs(k.tmp, key, order, n, s, 47)) return false;
    DBuf<int64_t> s_lo;
    if (!s_lo.alloc(n)) return false;
    k_gather_i64<<<grid_of(n), kB, 0, s>>>(lo.p, order.p, s_lo.p, n);     // key.p holds the sorted hi
    k_key_flags<<<grid_of(n), kB, 0, s>>>(s_lo.p, key.p, flag.p, n);
    if (!inclusive_sum(k.tmp, flag.p, rank.p, n, s)) return false;
    uint32_t nd = 0;
    CK(cudaMemcpyAsync(&nd, rank.p + n - 1, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    n_distinct = nd;
    DBuf<uint32_t> st;
    if (!u_lo.alloc(nd) || !u_hi.alloc(nd) || !st.alloc(nd)) return false;
    k_run_starts<<<grid_of(n), kB, 0, s>>>(flag.p, rank.p, s_lo.p, key.p, u_lo.p, u_hi.p, st.p, n);
    if (starts) *starts = std::move(st);
    return true;
}

bool count_path_kmers(btg_counter &k) {
    cudaStream_t s = k.s;
    const uint64_t P = k.P;
    DBuf<uint32_t> n_occ, n_cov, status;
    if (!n_occ.alloc(P, true, s) || !n_cov.alloc(P, true, s) || !status.alloc(k.C, true, s)) return false;
    if (btg_walk_paths_dev(&k.walk, 0, n_occ.p, n_cov.p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, status.p, s) != BTG_OK) return false;
    // exclusive prefix sums as uint64
    DBuf<uint64_t> occ_off, cov_off, tmp64;
    if (!occ_off.alloc(P + 1, true, s) || !cov_off.alloc(P + 1, true, s)) return false;
    {
        std::vector<uint32_t> h_occ = n_occ.download(s), h_cov = n_cov.download(s);
        std::vector<uint64_t> a(P + 1, 0), b(P + 1, 0);
        for (uint64_t i = 0; i < P; i++) { a[i + 1] = a[i] + h_occ[i]; b[i + 1] = b[i] + h_cov[i]; }
        k.N = a[P]; k.NC = b[P];
        CK(cudaMemcpyAsync(occ_off.p, a.data(), (P + 1) * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(cov_off.p, b.data(), (P + 1) * 8, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    const uint64_t N = k.N, NC = k.NC;
    DBuf<int64_t> w0, w1;
    if (!w0.alloc(N) || !w1.alloc(N) || !k.occ_path.alloc(N) || !k.occ_nt.alloc(N) || !k.cov_occ.alloc(NC) || !k.cov_var.alloc(NC)) return false;
    if (btg_walk_paths_dev(&k.walk, 1, nullptr, nullptr, occ_off.p, cov_off.p, w0.p, w1.p, k.occ_path.p, k.occ_nt.p, k.cov_occ.p, k.cov_var.p, status.p, s) != BTG_OK) return false;
    {
        std::vector<uint32_t> st = status.download(s);
        for (uint32_t c = 0; c < k.C; c++) if (st[c]) { set_error("path walk: running-variant capacity exceeded in cluster %u", c); return false; }
    }
    // distinct keys, ascending (key_hi, key_lo): two stable passes
    DBuf<uint32_t> order;
    if (!order.alloc(N)) return false;
    if (N) k_iota<<<grid_of(N), kB, 0, s>>>(order.p, N);
    {
        DBuf<int64_t> key;
        if (!key.alloc(N)) return false;
        CK(cudaMemcpyAsync(key.p, w0.p, N * 8, cudaMemcpyDeviceToDevice, s));
        if (!sort_pairs(k.tmp, key, order, N, s)) return false;                     // by key_lo
        if (N) k_gather_i64<<<grid_of(N), kB, 0, s>>>(w1.p, order.p, key.p, N);
        if (!sort_pairs(k.tmp, key, order, N, s, 47)) return false;                 // by key_hi (46 bits, non-negative)
    }
    DBuf<int64_t> s0, s1;
    DBuf<uint32_t> flag, rank;
    if (!s0.alloc(N) || !s1.alloc(N) || !flag.alloc(N) || !rank.alloc(N)) return false;
    if (N) {
        k_gather_i64<<<grid_of(N), kB, 0, s>>>(w0.p, order.p, s0.p, N);
        k_gather_i64<<<grid_of(N), kB, 0, s>>>(w1.p, order.p, s1.p, N);
        k_key_flags<<<grid_of(N), kB, 0, s>>>(s0.p, s1.p, flag.p, N);
    }
    if (!inclusive_sum(k.tmp, flag.p, rank.p, N, s)) return false;
    uint32_t nk = 0;
    if (N) { CK(cudaMemcpyAsync(&nk, rank.p + N - 1, 4, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
    k.n_keys = nk;
    if (!k.kw0.alloc(nk) || !k.kw1.alloc(nk) || !k.occ_key.alloc(N)) return false;
    if (N) k_emit_keys<<<grid_of(N), kB, 0, s>>>(s0.p, s1.p, flag.p, rank.p, order.p, k.kw0.p, k.kw1.p, k.occ_key.p, N);
    // prefix index over key_hi (46 bits = the first 23 nucleotides): ~1 key per bucket
    int bits = 8;
    while (bits < 24 && (1ull << bits) < std::max<uint64_t>(nk, 2)) bits++;
    k.lut_bits = bits;
    const uint32_t n_buckets = 1u << bits;
    if (!k.lut.alloc((size_t)n_buckets + 1, true, s)) return false;
    k_build_lut<<<grid_of((size_t)nk + 1), kB, 0, s>>>(k.kw1.p, nk, 2 * K - 64 - bits, k.lut.p, n_buckets);
    const uint32_t S = k.S;
    if (!k.counts.alloc((size_t)nk * S + 4, true, s) || !k.ic.alloc((size_t)nk * 2, true, s) || !k.max_mult.alloc((size_t)nk + 4, true, s) || !k.decoy.alloc(nk, true, s) ||
        !k.has_record.alloc(nk, true, s))
        return false;
    BTG_LAUNCHED();
    CK(cudaStreamSynchronize(s));
    k.counted = true;
    return true;
}

// HaplotypeInfo::nested_variant_cluster_indices and nested_variant_cluster_dependency (VariantClusterGraph.cpp:1006-1010,1112-1133): only clusters
// holding a vertex that stands for a nested cluster contribute — a short host pass over those clusters
void nested_tables(btg_counter &k) {
    const uint64_t P = k.P;
    k.h_hap_nested_off.assign(P + 1, 0);
    k.h_cl_dep_off.assign((size_t)k.C + 1, 0);
    k.h_hap_nested.clear(); k.h_dep_cluster.clear(); k.h_dep_var.clear();
    k.h_dep_var_off.assign(1, 0);
    if (k.h_v_nested.empty()) return;
    std::vector<std::vector<uint32_t>> per_hap(P);
    for (uint32_t c = 0; c < k.C; c++) {
        const uint64_t v0 = k.h_cl_vertex_off[c], v1 = k.h_cl_vertex_off[c + 1], nv = v1 - v0;
        bool any = false;
        for (uint64_t v = v0; v < v1; v++) any = any || k.h_v_nested[v] != 0xFFFFFFFFu;
        if (!any) continue;
        const uint8_t *bits = k.h_path_mem.data() + k.h_path_mem_off[c];
        for (uint32_t p = 0; p < k.h_n_paths[c]; p++) {
            std::vector<uint32_t> &lst = per_hap[k.h_cl_path_off[c] + p];
            for (uint64_t v = 0; v < nv; v++) if (k.h_v_nested[v0 + v] != 0xFFFFFFFFu && bits[(size_t)p * nv + v]) lst.push_back(k.h_v_nested[v0 + v]);
            std::sort(lst.begin(), lst.end());
        }
        std::vector<std::pair<uint32_t, std::vector<uint16_t>>> deps;
        for (uint64_t v = 0; v < nv; v++) {
            if (k.h_v_nested[v0 + v] == 0xFFFFFFFFu) continue;
            std::vector<uint16_t> vars;
            if (k.h_v_var[v0 + v] != 0xFFFF) vars.push_back(k.h_v_var[v0 + v]);
            for (uint64_t e = k.h_v_refvar_off[v0 + v]; e < k.h_v_refvar_off[v0 + v + 1]; e++) vars.push_back(k.h_v_refvar[e]);
            std::sort(vars.begin(), vars.end(), std::greater<uint16_t>());
            bool found = false;
            for (auto &d : deps) if (d.first == k.h_v_nested[v0 + v]) { d.second = vars; found = true; }   // a later vertex of the same nested cluster replaces (dict semantics)
            if (!found) deps.emplace_back(k.h_v_nested[v0 + v], vars);
        }
        std::sort(deps.begin(), deps.end(), [](auto &a, auto &b) { return a.first < b.first; });
        for (auto &d : deps) { k.h_dep_cluster.push_back(d.first); k.h_dep_var.insert(k.h_dep_var.end(), d.second.begin(), d.second.end()); k.h_dep_var_off.push_back(k.h_dep_var.size()); }
        k.h_cl_dep_off[c + 1] = deps.size();
    }
    for (uint64_t p = 0; p < P; p++) { k.h_hap_nested_off[p + 1] = k.h_hap_nested_off[p] + per_hap[p].size(); k.h_hap_nested.insert(k.h_hap_nested.end(), per_hap[p].begin(), per_hap[p].end()); }
    for (uint32_t c = 0; c < k.C; c++) k.h_cl_dep_off[c + 1] += k.h_cl_dep_off[c];
}

// classifyPathKmers + getHaplotypeCandidates (VariantClusterGraph.cpp:848-1135) for every cluster of the unit
bool build_unit_arrays(btg_counter &k, const btg_bloom *multigroup) {
    cudaStream_t s = k.s;
    const uint64_t N = k.N, nk = k.n_keys;
    const uint32_t C = k.C, S = k.S;
    // rows = distinct (cluster, key)
    DBuf<uint64_t> pair;
    DBuf<uint32_t> local_path, order;
    if (!pair.alloc(N) || !local_path.alloc(N) || !order.alloc(N)) return false;
    if (N) {
        k_occ_pairs<<<grid_of(N), kB, 0, s>>>(k.occ_path.p, k.occ_key.p, k.path_cluster.p, k.cl_path_off.p, std::max<uint64_t>(nk, 1), pair.p, local_path.p, N);
        k_iota<<<grid_of(N), kB, 0, s>>>(order.p, N);
    }
    if (!sort_pairs(k.tmp, pair, order, N, s)) return false;
    DBuf<uint32_t> flag, rank, occ_row, row_cluster, row_key, first_occ;
    if (!flag.alloc(N) || !rank.alloc(N) || !occ_row.alloc(N)) return false;
    if (N) k_pair_flags<<<grid_of(N), kB, 0, s>>>(pair.p, flag.p, N);
    if (!inclusive_sum(k.tmp, flag.p, rank.p, N, s)) return false;
    uint32_t R = 0;
    if (N) { CK(cudaMemcpyAsync(&R, rank.p + N - 1, 4, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
    if (!row_cluster.alloc(R) || !row_key.alloc(R) || !first_occ.alloc(R)) return false;
    CK(cudaMemsetAsync(first_occ.p, 0xFF, std::max<size_t>(R, 1) * 4, s));
    if (N) k_rows_of_occ<<<grid_of(N), kB, 0, s>>>(pair.p, flag.p, rank.p, order.p, std::max<uint64_t>(nk, 1), occ_row.p, row_cluster.p, row_key.p, first_occ.p, N);
    pair.release(); flag.release(); rank.release(); order.release();
    // multiplicities = occurrences per (row, path)
    uint32_t max_h = 1;
    for (uint32_t c = 0; c < C; c++) max_h = std::max(max_h, k.h_n_paths[c]);
    DBuf<uint64_t> trip, trip_u;
    DBuf<uint32_t> trip_cnt, dummy;
    if (!trip.alloc(N) || !dummy.alloc(N)) return false;
    if (N) k_trips<<<grid_of(N), kB, 0, s>>>(occ_row.p, local_path.p, max_h, trip.p, N);
    if (!sort_pairs(k.tmp, trip, dummy, N, s)) return false;
    dummy.release();
    size_t n_trips = 0;
    if (!run_length(k.tmp, trip.p, N, trip_u, trip_cnt, n_trips, s)) return false;
    trip.release();
    DBuf<uint32_t> t_row, t_path, row_max;
    DBuf<uint8_t> t_cnt;
    if (!t_row.alloc(n_trips) || !t_path.alloc(n_trips) || !t_cnt.alloc(n_trips) || !row_max.alloc(R, true, s)) return false;
    if (n_trips) k_trip_stats<<<grid_of(n_trips), kB, 0, s>>>(trip_u.p, trip_cnt.p, max_h, t_row.p, t_path.p, t_cnt.p, row_max.p, n_trips);
    trip_u.release(); trip_cnt.release();
    // per key
    DBuf<uint32_t> n_cl, sum_max, max_max, n_gr;
    if (!n_cl.alloc(nk, true, s) || !sum_max.alloc(nk, true, s) || !max_max.alloc(nk, true, s) || !n_gr.alloc(nk, true, s)) return false;
    if (R) k_key_stats<<<grid_of(R), kB, 0, s>>>(row_key.p, row_max.p, n_cl.p, sum_max.p, max_max.p, R);
    DBuf<uint8_t> mg_hit;
    if (multigroup) {   // the reference's own multigroup filter (KmerCounter.cpp:541; Bloom false positives included)
        DBuf<uint64_t> kmers;
        if (!kmers.alloc((size_t)nk * 2) || !mg_hit.alloc(nk, true, s)) return false;
        if (nk) {
            if (btg_table_keys_to_kmers_dev(k.kw0.p, k.kw1.p, nk, kmers.p, s) != BTG_OK) return false;
            if (btg_bloom_lookup_dev(multigroup, kmers.p, nk, mg_hit.p, s) != BTG_OK) return false;
        }
        CK(cudaStreamSynchronize(s));
    } else if (R) {     // exact: the k-mer occurs in more than one group
        DBuf<uint64_t> gk, gk_u;
        DBuf<uint32_t> d2, cnt;
        if (!gk.alloc(R) || !d2.alloc(R)) return false;
        k_group_key<<<grid_of(R), kB, 0, s>>>(row_cluster.p, row_key.p, k.cl_group.p, std::max<uint64_t>(nk, 1), gk.p, R);
        if (!sort_pairs(k.tmp, gk, d2, R, s)) return false;
        size_t n_gk = 0;
        if (!run_length(k.tmp, gk.p, R, gk_u, cnt, n_gk, s)) return false;
        if (n_gk) k_count_groups<<<grid_of(n_gk), kB, 0, s>>>(gk_u.p, std::max<uint64_t>(nk, 1), n_gr.p, n_gk);
    }
    if (!k.key_flags.alloc(nk, true, s)) return false;
    if (nk) k_key_flags_final<<<grid_of(nk), kB, 0, s>>>(k.has_record.p, k.decoy.p, k.max_mult.p, n_cl.p, sum_max.p, max_max.p, n_gr.p, multigroup ? mg_hit.p : nullptr, k.key_flags.p, nk);
    // kept rows in first-seen order within their cluster
    DBuf<uint64_t> sort_key;
    DBuf<uint32_t> keep, kept;
    if (!sort_key.alloc(R) || !keep.alloc(R) || !kept.alloc(R)) return false;
    if (R) {
        k_keep_rows<<<grid_of(R), kB, 0, s>>>(row_cluster.p, row_key.p, first_occ.p, k.key_flags.p, N, sort_key.p, keep.p, R);
        k_iota<<<grid_of(R), kB, 0, s>>>(kept.p, R);
    }
    if (!sort_pairs(k.tmp, sort_key, kept, R, s)) return false;
    uint64_t Rk = 0;
    {
        DBuf<uint32_t> ks;
        if (!ks.alloc(R)) return false;
        if (!inclusive_sum(k.tmp, keep.p, ks.p, R, s)) return false;
        uint32_t h = 0;
        if (R) { CK(cudaMemcpyAsync(&h, ks.p + R - 1, 4, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
        Rk = h;
    }
    k.n_rows = Rk;
    DBuf<uint32_t> new_row, k_cluster, k_key;
    DBuf<uint64_t> cl_rows, cl_kmer_off, cl_mult_rows, cl_mult_off;
    if (!new_row.alloc(R) || !k_cluster.alloc(Rk) || !k_key.alloc(Rk) || !cl_rows.alloc((size_t)C + 1, true, s) || !cl_kmer_off.alloc((size_t)C + 1) ||
        !cl_mult_rows.alloc((size_t)C + 1, true, s) || !cl_mult_off.alloc((size_t)C + 1))
        return false;
    CK(cudaMemsetAsync(new_row.p, 0xFF, std::max<size_t>(R, 1) * 4, s));
    if (Rk) k_new_rows<<<grid_of(Rk), kB, 0, s>>>(kept.p, row_cluster.p, row_key.p, new_row.p, k_cluster.p, k_key.p, cl_rows.p, Rk);
    if (!exclusive_sum(k.tmp, cl_rows.p, cl_kmer_off.p, (size_t)C + 1, s)) return false;
    if (C) k_mul_u64<<<grid_of(C), kB, 0, s>>>(cl_rows.p, k.n_paths.p, cl_mult_rows.p, C);
    if (!exclusive_sum(k.tmp, cl_mult_rows.p, cl_mult_off.p, (size_t)C + 1, s)) return false;
    k.h_cl_kmer_off = cl_kmer_off.download(s);
    k.h_cl_mult_off = cl_mult_off.download(s);
    if (!k.u_mult.alloc(k.h_cl_mult_off[C], true, s)) return false;
    if (n_trips) k_fill_mult<<<grid_of(n_trips), kB, 0, s>>>(t_row.p, t_path.p, t_cnt.p, new_row.p, k_cluster.p, cl_kmer_off.p, cl_mult_off.p, k.n_paths.p, k.u_mult.p, n_trips);
    // row columns, unique / multicluster lists
    DBuf<uint32_t> mc_flag, shared_id, is_uniq, is_multi, rank_u, rank_m;
    DBuf<uint64_t> cl_uniq, cl_multi, cl_uniq_off, cl_multi_off;
    if (!mc_flag.alloc(nk) || !shared_id.alloc(nk) || !is_uniq.alloc(Rk) || !is_multi.alloc(Rk) || !rank_u.alloc(Rk) || !rank_m.alloc(Rk) || !cl_uniq.alloc((size_t)C + 1, true, s) ||
        !cl_multi.alloc((size_t)C + 1, true, s) || !cl_uniq_off.alloc((size_t)C + 1) || !cl_multi_off.alloc((size_t)C + 1))
        return false;
    if (nk) k_flag_bit<<<grid_of(nk), kB, 0, s>>>(k.key_flags.p, 2, mc_flag.p, nk);
    if (!exclusive_sum(k.tmp, mc_flag.p, shared_id.p, nk, s)) return false;   // one shared KmerCounts record per multicluster key (KmerCounts.cpp:205-224)
    if (!k.u_has_counts.alloc(Rk) || !k.u_counts.alloc(Rk * S) || !k.u_ic.alloc(Rk * 2) || !k.u_shared.alloc(Rk)) return false;
    if (Rk) k_row_columns<<<grid_of(Rk), kB, 0, s>>>(k_cluster.p, k_key.p, k.key_flags.p, k.counts.p, k.ic.p, shared_id.p, S, k.u_has_counts.p, k.u_counts.p, k.u_ic.p, k.u_shared.p, is_uniq.p,
                                                    is_multi.p, cl_uniq.p, cl_multi.p, Rk);
    if (!exclusive_sum(k.tmp, is_uniq.p, rank_u.p, Rk, s) || !exclusive_sum(k.tmp, is_multi.p, rank_m.p, Rk, s)) return false;
    if (!exclusive_sum(k.tmp, cl_uniq.p, cl_uniq_off.p, (size_t)C + 1, s) || !exclusive_sum(k.tmp, cl_multi.p, cl_multi_off.p, (size_t)C + 1, s)) return false;
    k.h_cl_uniq_off = cl_uniq_off.download(s);
    k.h_cl_multi_off = cl_multi_off.download(s);
    if (!k.u_uniq_idx.alloc(k.h_cl_uniq_off[C]) || !k.u_multi_idx.alloc(k.h_cl_multi_off[C])) return false;
    if (Rk) {
        k_list_rows<<<grid_of(Rk), kB, 0, s>>>(is_uniq.p, rank_u.p, k_cluster.p, cl_kmer_off.p, k.u_uniq_idx.p, Rk);
        k_list_rows<<<grid_of(Rk), kB, 0, s>>>(is_multi.p, rank_m.p, k_cluster.p, cl_kmer_off.p, k.u_multi_idx.p, Rk);
    }
    k.h_multi_idx = k.u_multi_idx.download(s);
    // coverage bitmaps: (row, variant) -> haplotypes (KmerInfo::variant_haplotype_indices)
    const uint64_t NC = k.NC;
    DBuf<uint64_t> kmer_vh;
    if (!kmer_vh.alloc(Rk + 1, true, s) || !k.u_kmer_vh_off.alloc(Rk + 1)) return false;
    k.n_vh = 0; k.n_vh_bits = 0;
    if (NC) {
        DBuf<uint64_t> ev, ev_u;
        DBuf<uint32_t> ev_order, ev_cnt, ev_flag, ev_rank;
        if (!ev.alloc(NC) || !ev_order.alloc(NC)) return false;
        k_cov_events<<<grid_of(NC), kB, 0, s>>>(k.cov_occ.p, k.cov_var.p, occ_row.p, new_row.p, ev.p, NC);
        k_iota<<<grid_of(NC), kB, 0, s>>>(ev_order.p, NC);
        if (!sort_pairs(k.tmp, ev, ev_order, NC, s)) return false;
        size_t n_runs = 0;
        if (!run_length(k.tmp, ev.p, NC, ev_u, ev_cnt, n_runs, s)) return false;
        // the dropped rows sort last as one run of ~0
        uint64_t last = 0, n_valid = NC;
        if (n_runs) {
            CK(cudaMemcpyAsync(&last, ev_u.p + n_runs - 1, 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (last == ~0ull) {
                uint32_t c_last = 0;
                CK(cudaMemcpyAsync(&c_last, ev_cnt.p + n_runs - 1, 4, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                n_valid -= c_last; n_runs--;
            }
        }
        k.n_vh = n_runs;
        DBuf<uint64_t> e_h;
        if (!k.u_vh_var.alloc(n_runs) || !e_h.alloc(n_runs + 1, true, s) || !k.u_vh_bits_off.alloc(n_runs + 1)) return false;
        if (n_runs) k_ev_heads<<<grid_of(n_runs), kB, 0, s>>>(ev_u.p, k_cluster.p, k.n_paths.p, k.u_vh_var.p, e_h.p, kmer_vh.p, n_runs);
        if (!exclusive_sum(k.tmp, e_h.p, k.u_vh_bits_off.p, n_runs + 1, s)) return false;
        uint64_t tot = 0;
        CK(cudaMemcpyAsync(&tot, k.u_vh_bits_off.p + n_runs, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        k.n_vh_bits = tot;
        if (!k.u_vh_bits.alloc(tot, true, s)) return false;
        if (n_valid) {
            if (!ev_flag.alloc(n_valid) || !ev_rank.alloc(n_valid)) return false;
            k_pair_flags<<<grid_of(n_valid), kB, 0, s>>>(ev.p, ev_flag.p, n_valid);
            if (!inclusive_sum(k.tmp, ev_flag.p, ev_rank.p, n_valid, s)) return false;
            k_vh_bits<<<grid_of(n_valid), kB, 0, s>>>(ev_rank.p, ev_order.p, k.cov_occ.p, local_path.p, k.u_vh_bits_off.p, k.u_vh_bits.p, n_valid);
        }
    } else {
        if (!k.u_vh_var.alloc(0) || !k.u_vh_bits_off.alloc(1, true, s) || !k.u_vh_bits.alloc(0)) return false;
    }
    if (!exclusive_sum(k.tmp, kmer_vh.p, k.u_kmer_vh_off.p, Rk + 1, s)) return false;
    // haplotype -> allele (HaplotypeInfo::variant_allele_indices)
    k.h_cl_hapvar_off.assign((size_t)C + 1, 0);
    for (uint32_t c = 0; c < C; c++) k.h_cl_hapvar_off[c + 1] = k.h_cl_hapvar_off[c] + (uint64_t)k.h_n_paths[c] * (k.h_cl_var_off[c + 1] - k.h_cl_var_off[c]);
    DBuf<uint64_t> hapvar_off;
    if (!hapvar_off.upload(k.h_cl_hapvar_off.data(), (size_t)C + 1, s) || !k.u_hap_alleles.alloc(k.h_cl_hapvar_off[C] + 1, true, s)) return false;
    if (btg_path_alleles_dev(&k.walk, k.cl_var_off.p, k.var_nalleles.p, hapvar_off.p, k.u_hap_alleles.p, s) != BTG_OK) return false;
    BTG_LAUNCHED();
    CK(cudaStreamSynchronize(s));
    if (cudaGetLastError() != cudaSuccess) return false;
    nested_tables(k);
    return true;
}

}  // namespace

extern "C" {

btg_counter *btg_counter_create(const btg_counter_desc *d) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (!d || d->n_samples == 0 || d->n_samples > BTG_MAX_SAMPLES) { set_error("bad counter descriptor"); return nullptr; }
    auto *k = new btg_counter();
    k->S = d->n_samples; k->G = d->n_groups; k->C = d->n_clusters;
    k->s = ctx().stream;
    const uint32_t C = k->C, G = k->G;
    const uint64_t V = d->cl_vertex_off[C];
    k->V = V;
    copy_vec(k->h_gender, d->sample_gender, k->S);
    copy_vec(k->h_group_cluster_off, d->group_cluster_off, (size_t)G + 1);
    copy_vec(k->h_group_src_off, d->group_src_off, (size_t)G + 1);
    copy_vec(k->h_group_src, d->group_src, d->group_src_off[G]);
    copy_vec(k->h_group_edge_off, d->group_edge_off, (size_t)G + 1);
    copy_vec(k->h_group_edge_src, d->group_edge_src, d->group_edge_off[G]);
    copy_vec(k->h_group_edge_dst, d->group_edge_dst, d->group_edge_off[G]);
    copy_vec(k->h_cluster_idx, d->cluster_idx, C);
    copy_vec(k->h_cl_vertex_off, d->cl_vertex_off, (size_t)C + 1);
    copy_vec(k->h_cl_var_off, d->cl_var_off, (size_t)C + 1);
    k->n_var = d->cl_var_off[C];
    copy_vec(k->h_var_nalleles, d->var_nalleles, k->n_var);
    copy_vec(k->h_var_dep, d->var_dep, k->n_var);
    copy_vec(k->h_n_paths, d->n_paths, C);
    copy_vec(k->h_v_var, d->v_var, V);
    copy_vec(k->h_v_refvar_off, d->v_refvar_off, V + 1);
    copy_vec(k->h_v_refvar, d->v_refvar, d->v_refvar_off[V]);
    if (d->v_nested) copy_vec(k->h_v_nested, d->v_nested, V);
    k->h_cl_path_off.assign((size_t)C + 1, 0);
    k->h_path_mem_off.assign((size_t)C + 1, 0);
    k->h_cl_group.assign(C, 0);
    for (uint32_t g = 0; g < G; g++) for (uint64_t c = d->group_cluster_off[g]; c < d->group_cluster_off[g + 1]; c++) k->h_cl_group[c] = g;
    for (uint32_t c = 0; c < C; c++) {
        k->h_cl_path_off[c + 1] = k->h_cl_path_off[c] + d->n_paths[c];
        k->h_path_mem_off[c + 1] = k->h_path_mem_off[c] + (uint64_t)d->n_paths[c] * (d->cl_vertex_off[c + 1] - d->cl_vertex_off[c]);
    }
    k->P = k->h_cl_path_off[C];
    copy_vec(k->h_path_mem, d->path_mem, k->h_path_mem_off[C]);
    std::vector<uint32_t> path_cluster(k->P);
    for (uint32_t c = 0; c < C; c++) for (uint64_t p = k->h_cl_path_off[c]; p < k->h_cl_path_off[c + 1]; p++) path_cluster[p] = c;
    cudaStream_t s = k->s;
    bool ok = k->cl_vertex_off.upload(d->cl_vertex_off, (size_t)C + 1, s) && k->v_seq_off.upload(d->v_seq_off, V + 1, s) && k->seq.upload(d->seq, d->v_seq_off[V], s) &&
              k->v_flags.upload(d->v_flags, V, s) && k->v_var.upload(d->v_var, V, s) && k->v_allele.upload(d->v_allele, V, s) && k->v_refvar_off.upload(d->v_refvar_off, V + 1, s) &&
              k->v_refvar.upload(d->v_refvar, d->v_refvar_off[V], s) && k->cl_path_off.upload(k->h_cl_path_off.data(), (size_t)C + 1, s) &&
              k->path_mem_off.upload(k->h_path_mem_off.data(), (size_t)C + 1, s) && k->path_mem.upload(d->path_mem, k->h_path_mem_off[C], s) &&
              k->path_cluster.upload(path_cluster.data(), k->P, s) && k->n_paths.upload(d->n_paths, C, s) && k->cl_group.upload(k->h_cl_group.data(), C, s) &&
              k->cl_var_off.upload(d->cl_var_off, (size_t)C + 1, s) && k->var_nalleles.upload(d->var_nalleles, k->n_var, s);
    if (!ok || cudaStreamSynchronize(s) != cudaSuccess) { set_error("counter: upload failed (%s)", cudaGetErrorString(cudaGetLastError())); delete k; return nullptr; }
    k->walk = btg_pathwalk_desc{C, k->P, k->cl_vertex_off.p, k->v_seq_off.p, k->seq.p, k->v_flags.p, k->v_var.p, k->v_allele.p, k->v_refvar_off.p, k->v_refvar.p,
                                k->cl_path_off.p, k->path_mem_off.p, k->path_mem.p, k->path_cluster.p};
    return k;
}

void btg_counter_free(btg_counter *k) {
    if (!k) return;
    cudaStreamSynchronize(k->s);
    btg_table_set_index_dev(nullptr, 0);
    delete k;
}

int btg_counter_count_path_kmers(btg_counter *k, uint64_t *n_path_kmers_out) {
    BTG_REQUIRE_INIT();
    if (!k) { set_error("null argument"); return BTG_EINVAL; }
    if (!count_path_kmers(*k)) { if (!*btg_last_error()) set_error("countPathKmers failed (%s)", cudaGetErrorString(cudaGetLastError())); return BTG_ECUDA; }
    if (n_path_kmers_out) *n_path_kmers_out = k->n_keys;
    return BTG_OK;
}

int btg_counter_count_intercluster_kmers(btg_counter *k, const char *seq_dev, size_t len, int is_decoy, uint32_t ploidy_female, uint32_t ploidy_male) {
    BTG_REQUIRE_INIT();
    if (!k || !k->counted) { set_error("countPathKmers has not run"); return BTG_ESTATE; }
    if (len == 0 || k->n_keys == 0) return BTG_OK;
    k->use_index();
    return btg_table_scan_region_dev(k->kw0.p, k->kw1.p, (int64_t)k->n_keys, seq_dev, len, is_decoy, ploidy_female, ploidy_male, k->ic.p, k->max_mult.p, k->decoy.p, k->has_record.p, k->s);
}

int btg_counter_parse_sample_kmers(btg_counter *k, uint32_t sample_idx, const uint64_t *kmers_dev, const uint8_t *counts_dev, size_t n) {
    BTG_REQUIRE_INIT();
    if (!k || !k->counted) { set_error("countPathKmers has not run"); return BTG_ESTATE; }
    if (n == 0 || k->n_keys == 0) return BTG_OK;
    k->use_index();
    return btg_table_add_sample_kmers_dev(k->kw0.p, k->kw1.p, (int64_t)k->n_keys, kmers_dev, counts_dev, n, k->S, sample_idx, k->counts.p, k->has_record.p, k->s);
}

btg_unit *btg_counter_build_unit(btg_counter *k, const btg_bloom *multigroup, const uint8_t *group_ploidy) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (!k || !k->counted || !group_ploidy) { set_error("countPathKmers has not run, or null argument"); return nullptr; }
    if (!build_unit_arrays(*k, multigroup)) { if (!*btg_last_error()) set_error("classifyPathKmers failed (%s)", cudaGetErrorString(cudaGetLastError())); return nullptr; }
    const uint32_t C = k->C, G = k->G;
    // units with multicluster k-mers pass k_shared / k_has_counts through the host (btg_unit_upload_dev validates them there)
    const bool has_multi = k->h_cl_multi_off[C] != 0;
    std::vector<uint32_t> h_shared;
    std::vector<uint8_t> h_has;
    if (has_multi) { h_shared = k->u_shared.download(k->s); h_has = k->u_has_counts.download(k->s); }
    btg_unit_desc host{}, dev{};
    host.n_samples = k->S; host.n_groups = G; host.n_clusters = C;
    host.sample_gender = k->h_gender.data(); host.group_ploidy = group_ploidy;
    host.group_cluster_off = k->h_group_cluster_off.data(); host.group_src_off = k->h_group_src_off.data(); host.group_src = k->h_group_src.data();
    host.group_edge_off = k->h_group_edge_off.data(); host.group_edge_src = k->h_group_edge_src.data(); host.group_edge_dst = k->h_group_edge_dst.data();
    host.cluster_idx = k->h_cluster_idx.data(); host.cl_nhap = k->h_n_paths.data();
    host.cl_kmer_off = k->h_cl_kmer_off.data(); host.cl_var_off = k->h_cl_var_off.data(); host.cl_mult_off = k->h_cl_mult_off.data();
    host.cl_uniq_off = k->h_cl_uniq_off.data(); host.cl_multi_off = k->h_cl_multi_off.data(); host.multi_idx = k->h_multi_idx.data();
    host.cl_hapvar_off = k->h_cl_hapvar_off.data(); host.var_nalleles = k->h_var_nalleles.data(); host.var_dep = k->h_var_dep.data();
    host.hap_nested_off = k->h_hap_nested_off.data(); host.hap_nested = k->h_hap_nested.data(); host.cl_dep_off = k->h_cl_dep_off.data();
    host.dep_cluster = k->h_dep_cluster.data(); host.dep_var_off = k->h_dep_var_off.data(); host.dep_var = k->h_dep_var.data();
    dev.mult = k->u_mult.p; dev.k_counts = k->u_counts.p; dev.k_ic = k->u_ic.p; dev.uniq_idx = k->u_uniq_idx.p;
    dev.kmer_vh_off = k->u_kmer_vh_off.p; dev.vh_var = k->u_vh_var.p; dev.vh_bits_off = k->u_vh_bits_off.p; dev.vh_bits = k->u_vh_bits.p; dev.hap_alleles = k->u_hap_alleles.p;
    if (has_multi) { host.k_shared = h_shared.data(); host.k_has_counts = h_has.data(); }
    else { dev.k_shared = k->u_shared.p; dev.k_has_counts = k->u_has_counts.p; }
    return btg_unit_upload_dev(&host, &dev, k->n_vh, k->n_vh_bits);
}

int btg_counter_fit_nb(btg_counter *k, const char *seq_dev, size_t len, uint32_t ploidy_female, uint32_t ploidy_male, const uint64_t *const *sample_kmers_dev,
                       const uint8_t *const *sample_counts_dev, const size_t *sample_n, const uint64_t *parameter_kmers, size_t n_parameter_kmers, uint32_t random_seed,
                       uint64_t max_parameter_kmers, double *nb_p_out, double *nb_size_out, uint32_t *modal_multiplicity_out, uint64_t *n_modal_kmers_out) {
    BTG_REQUIRE_INIT();
    if (!k || !k->counted || !nb_p_out || !nb_size_out || (k->S && (!sample_kmers_dev || !sample_counts_dev || !sample_n))) { set_error("countPathKmers has not run, or null argument"); return BTG_EINVAL; }
    cudaStream_t s = k->s;
    const uint32_t S = k->S;
    auto fail = [&](const char *what) { if (!*btg_last_error()) set_error("NB fit: %s (%s)", what, cudaGetErrorString(cudaGetLastError())); return BTG_ECUDA; };
    // every valid canonical k-mer of the inter-cluster regions
    DBuf<uint64_t> km, kmv;
    DBuf<uint8_t> valid;
    DBuf<uint32_t> flag, rank;
    if (!km.alloc(len * 2) || !valid.alloc(len) || !flag.alloc(len) || !rank.alloc(len)) return fail("allocation");
    size_t n_valid = 0;
    if (len) {
        if (btg_scan_sequence_dev(seq_dev, len, km.p, valid.p, s) != BTG_OK) return BTG_ECUDA;
        k_valid_flags<<<grid_of(len), kB, 0, s>>>(valid.p, flag.p, len);
        if (!exclusive_sum(k->tmp, flag.p, rank.p, len, s)) return fail("scan");
        uint32_t last_r = 0, last_f = 0;
        cudaMemcpyAsync(&last_r, rank.p + len - 1, 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(&last_f, flag.p + len - 1, 4, cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) return fail("scan");
        n_valid = (size_t)last_r + last_f;
        if (!kmv.alloc(std::max<size_t>(n_valid, 1) * 2)) return fail("allocation");
        k_compact_kmers<<<grid_of(len), kB, 0, s>>>((const ulonglong2 *)km.p, flag.p, rank.p, (ulonglong2 *)kmv.p, len);
    }
    km.release(); valid.release(); flag.release(); rank.release();
    // distinct k-mers with their genomic multiplicity; not path k-mers; on the list / Bernoulli
    DBuf<int64_t> u_lo, u_hi;
    DBuf<uint32_t> starts;
    size_t nd = 0;
    if (!distinct_keys(*k, kmv.p, n_valid, u_lo, u_hi, &starts, nd)) return fail("sort");
    kmv.release();
    DBuf<uint64_t> cand;
    DBuf<int64_t> path_idx, list_idx, l_lo, l_hi;
    if (!cand.alloc(std::max<size_t>(nd, 1) * 2) || !path_idx.alloc(nd)) return fail("allocation");
    if (nd) {
        if (btg_table_keys_to_kmers_dev(u_lo.p, u_hi.p, nd, cand.p, s) != BTG_OK) return BTG_ECUDA;
        k->use_index();
        if (btg_table_lookup_dev(k->kw0.p, k->kw1.p, (int64_t)k->n_keys, cand.p, nd, path_idx.p, s) != BTG_OK) return BTG_ECUDA;
    }
    if (parameter_kmers) {
        DBuf<uint64_t> lk;
        size_t nl = 0;
        if (!lk.upload(parameter_kmers, n_parameter_kmers * 2, s) || !distinct_keys(*k, lk.p, n_parameter_kmers, l_lo, l_hi, nullptr, nl) || !list_idx.alloc(nd)) return fail("parameter k-mer list");
        if (nd) {
            btg_table_set_index_dev(nullptr, 0);
            if (btg_table_lookup_dev(l_lo.p, l_hi.p, (int64_t)nl, cand.p, nd, list_idx.p, s) != BTG_OK) return BTG_ECUDA;
        }
    }
    DBuf<uint32_t> sel, occ, srank;
    if (!sel.alloc(nd) || !occ.alloc(nd) || !srank.alloc(nd)) return fail("allocation");
    const double frac = std::min(1.0, 3.0 * (double)max_parameter_kmers / (double)std::max<size_t>(n_valid, 1));   // at most 3 x the cap is drawn (main.cpp:330-333)
    if (nd) k_param_select<<<grid_of(nd), kB, 0, s>>>(starts.p, nd, n_valid, path_idx.p, parameter_kmers ? list_idx.p : nullptr, u_lo.p, u_hi.p, random_seed, frac, sel.p, occ.p);
    if (!exclusive_sum(k->tmp, sel.p, srank.p, nd, s)) return fail("scan");
    size_t n_par = 0;
    if (nd) {
        uint32_t a = 0, b = 0;
        cudaMemcpyAsync(&a, srank.p + nd - 1, 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(&b, sel.p + nd - 1, 4, cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) return fail("scan");
        n_par = (size_t)a + b;
    }
    DBuf<int64_t> p_lo, p_hi;
    DBuf<uint32_t> p_occ;
    DBuf<uint8_t> p_counts, p_rec;
    if (!p_lo.alloc(n_par) || !p_hi.alloc(n_par) || !p_occ.alloc(n_par) || !p_counts.alloc(n_par * S + 4, true, s) || !p_rec.alloc(n_par + 4, true, s)) return fail("allocation");
    if (nd) k_compact_params<<<grid_of(nd), kB, 0, s>>>(sel.p, srank.p, u_lo.p, u_hi.p, occ.p, p_lo.p, p_hi.p, p_occ.p, nd);
    if (!parameter_kmers && n_par > max_parameter_kmers) n_par = max_parameter_kmers;   // the cap of parameter_kmers.fa.gz (main.cpp:543-581): the first in key order
    btg_table_set_index_dev(nullptr, 0);                                                   // the parameter k-mers are a second, un-indexed table
    for (uint32_t si = 0; si < S && n_par; si++)
        if (btg_table_add_sample_kmers_dev(p_lo.p, p_hi.p, (int64_t)n_par, sample_kmers_dev[si], sample_counts_dev[si], sample_n[si], S, si, p_counts.p, p_rec.p, s) != BTG_OK) return BTG_ECUDA;
    // modal multiplicity and the moments of its class, per sample (KmerHash.cpp:257-347, CountDistribution.cpp:66-141), in f64 on the host
    std::vector<uint32_t> h_occ(n_par);
    std::vector<uint8_t> h_cnt(n_par * S);
    if (n_par) { cudaMemcpyAsync(h_occ.data(), p_occ.p, n_par * 4, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(h_cnt.data(), p_counts.p, n_par * S, cudaMemcpyDeviceToHost, s); }
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail("copy");
    for (uint32_t si = 0; si < S; si++) {
        const uint32_t ploidy = k->h_gender[si] == 0 ? ploidy_female : ploidy_male;
        uint64_t hist[33] = {0};
        for (size_t i = 0; i < n_par; i++) { const uint64_t m = std::min<uint64_t>(255, (uint64_t)h_occ[i] * ploidy); if (m >= 1 && m <= 32) hist[m]++; }   // max_nb_kmer_multiplicity = 32
        uint32_t modal = 0; uint64_t best = 0;
        for (uint32_t m = 1; m <= 32; m++) if (hist[m] > best) { best = hist[m]; modal = m; }
        if (best == 0) { set_error("sample %u: no parameter k-mer with genomic multiplicity 1..32 (ploidy %u on this contig, %zu parameter k-mers): the negative binomial cannot be fitted", si, ploidy, n_par); return BTG_ESTATE; }
        if (best < 2) { set_error("sample %u: %llu parameter k-mer(s) at the modal multiplicity %u: mean and variance undefined", si, (unsigned long long)best, modal); return BTG_ESTATE; }
        double sum = 0;
        for (size_t i = 0; i < n_par; i++) if (std::min<uint64_t>(255, (uint64_t)h_occ[i] * ploidy) == modal) sum += h_cnt[i * S + si];
        const double mean = sum / (double)best;
        double ss = 0;
        for (size_t i = 0; i < n_par; i++) if (std::min<uint64_t>(255, (uint64_t)h_occ[i] * ploidy) == modal) { const double dlt = h_cnt[i * S + si] - mean; ss += dlt * dlt; }
        btg_nb_moments_to_parameters(mean, ss / (double)(best - 1), modal, nb_p_out + si, nb_size_out + si);
        if (modal_multiplicity_out) modal_multiplicity_out[si] = modal;
        if (n_modal_kmers_out) n_modal_kmers_out[si] = best;
    }
    return BTG_OK;
}

// sizes (elements) and copies of the arrays of the last btg_counter_build_unit, by the field names of btg_unit_desc (tests, fixtures, hosts that
// want the descriptor).  Also "key_flags" (bit0 record, bit1 multicluster, bit2 multigroup, bit3 excluded) and "key_lo" / "key_hi".
int64_t btg_counter_array(btg_counter *k, const char *field, void *out, uint64_t out_bytes) {
    if (!k || !field) { set_error("null argument"); return BTG_EINVAL; }
    const std::string f(field);
    const void *src = nullptr; size_t n = 0, esz = 1; bool on_dev = true;
    auto dv = [&](auto &b, size_t count) { src = b.p; n = count; esz = sizeof(*b.p); on_dev = true; };
    auto hv = [&](auto &v) { src = v.data(); n = v.size(); esz = sizeof(v[0]); on_dev = false; };
    const uint32_t C = k->C;
    if (f == "mult") dv(k->u_mult, k->h_cl_mult_off.empty() ? 0 : k->h_cl_mult_off[C]);
    else if (f == "k_has_counts") dv(k->u_has_counts, k->n_rows);
    else if (f == "k_counts") dv(k->u_counts, k->n_rows * k->S);
    else if (f == "k_ic") dv(k->u_ic, k->n_rows * 2);
    else if (f == "k_shared") dv(k->u_shared, k->n_rows);
    else if (f == "uniq_idx") dv(k->u_uniq_idx, k->h_cl_uniq_off.empty() ? 0 : k->h_cl_uniq_off[C]);
    else if (f == "kmer_vh_off") dv(k->u_kmer_vh_off, k->n_rows + 1);
    else if (f == "vh_var") dv(k->u_vh_var, k->n_vh);
    else if (f == "vh_bits_off") dv(k->u_vh_bits_off, k->n_vh + 1);
    else if (f == "vh_bits") dv(k->u_vh_bits, k->n_vh_bits);
    else if (f == "hap_alleles") dv(k->u_hap_alleles, k->h_cl_hapvar_off.empty() ? 0 : k->h_cl_hapvar_off[C]);
    else if (f == "key_flags") dv(k->key_flags, k->n_keys);
    else if (f == "key_lo") dv(k->kw0, k->n_keys);
    else if (f == "key_hi") dv(k->kw1, k->n_keys);
    else if (f == "cl_kmer_off") hv(k->h_cl_kmer_off);
    else if (f == "cl_mult_off") hv(k->h_cl_mult_off);
    else if (f == "cl_uniq_off") hv(k->h_cl_uniq_off);
    else if (f == "cl_multi_off") hv(k->h_cl_multi_off);
    else if (f == "multi_idx") hv(k->h_multi_idx);
    else if (f == "cl_hapvar_off") hv(k->h_cl_hapvar_off);
    else if (f == "hap_nested_off") hv(k->h_hap_nested_off);
    else if (f == "hap_nested") hv(k->h_hap_nested);
    else if (f == "cl_dep_off") hv(k->h_cl_dep_off);
    else if (f == "dep_cluster") hv(k->h_dep_cluster);
    else if (f == "dep_var_off") hv(k->h_dep_var_off);
    else if (f == "dep_var") hv(k->h_dep_var);
    else { set_error("unknown counter array '%s'", field); return BTG_EINVAL; }
    if (!out) return (int64_t)n;
    if (out_bytes < n * esz) { set_error("buffer too small for '%s'", field); return BTG_EINVAL; }
    if (n == 0) return 0;
    if (on_dev) {
        if (cudaMemcpyAsync(out, src, n * esz, cudaMemcpyDeviceToHost, k->s) != cudaSuccess || cudaStreamSynchronize(k->s) != cudaSuccess) { set_error("copy of '%s' failed", field); return BTG_ECUDA; }
    } else memcpy(out, src, n * esz);
    return (int64_t)n;
}

}  // extern "C"
