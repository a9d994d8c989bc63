// table.cu — path k-mer enumeration and the exact k-mer count table (k-mer match, genotype side).
//
// Replaces, for the k-mers that can influence inference (the path k-mers of the unit):
//   VariantClusterGraph::countPathKmers / classifyPathKmers / getHaplotypeCandidates' path walks
//       src/bayesTyper/VariantClusterGraph.cpp:800-846, 848-939, 941-1135 (+ updateVariantPathIndices :1137-1184)
//   KmerCounter::parseSampleKmersCallBack (sample k-mer stream -> counts)        src/bayesTyper/KmerCounter.cpp:388-429
//   KmerCounter::countInterclusterKmersCallback (genome scan -> multiplicities)  src/bayesTyper/KmerCounter.cpp:291-334
//   KmerCounts::{addSampleCount,addInterclusterMultiplicity}                     src/bayesTyper/KmerCounts.cpp:98-118,178-189
//
// The reference funnels every k-mer through an in-memory Bloom filter of the path k-mers and then a mutex-guarded
// hash of vectors.  Filter false positives only ever create table entries that no path k-mer looks up, so the
// device keeps an EXACT table instead: the distinct path k-mers as a sorted key array (built by the caller from
// the emitted occurrences) that the 17 B/record sample stream and the 1 B/nt genome scan probe directly.
// Keys are kept in the internal MSB-first form of kmer.cuh (V = hi:lo, nucleotide 0 on top), ascending, i.e. in
// lexicographic order of the k-mer strings — the order in which a KMC1 database lists its records
// (external/kmc_api/kmc_file.cpp:428-515: prefix LUT, then sorted suffixes; a KMC2 database lists signature bin by
// signature bin, each bin in this order), so a sample's stream walks the table front to back (once per bin).  The two key columns are signed 64-bit for the host glue's sort: key_hi = hi (46 bits, >= 0) and
// key_lo = lo ^ 2^63 (biased, so that signed order == unsigned order of lo).
#include "common.cuh"
#include "kmer.cuh"

using namespace btg;

namespace {

constexpr uint32_t NONE16 = 0xFFFF;
constexpr uint64_t kLoBias = 0x8000000000000000ULL;  // key_lo = lo ^ kLoBias

struct TableKey { int64_t lo, hi; };  // (key_lo, key_hi) of one k-mer
__device__ __forceinline__ TableKey key_of(const Kmer128 &v) { return TableKey{(int64_t)(v.lo ^ kLoBias), (int64_t)v.hi}; }
// a packed k-mer in the boundary layout (the ABI's 2 x uint64) -> table key
__device__ __forceinline__ TableKey key_of_boundary(int64_t w0, int64_t w1) { return key_of(from_boundary((uint64_t)w0, (uint64_t)w1)); }
constexpr int kMaxRunning = 48;  // variants whose window covers the current k-mer (running_variants)

struct PathWalkGraphs {
    uint32_t C;
    const uint64_t *cl_vertex_off, *v_seq_off;
    const uint8_t *seq, *v_flags;
    const uint16_t *v_var, *v_allele;   // variant_allele_idx (0xFFFF = none)
    const uint64_t *v_refvar_off;
    const uint16_t *v_refvar;           // reference_variant_indices
    const uint64_t *cl_path_off;        // [C+1] first best path of each cluster (global path index)
    const uint64_t *path_mem_off;       // [C+1] byte offset of the cluster's membership rows
    const uint8_t *path_mem;            // path x vertex membership (1 = on path)
    const uint32_t *path_cluster;       // [P] cluster of each path
};

struct Running { uint32_t var, allele, first, second; };

// One best path: canonical k-mer of every window (reset at disconnected vertices) + which variants each window
// covers.  EMIT = false: count only.
template <bool EMIT>
__global__ void __launch_bounds__(128) k_walk_paths(PathWalkGraphs g, uint64_t n_paths, uint32_t *__restrict__ n_occ, uint32_t *__restrict__ n_cov,
                                                    const uint64_t *__restrict__ occ_off, const uint64_t *__restrict__ cov_off,
                                                    int64_t *__restrict__ key_w0, int64_t *__restrict__ key_w1, uint32_t *__restrict__ occ_path,
                                                    uint32_t *__restrict__ occ_nt, int64_t *__restrict__ cov_occ, uint16_t *__restrict__ cov_var,
                                                    uint32_t *__restrict__ status) {
    const uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const uint32_t c = g.path_cluster[p];
    const uint64_t v0 = g.cl_vertex_off[c];
    const uint32_t V = (uint32_t)(g.cl_vertex_off[c + 1] - v0);
    const uint8_t *mem = g.path_mem + g.path_mem_off[c] + (p - g.cl_path_off[c]) * V;
    Roller roll;
    roll.reset();
    Running run[kMaxRunning];
    int n_run = 0;
    uint32_t num_nt = 0, occ = 0, cov = 0;
    uint64_t o_base = EMIT ? occ_off[p] : 0, c_base = EMIT ? cov_off[p] : 0;
    bool overflow = false;
    for (uint32_t v = 0; v < V; v++) {
        if (!mem[v]) continue;
        const uint64_t s0 = g.v_seq_off[v0 + v];
        const uint32_t len = (uint32_t)(g.v_seq_off[v0 + v + 1] - s0);
        const uint32_t var = g.v_var[v0 + v], allele = g.v_allele[v0 + v];
        if (var != NONE16) {  // VariantClusterGraph.cpp:987-999
            int f = -1;
            for (int i = 0; i < n_run; i++) if (run[i].var == var && run[i].allele == allele) { f = i; break; }
            if (f < 0) {
                if (n_run == kMaxRunning) { overflow = true; break; }
                f = n_run++;
                run[f] = Running{var, allele, num_nt + (g.v_flags[v0 + v] & 1u), num_nt + K - 1};
            }
            run[f].second += len;
        }
        for (uint64_t e = g.v_refvar_off[v0 + v]; e < g.v_refvar_off[v0 + v + 1]; e++) {  // :1001-1012
            const uint32_t rv = g.v_refvar[e];
            for (int i = 0; i < n_run; i++) if (run[i].var == rv && run[i].allele == 0) { run[i].second += len; break; }
        }
        if (g.v_flags[v0 + v] & 2u) roll.reset();
        for (uint32_t i = 0; i < len; i++) {
            if (roll.push(g.seq[s0 + i])) {
                // expire windows (updateVariantPathIndices :1141-1149), then record the covering variants
                int w = 0;
                for (int j = 0; j < n_run; j++) if (run[j].second > num_nt) run[w++] = run[j];
                n_run = w;
                if (EMIT) {
                    const Kmer128 cn = roll.canonical();
                    key_w0[o_base + occ] = (int64_t)(cn.lo ^ kLoBias);
                    key_w1[o_base + occ] = (int64_t)cn.hi;
                    occ_path[o_base + occ] = (uint32_t)p;
                    occ_nt[o_base + occ] = num_nt;
                }
                for (int j = 0; j < n_run; j++)
                    if (run[j].first <= num_nt) {
                        if (EMIT) { cov_occ[c_base + cov] = (int64_t)(o_base + occ); cov_var[c_base + cov] = (uint16_t)run[j].var; }
                        cov++;
                    }
                occ++;
            }
            num_nt++;
        }
    }
    if (overflow) status[c] = 2;
    if (!EMIT) { n_occ[p] = occ; n_cov[p] = cov; }
}

// haplotype -> allele table: HaplotypeInfo::variant_allele_indices (VariantClusterGraph.cpp:983-992,1091-1098)
__global__ void __launch_bounds__(128) k_path_alleles(PathWalkGraphs g, uint64_t n_paths, const uint64_t *__restrict__ cl_var_off,
                                                      const uint16_t *__restrict__ var_nalleles, const uint64_t *__restrict__ hapvar_off,
                                                      uint16_t *__restrict__ hap_alleles) {
    const uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const uint32_t c = g.path_cluster[p];
    const uint64_t v0 = g.cl_vertex_off[c];
    const uint32_t V = (uint32_t)(g.cl_vertex_off[c + 1] - v0);
    const uint32_t nvar = (uint32_t)(cl_var_off[c + 1] - cl_var_off[c]);
    const uint8_t *mem = g.path_mem + g.path_mem_off[c] + (p - g.cl_path_off[c]) * V;
    uint16_t *out = hap_alleles + hapvar_off[c] + (p - g.cl_path_off[c]) * nvar;
    for (uint32_t i = 0; i < nvar; i++) out[i] = NONE16;
    for (uint32_t v = 0; v < V; v++) {
        if (!mem[v]) continue;
        const uint32_t var = g.v_var[v0 + v];
        if (var != NONE16 && !(g.v_flags[v0 + v] & 2u)) out[var] = g.v_allele[v0 + v];
    }
    for (uint32_t i = 0; i < nvar; i++) if (out[i] == NONE16) out[i] = var_nalleles[cl_var_off[c] + i] - 1;
}

// ---- exact table probes ------------------------------------------------------------------------
// Optional prefix index (like KMC's prefix LUT): lut[b] = first key whose top `lut_bits` bits of the 46-bit key_hi
// (= the first lut_bits/2 nucleotides) are >= b; narrows the search to the bucket (typically 1-2 keys) at the cost of
// one 8 B read.
struct TableIndex {
    const int64_t *lut;  // [2^lut_bits + 1] or nullptr
    int shift;           // 46 - lut_bits
};

__device__ __forceinline__ int64_t table_find(const int64_t *__restrict__ kw0, const int64_t *__restrict__ kw1, int64_t n, int64_t w0, int64_t w1,
                                              const TableIndex ix = TableIndex{nullptr, 0}) {
    int64_t lo = 0, hi = n;
    if (ix.lut) {
        const int64_t b = w1 >> ix.shift;
        lo = __ldg(ix.lut + b);
        hi = __ldg(ix.lut + b + 1);
    }
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t m1 = __ldg(kw1 + mid);
        bool less;  // key[mid] < (w1, w0)
        if (m1 != w1) less = m1 < w1;
        else less = __ldg(kw0 + mid) < w0;
        if (less) lo = mid + 1; else hi = mid;
    }
    if (lo < end && __ldg(kw1 + lo) == w1 && __ldg(kw0 + lo) == w0) return lo;
    return -1;
}

__device__ __forceinline__ void sat_add_u8(uint8_t *p, uint32_t add) {  // updateMultiplicity / addSampleCount (KmerCounts.cpp:161-189)
    uint32_t *word = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(3));
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
    uint32_t old = *word, assumed;
    do {
        assumed = old;
        const uint32_t cur = (assumed >> sh) & 0xFFu;
        const uint32_t nv = (255u - cur) <= add ? 255u : cur + add;
        old = atomicCAS(word, assumed, (assumed & ~(0xFFu << sh)) | (nv << sh));
    } while (old != assumed);
}

// several saturating byte adds that fall into the same aligned 32-bit word are applied by ONE lane in one CAS loop: with
// S < 4 neighbouring keys share a word, and a warp walking the table front to back would otherwise collide with itself
__device__ __forceinline__ void sat_add_u8_warp(bool has, uint8_t *p, uint32_t add) {
    const unsigned active = __ballot_sync(0xFFFFFFFFu, has);
    if (!has) return;
    uint32_t *word = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(3));
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
    const unsigned peers = __match_any_sync(active, reinterpret_cast<unsigned long long>(word));
    const int leader = __ffs(peers) - 1;
    uint32_t shs[4], adds[4];  // at most 4 bytes per word
    int n = 0;
    for (unsigned m = peers; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        const uint32_t v = __shfl_sync(peers, (add << 8) | sh, src);
        if (n < 4) { shs[n] = v & 0xFFu; adds[n] = v >> 8; }
        n++;
    }
    if ((int)(threadIdx.x & 31u) != leader) return;
    if (n > 4) n = 4;  // cannot happen: distinct lanes of one group address distinct (key, sample) bytes
    uint32_t old = *word, assumed;
    do {
        assumed = old;
        uint32_t nw = assumed;
        for (int i = 0; i < n; i++) {
            const uint32_t cur = (nw >> shs[i]) & 0xFFu;
            const uint32_t nv = (255u - cur) <= adds[i] ? 255u : cur + adds[i];
            nw = (nw & ~(0xFFu << shs[i])) | (nv << shs[i]);
        }
        old = atomicCAS(word, assumed, nw);
    } while (old != assumed);
}

// KmerCounter::parseSampleKmersCallBack as a tiled merge-join: the sample's (k-mer, count) records stream past the table
// once (17 B/record).  A warp takes a tile of 128 consecutive records (coalesced 16 B loads), reduces their prefix-index
// buckets to [bmin, bmax] and stages the table keys of that bucket range — with a KMC-ordered stream a few dozen
// consecutive keys — in shared memory with coalesced loads.  Every record is then located by a branch-free
// lower-bound search in shared memory with a warp-uniform trip count (no divergence, no dependent global loads), its
// count is accumulated per key in shared memory, and the tile's keys are updated in the table by one coalesced pass.
// Traffic = records + keys + index ends once = the algorithmic 17 B/record + 16 B/key.  Tiles whose bucket range
// exceeds kTileKeys keys (unsorted or sparse stream, no index) probe the table directly, record by record.
constexpr int kTileLanes = 4;                    // records per lane and tile -> 128 records per warp tile
constexpr int kTileKeys = 256;                   // staged keys per tile
constexpr int kStreamWarps = 8;                  // warps per CTA
__global__ void __launch_bounds__(kStreamWarps * 32) k_table_add_sample(const int64_t *__restrict__ kw0, const int64_t *__restrict__ kw1, int64_t n_keys,
                                                          const longlong2 *__restrict__ kmers, const uint8_t *__restrict__ counts, size_t n,
                                                          uint32_t S, uint32_t sample, uint8_t *table_counts, uint8_t *has_record, TableIndex ix) {
    __shared__ int64_t s_hi[kStreamWarps][kTileKeys];
    __shared__ int64_t s_lo[kStreamWarps][kTileKeys];
    __shared__ uint32_t s_acc[kStreamWarps][kTileKeys];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const size_t n_tiles = (n + 32 * kTileLanes - 1) / (32 * kTileLanes);
    const size_t warp_id = (size_t)blockIdx.x * kStreamWarps + w, n_warps = (size_t)gridDim.x * kStreamWarps;
    for (size_t tile = warp_id; tile < n_tiles; tile += n_warps) {  // warp-uniform trip count
        const size_t r0 = tile * (32 * kTileLanes);
        TableKey q[kTileLanes];
        bool valid[kTileLanes];
        uint32_t bmin = 0xFFFFFFFFu, bmax = 0;
#pragma unroll
        for (int j = 0; j < kTileLanes; j++) {
            const size_t i = r0 + (size_t)j * 32 + lane;
            valid[j] = i < n;
            if (valid[j]) {
                const longlong2 k = __ldg(kmers + i);
                q[j] = key_of_boundary(k.x, k.y);
                if (ix.lut) { const uint32_t b = (uint32_t)(q[j].hi >> ix.shift); bmin = min(bmin, b); bmax = max(bmax, b); }
            } else q[j] = TableKey{0, 0};
        }
        int64_t t0 = 0, t1 = n_keys;
        if (ix.lut) {
            bmin = __reduce_min_sync(0xFFFFFFFFu, bmin);
            bmax = __reduce_max_sync(0xFFFFFFFFu, bmax);
            int64_t v = 0;
            if (lane == 0) v = __ldg(ix.lut + bmin); else if (lane == 1) v = __ldg(ix.lut + bmax + 1);
            t0 = __shfl_sync(0xFFFFFFFFu, v, 0);
            t1 = __shfl_sync(0xFFFFFFFFu, v, 1);
        }
        const int64_t nk = t1 - t0;
        if (nk < kTileKeys) {
            uint32_t top = 1;
            while ((int64_t)top <= nk) top <<= 1;  // warp-uniform; the search below covers top - 1 >= nk slots
            for (uint32_t i = lane; i < top; i += 32) {  // slots past the last key hold +inf: no bound checks in the search
                const bool in = (int64_t)i < nk;
                s_hi[w][i] = in ? __ldg(kw1 + t0 + i) : INT64_MAX;
                s_lo[w][i] = in ? __ldg(kw0 + t0 + i) : INT64_MAX;
                s_acc[w][i] = 0;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kTileLanes; j++) {
                // lower bound on key_hi alone (the first 23 nucleotides); ties on key_hi are walked linearly
                uint32_t pos = 0;
                for (uint32_t step = top >> 1; step; step >>= 1)
                    if (s_hi[w][pos + step - 1] < q[j].hi) pos += step;
                while (s_hi[w][pos] == q[j].hi && s_lo[w][pos] < q[j].lo) pos++;  // s_hi[top - 1] = +inf ends the walk
                if (valid[j] && s_hi[w][pos] == q[j].hi && s_lo[w][pos] == q[j].lo)
                    atomicAdd(&s_acc[w][pos], (uint32_t)counts[r0 + (size_t)j * 32 + lane]);
            }
            __syncwarp();
            if (S == 1) {
                // one lane per aligned 32-bit word of the count column (4 consecutive keys): no collisions inside the warp
                const int64_t w_first = t0 >> 2, w_last = (t1 - 1) >> 2;
                for (int64_t wd = w_first + lane; wd <= w_last; wd += 32) {
                    uint32_t add[4];
                    bool any = false;
#pragma unroll
                    for (int bt = 0; bt < 4; bt++) {
                        const int64_t i = wd * 4 + bt - t0;
                        add[bt] = (i >= 0 && i < nk) ? min(s_acc[w][i], 255u) : 0u;
                        if (add[bt]) { any = true; has_record[t0 + i] = 1; }
                    }
                    if (!any) continue;
                    uint32_t *word = reinterpret_cast<uint32_t *>(table_counts) + wd;  // cudaMalloc'd column: 4-byte aligned
                    uint32_t old = *word, assumed;
                    do {
                        assumed = old;
                        uint32_t nw = 0;
#pragma unroll
                        for (int bt = 0; bt < 4; bt++) {
                            const uint32_t cur = (assumed >> (8 * bt)) & 0xFFu;
                            nw |= ((255u - cur) <= add[bt] ? 255u : cur + add[bt]) << (8 * bt);
                        }
                        old = atomicCAS(word, assumed, nw);
                    } while (old != assumed);
                }
            } else {
                for (int64_t i0 = 0; i0 < nk; i0 += 32) {  // one coalesced pass over the tile's keys
                    const int64_t i = i0 + lane;
                    const uint32_t add = i < nk ? s_acc[w][i] : 0;
                    if (add) has_record[t0 + i] = 1;
                    uint8_t *p = table_counts + (size_t)(t0 + (i < nk ? i : 0)) * S + sample;
                    if (S >= 4) { if (add) sat_add_u8(p, add > 255u ? 255u : add); }   // distinct keys, distinct words
                    else sat_add_u8_warp(add != 0, p, add > 255u ? 255u : add);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < kTileLanes; j++) {
                if (!valid[j]) continue;
                const int64_t idx = table_find(kw0, kw1, n_keys, q[j].lo, q[j].hi, ix);
                if (idx >= 0) {
                    sat_add_u8(table_counts + (size_t)idx * S + sample, counts[r0 + (size_t)j * 32 + lane]);
                    has_record[idx] = 1;
                }
            }
        }
        __syncwarp();
    }
}

// ---- second form of the same merge-join (the default) ------------------------------------------------------------
// What the profile of the form above showed (profiles/r2_stream_bench_ncu_full.txt: ~1000 warp instructions per tile,
// issue slots 60 % busy, long-scoreboard stalls on four dependent loads per tile) and what this form changes:
//  * the search runs on a 32-bit delta of key_hi against the tile's bucket base (one LDS.32 + one ISETP per step) and is
//    unrolled for the tile's key count (switch on log2), 3 instructions per step instead of 11;
//  * the records' counts are loaded together with the records, not after each match (one dependent load less per record);
//  * key_hi is not staged: with the delta exact (bucket range < 2^32) equal deltas mean equal key_hi, so a tile needs
//    4 KB of shared memory instead of 5 and more CTAs fit;
//  * tiles without keys end after the index read.
// Tiles that do not qualify (no index, >= kTileKeys keys, bucket range >= 2^32: a sparse or unsorted stream) probe the
// table record by record as before.  Results are identical (tests/test_gpu_kmer.py runs every variant).
template <int LOG>
__device__ __forceinline__ uint32_t lower_bound_delta(const uint32_t *__restrict__ sd, uint32_t q) {
    uint32_t pos = 0;
#pragma unroll
    for (int b = LOG - 1; b >= 0; --b)
        if (sd[pos + (1u << b) - 1u] < q) pos += 1u << b;
    return pos;
}

template <int LOG, int LANES>
__device__ __forceinline__ void tile_match(const uint32_t *__restrict__ sd, const int64_t *__restrict__ slo, uint32_t *sacc,
                                           const uint32_t (&qd)[LANES], const int64_t (&qlo)[LANES], const uint32_t (&cnt)[LANES]) {
#pragma unroll
    for (int j = 0; j < LANES; j++) {
        uint32_t pos = lower_bound_delta<LOG>(sd, qd[j]);
        while (sd[pos] == qd[j] && slo[pos] < qlo[j]) pos++;  // ties on key_hi; the +inf slot past the last key ends the walk
        if (cnt[j] && sd[pos] == qd[j] && slo[pos] == qlo[j]) atomicAdd(&sacc[pos], cnt[j]);
    }
}

// adds the tile's accumulated counts (sacc[0..nk)) to column `sample` of table rows [t0, t0 + nk), saturating at 255
__device__ __forceinline__ void tile_write_back(const uint32_t *sacc, int64_t t0, int64_t nk, uint32_t lane, uint32_t S, uint32_t sample,
                                                uint8_t *table_counts, uint8_t *has_record) {
    if (S == 1) {
        // one lane per aligned 32-bit word of the count column (4 consecutive keys): no collisions inside the warp
        const int64_t t1 = t0 + nk;
        const int64_t w_first = t0 >> 2, w_last = (t1 - 1) >> 2;
        for (int64_t wd = w_first + lane; wd <= w_last; wd += 32) {
            uint32_t add[4];
            bool any = false;
#pragma unroll
            for (int bt = 0; bt < 4; bt++) {
                const int64_t i = wd * 4 + bt - t0;
                add[bt] = (i >= 0 && i < nk) ? min(sacc[i], 255u) : 0u;
                if (add[bt]) { any = true; has_record[t0 + i] = 1; }
            }
            if (!any) continue;
            uint32_t *word = reinterpret_cast<uint32_t *>(table_counts) + wd;  // cudaMalloc'd column: 4-byte aligned
            uint32_t old = *word, assumed;
            do {
                assumed = old;
                uint32_t nw = 0;
#pragma unroll
                for (int bt = 0; bt < 4; bt++) {
                    const uint32_t cur = (assumed >> (8 * bt)) & 0xFFu;
                    nw |= ((255u - cur) <= add[bt] ? 255u : cur + add[bt]) << (8 * bt);
                }
                old = atomicCAS(word, assumed, nw);
            } while (old != assumed);
        }
    } else {
        for (int64_t i0 = 0; i0 < nk; i0 += 32) {  // one coalesced pass over the tile's keys
            const int64_t i = i0 + lane;
            const uint32_t add = i < nk ? sacc[i] : 0;
            if (add) has_record[t0 + i] = 1;
            uint8_t *p = table_counts + (size_t)(t0 + (i < nk ? i : 0)) * S + sample;
            if (S >= 4) { if (add) sat_add_u8(p, add > 255u ? 255u : add); }   // distinct keys, distinct words
            else sat_add_u8_warp(add != 0, p, add > 255u ? 255u : add);
        }
    }
}

template <int LANES, int MIN_CTAS>
__global__ void __launch_bounds__(kStreamWarps * 32, MIN_CTAS) k_table_add_sample_v2(const int64_t *__restrict__ kw0, const int64_t *__restrict__ kw1, int64_t n_keys,
                                                          const longlong2 *__restrict__ kmers, const uint8_t *__restrict__ counts, size_t n,
                                                          uint32_t S, uint32_t sample, uint8_t *table_counts, uint8_t *has_record, TableIndex ix) {
    __shared__ uint32_t s_d[kStreamWarps][kTileKeys];
    __shared__ int64_t s_lo[kStreamWarps][kTileKeys];
    __shared__ uint32_t s_acc[kStreamWarps][kTileKeys];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const size_t n_tiles = (n + 32 * LANES - 1) / (32 * LANES);
    const size_t warp_id = (size_t)blockIdx.x * kStreamWarps + w, n_warps = (size_t)gridDim.x * kStreamWarps;
    uint32_t *sd = s_d[w], *sacc = s_acc[w];
    int64_t *slo = s_lo[w];
    for (size_t tile = warp_id; tile < n_tiles; tile += n_warps) {  // warp-uniform trip count
        const size_t r0 = tile * (32 * LANES);
        int64_t qhi[LANES], qlo[LANES];
        uint32_t cnt[LANES];
        uint32_t bmin = 0xFFFFFFFFu, bmax = 0;
        {
            // all loads of the tile first, branch-free (the last tile re-reads record n - 1 with count 0): one DRAM round trip per
            // tile instead of one per record row (with `if (i < n)` around each load the compiler kept load -> use -> load order)
            longlong2 k[LANES];
            const bool full = r0 + 32 * LANES <= n;
#pragma unroll
            for (int j = 0; j < LANES; j++) {
                const size_t i = r0 + (size_t)j * 32 + lane;
                const size_t ii = full || i < n ? i : n - 1;
                k[j] = __ldg(kmers + ii);
                cnt[j] = __ldg(counts + ii);
                if (!(full || i < n)) cnt[j] = 0;
            }
#pragma unroll
            for (int j = 0; j < LANES; j++) {
                const TableKey q = key_of_boundary(k[j].x, k[j].y);
                qhi[j] = q.hi; qlo[j] = q.lo;
                const uint32_t b = (uint32_t)(q.hi >> ix.shift);
                bmin = min(bmin, b); bmax = max(bmax, b);       // record n - 1 again in the last tile: inside the tile's range anyway
            }
        }
        bool tiled = false;
        int64_t t0 = 0, nk = 0;
        if (ix.lut) {
            bmin = __reduce_min_sync(0xFFFFFFFFu, bmin);
            bmax = __reduce_max_sync(0xFFFFFFFFu, bmax);
            int64_t v = 0;
            if (lane == 0) v = __ldg(ix.lut + bmin); else if (lane == 1) v = __ldg(ix.lut + bmax + 1);
            t0 = __shfl_sync(0xFFFFFFFFu, v, 0);
            nk = __shfl_sync(0xFFFFFFFFu, v, 1) - t0;
            if (nk == 0) continue;                          // no table key in the tile's buckets (warp-uniform)
            tiled = nk < kTileKeys && ((uint64_t)(bmax - bmin + 1u) << ix.shift) < 0xFFFFFFFFull;
        }
        if (tiled) {
            const int64_t base = (int64_t)bmin << ix.shift;
            const int log_top = 32 - __clz((uint32_t)nk);   // 2^log_top > nk: the search covers 2^log_top - 1 >= nk slots
            const uint32_t top = 1u << log_top;
            for (uint32_t i = lane; i < top; i += 64) {     // slots past the last key hold +inf: no bound checks in the search
                const uint32_t i2 = i + 32;                 // two rows per round, their four loads in flight together
                const bool in = (int64_t)i < nk, in2 = (int64_t)i2 < nk;
                const int64_t h1 = in ? __ldg(kw1 + t0 + i) : 0, l1 = in ? __ldg(kw0 + t0 + i) : INT64_MAX;
                const int64_t h2 = in2 ? __ldg(kw1 + t0 + i2) : 0, l2 = in2 ? __ldg(kw0 + t0 + i2) : INT64_MAX;
                sd[i] = in ? (uint32_t)(h1 - base) : 0xFFFFFFFFu;
                slo[i] = l1;
                sacc[i] = 0;
                if (i2 < top) {
                    sd[i2] = in2 ? (uint32_t)(h2 - base) : 0xFFFFFFFFu;
                    slo[i2] = l2;
                    sacc[i2] = 0;
                }
            }
            uint32_t qd[LANES];
#pragma unroll
            for (int j = 0; j < LANES; j++) qd[j] = cnt[j] ? (uint32_t)(qhi[j] - base) : 0xFFFFFFFFu;   // in-range deltas are < 2^32 - 1
            __syncwarp();
            switch (log_top) {
                case 1: tile_match<1, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
                case 2: tile_match<2, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
                case 3: tile_match<3, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
                case 4: tile_match<4, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
                case 5: tile_match<5, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
                case 6: tile_match<6, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
                case 7: tile_match<7, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
                default: tile_match<8, LANES>(sd, slo, sacc, qd, qlo, cnt); break;
            }
            __syncwarp();
            tile_write_back(sacc, t0, nk, lane, S, sample, table_counts, has_record);
        } else {
#pragma unroll
            for (int j = 0; j < LANES; j++) {
                if (r0 + (size_t)j * 32 + lane >= n) continue;
                const int64_t idx = table_find(kw0, kw1, n_keys, qlo[j], qhi[j], ix);
                if (idx >= 0 && cnt[j]) {
                    sat_add_u8(table_counts + (size_t)idx * S + sample, cnt[j]);
                    has_record[idx] = 1;
                }
            }
        }
        __syncwarp();
    }
}

// table keys <-> packed k-mers in the ABI's boundary layout
__global__ void __launch_bounds__(256) k_keys_from_kmers(const longlong2 *__restrict__ kmers, size_t n, int64_t *__restrict__ kw0, int64_t *__restrict__ kw1) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const longlong2 k = __ldg(kmers + i);
        const TableKey q = key_of_boundary(k.x, k.y);
        kw0[i] = q.lo; kw1[i] = q.hi;
    }
}
__global__ void __launch_bounds__(256) k_keys_to_kmers(const int64_t *__restrict__ kw0, const int64_t *__restrict__ kw1, size_t n, longlong2 *__restrict__ kmers) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t w0, w1;
        to_boundary(Kmer128{(uint64_t)kw1[i], (uint64_t)kw0[i] ^ kLoBias}, w0, w1);
        kmers[i] = longlong2{(long long)w0, (long long)w1};
    }
}

// KmerCounter::countInterclusterKmersCallback: rolling scan of one region, probe, addInterclusterMultiplicity
constexpr int kScanChunk = 64;
__global__ void __launch_bounds__(256) k_table_scan_region(const int64_t *__restrict__ kw0, const int64_t *__restrict__ kw1, int64_t n_keys,
                                                           const char *__restrict__ seq, size_t len, uint32_t is_decoy, uint32_t ploidy_f, uint32_t ploidy_m,
                                                           uint8_t *ic, uint8_t *max_mult, uint8_t *decoy, uint8_t *has_record, TableIndex ix) {
    const size_t nchunks = (len + kScanChunk - 1) / kScanChunk;
    for (size_t ch = blockIdx.x * (size_t)blockDim.x + threadIdx.x; ch < nchunks; ch += (size_t)gridDim.x * blockDim.x) {
        const size_t p0 = ch * kScanChunk, p1 = p0 + kScanChunk < len ? p0 + kScanChunk : len;
        Roller roll;
        roll.reset();
        const size_t start = p0 >= (size_t)(K - 1) ? p0 - (K - 1) : 0;
        for (size_t p = start; p < p1; p++) {
            const unsigned c = nt_code(__ldg(seq + p));
            bool complete = false;
            if (c > 3) roll.reset(); else complete = roll.push(c);
            if (complete && p >= p0) {
                const TableKey q = key_of(roll.canonical());
                const int64_t idx = table_find(kw0, kw1, n_keys, q.lo, q.hi, ix);
                if (idx >= 0) {
                    has_record[idx] = 1;
                    sat_add_u8(max_mult + idx, 1);  // max_haploid_multiplicity (KmerCounts.cpp:100)
                    if (is_decoy) decoy[idx] = 1;
                    else { sat_add_u8(ic + idx * 2, ploidy_f); sat_add_u8(ic + idx * 2 + 1, ploidy_m); }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_table_lookup(const int64_t *__restrict__ kw0, const int64_t *__restrict__ kw1, int64_t n_keys,
                                                      const longlong2 *__restrict__ kmers, size_t n, int64_t *__restrict__ idx_out, TableIndex ix) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const longlong2 k = __ldg(kmers + i);
        const TableKey q = key_of_boundary(k.x, k.y);
        idx_out[i] = table_find(kw0, kw1, n_keys, q.lo, q.hi, ix);
    }
}

}  // namespace

extern "C" {

// `g` fields are device pointers (the caller keeps graphs + best paths resident); see btgpu.h
int btg_walk_paths_dev(const btg_pathwalk_desc *d, int emit, uint32_t *n_occ, uint32_t *n_cov, const uint64_t *occ_off, const uint64_t *cov_off,
                       int64_t *key_w0, int64_t *key_w1, uint32_t *occ_path, uint32_t *occ_nt, int64_t *cov_occ, uint16_t *cov_var,
                       uint32_t *status, void *stream) {
    BTG_REQUIRE_INIT();
    if (!d) { set_error("null argument"); return BTG_EINVAL; }
    if (d->n_paths == 0) return BTG_OK;
    PathWalkGraphs g{d->n_clusters, d->cl_vertex_off, d->v_seq_off, d->seq, d->v_flags, d->v_var, d->v_allele, d->v_refvar_off, d->v_refvar,
                     d->cl_path_off, d->path_mem_off, d->path_mem, d->path_cluster};
    const unsigned grid = (unsigned)((d->n_paths + 127) / 128);
    if (emit) k_walk_paths<true><<<grid, 128, 0, pick_stream(stream)>>>(g, d->n_paths, n_occ, n_cov, occ_off, cov_off, key_w0, key_w1, occ_path, occ_nt, cov_occ, cov_var, status);
    else k_walk_paths<false><<<grid, 128, 0, pick_stream(stream)>>>(g, d->n_paths, n_occ, n_cov, occ_off, cov_off, key_w0, key_w1, occ_path, occ_nt, cov_occ, cov_var, status);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_path_alleles_dev(const btg_pathwalk_desc *d, const uint64_t *cl_var_off, const uint16_t *var_nalleles, const uint64_t *hapvar_off,
                         uint16_t *hap_alleles, void *stream) {
    BTG_REQUIRE_INIT();
    if (!d) { set_error("null argument"); return BTG_EINVAL; }
    if (d->n_paths == 0) return BTG_OK;
    PathWalkGraphs g{d->n_clusters, d->cl_vertex_off, d->v_seq_off, d->seq, d->v_flags, d->v_var, d->v_allele, d->v_refvar_off, d->v_refvar,
                     d->cl_path_off, d->path_mem_off, d->path_mem, d->path_cluster};
    k_path_alleles<<<(unsigned)((d->n_paths + 127) / 128), 128, 0, pick_stream(stream)>>>(g, d->n_paths, cl_var_off, var_nalleles, hapvar_off, hap_alleles);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

// Per host thread: two threads (or two KmerPipelines driven from different threads) never see each other's index.  A probe without an
// index is correct (binary search over all keys), only slower; btg_counter installs its own index around each of its passes.
static thread_local TableIndex g_index{nullptr, 0};

// Installs (or clears, lut = NULL) the prefix index used by the table probes that follow: lut[b] = index of the first
// key whose word-1 top `lut_bits` bits (of 46) are >= b, for b in [0, 2^lut_bits]; built by the caller from the keys.
int btg_table_set_index_dev(const int64_t *lut, int lut_bits) {
    if (lut && (lut_bits < 1 || lut_bits > 30)) { set_error("lut_bits must be 1..30"); return BTG_EINVAL; }
    g_index = TableIndex{lut, lut ? 2 * K - 64 - lut_bits : 0};
    return BTG_OK;
}

int btg_table_keys_from_kmers_dev(const uint64_t *kmers, size_t n, int64_t *key_lo, int64_t *key_hi, void *stream) {
    BTG_REQUIRE_INIT();
    if (n == 0) return BTG_OK;
    k_keys_from_kmers<<<btg_grid_for(n, 256, 8), 256, 0, pick_stream(stream)>>>((const longlong2 *)kmers, n, key_lo, key_hi);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_table_keys_to_kmers_dev(const int64_t *key_lo, const int64_t *key_hi, size_t n, uint64_t *kmers, void *stream) {
    BTG_REQUIRE_INIT();
    if (n == 0) return BTG_OK;
    k_keys_to_kmers<<<btg_grid_for(n, 256, 8), 256, 0, pick_stream(stream)>>>(key_lo, key_hi, n, (longlong2 *)kmers);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_table_lookup_dev(const int64_t *key_w0, const int64_t *key_w1, int64_t n_keys, const uint64_t *kmers, size_t n, int64_t *idx_out, void *stream) {
    BTG_REQUIRE_INIT();
    if (n == 0) return BTG_OK;
    k_table_lookup<<<btg_grid_for(n, 256, 8), 256, 0, pick_stream(stream)>>>(key_w0, key_w1, n_keys, (const longlong2 *)kmers, n, idx_out, g_index);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_table_add_sample_kmers_dev(const int64_t *key_w0, const int64_t *key_w1, int64_t n_keys, const uint64_t *kmers, const uint8_t *counts, size_t n,
                                   uint32_t n_samples, uint32_t sample_idx, uint8_t *table_counts, uint8_t *has_record, void *stream) {
    BTG_REQUIRE_INIT();
    if (sample_idx >= n_samples) { set_error("sample index out of range"); return BTG_EINVAL; }
    if (reinterpret_cast<uintptr_t>(table_counts) & 3u) { set_error("table_counts must be 4-byte aligned"); return BTG_EINVAL; }
    if (n == 0) return BTG_OK;
    static const int variant = [] { const char *e = getenv("BTG_STREAM_VARIANT"); return e ? atoi(e) : 1; }();
    const cudaStream_t st = pick_stream(stream);
#define BTG_STREAM_LAUNCH(KERNEL, LANES, CTAS)                                                                                              \
    KERNEL<<<btg_grid_for((n + (LANES) - 1) / (LANES), kStreamWarps * 32, CTAS), kStreamWarps * 32, 0, st>>>(                                   \
        key_w0, key_w1, n_keys, (const longlong2 *)kmers, counts, n, n_samples, sample_idx, table_counts, has_record, g_index)
    switch (variant) {
        case 0: BTG_STREAM_LAUNCH(k_table_add_sample, kTileLanes, 5); break;
        case 2: BTG_STREAM_LAUNCH((k_table_add_sample_v2<4, 6>), 4, 6); break;     // 40 registers, 6 CTAs per SM: measured equal to the default
        default: BTG_STREAM_LAUNCH((k_table_add_sample_v2<4, 5>), 4, 5); break;
    }
#undef BTG_STREAM_LAUNCH
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_table_scan_region_dev(const int64_t *key_w0, const int64_t *key_w1, int64_t n_keys, const char *seq, size_t len, int is_decoy,
                              uint32_t ploidy_female, uint32_t ploidy_male, uint8_t *ic, uint8_t *max_mult, uint8_t *decoy, uint8_t *has_record, void *stream) {
    BTG_REQUIRE_INIT();
    if (len == 0) return BTG_OK;
    const size_t nchunks = (len + kScanChunk - 1) / kScanChunk;
    k_table_scan_region<<<btg_grid_for(nchunks, 256, 4), 256, 0, pick_stream(stream)>>>(key_w0, key_w1, n_keys, seq, len, is_decoy ? 1u : 0u, ploidy_female,
                                                                                       ploidy_male, ic, max_mult, decoy, has_record, g_index);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

}  // extern "C"
