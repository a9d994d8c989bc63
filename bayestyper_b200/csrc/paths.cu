// paths.cu — per-(cluster, sample) candidate-path search on the device.
//
// Replaces VariantClusterGraph::findSamplePaths + mergePaths / isPathsRedundant / filterPaths / addPathIndices
// (src/bayesTyper/VariantClusterGraph.cpp:389-798), VariantClusterGraphPath::{addVertex,updateScore,getKmerScore,
// getVertexScore,updateObservedCoveredVertices} (src/bayesTyper/VariantClusterGraphPath.cpp:46-225) and the driver
// KmerCounter::findVariantClusterPaths (src/bayesTyper/KmerCounter.cpp:59-103).
//
// The north star asks for BIT-EXACT path enumeration, and the reference's result depends on the order in which
// std::shuffle (libstdc++ 13: pairwise swaps drawn with Lemire's method from std::mt19937) leaves the candidate
// paths before the greedy filter.  The kernel therefore carries a real mt19937 and the exact libstdc++ shuffle.
//
// Mapping: clusters are independent; ONE WARP walks one cluster's vertex DP for the current sample.  The DP itself (merge,
// shuffle, greedy filter) is a sequential programme over small lists: lane 0 runs it on a working set that lives in SHARED
// memory (path slots, per-vertex lists; clusters whose working set exceeds the per-warp budget fall back to a global scratch
// arena).  The part that costs memory latency — one Bloom lookup per completed k-mer of every candidate path that is extended
// by a vertex — is spread over the lanes: lane j rebuilds the window that ends at the j-th nucleotide of the vertex from the
// path's 110-bit tail register, hashes its canonical form from the 4-nucleotide ntHash table and issues all its probes, so 32
// lookups (up to 320 probes) are in flight per warp; the hit mask comes back through a ballot and lane 0 books the scores in
// sequence order (VariantClusterGraphPath::updateScore).  Round 1 ran one THREAD per cluster on a global scratch arena:
// 7.9 of 32 lanes active, every list access an L2 round trip (profiles/r1_find_sample_paths_ncu_full.txt).  The kernel is
// persistent: warps draw clusters (largest first) from an atomic counter, so a few huge clusters do not leave a tail.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kmer.cuh"

using namespace btg;

struct btg_bloom;
namespace btg_internal {
BloomView bloom_view(const btg_bloom *b);  // bloom.cu
}

namespace {

constexpr uint8_t kMinObserved = 2;   // min_observed_kmers (VariantClusterGraphPath.cpp:36)
constexpr uint32_t kMinSamplePaths = 1;  // min_num_sample_paths (VariantClusterGraph.cpp:60)
constexpr uint8_t kAbsent = 0xFF;

struct DevGraphs {
    uint32_t C;
    const uint64_t *cl_vertex_off;  // [C+1]
    const uint64_t *v_seq_off;      // [V+1]
    const uint8_t *seq;             // nucleotide codes 0..3
    const uint8_t *v_flags;         // bit0 is_first_nucleotides_redundant, bit1 is_disconnected
    const uint64_t *v_in_off;       // [V+1]
    const uint32_t *v_in_src;       // local ids, in_edges order
    const uint32_t *v_max_target;   // [V] max(v, targets of v)   (visited_vertices, VariantClusterGraph.cpp:462-476)
    const uint32_t *cl_group;       // group index of the cluster (seed)
    const uint32_t *cl_idx;         // variant_cluster_idx within the group (seed)
    // scratch layout per cluster
    const uint64_t *scr_off;        // [C+1] byte offsets into scratch
    const uint32_t *cl_pool;        // [C] number of path slots
    const uint32_t *cl_tmp;         // [C] capacity of the merge list
    uint8_t *scratch;
    // results: best paths per cluster, merged over samples (best_paths_indices)
    const uint64_t *best_off;       // [C+1] byte offsets into best (capacity rows x V bytes)
    const uint32_t *best_cap;       // [C] capacity in paths
    uint32_t *best_n;               // [C] current number of best paths
    uint8_t *best;                  // path x vertex membership bytes
    uint32_t *status;               // [0] = 1 + first cluster whose scratch overflowed (0 = none)
    const uint32_t *order;          // [C] clusters sorted by (vertices, sequence length), largest first: the draw order of the warps
    const uint32_t *cl_smem;        // [C] bytes of the cluster's working set when it fits the per-warp shared-memory budget, else 0
    uint32_t *mt_pool;              // [resident warps][624] Mersenne states (materialised only when a cluster draws > kMtWindow numbers)
    uint32_t *next;                 // work counter
    unsigned long long *stats;      // [3] k-mer lookups, Bloom probes executed under the reference's early-exit order, nucleotides walked (roofline)
    // several samples in ONE launch (btg_find_sample_paths_batch): work item = (cluster, sample); 0 = one sample per launch
    uint32_t batch_n;               // samples in this launch
    uint32_t batch_first;           // sample index of the first one
    const BloomView *batch_blooms;  // [batch_n]
    uint32_t *turn;                 // [C] samples of the cluster whose paths have been merged into `best` so far (addPathIndices is order-dependent)
    uint8_t *warp_scratch;          // [resident warps][warp_scratch_bytes]: working sets that exceed the shared-memory budget (one per warp, not per cluster)
    uint64_t warp_scratch_bytes;
};

// ---- std::mt19937 -------------------------------------------------------------------------------
// Bit-exact Mersenne Twister.  A cluster usually draws a handful of numbers, so the first kMtWindow outputs are
// produced without materialising the 624-word state: output i < 227 only needs the seeded words mt[i], mt[i+1] and
// mt[i+397], which come out of the (register-resident) seeding recurrence.  Only when a cluster draws more than that
// is the full state built in the scratch arena and twisted the standard way.
constexpr uint32_t kMtWindow = 24;
struct Mt19937 {
    uint32_t *mt;      // [624] scratch, used only after the window is exhausted
    uint32_t seed_value;
    uint32_t consumed; // outputs handed out so far
    uint32_t idx;      // position in the materialised state
    bool windowed, full;
    uint32_t wa[kMtWindow + 1], wb[kMtWindow];

    __device__ void seed(uint32_t s) { seed_value = s; consumed = 0; idx = 624; windowed = false; full = false; }
    __device__ void build_window() {
        uint32_t x = seed_value;
        wa[0] = x;
        for (uint32_t i = 1; i < 397 + kMtWindow; i++) {
            x = 1812433253u * (x ^ (x >> 30)) + i;
            if (i <= kMtWindow) wa[i] = x;
            if (i >= 397) wb[i - 397] = x;
        }
        windowed = true;
    }
    __device__ void materialise() {
        uint32_t x = seed_value;
        mt[0] = x;
        for (uint32_t i = 1; i < 624; i++) { x = 1812433253u * (x ^ (x >> 30)) + i; mt[i] = x; }
        for (uint32_t i = 0; i < 624; i++) {
            const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx = consumed;  // the first `consumed` (< 624) outputs of this block were served from the window
        full = true;
    }
    __device__ void twist() {
        for (uint32_t i = 0; i < 624; i++) {
            const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx = 0;
    }
    __device__ static uint32_t temper(uint32_t y) {
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    __device__ uint32_t next() {
        if (!full) {
            if (consumed < kMtWindow) {
                if (!windowed) build_window();
                const uint32_t i = consumed++;
                const uint32_t y = (wa[i] & 0x80000000u) | (wa[i + 1] & 0x7fffffffu);
                return temper(wb[i] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u));
            }
            materialise();
        }
        if (idx >= 624) twist();
        consumed++;
        return temper(mt[idx++]);
    }
    // libstdc++ uniform_int_distribution<unsigned long>{0, range-1} on a 32-bit engine: Lemire (uniform_int_dist.h:250-274)
    __device__ uint32_t below(uint32_t range) {
        uint64_t product = (uint64_t)next() * range;
        uint32_t low = (uint32_t)product;
        if (low < range) {
            const uint32_t threshold = (0u - range) % range;
            while (low < threshold) { product = (uint64_t)next() * range; low = (uint32_t)product; }
        }
        return (uint32_t)(product >> 32);
    }
};

// libstdc++ 13 std::shuffle (stl_algo.h): two swap positions per engine call
__device__ void std_shuffle(uint16_t *a, uint32_t n, Mt19937 &g) {
    if (n == 0) return;
    uint32_t i = 1;
    if ((n % 2) == 0) {
        const uint32_t j = g.below(2);
        const uint16_t t = a[i]; a[i] = a[j]; a[j] = t;
        i++;
    }
    while (i != n) {
        const uint32_t swap_range = i + 1;
        const uint32_t x = g.below(swap_range * (swap_range + 1));  // __gen_two_uniform_ints(b0, b1): {x / b1, x % b1}
        const uint32_t p0 = x / (swap_range + 1), p1 = x % (swap_range + 1);
        uint16_t t = a[i]; a[i] = a[p0]; a[p0] = t; i++;
        t = a[i]; a[i] = a[p1]; a[p1] = t; i++;
    }
}

// ---- one cluster's working set -------------------------------------------------------------------
struct PathHdr {        // 32 bytes, followed by V bytes of per-vertex state (kAbsent or num_observed_kmers 0..2)
    uint64_t fhi, flo;  // the last <= 55 nucleotides of the path (KmerPair's forward register, kmer.cuh internal form)
    uint32_t filled, score_first, score_second, nverts;
};

struct Work {
    const DevGraphs *g;
    uint32_t c, V;
    uint64_t v0;
    uint32_t slot_bytes, pool, tmp_cap;
    uint8_t *slots;       // pool x slot_bytes
    uint16_t *lists;      // V x 32 slot ids
    uint8_t *list_n;      // V
    uint16_t *tmp;        // merge list
    uint16_t *free_stack; // pool
    uint32_t n_free;
    uint8_t *covered;     // V (filterPaths' observed_covered_vertices)
    uint32_t *mt;         // 624
    bool overflow;

    __device__ PathHdr *hdr(uint32_t s) const { return reinterpret_cast<PathHdr *>(slots + (size_t)s * slot_bytes); }
    __device__ uint8_t *verts(uint32_t s) const { return slots + (size_t)s * slot_bytes + sizeof(PathHdr); }
    __device__ uint32_t alloc() {
        if (n_free == 0) { overflow = true; return 0; }
        return free_stack[--n_free];
    }
    __device__ void release(uint32_t s) { free_stack[n_free++] = (uint16_t)s; }
    __device__ uint32_t seq_len(uint32_t v) const { return (uint32_t)(g->v_seq_off[v0 + v + 1] - g->v_seq_off[v0 + v]); }
    __device__ const uint8_t *seq(uint32_t v) const { return g->seq + g->v_seq_off[v0 + v]; }
    __device__ bool disconnected(uint32_t v) const { return g->v_flags[v0 + v] & 2; }
    __device__ bool first_redundant(uint32_t v) const { return g->v_flags[v0 + v] & 1; }
};

__device__ void copy_slot(Work &w, uint32_t dst, uint32_t src) {
    const uint64_t *s = reinterpret_cast<const uint64_t *>(w.slots + (size_t)src * w.slot_bytes);
    uint64_t *d = reinterpret_cast<uint64_t *>(w.slots + (size_t)dst * w.slot_bytes);
    for (uint32_t i = 0; i < w.slot_bytes / 8; i++) d[i] = s[i];
}

// previous vertex of a path strictly below `v` (paths hold ascending vertex indices); V = none
__device__ __forceinline__ uint32_t prev_vertex(const uint8_t *pv, uint32_t v) {
    while (v > 0) { v--; if (pv[v] != kAbsent) return v; }
    return 0xFFFFFFFFu;
}
__device__ __forceinline__ uint32_t last_vertex(const uint8_t *pv, uint32_t V) { return prev_vertex(pv, V); }

// VariantClusterGraph::isPathsRedundant (VariantClusterGraph.cpp:534-629): both paths spell the same sequence with
// disconnections at the same places, compared from the back
__device__ bool paths_redundant(const Work &w, const uint8_t *p1, const uint8_t *p2) {
    uint32_t v1 = last_vertex(p1, w.V), v2 = last_vertex(p2, w.V);
    int64_t i1 = (int64_t)w.seq_len(v1) - 1, i2 = (int64_t)w.seq_len(v2) - 1;  // index of the next nucleotide to compare, -1 = vertex exhausted
    bool d1 = false, d2 = false;
    const uint32_t END = 0xFFFFFFFFu;
    while (true) {
        while (i1 < 0) {
            if (w.disconnected(v1)) d1 = true;
            v1 = prev_vertex(p1, v1);
            if (v1 != END) i1 = (int64_t)w.seq_len(v1) - 1; else break;
        }
        while (i2 < 0) {
            if (w.disconnected(v2)) d2 = true;
            v2 = prev_vertex(p2, v2);
            if (v2 != END) i2 = (int64_t)w.seq_len(v2) - 1; else break;
        }
        if (d1 != d2) return false;
        d1 = d2 = false;
        if (v1 == END || v2 == END) break;
        if (v1 == v2 && i1 == i2) { i1 = -1; i2 = -1; }  // same vertex, same offset: identical from here to its start
        const uint8_t *s1 = w.seq(v1), *s2 = w.seq(v2);
        while (i1 >= 0 && i2 >= 0) {
            if (s1[i1] != s2[i2]) return false;
            i1--; i2--;
        }
    }
    return v1 == END && v2 == END;
}

// VariantClusterGraph::mergePaths (VariantClusterGraph.cpp:484-532).  Paths are always copied into fresh slots;
// a source list is released after its last consumer (the reference moves instead of copying there: same result).
__device__ void merge_paths(Work &w, uint32_t &n_main, const uint16_t *input, uint32_t n_in) {
    const uint32_t main_size = n_main;
    for (uint32_t i = 0; i < n_in && !w.overflow; i++) {
        const uint32_t in_slot = input[i];
        const uint8_t *pin = w.verts(in_slot);
        bool redundant = false;
        for (uint32_t m = 0; m < main_size; m++) {
            const uint32_t ms = w.tmp[m];
            if (paths_redundant(w, w.verts(ms), pin)) {
                if (w.hdr(ms)->nverts < w.hdr(in_slot)->nverts) copy_slot(w, ms, in_slot);  // keep the longer vertex list
                redundant = true;
                break;
            }
        }
        if (!redundant) {
            if (n_main >= w.tmp_cap) { w.overflow = true; return; }
            const uint32_t s = w.alloc();
            if (w.overflow) return;
            copy_slot(w, s, in_slot);
            w.tmp[n_main++] = (uint16_t)s;
        }
    }
}

// VariantClusterGraphPath::updateScore (VariantClusterGraphPath.cpp:89-130)
__device__ void update_score(Work &w, PathHdr *h, uint8_t *pv, uint32_t cur_vertex, bool observed, uint32_t cur_len) {
    if (observed) {
        h->score_first++;
        uint32_t v = cur_vertex;
        if (cur_len > 1 || !w.first_redundant(v)) if (pv[v] < kMinObserved) pv[v]++;
        v = prev_vertex(pv, v);
        while (v != 0xFFFFFFFFu) {
            if ((uint32_t)K <= cur_len || pv[v] == kMinObserved) break;
            if (pv[v] < kMinObserved) pv[v]++;
            cur_len += w.seq_len(v);
            v = prev_vertex(pv, v);
        }
    }
    h->score_second++;
}

// VariantClusterGraphPath::addVertex (VariantClusterGraphPath.cpp:46-87) by the whole warp: lane j takes the k-mer window that ends
// at nucleotide base + j of the vertex; lane 0 books the scores in order.  T = 4-nucleotide ntHash table (shared).
__device__ void add_vertex_warp(Work &w, uint32_t slot, uint32_t v, const BloomView &bloom, const uint64_t *T, uint32_t lane) {
    PathHdr *h = w.hdr(slot);
    uint8_t *pv = w.verts(slot);
    const uint32_t len = w.seq_len(v);
    const bool disc = w.disconnected(v);
    if (lane == 0) {
        pv[v] = 0;
        h->nverts++;
        if (disc) {
            if (len == 0) pv[v] = kMinObserved;
            h->fhi = h->flo = 0; h->filled = 0;  // KmerPair::reset
        }
    }
    __syncwarp();
    Kmer128 f0{h->fhi, h->flo};
    uint32_t filled0 = h->filled;
    const uint8_t *s = w.seq(v);
    for (uint32_t base = 0; base < len; base += 32) {
        const uint32_t n = len - base < 32 ? len - base : 32;   // nucleotides of this round
        const uint32_t c_mine = lane < n ? s[base + lane] : 0u;
        Kmer128 f = f0;
        for (uint32_t t = 0; t < n; t++) {                      // shift in nucleotides base .. base + lane
            const uint32_t ct = __shfl_sync(0xFFFFFFFFu, c_mine, t);
            if (t <= lane) {
                f.hi = ((f.hi << 2) | (f.lo >> 62)) & kHiMask;
                f.lo = (f.lo << 2) | ct;
            }
        }
        const bool complete = lane < n && filled0 + lane + 1 >= (uint32_t)K;
        bool hit = false;
        unsigned probes = 0;
        if (complete) {
            const Kmer128 r = revcomp(f);
            hit = bloom_contains(bloom, ntp64(forward_is_canonical(f, r) ? f : r, T), &probes);
        }
        const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, complete), hmask = __ballot_sync(0xFFFFFFFFu, hit);
        const uint32_t np = __reduce_add_sync(0xFFFFFFFFu, probes);
        if (lane == 0) { atomicAdd(w.g->stats, (unsigned long long)__popc(cmask)); atomicAdd(w.g->stats + 1, (unsigned long long)np); atomicAdd(w.g->stats + 2, (unsigned long long)n); }
        f0.hi = __shfl_sync(0xFFFFFFFFu, f.hi, n - 1);
        f0.lo = __shfl_sync(0xFFFFFFFFu, f.lo, n - 1);
        filled0 = filled0 + n < (uint32_t)K ? filled0 + n : (uint32_t)K;
        if (lane == 0)
            for (uint32_t t = 0; t < n; t++)
                if ((cmask >> t) & 1u) update_score(w, h, pv, v, (hmask >> t) & 1u, base + t + 1);
        __syncwarp();
    }
    __syncwarp();   // every lane has read the header (a vertex without nucleotides runs no round above)
    if (lane == 0) { h->fhi = f0.hi; h->flo = f0.lo; h->filled = filled0; }
    __syncwarp();
}

__device__ __forceinline__ double kmer_score(const PathHdr *h) {  // getKmerScore (…GraphPath.cpp:137-149)
    return h->score_second > 0 ? h->score_first / static_cast<double>(h->score_second) : 1.0;
}
__host__ __device__ inline bool dbl_compare(double a, double b) {  // Utils.hpp:81-87
    return (a == b) || (fabs(a - b) < fabs(a < b ? a : b) * 2.220446049250313e-16 * 100);
}

// getVertexScore / updateObservedCoveredVertices (…GraphPath.cpp:151-225)
__device__ uint32_t vertex_score(const Work &w, const uint8_t *pv, bool is_complete) {
    uint32_t score = 0;
    for (uint32_t v = 0; v < w.V; v++) if (pv[v] == kMinObserved && !w.covered[v]) score++;
    if (!is_complete) {
        uint32_t v = last_vertex(pv, w.V), cur_len = 0;
        while (v != 0xFFFFFFFFu) {
            if ((uint32_t)(K - 1) <= cur_len || pv[v] == kMinObserved) break;
            if (!w.covered[v]) score++;
            cur_len += w.seq_len(v);
            v = prev_vertex(pv, v);
        }
    }
    return score;
}
__device__ void update_covered(Work &w, const uint8_t *pv, bool is_complete) {
    for (uint32_t v = 0; v < w.V; v++) if (pv[v] == kMinObserved) w.covered[v] = 1;
    if (!is_complete) {
        uint32_t v = last_vertex(pv, w.V), cur_len = 0;
        while (v != 0xFFFFFFFFu) {
            if ((uint32_t)(K - 1) <= cur_len || pv[v] == kMinObserved) break;
            w.covered[v] = 1;
            cur_len += w.seq_len(v);
            v = prev_vertex(pv, v);
        }
    }
}

// VariantClusterGraph::filterPaths (VariantClusterGraph.cpp:631-724): greedy selection in place; returns new size
__device__ uint32_t filter_paths(Work &w, uint16_t *paths, uint32_t n, uint32_t max_paths, bool is_complete) {
    if (!((n > max_paths) || (is_complete && n > kMinSamplePaths))) return n;
    bool first_pass = true;
    for (uint32_t v = 0; v < w.V; v++) w.covered[v] = 0;
    uint32_t sorted_end = 0;
    while (sorted_end != n) {
        uint32_t best = sorted_end;
        double best_k = kmer_score(w.hdr(paths[best]));
        uint32_t best_v = vertex_score(w, w.verts(paths[best]), is_complete);
        for (uint32_t i = sorted_end + 1; i < n; i++) {
            const double ck = kmer_score(w.hdr(paths[i]));
            const uint32_t cv = vertex_score(w, w.verts(paths[i]), is_complete);
            if (first_pass) {
                if (cv > 0) {
                    if ((dbl_compare(ck, best_k) && cv > best_v) || ck > best_k || best_v == 0) { best = i; best_k = ck; best_v = cv; }
                }
            } else if (!is_complete || cv == w.hdr(paths[i])->nverts) {
                if (ck > best_k) { best = i; best_k = ck; best_v = cv; }
            }
        }
        if (first_pass) {
            update_covered(w, w.verts(paths[best]), is_complete);
        } else if (is_complete && sorted_end >= kMinSamplePaths && best_v < w.hdr(paths[best])->nverts) {
            break;
        }
        if (sorted_end != best) { const uint16_t t = paths[sorted_end]; paths[sorted_end] = paths[best]; paths[best] = t; }
        if (first_pass && best_v == 0) {
            first_pass = false;
            for (uint32_t v = 0; v < w.V; v++) w.covered[v] = 0;
        } else {
            sorted_end++;
            if (sorted_end == max_paths) break;
        }
    }
    for (uint32_t i = sorted_end; i < n; i++) w.release(paths[i]);
    return sorted_end;
}

// redundancy between a stored best path (row of membership bytes) and a sample path: same predicate, the row
// only needs != kAbsent semantics, which paths_redundant uses
// VariantClusterGraph::addPathIndices (VariantClusterGraph.cpp:726-798)
__device__ void add_path_indices(Work &w, const uint16_t *paths, uint32_t n) {
    const DevGraphs &g = *w.g;
    uint8_t *best = g.best + g.best_off[w.c];
    uint32_t n_best = g.best_n[w.c];
    const uint32_t cap = g.best_cap[w.c];
    uint32_t redundant_mask_lo = 0;  // n <= 32
    for (uint32_t b = 0; b < n_best; b++) {
        uint8_t *row = best + (size_t)b * w.V;
        uint32_t row_n = 0;
        for (uint32_t v = 0; v < w.V; v++) row_n += row[v] != kAbsent;
        for (uint32_t i = 0; i < n; i++) {
            if (redundant_mask_lo & (1u << i)) continue;
            const uint8_t *pv = w.verts(paths[i]);
            if (paths_redundant(w, row, pv)) {
                if (row_n < w.hdr(paths[i])->nverts)
                    for (uint32_t v = 0; v < w.V; v++) row[v] = pv[v] != kAbsent ? 0 : kAbsent;
                redundant_mask_lo |= 1u << i;
                break;
            }
        }
    }
    for (uint32_t i = 0; i < n; i++) {
        if (!(redundant_mask_lo & (1u << i))) {
            if (n_best >= cap) { w.overflow = true; break; }
            const uint8_t *pv = w.verts(paths[i]);
            uint8_t *row = best + (size_t)n_best * w.V;
            for (uint32_t v = 0; v < w.V; v++) row[v] = pv[v] != kAbsent ? 0 : kAbsent;
            n_best++;
        }
    }
    g.best_n[w.c] = n_best;
}

// VariantClusterGraph::findSamplePaths (VariantClusterGraph.cpp:389-482): persistent kernel, one warp per cluster at a time
constexpr uint32_t kPathWarps = 4;            // warps per block
constexpr uint32_t kPathSmemPerWarp = 8192;   // shared-memory budget of one cluster's working set

// BATCH: work item = (cluster, sample of the batch) — the samples of a cluster run side by side on different warps and merge their paths into
// the cluster's best paths in sample order (btg_find_sample_paths_batch); otherwise one sample per launch, exactly the round-2 kernel.
template <bool BATCH>
__global__ void __launch_bounds__(kPathWarps * 32) k_find_sample_paths(DevGraphs g, BloomView bloom, uint32_t random_seed, uint32_t sample_idx, uint32_t max_paths) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *T = reinterpret_cast<uint64_t *>(smem);                    // [256] ntHash 4-nucleotide table
    build_hash_table(T);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    uint8_t *my_smem = smem + 2048 + (size_t)wib * kPathSmemPerWarp;
    uint32_t *my_mt = g.mt_pool + ((size_t)blockIdx.x * kPathWarps + wib) * 624;
    const uint32_t per_cluster = BATCH ? g.batch_n : 1u;           // work items per cluster: the samples of the batch, handed out in sample order
    const uint64_t n_items = (uint64_t)g.C * per_cluster;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(g.next, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_items) break;
        const uint32_t c = g.order[BATCH ? t / per_cluster : t];
        const uint32_t batch_pos = BATCH ? t % per_cluster : 0u;
        if constexpr (BATCH) { sample_idx = g.batch_first + batch_pos; bloom = g.batch_blooms[batch_pos]; }
        Work w;
        w.g = &g; w.c = c;
        w.v0 = g.cl_vertex_off[c];
        w.V = (uint32_t)(g.cl_vertex_off[c + 1] - w.v0);
        w.pool = g.cl_pool[c]; w.tmp_cap = g.cl_tmp[c];
        w.slot_bytes = (uint32_t)(sizeof(PathHdr) + ((w.V + 7) & ~7u));
        uint8_t *p = g.cl_smem[c] ? my_smem
                     : BATCH ? g.warp_scratch + ((size_t)blockIdx.x * kPathWarps + wib) * g.warp_scratch_bytes   // the samples of a cluster run side by side
                                 : g.scratch + g.scr_off[c];
        w.mt = my_mt;
        w.slots = p; p += (size_t)w.pool * w.slot_bytes;
        w.lists = reinterpret_cast<uint16_t *>(p); p += (size_t)w.V * 32 * 2;
        w.tmp = reinterpret_cast<uint16_t *>(p); p += (size_t)((w.tmp_cap + 3) & ~3u) * 2;
        w.free_stack = reinterpret_cast<uint16_t *>(p); p += (size_t)((w.pool + 3) & ~3u) * 2;
        w.list_n = p; p += (w.V + 7) & ~7u;
        w.covered = p;
        w.overflow = false;
        w.n_free = w.pool;
        // seed = r + (g+1)(s+1) + cluster_idx (KmerCounter.cpp:65, VariantClusterGroup.cpp:142)
        Mt19937 rng;
        rng.mt = w.mt;
        rng.seed(random_seed + (g.cl_group[c] + 1) * (sample_idx + 1) + g.cl_idx[c]);
        if (lane == 0) {
            for (uint32_t i = 0; i < w.pool; i++) w.free_stack[i] = (uint16_t)(w.pool - 1 - i);
            for (uint32_t v = 0; v < w.V; v++) w.list_n[v] = 0;
        }
        __syncwarp();
        for (uint32_t v = 0; v < w.V; v++) {
            uint32_t n = 0;
            if (lane == 0 && !w.overflow) {
                const uint64_t e0 = g.v_in_off[w.v0 + v], e1 = g.v_in_off[w.v0 + v + 1];
                if (e0 == e1) {
                    const uint32_t s = w.alloc();
                    PathHdr *h = w.hdr(s);
                    h->fhi = h->flo = 0;
                    h->filled = h->score_first = h->score_second = h->nverts = 0;
                    uint8_t *pv = w.verts(s);
                    for (uint32_t i = 0; i < w.V; i++) pv[i] = kAbsent;
                    w.tmp[n++] = (uint16_t)s;
                } else {
                    for (uint64_t e = e0; e < e1 && !w.overflow; e++) {
                        const uint32_t u = g.v_in_src[e];
                        merge_paths(w, n, w.lists + (size_t)u * 32, w.list_n[u]);
                        if (g.v_max_target[w.v0 + u] == v) {  // last consumer of u's paths
                            for (uint32_t i = 0; i < w.list_n[u]; i++) w.release(w.lists[(size_t)u * 32 + i]);
                            w.list_n[u] = 0;
                        }
                    }
                }
                if (!w.overflow) std_shuffle(w.tmp, n, rng);
            }
            // lane 0's view of the DP state (n, overflow) for the cooperative part
            n = __shfl_sync(0xFFFFFFFFu, n, 0);
            const bool ovf = __shfl_sync(0xFFFFFFFFu, (uint32_t)w.overflow, 0) != 0;
            __syncwarp();
            if (ovf) { w.overflow = true; break; }
            for (uint32_t i = 0; i < n; i++) add_vertex_warp(w, w.tmp[i], v, bloom, T, lane);
            if (lane == 0) {
                n = filter_paths(w, w.tmp, n, max_paths, false);
                for (uint32_t i = 0; i < n; i++) w.lists[(size_t)v * 32 + i] = w.tmp[i];
                w.list_n[v] = (uint8_t)n;
            }
            __syncwarp();
        }
        if (lane == 0) {
            uint32_t n = 0;
            uint16_t *fin = w.lists + (size_t)(w.V - 1) * 32;
            if (!w.overflow) n = filter_paths(w, fin, w.list_n[w.V - 1], max_paths, true);
            if constexpr (BATCH) {
                // addPathIndices merges order-dependently (VariantClusterGraph.cpp:726-798): wait until the earlier samples of the batch have merged
                // theirs.  Items are handed out in (cluster, sample) order and every warp that holds an earlier one is running (persistent,
                // co-resident grid), so the wait ends; the fence makes their rows of `best` visible to this warp's plain loads.
                volatile uint32_t *turn = g.turn + c;
                unsigned long long t0 = 0;
                for (uint32_t spins = 0; *turn != batch_pos; spins++) {
                    __nanosleep(64);
                    if ((spins & 0xFFFFu) == 0xFFFFu) {      // a wait of 20 s means a lost turn, not a slow cluster: report it instead of hanging the device
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > 20000000000ull) { atomicMax(g.status, 0xFFFFFFFFu); break; }
                    }
                }
                __threadfence();
            }
            if (!w.overflow) add_path_indices(w, fin, n);
            if (w.overflow) atomicMax(g.status, c + 1);
            if constexpr (BATCH) {
                __threadfence();
                atomicExch(g.turn + c, batch_pos + 1);       // also after an overflow: the later samples must not wait for ever
            }
        }
        __syncwarp();
    }
}

// best_paths_indices as one byte per (path, vertex), cluster-major, without the unused capacity rows
__global__ void k_compact_best(DevGraphs g, const uint64_t *out_off, uint8_t *out) {
    const uint32_t c = blockIdx.x;
    const uint64_t n = out_off[c + 1] - out_off[c];
    const uint8_t *src = g.best + g.best_off[c];
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) out[out_off[c] + i] = src[i] != kAbsent;
}

template <class T> T *upload(const T *h, size_t n, bool &ok) {
    T *d = nullptr;
    if (btg::dmalloc(&d, (n ? n : 1) * sizeof(T)) != cudaSuccess) { ok = false; return nullptr; }
    if (n && cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
    return d;
}

}  // namespace

struct btg_graphs {
    DevGraphs g{};
    std::vector<void *> allocs;
    std::vector<uint64_t> h_best_off, h_vertex_off;
    std::vector<uint32_t> h_best_cap;
    uint32_t n_samples_cap = 0;
    uint32_t grid = 1, grid_batch = 1;
    uint64_t scratch_max = 0;          // largest working set that does not fit the shared-memory budget (bytes, 16-aligned)
    BloomView *d_batch_blooms = nullptr;   // [n_samples_cap]
};

extern "C" {

btg_graphs *btg_graphs_upload(const btg_graphs_desc *d, uint32_t max_samples, uint32_t max_sample_haplotypes) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (!d || max_samples == 0 || max_samples > BTG_MAX_SAMPLES || max_sample_haplotypes == 0 || max_sample_haplotypes > 32) {
        set_error("bad graph descriptor (max_sample_haplotypes must be 1..32)");
        return nullptr;
    }
    const uint32_t C = d->n_clusters;
    const uint64_t Vtot = d->cl_vertex_off[C];
    auto *gr = new btg_graphs();
    bool ok = true;
    auto keep = [&](auto *p) { gr->allocs.push_back((void *)p); return p; };
    DevGraphs &g = gr->g;
    g.C = C;
    g.cl_vertex_off = keep(upload(d->cl_vertex_off, C + 1, ok));
    g.v_seq_off = keep(upload(d->v_seq_off, Vtot + 1, ok));
    g.seq = keep(upload(d->seq, d->v_seq_off[Vtot], ok));
    g.v_flags = keep(upload(d->v_flags, Vtot, ok));
    g.v_in_off = keep(upload(d->v_in_off, Vtot + 1, ok));
    g.v_in_src = keep(upload(d->v_in_src, d->v_in_off[Vtot], ok));
    g.cl_group = keep(upload(d->cl_group, C, ok));
    g.cl_idx = keep(upload(d->cl_idx, C, ok));
    // per-cluster sizing: simulate the candidate counts of the vertex DP (upper bounds)
    std::vector<uint32_t> max_target(Vtot), pool(C), tmpc(C), best_cap(C), cl_smem(C, 0);
    std::vector<uint64_t> scr_off(C + 1, 0), best_off(C + 1, 0);
    for (uint32_t c = 0; c < C; c++) {
        const uint64_t v0 = d->cl_vertex_off[c];
        const uint32_t V = (uint32_t)(d->cl_vertex_off[c + 1] - v0);
        if (V == 0 || V > 60000) { set_error("cluster %u: unsupported vertex count %u", c, V); ok = false; break; }
        for (uint32_t v = 0; v < V; v++) max_target[v0 + v] = v;
        for (uint32_t v = 0; v < V; v++)
            for (uint64_t e = d->v_in_off[v0 + v]; e < d->v_in_off[v0 + v + 1]; e++) {
                const uint32_t u = d->v_in_src[e];
                if (u >= v) { set_error("cluster %u: edge %u->%u is not topological", c, u, v); ok = false; }
                else max_target[v0 + u] = std::max(max_target[v0 + u], v);
            }
        if (!ok) break;
        std::vector<uint32_t> cnt(V, 0);
        uint32_t need_pool = 1, need_tmp = 1;
        for (uint32_t v = 0; v < V; v++) {
            uint32_t pre = 0;
            for (uint64_t e = d->v_in_off[v0 + v]; e < d->v_in_off[v0 + v + 1]; e++) pre += cnt[d->v_in_src[e]];
            if (d->v_in_off[v0 + v] == d->v_in_off[v0 + v + 1]) pre = 1;
            uint32_t alive = pre;
            for (uint32_t u = 0; u < v; u++) if (max_target[v0 + u] >= v) alive += cnt[u];
            need_pool = std::max(need_pool, alive);
            need_tmp = std::max(need_tmp, pre);
            cnt[v] = std::min(pre, max_sample_haplotypes);
        }
        if (need_pool > 60000) { set_error("cluster %u: candidate path pool too large (%u)", c, need_pool); ok = false; break; }
        pool[c] = need_pool + 1;
        tmpc[c] = need_tmp + 1;
        best_cap[c] = max_sample_haplotypes * max_samples;
        const uint64_t slot_bytes = sizeof(PathHdr) + ((V + 7) & ~7u);
        const uint64_t bytes = (uint64_t)pool[c] * slot_bytes + (uint64_t)V * 64 + (uint64_t)((tmpc[c] + 3) & ~3u) * 2 +
                               (uint64_t)((pool[c] + 3) & ~3u) * 2 + 2 * (uint64_t)((V + 7) & ~7u);
        // the working set of most clusters fits the per-warp shared-memory budget; only the others get a slice of the global arena
        if (bytes <= kPathSmemPerWarp) cl_smem[c] = (uint32_t)bytes;
        scr_off[c + 1] = scr_off[c] + (cl_smem[c] ? 0 : ((bytes + 15) & ~15ull));
        if (!cl_smem[c]) gr->scratch_max = std::max<uint64_t>(gr->scratch_max, (bytes + 15) & ~15ull);
        best_off[c + 1] = best_off[c] + (uint64_t)best_cap[c] * V;
    }
    if (ok) {
        std::vector<uint32_t> order(C);
        for (uint32_t c = 0; c < C; c++) order[c] = c;
        auto shape = [&](uint32_t c) {
            const uint64_t v0 = d->cl_vertex_off[c], v1 = d->cl_vertex_off[c + 1];
            return std::make_pair((uint64_t)(v1 - v0), (uint64_t)(d->v_seq_off[v1] - d->v_seq_off[v0]));
        };
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return shape(a) > shape(b); });  // largest first
        g.order = keep(upload(order.data(), C, ok));
        g.v_max_target = keep(upload(max_target.data(), Vtot, ok));
        g.scr_off = keep(upload(scr_off.data(), C + 1, ok));
        g.cl_pool = keep(upload(pool.data(), C, ok));
        g.cl_tmp = keep(upload(tmpc.data(), C, ok));
        g.best_off = keep(upload(best_off.data(), C + 1, ok));
        g.best_cap = keep(upload(best_cap.data(), C, ok));
        g.cl_smem = keep(upload(cl_smem.data(), C, ok));
        // persistent grid: as many blocks as are co-resident; one Mersenne state per warp of the grid
        int per_sm = 0;
        const size_t smem = 2048 + (size_t)kPathWarps * kPathSmemPerWarp;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_find_sample_paths<false>, kPathWarps * 32, smem);
        gr->grid = (uint32_t)std::max(1, per_sm) * (uint32_t)ctx().sm_count;
        int per_sm_batch = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_batch, k_find_sample_paths<true>, kPathWarps * 32, smem);
        gr->grid_batch = (uint32_t)std::max(1, std::min(per_sm, per_sm_batch)) * (uint32_t)ctx().sm_count;   // <= grid: the Mersenne pool is sized by `grid`
        uint8_t *scratch = nullptr, *best = nullptr;
        uint32_t *best_n = nullptr, *status = nullptr, *mt_pool = nullptr;
        unsigned long long *stats = nullptr;
        ok = ok && btg::dmalloc(&scratch, scr_off[C] + 16) == cudaSuccess;
        ok = ok && btg::dmalloc(&best, best_off[C] + 16) == cudaSuccess;
        ok = ok && btg::dmalloc(&best_n, (C + 1) * 4) == cudaSuccess && btg::dmalloc(&status, 2 * 4) == cudaSuccess;
        ok = ok && btg::dmalloc(&mt_pool, (size_t)gr->grid * kPathWarps * 624 * 4) == cudaSuccess;
        ok = ok && btg::dmalloc(&stats, 3 * 8) == cudaSuccess;
        keep(scratch); keep(best); keep(best_n); keep(status); keep(mt_pool); keep(stats);
        uint32_t *turn = nullptr;
        ok = ok && btg::dmalloc(&turn, ((size_t)C + 1) * 4) == cudaSuccess && btg::dmalloc(&gr->d_batch_blooms, (size_t)max_samples * sizeof(BloomView)) == cudaSuccess;
        keep(turn); keep(gr->d_batch_blooms);
        g.turn = turn;
        if (ok) cudaMemsetAsync(stats, 0, 3 * 8, ctx().stream);
        g.stats = stats;
        if (ok) {
            cudaMemsetAsync(best_n, 0, (C + 1) * 4, ctx().stream);
            cudaMemsetAsync(status, 0, 2 * 4, ctx().stream);
        }
        g.scratch = scratch; g.best = best; g.best_n = best_n; g.status = status; g.mt_pool = mt_pool; g.next = status + 1;
    }
    if (!ok) {
        if (!*btg_last_error()) set_error("graph upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
        btg_graphs_free(gr);
        return nullptr;
    }
    gr->h_best_off = best_off;
    gr->h_best_cap = best_cap;
    gr->h_vertex_off.assign(d->cl_vertex_off, d->cl_vertex_off + C + 1);
    gr->n_samples_cap = max_samples;
    return gr;
}

int btg_graphs_path_stats(const btg_graphs *gr, uint64_t *out3) {
    BTG_REQUIRE_INIT();
    if (!gr || !out3) { set_error("null argument"); return BTG_EINVAL; }
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    BTG_CUDA(cudaMemcpy(out3, gr->g.stats, 3 * 8, cudaMemcpyDeviceToHost));
    return BTG_OK;
}

int btg_graphs_reset(btg_graphs *gr) {
    BTG_REQUIRE_INIT();
    if (!gr) { set_error("null argument"); return BTG_EINVAL; }
    BTG_CUDA(cudaMemsetAsync(gr->g.best_n, 0, ((size_t)gr->g.C + 1) * 4, ctx().stream));
    BTG_CUDA(cudaMemsetAsync(gr->g.status, 0, 2 * 4, ctx().stream));
    BTG_CUDA(cudaMemsetAsync(gr->g.stats, 0, 3 * 8, ctx().stream));
    return BTG_OK;
}

void btg_graphs_free(btg_graphs *gr) {
    if (!gr) return;
    cudaStreamSynchronize(ctx().stream);
    for (void *p : gr->allocs) btg::dfree(p);
    delete gr;
}

int btg_find_sample_paths(btg_graphs *gr, const btg_bloom *sample_bloom, uint32_t sample_idx, uint32_t random_seed, uint32_t max_sample_haplotypes) {
    BTG_REQUIRE_INIT();
    if (!gr || !sample_bloom) { set_error("null argument"); return BTG_EINVAL; }
    if (sample_idx >= gr->n_samples_cap) { set_error("sample index %u beyond the capacity given at upload (%u)", sample_idx, gr->n_samples_cap); return BTG_EINVAL; }
    if (max_sample_haplotypes == 0 || max_sample_haplotypes > 32) { set_error("max_sample_haplotypes must be 1..32"); return BTG_EINVAL; }
    if (gr->g.C == 0) return BTG_OK;
    // asynchronous on the library stream: the samples of a unit queue up behind each other (addPathIndices merges in sample order);
    // a scratch overflow is reported by btg_get_best_paths
    auto s = ctx().stream;
    BTG_CUDA(cudaMemsetAsync(gr->g.next, 0, 4, s));
    const size_t smem = 2048 + (size_t)kPathWarps * kPathSmemPerWarp;
    k_find_sample_paths<false><<<gr->grid, kPathWarps * 32, smem, s>>>(gr->g, btg_internal::bloom_view(sample_bloom), random_seed, sample_idx, max_sample_haplotypes);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

// KmerCounter::findVariantClusterPaths for SEVERAL samples in one launch.  The reference searches one sample at a time because it holds one
// sample's Bloom filter at a time (main.cpp:219-247); the searches themselves are independent — only addPathIndices, which merges a sample's
// paths into the cluster's best paths, depends on the sample order — so with the filters of the batch resident the (cluster, sample) pairs
// run side by side and each cluster merges in sample order (k_find_sample_paths<true>).  A launch of one sample is bounded by the sequential
// vertex DP of the unit's slowest cluster; a batch pays that latency once instead of once per sample.  Same best paths, bit for bit.
int btg_find_sample_paths_batch(btg_graphs *gr, const btg_bloom *const *sample_blooms, uint32_t sample_first, uint32_t n_samples, uint32_t random_seed,
                                uint32_t max_sample_haplotypes) {
    BTG_REQUIRE_INIT();
    if (!gr || !sample_blooms) { set_error("null argument"); return BTG_EINVAL; }
    if (n_samples == 0) return BTG_OK;
    if ((uint64_t)sample_first + n_samples > gr->n_samples_cap) {
        set_error("samples %u..%u beyond the capacity given at upload (%u)", sample_first, sample_first + n_samples - 1, gr->n_samples_cap);
        return BTG_EINVAL;
    }
    if (max_sample_haplotypes == 0 || max_sample_haplotypes > 32) { set_error("max_sample_haplotypes must be 1..32"); return BTG_EINVAL; }
    for (uint32_t i = 0; i < n_samples; i++) if (!sample_blooms[i]) { set_error("null Bloom filter for sample %u", sample_first + i); return BTG_EINVAL; }
    if (gr->g.C == 0) return BTG_OK;
    // working sets beyond the shared-memory budget: one slice per resident warp (the per-cluster slices of the one-sample kernel would be shared
    // by the samples of a cluster); allocated on first use, kept with the graphs.  Too large (or a batch of one): the samples one after the other.
    const uint64_t n_warps = (uint64_t)gr->grid_batch * kPathWarps;
    const uint64_t need = gr->scratch_max * n_warps;
    bool batch = n_samples > 1 && (uint64_t)gr->g.C * n_samples < 0xFFFFFFFFull;
    if (batch && need && !gr->g.warp_scratch) {
        const uint64_t cap = std::min<uint64_t>(btg::free_device_memory() / 4, 16ull << 30);
        uint8_t *ws = nullptr;
        if (need <= cap && btg::dmalloc(&ws, need + 16) == cudaSuccess) {
            gr->allocs.push_back(ws);
            gr->g.warp_scratch = ws;
            gr->g.warp_scratch_bytes = gr->scratch_max;
        } else {
            cudaGetLastError();
            batch = false;
        }
    }
    if (!batch) {
        for (uint32_t i = 0; i < n_samples; i++) {
            const int rc = btg_find_sample_paths(gr, sample_blooms[i], sample_first + i, random_seed, max_sample_haplotypes);
            if (rc != BTG_OK) return rc;
        }
        return BTG_OK;
    }
    auto s = ctx().stream;
    std::vector<BloomView> views(n_samples);
    for (uint32_t i = 0; i < n_samples; i++) views[i] = btg_internal::bloom_view(sample_blooms[i]);
    BTG_CUDA(cudaMemcpyAsync(gr->d_batch_blooms, views.data(), (size_t)n_samples * sizeof(BloomView), cudaMemcpyHostToDevice, s));   // pageable source: staged before the call returns
    BTG_CUDA(cudaMemsetAsync(gr->g.next, 0, 4, s));
    BTG_CUDA(cudaMemsetAsync(gr->g.turn, 0, ((size_t)gr->g.C + 1) * 4, s));
    DevGraphs g = gr->g;
    g.batch_n = n_samples;
    g.batch_first = sample_first;
    g.batch_blooms = gr->d_batch_blooms;
    const size_t smem = 2048 + (size_t)kPathWarps * kPathSmemPerWarp;
    k_find_sample_paths<true><<<gr->grid_batch, kPathWarps * 32, smem, s>>>(g, BloomView{}, random_seed, sample_first, max_sample_haplotypes);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_get_best_paths(const btg_graphs *gr, uint32_t *n_paths_out, uint64_t *path_off_out, uint8_t *membership_out, uint64_t membership_bytes) {
    BTG_REQUIRE_INIT();
    if (!gr || !n_paths_out) { set_error("null argument"); return BTG_EINVAL; }
    const uint32_t C = gr->g.C;
    auto s = ctx().stream;
    BTG_CUDA(cudaStreamSynchronize(s));
    uint32_t status = 0;
    BTG_CUDA(cudaMemcpy(&status, gr->g.status, 4, cudaMemcpyDeviceToHost));
    if (status == 0xFFFFFFFFu) { set_error("batched path search: a cluster waited 20 s for the paths of an earlier sample (lost turn)"); return BTG_ECUDA; }
    if (status) { set_error("cluster %u: path-search scratch overflow", status - 1); return BTG_ESTATE; }
    BTG_CUDA(cudaMemcpy(n_paths_out, gr->g.best_n, C * 4, cudaMemcpyDeviceToHost));
    if (!path_off_out) return BTG_OK;
    path_off_out[0] = 0;
    for (uint32_t c = 0; c < C; c++) path_off_out[c + 1] = path_off_out[c] + (uint64_t)n_paths_out[c] * (gr->h_vertex_off[c + 1] - gr->h_vertex_off[c]);
    if (!membership_out) return BTG_OK;
    if (membership_bytes < path_off_out[C]) { set_error("membership buffer too small"); return BTG_EINVAL; }
    if (C == 0 || path_off_out[C] == 0) return BTG_OK;
    // compact on the device (the capacity rows of best_paths_indices never cross the bus), then one copy out
    uint64_t *d_off = nullptr;
    uint8_t *d_out = nullptr;
    int rc = BTG_OK;
    if (btg::dmalloc(&d_off, (C + 1) * 8) != cudaSuccess || btg::dmalloc(&d_out, path_off_out[C]) != cudaSuccess) { set_error("best-path compaction: allocation failed"); rc = BTG_ENOMEM; }
    if (rc == BTG_OK) {
        cudaMemcpyAsync(d_off, path_off_out, (C + 1) * 8, cudaMemcpyHostToDevice, s);
        k_compact_best<<<C, 64, 0, s>>>(gr->g, d_off, d_out);
        BTG_LAUNCHED();
        cudaMemcpyAsync(membership_out, d_out, path_off_out[C], cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("best-path compaction failed: %s", cudaGetErrorString(cudaGetLastError())); rc = BTG_ECUDA; }
    }
    btg::dfree(d_off); btg::dfree(d_out);
    return rc;
}

}  // extern "C"
