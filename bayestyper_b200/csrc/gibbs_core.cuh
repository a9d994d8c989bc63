// gibbs_core.cuh — device-side data model and per-cluster steps of the Gibbs sampler, shared by the kernels of
// gibbs.cu (one thread per cluster / group) and gibbs_wide.cu (one warp per cluster / group, lane = sample).
// See gibbs.cu for the reference files this restates.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <type_traits>
#include <vector>
#include <chrono>
#include <functional>

#include <cooperative_groups.h>

#include "common.cuh"
#include "gibbs_rng.cuh"
#include "comm.cuh"

namespace cg = cooperative_groups;

// -DBTG_NOISE_TIMING=1 compiles per-cluster timers into the chain kernel (slowest cluster per iteration, clock sums of the
// sub-steps); they cost registers, so the default build only keeps block 0's four phase laps (BTG_NOISE_PHASES=1 at run time)
#ifndef BTG_NOISE_TIMING
#define BTG_NOISE_TIMING 0
#endif

using namespace btg;

namespace btg_gibbs {


constexpr uint16_t NONE = 0xFFFF;  // Utils::ushort_overflow
// index in the whole unit of group g of this btg_unit (btg_gibbs_opts: base + g * stride)
__host__ __device__ inline uint64_t group_index(const btg_gibbs_opts &o, uint32_t g) {
    return o.group_index_base + (uint64_t)g * (o.group_index_stride ? o.group_index_stride : 1u);
}
constexpr double kDoubleEps100 = 2.220446049250313e-16 * 100;
constexpr float kFloatEps100 = 1.1920929e-07f * 100;

__host__ __device__ inline bool doubleCompare(double a, double b) {  // Utils.hpp:81-87
    return (a == b) || (fabs(a - b) < fabs(a < b ? a : b) * kDoubleEps100);
}
__host__ __device__ inline bool floatCompare(float a, float b) {  // Utils.hpp:89-95
    return (a == b) || (fabsf(a - b) < fabsf(a < b ? a : b) * kFloatEps100);
}
__host__ __device__ inline bool floatLess(float a, float b) { return (a < b) && !floatCompare(a, b); }
// Utils::logAddition (Utils.hpp:105-124): hi + log1p(exp(lo - hi)).  When lo - hi < -45 the addend is below exp(-45) = 2.9e-20, less than
// half an ulp of any |hi| >= 2^-9, so the sum rounds to hi whatever the libm: the two transcendental calls are skipped and the result is the
// same bits.  Most diplotypes of a large cluster are that far below the best one, and the running sum is a chain of these calls.
BTG_LEAF double logAddition(double a, double b) {
    const double hi = a < b ? b : a, lo = a < b ? a : b;
    if (lo - hi < -45.0 && fabs(hi) >= 0.001953125) return hi;
    return hi + m_log1p(m_exp(lo - hi));
}

static __device__ double nbLogPmf(double p, double size, uint32_t obs, uint32_t scale) {  // NegativeBinomialDistribution.cpp:121-147
    const double coef = lgamma(obs + size * scale) - lgamma(size * scale) - lgamma((double)(obs + 1));
    return coef + log(p) * size * scale + log(1 - p) * obs;
}
static __device__ double poissonLogProb(uint32_t value, double rate) {  // CountDistribution.cpp:349-352
    return value * log(rate) - rate - lgamma((double)(value + 1));
}

// CountDistribution::updateNoiseCache / noiseCountLogPmf (CountDistribution.cpp:240-253,314-347): one thread per (s, c)
static __device__ double noiseCountLogPmf(double rate, uint32_t c) {
    double v = poissonLogProb(c, rate);
    if (c == 255) {
        uint32_t limit = c;
        double prev;
        do {
            limit++;
            prev = v;
            v = logAddition(v, poissonLogProb(limit, rate));
            if (v > 0) { v = 0; break; }
        } while (!doubleCompare(prev, v));
    }
    return v;
}
}  // namespace btg_gibbs

struct btg_count_dist {
    uint32_t S = 0;
    double *p = nullptr, *size = nullptr, *rates = nullptr;  // device [S]
    double *genomic = nullptr;                               // device [S][256][256]
    double *noise = nullptr;                                 // device [S][256]
    float prior_shape = 1.f, prior_scale = 0.01f;
    std::vector<double> h_p, h_size;
};

// ---------------------------------------------------------------------------------------------
// unit on the device
// ---------------------------------------------------------------------------------------------
namespace btg_gibbs {


// Arena layout.  Clusters are sorted by cost and dealt to lanes in that order; the 32 clusters that share a
// warp share one arena SLOT whose arrays are interleaved across lanes (element e of lane l lives at
// base + e*32 + l).  A warp reading "the same field" of its 32 clusters therefore touches one or two 128 B
// lines instead of 32 scattered ones, and the hot state of the resident warps stays in L1.
struct ClusterLayout {
    uint32_t group;                     // owning group
    uint32_t n_alleles;                 // sum of numberOfAlleles over the cluster's variants
    uint32_t Dall;                      // (H+1)(H+2)/2 diplotype slots (index H = "missing")
    uint32_t pos;                       // position in the cost order: slot = pos >> 5, lane = pos & 31
};
struct SlotLayout {
    uint64_t f64_off, u32_off, u8_off;  // element offsets of the slot in the three pools
    uint32_t H, K, nvar, n_uniq, n_alleles, Dall;  // per-lane capacities = max over the slot's clusters
    uint32_t n_multi;                   // multicluster k-mers (0 for every slot of single-cluster groups)
    uint32_t has_cache;                 // dense per-(sample, diplotype) caches allocated (always in the interleaved layout)
    uint32_t cum_rows;                  // one row of cumulative log-probs per sample (wide layout; slots holding clusters of nested groups,
                                        // which the warp-per-group kernel of the joint mode works on with one lane per sample)
};
// st = 32 for the lane-interleaved slots, 1 for the dense slots of the wide layout (one cluster per slot, worked on by a whole warp)
template <class T> struct LaneArr {
    T *p;
    uint32_t st;
    __device__ __forceinline__ T &operator[](uint32_t i) const { return p[(size_t)i * st]; }
    __device__ __forceinline__ LaneArr<T> operator+(size_t i) const { return LaneArr<T>{p + i * st, st}; }
};

// k-mer tile accessor: lane-interleaved (stride 32) for clusters that run one per thread — the 32 clusters of a warp
// read one sector per element — and dense (stride 1) for the large clusters that a whole warp works on, where all lanes
// read the same row and an interleaved layout would cost one sector per byte
struct TileArr {
    uint8_t *p;
    uint32_t stride;
    __device__ __forceinline__ uint8_t &operator[](uint32_t i) const { return p[(size_t)i * stride]; }
};
constexpr uint32_t kBigFillCost = 384;   // sweep on configs[1] with the hot state and the merged phases: profiles/r2_noise_chain_big_sweep.txt (128 was the round-1 value)
constexpr uint32_t kBigFillCostWarp = 128;   // the same bound for the warp-per-cluster kernel (lane = sample): per-lane lookups of one fill
constexpr uint32_t kChainSplit = 20;      // virtual threads of a chain-split cluster (chain c runs on thread c % kChainSplit)
constexpr uint32_t kSplitFillCost = 64;   // clusters above this fill cost are chain-split in the default mode (sweep: profiles/r1_gibbs_tail.txt)
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

struct DevUnit {
    uint32_t S, G, C;
    uint32_t wide;               // arena layout: 0 = 32 clusters interleaved per slot (one thread per cluster), 1 = one dense slot per cluster (one warp per cluster, lane = sample)
    const uint8_t *sample_gender, *group_ploidy;
    const uint64_t *group_cluster_off;
    const uint32_t *cluster_idx, *cl_nhap;
    const uint64_t *cl_kmer_off, *cl_var_off, *cl_mult_off;
    const uint8_t *mult, *k_has_counts, *k_counts, *k_ic;
    const uint64_t *cl_uniq_off;
    const uint32_t *uniq_idx;
    const uint64_t *kmer_vh_off;
    const uint16_t *vh_var;
    const uint64_t *vh_bits_off;
    const uint8_t *vh_bits;
    const uint64_t *cl_hapvar_off;
    const uint16_t *hap_alleles, *var_nalleles;
    const uint8_t *var_dep;
    const uint64_t *valt_off;    // prefix sums of numberOfAlleles over all variants
    const ClusterLayout *layout;
    const SlotLayout *slots;
    const uint32_t *order;       // clusters sorted by decreasing cost
    double *f64_pool;
    uint32_t *u32_pool;
    uint8_t *u8_pool;
    const double *lgamma_int;
    const uint64_t *big_tile_off;  // [C] offset of the cluster's dense tile in big_tile_pool, ~0 for one-thread clusters
    uint8_t *big_tile_pool;
    // nested groups / multicluster k-mers
    uint32_t n_regular;          // order[0 .. n_regular): clusters of single-cluster groups
    // chain-split clusters: large single-cluster groups whose chains run as kChainSplit independent threads (default mode)
    uint32_t n_split;
    const uint32_t *split_cluster;   // [n_split] cluster
    const uint32_t *split_pos;       // [n_split][kChainSplit] arena position of virtual thread v (v = 0: the cluster's own position)
    const uint32_t *split_of;        // [C] index into split_cluster or NONE32
    uint32_t n_nested_groups;
    const uint32_t *nested_groups;
    const uint32_t *k_shared;    // [rows] shared multiplicity record of a multicluster k-mer
    const uint64_t *cl_multi_off;
    const uint32_t *multi_idx;
    const uint64_t *hap_start;   // [C+1] first haplotype of each cluster in hap_nested_off
    const uint64_t *hap_nested_off;
    const uint32_t *hap_nested;
    const uint64_t *cl_dep_off;
    const uint32_t *dep_cluster;
    const uint64_t *dep_var_off;
    const uint16_t *dep_var;
    const uint64_t *group_src_off;
    const uint32_t *group_src;
    const uint64_t *cl_edge_off; // [C+1] out-edges of each cluster (CSR over clusters, targets = local indices)
    const uint32_t *edge_dst;
    uint32_t *src_mut, *edge_mut;        // branch orderings as shuffled so far (VariantClusterGroup::shuffleBranchOrdering)
    uint32_t *dfs_order, *dfs_stack;     // [C] scratch of the group threads
    uint8_t *shared_mult;                // [records][S] KmerCounts::multiplicities
    const uint32_t *nest_slot;           // [C] index into the nested-info arrays, NONE for single-cluster groups
    uint8_t *nest_pl, *nest_k;           // [slots][S] NestedVariantClusterInfo: nested_ploidy, number of nested_kmer_stats
    uint32_t *nest_n;                    // [slots][S][2] KmerStats count
    double *nest_f;                      // [slots][S][2][2] KmerStats (fraction, mean)
};

// arena sizes (elements) of one cluster — must match the pointer carving in Cl::bind
constexpr uint32_t kSimplexTableMaxH = 32;
struct ArenaSizes { uint64_t f64, u32, u8; };
// wide layout: the cumulative log-probs are kept per sample (the S draws of an iteration run side by side); clusters whose dense
// caches would exceed kWideCacheCap entries get none (has_cache = 0): their diplotype log-probs are recomputed when needed —
// the value of an entry does not depend on whether it was cached
constexpr uint64_t kWideCacheCap = 16384;
constexpr uint32_t kWideMinSamples = 8;   // units with at least this many samples get the wide layout (BTG_WIDE=0/1 overrides)
__host__ __device__ inline bool arena_has_cache(uint32_t S, uint32_t Dall, uint32_t n_multi, bool wide, uint64_t cap = kWideCacheCap) {
    return !wide || n_multi > 0 || (uint64_t)S * Dall <= cap;
}
__host__ __device__ inline ArenaSizes arena_sizes(uint32_t S, uint32_t H, uint32_t K, uint32_t nvar, uint32_t n_uniq, uint32_t n_alleles, uint32_t Dall,
                                                  uint32_t n_multi, bool wide = false, uint64_t cache_cap = kWideCacheCap, bool cum_rows = false) {
    const bool cache = arena_has_cache(S, Dall, n_multi, wide, cache_cap);
    ArenaSizes a;
    a.f64 = (uint64_t)H * 2        /* freq, log(freq) */
          + (H + 1)                /* simplex prob vector (H > kSimplexTableMaxH: most recent key only) */
          + (H <= kSimplexTableMaxH ? (uint64_t)H * (H + 1) : 0) /* else: one vector per plus-count, + lengths */
          + (cache ? (uint64_t)S * Dall : 0)     /* unique diplotype log-prob cache */
          + (cache ? (wide || cum_rows ? (uint64_t)S * Dall : Dall) : 0)   /* cumulative log-probs of one draw (or of one draw per sample) */
          + (uint64_t)S * 2 * nvar * 2   /* k-mer stats cache (fraction, mean) */
          + (uint64_t)n_alleles * S * 3  /* allele k-mer stats: 3 sums (count, fraction, mean) */
          + (n_multi ? (uint64_t)S * Dall : 0)  /* multicluster diplotype log-prob cache */
          + 2;                     /* sparsity, spare */
    a.u32 = (uint64_t)H            /* observation counts */
          + n_uniq * 2ull          /* unique k-mer order, subset */
          + (uint64_t)H * nvar     /* subset counters per (haplotype, variant) */
          + (uint64_t)Dall * S     /* diplotype tallies */
          + (uint64_t)S * 2 * nvar /* k-mer stats cache counts */
          + (uint64_t)n_alleles * S * 3  /* allele stats counts */
          + S                      /* current diplotypes (first | second << 16) */
          + n_multi * 2ull         /* multicluster k-mer order, subset */
          + 32;                    /* misc + rng states */
    a.u8 = (uint64_t)H + K + S     /* non-zero flags, uncovered rows, stats-cache update flags */
         + (uint64_t)n_multi * S   /* sample_multicluster_kmer_multiplicities */
         + (uint64_t)n_uniq * (H + S + 2);  /* k-mer tile of the current subsample: multiplicities, counts, (F, M) inter-cluster multiplicity */
    a.u8 = (a.u8 + 7) & ~7ull;
    return a;
}

enum Misc { kNSub = 0, kNumHap = 1, kNumMissing = 2, kSimplexNobs = 3, kSimplexPlus = 4, kSimplexLen = 5, kSparse = 6, kCover = 7, kNMultiSub = 8, kUseMulti = 9,
            kMiscWords = 12 };
enum RngSave { kRng0 = 0, kRng1 = 9, kRngWords = 20 };   // a saved stream takes 9 words (Philox::save)

// HOT STATE IN SHARED MEMORY.  The sampler's step is a chain of dependent accesses to the cluster's own small arrays (frequencies,
// counts, caches, cumulative log-probs, the k-mer tile); in the arena every one of them is an L2 round trip, and that latency was the
// whole cost of an iteration (85-110 k cycles per one-thread SNV cluster, profiles/r1_noise_chain_phases.txt).  A thread that keeps ONE
// cluster for a whole chain binds these arrays to a slice of the block's shared memory instead (element e of thread t at e * blockDim + t:
// conflict-free), when the cluster has at most kHotH haplotype candidates — > 90 % of the clusters of a real unit.  Same code, same
// arithmetic: only the addresses change.
constexpr uint32_t kHotH = 2, kHotDall = (kHotH + 1) * (kHotH + 2) / 2, kHotTile = 24;
__host__ __device__ inline uint32_t hot_f64(uint32_t S) { return kHotH * 2 + (kHotH + 1) + kHotH * (kHotH + 1) + S * kHotDall + kHotDall + 2; }
__host__ __device__ inline uint32_t hot_u32(uint32_t S) { return kHotH + S + kMiscWords; }
__host__ __device__ inline uint32_t hot_u8(uint32_t S) { return (kHotH + S + kHotTile * (kHotH + S + 2) + 3u) & ~3u; }
__host__ __device__ inline uint32_t hot_bytes(uint32_t S) { return hot_f64(S) * 8 + hot_u32(S) * 4 + hot_u8(S); }

// per-cluster view
struct Cl {
    uint32_t S, H, K, nvar, n_uniq, Dall, n_alleles, c, g, n_multi;
    bool has_simplex_tab, has_cache;
    uint32_t cum_stride;   // wide layout: the cumulative log-probs of sample s start at cum[s * cum_stride] (0: one shared row)
    uint64_t row0, var0;
    const DevUnit *u;
    const uint8_t *M;
    LaneArr<double> freq, logf, simplex, simplex_tab, ucache, cum, kc_f, as_f, mcache, fmisc;
    LaneArr<uint32_t> obs, uniq, uniq_sub, cnt, tally, kc_n, as_n, dipl, multi, multi_sub, misc, rng;
    LaneArr<uint8_t> nz, uncovered, stats_update, sample_multi;
    TileArr tile_m, tile_c, tile_ic;
    uint8_t *hot_tile = nullptr;   // shared-memory tile of a cluster whose hot state is bound to shared memory (bind_hot)
    uint32_t hot_stride = 0;

    __device__ void bind(const DevUnit &du, uint32_t cluster, uint32_t pos_override = 0xFFFFFFFFu) {
        u = &du; c = cluster;
        ClusterLayout L = du.layout[c];
        if (pos_override != 0xFFFFFFFFu) L.pos = pos_override;
        g = L.group;
        S = du.S;
        H = du.cl_nhap[c];
        row0 = du.cl_kmer_off[c];
        K = (uint32_t)(du.cl_kmer_off[c + 1] - row0);
        var0 = du.cl_var_off[c];
        nvar = (uint32_t)(du.cl_var_off[c + 1] - var0);
        n_uniq = (uint32_t)(du.cl_uniq_off[c + 1] - du.cl_uniq_off[c]);
        n_multi = (uint32_t)(du.cl_multi_off[c + 1] - du.cl_multi_off[c]);
        Dall = L.Dall;
        n_alleles = L.n_alleles;
        M = du.mult + du.cl_mult_off[c];
        const SlotLayout SL = du.slots[du.wide ? L.pos : L.pos >> 5];
        const uint32_t lane = du.wide ? 0u : L.pos & 31u, st = du.wide ? 1u : 32u;
        has_cache = SL.has_cache;
        cum_stride = SL.cum_rows ? SL.Dall : 0u;
        LaneArr<double> f{du.f64_pool + SL.f64_off + lane, st};
        freq = f; f = f + SL.H;
        logf = f; f = f + SL.H;
        simplex = f; f = f + (SL.H + 1);
        has_simplex_tab = SL.H <= kSimplexTableMaxH;
        simplex_tab = f; f = f + (has_simplex_tab ? (uint64_t)SL.H * (SL.H + 1) : 0);
        ucache = f; f = f + (has_cache ? (uint64_t)S * SL.Dall : 0);
        cum = f; f = f + (has_cache ? (SL.cum_rows ? (uint64_t)S * SL.Dall : SL.Dall) : 0);
        kc_f = f; f = f + (uint64_t)S * 2 * SL.nvar * 2;
        as_f = f; f = f + (uint64_t)SL.n_alleles * S * 3;
        mcache = f; f = f + (SL.n_multi ? (uint64_t)S * SL.Dall : 0);
        fmisc = f;
        LaneArr<uint32_t> w{du.u32_pool + SL.u32_off + lane, st};
        obs = w; w = w + SL.H;
        uniq = w; w = w + SL.n_uniq;
        uniq_sub = w; w = w + SL.n_uniq;
        cnt = w; w = w + (uint64_t)SL.H * SL.nvar;
        tally = w; w = w + (uint64_t)SL.Dall * S;
        kc_n = w; w = w + (uint64_t)S * 2 * SL.nvar;
        as_n = w; w = w + (uint64_t)SL.n_alleles * S * 3;
        dipl = w; w = w + S;
        multi = w; w = w + SL.n_multi;
        multi_sub = w; w = w + SL.n_multi;
        misc = w; w = w + kMiscWords;
        rng = w;
        LaneArr<uint8_t> b{du.u8_pool + SL.u8_off + lane, st};
        nz = b; b = b + SL.H;
        uncovered = b; b = b + SL.K;
        stats_update = b; b = b + S;
        sample_multi = b; b = b + (uint64_t)SL.n_multi * S;
        const uint64_t dense = du.big_tile_off[c];
        if (dense != ~0ull) {
            uint8_t *t = du.big_tile_pool + dense;
            tile_m = TileArr{t, 1}; t += (size_t)n_uniq * H;
            tile_c = TileArr{t, 1}; t += (size_t)n_uniq * S;
            tile_ic = TileArr{t, 1};
        } else {
            tile_m = TileArr{b.p, st}; b = b + (uint64_t)SL.n_uniq * SL.H;
            tile_c = TileArr{b.p, st}; b = b + (uint64_t)SL.n_uniq * S;
            tile_ic = TileArr{b.p, st};
        }
    }
    // rebinds the hot arrays to this thread's slice of the block's shared memory (smem: hot_bytes(S) * nthreads bytes); requires H <= kHotH
    // and no multicluster k-mers.  Whatever the arrays held in the arena is NOT carried over (see hot_copy).
    __device__ void bind_hot(uint8_t *smem, uint32_t tid, uint32_t nthreads) {
        LaneArr<double> f{reinterpret_cast<double *>(smem) + tid, nthreads};
        freq = f; f = f + kHotH;
        logf = f; f = f + kHotH;
        simplex = f; f = f + (kHotH + 1);
        has_simplex_tab = true;
        simplex_tab = f; f = f + kHotH * (kHotH + 1);
        ucache = f; f = f + S * kHotDall;
        cum = f; f = f + kHotDall;
        cum_stride = 0;
        fmisc = f;
        LaneArr<uint32_t> w{reinterpret_cast<uint32_t *>(smem + (size_t)hot_f64(S) * 8 * nthreads) + tid, nthreads};
        obs = w; w = w + kHotH;
        dipl = w; w = w + S;
        misc = w;
        LaneArr<uint8_t> b{smem + ((size_t)hot_f64(S) * 8 + (size_t)hot_u32(S) * 4) * nthreads + tid, nthreads};
        nz = b; b = b + kHotH;
        stats_update = b; b = b + S;
        hot_tile = b.p;
        hot_stride = nthreads;
    }
    // the k-mer tile of the current subsample moves into the shared slice when it has at most kHotTile rows (after cl_reset built it)
    __device__ __forceinline__ bool tile_fits_hot() const { return hot_tile != nullptr && misc[kNSub] <= kHotTile; }
    __device__ void bind_hot_tile() {
        uint8_t *t = hot_tile;
        tile_m = TileArr{t, hot_stride}; t += (size_t)kHotTile * kHotH * hot_stride;
        tile_c = TileArr{t, hot_stride}; t += (size_t)kHotTile * S * hot_stride;
        tile_ic = TileArr{t, hot_stride};
    }
    __device__ void move_tile_to_hot() {
        const uint32_t n_sub = misc[kNSub];
        const TileArr am = tile_m, ac = tile_c, ai = tile_ic;
        bind_hot_tile();
        for (uint32_t i = 0; i < n_sub * H; i++) tile_m[i] = am[i];
        for (uint32_t i = 0; i < n_sub * S; i++) tile_c[i] = ac[i];
        for (uint32_t i = 0; i < n_sub * 2; i++) tile_ic[i] = ai[i];
    }
    // k-mer tile (lock-step modes, where the diplotype caches are cleared every iteration): row i holds everything the
    // likelihood reads about the i-th k-mer of the current subsample, so the per-iteration gathers touch three compact
    // lane-interleaved byte arrays instead of five scattered unit arrays
    __device__ __forceinline__ uint8_t tileDiplMult(uint32_t i, uint32_t a, uint32_t b) const {
        uint8_t r = 0;
        if (a != NONE) r += tile_m[i * H + a];
        if (b != NONE) r += tile_m[i * H + b];
        return r;
    }
    __device__ __forceinline__ uint8_t m(uint32_t k, uint32_t h) const { return M[(size_t)k * H + h]; }
    __device__ __forceinline__ uint8_t count(uint32_t k, uint32_t s) const { return u->k_has_counts[row0 + k] ? u->k_counts[(row0 + k) * S + s] : 0; }
    __device__ __forceinline__ uint8_t ic(uint32_t k, uint32_t s) const { return u->k_has_counts[row0 + k] ? u->k_ic[(row0 + k) * 2 + u->sample_gender[s]] : 0; }
    __device__ __forceinline__ uint16_t nalleles(uint32_t v) const { return u->var_nalleles[var0 + v]; }
    __device__ __forceinline__ bool isMissing(uint32_t v, uint16_t a) const { return u->var_dep[var0 + v] && a == nalleles(v) - 1; }  // VariantInfo.hpp:82-94
    __device__ __forceinline__ uint16_t hapAllele(uint32_t h, uint32_t v) const { return u->hap_alleles[u->cl_hapvar_off[c] + (size_t)h * nvar + v]; }
    __device__ __forceinline__ uint32_t alleleBase(uint32_t v, uint32_t s) const {  // index of (v, s, allele 0) in allele-major arrays
        return (uint32_t)(u->valt_off[var0 + v] - u->valt_off[var0]) * S + s * nalleles(v);
    }
    // dense diplotype slot: h in [0,H], H = missing; first <= second
    __device__ __forceinline__ uint32_t slot(uint32_t a, uint32_t b) const { return b * (b + 1) / 2 + a; }
    __device__ __forceinline__ uint8_t diplMult(uint32_t k, uint32_t a, uint32_t b) const {  // …Haplotypes.cpp:45-61
        uint8_t r = 0;
        if (a != NONE) r += m(k, a);
        if (b != NONE) r += m(k, b);
        return r;
    }
    // KmerCounts::getSampleMultiplicity of a multicluster k-mer (KmerCounts.cpp:205-224)
    __device__ __forceinline__ uint8_t &sharedMult(uint32_t k, uint32_t s) const { return u->shared_mult[(size_t)u->k_shared[row0 + k] * S + s]; }
    // VariantClusterHaplotypes::getMulticlusterKmerMultiplicity (VariantClusterHaplotypes.cpp:76-93); (pa, pb) = current diplotype
    __device__ __forceinline__ uint8_t multiMult(uint32_t k, uint32_t a, uint32_t b, uint32_t pa, uint32_t pb, uint32_t s) const {
        if (count(k, s) == 0) return (uint8_t)(diplMult(k, a, b) + ic(k, s));
        return (uint8_t)(sharedMult(k, s) - diplMult(k, pa, pb) + diplMult(k, a, b) + ic(k, s));
    }
};

// KmerStats (KmerStats.cpp:51-63) keeps Welford running means; only the means (count, fraction of non-zero,
// mean) are ever read on this path, so the kernel keeps plain sums and divides once when a value is consumed
// (two f64 divisions per addValue become one addition; the quotient differs from Welford's by rounding only).
// k-mer stats cache entry: kc_n = #values, kc_f[2i] = #non-zero values -> fraction, kc_f[2i+1] = sum -> mean
// (both finalised in place after a cache rebuild); allele stats entry: as_n = #values, as_f = sum.
__device__ __forceinline__ void kc_add(uint32_t &n, double &nonzero, double &sum, double v) {
    n++;
    nonzero += v != 0.0 ? 1.0 : 0.0;
    sum += v;
}

struct Tables {
    const double *genomic;  // [S][256][256]
    const double *noise;    // [S][256]
    // address of the table entry (no branch between the byte loads that produce (m, c) and the gather, so the gathers of
    // consecutive k-mers can be in flight together)
    __device__ __forceinline__ const double *entry(uint32_t s, uint8_t m, uint8_t c) const {
        return m == 0 ? noise + s * 256u + c : genomic + ((size_t)s * 256 + m) * 256 + c;
    }
    __device__ __forceinline__ double logProb(uint32_t s, uint8_t m, uint8_t c) const {  // CountDistribution.cpp:255-265
        return *entry(s, m, c);
    }
};

// Sum over the k-mer tile of one (sample, diplotype) cache entry, in subsample order (bit-identical to the sequential
// loop): the (multiplicity, count) bytes of eight k-mers are read first, then their eight table entries are gathered
// together, then added in order — eight L2 round trips overlap instead of queueing behind each other.
__device__ __forceinline__ const double *tile_term(const Cl &cl, const Tables &T, uint32_t s, uint32_t g, uint32_t a, uint32_t b, uint32_t i) {
    return T.entry(s, (uint8_t)(cl.tileDiplMult(i, a, b) + cl.tile_ic[i * 2 + g]), cl.tile_c[i * cl.S + s]);
}
__device__ __forceinline__ double tile_entry_sum(const Cl &cl, const Tables &T, uint32_t s, uint32_t a, uint32_t b, uint32_t n_sub) {
    const uint32_t g = cl.u->sample_gender[s];
    double acc = 0;
    uint32_t i = 0;
    for (; i + 8 <= n_sub; i += 8) {
        const double *p0 = tile_term(cl, T, s, g, a, b, i), *p1 = tile_term(cl, T, s, g, a, b, i + 1), *p2 = tile_term(cl, T, s, g, a, b, i + 2),
                     *p3 = tile_term(cl, T, s, g, a, b, i + 3), *p4 = tile_term(cl, T, s, g, a, b, i + 4), *p5 = tile_term(cl, T, s, g, a, b, i + 5),
                     *p6 = tile_term(cl, T, s, g, a, b, i + 6), *p7 = tile_term(cl, T, s, g, a, b, i + 7);
        const double v0 = *(p0), v1 = *(p1), v2 = *(p2), v3 = *(p3), v4 = *(p4), v5 = *(p5), v6 = *(p6), v7 = *(p7);
        acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
    }
    for (; i < n_sub; i++) acc += *(tile_term(cl, T, s, g, a, b, i));
    return acc;
}

// Row-major fill of ALL cache entries of a one-thread cluster with at most 4 live haplotypes (<= 10 diplotypes): the
// tile is walked once per sample and every k-mer updates all diplotype sums, so the gathers of one k-mer (up to 10,
// independent) are in flight together and each tile byte is read once instead of once per diplotype.  Every entry is
// still the sum of its terms in subsample order.  Used where the caches are cleared every iteration (lock-step modes).
// The common case — at most two live haplotypes, i.e. three diplotype sums — walks the tile FOUR k-mers at a time: the (up to) twelve table
// gathers of a round are issued before any of them is added, so a cluster-iteration waits for ceil(n_sub / 4) L2 round trips instead of n_sub
// (50 k cycles per SNV cluster-iteration were spent waiting for one gather after the other, profiles/r2_noise_chain_phases.txt).  The sums
// are still taken in subsample order.
__device__ __forceinline__ void cl_fill_cache_rows2(Cl &cl, const Tables &T, const uint8_t *ploidy, uint32_t hs0, uint32_t hs1, uint32_t n) {
    const uint32_t H = cl.H, n_sub = cl.misc[kNSub];
    for (uint32_t s = 0; s < cl.S; s++) {
        const uint8_t pl = ploidy[s];
        if (pl == 0) continue;
        const uint32_t g = cl.u->sample_gender[s];
        double a00 = 0, a01 = 0, a11 = 0;
        for (uint32_t i0 = 0; i0 < n_sub; i0 += 4) {
            const double *p00[4], *p01[4], *p11[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                const uint32_t i = i0 + j < n_sub ? i0 + j : n_sub - 1;   // the tail repeats the last row; its values are not added
                const uint8_t c = cl.tile_c[i * cl.S + s], base = cl.tile_ic[i * 2 + g];
                const uint8_t m0 = cl.tile_m[i * H + hs0], m1 = n > 1 ? cl.tile_m[i * H + hs1] : 0;
                if (pl == 2) {
                    p00[j] = T.entry(s, (uint8_t)(m0 + m0 + base), c);
                    p01[j] = T.entry(s, (uint8_t)(m0 + m1 + base), c);
                    p11[j] = T.entry(s, (uint8_t)(m1 + m1 + base), c);
                } else {
                    p00[j] = T.entry(s, (uint8_t)(m0 + base), c);
                    p11[j] = T.entry(s, (uint8_t)(m1 + base), c);
                    p01[j] = p00[j];
                }
            }
            double v00[4], v01[4], v11[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                v00[j] = *p00[j];
                v01[j] = n > 1 && pl == 2 ? *p01[j] : 0.0;
                v11[j] = n > 1 ? *p11[j] : 0.0;
            }
#pragma unroll
            for (uint32_t j = 0; j < 4; j++)
                if (i0 + j < n_sub) { a00 += v00[j]; a01 += v01[j]; a11 += v11[j]; }
        }
        const size_t cb = (size_t)s * cl.Dall;
        if (pl == 2) {
            cl.ucache[cb + cl.slot(hs0, hs0)] = a00;
            if (n > 1) { cl.ucache[cb + cl.slot(hs0, hs1)] = a01; cl.ucache[cb + cl.slot(hs1, hs1)] = a11; }
        } else {  // haploid entries live in the (h, "missing") slots
            cl.ucache[cb + cl.slot(hs0, H)] = a00;
            if (n > 1) cl.ucache[cb + cl.slot(hs1, H)] = a11;
        }
    }
}

__device__ __forceinline__ bool cl_fill_cache_rows(Cl &cl, const Tables &T, const uint8_t *ploidy) {
    const uint32_t H = cl.H, n_sub = cl.misc[kNSub];
    uint32_t hs0 = 0, hs1 = 0, hs2 = 0, hs3 = 0, n = 0;
    for (uint32_t h = 0; h < H; h++)
        if (cl.nz[h]) {
            if (n == 0) hs0 = h; else if (n == 1) hs1 = h; else if (n == 2) hs2 = h; else if (n == 3) hs3 = h; else return false;
            n++;
        }
    if (n <= 2 && n_sub > 0) { cl_fill_cache_rows2(cl, T, ploidy, hs0, hs1, n); return true; }
    for (uint32_t s = 0; s < cl.S; s++) {
        const uint8_t pl = ploidy[s];
        if (pl == 0) continue;
        const uint32_t g = cl.u->sample_gender[s];
        // pair slots: (0,0) (0,1) (0,2) (0,3) (1,1) (1,2) (1,3) (2,2) (2,3) (3,3)
        double a00 = 0, a01 = 0, a02 = 0, a03 = 0, a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
        for (uint32_t i = 0; i < n_sub; i++) {
            const uint8_t c = cl.tile_c[i * cl.S + s];
            const uint8_t base = cl.tile_ic[i * 2 + g];
            const uint8_t m0 = cl.tile_m[i * H + hs0];
            const uint8_t m1 = n > 1 ? cl.tile_m[i * H + hs1] : 0, m2 = n > 2 ? cl.tile_m[i * H + hs2] : 0, m3 = n > 3 ? cl.tile_m[i * H + hs3] : 0;
            if (pl == 2) {
                const double *p00 = T.entry(s, (uint8_t)(m0 + m0 + base), c), *p01 = T.entry(s, (uint8_t)(m0 + m1 + base), c),
                             *p02 = T.entry(s, (uint8_t)(m0 + m2 + base), c), *p03 = T.entry(s, (uint8_t)(m0 + m3 + base), c),
                             *p11 = T.entry(s, (uint8_t)(m1 + m1 + base), c), *p12 = T.entry(s, (uint8_t)(m1 + m2 + base), c),
                             *p13 = T.entry(s, (uint8_t)(m1 + m3 + base), c), *p22 = T.entry(s, (uint8_t)(m2 + m2 + base), c),
                             *p23 = T.entry(s, (uint8_t)(m2 + m3 + base), c), *p33 = T.entry(s, (uint8_t)(m3 + m3 + base), c);
                const double v00 = *(p00);
                double v01 = 0, v02 = 0, v03 = 0, v11 = 0, v12 = 0, v13 = 0, v22 = 0, v23 = 0, v33 = 0;
                if (n > 1) { v01 = *(p01); v11 = *(p11); }
                if (n > 2) { v02 = *(p02); v12 = *(p12); v22 = *(p22); }
                if (n > 3) { v03 = *(p03); v13 = *(p13); v23 = *(p23); v33 = *(p33); }
                a00 += v00; a01 += v01; a02 += v02; a03 += v03; a11 += v11; a12 += v12; a13 += v13; a22 += v22; a23 += v23; a33 += v33;
            } else {
                const double v0 = *(T.entry(s, (uint8_t)(m0 + base), c));
                double v1 = 0, v2 = 0, v3 = 0;
                if (n > 1) v1 = *(T.entry(s, (uint8_t)(m1 + base), c));
                if (n > 2) v2 = *(T.entry(s, (uint8_t)(m2 + base), c));
                if (n > 3) v3 = *(T.entry(s, (uint8_t)(m3 + base), c));
                a00 += v0; a11 += v1; a22 += v2; a33 += v3;
            }
        }
        const size_t cb = (size_t)s * cl.Dall;
        if (pl == 2) {
            cl.ucache[cb + cl.slot(hs0, hs0)] = a00;
            if (n > 1) { cl.ucache[cb + cl.slot(hs0, hs1)] = a01; cl.ucache[cb + cl.slot(hs1, hs1)] = a11; }
            if (n > 2) { cl.ucache[cb + cl.slot(hs0, hs2)] = a02; cl.ucache[cb + cl.slot(hs1, hs2)] = a12; cl.ucache[cb + cl.slot(hs2, hs2)] = a22; }
            if (n > 3) { cl.ucache[cb + cl.slot(hs0, hs3)] = a03; cl.ucache[cb + cl.slot(hs1, hs3)] = a13; cl.ucache[cb + cl.slot(hs2, hs3)] = a23; cl.ucache[cb + cl.slot(hs3, hs3)] = a33; }
        } else {  // haploid entries live in the (h, "missing") slots
            cl.ucache[cb + cl.slot(hs0, H)] = a00;
            if (n > 1) cl.ucache[cb + cl.slot(hs1, H)] = a11;
            if (n > 2) cl.ucache[cb + cl.slot(hs2, H)] = a22;
            if (n > 3) cl.ucache[cb + cl.slot(hs3, H)] = a33;
        }
    }
    return true;
}

// ---- VariantClusterGenotyper ctor: sparsity estimate + frequency reset ------------------------
__device__ __forceinline__ void cl_reset_frequencies(Cl &cl) {  // FrequencyDistribution.cpp:46-51,104-115
    const double f0 = 1 / static_cast<double>(cl.H);
    for (uint32_t h = 0; h < cl.H; h++) { cl.obs[h] = 0; cl.freq[h] = f0; cl.nz[h] = 1; }
}

__device__ __forceinline__ void cl_construct(Cl &cl, const btg_gibbs_opts &o, uint64_t group_index, uint32_t chain) {
    const uint32_t H = cl.H, K = cl.K, S = cl.S;
    const uint32_t *src = cl.u->uniq_idx + cl.u->cl_uniq_off[cl.c];
    for (uint32_t i = 0; i < cl.n_uniq; i++) cl.uniq[i] = src[i];
    for (uint32_t i = 0; i < cl.Dall * S; i++) cl.tally[i] = 0;
    for (uint32_t s = 0; s < S; s++) { cl.dipl[s] = 0xFFFFFFFFu; cl.stats_update[s] = 1; }
    for (uint32_t i = 0; i < S * 2 * cl.nvar; i++) { cl.kc_n[i] = 0; cl.kc_f[2 * i] = 0; cl.kc_f[2 * i + 1] = 0; }
    for (uint32_t i = 0; i < cl.n_alleles * S * 3; i++) { cl.as_n[i] = 0; cl.as_f[i] = 0; }
    for (int i = 0; i < 8; i++) cl.misc[i] = 0;
    cl.misc[kSimplexNobs] = 0xFFFFFFFFu;
    cl.misc[kNMultiSub] = 0;
    cl.misc[kUseMulti] = 0;
    for (uint32_t i = 0; i < cl.n_multi; i++) cl.multi[i] = cl.u->multi_idx[cl.u->cl_multi_off[cl.c] + i];
    // SparsityEstimator::estimateMinimumColumnCover (SparsityEstimator.cpp:41-90), stream kind 1
    Philox sp;
    sp.init(o.random_seed, group_index, cl.u->cluster_idx[cl.c], kRngSparsity, chain);
    uint32_t n_unc = 0;
    for (uint32_t k = 0; k < K; k++) { cl.uncovered[k] = cl.u->k_has_counts[cl.row0 + k]; n_unc += cl.uncovered[k]; }
    uint32_t cover = 0;
    while (n_unc > 0) {
        // column cover = sum of multiplicities over uncovered rows; cnt[] doubles as the scratch row
        uint32_t mx = 0, ties = 0;
        for (uint32_t h = 0; h < H; h++) {
            uint32_t col = 0;
            for (uint32_t k = 0; k < K; k++) if (cl.uncovered[k]) col += cl.m(k, h);
            cl.cnt[h] = col;
            if (col > mx) { mx = col; ties = 1; } else if (col == mx) ties++;
        }
        // DiscreteSampler with unit weights (DiscreteSampler.cpp:61-87): u * n against cum = 1..n
        const double x = sp.u01() * (double)ties;
        uint32_t idx = 0;
        if (ties > 1) while (idx + 1 < ties && !(x < (double)(idx + 1))) idx++;
        uint32_t pick = 0, seen = 0;
        for (uint32_t h = 0; h < H; h++) if (cl.cnt[h] == mx) { if (seen == idx) { pick = h; break; } seen++; }
        cover++;
        for (uint32_t k = 0; k < K; k++) if (cl.uncovered[k] && cl.m(k, pick)) { cl.uncovered[k] = 0; n_unc--; }
    }
    cl.misc[kCover] = cover;
    cl.misc[kSparse] = cover > 0;
    if (cover > 0) {  // SparseFrequencyDistribution ctor (FrequencyDistribution.cpp:97-103)
        const double sp_in = cover / static_cast<double>(H), cap = 1 - 2.220446049250313e-16 * 100;
        cl.fmisc[0] = sp_in < cap ? sp_in : cap;
    } else cl.fmisc[0] = 0;
    cl_reset_frequencies(cl);
}

// cl_construct by a whole warp (large clusters of the lock-step chain, where the slowest constructor holds every chain's first
// barrier: 18 ms per chain, profiles/r1_noise_chain_phases.txt): lanes take haplotypes for the column sums of the greedy cover and
// rows / array elements for everything else; the picks and the random draws are those of the sequential code.
__device__ __forceinline__ void cl_construct_warp(Cl &cl, const btg_gibbs_opts &o, uint64_t group_index, uint32_t chain, uint32_t lane) {
    const uint32_t H = cl.H, K = cl.K, S = cl.S, FULL = 0xFFFFFFFFu;
    const uint32_t *src = cl.u->uniq_idx + cl.u->cl_uniq_off[cl.c];
    for (uint32_t i = lane; i < cl.n_uniq; i += 32) cl.uniq[i] = src[i];
    for (uint32_t i = lane; i < cl.Dall * S; i += 32) cl.tally[i] = 0;
    for (uint32_t s = lane; s < S; s += 32) { cl.dipl[s] = 0xFFFFFFFFu; cl.stats_update[s] = 1; }
    for (uint32_t i = lane; i < S * 2 * cl.nvar; i += 32) { cl.kc_n[i] = 0; cl.kc_f[2 * i] = 0; cl.kc_f[2 * i + 1] = 0; }
    for (uint32_t i = lane; i < cl.n_alleles * S * 3; i += 32) { cl.as_n[i] = 0; cl.as_f[i] = 0; }
    if (lane < 8) cl.misc[lane] = 0;
    __syncwarp();
    if (lane == 0) { cl.misc[kSimplexNobs] = 0xFFFFFFFFu; cl.misc[kNMultiSub] = 0; cl.misc[kUseMulti] = 0; }
    for (uint32_t i = lane; i < cl.n_multi; i += 32) cl.multi[i] = cl.u->multi_idx[cl.u->cl_multi_off[cl.c] + i];
    // SparsityEstimator::estimateMinimumColumnCover (SparsityEstimator.cpp:41-90), stream kind 1 (lane 0 draws)
    Philox sp;
    sp.init(o.random_seed, group_index, cl.u->cluster_idx[cl.c], kRngSparsity, chain);
    uint32_t n_part = 0;
    for (uint32_t k = lane; k < K; k += 32) { const uint8_t un = cl.u->k_has_counts[cl.row0 + k]; cl.uncovered[k] = un; n_part += un; }
    uint32_t n_unc = __reduce_add_sync(FULL, n_part);
    __syncwarp();
    uint32_t cover = 0;
    while (n_unc > 0) {  // warp-uniform
        uint32_t mx = 0;
        for (uint32_t hb = 0; hb < H; hb += 32) {  // column cover of haplotype hb + lane over the uncovered rows
            const uint32_t h = hb + lane;
            uint32_t col = 0;
            if (h < H) {
                for (uint32_t k = 0; k < K; k++) if (cl.uncovered[k]) col += cl.m(k, h);
                cl.cnt[h] = col;
            }
            mx = max(mx, __reduce_max_sync(FULL, col));
        }
        __syncwarp();
        uint32_t ties = 0;
        for (uint32_t hb = 0; hb < H; hb += 32) ties += __popc(__ballot_sync(FULL, hb + lane < H && cl.cnt[hb + lane] == mx));
        // DiscreteSampler with unit weights (DiscreteSampler.cpp:61-87): u * n against cum = 1..n
        uint32_t idx = 0;
        if (lane == 0) {
            const double x = sp.u01() * (double)ties;
            if (ties > 1) while (idx + 1 < ties && !(x < (double)(idx + 1))) idx++;
        }
        idx = __shfl_sync(FULL, idx, 0);
        uint32_t pick = 0;
        bool found = false;
        for (uint32_t hb = 0; hb < H; hb += 32) {  // the idx-th haplotype (ascending) whose column cover is the maximum
            const uint32_t bal = __ballot_sync(FULL, hb + lane < H && cl.cnt[hb + lane] == mx);
            const uint32_t n = __popc(bal);
            if (!found) {
                if (idx < n) { pick = hb + __fns(bal, 0, idx + 1); found = true; }
                else idx -= n;
            }
        }
        cover++;
        uint32_t removed = 0;
        for (uint32_t k = lane; k < K; k += 32) if (cl.uncovered[k] && cl.m(k, pick)) { cl.uncovered[k] = 0; removed++; }
        n_unc -= __reduce_add_sync(FULL, removed);
        __syncwarp();
    }
    if (lane == 0) {
        cl.misc[kCover] = cover;
        cl.misc[kSparse] = cover > 0;
        if (cover > 0) {  // SparseFrequencyDistribution ctor (FrequencyDistribution.cpp:97-103)
            const double sp_in = cover / static_cast<double>(H), cap = 1 - 2.220446049250313e-16 * 100;
            cl.fmisc[0] = sp_in < cap ? sp_in : cap;
        } else cl.fmisc[0] = 0;
        cl_reset_frequencies(cl);
    }
    __syncwarp();
}

// VariantClusterHaplotypes::isMaxHaplotypeVariantKmer (VariantClusterHaplotypes.cpp:159-178)
__device__ __forceinline__ bool cl_is_max_hap_var_kmer(Cl &cl, uint32_t k, uint32_t max_kmers) {
    bool is_max = true;
    const DevUnit &u = *cl.u;
    for (uint64_t e = u.kmer_vh_off[cl.row0 + k]; e < u.kmer_vh_off[cl.row0 + k + 1]; e++) {
        const uint32_t v = u.vh_var[e];
        const uint8_t *bits = u.vh_bits + u.vh_bits_off[e];
        for (uint32_t h = 0; h < cl.H; h++)
            if (bits[h] && cl.cnt[(size_t)h * cl.nvar + v] < max_kmers) { cl.cnt[(size_t)h * cl.nvar + v]++; is_max = false; }
    }
    return is_max;
}

// VariantClusterGenotyper::reset + VariantClusterHaplotypes::sampleKmerSubset (…Genotyper.cpp:113-129, …Haplotypes.cpp:110-157)
template <bool MC = false, bool TILE = false, bool FRESH = false>
__device__ __forceinline__ void cl_reset(Cl &cl, const btg_gibbs_opts &o, Philox &prng) {
    const double rate = (double)o.kmer_subsampling_rate;
    if constexpr (FRESH) {  // chains are independent in the default mode: every chain shuffles the original order
        const uint32_t *src = cl.u->uniq_idx + cl.u->cl_uniq_off[cl.c];
        for (uint32_t i = 0; i < cl.n_uniq; i++) cl.uniq[i] = src[i];
    }
    for (uint32_t i = 0; i < cl.H * cl.nvar; i++) cl.cnt[i] = 0;
    for (uint32_t i = cl.n_uniq; i > 1; i--) {  // Fisher-Yates from the back
        const uint32_t j = prng.uniform_int(i);
        const uint32_t t = cl.uniq[i - 1]; cl.uniq[i - 1] = cl.uniq[j]; cl.uniq[j] = t;
    }
    uint32_t n_sub = 0;
    for (uint32_t i = 0; i < cl.n_uniq; i++) {
        const uint32_t k = cl.uniq[i];
        if (prng.u01() < rate)
            if (!cl_is_max_hap_var_kmer(cl, k, o.max_haplotype_variant_kmers)) cl.uniq_sub[n_sub++] = k;
    }
    cl.misc[kNSub] = n_sub;
    if constexpr (TILE) {
        for (uint32_t i = 0; i < n_sub; i++) {
            const uint32_t k = cl.uniq_sub[i];
            const bool has = cl.u->k_has_counts[cl.row0 + k];
            for (uint32_t h = 0; h < cl.H; h++) cl.tile_m[i * cl.H + h] = cl.m(k, h);
            for (uint32_t s = 0; s < cl.S; s++) cl.tile_c[i * cl.S + s] = has ? cl.u->k_counts[(cl.row0 + k) * cl.S + s] : 0;
            cl.tile_ic[i * 2] = has ? cl.u->k_ic[(cl.row0 + k) * 2] : 0;
            cl.tile_ic[i * 2 + 1] = has ? cl.u->k_ic[(cl.row0 + k) * 2 + 1] : 0;
        }
    }
    for (uint32_t s = 0; s < cl.S; s++) cl.stats_update[s] = 1;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (cl.has_cache)
        for (uint32_t i = 0; i < cl.S * cl.Dall; i++) cl.ucache[i] = nan;  // clear the per-sample diplotype caches
    if constexpr (MC) {
        if (cl.n_multi) {
            for (uint32_t i = cl.n_multi; i > 1; i--) {
                const uint32_t j = prng.uniform_int(i);
                const uint32_t t = cl.multi[i - 1]; cl.multi[i - 1] = cl.multi[j]; cl.multi[j] = t;
            }
            uint32_t n_msub = 0;
            for (uint32_t i = 0; i < cl.n_multi; i++) {
                const uint32_t k = cl.multi[i];
                if (prng.u01() < rate)
                    if (!cl_is_max_hap_var_kmer(cl, k, o.max_haplotype_variant_kmers)) cl.multi_sub[n_msub++] = k;
            }
            cl.misc[kNMultiSub] = n_msub;
            for (uint32_t i = 0; i < n_msub * cl.S; i++) cl.sample_multi[i] = 0;
            for (uint32_t i = 0; i < cl.S * cl.Dall; i++) cl.mcache[i] = nan;
        }
        cl.misc[kUseMulti] = 0;
    }
    cl_reset_frequencies(cl);
}

// VariantClusterGenotyper::updateMulticlusterDiplotypeLogProb (…Genotyper.cpp:569-595): cached terms of the k-mers whose
// shared multiplicity another cluster of the group has changed are replaced in place (NaN = diplotype not cached)
__device__ __forceinline__ void cl_update_multi_log_prob(Cl &cl, const Tables &T, uint32_t s) {
    const uint32_t n_msub = cl.misc[kNMultiSub], H = cl.H;
    const uint32_t pa = cl.dipl[s] & 0xFFFFu, pb = cl.dipl[s] >> 16;
    for (uint32_t sub = 0; sub < n_msub; sub++) {
        const uint32_t k = cl.multi_sub[sub];
        const uint8_t cnt = cl.count(k, s);
        const uint8_t seen = cl.sample_multi[sub * cl.S + s];
        if (!(cnt > 0 && cl.sharedMult(k, s) != seen)) continue;  // isMulticlusterKmerUpdated (…Haplotypes.cpp:180-195)
        const uint8_t base_prev = (uint8_t)(seen - cl.diplMult(k, pa, pb) + cl.ic(k, s));
        for (uint32_t b = 0; b <= H; b++) {
            for (uint32_t a = 0; a <= b && a < H; a++) {
                const size_t ci = (size_t)s * cl.Dall + cl.slot(a, b);
                double v = cl.mcache[ci];
                if (v != v) continue;
                const uint32_t bb = b == H ? NONE : b;
                v -= T.logProb(s, (uint8_t)(base_prev + cl.diplMult(k, a, bb)), cnt);  // getPreviousMulticlusterKmerMultiplicity
                v += T.logProb(s, cl.multiMult(k, a, bb, pa, pb, s), cnt);
                cl.mcache[ci] = v;
            }
        }
    }
}

// VariantClusterHaplotypes::updateMulticlusterKmerMultiplicities (VariantClusterHaplotypes.cpp:197-233)
__device__ __forceinline__ void cl_update_multi_multiplicities(Cl &cl, uint32_t s, uint32_t prev) {
    const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16;
    if (cl.dipl[s] != prev) {
        cl.stats_update[s] = 1;
        const uint32_t pa = prev & 0xFFFFu, pb = prev >> 16;
        for (uint32_t i = 0; i < cl.n_multi; i++) {
            const uint32_t k = cl.multi[i];
            const uint8_t cur = cl.diplMult(k, da, db), old = cl.diplMult(k, pa, pb);
            if (cur != old) { uint8_t &m = cl.sharedMult(k, s); m = (uint8_t)(m - old + cur); }
        }
    }
    const uint32_t n_msub = cl.misc[kNMultiSub];
    for (uint32_t sub = 0; sub < n_msub; sub++) {
        const uint32_t k = cl.multi_sub[sub];
        const uint8_t m = cl.sharedMult(k, s);
        if (cl.diplMult(k, da, db) > 0 && cl.count(k, s) > 0 && m != cl.sample_multi[sub * cl.S + s]) cl.stats_update[s] = 1;
        cl.sample_multi[sub * cl.S + s] = m;
    }
}

// VariantClusterGenotyper::calcDiplotypeLogProb (VariantClusterGenotyper.cpp:597-666)
template <bool MC = false, bool TILE = false>
__device__ __forceinline__ double cl_dipl_log_prob(Cl &cl, const Tables &T, uint32_t s, uint32_t a, uint32_t b) {
    double lp = 0;  // logf[] = log(freq[]) of this iteration (cl_sample_diplotypes)
    if (b == NONE) lp += cl.logf[a];
    else if (a == b) lp += 2 * cl.logf[a];
    else lp += 0.6931471805599453 + cl.logf[a] + cl.logf[b];
    const size_t ci = (size_t)s * cl.Dall + cl.slot(a, b == NONE ? cl.H : b);
    double acc = cl.has_cache ? cl.ucache[ci] : __longlong_as_double(0x7ff8000000000000LL);
    if (acc != acc) {  // not cached yet (or a cluster without dense caches: the sum is the same whenever it is taken)
        acc = 0;
        const uint32_t n_sub = cl.misc[kNSub];
        if constexpr (TILE) {
            acc = tile_entry_sum(cl, T, s, a, b, n_sub);
        } else {
            for (uint32_t i = 0; i < n_sub; i++) {
                const uint32_t k = cl.uniq_sub[i];
                acc += T.logProb(s, (uint8_t)(cl.diplMult(k, a, b) + cl.ic(k, s)), cl.count(k, s));
            }
        }
        if (cl.has_cache) cl.ucache[ci] = acc;
    }
    lp += acc;
    if constexpr (MC) {
        if (cl.misc[kUseMulti]) {
            double macc = cl.mcache[ci];
            if (macc != macc) {
                macc = 0;
                const uint32_t n_msub = cl.misc[kNMultiSub], pa = cl.dipl[s] & 0xFFFFu, pb = cl.dipl[s] >> 16;
                for (uint32_t i = 0; i < n_msub; i++) {
                    const uint32_t k = cl.multi_sub[i];
                    macc += T.logProb(s, cl.multiMult(k, a, b, pa, pb, s), cl.count(k, s));
                }
                cl.mcache[ci] = macc;
            }
            lp += macc;
        }
    }
    return lp;
}

__device__ __forceinline__ void cl_increment(Cl &cl, uint32_t h) {  // HaplotypeFrequencyDistribution.cpp:114-126
    if (h == NONE) { cl.misc[kNumMissing]++; return; }
    cl.misc[kNumHap]++;
    cl.obs[h]++;
}

// VariantClusterGenotyper::sampleDiplotype (VariantClusterGenotyper.cpp:707-755) + LogDiscreteSampler (DiscreteSampler.cpp:106-126)
// The uniform of the draw comes from the block that (sampleDiplotypes call, sample) owns in the genotyper's stream (gibbs_rng.cuh).
// WIDE: one lane per sample — the cumulative log-probs live in the sample's own row and the haplotype counts are added up by
// the caller (clw_sample_diplotypes).  Clusters without dense caches (wide layout, large H) keep no cumulative row either: the
// outcome is found by walking the enumeration a second time (the same sums in the same order give the same running values; the
// cumulative log-probs never decrease, so "first value above x" is what upper_bound returns).
template <bool MC = false, bool TILE = false, bool WIDE = false>
__device__ __forceinline__ void cl_sample_diplotype(Cl &cl, const Tables &T, uint32_t s, uint8_t ploidy, const Philox &prng, uint64_t nzm) {
    uint32_t n = 0;
    double run = 0;
    const uint32_t H = cl.H;
    const bool keep = cl.has_cache;
    LaneArr<double> cum = cl.cum + (size_t)s * cl.cum_stride;
    // next haplotype >= from with a non-zero frequency (H if none).  For H <= 64 the flags arrive as the bit mask nzm, so the
    // pair enumeration visits only live pairs instead of testing H^2/2 flags in the arena (a cluster with 16 haplotypes of
    // which 4 are live: 10 steps instead of 136 loads); the order of enumeration is unchanged.
    const bool use_mask = H <= 64;
    auto next = [&](uint32_t from) -> uint32_t {
        if (use_mask) {
            const uint64_t m = from < 64 ? nzm >> from : 0;
            return m ? from + (uint32_t)__ffsll((long long)m) - 1 : H;
        }
        while (from < H && !cl.nz[from]) from++;
        return from;
    };
    if (ploidy == 2) {
        for (uint32_t a = next(0); a < H; a = next(a + 1)) {
            for (uint32_t b = a; b < H; b = next(b + 1)) {
                const double lp = cl_dipl_log_prob<MC, TILE>(cl, T, s, a, b);
                run = n == 0 ? lp : logAddition(lp, run);
                if (keep) cum[n] = run;
                n++;
            }
        }
    } else if (ploidy == 1) {
        for (uint32_t a = next(0); a < H; a = next(a + 1)) {
            const double lp = cl_dipl_log_prob<MC, TILE>(cl, T, s, a, NONE);
            run = n == 0 ? lp : logAddition(lp, run);
            if (keep) cum[n] = run;
            n++;
        }
    } else {
        n = 1;
    }
    const double x = m_log(prng.u01_draw(s)) + run;
    uint32_t idx = 0;
    uint32_t da = NONE, db = NONE;
    if (n > 1 && keep) {  // upper_bound
        uint32_t lo = 0, hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (x < cum[mid]) hi = mid; else lo = mid + 1; }
        idx = lo < n ? lo : n - 1;
    }
    if (n > 1 && !keep) {  // second walk
        double r2 = 0;
        uint32_t i = 0;
        idx = n - 1;
        if (ploidy == 2) {
            for (uint32_t a = next(0); a < H && i < n; a = next(a + 1))
                for (uint32_t b = a; b < H; b = next(b + 1)) {
                    const double lp = cl_dipl_log_prob<MC, TILE>(cl, T, s, a, b);
                    r2 = i == 0 ? lp : logAddition(lp, r2);
                    if (x < r2) { idx = i; i = n; break; }
                    i++;
                }
        } else {
            for (uint32_t a = next(0); a < H; a = next(a + 1)) {
                const double lp = cl_dipl_log_prob<MC, TILE>(cl, T, s, a, NONE);
                r2 = i == 0 ? lp : logAddition(lp, r2);
                if (x < r2) { idx = i; break; }
                i++;
            }
        }
    }
    // map the outcome index back to (a, b) in enumeration order
    if (ploidy == 2) {
        uint32_t i = 0;
        for (uint32_t a = next(0); a < H && da == NONE; a = next(a + 1)) {
            for (uint32_t b = a; b < H; b = next(b + 1)) {
                if (i == idx) { da = a; db = b; break; }
                i++;
            }
        }
    } else if (ploidy == 1) {
        uint32_t i = 0;
        for (uint32_t a = next(0); a < H; a = next(a + 1)) { if (i == idx) { da = a; break; } i++; }
    }
    cl.dipl[s] = (da & 0xFFFFu) | (db << 16);
    if constexpr (!WIDE) {
        cl_increment(cl, da);
        cl_increment(cl, db);
    }
}

// VariantClusterHaplotypes::updateAlleleKmerStats (VariantClusterHaplotypes.cpp:235-372), single-cluster groups
__device__ __forceinline__ void cl_add_haplotype_stats(Cl &cl, uint32_t s, uint32_t which, uint32_t h) {  // addHaplotypeKmerStats
    uint32_t last = NONE;
    for (uint32_t v = 0; v < cl.nvar; v++) {
        const uint16_t a = cl.hapAllele(h, v);
        uint32_t srcv;
        if (cl.isMissing(v, a)) srcv = last; else { srcv = v; last = v; }
        const uint32_t ci = (s * 2 + which) * cl.nvar + srcv;
        const uint32_t n = cl.kc_n[ci];
        const uint32_t ai = cl.alleleBase(v, s) + a;
        // AlleleKmerStats::addKmerStats (KmerStats.cpp:115-122): count, fraction (if any), mean (if any)
        cl.as_n[ai * 3 + 0]++; cl.as_f[ai * 3 + 0] += (double)n;
        if (n > 0) {
            cl.as_n[ai * 3 + 1]++; cl.as_f[ai * 3 + 1] += cl.kc_f[2 * ci];      // getFraction()
            cl.as_n[ai * 3 + 2]++; cl.as_f[ai * 3 + 2] += cl.kc_f[2 * ci + 1];  // getMean()
        }
    }
}

// updateKmerStatsCache (…Haplotypes.cpp:302-333)
__device__ __forceinline__ void cl_stats_cache_add(Cl &cl, uint32_t k, uint32_t s, uint32_t da, uint32_t db, uint8_t mult) {
    const DevUnit &u = *cl.u;
    const double kc = u.k_has_counts[cl.row0 + k] ? cl.count(k, s) / static_cast<double>(mult) : 0.0;
    for (uint64_t e = u.kmer_vh_off[cl.row0 + k]; e < u.kmer_vh_off[cl.row0 + k + 1]; e++) {
        const uint32_t v = u.vh_var[e];
        const uint8_t *bits = u.vh_bits + u.vh_bits_off[e];
        if (bits[da]) { const uint32_t ci = (s * 2 + 0) * cl.nvar + v; kc_add(cl.kc_n[ci], cl.kc_f[2 * ci], cl.kc_f[2 * ci + 1], kc); }
        if (db != NONE && bits[db]) { const uint32_t ci = (s * 2 + 1) * cl.nvar + v; kc_add(cl.kc_n[ci], cl.kc_f[2 * ci], cl.kc_f[2 * ci + 1], kc); }
    }
}

template <bool MC = false>
__device__ __forceinline__ void cl_update_allele_stats_sample(Cl &cl, uint32_t s) {
    {
        const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16;
        if (cl.stats_update[s]) {
            cl.stats_update[s] = 0;
            for (uint32_t i = s * 2 * cl.nvar; i < (s + 1) * 2 * cl.nvar; i++) { cl.kc_n[i] = 0; cl.kc_f[2 * i] = 0; cl.kc_f[2 * i + 1] = 0; }
            if (da != NONE) {
                const uint32_t n_sub = cl.misc[kNSub];
                for (uint32_t i = 0; i < n_sub; i++) {
                    const uint32_t k = cl.uniq_sub[i];
                    const uint8_t dm = cl.diplMult(k, da, db);
                    if (dm == 0) continue;
                    cl_stats_cache_add(cl, k, s, da, db, (uint8_t)(dm + cl.ic(k, s)));
                }
                if constexpr (MC) {
                    const uint32_t n_msub = cl.misc[kNMultiSub];
                    for (uint32_t i = 0; i < n_msub; i++) {
                        const uint32_t k = cl.multi_sub[i];
                        if (cl.diplMult(k, da, db) == 0) continue;
                        cl_stats_cache_add(cl, k, s, da, db, cl.multiMult(k, da, db, da, db, s));
                    }
                }
            }
            // finalise: (#non-zero, sum) -> (fraction, mean)
            for (uint32_t i = s * 2 * cl.nvar; i < (s + 1) * 2 * cl.nvar; i++) {
                const uint32_t n = cl.kc_n[i];
                if (n) { cl.kc_f[2 * i] = cl.kc_f[2 * i] / n; cl.kc_f[2 * i + 1] = cl.kc_f[2 * i + 1] / n; }
            }
        }
        if (da != NONE) cl_add_haplotype_stats(cl, s, 0, da);
        if (db != NONE) cl_add_haplotype_stats(cl, s, 1, db);
    }
}
template <bool MC = false>
__device__ __forceinline__ void cl_update_allele_stats(Cl &cl) {
    for (uint32_t s = 0; s < cl.S; s++) cl_update_allele_stats_sample<MC>(cl, s);
}

// VariantClusterGenotyper::sampleDiplotypes (VariantClusterGenotyper.cpp:668-705)
template <bool MC = false, bool TILE = false>
__device__ __forceinline__ void cl_sample_diplotypes(Cl &cl, const Tables &T, const uint8_t *ploidy, bool collect, Philox &prng) {
    uint64_t nzm = 0;  // non-zero flags of the first 64 haplotypes as a bit mask
    for (uint32_t h = 0; h < cl.H; h++)
        if (cl.nz[h]) { cl.logf[h] = m_log(cl.freq[h]); if (h < 64) nzm |= 1ull << h; }  // one log per haplotype per iteration
    for (uint32_t s = 0; s < cl.S; s++) {
        const uint32_t prev = cl.dipl[s];
        if constexpr (MC) { if (cl.misc[kUseMulti]) cl_update_multi_log_prob(cl, T, s); }
        cl_sample_diplotype<MC, TILE>(cl, T, s, ploidy[s], prng, nzm);
        if constexpr (MC) cl_update_multi_multiplicities(cl, s, prev);
        else if (cl.dipl[s] != prev) cl.stats_update[s] = 1;  // …Haplotypes.cpp:199-201
        if (collect) {
            const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16;
            cl.tally[(size_t)cl.slot(da == NONE ? cl.H : da, db == NONE ? cl.H : db) * cl.S + s]++;
        }
    }
    prng.t_draw++;
    if (collect) cl_update_allele_stats<MC>(cl);
    if constexpr (MC) cl.misc[kUseMulti] = cl.misc[kNMultiSub] > 0;
}

// SparseFrequencyDistribution::updateCachedSimplexProbVector (FrequencyDistribution.cpp:143-196);
// lgamma of the integer arguments comes from a table shared by all clusters
__device__ __forceinline__ uint32_t cl_simplex_vector(Cl &cl, LaneArr<double> out, uint32_t n_obs, uint32_t plus) {
    const double *lg = cl.u->lgamma_int;
    const uint32_t H = cl.H;
    const double sparsity = cl.fmisc[0];
    const double ls = m_log(sparsity), l1s = m_log(1 - sparsity);
    double prob_z = plus * ls + (H - plus) * l1s;
    double prob_t = lg[plus] - lg[n_obs + plus];
    double row_sum = 0 + prob_z + prob_t;
    uint32_t len = 0;
    out[len++] = row_sum;
    for (uint32_t j = plus + 1; j < H + 1; j++) {
        const double cardinal = lg[H - plus + 1] - (lg[j - plus + 1] + lg[H - j + 1]);
        prob_z = j * ls + (H - j) * l1s;
        prob_t = lg[j] - lg[n_obs + j];
        const double prob_eq = cardinal + prob_z + prob_t;
        row_sum += m_log(1 + m_exp(prob_eq - row_sum));
        out[len++] = row_sum;
        if (doubleCompare(out[len - 1], out[len - 2])) break;
    }
    for (uint32_t i = 0; i < len; i++) out[i] = m_exp(out[i] - row_sum);
    return len;
}

// VariantClusterGenotyper::sampleHaplotypeFrequencies (…Genotyper.cpp:781-785) ->
// (Sparse)FrequencyDistribution::sampleFrequencies (FrequencyDistribution.cpp:75-94,209-304)
__device__ __forceinline__ void cl_sample_frequencies(Cl &cl, Philox &fr) {
    const uint32_t H = cl.H;
    const uint32_t n_obs = cl.misc[kNumHap];
    if (n_obs > 0) {
        if (!cl.misc[kSparse]) {
            double norm = 0;
            for (uint32_t h = 0; h < H; h++) { const double f = fr.gamma(cl.obs[h] + 1.0); cl.freq[h] = f; norm += f; cl.obs[h] = 0; }
            for (uint32_t h = 0; h < H; h++) cl.freq[h] = m_div(cl.freq[h], norm);
        } else {
            uint32_t plus = 0;
            for (uint32_t h = 0; h < H; h++) plus += cl.obs[h] > 0;
            // cached_simplex_prob_vectors (FrequencyDistribution.cpp:211-229): one vector per (n_obs, plus).  Small
            // clusters keep a row per plus-count for the current n_obs; large ones only the most recent key.
            LaneArr<double> vec = cl.simplex;
            uint32_t len;
            if (cl.has_simplex_tab) {
                if (cl.misc[kSimplexNobs] != n_obs) {
                    for (uint32_t p = 0; p < H; p++) cl.simplex_tab[p * (H + 1)] = 0;
                    cl.misc[kSimplexNobs] = n_obs;
                }
                LaneArr<double> row = cl.simplex_tab + (size_t)(plus - 1) * (H + 1);
                len = (uint32_t)row[0];
                if (len == 0) { len = cl_simplex_vector(cl, row + 1, n_obs, plus); row[0] = (double)len; }
                vec = row + 1;
            } else {
                if (cl.misc[kSimplexNobs] != n_obs || cl.misc[kSimplexPlus] != plus) {
                    cl.misc[kSimplexLen] = cl_simplex_vector(cl, cl.simplex, n_obs, plus);
                    cl.misc[kSimplexNobs] = n_obs;
                    cl.misc[kSimplexPlus] = plus;
                }
                len = cl.misc[kSimplexLen];
            }
            const double uu = fr.u01();
            uint32_t ub = 0;
            while (ub < len && !(uu < vec[ub])) ub++;  // upper_bound
            const uint32_t simplex_size = ub + plus;
            double norm = 0;
            // observed haplotypes, ascending index; nz[] marks membership of the (growing) plus set
            for (uint32_t h = 0; h < H; h++) {
                if (cl.obs[h] > 0) { const double f = fr.gamma(cl.obs[h] + 1.0); cl.freq[h] = f; norm += f; cl.nz[h] = 1; }
                else cl.nz[h] = 0;
            }
            uint32_t n_zero = H - plus;
            while (plus < simplex_size) {
                const uint32_t posn = fr.uniform_int(n_zero);
                uint32_t seen = 0, pick = 0;
                for (uint32_t h = 0; h < H; h++) if (!cl.nz[h]) { if (seen == posn) { pick = h; break; } seen++; }
                const double f = fr.gamma(1.0);
                cl.freq[pick] = f; norm += f; cl.nz[pick] = 1;
                plus++; n_zero--;
            }
            for (uint32_t h = 0; h < H; h++) {
                if (cl.nz[h]) cl.freq[h] = m_div(cl.freq[h], norm); else cl.freq[h] = 0;
                cl.obs[h] = 0;
            }
        }
    }
    cl.misc[kNumHap] = 0;
    cl.misc[kNumMissing] = 0;
}

// VariantClusterGenotyper::getGenotypes & co. (VariantClusterGenotyper.cpp:208-567)
struct ResultView {
    const uint64_t *allele_off, *geno_off, *valt_off;
    uint16_t *gt; uint32_t *gq; float *gpp, *app, *nak, *fak, *mac; uint16_t *saf; uint8_t *ploidy;
    uint32_t *an, *ac; float *af, *acp; uint8_t *anc; uint16_t *hc;
};

__device__ __forceinline__ void cl_summarise(Cl &cl, const btg_gibbs_opts &o, const uint8_t *ploidy, const ResultView &R, uint32_t v_first = 0, uint32_t v_step = 1) {
    const uint32_t S = cl.S, H = cl.H;
    for (uint32_t v = v_first; v < cl.nvar; v += v_step) {
        const uint64_t gv = cl.var0 + v;
        const uint32_t nA = cl.nalleles(v), nG = nA * (nA + 1) / 2;
        R.hc[gv] = (uint16_t)H;
        uint8_t *anc = R.anc + R.valt_off[gv];
        uint32_t *ac = R.ac + R.valt_off[gv];
        float *acp = R.acp + R.valt_off[gv], *af = R.af + R.valt_off[gv];
        for (uint32_t a = 0; a < nA; a++) { anc[a] = 1; ac[a] = 0; acp[a] = 0; }
        for (uint32_t h = 0; h < H; h++) anc[cl.hapAllele(h, v)] = 0;  // getNonCoveredAlleles (…Genotyper.cpp:221-247)
        if (cl.u->var_dep[gv]) anc[nA - 1] = 0;
        uint32_t total_count = 0;
        for (uint32_t s = 0; s < S; s++) {
            float *gpp = R.gpp + R.geno_off[gv] + (size_t)s * nG;
            const uint64_t ab = R.allele_off[gv] + (size_t)s * nA;
            float *app = R.app + ab, *nak = R.nak + ab, *fak = R.fak + ab, *mac = R.mac + ab;
            uint16_t *saf = R.saf + ab;
            const uint8_t pl = ploidy[s];
            R.ploidy[gv * S + s] = pl;
            const uint32_t n_geno = pl == 2 ? nG : (pl == 1 ? nA : 0), n_all = pl == 0 ? 0 : nA;
            for (uint32_t i = 0; i < nG; i++) gpp[i] = 0;
            for (uint32_t i = 0; i < nA; i++) { app[i] = 0; saf[i] = 0; }
            uint32_t n_it = 0, best_n = 0, best_a = NONE, best_b = NONE;
            float best_p = 0;
            for (uint32_t b = 0; b <= H; b++) {
                for (uint32_t a = 0; a <= b; a++) {
                    const uint32_t cnt = cl.tally[(size_t)cl.slot(a, b) * S + s];
                    if (cnt == 0) continue;
                    uint32_t ga = NONE, gb = NONE, gi = 0;
                    if (pl == 2) {
                        ga = a == H ? nA - 1 : cl.hapAllele(a, v);  // haplotypeToAlleleIndex (…Genotyper.cpp:208-219)
                        gb = b == H ? nA - 1 : cl.hapAllele(b, v);
                        if (ga > gb) { const uint32_t t = ga; ga = gb; gb = t; }
                        gi = gb * (gb + 1) / 2 + ga;
                        gpp[gi] += cnt;
                        app[ga] += cnt;
                        if (ga != gb) app[gb] += cnt;
                    } else if (pl == 1) {
                        ga = a == H ? nA - 1 : cl.hapAllele(a, v);
                        gi = ga;
                        gpp[gi] += cnt;
                        app[gi] += cnt;
                    }
                    n_it += cnt;
                    if (pl != 0) {
                        if (floatCompare(best_p, gpp[gi])) best_n++;
                        else if (best_p < gpp[gi]) { best_n = 1; best_a = ga; best_b = gb; best_p = gpp[gi]; }
                    }
                }
            }
            best_p /= n_it;
            for (uint32_t i = 0; i < n_geno; i++) gpp[i] /= n_it;
            for (uint32_t i = 0; i < n_all; i++) app[i] /= n_it;
            const uint32_t ai0 = cl.alleleBase(v, s);
            for (uint32_t a = 0; a < nA; a++) {
                const uint32_t ai = ai0 + a;
                nak[a] = cl.as_n[ai * 3 + 0] ? (float)(cl.as_f[ai * 3 + 0] / cl.as_n[ai * 3 + 0]) : -1.f;
                fak[a] = cl.as_n[ai * 3 + 1] ? (float)(cl.as_f[ai * 3 + 1] / cl.as_n[ai * 3 + 1]) : -1.f;
                mac[a] = cl.as_n[ai * 3 + 2] ? (float)(cl.as_f[ai * 3 + 2] / cl.as_n[ai * 3 + 2]) : -1.f;
            }
            for (uint32_t a = 0; a < n_all; a++) {
                if (!floatCompare(app[a], 0)) {
                    if (floatLess(nak[a], o.min_number_of_kmers)) saf[a] += 1;
                    if (!floatCompare(nak[a], 0))
                        if (floatLess(fak[a], o.min_fraction_observed_kmers[s])) saf[a] += 2;
                }
            }
            uint32_t gq;
            if (floatCompare(best_p, 1)) gq = 99;
            else if (floatCompare(best_p, 0)) gq = 0;
            else gq = (uint32_t)(-10 * log10f(1 - best_p));
            R.gq[gv * S + s] = gq;
            uint16_t *gt = R.gt + (gv * S + s) * 2;
            gt[0] = NONE;
            gt[1] = pl == 2 ? NONE : 0xFFFE;
            if (pl == 2) {
                if (best_n == 1 && !floatLess(best_p, o.min_genotype_posterior))
                    if (saf[best_a] == 0 && saf[best_b] == 0) { gt[0] = (uint16_t)best_a; gt[1] = (uint16_t)best_b; }
            } else if (pl == 1) {
                if (best_n == 1 && !floatLess(best_p, o.min_genotype_posterior))
                    if (saf[best_a] == 0) gt[0] = (uint16_t)best_a;
            }
            for (int i = 0; i < 2; i++)  // getGenotypeVariantStats (…Genotyper.cpp:470-526)
                if (gt[i] < 0xFFFE) { total_count++; if (gt[i] > 0) ac[gt[i]]++; }
            for (uint32_t a = 0; a < n_all; a++)
                if (saf[a] == 0) acp[a] = fmaxf(acp[a], app[a]);
        }
        R.an[gv] = total_count;
        for (uint32_t a = 0; a < nA; a++) af[a] = total_count > 0 ? ac[a] / static_cast<float>(total_count) : 0.f;
    }
}

// ---- groups with nested clusters ------------------------------------------------------------------
// VariantClusterGenotyper::updateNestedVariantClusterInfo / updateNestedPloidy / addNestedKmerStats (…Genotyper.cpp:140-206):
// `cl` is the parent that has just been sampled; the child's incoming info (slot nt) already holds a copy of the parent's own
__device__ __forceinline__ void cl_update_nested_info(Cl &cl, uint32_t nt, uint32_t child_cluster_idx, uint32_t s_begin = 0, uint32_t s_end = 0xFFFFFFFFu) {
    const DevUnit &u = *cl.u;
    uint64_t dep = u.cl_dep_off[cl.c];
    while (dep < u.cl_dep_off[cl.c + 1] && u.dep_cluster[dep] != child_cluster_idx) dep++;
    for (uint32_t s = s_begin; s < cl.S && s < s_end; s++) {
        for (uint32_t which = 0; which < 2; which++) {
            const uint32_t h = which == 0 ? (cl.dipl[s] & 0xFFFFu) : (cl.dipl[s] >> 16);
            if (h == NONE) continue;
            bool contains = false;  // haplotype runs through the child cluster's position
            for (uint64_t e = u.hap_nested_off[u.hap_start[cl.c] + h]; e < u.hap_nested_off[u.hap_start[cl.c] + h + 1]; e++)
                if (u.hap_nested[e] == child_cluster_idx) { contains = true; break; }
            if (contains) continue;
            uint8_t &pl = u.nest_pl[(size_t)nt * cl.S + s];
            pl = pl == 2 ? 1 : 0;
            uint32_t v = NONE;
            if (dep < u.cl_dep_off[cl.c + 1])
                for (uint64_t e = u.dep_var_off[dep]; e < u.dep_var_off[dep + 1]; e++) {
                    const uint32_t nv = u.dep_var[e];
                    if (!cl.isMissing(nv, cl.hapAllele(h, nv))) { v = nv; break; }
                }
            uint8_t &k = u.nest_k[(size_t)nt * cl.S + s];
            if (v == NONE || k >= 2) continue;  // the reference asserts both
            const uint32_t ci = (s * 2 + which) * cl.nvar + v;
            const size_t o = ((size_t)nt * cl.S + s) * 2 + k;
            u.nest_n[o] = cl.kc_n[ci];
            u.nest_f[2 * o] = cl.kc_f[2 * ci];
            u.nest_f[2 * o + 1] = cl.kc_f[2 * ci + 1];
            k++;
        }
    }
}

// VariantClusterHaplotypes::addNestedHaplotypeKmerStats (VariantClusterHaplotypes.cpp:363-372): the k-mer stats of the enclosing
// allele(s) are booked on the "missing" allele of every variant of this cluster
__device__ __forceinline__ void cl_add_nested_stats(Cl &cl, uint32_t ns, uint32_t s_begin = 0, uint32_t s_end = 0xFFFFFFFFu) {
    const DevUnit &u = *cl.u;
    for (uint32_t s = s_begin; s < cl.S && s < s_end; s++) {
        const uint32_t nk = u.nest_k[(size_t)ns * cl.S + s];
        for (uint32_t k = 0; k < nk; k++) {
            const size_t o = ((size_t)ns * cl.S + s) * 2 + k;
            const uint32_t n = u.nest_n[o];
            for (uint32_t v = 0; v < cl.nvar; v++) {
                const uint32_t ai = cl.alleleBase(v, s) + cl.nalleles(v) - 1;
                cl.as_n[ai * 3 + 0]++; cl.as_f[ai * 3 + 0] += (double)n;
                if (n > 0) {
                    cl.as_n[ai * 3 + 1]++; cl.as_f[ai * 3 + 1] += u.nest_f[2 * o];
                    cl.as_n[ai * 3 + 2]++; cl.as_f[ai * 3 + 2] += u.nest_f[2 * o + 1];
                }
            }
        }
    }
}

// ---- estimateNoise: lock-step iterations over the selected single-cluster groups ---------------
struct NoiseState {
    uint64_t *hist;        // [S][2] sufficient statistics (n_obs, sum of counts) of the CountAllocation histogram
                           // (CountAllocation.cpp:34-57, CountDistribution::calcCountSuffStats :188-200) — the only thing read from it
    double *rates;         // [S] current noise rates (device copy owned by the count dist)
    double *noise_table;   // [S][256]
    double *mean_rates;    // [S]
    double *trace;         // rows of (chain, iteration, rates...) or nullptr
    uint32_t *rng;         // persisted Philox state of CountDistribution::prng (kind 4)
    uint32_t *trace_row;
    const double *lg;      // lg[i] = lgamma((double)i), i < n_lg (the Poisson rows need lgamma(count + 1) of integers only)
    uint32_t n_lg;
    unsigned long long *phase_ns;  // optional (BTG_NOISE_PHASES=1): time block 0 spends in [fill, sample, exchange+update, release] per chain
};

// noiseCountLogPmf with log(rate) hoisted and lgamma of the integer argument read from the table (same values, same order of
// operations as poissonLogProb: value * log(rate) - rate - lgamma(value + 1))
__device__ __forceinline__ double poissonLogProbT(uint32_t value, double rate, double log_rate, const double *lg, uint32_t n_lg) {
    return value * log_rate - rate - (value + 1 < n_lg ? lg[value + 1] : lgamma((double)(value + 1)));
}
static __device__ double noiseCountLogPmfT(double rate, double log_rate, uint32_t c, const double *lg, uint32_t n_lg) {
    double v = poissonLogProbT(c, rate, log_rate, lg, n_lg);
    if (c == 255) {
        uint32_t limit = c;
        double prev;
        do {
            limit++;
            prev = v;
            v = logAddition(v, poissonLogProbT(limit, rate, log_rate, lg, n_lg));
            if (v > 0) { v = 0; break; }
        } while (!doubleCompare(prev, v));
    }
    return v;
}

// CountDistribution::sampleNoiseParameters / resetNoiseRates + updateNoiseCache, on the device so that the
// iteration loop never synchronises with the host.  mode 0: reset from the prior; 1: posterior draw from hist;
// 2: set to the accumulated mean.  One block; thread 0 draws, then all threads rebuild the Poisson rows.
static __device__ void noise_update_block(const NoiseState &ns, uint32_t S, float prior_shape, float prior_scale, uint32_t seed, int mode, int accumulate,
                                   double chain_label, double iter_label, double mean_div, double *sh_rates) {
    if (threadIdx.x == 0) {
        Philox rng;
        rng.load(ns.rng, 0, seed, (uint64_t)-1, 0);
        for (uint32_t s = 0; s < S; s++) {
            double r;
            if (mode == 0) {
                r = rng.gamma((double)prior_shape) * (double)prior_scale;  // CountDistribution.cpp:163-171,202-213
            } else if (mode == 1) {
                const unsigned long long n_obs = ns.hist[s * 2], sum = ns.hist[s * 2 + 1];  // calcCountSuffStats (CountDistribution.cpp:188-200)
                ns.hist[s * 2] = 0; ns.hist[s * 2 + 1] = 0;
                const float shape_f = prior_shape + (float)sum;                                   // float arithmetic as in the
                const float scale_f = prior_scale / ((float)n_obs * prior_scale + 1);             // reference (CountDistribution.cpp:182)
                r = rng.gamma((double)shape_f) * (double)scale_f;
            } else if (mode == 2) {
                r = ns.mean_rates[s] / mean_div;
            } else {
                r = ns.rates[s];  // mode 3: record the current rates, no draw
            }
            ns.rates[s] = r;
            sh_rates[s] = r;
            if (accumulate) ns.mean_rates[s] += r;
        }
        rng.save(ns.rng, 0);
        if (ns.trace) {
            double *row = ns.trace + (size_t)(*ns.trace_row) * (2 + S);
            row[0] = chain_label; row[1] = iter_label;
            for (uint32_t s = 0; s < S; s++) row[2 + s] = sh_rates[s];
            (*ns.trace_row)++;
        }
    }
    __syncthreads();
    if (ns.lg) {
        __shared__ double sh_log_rates[BTG_MAX_SAMPLES];
        if (threadIdx.x < S) sh_log_rates[threadIdx.x] = log(sh_rates[threadIdx.x]);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < S * 256u; i += blockDim.x) ns.noise_table[i] = noiseCountLogPmfT(sh_rates[i >> 8], sh_log_rates[i >> 8], i & 255u, ns.lg, ns.n_lg);
    } else {
        for (uint32_t i = threadIdx.x; i < S * 256u; i += blockDim.x) ns.noise_table[i] = noiseCountLogPmf(sh_rates[i >> 8], i & 255u);
    }
}

// Warp-cooperative fill of the per-(sample, diplotype) k-mer log-likelihood cache for the current non-zero haplotypes:
// entry e of the enumeration goes to lane e % 32, which sums over the k-mer subset in the same order as the
// sequential code (so the cached value is bit-identical).  Used for large clusters in the lock-step noise chain,
// where the slowest cluster sets the pace of every iteration.
// A cluster may be shared by `parts` warps (anywhere in the grid): entries are dealt to them in rounds of 32.
// With many k-mers per entry (n_sub >= 16) the roles turn: the warp takes its entries one at a time, the 32 lanes gather 32
// TERMS of the entry at once, and the terms are then added in subsample order through shuffles — the sum is still the
// sequential one, but an entry costs n_sub/32 gather rounds instead of n_sub dependent gathers in one lane (the slowest
// fill task of an iteration was a lane walking ~60 k-mers: 214 us, profiles/r1_noise_chain_phases.txt).
__device__ __forceinline__ void cl_fill_cache_warp(Cl &cl, const Tables &T, const uint8_t *ploidy, uint32_t lane, uint32_t part, uint32_t parts) {
    const uint32_t H = cl.H, n_sub = cl.misc[kNSub];
    const bool by_terms = n_sub >= 16;
    uint32_t e = 0;
    for (uint32_t s = 0; s < cl.S; s++) {
        const uint8_t pl = ploidy[s];
        if (pl == 0) continue;
        const uint32_t g = cl.u->sample_gender[s];
        for (uint32_t a = 0; a < H; a++) {
            if (!cl.nz[a]) continue;
            const uint32_t b_end = pl == 2 ? H : a + 1;
            for (uint32_t b = a; b < b_end; b++) {
                if (pl == 2 && !cl.nz[b]) continue;
                const uint32_t mine = e++;
                const uint32_t bb = pl == 2 ? b : NONE;
                const size_t ci = (size_t)s * cl.Dall + cl.slot(a, bb == NONE ? H : bb);
                if (by_terms) {  // warp-uniform control flow from here on
                    if ((mine % parts) != part) continue;
                    if (cl.ucache[ci] == cl.ucache[ci]) continue;  // already cached (same address in every lane)
                    double acc = 0;
                    for (uint32_t base = 0; base < n_sub; base += 32) {
                        const uint32_t i = base + lane;
                        const double v = i < n_sub ? *tile_term(cl, T, s, g, a, bb, i) : 0.0;
                        const uint32_t m = n_sub - base < 32 ? n_sub - base : 32;
                        for (uint32_t j = 0; j < m; j++) acc += __shfl_sync(0xFFFFFFFFu, v, j);  // in subsample order
                    }
                    if (lane == 0) cl.ucache[ci] = acc;
                } else {
                    if ((mine & 31u) != lane || ((mine >> 5) % parts) != part) continue;
                    if (cl.ucache[ci] == cl.ucache[ci]) continue;  // already cached
                    cl.ucache[ci] = tile_entry_sum(cl, T, s, a, bb, n_sub);
                }
            }
        }
    }
}

// one lock-step iteration of one cluster by a single thread (sampleGenotypesCallback body without the noise counts)
__device__ __forceinline__ void noise_iteration_thread(Cl &cl, const DevUnit &du, const Tables &T, const btg_gibbs_opts &o, bool collect) {
    const uint64_t gidx = group_index(o, cl.g);
    const uint8_t *ploidy = du.group_ploidy + (size_t)cl.g * du.S;
    Philox prng, fr;
    prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
    fr.load(cl.rng, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
    cl_sample_diplotypes<false, true>(cl, T, ploidy, collect, prng);
    cl_sample_frequencies(cl, fr);
    prng.save(cl.rng, kRng0);
    fr.save(cl.rng, kRng1);
}

// Grid-wide barrier of the persistent chain kernel (all blocks are co-resident: cooperative launch).  One thread per block
// arrives on a counter and then polls a generation word WITH BACK-OFF.  cooperative_groups' grid.sync() polls without
// pause: with ~1200 blocks waiting for the few warps that still work, the polls queue up on the one L2 slice that holds the
// barrier word and every load of the working warps that maps to that slice waits behind them (measured: ~50 us of fixed cost
// per phase and 0.5 us per dependent load, profiles/r1_noise_chain_phases.txt).
struct GridBarrier {
    unsigned int *count, *gen;
};
__device__ __forceinline__ void grid_barrier(const GridBarrier &b) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int g;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(b.gen) : "memory");
        __threadfence();  // this block's writes are visible before its arrival
        if (atomicAdd(b.count, 1u) == gridDim.x - 1) {
            *b.count = 0;
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(b.gen), "r"(g + 1) : "memory");
        } else {
            unsigned int now, ns_sleep = 64;
            const unsigned long long t0 = global_timer_ns();
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(b.gen) : "memory");
                if (now != g) break;
                __nanosleep(ns_sleep);
                if (ns_sleep < 1024) ns_sleep *= 2;
                // never hang the GPU: if the grid is not co-resident (it always is: cooperative launch) give up after 60 s; the
                // host sees the flag in the generation word's neighbour and reports an error
                if (ns_sleep == 1024 && global_timer_ns() - t0 > 60000000000ull) { atomicExch(b.gen + 1, 1u); break; }
            }
        }
    }
    __syncthreads();
}

template <class T> T *upload(const T *h, size_t n, bool &ok) {
    T *d = nullptr;
    if (btg::dmalloc(&d, (n ? n : 1) * sizeof(T)) != cudaSuccess) { ok = false; return nullptr; }
    if (n && cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
    return d;
}

// gibbs_wide.cu: the warp-per-cluster kernels (lane = sample)
cudaError_t wide_estimate_genotypes(const DevUnit &du, const Tables &T, const btg_gibbs_opts &o, const ResultView &R, cudaStream_t st);
cudaError_t wide_noise_chain(const DevUnit &du, const Tables &T, const btg_gibbs_opts &o, const uint32_t *d_sel, uint32_t n_sel, uint32_t n_big,
                             const uint32_t *d_tasks, uint32_t n_tasks, uint32_t chain, uint32_t iters,
                             const NoiseState &ns, float prior_shape, float prior_scale, unsigned long long *hist, int joint, const PeerExchange &px,
                             const GridBarrier &gb, uint32_t share, int sm_count, cudaStream_t st);
}  // namespace btg_gibbs
using namespace btg_gibbs;

struct btg_unit {
    DevUnit du{};
    void *res_view = nullptr;  // ResultView* (host struct with device pointers), allocated on first use
    std::vector<void *> allocs;
    std::vector<uint64_t> h_valt_off, h_allele_off, h_geno_off;
    std::vector<uint32_t> h_nhap, h_group_nvar;
    std::vector<uint64_t> h_group_cluster_off, h_cl_var_off;
    std::vector<ClusterLayout> h_layout;
    std::vector<SlotLayout> h_slots;
    std::vector<uint32_t> h_fill_cost;  // table lookups of one full cache fill: S * D * n_uniq / 10
    uint64_t n_variants = 0, n_alleles_total = 0, n_shared = 0;
    uint32_t max_h = 0;
    // sizes of the mutable arena pools, and shadow copies of them: estimateNoise runs several chains at once, each on its own
    // arena (shadow_du[k] = du with private pools; the descriptors are shared)
    uint64_t f64_total = 0, u32_total = 0, u8_total = 0, tile_total = 0;
    std::vector<DevUnit> shadow_du;
};
