// gibbs.cu — the per-cluster Gibbs sampler on the device (sm_100a).
//
// Replaces, for one inference unit resident in HBM:
//   InferenceEngine::estimateGenotypes / estimateNoise        src/bayesTyper/InferenceEngine.cpp:135-382
//   VariantClusterGroup::estimateGenotypes / runGibbsSample   src/bayesTyper/VariantClusterGroup.cpp:220-250
//   VariantClusterGenotyper (ctor, reset, sampleDiplotypes, sampleDiplotype, calcDiplotypeLogProb,
//     sampleHaplotypeFrequencies, getNoiseCounts, getGenotypes…)  src/bayesTyper/VariantClusterGenotyper.cpp
//   VariantClusterHaplotypes (sampleKmerSubset, updateAlleleKmerStats…) src/bayesTyper/VariantClusterHaplotypes.cpp
//   CountDistribution / NegativeBinomialDistribution tables   src/bayesTyper/CountDistribution.cpp
//   (Sparse)FrequencyDistribution, SparsityEstimator, DiscreteSampler, KmerStats, CountAllocation
//
// Mapping.  Variant-cluster groups are independent and, inside a group, the sampler is
// a strictly sequential chain (sample s+1 sees the haplotype counts left by sample s; iteration i+1 sees
// the frequencies drawn in iteration i).  The unit of parallelism is therefore the GROUP: one thread
// walks one group through all chains x iterations with its state in a private arena slice; groups are
// sorted by cost so that the 32 lanes of a warp carry similar work.  All arithmetic is f64 (the
// reference's), table lookups go to the L2-resident log-pmf cache ([S][256][256] doubles).
//
// Groups made of ONE cluster (the bulk of any unit) run one thread per cluster in k_estimate_genotypes.  Groups
// with nested clusters (VariantClusterGroup::runGibbsSample recursion, multicluster k-mers sharing a multiplicity
// record) run one thread per GROUP in k_estimate_genotypes_nested, which walks the group's clusters in the
// reference's depth-first order every iteration.  The joint noise mode still requires single-cluster groups.
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

namespace {

constexpr uint16_t NONE = 0xFFFF;  // Utils::ushort_overflow
constexpr double kDoubleEps100 = 2.220446049250313e-16 * 100;
constexpr float kFloatEps100 = 1.1920929e-07f * 100;

__host__ __device__ inline bool doubleCompare(double a, double b) {  // Utils.hpp:81-87
    return (a == b) || (fabs(a - b) < fabs(a < b ? a : b) * kDoubleEps100);
}
__host__ __device__ inline bool floatCompare(float a, float b) {  // Utils.hpp:89-95
    return (a == b) || (fabsf(a - b) < fabsf(a < b ? a : b) * kFloatEps100);
}
__host__ __device__ inline bool floatLess(float a, float b) { return (a < b) && !floatCompare(a, b); }
BTG_LEAF double logAddition(double a, double b) {  // Utils.hpp:105-124
    return a < b ? b + m_log1p(m_exp(a - b)) : a + m_log1p(m_exp(b - a));
}

// ---------------------------------------------------------------------------------------------
// CountDistribution tables
// ---------------------------------------------------------------------------------------------
__device__ double nbLogPmf(double p, double size, uint32_t obs, uint32_t scale) {  // NegativeBinomialDistribution.cpp:121-147
    const double coef = lgamma(obs + size * scale) - lgamma(size * scale) - lgamma((double)(obs + 1));
    return coef + log(p) * size * scale + log(1 - p) * obs;
}
__device__ double poissonLogProb(uint32_t value, double rate) {  // CountDistribution.cpp:349-352
    return value * log(rate) - rate - lgamma((double)(value + 1));
}

// CountDistribution::updateGenomicCache / genomicCountLogPmf (CountDistribution.cpp:215-238,267-312): one thread per (s, m, c)
__global__ void k_genomic_table(const double *__restrict__ p, const double *__restrict__ size, uint32_t S, double *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * 65536u) return;
    const uint32_t s = i >> 16, m = (i >> 8) & 255u, c = i & 255u;
    double v;
    if (m == 0) {
        v = c == 0 ? 0.0 : -INFINITY;
    } else {
        v = nbLogPmf(p[s], size[s], c, m);
        if (c == 255) {
            uint32_t limit = c;
            double prev;
            do {
                limit++;
                prev = v;
                v = logAddition(v, nbLogPmf(p[s], size[s], limit, m));
                if (v > 0) { v = 0; break; }
            } while (!doubleCompare(prev, v));
        }
    }
    out[i] = v;
}

// CountDistribution::updateNoiseCache / noiseCountLogPmf (CountDistribution.cpp:240-253,314-347): one thread per (s, c)
__device__ double noiseCountLogPmf(double rate, uint32_t c) {
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
__global__ void k_noise_table(const double *__restrict__ rates, uint32_t S, double *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * 256u) return;
    out[i] = noiseCountLogPmf(rates[i >> 8], i & 255u);
}

__global__ void k_lgamma_int(double *out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lgamma((double)i);  // out[0] = inf, never read
}

}  // namespace

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
namespace {

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
};
template <class T> struct LaneArr {
    T *p;
    __device__ __forceinline__ T &operator[](uint32_t i) const { return p[i * 32u]; }
    __device__ __forceinline__ LaneArr<T> operator+(size_t i) const { return LaneArr<T>{p + i * 32}; }
};

// k-mer tile accessor: lane-interleaved (stride 32) for clusters that run one per thread — the 32 clusters of a warp
// read one sector per element — and dense (stride 1) for the large clusters that a whole warp works on, where all lanes
// read the same row and an interleaved layout would cost one sector per byte
struct TileArr {
    uint8_t *p;
    uint32_t stride;
    __device__ __forceinline__ uint8_t &operator[](uint32_t i) const { return p[(size_t)i * stride]; }
};
constexpr uint32_t kBigFillCost = 128;
constexpr uint32_t kChainSplit = 20;      // virtual threads of a chain-split cluster (chain c runs on thread c % kChainSplit)
constexpr uint32_t kSplitFillCost = 64;   // clusters above this fill cost are chain-split in the default mode (sweep: profiles/r1_gibbs_tail.txt)
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

struct DevUnit {
    uint32_t S, G, C;
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
__host__ __device__ inline ArenaSizes arena_sizes(uint32_t S, uint32_t H, uint32_t K, uint32_t nvar, uint32_t n_uniq, uint32_t n_alleles, uint32_t Dall,
                                                  uint32_t n_multi) {
    ArenaSizes a;
    a.f64 = (uint64_t)H * 2        /* freq, log(freq) */
          + (H + 1)                /* simplex prob vector (H > kSimplexTableMaxH: most recent key only) */
          + (H <= kSimplexTableMaxH ? (uint64_t)H * (H + 1) : 0) /* else: one vector per plus-count, + lengths */
          + (uint64_t)S * Dall     /* unique diplotype log-prob cache */
          + Dall                   /* cumulative log-probs of one draw */
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

enum Misc { kNSub = 0, kNumHap = 1, kNumMissing = 2, kSimplexNobs = 3, kSimplexPlus = 4, kSimplexLen = 5, kSparse = 6, kCover = 7, kRng0 = 8, kRng1 = 16,
            kNMultiSub = 24, kUseMulti = 25 };

// per-cluster view
struct Cl {
    uint32_t S, H, K, nvar, n_uniq, Dall, n_alleles, c, g, n_multi;
    bool has_simplex_tab;
    uint64_t row0, var0;
    const DevUnit *u;
    const uint8_t *M;
    LaneArr<double> freq, logf, simplex, simplex_tab, ucache, cum, kc_f, as_f, mcache, fmisc;
    LaneArr<uint32_t> obs, uniq, uniq_sub, cnt, tally, kc_n, as_n, dipl, multi, multi_sub, misc;
    LaneArr<uint8_t> nz, uncovered, stats_update, sample_multi;
    TileArr tile_m, tile_c, tile_ic;

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
        const SlotLayout SL = du.slots[L.pos >> 5];
        const uint32_t lane = L.pos & 31u;
        LaneArr<double> f{du.f64_pool + SL.f64_off + lane};
        freq = f; f = f + SL.H;
        logf = f; f = f + SL.H;
        simplex = f; f = f + (SL.H + 1);
        has_simplex_tab = SL.H <= kSimplexTableMaxH;
        simplex_tab = f; f = f + (has_simplex_tab ? (uint64_t)SL.H * (SL.H + 1) : 0);
        ucache = f; f = f + (uint64_t)S * SL.Dall;
        cum = f; f = f + SL.Dall;
        kc_f = f; f = f + (uint64_t)S * 2 * SL.nvar * 2;
        as_f = f; f = f + (uint64_t)SL.n_alleles * S * 3;
        mcache = f; f = f + (SL.n_multi ? (uint64_t)S * SL.Dall : 0);
        fmisc = f;
        LaneArr<uint32_t> w{du.u32_pool + SL.u32_off + lane};
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
        misc = w;
        LaneArr<uint8_t> b{du.u8_pool + SL.u8_off + lane};
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
            tile_m = TileArr{b.p, 32}; b = b + (uint64_t)SL.n_uniq * SL.H;
            tile_c = TileArr{b.p, 32}; b = b + (uint64_t)SL.n_uniq * S;
            tile_ic = TileArr{b.p, 32};
        }
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
__device__ __forceinline__ bool cl_fill_cache_rows(Cl &cl, const Tables &T, const uint8_t *ploidy) {
    const uint32_t H = cl.H, n_sub = cl.misc[kNSub];
    uint32_t hs0 = 0, hs1 = 0, hs2 = 0, hs3 = 0, n = 0;
    for (uint32_t h = 0; h < H; h++)
        if (cl.nz[h]) {
            if (n == 0) hs0 = h; else if (n == 1) hs1 = h; else if (n == 2) hs2 = h; else if (n == 3) hs3 = h; else return false;
            n++;
        }
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
    double acc = cl.ucache[ci];
    if (acc != acc) {  // not cached yet
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
        cl.ucache[ci] = acc;
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
template <bool MC = false, bool TILE = false>
__device__ __forceinline__ void cl_sample_diplotype(Cl &cl, const Tables &T, uint32_t s, uint8_t ploidy, Philox &prng, uint64_t nzm) {
    uint32_t n = 0;
    double run = 0;
    const uint32_t H = cl.H;
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
                cl.cum[n++] = run;
            }
        }
    } else if (ploidy == 1) {
        for (uint32_t a = next(0); a < H; a = next(a + 1)) {
            const double lp = cl_dipl_log_prob<MC, TILE>(cl, T, s, a, NONE);
            run = n == 0 ? lp : logAddition(lp, run);
            cl.cum[n++] = run;
        }
    } else {
        cl.cum[n++] = 0;
    }
    const double x = m_log(prng.u01()) + run;
    uint32_t idx = 0;
    if (n > 1) {  // upper_bound
        uint32_t lo = 0, hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (x < cl.cum[mid]) hi = mid; else lo = mid + 1; }
        idx = lo < n ? lo : n - 1;
    }
    // map the outcome index back to (a, b) in enumeration order
    uint32_t da = NONE, db = NONE;
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
    cl_increment(cl, da);
    cl_increment(cl, db);
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
__device__ __forceinline__ void cl_update_allele_stats(Cl &cl) {
    for (uint32_t s = 0; s < cl.S; s++) {
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

__device__ __forceinline__ void cl_summarise(Cl &cl, const btg_gibbs_opts &o, const uint8_t *ploidy, const ResultView &R) {
    const uint32_t S = cl.S, H = cl.H;
    for (uint32_t v = 0; v < cl.nvar; v++) {
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

// InferenceEngine::estimateGenotypesCallback (InferenceEngine.cpp:278-333): one thread = one group, all chains
// Chains of one cluster are independent in this mode: every chain re-keys the cluster's two random streams with its chain
// index, shuffles the original k-mer order and starts from reset frequencies; tallies and allele statistics are sums over
// chains.  (The reference keeps ONE mt19937 running through all chains of a genotyper, InferenceEngine.cpp:292-306 — a
// property of its generator, not of the model; oracle-P follows the per-chain contract.)  That lets the few large clusters,
// which otherwise set the kernel's tail (one thread: 1.3 s while the mean thread takes 0.16 s, profiles/r1_gibbs_tail.txt), run
// their chains on kChainSplit threads with private arena positions; k_merge_split adds the pieces up and summarises.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(64, MIN_BLOCKS) k_estimate_genotypes(DevUnit du, Tables T, btg_gibbs_opts o, ResultView R, int reconverge, unsigned long long *dbg) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long t_in = dbg ? global_timer_ns() : 0;
    // no early return: every lane of the warp reaches the __syncwarp()s below
    const uint32_t n_virtual = du.n_split * kChainSplit;
    bool live;
    uint32_t cluster = 0, pos = NONE32, chain0 = 0, chain_step = 1;
    bool split = false;
    if (t < n_virtual) {  // the large clusters first: their blocks start before everything else
        const uint32_t j = t / kChainSplit, v = t % kChainSplit;
        cluster = du.split_cluster[j];
        pos = du.split_pos[(size_t)j * kChainSplit + v];
        chain0 = v; chain_step = kChainSplit;
        split = true; live = true;
    } else {
        const uint32_t i = t - n_virtual;
        live = i < du.n_regular;
        cluster = du.order[live ? i : 0];
        if (live && du.split_of[cluster] != NONE32) live = false;  // handled above
    }
    Cl cl;
    cl.bind(du, cluster, pos);
    const uint64_t gidx = o.group_index_base + cl.g;
    const uint8_t *ploidy = du.group_ploidy + (size_t)cl.g * du.S;
    Philox prng, fr;
    prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, 0);
    fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, 0);
    if (live) cl_construct(cl, o, gidx, 0);
    const uint32_t iters = (uint32_t)o.gibbs_burn_in + o.gibbs_samples;
    // warp-uniform trip count (lanes of split clusters take fewer chains; they idle through the rest)
    for (uint32_t round = 0; round < o.n_chains; round++) {
        const uint32_t chain = chain0 + round * chain_step;
        const bool run = live && chain < o.n_chains;
        if (run) {
            prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, chain);
            fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, chain);
            cl_reset<false, false, true>(cl, o, prng);
        }
        if (reconverge) __syncwarp();
        if (__all_sync(0xFFFFFFFFu, !run)) continue;  // nothing left for this warp in this round
        for (uint32_t it = 0; it < iters; it++) {
            if (run) cl_sample_diplotypes(cl, T, ploidy, it >= o.gibbs_burn_in, prng);
            if (reconverge) __syncwarp();
            if (run) cl_sample_frequencies(cl, fr);
            if (reconverge) __syncwarp();
        }
    }
    if (live && !split) cl_summarise(cl, o, ploidy, R);
    if (dbg && live) {  // BTG_GIBBS_TIMING=1: slowest thread and the sum of all per-thread times
        const unsigned long long dt = global_timer_ns() - t_in;
        atomicMax(dbg, (dt << 32) | cl.c);
        atomicAdd(dbg + 1, dt);
        if (cl.H > 4) atomicAdd(dbg + 2, dt);
    }
}

// chain-split clusters: add the tallies and allele statistics of the virtual threads 1.. into thread 0's arena position
// (ascending order, so the f64 sums are reproducible) and summarise
__global__ void __launch_bounds__(64) k_merge_split(DevUnit du, btg_gibbs_opts o, ResultView R) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= du.n_split) return;
    Cl cl, part;
    cl.bind(du, du.split_cluster[j]);
    for (uint32_t v = 1; v < kChainSplit; v++) {
        part.bind(du, du.split_cluster[j], du.split_pos[(size_t)j * kChainSplit + v]);
        for (uint32_t i = 0; i < cl.Dall * cl.S; i++) cl.tally[i] += part.tally[i];
        for (uint32_t i = 0; i < cl.n_alleles * cl.S * 3; i++) { cl.as_n[i] += part.as_n[i]; cl.as_f[i] += part.as_f[i]; }
    }
    cl_summarise(cl, o, du.group_ploidy + (size_t)cl.g * du.S, R);
}


// ---- groups with nested clusters ------------------------------------------------------------------
// VariantClusterGenotyper::updateNestedVariantClusterInfo / updateNestedPloidy / addNestedKmerStats (…Genotyper.cpp:140-206):
// `cl` is the parent that has just been sampled; the child's incoming info (slot nt) already holds a copy of the parent's own
__device__ __forceinline__ void cl_update_nested_info(Cl &cl, uint32_t nt, uint32_t child_cluster_idx) {
    const DevUnit &u = *cl.u;
    uint64_t dep = u.cl_dep_off[cl.c];
    while (dep < u.cl_dep_off[cl.c + 1] && u.dep_cluster[dep] != child_cluster_idx) dep++;
    for (uint32_t s = 0; s < cl.S; s++) {
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
__device__ __forceinline__ void cl_add_nested_stats(Cl &cl, uint32_t ns) {
    const DevUnit &u = *cl.u;
    for (uint32_t s = 0; s < cl.S; s++) {
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

// InferenceEngine::estimateGenotypesCallback for a group with several clusters: one thread = one GROUP.  Per chain the branch
// orderings are shuffled (cumulatively, VariantClusterGroup.cpp:208-218) and flattened into the depth-first order that
// runGibbsSample's recursion (…Group.cpp:236-250) visits; every iteration walks that order, each cluster passing the
// NestedVariantClusterInfo of its children on before they run.
__global__ void __launch_bounds__(64) k_estimate_genotypes_nested(DevUnit du, Tables T, btg_gibbs_opts o, ResultView R) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= du.n_nested_groups) return;
    const uint32_t g = du.nested_groups[i], S = du.S;
    const uint64_t c0 = du.group_cluster_off[g];
    const uint32_t n = (uint32_t)(du.group_cluster_off[g + 1] - c0);
    const uint64_t gidx = o.group_index_base + g;
    const uint8_t *ploidy = du.group_ploidy + (size_t)g * S;
    const uint64_t s0 = du.group_src_off[g], s1 = du.group_src_off[g + 1];
    const uint64_t e0 = du.cl_edge_off[c0], e1 = du.cl_edge_off[c0 + n];
    for (uint64_t e = s0; e < s1; e++) du.src_mut[e] = du.group_src[e];
    for (uint64_t e = e0; e < e1; e++) du.edge_mut[e] = du.edge_dst[e];
    Cl cl;
    for (uint32_t j = 0; j < n; j++) {  // VariantClusterGroup::initGenotyper: genotypers are constructed once
        cl.bind(du, (uint32_t)(c0 + j));
        cl_construct(cl, o, gidx, 0);
        Philox prng, fr;
        prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, 0);
        fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, 0);
        prng.save(cl.misc, kRng0);
        fr.save(cl.misc, kRng1);
    }
    const uint32_t iters = (uint32_t)o.gibbs_burn_in + o.gibbs_samples;
    for (uint32_t chain = 0; chain < o.n_chains; chain++) {
        for (uint32_t j = 0; j < n; j++) {
            cl.bind(du, (uint32_t)(c0 + j));
            Philox prng;
            prng.load(cl.misc, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
            cl_reset<true>(cl, o, prng);
            prng.save(cl.misc, kRng0);
        }
        {   // shuffleBranchOrdering: sources, then every vertex's out-edges in vertex order (stream kind 3)
            Philox br;
            br.init(o.random_seed, gidx, 0, kRngBranch, chain);
            for (uint64_t m = s1 - s0; m > 1; m--) {
                const uint32_t j = br.uniform_int((uint32_t)m);
                const uint32_t t = du.src_mut[s0 + m - 1]; du.src_mut[s0 + m - 1] = du.src_mut[s0 + j]; du.src_mut[s0 + j] = t;
            }
            for (uint32_t v = 0; v < n; v++) {
                const uint64_t b0 = du.cl_edge_off[c0 + v];
                for (uint64_t m = du.cl_edge_off[c0 + v + 1] - b0; m > 1; m--) {
                    const uint32_t j = br.uniform_int((uint32_t)m);
                    const uint32_t t = du.edge_mut[b0 + m - 1]; du.edge_mut[b0 + m - 1] = du.edge_mut[b0 + j]; du.edge_mut[b0 + j] = t;
                }
            }
        }
        {   // depth-first pre-order of the forest
            uint32_t top = 0, len = 0;
            for (uint64_t e = s1; e > s0; e--) du.dfs_stack[c0 + top++] = du.src_mut[e - 1];
            while (top > 0) {
                const uint32_t v = du.dfs_stack[c0 + --top];
                du.dfs_order[c0 + len++] = v;
                for (uint64_t e = du.cl_edge_off[c0 + v + 1]; e > du.cl_edge_off[c0 + v]; e--) du.dfs_stack[c0 + top++] = du.edge_mut[e - 1];
            }
        }
        for (uint64_t e = s0; e < s1; e++) {  // the info a source vertex receives: the chromosome ploidy, no enclosing allele
            const uint32_t ns = du.nest_slot[c0 + du.src_mut[e]];
            for (uint32_t s = 0; s < S; s++) { du.nest_pl[(size_t)ns * S + s] = ploidy[s]; du.nest_k[(size_t)ns * S + s] = 0; }
        }
        for (uint32_t it = 0; it < iters; it++) {
            const bool collect = it >= o.gibbs_burn_in;
            for (uint32_t pos = 0; pos < n; pos++) {
                const uint32_t v = du.dfs_order[c0 + pos];
                cl.bind(du, (uint32_t)(c0 + v));
                const uint32_t ns = du.nest_slot[c0 + v];
                Philox prng, fr;
                prng.load(cl.misc, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
                fr.load(cl.misc, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
                cl_sample_diplotypes<true>(cl, T, du.nest_pl + (size_t)ns * S, collect, prng);
                if (collect) cl_add_nested_stats(cl, ns);
                cl_sample_frequencies(cl, fr);
                prng.save(cl.misc, kRng0);
                fr.save(cl.misc, kRng1);
                for (uint64_t e = du.cl_edge_off[c0 + v]; e < du.cl_edge_off[c0 + v + 1]; e++) {
                    const uint32_t t = du.edge_mut[e], nt = du.nest_slot[c0 + t];
                    for (uint32_t s = 0; s < S; s++) {
                        du.nest_pl[(size_t)nt * S + s] = du.nest_pl[(size_t)ns * S + s];
                        const uint32_t nk = du.nest_k[(size_t)ns * S + s];
                        du.nest_k[(size_t)nt * S + s] = (uint8_t)nk;
                        for (uint32_t k = 0; k < nk; k++) {
                            const size_t a = ((size_t)ns * S + s) * 2 + k, b = ((size_t)nt * S + s) * 2 + k;
                            du.nest_n[b] = du.nest_n[a]; du.nest_f[2 * b] = du.nest_f[2 * a]; du.nest_f[2 * b + 1] = du.nest_f[2 * a + 1];
                        }
                    }
                    cl_update_nested_info(cl, nt, du.cluster_idx[c0 + t]);
                }
            }
        }
    }
    for (uint32_t j = 0; j < n; j++) {  // collectGenotypes: every cluster is summarised with the chromosome ploidy
        cl.bind(du, (uint32_t)(c0 + j));
        cl_summarise(cl, o, ploidy, R);
    }
}

// VariantClusterGroup::collectGenotypes for every cluster (joint mode collects after all chains)
__global__ void __launch_bounds__(64) k_summarise(DevUnit du, btg_gibbs_opts o, ResultView R) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= du.C) return;
    Cl cl;
    cl.bind(du, du.order[i]);
    cl_summarise(cl, o, du.group_ploidy + (size_t)cl.g * du.S, R);
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
__device__ double noiseCountLogPmfT(double rate, double log_rate, uint32_t c, const double *lg, uint32_t n_lg) {
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
__device__ void noise_update_block(const NoiseState &ns, uint32_t S, float prior_shape, float prior_scale, uint32_t seed, int mode, int accumulate,
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

__global__ void k_noise_update(NoiseState ns, uint32_t S, float prior_shape, float prior_scale, uint32_t seed, int mode, int accumulate,
                               double chain_label, double iter_label, double mean_div) {
    __shared__ double sh_rates[BTG_MAX_SAMPLES];
    noise_update_block(ns, S, prior_shape, prior_scale, seed, mode, accumulate, chain_label, iter_label, mean_div, sh_rates);
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
    const uint64_t gidx = o.group_index_base + cl.g;
    const uint8_t *ploidy = du.group_ploidy + (size_t)cl.g * du.S;
    Philox prng, fr;
    prng.load(cl.misc, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
    fr.load(cl.misc, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
    cl_sample_diplotypes<false, true>(cl, T, ploidy, collect, prng);
    cl_sample_frequencies(cl, fr);
    prng.save(cl.misc, kRng0);
    fr.save(cl.misc, kRng1);
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

// One whole chain of estimateNoise as ONE persistent cooperative kernel (InferenceEngine.cpp:191-253): every thread keeps its
// clusters' state hot in L1 across the 350 iterations; the per-iteration "join + merge + sampleNoiseParameters" of the
// reference (thread spawn/join per iteration, InferenceEngine.cpp:213-226) becomes two grid-wide barriers around block 0's
// histogram -> Gamma draw -> Poisson-row rebuild.
// joint = 0: estimateNoise (fresh genotypers each chain, streams of chain `chain`, nothing collected)
// joint = 1: estimateNoiseAndGenotypes (InferenceEngine.cpp:384-472): genotypers are constructed in the first chain only and
//            persist (streams of chain 0), samples are collected after the burn-in
#ifndef BTG_NOISE_MINBLOCKS
#define BTG_NOISE_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(256, BTG_NOISE_MINBLOCKS) k_noise_chain(DevUnit du, Tables T, btg_gibbs_opts o, const uint32_t *sel, uint32_t n_sel, uint32_t n_big, uint32_t chain,
                                                       uint32_t iters, NoiseState ns, float prior_shape, float prior_scale, unsigned long long *hist, int joint,
                                                       PeerExchange px, const uint32_t *fill_tasks, uint32_t n_fill_tasks, GridBarrier gb) {
    __shared__ unsigned long long sh_tot[kMailRow];
    __shared__ double sh_rates[BTG_MAX_SAMPLES];
    // getNoiseCounts of the block's clusters: only (n_obs, sum) per sample are ever read from the merged CountAllocation,
    // so they are summed in shared memory and leave the block as <= 2S global atomics per iteration
    __shared__ unsigned long long sh_stat[BTG_MAX_SAMPLES * 2];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const unsigned long long t_start = ns.phase_ns && blockIdx.x == 0 && threadIdx.x == 0 ? global_timer_ns() : 0;
    for (uint32_t i = tid >> 5; i < n_big; i += nthreads >> 5) {  // large clusters: constructed by a warp, reset by its lane 0
        Cl cl;
        cl.bind(du, sel[i]);
        const uint64_t gidx = o.group_index_base + cl.g;
        const uint32_t stream_chain = joint ? 0 : chain;
        if (!joint || chain == 1) cl_construct_warp(cl, o, gidx, stream_chain, tid & 31u);
        if ((tid & 31u) == 0) {
            Philox prng, fr;
            if (!joint || chain == 1) {
                prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, stream_chain);
                fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, stream_chain);
            } else {
                prng.load(cl.misc, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
                fr.load(cl.misc, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
            }
            cl_reset<false, true>(cl, o, prng);
            prng.save(cl.misc, kRng0);
            fr.save(cl.misc, kRng1);
        }
        __syncwarp();
    }
    for (uint32_t i = n_big + tid; i < n_sel; i += nthreads) {  // initGenotypersCallback: fresh genotypers every chain
        Cl cl;
        cl.bind(du, sel[i]);
        const uint64_t gidx = o.group_index_base + cl.g;
        Philox prng, fr;
        if (!joint || chain == 1) {
            const uint32_t stream_chain = joint ? 0 : chain;
            cl_construct(cl, o, gidx, stream_chain);
            prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, stream_chain);
            fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, stream_chain);
        } else {
            prng.load(cl.misc, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
            fr.load(cl.misc, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
        }
        cl_reset<false, true>(cl, o, prng);
        prng.save(cl.misc, kRng0);
        fr.save(cl.misc, kRng1);
    }
    if (blockIdx.x == 0 && ns.trace) noise_update_block(ns, du.S, prior_shape, prior_scale, o.random_seed, 3, 0, (double)chain, 0, 1, sh_rates);
    grid_barrier(gb);
    if (t_start) ns.phase_ns[7] += global_timer_ns() - t_start;  // construct + reset of every selected cluster
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
#if BTG_NOISE_TIMING
    unsigned long long sub_acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // per-thread clock sums of the one-thread sub-steps, flushed once at the end
#endif
    const bool timing = ns.phase_ns && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t_prev = timing ? global_timer_ns() : 0;
    auto lap = [&](int phase) {
        if (timing) { const unsigned long long t = global_timer_ns(); ns.phase_ns[phase] += t - t_prev; t_prev = t; }
    };
    for (uint32_t it = 1; it <= iters; it++) {
        if (threadIdx.x < 2 * du.S) sh_stat[threadIdx.x] = 0;
        __syncthreads();
        // phase A: the diplotype caches of the large clusters sel[0 .. n_big) are filled by the whole grid: fill task t =
        // (cluster, part, parts) gives one warp every parts-th round of 32 cache entries of that cluster, so the slowest
        // cluster no longer sets the pace of the iteration with a single warp
        // (tasks are dealt from the LAST warp downwards: the one-thread clusters below occupy the first threads of the grid)
        for (uint32_t t = (nthreads >> 5) - 1 - (tid >> 5); t < n_fill_tasks; t += nthreads >> 5) {
            const unsigned long long t_in = BTG_NOISE_TIMING && ns.phase_ns ? global_timer_ns() : 0;
            Cl cl;
            cl.bind(du, sel[fill_tasks[3 * t]]);
            cl_fill_cache_warp(cl, T, du.group_ploidy + (size_t)cl.g * du.S, tid & 31u, fill_tasks[3 * t + 1], fill_tasks[3 * t + 2]);
            if (BTG_NOISE_TIMING && ns.phase_ns) { const unsigned long long v = ((global_timer_ns() - t_in) << 32) | cl.c; atomicMax(ns.phase_ns + 4, v); atomicMax(ns.phase_ns + 8 + 3 * (it - 1), v); }
        }
        // ... while the one-thread clusters sel[n_big .. n_sel) take their whole step in the same phase (they do not depend on the
        // fill tasks): the warps that hold no fill task are not idle at the barrier while the large caches are filled
        for (uint32_t i = n_big + tid; i < n_sel; i += nthreads) {  // sampleGenotypesCallback
            const unsigned long long t_in = BTG_NOISE_TIMING && ns.phase_ns ? global_timer_ns() : 0;
#if BTG_NOISE_TIMING
            const bool sub = ns.phase_ns != nullptr;
            long long ck = sub ? clock64() : 0;
            auto tick = [&](int k) { if (sub) { const long long now = clock64(); sub_acc[k] += (unsigned int)(now - ck); ck = now; } };
#else
            auto tick = [](int) {};
#endif
            Cl cl;
            cl.bind(du, sel[i]);
            tick(0);
            const uint8_t *ploidy_i = du.group_ploidy + (size_t)cl.g * du.S;
            cl_fill_cache_rows(cl, T, ploidy_i);  // > 4 live haplotypes: entries are filled on demand
            tick(1);
            {
                const uint64_t gidx = o.group_index_base + cl.g;
                Philox prng, fr;
                prng.load(cl.misc, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
                fr.load(cl.misc, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
                tick(2);
                cl_sample_diplotypes<false, true>(cl, T, ploidy_i, joint && it > o.gibbs_burn_in, prng);
                tick(3);
                cl_sample_frequencies(cl, fr);
                tick(4);
                prng.save(cl.misc, kRng0);
                fr.save(cl.misc, kRng1);
                tick(5);
            }
#if BTG_NOISE_TIMING
            if (sub) sub_acc[8]++;
#endif
            if (BTG_NOISE_TIMING && ns.phase_ns) { const unsigned long long v = ((global_timer_ns() - t_in) << 32) | cl.c; atomicMax(ns.phase_ns + 6, v); atomicMax(ns.phase_ns + 8 + 3 * (it - 1) + 2, v); }
            const uint32_t n_sub = cl.misc[kNSub];
            for (uint32_t s = 0; s < cl.S; s++) {  // getNoiseCounts
                const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16;
                uint32_t n0 = 0, c0 = 0;
                const uint32_t g = du.sample_gender[s];
#pragma unroll 4
                for (uint32_t j = 0; j < n_sub; j++) {
                    const uint8_t mm = (uint8_t)(cl.tileDiplMult(j, da, db) + cl.tile_ic[j * 2 + g]), cc = cl.tile_c[j * cl.S + s];
                    n0 += mm == 0; c0 += mm == 0 ? cc : 0u;
                }
                if (n0) { atomicAdd(sh_stat + 2 * s, (unsigned long long)n0); atomicAdd(sh_stat + 2 * s + 1, (unsigned long long)c0); }
            }
            tick(6);
            for (uint32_t j = 0; j < cl.S * cl.Dall; j++) cl.ucache[j] = nan;  // clearGenotyperCache
            tick(7);
        }
        if (n_fill_tasks) grid_barrier(gb);
        lap(0);
        // phase B: sel[0 .. n_big): one WARP each (lane 0 samples from the filled cache; counts and cache clear by all lanes)
        for (uint32_t i = tid >> 5; i < n_big; i += nthreads >> 5) {
            const uint32_t lane = tid & 31u;
            const unsigned long long t_in = BTG_NOISE_TIMING && ns.phase_ns ? global_timer_ns() : 0;
            Cl cl;
            cl.bind(du, sel[i]);
            if (lane == 0) noise_iteration_thread(cl, du, T, o, joint && it > o.gibbs_burn_in);
            if (BTG_NOISE_TIMING && ns.phase_ns && lane == 0) { const unsigned long long v = ((global_timer_ns() - t_in) << 32) | cl.c; atomicMax(ns.phase_ns + 5, v); atomicMax(ns.phase_ns + 8 + 3 * (it - 1) + 1, v); }
            __syncwarp();
            const uint32_t n_sub = cl.misc[kNSub];
            for (uint32_t s = 0; s < cl.S; s++) {  // getNoiseCounts
                const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16;
                uint32_t n0 = 0, c0 = 0;
                const uint32_t g = du.sample_gender[s];
                for (uint32_t j = lane; j < n_sub; j += 32)
                    if ((uint8_t)(cl.tileDiplMult(j, da, db) + cl.tile_ic[j * 2 + g]) == 0) { n0++; c0 += cl.tile_c[j * cl.S + s]; }
                if (n0) { atomicAdd(sh_stat + 2 * s, (unsigned long long)n0); atomicAdd(sh_stat + 2 * s + 1, (unsigned long long)c0); }
            }
            for (uint32_t j = lane; j < cl.S * cl.Dall; j += 32) cl.ucache[j] = nan;  // clearGenotyperCache
            __syncwarp();
        }
        __syncthreads();
        if (threadIdx.x < 2 * du.S && sh_stat[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh_stat[threadIdx.x]);
        grid_barrier(gb);
        lap(1);
        if (blockIdx.x == 0) {
            if (px.world > 1) {  // sharded unit: add up the ranks' statistics over peer memory (comm.cuh) before the draw
                if (threadIdx.x < 2 * du.S) sh_tot[threadIdx.x] = hist[threadIdx.x];
                peer_allreduce_block(px, px.seq0 + it, sh_tot, 2 * du.S);
                if (threadIdx.x < 2 * du.S) hist[threadIdx.x] = sh_tot[threadIdx.x];
                __syncthreads();
            }
            noise_update_block(ns, du.S, prior_shape, prior_scale, o.random_seed, 1, o.gibbs_burn_in < it, (double)chain, (double)it, 1, sh_rates);
        }
        lap(2);
        grid_barrier(gb);
        lap(3);
    }
#if BTG_NOISE_TIMING
    if (ns.phase_ns && sub_acc[8])
        for (int k = 0; k < 9; k++) atomicAdd(ns.phase_ns + 8 + 3 * (size_t)iters + k, sub_acc[k]);
#endif
}

__global__ void k_noise_rng_init(uint32_t *rng, uint32_t seed, uint32_t chain = 0) {
    Philox r;
    r.init(seed, (uint64_t)-1, 0, kRngNoise, chain);
    r.save(rng, 0);
}

// estimateNoise, end: mean of the post-burn-in rates of all chains (added up in chain order) -> setNoiseRates
// (InferenceEngine.cpp:259-264), Poisson rows rebuilt, final trace row "0 0"
__global__ void k_noise_finish(const double *chain_means /* [n_chains][S] sums */, uint32_t n_chains, uint32_t S, double div, double *rates, double *noise_table,
                               double *trace_row) {
    __shared__ double sh[BTG_MAX_SAMPLES];
    if (threadIdx.x < S) {
        double acc = 0;
        for (uint32_t b = 0; b < n_chains; b++) acc += chain_means[(size_t)b * S + threadIdx.x];
        const double r = acc / div;
        rates[threadIdx.x] = r;
        sh[threadIdx.x] = r;
        if (trace_row) { trace_row[0] = 0; trace_row[1] = 0; trace_row[2 + threadIdx.x] = r; }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < S * 256u; i += blockDim.x) noise_table[i] = noiseCountLogPmf(sh[i >> 8], i & 255u);
}

template <class T> T *upload(const T *h, size_t n, bool &ok) {
    T *d = nullptr;
    if (cudaMalloc(&d, (n ? n : 1) * sizeof(T)) != cudaSuccess) { ok = false; return nullptr; }
    if (n && cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
    return d;
}

}  // namespace

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

void btg_unit_free_result(btg_unit *u);

extern "C" {

// ---- count distribution ---------------------------------------------------------------------
btg_count_dist *btg_count_dist_create(uint32_t S, const double *nb_p, const double *nb_size, float prior_shape, float prior_scale) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (S == 0 || S > BTG_MAX_SAMPLES || !nb_p || !nb_size) { set_error("bad count distribution arguments"); return nullptr; }
    for (uint32_t s = 0; s < S; s++)
        if (!(nb_p[s] > 0 && nb_p[s] < 1 && nb_size[s] > 0)) { set_error("negative binomial parameters out of range for sample %u", s); return nullptr; }
    auto *cd = new btg_count_dist();
    cd->S = S;
    cd->prior_shape = prior_shape;
    cd->prior_scale = prior_scale;
    cd->h_p.assign(nb_p, nb_p + S);
    cd->h_size.assign(nb_size, nb_size + S);
    bool ok = true;
    cd->p = upload(nb_p, S, ok);
    cd->size = upload(nb_size, S, ok);
    std::vector<double> ones(S, 1.0);
    cd->rates = upload(ones.data(), S, ok);
    ok = ok && cudaMalloc(&cd->genomic, (size_t)S * 65536 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&cd->noise, (size_t)S * 256 * sizeof(double)) == cudaSuccess;
    if (!ok) { set_error("count distribution allocation failed"); btg_count_dist_free(cd); return nullptr; }
    auto s = ctx().stream;
    k_genomic_table<<<(S * 65536 + 255) / 256, 256, 0, s>>>(cd->p, cd->size, S, cd->genomic);
    BTG_LAUNCHED();
    k_noise_table<<<(S * 256 + 255) / 256, 256, 0, s>>>(cd->rates, S, cd->noise);
    BTG_LAUNCHED();
    if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("count table kernels failed: %s", cudaGetErrorString(cudaGetLastError())); btg_count_dist_free(cd); return nullptr; }
    return cd;
}

void btg_nb_moments_to_parameters(double mean, double var, uint32_t multiplicity, double *p_out, double *size_out) {
    const double max_p = 0.99;  // NegativeBinomialDistribution.cpp:38,68-79
    if (max_p < (mean / var)) var = mean / max_p;
    if (p_out) *p_out = mean / var;
    if (size_out) *size_out = std::pow(mean, 2) / (var - mean) / multiplicity;  // CountDistribution.cpp:115-116
}

int btg_count_dist_set_noise_rates(btg_count_dist *cd, const double *rates) {
    BTG_REQUIRE_INIT();
    if (!cd || !rates) { set_error("null argument"); return BTG_EINVAL; }
    auto s = ctx().stream;
    BTG_CUDA(cudaMemcpyAsync(cd->rates, rates, cd->S * sizeof(double), cudaMemcpyHostToDevice, s));
    k_noise_table<<<(cd->S * 256 + 255) / 256, 256, 0, s>>>(cd->rates, cd->S, cd->noise);
    BTG_LAUNCHED();
    BTG_CUDA(cudaStreamSynchronize(s));
    return BTG_OK;
}

int btg_count_dist_get_noise_rates(const btg_count_dist *cd, double *out) {
    BTG_REQUIRE_INIT();
    if (!cd || !out) { set_error("null argument"); return BTG_EINVAL; }
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    BTG_CUDA(cudaMemcpy(out, cd->rates, cd->S * sizeof(double), cudaMemcpyDeviceToHost));
    return BTG_OK;
}

int btg_count_dist_tables(const btg_count_dist *cd, double *genomic_out, double *noise_out) {
    BTG_REQUIRE_INIT();
    if (!cd) { set_error("null argument"); return BTG_EINVAL; }
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    if (genomic_out) BTG_CUDA(cudaMemcpy(genomic_out, cd->genomic, (size_t)cd->S * 65536 * sizeof(double), cudaMemcpyDeviceToHost));
    if (noise_out) BTG_CUDA(cudaMemcpy(noise_out, cd->noise, (size_t)cd->S * 256 * sizeof(double), cudaMemcpyDeviceToHost));
    return BTG_OK;
}

void btg_count_dist_free(btg_count_dist *cd) {
    if (!cd) return;
    cudaFree(cd->p); cudaFree(cd->size); cudaFree(cd->rates); cudaFree(cd->genomic); cudaFree(cd->noise);
    delete cd;
}

// ---- unit -------------------------------------------------------------------------------------
btg_unit *btg_unit_upload(const btg_unit_desc *d) { return btg_unit_upload_dev(d, nullptr, 0, 0); }

btg_unit *btg_unit_upload_dev(const btg_unit_desc *d, const btg_unit_desc *dev, uint64_t n_vh_dev, uint64_t n_vh_bits_dev) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (!d || d->n_samples == 0 || d->n_samples > BTG_MAX_SAMPLES) { set_error("bad unit descriptor"); return nullptr; }
    // row-level arrays may come from the device (dev->field != NULL): device-to-device copy instead of a host round trip
    auto from = [&](auto host_ptr, auto dev_ptr, size_t n, bool &ok_flag) {
        using T = std::remove_cv_t<std::remove_pointer_t<decltype(host_ptr)>>;
        if (!dev_ptr) return upload(host_ptr, n, ok_flag);
        T *p = nullptr;
        if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) { ok_flag = false; return (T *)nullptr; }
        if (n && cudaMemcpyAsync(p, dev_ptr, n * sizeof(T), cudaMemcpyDeviceToDevice, ctx().stream) != cudaSuccess) ok_flag = false;
        return p;
    };
#define BTG_DEVF(f) (dev ? dev->f : nullptr)
    const uint32_t S = d->n_samples, G = d->n_groups, C = d->n_clusters;
    auto *u = new btg_unit();
    bool ok = true;
    auto keep = [&](auto *p) { u->allocs.push_back((void *)p); return p; };
    const bool up_timing = getenv("BTG_UPLOAD_TIMING") != nullptr;
    auto up_t0 = std::chrono::steady_clock::now();
    auto up_lap = [&](const char *what) {
        if (!up_timing) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[btgpu] unit upload: %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - up_t0).count());
        up_t0 = now;
    };
    const uint64_t rows = d->cl_kmer_off[C], nvar = d->cl_var_off[C], n_vh = BTG_DEVF(kmer_vh_off) ? n_vh_dev : d->kmer_vh_off[rows];
    DevUnit &du = u->du;
    du.S = S; du.G = G; du.C = C;
    du.sample_gender = keep(upload(d->sample_gender, S, ok));
    du.group_ploidy = keep(upload(d->group_ploidy, (size_t)G * S, ok));
    du.group_cluster_off = keep(upload(d->group_cluster_off, G + 1, ok));
    du.cluster_idx = keep(upload(d->cluster_idx, C, ok));
    du.cl_nhap = keep(upload(d->cl_nhap, C, ok));
    du.cl_kmer_off = keep(upload(d->cl_kmer_off, C + 1, ok));
    du.cl_var_off = keep(upload(d->cl_var_off, C + 1, ok));
    du.cl_mult_off = keep(upload(d->cl_mult_off, C + 1, ok));
    du.mult = keep(from(d->mult, BTG_DEVF(mult), d->cl_mult_off[C], ok));
    du.k_has_counts = keep(from(d->k_has_counts, BTG_DEVF(k_has_counts), rows, ok));
    du.k_counts = keep(from(d->k_counts, BTG_DEVF(k_counts), rows * S, ok));
    du.k_ic = keep(from(d->k_ic, BTG_DEVF(k_ic), rows * 2, ok));
    du.cl_uniq_off = keep(upload(d->cl_uniq_off, C + 1, ok));
    du.uniq_idx = keep(from(d->uniq_idx, BTG_DEVF(uniq_idx), d->cl_uniq_off[C], ok));
    du.kmer_vh_off = keep(from(d->kmer_vh_off, BTG_DEVF(kmer_vh_off), rows + 1, ok));
    du.vh_var = keep(from(d->vh_var, BTG_DEVF(vh_var), n_vh, ok));
    du.vh_bits_off = keep(from(d->vh_bits_off, BTG_DEVF(vh_bits_off), n_vh + 1, ok));
    du.vh_bits = keep(from(d->vh_bits, BTG_DEVF(vh_bits), BTG_DEVF(vh_bits_off) ? n_vh_bits_dev : d->vh_bits_off[n_vh], ok));
    du.cl_hapvar_off = keep(upload(d->cl_hapvar_off, C + 1, ok));
    du.hap_alleles = keep(from(d->hap_alleles, BTG_DEVF(hap_alleles), d->cl_hapvar_off[C], ok));
    du.var_nalleles = keep(upload(d->var_nalleles, nvar, ok));
    du.var_dep = keep(upload(d->var_dep, nvar, ok));
    // nested groups and multicluster k-mers
    {
        const uint64_t n_multi = d->cl_multi_off[C];
        uint64_t n_hap = 0, n_shared = 0;
        std::vector<uint64_t> hap_start(C + 1, 0);
        for (uint32_t c = 0; c < C; c++) hap_start[c + 1] = hap_start[c] + d->cl_nhap[c];
        n_hap = hap_start[C];
        if (n_multi && (BTG_DEVF(k_shared) || BTG_DEVF(k_has_counts))) {
            set_error("units with multicluster k-mers must pass k_shared and k_has_counts as host arrays (they are validated on the host)");
            ok = false;
        }
        for (uint32_t c = 0; c < C && ok; c++)
            for (uint64_t i = d->cl_multi_off[c]; i < d->cl_multi_off[c + 1]; i++) {
                const uint64_t r = d->cl_kmer_off[c] + d->multi_idx[i];
                if (d->k_shared[r] == 0xFFFFFFFFu || !d->k_has_counts[r]) { set_error("cluster %u: multicluster k-mer row %llu has no shared count record (k_shared)", c, (unsigned long long)r); ok = false; break; }
                n_shared = std::max<uint64_t>(n_shared, (uint64_t)d->k_shared[r] + 1);
            }
        du.k_shared = keep(from(d->k_shared, BTG_DEVF(k_shared), rows, ok));
        du.cl_multi_off = keep(upload(d->cl_multi_off, C + 1, ok));
        du.multi_idx = keep(upload(d->multi_idx, n_multi, ok));
        du.hap_start = keep(upload(hap_start.data(), C + 1, ok));
        du.hap_nested_off = keep(upload(d->hap_nested_off, n_hap + 1, ok));
        du.hap_nested = keep(upload(d->hap_nested, d->hap_nested_off[n_hap], ok));
        du.cl_dep_off = keep(upload(d->cl_dep_off, C + 1, ok));
        const uint64_t n_dep = d->cl_dep_off[C];
        du.dep_cluster = keep(upload(d->dep_cluster, n_dep, ok));
        du.dep_var_off = keep(upload(d->dep_var_off, n_dep + 1, ok));
        du.dep_var = keep(upload(d->dep_var, d->dep_var_off[n_dep], ok));
        du.group_src_off = keep(upload(d->group_src_off, G + 1, ok));
        du.group_src = keep(upload(d->group_src, d->group_src_off[G], ok));
        // out-edges as a CSR over clusters, each vertex's targets in the order given (VariantClusterGroup.cpp:94-104)
        std::vector<uint64_t> cl_edge_off(C + 1, 0);
        const uint64_t n_edges = d->group_edge_off[G];
        std::vector<uint32_t> edge_dst(n_edges);
        std::vector<uint32_t> nested_groups, nest_slot(C, 0xFFFFFFFFu);
        uint32_t n_nest = 0;
        for (uint32_t g = 0; g < G && ok; g++) {
            const uint64_t c0 = d->group_cluster_off[g], n = d->group_cluster_off[g + 1] - c0;
            if (n == 0) { set_error("group %u has no cluster", g); ok = false; break; }
            for (uint64_t e = d->group_edge_off[g]; e < d->group_edge_off[g + 1]; e++) {
                if (d->group_edge_src[e] >= n || d->group_edge_dst[e] >= n) { set_error("group %u: edge %llu out of range", g, (unsigned long long)e); ok = false; break; }
                cl_edge_off[c0 + d->group_edge_src[e] + 1]++;
            }
            for (uint64_t e = d->group_src_off[g]; e < d->group_src_off[g + 1]; e++)
                if (d->group_src[e] >= n) { set_error("group %u: source vertex out of range", g); ok = false; break; }
            if (n > 1) {
                nested_groups.push_back(g);
                for (uint64_t c = c0; c < c0 + n; c++) nest_slot[c] = n_nest++;
                if (d->group_edge_off[g + 1] - d->group_edge_off[g] + (d->group_src_off[g + 1] - d->group_src_off[g]) != n) {
                    set_error("group %u: %llu clusters need a forest of %llu sources + edges", g, (unsigned long long)n, (unsigned long long)n); ok = false; break;
                }
            }
        }
        for (uint32_t c = 0; c < C; c++) cl_edge_off[c + 1] += cl_edge_off[c];
        if (ok) {
            std::vector<uint64_t> fill(cl_edge_off.begin(), cl_edge_off.end() - 1);
            for (uint32_t g = 0; g < G; g++) {
                const uint64_t c0 = d->group_cluster_off[g];
                for (uint64_t e = d->group_edge_off[g]; e < d->group_edge_off[g + 1]; e++) edge_dst[fill[c0 + d->group_edge_src[e]]++] = d->group_edge_dst[e];
            }
        }
        du.cl_edge_off = keep(upload(cl_edge_off.data(), C + 1, ok));
        du.edge_dst = keep(upload(edge_dst.data(), n_edges, ok));
        du.nested_groups = keep(upload(nested_groups.data(), nested_groups.size(), ok));
        du.n_nested_groups = (uint32_t)nested_groups.size();
        du.nest_slot = keep(upload(nest_slot.data(), C, ok));
        auto dmalloc = [&](auto *&dst, size_t n) {
            void *p = nullptr;
            if (cudaMalloc(&p, (n ? n : 1) * sizeof(*dst)) != cudaSuccess) { ok = false; dst = nullptr; return; }
            u->allocs.push_back(p);
            dst = static_cast<std::remove_reference_t<decltype(dst)>>(p);
        };
        dmalloc(du.src_mut, d->group_src_off[G]); dmalloc(du.edge_mut, n_edges);
        dmalloc(du.dfs_order, nested_groups.empty() ? 0 : C); dmalloc(du.dfs_stack, nested_groups.empty() ? 0 : C);
        dmalloc(du.shared_mult, n_shared * S);
        dmalloc(du.nest_pl, (size_t)n_nest * S); dmalloc(du.nest_k, (size_t)n_nest * S);
        dmalloc(du.nest_n, (size_t)n_nest * S * 2); dmalloc(du.nest_f, (size_t)n_nest * S * 4);
        u->n_shared = n_shared;
    }
    // result offsets
    u->n_variants = nvar;
    u->h_valt_off.assign(nvar + 1, 0);
    u->h_allele_off.assign(nvar + 1, 0);
    u->h_geno_off.assign(nvar + 1, 0);
    for (uint64_t v = 0; v < nvar; v++) {
        const uint64_t nA = d->var_nalleles[v];
        u->h_valt_off[v + 1] = u->h_valt_off[v] + nA;
        u->h_allele_off[v + 1] = u->h_allele_off[v] + S * nA;
        u->h_geno_off[v + 1] = u->h_geno_off[v] + S * nA * (nA + 1) / 2;
    }
    u->n_alleles_total = u->h_valt_off[nvar];
    du.valt_off = keep(upload(u->h_valt_off.data(), nvar + 1, ok));
    // arena layout + cost order
    u->h_layout.resize(C);
    u->h_fill_cost.assign(C, 0);
    u->h_nhap.assign(d->cl_nhap, d->cl_nhap + C);
    u->h_group_cluster_off.assign(d->group_cluster_off, d->group_cluster_off + G + 1);
    u->h_cl_var_off.assign(d->cl_var_off, d->cl_var_off + C + 1);
    struct Dims { uint32_t H, K, nv, nu, nal, Dall, nm; };
    up_lap("descriptor arrays -> device");
    std::vector<Dims> dims(C);
    std::vector<uint64_t> cost(C);
    for (uint32_t g = 0; g < G; g++) {
        for (uint64_t c = d->group_cluster_off[g]; c < d->group_cluster_off[g + 1]; c++) {
            const uint32_t H = d->cl_nhap[c];
            const uint32_t K = (uint32_t)(d->cl_kmer_off[c + 1] - d->cl_kmer_off[c]);
            const uint32_t nv = (uint32_t)(d->cl_var_off[c + 1] - d->cl_var_off[c]);
            const uint32_t nu = (uint32_t)(d->cl_uniq_off[c + 1] - d->cl_uniq_off[c]);
            const uint32_t nal = (uint32_t)(u->h_valt_off[d->cl_var_off[c + 1]] - u->h_valt_off[d->cl_var_off[c]]);
            if (H == 0 || H >= 0xFFFE) { set_error("cluster %llu: invalid number of haplotypes %u", (unsigned long long)c, H); ok = false; break; }
            ClusterLayout &L = u->h_layout[c];
            L.group = g;
            L.n_alleles = nal;
            L.Dall = (H + 1) * (H + 2) / 2;
            const uint32_t nm = (uint32_t)(d->cl_multi_off[c + 1] - d->cl_multi_off[c]);
            dims[c] = Dims{H, K, nv, nu, nal, L.Dall, nm};
            u->h_fill_cost[c] = (uint32_t)std::min<uint64_t>(0xFFFFFFFFu, (uint64_t)S * ((uint64_t)H * (H + 1) / 2) * (nu / 10 + 1));
            cost[c] = (uint64_t)S * ((uint64_t)H * (H + 1) / 2) * 8 + nu + (uint64_t)H * K / 16;
            u->max_h = std::max(u->max_h, H);
        }
    }
    std::vector<uint32_t> order(C);
    std::iota(order.begin(), order.end(), 0u);
    // clusters of single-cluster groups first (one thread each), then the clusters of nested groups (one thread per group)
    auto is_nested = [&](uint32_t c) { const uint32_t g = u->h_layout[c].group; return d->group_cluster_off[g + 1] - d->group_cluster_off[g] > 1; };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const bool na = is_nested(a), nb = is_nested(b);
        return na != nb ? nb : cost[a] > cost[b];
    });
    du.n_regular = 0;
    while (du.n_regular < C && !is_nested(order[du.n_regular])) du.n_regular++;
    // chain-split clusters (default mode): large clusters of single-cluster groups get kChainSplit arena positions; the extra
    // positions follow the C regular ones
    const uint32_t split_cost = getenv("BTG_SPLIT_COST") ? (uint32_t)strtoul(getenv("BTG_SPLIT_COST"), nullptr, 10) : kSplitFillCost;  // tests: 0 splits everything
    std::vector<uint32_t> split_cluster, split_of(C ? C : 1, NONE32);
    for (uint32_t i = 0; i < du.n_regular; i++)
        if (u->h_fill_cost[order[i]] > split_cost) { split_of[order[i]] = (uint32_t)split_cluster.size(); split_cluster.push_back(order[i]); }
    const uint32_t n_split = (uint32_t)split_cluster.size();
    std::vector<uint32_t> ext_cluster(order);  // cluster at every arena position
    std::vector<uint32_t> split_pos((size_t)n_split * kChainSplit + 1, NONE32);
    for (uint32_t j = 0; j < n_split; j++)
        for (uint32_t v = 1; v < kChainSplit; v++) { split_pos[(size_t)j * kChainSplit + v] = (uint32_t)ext_cluster.size(); ext_cluster.push_back(split_cluster[j]); }
    const uint32_t n_pos = (uint32_t)ext_cluster.size();
    // one arena slot per warp of the position order, sized by the largest cluster in it
    const uint32_t n_slots = (n_pos + 31) / 32;
    u->h_slots.assign(n_slots ? n_slots : 1, SlotLayout{});
    uint64_t f64_total = 0, u32_total = 0, u8_total = 0;
    for (uint32_t w = 0; w < n_slots; w++) {
        SlotLayout &SL = u->h_slots[w];
        for (uint32_t i = w * 32; i < std::min<uint64_t>(n_pos, (uint64_t)w * 32 + 32); i++) {
            const Dims &D = dims[ext_cluster[i]];
            if (i < C) u->h_layout[order[i]].pos = i;
            SL.H = std::max(SL.H, D.H); SL.K = std::max(SL.K, D.K); SL.nvar = std::max(SL.nvar, D.nv);
            SL.n_uniq = std::max(SL.n_uniq, D.nu); SL.n_alleles = std::max(SL.n_alleles, D.nal); SL.Dall = std::max(SL.Dall, D.Dall);
            SL.n_multi = std::max(SL.n_multi, D.nm);
        }
        const ArenaSizes a = arena_sizes(S, SL.H, SL.K, SL.nvar, SL.n_uniq, SL.n_alleles, SL.Dall, SL.n_multi);
        SL.f64_off = f64_total; SL.u32_off = u32_total; SL.u8_off = u8_total;
        f64_total += a.f64 * 32; u32_total += a.u32 * 32; u8_total += a.u8 * 32;
    }
    std::vector<uint64_t> tile_off(C ? C : 1, ~0ull);
    uint64_t tile_total = 0;
    for (uint32_t c = 0; c < C; c++)
        if (u->h_fill_cost[c] > (uint64_t)kBigFillCost * S) { tile_off[c] = tile_total; tile_total += ((uint64_t)dims[c].nu * (dims[c].H + S + 2) + 31) & ~31ull; }
    du.big_tile_off = keep(upload(tile_off.data(), C, ok));
    uint8_t *tile_pool = nullptr;
    ok = ok && cudaMalloc(&tile_pool, tile_total + 32) == cudaSuccess;
    keep(tile_pool);
    du.big_tile_pool = tile_pool;
    up_lap("host layout");
    for (uint32_t j = 0; j < n_split; j++) split_pos[(size_t)j * kChainSplit] = u->h_layout[split_cluster[j]].pos;
    du.n_split = n_split;
    du.split_cluster = keep(upload(split_cluster.data(), n_split, ok));
    du.split_pos = keep(upload(split_pos.data(), (size_t)n_split * kChainSplit, ok));
    du.split_of = keep(upload(split_of.data(), C, ok));
    du.layout = keep(upload(u->h_layout.data(), C, ok));
    du.slots = keep(upload(u->h_slots.data(), u->h_slots.size(), ok));
    du.order = keep(upload(order.data(), C, ok));
    double *f64_pool = nullptr; uint32_t *u32_pool = nullptr; uint8_t *u8_pool = nullptr; double *lg = nullptr;
    ok = ok && cudaMalloc(&f64_pool, (f64_total + 1) * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&u32_pool, (u32_total + 1) * sizeof(uint32_t)) == cudaSuccess;
    ok = ok && cudaMalloc(&u8_pool, u8_total + 8) == cudaSuccess;
    const uint32_t n_lg = u->max_h + 2 * S + 4;
    ok = ok && cudaMalloc(&lg, n_lg * sizeof(double)) == cudaSuccess;
    keep(f64_pool); keep(u32_pool); keep(u8_pool); keep(lg);
    u->f64_total = f64_total; u->u32_total = u32_total; u->u8_total = u8_total; u->tile_total = tile_total;
    du.f64_pool = f64_pool; du.u32_pool = u32_pool; du.u8_pool = u8_pool; du.lgamma_int = lg;
    up_lap("arena allocation");
    if (ok) {
        k_lgamma_int<<<(n_lg + 127) / 128, 128, 0, ctx().stream>>>(lg, n_lg);
        BTG_LAUNCHED();
        ok = cudaStreamSynchronize(ctx().stream) == cudaSuccess;
    }
    if (!ok) {
        if (!*btg_last_error()) set_error("unit upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
        btg_unit_free(u);
        return nullptr;
    }
    return u;
}

void btg_unit_free(btg_unit *u) {
    if (!u) return;
    cudaStreamSynchronize(ctx().stream);
    for (void *p : u->allocs) cudaFree(p);
    btg_unit_free_result(u);
    delete u;
}

}  // extern "C"

namespace {
// device-side result arrays owned by the unit (allocated on first use)
struct DevResult {
    ResultView R{};
    std::vector<void *> allocs;
    uint64_t nv = 0, nall = 0, ngen = 0, nalt = 0;
    ~DevResult() { for (void *p : allocs) cudaFree(p); }
};

DevResult *unit_result(btg_unit *u) {
    if (u->res_view) return static_cast<DevResult *>(u->res_view);
    auto *dr = new DevResult();
    const uint32_t S = u->du.S;
    const uint64_t nv = u->n_variants, nall = u->h_allele_off[nv], ngen = u->h_geno_off[nv], nalt = u->h_valt_off[nv];
    dr->nv = nv; dr->nall = nall; dr->ngen = ngen; dr->nalt = nalt;
    bool ok = true;
    auto mk = [&](auto *&dst, size_t n) {
        void *d = nullptr;
        if (cudaMalloc(&d, (n ? n : 1) * sizeof(*dst)) != cudaSuccess) { ok = false; return; }
        dr->allocs.push_back(d);
        dst = static_cast<std::remove_reference_t<decltype(dst)>>(d);
    };
    dr->R.allele_off = upload(u->h_allele_off.data(), nv + 1, ok); dr->allocs.push_back((void *)dr->R.allele_off);
    dr->R.geno_off = upload(u->h_geno_off.data(), nv + 1, ok); dr->allocs.push_back((void *)dr->R.geno_off);
    dr->R.valt_off = u->du.valt_off;
    mk(dr->R.gt, nv * S * 2); mk(dr->R.gq, nv * S); mk(dr->R.gpp, ngen); mk(dr->R.app, nall);
    mk(dr->R.nak, nall); mk(dr->R.fak, nall); mk(dr->R.mac, nall); mk(dr->R.saf, nall); mk(dr->R.ploidy, nv * S);
    mk(dr->R.an, nv); mk(dr->R.ac, nalt); mk(dr->R.af, nalt); mk(dr->R.acp, nalt); mk(dr->R.anc, nalt); mk(dr->R.hc, nv);
    if (!ok) { delete dr; return nullptr; }
    u->res_view = dr;
    return dr;
}
}  // namespace

extern "C" {

int btg_estimate_genotypes_async(btg_unit *u, const btg_count_dist *cd, const btg_gibbs_opts *opts, void *stream) {
    BTG_REQUIRE_INIT();
    if (!u || !cd || !opts) { set_error("null argument"); return BTG_EINVAL; }
    if (cd->S != u->du.S) { set_error("count distribution has %u samples, unit has %u", cd->S, u->du.S); return BTG_EINVAL; }
    DevResult *dr = unit_result(u);
    if (!dr) { set_error("result allocation failed"); return BTG_ENOMEM; }
    Tables T{cd->genomic, cd->noise};
    if (u->du.n_nested_groups) {
        // KmerCounts::multiplicities start at zero in a fresh run (KmerCounts.hpp:100)
        BTG_CUDA(cudaMemsetAsync(u->du.shared_mult, 0, (size_t)u->n_shared * u->du.S, pick_stream(stream)));
        k_estimate_genotypes_nested<<<(u->du.n_nested_groups + 63) / 64, 64, 0, pick_stream(stream)>>>(u->du, T, *opts, dr->R);
        BTG_LAUNCHED();
        BTG_CUDA(cudaGetLastError());
    }
    if (u->du.n_regular) {
        static const int occ = getenv("BTG_GIBBS_OCC") ? atoi(getenv("BTG_GIBBS_OCC")) : 8;
        const int reconverge = getenv("BTG_GIBBS_SYNC") ? atoi(getenv("BTG_GIBBS_SYNC")) : 1;
        const unsigned grid = (u->du.n_regular + u->du.n_split * kChainSplit + 63) / 64;
        unsigned long long *dbg = nullptr;
        if (getenv("BTG_GIBBS_TIMING")) { cudaMalloc(&dbg, 32); cudaMemset(dbg, 0, 32); }
        if (occ >= 16) k_estimate_genotypes<16><<<grid, 64, 0, pick_stream(stream)>>>(u->du, T, *opts, dr->R, reconverge, dbg);
        else if (occ >= 12) k_estimate_genotypes<12><<<grid, 64, 0, pick_stream(stream)>>>(u->du, T, *opts, dr->R, reconverge, dbg);
        else k_estimate_genotypes<8><<<grid, 64, 0, pick_stream(stream)>>>(u->du, T, *opts, dr->R, reconverge, dbg);
        BTG_LAUNCHED();
        BTG_CUDA(cudaGetLastError());
        if (u->du.n_split) {
            k_merge_split<<<(u->du.n_split + 63) / 64, 64, 0, pick_stream(stream)>>>(u->du, *opts, dr->R);
            BTG_LAUNCHED();
            BTG_CUDA(cudaGetLastError());
        }
        if (dbg) {
            unsigned long long h[4];
            cudaStreamSynchronize(pick_stream(stream));
            cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost);
            cudaFree(dbg);
            const uint32_t c = (uint32_t)(h[0] & 0xFFFFFFFFu);
            fprintf(stderr, "[btgpu] k_estimate_genotypes: slowest thread %.1f ms (cluster %u, H %u, fill cost %u); mean thread %.2f ms over %u clusters; clusters with H > 4 hold %.1f %% of the thread time\n",
                    (h[0] >> 32) / 1e6, c, c < u->du.C ? u->h_nhap[c] : 0, c < u->du.C ? u->h_fill_cost[c] : 0, h[1] / 1e6 / std::max(1u, u->du.n_regular), u->du.n_regular, 100.0 * h[2] / std::max<unsigned long long>(1, h[1]));
        }
    }
    return BTG_OK;
}

int btg_unit_download_result(btg_unit *u, btg_genotype_result *out, void *stream) {
    BTG_REQUIRE_INIT();
    if (!u || !out) { set_error("null argument"); return BTG_EINVAL; }
    if (out->n_variants != u->n_variants) { set_error("result sized for %llu variants, unit has %llu", (unsigned long long)out->n_variants, (unsigned long long)u->n_variants); return BTG_EINVAL; }
    DevResult *dr = unit_result(u);
    if (!dr) { set_error("result allocation failed"); return BTG_ENOMEM; }
    const uint32_t S = u->du.S;
    auto s = pick_stream(stream);
    const ResultView &R = dr->R;
#define BTG_DL(field, n) BTG_CUDA(cudaMemcpyAsync(out->field, R.field, (n) * sizeof(*R.field), cudaMemcpyDeviceToHost, s))
    BTG_DL(gt, dr->nv * S * 2); BTG_DL(gq, dr->nv * S); BTG_DL(gpp, dr->ngen); BTG_DL(app, dr->nall);
    BTG_DL(nak, dr->nall); BTG_DL(fak, dr->nall); BTG_DL(mac, dr->nall); BTG_DL(saf, dr->nall); BTG_DL(ploidy, dr->nv * S);
    BTG_DL(an, dr->nv); BTG_DL(ac, dr->nalt); BTG_DL(af, dr->nalt); BTG_DL(acp, dr->nalt); BTG_DL(anc, dr->nalt); BTG_DL(hc, dr->nv);
#undef BTG_DL
    BTG_CUDA(cudaStreamSynchronize(s));
    return BTG_OK;
}

int btg_estimate_genotypes(btg_unit *u, const btg_count_dist *cd, const btg_gibbs_opts *opts, btg_genotype_result *out) {
    if (!out) { set_error("null argument"); return BTG_EINVAL; }
    int rc = btg_estimate_genotypes_async(u, cd, opts, nullptr);
    if (rc != BTG_OK) return rc;
    return btg_unit_download_result(u, out, nullptr);
}

}  // extern "C"
void btg_unit_free_result(btg_unit *u) {
    if (u->res_view) { delete static_cast<DevResult *>(u->res_view); u->res_view = nullptr; }
}
extern "C" {

int btg_unit_cluster_tally(const btg_unit *u, uint32_t cluster, uint32_t *tally_out, uint64_t n) {
    BTG_REQUIRE_INIT();
    if (!u || cluster >= u->du.C || !tally_out) { set_error("bad argument"); return BTG_EINVAL; }
    const ClusterLayout &L = u->h_layout[cluster];
    const uint32_t H = u->h_nhap[cluster], S = u->du.S;
    const uint64_t need = (uint64_t)L.Dall * S;
    if (n < need) { set_error("tally buffer too small"); return BTG_EINVAL; }
    // tally sits after obs[H], uniq[2*n_uniq], cnt[H*nvar] in the slot's u32 arrays (see Cl::bind), lane-interleaved
    (void)H;
    const SlotLayout &SL = u->h_slots[L.pos >> 5];
    const uint64_t off = SL.u32_off + (L.pos & 31u) + ((uint64_t)SL.H + 2ull * SL.n_uniq + (uint64_t)SL.H * SL.nvar) * 32;
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    BTG_CUDA(cudaMemcpy2D(tally_out, sizeof(uint32_t), u->du.u32_pool + off, 32 * sizeof(uint32_t), sizeof(uint32_t), need, cudaMemcpyDeviceToHost));
    return BTG_OK;
}

static int noise_chains(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out, int joint);
static int estimate_noise_concurrent(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out);

int btg_estimate_noise(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, double *trace_out) {
    return estimate_noise_concurrent(u, cd, opts, nullptr, trace_out);
}

int btg_estimate_noise_sharded(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out) {
    return estimate_noise_concurrent(u, cd, opts, sh, trace_out);
}

int btg_estimate_noise_and_genotypes(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, btg_genotype_result *out, double *trace_out) {
    return btg_estimate_noise_and_genotypes_sharded(u, cd, opts, nullptr, out, trace_out);
}

int btg_estimate_noise_and_genotypes_sharded(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, btg_genotype_result *out,
                                             double *trace_out) {
    if (!out) { set_error("null argument"); return BTG_EINVAL; }
    int rc = noise_chains(u, cd, opts, sh, trace_out, 1);
    if (rc != BTG_OK) return rc;
    DevResult *dr = unit_result(u);
    if (!dr) { set_error("result allocation failed"); return BTG_ENOMEM; }
    if (u->du.C) {
        k_summarise<<<(u->du.C + 63) / 64, 64, 0, ctx().stream>>>(u->du, *opts, dr->R);
        BTG_LAUNCHED();
        BTG_CUDA(cudaGetLastError());
    }
    return btg_unit_download_result(u, out, nullptr);
}


// the engine's own stream (InferenceEngine.cpp:174) lives on the host: Fisher-Yates with the same Philox recipe as the device
struct HostEnginePhilox {
    uint32_t key[2], ctr[4], buf[4]; int pos;
    void init(uint32_t seed) { key[0] = seed; key[1] = 0; ctr[0] = ctr[1] = ctr[2] = 0; ctr[3] = kRngEngine; pos = 4; }
    uint32_t next() {
        if (pos == 4) {
            uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
            for (int r = 0; r < 10; r++) {
                const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
                c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
                k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
            }
            for (int i = 0; i < 4; i++) buf[i] = c[i];
            if (++ctr[0] == 0) ++ctr[1];
            pos = 0;
        }
        return buf[pos++];
    }
    uint32_t uniform_int(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};

// shadow arenas for chains that run at the same time: shadow 0 is the unit's own arena
static uint32_t ensure_shadows(btg_unit *u, uint32_t want) {
    if (u->shadow_du.empty()) u->shadow_du.push_back(u->du);
    while (u->shadow_du.size() < want) {
        double *f = nullptr; uint32_t *w = nullptr; uint8_t *b = nullptr, *t = nullptr;
        const bool ok = cudaMalloc(&f, (u->f64_total + 1) * sizeof(double)) == cudaSuccess && cudaMalloc(&w, (u->u32_total + 1) * sizeof(uint32_t)) == cudaSuccess &&
                        cudaMalloc(&b, u->u8_total + 8) == cudaSuccess && cudaMalloc(&t, u->tile_total + 32) == cudaSuccess;
        if (!ok) { cudaFree(f); cudaFree(w); cudaFree(b); cudaFree(t); cudaGetLastError(); break; }  // fewer concurrent chains
        for (void *p : {(void *)f, (void *)w, (void *)b, (void *)t}) u->allocs.push_back(p);
        DevUnit d = u->du;
        d.f64_pool = f; d.u32_pool = w; d.u8_pool = b; d.big_tile_pool = t;
        u->shadow_du.push_back(d);
    }
    return (uint32_t)std::min<size_t>(want, u->shadow_du.size());
}

// InferenceEngine::estimateNoise (InferenceEngine.cpp:135-276).  The chains are independent under this library's stream contract
// (fresh genotypers per chain as in the reference, InferenceEngine.cpp:240-251; the noise rates of chain c come from their own
// stream, kind 4 / chain c + 1, starting with the prior draw), so several chains run AT THE SAME TIME: each is one persistent
// cooperative k_noise_chain on its own CUDA stream, arena and noise state, with a share of the SMs (BTG_NOISE_CONCURRENCY = K).
// Measured (profiles/r1_noise_chain_phases.txt): K = 4 or 8 chains side by side take as long as one after the other — an iteration
// costs (clusters per thread) x (slowest lane's step), so a chain on 1/K of the SMs is K times slower — hence the default K = 1;
// the per-chain contract is what makes the results independent of K.
static int estimate_noise_concurrent(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out) {
    BTG_REQUIRE_INIT();
    if (!u || !cd || !opts) { set_error("null argument"); return BTG_EINVAL; }
    btg_comm *comm = sh ? sh->comm : nullptr;
    const uint32_t world = comm ? comm->world : 1;
    if (sh && (!sh->group_n_clusters || !sh->group_n_variants || opts->group_index_base + u->du.G > sh->n_groups_total)) {
        set_error("bad shard descriptor: this rank's groups [%llu, %llu) do not fit the %llu groups of the unit", (unsigned long long)opts->group_index_base,
                  (unsigned long long)(opts->group_index_base + u->du.G), (unsigned long long)(sh ? sh->n_groups_total : 0));
        return BTG_EINVAL;
    }
    if (comm && !comm->connected) { set_error("communicator is not connected (btg_comm_connect)"); return BTG_ESTATE; }
    if (cd->S != u->du.S) { set_error("count distribution / unit sample mismatch"); return BTG_EINVAL; }
    const uint32_t S = u->du.S, G = u->du.G, n_chains = opts->n_chains;
    const uint32_t iters = (uint32_t)opts->gibbs_burn_in + opts->gibbs_samples;
    const size_t row_len = 2 + S, trace_rows = (size_t)n_chains * (iters + 1) + 1;
    auto s0 = ctx().stream;
    const bool want_phases = getenv("BTG_NOISE_PHASES") && atoi(getenv("BTG_NOISE_PHASES"));
    // one mailbox per communicator: a sharded unit runs its chains one after the other (the exchange is inside the kernel)
    uint32_t K = world > 1 || want_phases ? 1u : (uint32_t)std::max(1, getenv("BTG_NOISE_CONCURRENCY") ? atoi(getenv("BTG_NOISE_CONCURRENCY")) : 1);
    K = ensure_shadows(u, std::min(K, std::max(1u, n_chains)));
    int rc = BTG_OK;
    std::vector<void *> tmp;
    auto dalloc = [&](size_t bytes) { void *p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) { rc = BTG_ENOMEM; return (void *)nullptr; } tmp.push_back(p); cudaMemsetAsync(p, 0, bytes ? bytes : 8, s0); return p; };
    // per-chain state, one allocation each: [n_chains][...]
    const size_t nc = std::max(1u, n_chains);
    auto *d_hist = (unsigned long long *)dalloc(nc * 2 * S * 8);
    auto *d_rates = (double *)dalloc(nc * S * 8);
    auto *d_tables = (double *)dalloc(nc * S * 256 * 8);
    auto *d_means = (double *)dalloc(nc * S * 8);
    auto *d_rng = (uint32_t *)dalloc(nc * 8 * 4);
    auto *d_rows = (uint32_t *)dalloc(nc * 4);
    auto *d_bar = (unsigned int *)dalloc(nc * 256);
    auto *d_trace = trace_out ? (double *)dalloc(trace_rows * row_len * 8) : nullptr;
    const uint32_t n_lg = 1024;
    auto *lg_tab = (double *)dalloc(n_lg * 8);
    unsigned long long *d_phase = want_phases ? (unsigned long long *)dalloc((8 + 3 * (size_t)iters + 16) * 8) : nullptr;
    std::vector<cudaStream_t> streams(K, nullptr);
    cudaEvent_t ev_ready = nullptr;
    std::vector<std::vector<uint32_t>> sels(n_chains), tasks(n_chains);
    std::vector<uint32_t> n_bigs(n_chains, 0);
    uint32_t *d_sel = nullptr, *d_tasks = nullptr;
    size_t sel_cap = 0, task_cap = 0;
    std::function<void(uint32_t)> select_chain;  // group selection of chain b (host; chains in order: the engine stream is sequential)
    if (rc == BTG_OK) {
        // ---- group selection of every chain (InferenceEngine.cpp:174-189), on the host, in chain order ----
        HostEnginePhilox engine;
        engine.init(opts->random_seed);
        const uint64_t base = sh ? opts->group_index_base : 0;
        std::vector<uint32_t> noise_groups;  // single-cluster groups of the WHOLE unit (InferenceEngine.cpp:144-151)
        if (sh) { for (uint64_t g = 0; g < sh->n_groups_total; g++) if (sh->group_n_clusters[g] == 1) noise_groups.push_back((uint32_t)g); }
        else { for (uint32_t g = 0; g < G; g++) if (u->h_group_cluster_off[g + 1] - u->h_group_cluster_off[g] == 1) noise_groups.push_back(g); }
        auto group_variants = [&](uint32_t g) {
            if (sh) return sh->group_n_variants[g];
            uint64_t n = 0;
            for (uint64_t c = u->h_group_cluster_off[g]; c < u->h_group_cluster_off[g + 1]; c++) n += u->h_cl_var_off[c + 1] - u->h_cl_var_off[c];
            return (uint32_t)n;
        };
        const uint32_t noise_variants_batch_size = 100000;  // InferenceEngine.cpp:50
        auto is_big = [&](uint32_t c) { return u->h_fill_cost[c] > (uint64_t)kBigFillCost * S; };  // the clusters with a dense tile: cost PER SAMPLE (a one-thread cluster walks its samples in turn)
        // upper bounds per chain: every local single-cluster group selected; every large cluster with its maximal number of fill tasks
        for (uint32_t g = 0; g < G; g++) {
            if (u->h_group_cluster_off[g + 1] - u->h_group_cluster_off[g] != 1) continue;
            const uint32_t c = (uint32_t)u->h_group_cluster_off[g];
            sel_cap++;
            if (is_big(c)) task_cap += 3 * (size_t)std::min<uint64_t>(64, std::max<uint64_t>(1, ((uint64_t)S * ((uint64_t)u->h_nhap[c] * (u->h_nhap[c] + 1) / 2) + 3) / 4));
        }
        d_sel = (uint32_t *)dalloc(std::max<size_t>(1, sel_cap) * nc * 4);
        d_tasks = (uint32_t *)dalloc(std::max<size_t>(1, task_cap) * nc * 4);
        select_chain = [&, base, noise_groups, engine, group_variants, is_big, noise_variants_batch_size](uint32_t b) mutable {
            uint32_t end = 0, nvv = 0;
            for (size_t i = noise_groups.size(); i > 1; i--) std::swap(noise_groups[i - 1], noise_groups[engine.uniform_int((uint32_t)i)]);
            while (nvv < noise_variants_batch_size && end < noise_groups.size()) { nvv += group_variants(noise_groups[end]); end++; }
            std::sort(noise_groups.begin(), noise_groups.begin() + end);
            auto &sel = sels[b];
            for (uint32_t i = 0; i < end; i++) {
                const uint64_t g = noise_groups[i];
                if (g >= base && g < base + G) sel.push_back((uint32_t)u->h_group_cluster_off[g - base]);  // this rank's share
            }
            // large clusters first (fill tasks + one warp each), then by position in the cost order (neighbours share arena slots)
            std::sort(sel.begin(), sel.end(), [&](uint32_t a, uint32_t c) {
                const bool ba = is_big(a), bc = is_big(c);
                return ba != bc ? ba : u->h_layout[a].pos < u->h_layout[c].pos;
            });
            uint32_t n_big = 0;
            while (n_big < sel.size() && is_big(sel[n_big])) n_big++;
            n_bigs[b] = n_big;
            for (uint32_t i = 0; i < n_big; i++) {  // a large cluster gets one warp per 32 cache entries (upper bound), at most 64
                const uint64_t H = u->h_nhap[sel[i]], entries = (uint64_t)S * (H * (H + 1) / 2);
                // entry-parallel fill: 32 entries per warp; term-parallel fill (>= 16 k-mers per entry expected): 4 entries per warp
                const bool by_terms = (u->h_fill_cost[sel[i]] / std::max<uint64_t>(1, entries)) >= 16;
                const uint32_t parts = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, by_terms ? (entries + 3) / 4 : (entries + 31) / 32));
                for (uint32_t p = 0; p < parts; p++) { tasks[b].push_back(i); tasks[b].push_back(p); tasks[b].push_back(parts); }
            }
        };
    }
    if (rc == BTG_OK) {
        for (auto &st : streams) if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) rc = BTG_ECUDA;
        if (cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming) != cudaSuccess) rc = BTG_ECUDA;
    }
    if (rc == BTG_OK) {
        k_lgamma_int<<<(n_lg + 127) / 128, 128, 0, s0>>>(lg_tab, n_lg);
        BTG_LAUNCHED();
        cudaEventRecord(ev_ready, s0);  // allocations zeroed, tables of cd complete
        PeerExchange px{};
        px.world = world; px.rank = comm ? comm->rank : 0;
        if (comm) {
            for (uint32_t r = 0; r < world; r++) px.mail[r] = comm->peers[r];
            px.error = comm->error;
            const char *tmo = getenv("BTG_PEER_TIMEOUT_MS");
            px.timeout_ns = (tmo ? strtoull(tmo, nullptr, 10) : 20000ull) * 1000000ull;
        }
        const uint32_t bs = 256;
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_noise_chain, bs, 0);
        const uint32_t capacity = (uint32_t)std::max(1, per_sm) * (uint32_t)ctx().sm_count;
        const uint32_t max_blocks = std::max(1u, (K > 1 ? capacity - capacity / 16 : capacity) / K);  // this chain's share of the SMs (a few block slots stay free)
        for (auto &st : streams) cudaStreamWaitEvent(st, ev_ready, 0);
        for (uint32_t b = 0; b < n_chains && rc == BTG_OK; b++) {
            select_chain(b);  // while the previous chains run on the device
            const uint32_t k = b % K;
            cudaStream_t st = streams[k];
            NoiseState ns{};
            ns.hist = (uint64_t *)(d_hist + (size_t)b * 2 * S);
            ns.rates = d_rates + (size_t)b * S;
            ns.noise_table = d_tables + (size_t)b * S * 256;
            ns.mean_rates = d_means + (size_t)b * S;
            ns.trace = d_trace ? d_trace + (size_t)b * (iters + 1) * row_len : nullptr;
            ns.rng = d_rng + (size_t)b * 8;
            ns.trace_row = d_rows + b;
            ns.lg = lg_tab; ns.n_lg = n_lg;
            ns.phase_ns = d_phase;
            GridBarrier gb{d_bar + (size_t)b * 64, d_bar + (size_t)b * 64 + 32};
            uint32_t *sel_b = d_sel + (size_t)b * std::max<size_t>(1, sel_cap), *tasks_b = d_tasks + (size_t)b * std::max<size_t>(1, task_cap);
            if (!sels[b].empty() && cudaMemcpyAsync(sel_b, sels[b].data(), sels[b].size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = BTG_ECUDA; break; }
            if (!tasks[b].empty() && cudaMemcpyAsync(tasks_b, tasks[b].data(), tasks[b].size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = BTG_ECUDA; break; }
            // this chain's noise stream, and its first draw: the rates of the prior (CountDistribution.cpp:62,163-171)
            k_noise_rng_init<<<1, 1, 0, st>>>(ns.rng, opts->random_seed, b + 1);
            BTG_LAUNCHED();
            NoiseState ns_quiet = ns;
            ns_quiet.trace = nullptr;
            k_noise_update<<<1, 256, 0, st>>>(ns_quiet, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 0, 0, 0, 0, 1);
            BTG_LAUNCHED();
            Tables T{cd->genomic, ns.noise_table};
            uint32_t n_sel = (uint32_t)sels[b].size(), n_big = n_bigs[b], n_tasks = (uint32_t)(tasks[b].size() / 3), chain_id = b + 1, iters_arg = iters;
            const uint32_t want_threads = std::max<uint32_t>({n_big * 32u, n_sel - n_big, n_tasks * 32u, 1u});
            const uint32_t grid = std::max(1u, std::min((want_threads + bs - 1) / bs, max_blocks));
            float ps = cd->prior_shape, pc = cd->prior_scale;
            btg_gibbs_opts o = *opts;
            int joint = 0;
            unsigned long long *hist = (unsigned long long *)ns.hist;
            if (comm) { px.seq0 = comm->seq; comm->seq += iters; }
            DevUnit du_k = u->shadow_du[k];
            void *args[] = {&du_k, &T, &o, &sel_b, &n_sel, &n_big, &chain_id, &iters_arg, &ns, &ps, &pc, &hist, &joint, &px, &tasks_b, &n_tasks, &gb};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_noise_chain, dim3(grid), dim3(bs), args, 0, st);
            BTG_LAUNCHED();
            if (e != cudaSuccess) { set_error("cooperative launch failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; break; }
        }
        // join the chain streams, then: mean of the post-burn-in rates -> setNoiseRates, final trace row "0 0"
        for (auto &st : streams) { cudaEvent_t ev; if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) { cudaEventRecord(ev, st); cudaStreamWaitEvent(s0, ev, 0); cudaEventDestroy(ev); } }
        if (rc == BTG_OK) {
            k_noise_finish<<<1, 256, 0, s0>>>(d_means, n_chains, S, (double)opts->gibbs_samples * n_chains, cd->rates, cd->noise,
                                              d_trace ? d_trace + (trace_rows - 1) * row_len : nullptr);
            BTG_LAUNCHED();
            if (trace_out) cudaMemcpyAsync(trace_out, d_trace, trace_rows * row_len * 8, cudaMemcpyDeviceToHost, s0);
        }
        cudaError_t e = cudaStreamSynchronize(s0);
        for (auto &st : streams) if (st) { cudaError_t e2 = cudaStreamSynchronize(st); if (e == cudaSuccess) e = e2; }
        if (e != cudaSuccess && rc == BTG_OK) { set_error("noise estimation failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; }
        if (rc == BTG_OK && comm && world > 1) {
            uint32_t flag = 0;
            cudaMemcpy(&flag, comm->error, 4, cudaMemcpyDeviceToHost);
            if (flag) { set_error("peer exchange timed out waiting for rank %u (a rank failed or left the lock-step)", flag - 1); cudaMemset(comm->error, 0, 4); rc = BTG_ECUDA; }
        }
        if (rc == BTG_OK) {
            std::vector<unsigned int> bar(nc * 64);
            cudaMemcpy(bar.data(), d_bar, bar.size() * 4, cudaMemcpyDeviceToHost);
            for (size_t b = 0; b < nc; b++) if (bar[b * 64 + 33]) { set_error("grid barrier of chain %zu timed out (grid not co-resident)", b); rc = BTG_ECUDA; break; }
        }
        if (rc == BTG_OK && d_phase) {
            unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            cudaMemcpy(ph, d_phase, sizeof ph, cudaMemcpyDeviceToHost);
            const double n_it = (double)n_chains * iters;
            fprintf(stderr, "[btgpu] noise chain phases, us per iteration (block 0, chains one after the other): fill+one-thread %.1f  large owners %.1f  exchange+update %.1f  release %.1f; construct+reset %.2f ms per chain\n",
                    ph[0] / n_it / 1e3, ph[1] / n_it / 1e3, ph[2] / n_it / 1e3, ph[3] / n_it / 1e3, ph[7] / 1e6 / std::max(1u, n_chains));
        }
    } else if (rc != BTG_OK && !*btg_last_error()) {
        set_error("noise estimation allocation failed");
    }
    cudaStreamSynchronize(s0);
    for (auto &st : streams) if (st) cudaStreamDestroy(st);
    if (ev_ready) cudaEventDestroy(ev_ready);
    for (void *p : tmp) cudaFree(p);
    return rc;
}

static int noise_chains(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out, int joint) {
    BTG_REQUIRE_INIT();
    if (!u || !cd || !opts) { set_error("null argument"); return BTG_EINVAL; }
    btg_comm *comm = sh ? sh->comm : nullptr;
    const uint32_t world = comm ? comm->world : 1;
    if (sh && (!sh->group_n_clusters || !sh->group_n_variants || opts->group_index_base + u->du.G > sh->n_groups_total)) {
        set_error("bad shard descriptor: this rank's groups [%llu, %llu) do not fit the %llu groups of the unit", (unsigned long long)opts->group_index_base,
                  (unsigned long long)(opts->group_index_base + u->du.G), (unsigned long long)(sh ? sh->n_groups_total : 0));
        return BTG_EINVAL;
    }
    if (comm && !comm->connected) { set_error("communicator is not connected (btg_comm_connect)"); return BTG_ESTATE; }
    if (cd->S != u->du.S) { set_error("count distribution / unit sample mismatch"); return BTG_EINVAL; }
    const uint32_t S = u->du.S, G = u->du.G;
    if (joint && u->du.n_nested_groups) { set_error("the joint noise-genotyping mode does not support nested variant-cluster groups in this build (%u groups)", u->du.n_nested_groups); return BTG_EINVAL; }
    const uint32_t iters = (uint32_t)opts->gibbs_burn_in + opts->gibbs_samples;
    const size_t trace_rows = (size_t)opts->n_chains * (iters + 1) + (joint ? 0 : 1);
    auto s = ctx().stream;
    NoiseState ns{};
    unsigned long long *hist = nullptr;
    uint32_t *d_sel = nullptr, *d_tasks = nullptr;
    size_t tasks_cap = 0;
    std::vector<void *> tmp;
    auto dalloc = [&](size_t bytes) { void *p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) return (void *)nullptr; tmp.push_back(p); cudaMemsetAsync(p, 0, bytes ? bytes : 8, s); return p; };
    hist = (unsigned long long *)dalloc((size_t)S * 2 * 8);
    ns.hist = (uint64_t *)hist;
    ns.rates = cd->rates;
    ns.noise_table = cd->noise;
    ns.mean_rates = (double *)dalloc(S * 8);
    ns.trace = trace_out ? (double *)dalloc(trace_rows * (2 + S) * 8) : nullptr;
    ns.rng = (uint32_t *)dalloc(8 * 4);
    GridBarrier gb{};
    gb.count = (unsigned int *)dalloc(256);   // count and generation in separate 128-byte lines
    gb.gen = gb.count ? gb.count + 32 : nullptr;
    const bool want_phases = getenv("BTG_NOISE_PHASES") && atoi(getenv("BTG_NOISE_PHASES"));
    ns.phase_ns = want_phases ? (unsigned long long *)dalloc((8 + 3 * (size_t)iters + 16) * 8) : nullptr;  // [4 phases, 3 maxima, spare][iteration][3 maxima]
    const uint32_t n_lg = 1024;
    double *lg_tab = (double *)dalloc(n_lg * 8);
    ns.lg = lg_tab; ns.n_lg = n_lg;
    ns.trace_row = (uint32_t *)dalloc(4);
    d_sel = (uint32_t *)dalloc((size_t)G * 4);
    int rc = BTG_OK;
    if (!hist || !ns.mean_rates || !ns.rng || !ns.trace_row || !d_sel || (trace_out && !ns.trace)) { set_error("noise estimation allocation failed"); rc = BTG_ENOMEM; }
    if (rc == BTG_OK) {
        // the engine's own stream (InferenceEngine.cpp:174) lives on the host: Fisher-Yates with the same Philox recipe
        struct HostPhilox {
            uint32_t key[2], ctr[4], buf[4]; int pos;
            void init(uint32_t seed) { key[0] = seed; key[1] = 0; ctr[0] = ctr[1] = ctr[2] = 0; ctr[3] = kRngEngine; pos = 4; }
            uint32_t next() {
                if (pos == 4) {
                    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
                    for (int r = 0; r < 10; r++) {
                        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
                        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
                        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
                        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
                    }
                    for (int i = 0; i < 4; i++) buf[i] = c[i];
                    if (++ctr[0] == 0) ++ctr[1];
                    pos = 0;
                }
                return buf[pos++];
            }
            uint32_t uniform_int(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
        } engine;
        engine.init(opts->random_seed);
        // single-cluster groups (InferenceEngine.cpp:144-151) of the WHOLE unit: with a shard descriptor every rank walks the
        // same global list with the same engine stream, so the selection does not depend on the sharding
        const uint64_t base = sh ? opts->group_index_base : 0;
        std::vector<uint32_t> noise_groups;
        if (sh) {
            for (uint64_t g = 0; g < sh->n_groups_total; g++) if (sh->group_n_clusters[g] == 1) noise_groups.push_back((uint32_t)g);
        } else {
            for (uint32_t g = 0; g < G; g++)
                if (u->h_group_cluster_off[g + 1] - u->h_group_cluster_off[g] == 1) noise_groups.push_back(g);
        }
        auto group_variants = [&](uint32_t g) {
            if (sh) return sh->group_n_variants[g];
            uint64_t n = 0;
            for (uint64_t c = u->h_group_cluster_off[g]; c < u->h_group_cluster_off[g + 1]; c++) n += u->h_cl_var_off[c + 1] - u->h_cl_var_off[c];
            return (uint32_t)n;
        };
        PeerExchange px{};
        px.world = world; px.rank = comm ? comm->rank : 0;
        if (comm) {
            for (uint32_t r = 0; r < world; r++) px.mail[r] = comm->peers[r];
            px.error = comm->error;
            const char *tmo = getenv("BTG_PEER_TIMEOUT_MS");
            px.timeout_ns = (tmo ? strtoull(tmo, nullptr, 10) : 20000ull) * 1000000ull;
        }
        Tables T{cd->genomic, cd->noise};
        k_noise_rng_init<<<1, 1, 0, s>>>(ns.rng, opts->random_seed);
        BTG_LAUNCHED();
        if (lg_tab) { k_lgamma_int<<<(n_lg + 127) / 128, 128, 0, s>>>(lg_tab, n_lg); BTG_LAUNCHED(); } else ns.lg = nullptr;
        NoiseState ns_quiet = ns;  // same state, no trace row
        ns_quiet.trace = nullptr;
        // CountDistribution ctor draws the initial rates (CountDistribution.cpp:62)
        k_noise_update<<<1, 256, 0, s>>>(ns_quiet, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 0, 0, 0, 0, 1);
        BTG_LAUNCHED();
        const uint32_t noise_variants_batch_size = 100000;  // InferenceEngine.cpp:50
        std::vector<uint32_t> sel, tasks;
        for (uint32_t chain = 0; chain < opts->n_chains && rc == BTG_OK; chain++) {
            uint32_t end = 0, nvv = 0;
            if (joint) {
                end = (uint32_t)noise_groups.size();  // every group, every chain (InferenceEngine.cpp:407-408)
            } else {
                for (size_t i = noise_groups.size(); i > 1; i--) std::swap(noise_groups[i - 1], noise_groups[engine.uniform_int((uint32_t)i)]);
                while (nvv < noise_variants_batch_size && end < noise_groups.size()) { nvv += group_variants(noise_groups[end]); end++; }
                std::sort(noise_groups.begin(), noise_groups.begin() + end);
            }
            sel.clear();
            for (uint32_t i = 0; i < end; i++) {
                const uint64_t g = noise_groups[i];
                if (g >= base && g < base + G) sel.push_back((uint32_t)u->h_group_cluster_off[g - base]);  // this rank's share
            }
            // large clusters first (one warp each in the chain kernel), then by position in the cost order (neighbours share arena slots)
            auto is_big = [&](uint32_t c) { return u->h_fill_cost[c] > (uint64_t)kBigFillCost * S; };  // the clusters with a dense tile: cost PER SAMPLE (a one-thread cluster walks its samples in turn)
            std::sort(sel.begin(), sel.end(), [&](uint32_t a, uint32_t b) {
                const bool ba = is_big(a), bb = is_big(b);
                return ba != bb ? ba : u->h_layout[a].pos < u->h_layout[b].pos;
            });
            uint32_t n_big = 0;
            while (n_big < sel.size() && is_big(sel[n_big])) n_big++;
            // fill tasks: a large cluster gets one warp per 32 cache entries (S x diplotypes, upper bound), at most 64
            tasks.clear();
            for (uint32_t i = 0; i < n_big; i++) {
                const uint64_t H = u->h_nhap[sel[i]], entries = (uint64_t)S * (H * (H + 1) / 2);
                // entry-parallel fill: 32 entries per warp; term-parallel fill (>= 16 k-mers per entry expected): 4 entries per warp
                const bool by_terms = (u->h_fill_cost[sel[i]] / std::max<uint64_t>(1, entries)) >= 16;
                const uint32_t parts = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, by_terms ? (entries + 3) / 4 : (entries + 31) / 32));
                for (uint32_t p = 0; p < parts; p++) { tasks.push_back(i); tasks.push_back(p); tasks.push_back(parts); }
            }
            if (tasks.size() > tasks_cap) {
                if (d_tasks) cudaFree(d_tasks);
                tasks_cap = tasks.size() * 2;
                if (cudaMalloc(&d_tasks, tasks_cap * 4) != cudaSuccess) { d_tasks = nullptr; tasks_cap = 0; rc = BTG_ENOMEM; set_error("fill task allocation failed"); break; }
            }
            if (!tasks.empty() && cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BTG_ECUDA; break; }
            if (cudaMemcpyAsync(d_sel, sel.data(), sel.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BTG_ECUDA; break; }
            cudaStreamSynchronize(s);  // sel is reused by the host next chain
            const uint32_t n_sel = (uint32_t)sel.size();
            if (n_sel || world > 1) {  // a rank with nothing selected still takes part in every exchange
                // (a shared-memory window of the log-pmf tables was tried and made the fill slower: the chain is bound by instruction issue
                //  at 16 warps/SM, not by the gathers — profiles/r1_noise_chain_phases.txt)
                const size_t smem = 0;
                const uint32_t bs = 256;
                int per_sm = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_noise_chain, bs, smem);
                const uint32_t max_blocks = (uint32_t)std::max(1, per_sm) * (uint32_t)ctx().sm_count;
                const uint32_t want_threads = std::max<uint32_t>({n_big * 32u, n_sel - n_big, (uint32_t)(tasks.size() / 3) * 32u});
                const uint32_t grid = std::max(1u, std::min((want_threads + bs - 1) / bs, max_blocks));
                uint32_t chain_id = chain + 1, n_sel_arg = n_sel, iters_arg = iters;
                float ps = cd->prior_shape, pc = cd->prior_scale;
                btg_gibbs_opts o = *opts;
                if (comm) { px.seq0 = comm->seq; comm->seq += iters; }
                uint32_t n_tasks = (uint32_t)(tasks.size() / 3);
                void *args[] = {&u->du, &T, &o, &d_sel, &n_sel_arg, &n_big, &chain_id, &iters_arg, &ns, &ps, &pc, &hist, &joint, &px, &d_tasks, &n_tasks, &gb};
                cudaError_t e = cudaLaunchCooperativeKernel((void *)k_noise_chain, dim3(grid), dim3(bs), args, smem, s);
                BTG_LAUNCHED();
                if (e != cudaSuccess) { set_error("cooperative launch failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; break; }
            } else {
                if (ns.trace) { k_noise_update<<<1, 256, 0, s>>>(ns, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 3, 0, chain + 1, 0, 1); BTG_LAUNCHED(); }
                for (uint32_t it = 1; it <= iters; it++) {
                    k_noise_update<<<1, 256, 0, s>>>(ns, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 1, opts->gibbs_burn_in < it, chain + 1, it, 1);
                    BTG_LAUNCHED();
                }
            }
            // resetNoiseRates at the end of the chain (InferenceEngine.cpp:253); not a trace row
            k_noise_update<<<1, 256, 0, s>>>(ns_quiet, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 0, 0, 0, 0, 1);
            BTG_LAUNCHED();
            if (cudaGetLastError() != cudaSuccess) rc = BTG_ECUDA;
        }
        if (rc == BTG_OK) {
            // mean of the post-burn-in rates -> setNoiseRates (InferenceEngine.cpp:259-264); final trace row "0 0"
            if (!joint) k_noise_update<<<1, 256, 0, s>>>(ns, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 2, 0, 0, 0, (double)opts->gibbs_samples * opts->n_chains);
            BTG_LAUNCHED();
            if (trace_out) cudaMemcpyAsync(trace_out, ns.trace, trace_rows * (2 + S) * 8, cudaMemcpyDeviceToHost, s);
            cudaError_t e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { set_error("noise estimation failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; }
            if (rc == BTG_OK && ns.phase_ns) {
                unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                cudaMemcpy(ph, ns.phase_ns, sizeof ph, cudaMemcpyDeviceToHost);
                if (BTG_NOISE_TIMING) {
                    unsigned long long sub[16];
                    cudaMemcpy(sub, ns.phase_ns + 8 + 3 * (size_t)iters, sizeof sub, cudaMemcpyDeviceToHost);
                    const char *nm[8] = {"bind", "fill rows", "rng load", "sample diplotypes", "sample frequencies", "rng save", "noise counts", "clear cache"};
                    if (sub[8]) {
                        fprintf(stderr, "[btgpu]   one-thread clusters, mean cycles per cluster-iteration:");
                        for (int k = 0; k < 8; k++) fprintf(stderr, " %s %.0f;", nm[k], (double)sub[k] / (double)sub[8]);
                        fprintf(stderr, "\n");
                    }
                }
                const char *what[3] = {"fill task", "large-cluster owner", "one-thread cluster"};
                if (BTG_NOISE_TIMING) {   // per-iteration maxima (accumulated over the chains by atomicMax): median over iterations and the cluster that is most often the slowest
                    std::vector<unsigned long long> it_max(3 * (size_t)iters);
                    cudaMemcpy(it_max.data(), ns.phase_ns + 8, it_max.size() * 8, cudaMemcpyDeviceToHost);
                    for (int k = 0; k < 3; k++) {
                        std::vector<double> us;
                        std::vector<uint32_t> who;
                        for (uint32_t i = 10; i < iters; i++) { us.push_back((it_max[3 * i + k] >> 32) / 1e3); who.push_back((uint32_t)(it_max[3 * i + k] & 0xFFFFFFFFu)); }
                        if (us.empty()) continue;
                        std::vector<double> sorted_us = us;
                        std::sort(sorted_us.begin(), sorted_us.end());
                        std::sort(who.begin(), who.end());
                        uint32_t best = who[0], best_n = 0, run = 0;
                        for (size_t i = 0; i < who.size(); i++) { run = (i && who[i] == who[i - 1]) ? run + 1 : 1; if (run > best_n) { best_n = run; best = who[i]; } }
                        fprintf(stderr, "[btgpu]   per-iteration slowest %s: median %.1f us, p90 %.1f us; most often cluster %u (%u of %zu iterations; H %u, variants %llu, fill cost %u)\n",
                                what[k], sorted_us[sorted_us.size() / 2], sorted_us[sorted_us.size() * 9 / 10], best, best_n, who.size(), best < u->du.C ? u->h_nhap[best] : 0,
                                best < u->du.C ? (unsigned long long)(u->h_cl_var_off[best + 1] - u->h_cl_var_off[best]) : 0ull, best < u->du.C ? u->h_fill_cost[best] : 0);
                    }
                }
                for (int k = 0; k < 3 && BTG_NOISE_TIMING; k++) {
                    const uint32_t c = (uint32_t)(ph[4 + k] & 0xFFFFFFFFu);
                    if (ph[4 + k] && c < u->du.C)
                        fprintf(stderr, "[btgpu]   slowest %s: %.1f us, cluster %u (H %u, variants %llu, fill cost %u)\n", what[k], (ph[4 + k] >> 32) / 1e3, c, u->h_nhap[c],
                                (unsigned long long)(u->h_cl_var_off[c + 1] - u->h_cl_var_off[c]), u->h_fill_cost[c]);
                }
                const double n_it = (double)opts->n_chains * iters;
                fprintf(stderr, "[btgpu] noise chain phases, us per iteration (block 0): fill %.1f  sample %.1f  exchange+update %.1f  release %.1f\n",
                        ph[0] / n_it / 1e3, ph[1] / n_it / 1e3, ph[2] / n_it / 1e3, ph[3] / n_it / 1e3);
            }
            if (rc == BTG_OK && comm && world > 1) {
                uint32_t flag = 0;
                cudaMemcpy(&flag, comm->error, 4, cudaMemcpyDeviceToHost);
                if (flag) { set_error("peer exchange timed out waiting for rank %u (a rank failed or left the lock-step)", flag - 1); cudaMemset(comm->error, 0, 4); rc = BTG_ECUDA; }
            }
        } else {
            set_error("noise estimation failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    cudaStreamSynchronize(s);
    for (void *p : tmp) cudaFree(p);
    if (d_tasks) cudaFree(d_tasks);
    return rc;
}

}  // extern "C"
