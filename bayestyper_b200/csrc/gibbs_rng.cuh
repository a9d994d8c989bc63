// gibbs_rng.cuh — counter-based random streams of the Gibbs path (device side).
//
// The reference draws from std::mt19937 + libstdc++ distributions (one engine per
// VariantClusterGenotyper / FrequencyDistribution / SparsityEstimator / CountDistribution,
// SURVEY.md appendix C).  A 2.5 KB Mersenne state per cluster does not belong in registers;
// each of those engines becomes an independent Philox4x32-10 stream addressed by
//   key     = (random_seed, uint32(group_index + 1))
//   counter = (n_lo, n_hi, cluster_idx_in_group, kind | chain << 8)
// which keeps the reference's determinism contract (results independent of how groups are
// scheduled or sharded).  n counts the blocks a stream consumes in sequence (n_hi = 0 in practice).  The one draw
// that the model allows to be taken side by side — the diplotype of sample s in the t-th sampleDiplotypes call of
// a genotyper since its stream was keyed (VariantClusterGenotyper.cpp:668-705: the loop over samples only adds up
// haplotype counts) — owns its own block instead: counter = (t, 1 + s, cluster_idx, kind | chain << 8), so it
// does not depend on the order in which the samples are visited (one thread walking them, or one lane each).  The exact draw recipes are part of the parity contract with
// oracle/gibbs_oracle.cpp and are documented in DESIGN.md §RNG.
#pragma once
#include <cstdint>

namespace btg {

// The sampler kernels would inline every call, including libm's f64 log/exp/log1p/cos/pow (60-300 instructions each) and
// the ten Philox rounds at every call site: k_estimate_genotypes was 15.1k SASS instructions with a hot loop of ~63 KB, and
// at full occupancy (16 warps per SM, each at a different point of the loop) it stalled on instruction fetch
// (no_instruction = 16 of 23 stall cycles per issue, profiles/r1_gibbs_full_occupancy_ncu_full.txt).  Keeping these leaf
// functions out of line (scalar arguments in registers, no state spilled) runs the same arithmetic 1.4x faster there
// (161k clusters: 730 ms -> 529 ms); with few resident warps it is ~5 % slower.  BTG_OUTLINE=0 restores full inlining.
#ifndef BTG_OUTLINE
#define BTG_OUTLINE 1
#endif
#if BTG_OUTLINE
#define BTG_LEAF static __device__ __noinline__
#else
#define BTG_LEAF static __device__ __forceinline__
#endif
BTG_LEAF double m_log(double x) { return log(x); }
BTG_LEAF double m_exp(double x) { return exp(x); }
BTG_LEAF double m_log1p(double x) { return log1p(x); }
BTG_LEAF double m_cos(double x) { return cos(x); }
BTG_LEAF double m_pow(double x, double y) { return pow(x, y); }
BTG_LEAF double m_sqrt(double x) { return sqrt(x); }
BTG_LEAF double m_div(double x, double y) { return x / y; }
BTG_LEAF uint4 philox4x32_10(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const uint32_t n0 = hi1 ^ x1 ^ k0, n2 = hi0 ^ x3 ^ k1;
        x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
}

enum RngKind : uint32_t { kRngGenotyper = 0, kRngSparsity = 1, kRngFrequency = 2, kRngBranch = 3, kRngNoise = 4, kRngEngine = 5 };

struct Philox {
    uint32_t key0, key1;
    uint32_t c0, c1, c2, c3;  // counter
    uint32_t b0, b1, b2, b3;  // current block
    uint32_t pos;             // next word of the block (4 = exhausted)
    uint32_t t_draw;          // sampleDiplotypes calls since the stream was keyed (addresses the per-sample draw blocks)

    __device__ __forceinline__ void init(uint32_t seed, uint64_t group_index, uint32_t cluster_idx, uint32_t kind, uint32_t chain = 0) {
        key0 = seed;
        key1 = (uint32_t)(group_index + 1);
        c0 = c1 = 0;
        c2 = cluster_idx;
        c3 = kind | (chain << 8);
        pos = 4;
        b0 = b1 = b2 = b3 = 0;
        t_draw = 0;
    }
    // the diplotype draw of sample s in the current sampleDiplotypes call: own counter block (t_draw, 1 + s), words x0:x1
    __device__ __forceinline__ double u01_draw(uint32_t s) const {
        const uint4 b = philox4x32_10(t_draw, 1u + s, c2, c3, key0, key1);
        const uint64_t x = ((uint64_t)b.x << 32) | b.y;
        return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ void refill() {
        const uint4 b = philox4x32_10(c0, c1, c2, c3, key0, key1);
        b0 = b.x; b1 = b.y; b2 = b.z; b3 = b.w;
        if (++c0 == 0) ++c1;
        pos = 0;
    }
    __device__ __forceinline__ uint32_t next() {
        if (pos == 4) refill();
        const uint32_t r = pos == 0 ? b0 : (pos == 1 ? b1 : (pos == 2 ? b2 : b3));
        pos++;
        return r;
    }
    // ((hi:lo >> 11) + 0.5) * 2^-53, in (0,1)
    __device__ __forceinline__ double u01() {
        const uint64_t hi = next(), lo = next();
        const uint64_t x = (hi << 32) | lo;
        return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ uint32_t uniform_int(uint32_t n) { return __umulhi(next(), n); }
    __device__ __forceinline__ double normal() {
        const double u1 = u01(), u2 = u01();
        return m_sqrt(-2.0 * m_log(u1)) * m_cos(6.283185307179586476925286766559 * u2);
    }
    // Marsaglia-Tsang, scale 1
    __device__ double gamma(double a) {
        double boost = 1.0;
        if (a < 1.0) {
            // gamma(a) = gamma(a+1) * U^(1/a); the gamma(a+1) draw comes first
            const double g = gamma_ge1(a + 1.0);
            return g * m_pow(u01(), m_div(1.0, a));
        }
        return boost * gamma_ge1(a);
    }
    __device__ double gamma_ge1(double a) {
        const double d = a - 1.0 / 3.0, c = m_div(1.0, m_sqrt(9.0 * d));
        for (;;) {
            const double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = u01();
            const double x2 = x * x;
            if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
            if (m_log(u) < 0.5 * x2 + d * (1.0 - v + m_log(v))) return d * v;
        }
    }
    // persist / restore (noise modes run one iteration per launch)
    // A: anything indexable (plain pointer or a lane-interleaved accessor)
    template <class A> __device__ __forceinline__ void save(A p, uint32_t o) const {
        p[o + 0] = c0; p[o + 1] = c1; p[o + 2] = b0; p[o + 3] = b1; p[o + 4] = b2; p[o + 5] = b3; p[o + 6] = pos; p[o + 7] = c3; p[o + 8] = t_draw;
    }
    template <class A> __device__ __forceinline__ void load(A p, uint32_t o, uint32_t seed, uint64_t group_index, uint32_t cluster_idx) {
        key0 = seed; key1 = (uint32_t)(group_index + 1); c2 = cluster_idx;
        c0 = p[o + 0]; c1 = p[o + 1]; b0 = p[o + 2]; b1 = p[o + 3]; b2 = p[o + 4]; b3 = p[o + 5]; pos = p[o + 6]; c3 = p[o + 7]; t_draw = p[o + 8];
    }
};

}  // namespace btg
