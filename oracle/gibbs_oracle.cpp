// gibbs_oracle.cpp — "oracle-P": CPU restatement of the reference's per-cluster Gibbs
// sampler.  ONE restatement, TWO draw sources (bto_set_rng_mode):
//   mode 0 (Philox)  — the SAME counter-based streams as the CUDA kernels: the GPU must reproduce its tallies exactly;
//   mode 1 (mt19937) — std::mt19937 + the libstdc++ distributions, seeds, running-stream semantics and unordered-container
//                      walks of the reference: oracle-P must then reproduce the REFERENCE's
//                      diplotype tallies / noise trace exactly (tests/test_ref_parity_exact.py, against oracle-R dumps).
// The two modes share every line below except the bodies of struct Rng and the few `if (mt())` branches that
// express the reference's stream ownership (one engine running through all chains) against this library's
// per-chain / per-sample stream addressing.
//
// TEST INFRASTRUCTURE ONLY (see kmer_oracle.c).  It follows the reference function by
// function (citations are file:line under /root/reference) with std containers, and is
// deliberately written independently of bayestyper_b200/csrc/gibbs*.cu.
//
// Parity status: PINNED, exactly.
//   * vs the reference (oracle-R = the reference's own translation units, oracle/ref_build): in mt19937 mode this file
//     reproduces the reference's diplotype tallies, hence every genotype field, the allele statistics and the full
//     noise-rate trace EXACTLY — default mode, estimateNoise, --noise-genotyping with 2 and 30 samples, nested and deeply
//     nested groups (tests/test_ref_parity_exact.py, 12 cases against fixtures dumped by oracle-R).
//   * vs the CUDA path: in Philox mode the kernels must reproduce this file's tallies (posteriors to 1e-4; in practice
//     identical tallies): tests/test_gpu_gibbs.py, tests/test_gpu_gibbs_wide.py.
//   * the Philox mode is also compared statistically with reference runs (identical hard calls on confident sites, GPP
//     within Monte-Carlo error): tests/test_ref_parity.py — a second, independent check of the same restatement.
//
// Random-stream specification (shared with the kernels):
//   Philox4x32-10, key = (random_seed, uint32(group_index + 1)),
//   counter = (n_lo, n_hi, cluster_idx_in_group, kind | chain << 8); words are consumed in
//   order x0..x3 of successive counters n = 0,1,2...
//   kinds: 0 genotyper (k-mer subsample, diplotype draws), 1 sparsity estimator,
//          2 haplotype-frequency distribution, 3 branch shuffle, 4 noise rates (group 0, cluster 0)
//   u01  = ((hi32:lo32 >> 11) + 0.5) * 2^-53      (hi word drawn first)
//   uniform_int(n) = (u32 * n) >> 32
//   normal = sqrt(-2 ln u1) * cos(2 pi u2)          gamma(a>=1) = Marsaglia-Tsang
//   shuffle = Fisher-Yates from the back: for i = n-1..1: swap(a[i], a[uniform_int(i+1)])
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <random>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/btgpu.h"

namespace {

const uint16_t NONE = 0xFFFF;  // Utils::ushort_overflow

int g_rng_mode = 0;                          // 0 = Philox (the kernels' streams), 1 = mt19937 (the reference's)
const uint64_t *g_group_index = nullptr;     // optional: index of every group of the descriptor in the FULL unit (seeds)
inline bool mt() { return g_rng_mode == 1; }
inline uint64_t groupIndex(const btg_gibbs_opts *o, uint32_t g) { return g_group_index ? g_group_index[g] : o->group_index_base + (uint64_t)g * (o->group_index_stride ? o->group_index_stride : 1); }

// ------------------------------------------------------------------ Philox4x32-10
struct Philox {
    uint32_t key[2], ctr[4], buf[4];
    int pos;
    uint32_t t_draw;  // sampleDiplotypes calls since the stream was keyed
    void init(uint32_t seed, uint64_t group_index, uint32_t cluster_idx, uint32_t kind, uint32_t chain = 0) {
        key[0] = seed;
        key[1] = (uint32_t)(group_index + 1);
        ctr[0] = ctr[1] = 0;
        ctr[2] = cluster_idx;
        ctr[3] = kind | (chain << 8);
        pos = 4;
        t_draw = 0;
    }
    static void block(uint32_t *c, uint32_t k0, uint32_t k1) {
        uint32_t k[2] = {k0, k1};
        for (int r = 0; r < 10; r++) {
            round(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
    }
    // the diplotype draw of sample s in the current sampleDiplotypes call owns the block (t_draw, 1 + s): words x0:x1
    double u01_draw(uint32_t s) const {
        uint32_t c[4] = {t_draw, 1u + s, ctr[2], ctr[3]};
        block(c, key[0], key[1]);
        uint64_t x = ((uint64_t)c[0] << 32) | c[1];
        return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
    static void round(uint32_t *c, const uint32_t *k) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    uint32_t next() {
        if (pos == 4) {
            uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
            for (int r = 0; r < 10; r++) {
                round(c, k);
                k[0] += 0x9E3779B9u;
                k[1] += 0xBB67AE85u;
            }
            memcpy(buf, c, sizeof(buf));
            if (++ctr[0] == 0) ++ctr[1];
            pos = 0;
        }
        return buf[pos++];
    }
    double u01() {
        uint64_t hi = next(), lo = next();
        uint64_t x = (hi << 32) | lo;
        return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    }
    uint32_t uniform_int(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
    double normal() {
        double u1 = u01(), u2 = u01();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
    }
    double gamma(double a) {  // scale 1
        if (a < 1.0) {
            double g = gamma(a + 1.0);
            return g * std::pow(u01(), 1.0 / a);
        }
        const double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            double u = u01();
            double x2 = x * x;
            if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
            if (std::log(u) < 0.5 * x2 + d * (1.0 - v + std::log(v))) return d * v;
        }
    }
    template <class T> void shuffle(std::vector<T> &a) {
        for (size_t i = a.size(); i > 1; i--) std::swap(a[i - 1], a[uniform_int((uint32_t)i)]);
    }
};

// ------------------------------------------------------------------ one random stream, either draw source
// mt19937 seeds: genotyper / sparsity estimator / frequency distribution r + (g+1)(chain+1) + c
// (InferenceEngine.cpp:70, VariantClusterGroup.cpp:181, VariantClusterGenotyper.cpp:61,100,105 — three engines, one seed),
// branch shuffle r + (g+1)(chain+1) (InferenceEngine.cpp:71,295), CountDistribution and the engine's own stream r
// (CountDistribution.cpp:53, InferenceEngine.cpp:174).
struct Rng {
    Philox ph;
    std::mt19937 eng;
    std::gamma_distribution<double> gamma_dist;  // a member in the reference too (FrequencyDistribution.hpp, CountDistribution.hpp): its
                                                 // normal_distribution keeps a spare value between calls
    void init(uint32_t seed, uint64_t group_index, uint32_t cluster_idx, uint32_t kind, uint32_t chain = 0) {
        ph.init(seed, group_index, cluster_idx, kind, chain);
        if (mt()) {
            uint32_t sd = seed;
            if (kind <= 3) sd += (uint32_t)(group_index + 1) * (chain + 1);
            if (kind <= 2) sd += cluster_idx;
            eng = std::mt19937(sd);
            gamma_dist = std::gamma_distribution<double>();
        }
    }
    double canonical() { return std::generate_canonical<double, std::numeric_limits<double>::digits>(eng); }
    double u01() { return mt() ? canonical() : ph.u01(); }
    // LogDiscreteSampler::sample's uniform for sample s (DiscreteSampler.cpp:120-126): the reference takes it from the running engine
    double u01_draw(uint32_t s) { return mt() ? canonical() : ph.u01_draw(s); }
    void draws_done() { ph.t_draw++; }
    bool bernoulli(float rate) {  // VariantClusterHaplotypes.cpp:118,124
        if (mt()) { std::bernoulli_distribution d(rate); return d(eng); }
        return ph.u01() < (double)rate;
    }
    uint32_t uniform_int(uint32_t n) {  // FrequencyDistribution.cpp:253-254
        if (mt()) { std::uniform_int_distribution<> d(0, (int)n - 1); return (uint32_t)d(eng); }
        return ph.uniform_int(n);
    }
    double gamma(double shape, double scale = 1) {  // FrequencyDistribution.cpp:79-84,236-248, CountDistribution.cpp:202-213
        if (mt()) { gamma_dist.param(std::gamma_distribution<double>::param_type(shape, scale)); return gamma_dist(eng); }
        return ph.gamma(shape) * scale;
    }
    template <class T> void shuffle(std::vector<T> &a) {
        if (mt()) std::shuffle(a.begin(), a.end(), eng); else ph.shuffle(a);
    }
};

// boost::hash<pair<ushort, ushort>> as oracle-R's shim defines it (oracle/ref_build/shim/boost/functional/hash.hpp): fixes the
// walk order of diplotype_sampling_frequencies (VariantClusterGenotyper.hpp:112) in the mt19937 mode
struct PairHash {
    size_t operator()(const std::pair<uint16_t, uint16_t> &p) const {
        size_t seed = 0;
        seed ^= std::hash<uint16_t>()(p.first) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        seed ^= std::hash<uint16_t>()(p.second) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};
// walk of an unordered container: the reference's own order (same libstdc++, same history of insertions) in the mt19937 mode,
// ascending otherwise (the kernels walk ascending indices)
template <class Map> std::vector<typename Map::key_type> walkKeys(const Map &m) {
    std::vector<typename Map::key_type> v;
    for (auto &kv : m) v.push_back(kv.first);
    if (!mt()) std::sort(v.begin(), v.end());
    return v;
}
template <class Set> std::vector<typename Set::key_type> walk(const Set &s) {
    std::vector<typename Set::key_type> v(s.begin(), s.end());
    if (!mt()) std::sort(v.begin(), v.end());
    return v;
}

// ------------------------------------------------------------------ Utils.hpp:81-124
const double double_precision = std::numeric_limits<double>::epsilon();
const float float_precision = std::numeric_limits<float>::epsilon();
bool doubleCompare(double a, double b) { return (a == b) || (std::abs(a - b) < std::abs(std::min(a, b)) * double_precision * 100); }
bool floatCompare(float a, float b) { return (a == b) || (std::abs(a - b) < std::abs(std::min(a, b)) * float_precision * 100); }
bool floatLess(float a, float b) { return (a < b) && !floatCompare(a, b); }
double logAddition(double a, double b) { return a < b ? b + std::log1p(std::exp(a - b)) : a + std::log1p(std::exp(b - a)); }

// ------------------------------------------------------------------ KmerStats.cpp:34-124
struct KStats {
    uint32_t count = 0;
    double fraction = 0, mean = 0, M2 = 0;
    void reset() { count = 0; fraction = mean = M2 = 0; }
    void add(double v, bool valid) {
        if (!valid) return;
        count++;
        fraction += (static_cast<double>(!doubleCompare(v, 0)) - fraction) / count;
        double delta = v - mean;
        mean += delta / count;
        M2 += delta * (v - mean);
    }
    std::pair<double, bool> getFraction() const { return count == 0 ? std::make_pair(-1.0, false) : std::make_pair(fraction, true); }
    std::pair<double, bool> getMean() const { return count == 0 ? std::make_pair(-1.0, false) : std::make_pair(mean, true); }
};
struct AlleleKStats {
    std::vector<KStats> count_stats, fraction_stats, mean_stats;
    explicit AlleleKStats(int n = 0) : count_stats(n), fraction_stats(n), mean_stats(n) {}
    void addKmerStats(const KStats &k, uint16_t allele) {  // KmerStats.cpp:115-122
        count_stats[allele].add(k.count, true);
        auto f = k.getFraction();
        fraction_stats[allele].add(f.first, f.second);
        auto m = k.getMean();
        mean_stats[allele].add(m.first, m.second);
    }
};

// ------------------------------------------------------------------ CountDistribution
struct CountTables {
    uint32_t S = 0;
    std::vector<double> p, size, noise_rates;
    std::vector<double> genomic;  // [S][256][256]
    std::vector<double> noise;    // [S][256]
    double prior_shape = 1, prior_scale = 0.01;

    // NegativeBinomialDistribution::logPmf (NegativeBinomialDistribution.cpp:121-147)
    double nbLogPmf(uint32_t s, uint32_t obs, uint32_t scale) const {
        double coef = std::lgamma(obs + size[s] * scale) - std::lgamma(size[s] * scale) - std::lgamma(obs + 1);
        return coef + std::log(p[s]) * size[s] * scale + std::log(1 - p[s]) * obs;
    }
    // CountDistribution::genomicCountLogPmf (CountDistribution.cpp:267-312)
    double genomicCountLogPmf(uint32_t s, uint32_t m, uint32_t c) const {
        if (m == 0) return c == 0 ? 0 : -std::numeric_limits<double>::infinity();
        double v = nbLogPmf(s, c, m);
        if (c == 255) {
            uint32_t limit = c;
            double prev = 0;
            do {
                limit++;
                prev = v;
                v = logAddition(v, nbLogPmf(s, limit, m));
                if (v > 0) { v = 0; break; }
            } while (!doubleCompare(prev, v));
        }
        return v;
    }
    // CountDistribution::poissonLogProb / noiseCountLogPmf (CountDistribution.cpp:314-352)
    static double poissonLogProb(uint32_t value, double rate) { return value * std::log(rate) - rate - std::lgamma(value + 1); }
    double noiseCountLogPmf(uint32_t s, uint32_t c) const {
        double v = poissonLogProb(c, noise_rates[s]);
        if (c == 255) {
            uint32_t limit = c;
            double prev = 0;
            do {
                limit++;
                prev = v;
                v = logAddition(v, poissonLogProb(limit, noise_rates[s]));
                if (v > 0) { v = 0; break; }
            } while (!doubleCompare(prev, v));
        }
        return v;
    }
    void updateGenomic() {
        genomic.assign((size_t)S * 65536, 0);
        for (uint32_t s = 0; s < S; s++)
            for (uint32_t m = 0; m < 256; m++)
                for (uint32_t c = 0; c < 256; c++) genomic[((size_t)s * 256 + m) * 256 + c] = genomicCountLogPmf(s, m, c);
    }
    void updateNoise() {
        noise.assign((size_t)S * 256, 0);
        for (uint32_t s = 0; s < S; s++)
            for (uint32_t c = 0; c < 256; c++) noise[(size_t)s * 256 + c] = noiseCountLogPmf(s, c);
    }
    // CountDistribution::calcCountLogProb (CountDistribution.cpp:255-265)
    double logProb(uint32_t s, uint8_t m, uint8_t c) const { return m == 0 ? noise[(size_t)s * 256 + c] : genomic[((size_t)s * 256 + m) * 256 + c]; }
};

typedef std::pair<uint16_t, uint16_t> Dipl;

// ------------------------------------------------------------------ per-cluster genotyper
struct Genotyper {
    const btg_unit_desc *d;
    const btg_gibbs_opts *o;
    uint32_t c;  // global cluster index
    uint32_t S, H, K, nvar;
    uint64_t row0, var0;
    const uint8_t *M;
    Rng prng, prng_freq;
    std::vector<uint32_t> uniq, multi, uniq_sub, multi_sub;
    // SparseFrequencyDistribution / FrequencyDistribution state
    bool sparse = false;
    double sparsity = 0;
    std::vector<uint32_t> obs;
    std::vector<double> freq;
    std::vector<uint8_t> nz;
    std::unordered_set<uint32_t> plus, zero;  // plus_count_indices, zero_count_indices (FrequencyDistribution.hpp)
    uint32_t num_hap_count = 0, num_missing_count = 0;
    std::map<std::pair<uint32_t, uint32_t>, std::vector<double> > simplex_cache;
    // VariantClusterGenotyper state
    std::vector<std::map<Dipl, double> > unique_cache, multi_cache;
    bool use_multi = false;
    std::vector<uint8_t> sample_multi;  // sample_multicluster_kmer_multiplicities [multi_sub][S]
    uint8_t *shared = nullptr;          // KmerCounts::multiplicities of the group-shared k-mer records [id][S]
    uint64_t hap0 = 0;                  // index of this cluster's first haplotype in hap_nested_off
    std::vector<Dipl> dipl;
    std::unordered_map<Dipl, std::vector<uint32_t>, PairHash> tally;  // diplotype_sampling_frequencies
    struct StatsCache { bool update = true; std::vector<KStats> h1, h2; };
    std::vector<StatsCache> stats_cache;
    std::vector<std::vector<AlleleKStats> > allele_stats;  // [var][sample]

    uint8_t m(uint32_t k, uint16_t h) const { return M[(size_t)k * H + h]; }
    uint8_t count(uint32_t k, uint32_t s) const { return d->k_has_counts[row0 + k] ? d->k_counts[(row0 + k) * S + s] : 0; }
    uint8_t ic(uint32_t k, uint32_t s) const { return d->k_has_counts[row0 + k] ? d->k_ic[(row0 + k) * 2 + d->sample_gender[s]] : 0; }
    uint16_t nalleles(uint32_t v) const { return d->var_nalleles[var0 + v]; }
    bool isMissing(uint32_t v, uint16_t a) const { return d->var_dep[var0 + v] && a == nalleles(v) - 1; }  // VariantInfo.hpp:82-94
    uint16_t hapAllele(uint16_t h, uint32_t v) const { return d->hap_alleles[d->cl_hapvar_off[c] + (size_t)h * nvar + v]; }

    // VariantClusterHaplotypes::getDiplotypeKmerMultiplicity / getUniqueKmerMultiplicity (VariantClusterHaplotypes.cpp:45-76)
    uint8_t diplMult(uint32_t k, const Dipl &dp) const {
        uint8_t r = 0;
        if (dp.first != NONE) r += m(k, dp.first);
        if (dp.second != NONE) r += m(k, dp.second);
        return r;
    }
    uint8_t uniqueMult(uint32_t k, const Dipl &dp, uint32_t s) const { return (uint8_t)(diplMult(k, dp) + ic(k, s)); }
    // KmerCounts::getSampleMultiplicity of a multicluster k-mer (KmerCounts.cpp:205-224)
    uint8_t &sharedMult(uint32_t k, uint32_t s) const { return shared[(size_t)d->k_shared[row0 + k] * S + s]; }
    // VariantClusterHaplotypes::getMulticlusterKmerMultiplicity (VariantClusterHaplotypes.cpp:76-93)
    uint8_t multiMult(uint32_t k, const Dipl &dp, const Dipl &prev, uint32_t s) const {
        if (count(k, s) == 0) return (uint8_t)(diplMult(k, dp) + ic(k, s));
        return (uint8_t)(sharedMult(k, s) - diplMult(k, prev) + diplMult(k, dp) + ic(k, s));
    }
    // VariantClusterHaplotypes::getPreviousMulticlusterKmerMultiplicity (VariantClusterHaplotypes.cpp:95-108)
    uint8_t prevMultiMult(uint32_t sub, const Dipl &dp, const Dipl &prev, uint32_t s) const {
        const uint32_t k = multi_sub[sub];
        return (uint8_t)(sample_multi[(size_t)sub * S + s] - diplMult(k, prev) + diplMult(k, dp) + ic(k, s));
    }

    // VariantClusterGenotyper ctor (VariantClusterGenotyper.cpp:59-106)
    void init(const btg_unit_desc *desc, const btg_gibbs_opts *opts, uint32_t cluster, uint64_t group_index, uint32_t chain,
              uint8_t *shared_mult = nullptr, uint64_t first_haplotype = 0) {
        d = desc; o = opts; c = cluster;
        shared = shared_mult;
        hap0 = first_haplotype;
        S = d->n_samples;
        H = d->cl_nhap[c];
        row0 = d->cl_kmer_off[c];
        K = (uint32_t)(d->cl_kmer_off[c + 1] - row0);
        var0 = d->cl_var_off[c];
        nvar = (uint32_t)(d->cl_var_off[c + 1] - var0);
        M = d->mult + d->cl_mult_off[c];
        prng.init(o->random_seed, group_index, d->cluster_idx[c], 0, chain);
        prng_freq.init(o->random_seed, group_index, d->cluster_idx[c], 2, chain);
        uniq.assign(d->uniq_idx + d->cl_uniq_off[c], d->uniq_idx + d->cl_uniq_off[c + 1]);
        multi.assign(d->multi_idx + d->cl_multi_off[c], d->multi_idx + d->cl_multi_off[c + 1]);
        unique_cache.assign(S, std::map<Dipl, double>());
        multi_cache.assign(S, std::map<Dipl, double>());
        assert(multi.empty() || shared);
        dipl.assign(S, Dipl(NONE, NONE));
        stats_cache.assign(S, StatsCache());
        for (auto &sc : stats_cache) { sc.h1.assign(nvar, KStats()); sc.h2.assign(nvar, KStats()); }
        allele_stats.clear();
        for (uint32_t v = 0; v < nvar; v++) allele_stats.emplace_back(S, AlleleKStats(nalleles(v)));
        // SparsityEstimator::estimateMinimumColumnCover (SparsityEstimator.cpp:41-90), own stream (kind 1)
        Rng sp;
        sp.init(o->random_seed, group_index, d->cluster_idx[c], 1, chain);
        std::vector<uint8_t> uncovered(K);
        uint32_t n_unc = 0;
        for (uint32_t k = 0; k < K; k++) { uncovered[k] = d->k_has_counts[row0 + k]; n_unc += uncovered[k]; }
        std::vector<uint32_t> cover;
        while (n_unc > 0) {
            std::vector<uint32_t> col(H, 0);
            for (uint32_t k = 0; k < K; k++)
                if (uncovered[k]) for (uint32_t h = 0; h < H; h++) col[h] += m(k, h);
            uint32_t mx = *std::max_element(col.begin(), col.end());
            assert(mx > 0);
            std::vector<uint32_t> ties;
            for (uint32_t h = 0; h < H; h++) if (col[h] == mx) ties.push_back(h);
            // DiscreteSampler with unit weights: cum = 1..n, u*n, upper_bound (DiscreteSampler.cpp:61-87)
            double x = sp.u01() * (double)ties.size();
            uint32_t idx = 0;
            if (ties.size() > 1) { while (idx < ties.size() && !(x < (double)(idx + 1))) idx++; assert(idx < ties.size()); }
            uint32_t h = ties[idx];
            cover.push_back(h);
            for (uint32_t k = 0; k < K; k++)
                if (uncovered[k] && m(k, h)) { uncovered[k] = 0; n_unc--; }
        }
        // SparseHaplotypeFrequencyDistribution ctor (HaplotypeFrequencyDistribution.cpp:79-90)
        sparse = !cover.empty();
        if (sparse) sparsity = std::min(cover.size() / static_cast<double>(H), 1 - double_precision * 100);  // FrequencyDistribution.cpp:97-103
        obs.assign(H, 0);
        resetFrequencies();
    }
    // (Sparse)FrequencyDistribution::reset (FrequencyDistribution.cpp:46-51,104-115)
    void resetFrequencies() {
        assert(num_hap_count == 0 && num_missing_count == 0);
        obs.assign(H, 0);
        freq.assign(H, 1 / static_cast<double>(H));
        nz.assign(H, 1);
        plus.clear();
        zero.clear();
        for (uint32_t i = 0; i < H; i++) zero.insert(i);
    }
    // VariantClusterHaplotypes::isMaxHaplotypeVariantKmer (VariantClusterHaplotypes.cpp:159-178)
    bool isMaxHapVarKmer(std::vector<uint32_t> &cnt, uint32_t k) {
        bool is_max = true;
        for (uint64_t e = d->kmer_vh_off[row0 + k]; e < d->kmer_vh_off[row0 + k + 1]; e++) {
            const uint16_t v = d->vh_var[e];
            const uint8_t *bits = d->vh_bits + d->vh_bits_off[e];
            for (uint32_t h = 0; h < H; h++)
                if (bits[h] && cnt[(size_t)h * nvar + v] < o->max_haplotype_variant_kmers) { cnt[(size_t)h * nvar + v]++; is_max = false; }
        }
        return is_max;
    }
    // VariantClusterGenotyper::reset + VariantClusterHaplotypes::sampleKmerSubset (…Genotyper.cpp:113-129, …Haplotypes.cpp:110-157)
    void reset() {
        uniq_sub.clear();
        multi_sub.clear();
        const float rate = o->kmer_subsampling_rate;
        std::vector<uint32_t> cnt((size_t)H * nvar, 0);
        prng.shuffle(uniq);
        for (auto k : uniq) if (prng.bernoulli(rate)) if (!isMaxHapVarKmer(cnt, k)) uniq_sub.push_back(k);
        prng.shuffle(multi);
        for (auto k : multi) if (prng.bernoulli(rate)) if (!isMaxHapVarKmer(cnt, k)) multi_sub.push_back(k);
        sample_multi.assign(multi_sub.size() * (size_t)S, 0);
        for (auto &sc : stats_cache) sc.update = true;
        use_multi = false;
        clearCache();
        resetFrequencies();
    }
    void clearCache() {  // VariantClusterGenotyper::clearCache (…Genotyper.cpp:131-138)
        for (auto &mcache : unique_cache) mcache.clear();
        for (auto &mcache : multi_cache) mcache.clear();
    }
    // VariantClusterGenotyper::updateMulticlusterDiplotypeLogProb (…Genotyper.cpp:569-595): the cached terms of k-mers
    // whose shared multiplicity was changed by another cluster of the group are replaced in place
    void updateMulticlusterDiplotypeLogProb(const CountTables &T, uint32_t s) {
        if (multi_cache[s].empty()) return;
        for (uint32_t sub = 0; sub < multi_sub.size(); sub++) {
            const uint32_t k = multi_sub[sub];
            if (!(count(k, s) > 0 && sharedMult(k, s) != sample_multi[(size_t)sub * S + s])) continue;  // isMulticlusterKmerUpdated
            for (auto &e : multi_cache[s]) {
                e.second -= T.logProb(s, prevMultiMult(sub, e.first, dipl[s], s), count(k, s));
                e.second += T.logProb(s, multiMult(k, e.first, dipl[s], s), count(k, s));
            }
        }
    }
    // VariantClusterHaplotypes::updateMulticlusterKmerMultiplicities (VariantClusterHaplotypes.cpp:197-233)
    void updateMulticlusterKmerMultiplicities(const Dipl &dp, const Dipl &prev, uint32_t s) {
        if (dp != prev) {
            stats_cache[s].update = true;
            for (auto k : multi) {
                const uint8_t cur = diplMult(k, dp), old = diplMult(k, prev);
                if (cur != old) { assert(old <= sharedMult(k, s)); sharedMult(k, s) -= old; sharedMult(k, s) += cur; }
            }
        }
        for (uint32_t sub = 0; sub < multi_sub.size(); sub++) {
            const uint32_t k = multi_sub[sub];
            if (diplMult(k, dp) > 0 && count(k, s) > 0 && sharedMult(k, s) != sample_multi[(size_t)sub * S + s]) stats_cache[s].update = true;
            sample_multi[(size_t)sub * S + s] = sharedMult(k, s);
        }
    }
    // VariantClusterGenotyper::calcDiplotypeLogProb (VariantClusterGenotyper.cpp:597-666)
    double calcDiplotypeLogProb(const CountTables &T, uint32_t s, const Dipl &dp) {
        double lp = 0;
        if (dp.second == NONE) lp += std::log(freq[dp.first]);
        else if (dp.first == dp.second) lp += 2 * std::log(freq[dp.first]);
        else lp += std::log(2) + std::log(freq[dp.first]) + std::log(freq[dp.second]);
        auto ins = unique_cache[s].emplace(dp, 0.0);
        if (ins.second) {
            double acc = 0;
            for (auto k : uniq_sub) acc += T.logProb(s, uniqueMult(k, dp, s), count(k, s));
            ins.first->second = acc;
        }
        lp += ins.first->second;
        if (use_multi) {
            auto mins = multi_cache[s].emplace(dp, 0.0);
            if (mins.second) {
                double acc = 0;
                for (auto k : multi_sub) acc += T.logProb(s, multiMult(k, dp, dipl[s], s), count(k, s));
                mins.first->second = acc;
            }
            lp += mins.first->second;
        }
        assert(std::isfinite(lp));
        return lp;
    }
    // HaplotypeFrequencyDistribution::incrementCount (HaplotypeFrequencyDistribution.cpp:114-126)
    void incrementCount(uint16_t h) {
        if (h == NONE) { num_missing_count++; return; }
        num_hap_count++;
        if (sparse && obs[h] == 0) { plus.insert(h); zero.erase(h); }  // FrequencyDistribution.cpp:198-207
        obs[h]++;
    }
    // VariantClusterGenotyper::sampleDiplotype (VariantClusterGenotyper.cpp:707-755)
    void sampleDiplotype(const std::vector<uint16_t> &nzh, const CountTables &T, uint32_t s, uint8_t ploidy) {
        std::vector<double> cum;
        std::vector<Dipl> outcomes;
        auto add = [&](double lp) { cum.push_back(cum.empty() ? lp : logAddition(lp, cum.back())); };  // LogDiscreteSampler::addOutcome
        if (ploidy == 2) {
            for (size_t i = 0; i < nzh.size(); i++)
                for (size_t j = i; j < nzh.size(); j++) { add(calcDiplotypeLogProb(T, s, Dipl(nzh[i], nzh[j]))); outcomes.emplace_back(nzh[i], nzh[j]); }
        } else if (ploidy == 1) {
            for (auto h : nzh) { add(calcDiplotypeLogProb(T, s, Dipl(h, NONE))); outcomes.emplace_back(h, NONE); }
        } else {
            add(0);
            outcomes.emplace_back(NONE, NONE);
        }
        // LogDiscreteSampler::sample + DiscreteSampler::search (DiscreteSampler.cpp:120-126,68-87)
        double x = std::log(prng.u01_draw(s)) + cum.back();
        uint32_t idx = 0;
        if (cum.size() > 1) { idx = (uint32_t)(std::upper_bound(cum.begin(), cum.end(), x) - cum.begin()); assert(idx < cum.size()); }
        dipl[s] = outcomes[idx];
        incrementCount(dipl[s].first);
        incrementCount(dipl[s].second);
    }
    // VariantClusterHaplotypes::updateKmerStatsCache (VariantClusterHaplotypes.cpp:302-333)
    void updateKmerStatsCache(uint32_t k, const Dipl &dp, uint32_t s, uint8_t mult) {
        double kc = d->k_has_counts[row0 + k] ? count(k, s) / static_cast<double>(mult) : 0;
        for (uint64_t e = d->kmer_vh_off[row0 + k]; e < d->kmer_vh_off[row0 + k + 1]; e++) {
            const uint16_t v = d->vh_var[e];
            const uint8_t *bits = d->vh_bits + d->vh_bits_off[e];
            if (bits[dp.first]) stats_cache[s].h1[v].add(kc, true);
            if (dp.second != NONE && bits[dp.second]) stats_cache[s].h2[v].add(kc, true);
        }
    }
    // VariantClusterHaplotypes::addHaplotypeKmerStats (VariantClusterHaplotypes.cpp:335-361)
    void addHaplotypeKmerStats(const std::vector<KStats> &cache, uint32_t s, uint16_t h) {
        uint32_t last = NONE;
        for (uint32_t v = 0; v < nvar; v++) {
            uint16_t a = hapAllele(h, v);
            if (isMissing(v, a)) { assert(last != NONE); allele_stats[v][s].addKmerStats(cache[last], a); }
            else { allele_stats[v][s].addKmerStats(cache[v], a); last = v; }
        }
    }
    // VariantClusterHaplotypes::NestedVariantClusterInfo (VariantClusterHaplotypes.hpp:103-109)
    struct NestedInfo { uint8_t ploidy; std::vector<KStats> stats; };
    // VariantClusterHaplotypes::updateAlleleKmerStats (VariantClusterHaplotypes.cpp:235-300)
    void updateAlleleKmerStats(const std::vector<NestedInfo> &nested) {
        for (uint32_t s = 0; s < S; s++) {
            const Dipl dp = dipl[s];
            auto &sc = stats_cache[s];
            if (sc.update) {
                sc.update = false;
                for (uint32_t v = 0; v < nvar; v++) { sc.h1[v].reset(); sc.h2[v].reset(); }
                if (dp.first != NONE)
                    for (auto k : uniq_sub) if (diplMult(k, dp) > 0) updateKmerStatsCache(k, dp, s, uniqueMult(k, dp, s));
                if (dp.first != NONE)
                    for (auto k : multi_sub) if (diplMult(k, dp) > 0) updateKmerStatsCache(k, dp, s, multiMult(k, dp, dp, s));
            }
            if (dp.first != NONE) addHaplotypeKmerStats(sc.h1, s, dp.first);
            if (dp.second != NONE) addHaplotypeKmerStats(sc.h2, s, dp.second);
            for (auto &ks : nested[s].stats)  // addNestedHaplotypeKmerStats (VariantClusterHaplotypes.cpp:363-372)
                for (uint32_t v = 0; v < nvar; v++) allele_stats[v][s].addKmerStats(ks, (uint16_t)(nalleles(v) - 1));
        }
    }
    // VariantClusterGenotyper::updateNestedVariantClusterInfo / updateNestedPloidy / addNestedKmerStats (…Genotyper.cpp:140-206)
    void updateNestedVariantClusterInfo(std::vector<NestedInfo> &nested, uint32_t child_cluster_idx) const {
        uint64_t dep = d->cl_dep_off[c];
        while (dep < d->cl_dep_off[c + 1] && d->dep_cluster[dep] != child_cluster_idx) dep++;
        for (uint32_t s = 0; s < S; s++) {
            for (int which = 0; which < 2; which++) {
                const uint16_t h = which == 0 ? dipl[s].first : dipl[s].second;
                if (h == NONE) continue;
                const uint32_t *nb = d->hap_nested + d->hap_nested_off[hap0 + h], *ne = d->hap_nested + d->hap_nested_off[hap0 + h + 1];
                if (std::binary_search(nb, ne, child_cluster_idx)) continue;
                assert(nested[s].ploidy != 0);
                nested[s].ploidy = nested[s].ploidy == 2 ? 1 : 0;
                assert(dep < d->cl_dep_off[c + 1] && nested[s].stats.size() < 2);
                uint32_t v = NONE;
                for (uint64_t e = d->dep_var_off[dep]; e < d->dep_var_off[dep + 1]; e++) {
                    const uint16_t nv = d->dep_var[e];
                    if (!isMissing(nv, hapAllele(h, nv))) { v = nv; break; }
                }
                assert(v != NONE);
                nested[s].stats.push_back(which == 0 ? stats_cache[s].h1[v] : stats_cache[s].h2[v]);
            }
        }
    }
    // VariantClusterGenotyper::sampleDiplotypes (VariantClusterGenotyper.cpp:668-705)
    void sampleDiplotypes(const CountTables &T, const std::vector<NestedInfo> &nested, bool collect) {
        std::vector<uint16_t> nzh;
        for (uint16_t h = 0; h < H; h++) if (nz[h]) nzh.push_back(h);
        for (uint32_t s = 0; s < S; s++) {
            const Dipl prev = dipl[s];
            updateMulticlusterDiplotypeLogProb(T, s);
            sampleDiplotype(nzh, T, s, nested[s].ploidy);
            updateMulticlusterKmerMultiplicities(dipl[s], prev, s);
            if (collect) {
                auto it = tally.emplace(dipl[s], std::vector<uint32_t>(S, 0)).first;
                it->second[s]++;
            }
        }
        prng.draws_done();
        if (collect) updateAlleleKmerStats(nested);
        use_multi = !multi_sub.empty();
    }
    void sampleDiplotypes(const CountTables &T, const uint8_t *ploidy, bool collect) {  // a group of one cluster
        std::vector<NestedInfo> top(S);
        for (uint32_t s = 0; s < S; s++) top[s].ploidy = ploidy[s];
        sampleDiplotypes(T, top, collect);
    }
    // SparseFrequencyDistribution::updateCachedSimplexProbVector (FrequencyDistribution.cpp:143-196)
    void simplexProbVector(std::vector<double> &vec, uint32_t n_obs, uint32_t plus_size) const {
        const double dirichlet_parameter = 1;
        double cardinal = 0;
        double prob_z = plus_size * std::log(sparsity) + (H - plus_size) * std::log(1 - sparsity);
        double prob_t = std::lgamma(plus_size * dirichlet_parameter) - std::lgamma(n_obs + plus_size * dirichlet_parameter);
        double row_sum = cardinal + prob_z + prob_t;
        vec.push_back(row_sum);
        for (uint32_t j = plus_size + 1; j < H + 1; j++) {
            cardinal = std::lgamma(H - plus_size + 1) - (std::lgamma(j - plus_size + 1) + std::lgamma(H - j + 1));
            prob_z = j * std::log(sparsity) + (H - j) * std::log(1 - sparsity);
            prob_t = std::lgamma(j * dirichlet_parameter) - std::lgamma(n_obs + j * dirichlet_parameter);
            double prob_eq = cardinal + prob_z + prob_t;
            row_sum += std::log(1 + std::exp(prob_eq - row_sum));
            vec.push_back(row_sum);
            if (doubleCompare(vec.back(), *(vec.rbegin() + 1))) break;
        }
        for (auto &p : vec) p = std::exp(p - row_sum);
    }
    // VariantClusterGenotyper::sampleHaplotypeFrequencies -> SparseHaplotypeFrequencyDistribution::sampleFrequencies
    // (…Genotyper.cpp:781-785, HaplotypeFrequencyDistribution.cpp:128-137, FrequencyDistribution.cpp:75-94,209-304)
    void sampleHaplotypeFrequencies() {
        if (num_hap_count > 0) {
            if (!sparse) {
                double norm = 0;
                for (uint32_t i = 0; i < H; i++) { freq[i] = prng_freq.gamma(obs[i] + 1.0); norm += freq[i]; obs[i] = 0; }
                for (auto &f : freq) f /= norm;
            } else {
                const uint32_t n_obs = num_hap_count;
                auto key = std::make_pair(n_obs, (uint32_t)plus.size() - 1);
                auto it = simplex_cache.find(key);
                if (it == simplex_cache.end()) {
                    it = simplex_cache.emplace(key, std::vector<double>()).first;
                    simplexProbVector(it->second, n_obs, (uint32_t)plus.size());
                }
                const std::vector<double> &pv = it->second;
                uint32_t simplex_size = (uint32_t)(std::upper_bound(pv.begin(), pv.end(), prng_freq.u01()) - pv.begin()) + (uint32_t)plus.size();
                double norm = 0;
                for (auto h : walk(plus)) { freq[h] = prng_freq.gamma(obs[h] + 1.0); norm += freq[h]; nz[h] = 1; }
                while (plus.size() < simplex_size) {
                    uint32_t posn = prng_freq.uniform_int((uint32_t)zero.size());
                    uint32_t h = walk(zero)[posn];  // posn-th zero-count haplotype of the walk
                    freq[h] = prng_freq.gamma(1.0);
                    norm += freq[h];
                    nz[h] = 1;
                    plus.insert(h);
                    zero.erase(h);
                }
                for (auto h : zero) { freq[h] = 0; nz[h] = 0; obs[h] = 0; }
                for (auto h : walk(plus)) { freq[h] /= norm; zero.insert(h); obs[h] = 0; }
                plus.clear();
            }
        }
        num_hap_count = 0;
        num_missing_count = 0;
    }
    // VariantClusterGenotyper::getNoiseCounts (VariantClusterGenotyper.cpp:757-779)
    void getNoiseCounts(std::vector<uint64_t> &hist) const {
        for (uint32_t s = 0; s < S; s++)
            for (auto k : uniq_sub)
                if (uniqueMult(k, dipl[s], s) == 0) hist[(size_t)s * 256 + count(k, s)]++;
    }
};

// VariantClusterGroup (VariantClusterGroup.cpp:47-260): the clusters of one group, the nested-cluster forest over them
// and the shared multiplicity records of their multicluster k-mers
struct Group {
    const btg_unit_desc *d;
    uint32_t g, S;
    uint64_t c0;
    std::vector<Genotyper> gts;
    std::vector<uint32_t> src;
    std::vector<std::vector<uint32_t> > out_edges;
    void init(const btg_unit_desc *desc, const btg_gibbs_opts *o, uint32_t group, uint32_t chain, uint8_t *shared, const std::vector<uint64_t> &hap_start) {
        d = desc; g = group; S = d->n_samples;
        c0 = d->group_cluster_off[g];
        const uint32_t n = (uint32_t)(d->group_cluster_off[g + 1] - c0);
        gts.resize(n);
        for (uint32_t i = 0; i < n; i++) gts[i].init(d, o, (uint32_t)(c0 + i), groupIndex(o, g), chain, shared, hap_start[c0 + i]);
        src.assign(d->group_src + d->group_src_off[g], d->group_src + d->group_src_off[g + 1]);
        out_edges.assign(n, std::vector<uint32_t>());
        for (uint64_t e = d->group_edge_off[g]; e < d->group_edge_off[g + 1]; e++) out_edges[d->group_edge_src[e]].push_back(d->group_edge_dst[e]);
    }
    void reset() { for (auto &gt : gts) gt.reset(); }  // VariantClusterGroup::initGenotyper (…Group.cpp:175-186)
    // VariantClusterGroup::shuffleBranchOrdering (…Group.cpp:208-218): cumulative, own stream (kind 3)
    void shuffleBranchOrdering(const btg_gibbs_opts *o, uint32_t chain) {
        Rng br;
        br.init(o->random_seed, groupIndex(o, g), 0, 3, chain);
        br.shuffle(src);
        for (auto &e : out_edges) br.shuffle(e);
    }
    // VariantClusterGroup::estimateGenotypes / runGibbsSample (…Group.cpp:220-250)
    void run(uint32_t v, const CountTables &T, const std::vector<Genotyper::NestedInfo> &nested, bool collect) {
        gts[v].sampleDiplotypes(T, nested, collect);
        gts[v].sampleHaplotypeFrequencies();
        for (auto t : out_edges[v]) {
            auto child = nested;
            gts[v].updateNestedVariantClusterInfo(child, d->cluster_idx[c0 + t]);
            run(t, T, child, collect);
        }
    }
    void estimateGenotypes(const CountTables &T, bool collect) {
        std::vector<Genotyper::NestedInfo> top(S);
        for (uint32_t s = 0; s < S; s++) top[s].ploidy = d->group_ploidy[(size_t)g * S + s];
        for (auto v : src) run(v, T, top, collect);
    }
};

// index of each cluster's first haplotype in hap_nested_off, and the number of shared multiplicity records
std::vector<uint64_t> haplotypeStarts(const btg_unit_desc *d) {
    std::vector<uint64_t> h(d->n_clusters + 1, 0);
    for (uint32_t c = 0; c < d->n_clusters; c++) h[c + 1] = h[c] + d->cl_nhap[c];
    return h;
}
size_t sharedRecords(const btg_unit_desc *d) {
    size_t n = 0;
    const uint64_t rows = d->cl_kmer_off[d->n_clusters];
    for (uint64_t r = 0; r < rows; r++) if (d->k_shared[r] != 0xFFFFFFFFu) n = std::max(n, (size_t)d->k_shared[r] + 1);
    return n;
}

uint16_t hapToAllele(const Genotyper &g, uint16_t h, uint32_t v) { return h != NONE ? g.hapAllele(h, v) : g.nalleles(v) - 1; }  // …Genotyper.cpp:208-219

// VariantClusterGenotyper::getGenotypes / getGenotypeSampleStats / getGenotypeVariantStats (…Genotyper.cpp:249-567)
void summarise(const Genotyper &g, const uint8_t *ploidy, btg_genotype_result *out) {
    const uint32_t S = g.S;
    for (uint32_t v = 0; v < g.nvar; v++) {
        const uint64_t gv = g.var0 + v;
        const uint16_t nA = g.nalleles(v);
        const uint32_t nG = nA * (nA + 1) / 2;
        out->hc[gv] = (uint16_t)g.H;
        // getNonCoveredAlleles (…Genotyper.cpp:221-247)
        std::vector<uint8_t> covered(nA, 0);
        for (uint16_t h = 0; h < g.H; h++) covered[g.hapAllele(h, v)] = 1;
        if (g.d->var_dep[gv]) covered[nA - 1] = 1;
        for (uint16_t a = 0; a < nA; a++) out->anc[out->valt_off[gv] + a] = !covered[a];
        uint32_t total_count = 0;
        std::vector<uint32_t> alt_counts(nA, 0);
        std::vector<float> acp(nA, 0);
        for (uint32_t s = 0; s < S; s++) {
            float *gpp = out->gpp + out->geno_off[gv] + (size_t)s * nG;
            float *app = out->app + out->allele_off[gv] + (size_t)s * nA;
            float *nak = out->nak + out->allele_off[gv] + (size_t)s * nA, *fak = out->fak + out->allele_off[gv] + (size_t)s * nA,
                  *mac = out->mac + out->allele_off[gv] + (size_t)s * nA;
            uint16_t *saf = out->saf + out->allele_off[gv] + (size_t)s * nA;
            const uint8_t pl = ploidy[s];
            out->ploidy[gv * S + s] = pl;
            const uint32_t n_geno = pl == 2 ? nG : (pl == 1 ? nA : 0), n_all = pl == 0 ? 0 : nA;
            for (uint32_t i = 0; i < nG; i++) gpp[i] = 0;
            for (uint32_t i = 0; i < nA; i++) { app[i] = 0; saf[i] = 0; }
            uint32_t n_it = 0;
            std::vector<Dipl> best;
            float best_p = 0;
            for (auto &key : walkKeys(g.tally)) {  // mt19937 mode: the reference's walk of diplotype_sampling_frequencies (…Genotyper.cpp:277)
                const auto &kv = *g.tally.find(key);
                const uint32_t cnt = kv.second[s];
                if (cnt == 0) continue;
                Dipl ge(NONE, NONE);
                uint32_t gi = NONE;
                if (pl == 2) {
                    ge.first = hapToAllele(g, kv.first.first, v);
                    ge.second = hapToAllele(g, kv.first.second, v);
                    if (ge.first > ge.second) std::swap(ge.first, ge.second);
                    gi = (ge.second * (ge.second + 1)) / 2 + ge.first;
                    gpp[gi] += cnt;
                    app[ge.first] += cnt;
                    if (ge.first != ge.second) app[ge.second] += cnt;
                } else if (pl == 1) {
                    ge.first = hapToAllele(g, kv.first.first, v);
                    gi = ge.first;
                    gpp[gi] += cnt;
                    app[gi] += cnt;
                }
                n_it += cnt;
                if (pl != 0) {
                    if (floatCompare(best_p, gpp[gi])) best.push_back(ge);
                    else if (best_p < gpp[gi]) { best.clear(); best.push_back(ge); best_p = gpp[gi]; }
                }
            }
            best_p /= n_it;
            for (uint32_t i = 0; i < n_geno; i++) gpp[i] /= n_it;
            for (uint32_t i = 0; i < n_all; i++) app[i] /= n_it;
            const AlleleKStats &ak = g.allele_stats[v][s];
            for (uint16_t a = 0; a < nA; a++) {
                nak[a] = (float)ak.count_stats[a].getMean().first;
                fak[a] = (float)ak.fraction_stats[a].getMean().first;
                mac[a] = (float)ak.mean_stats[a].getMean().first;
            }
            for (uint16_t a = 0; a < n_all; a++) {
                if (!floatCompare(app[a], 0)) {
                    auto cs = ak.count_stats[a].getMean();
                    if (floatLess((float)cs.first, g.o->min_number_of_kmers)) saf[a] += 1;
                    auto fs = ak.fraction_stats[a].getMean();
                    if (!floatCompare((float)cs.first, 0))
                        if (floatLess((float)fs.first, g.o->min_fraction_observed_kmers[s])) saf[a] += 2;
                }
            }
            uint32_t gq;
            if (floatCompare(best_p, 1)) gq = 99;
            else if (floatCompare(best_p, 0)) gq = 0;
            else gq = (uint32_t)(-10 * std::log10(1 - best_p));  // float arithmetic, as in the reference (float operands)
            out->gq[gv * S + s] = gq;
            uint16_t *gt = out->gt + (gv * S + s) * 2;
            gt[0] = NONE;
            gt[1] = pl == 2 ? NONE : 0xFFFE;
            if (pl == 2) {
                if (best.size() == 1 && !floatLess(best_p, g.o->min_genotype_posterior))
                    if (saf[best[0].first] == 0 && saf[best[0].second] == 0) { gt[0] = best[0].first; gt[1] = best[0].second; }
            } else if (pl == 1) {
                if (best.size() == 1 && !floatLess(best_p, g.o->min_genotype_posterior))
                    if (saf[best[0].first] == 0) gt[0] = best[0].first;
            }
            // getGenotypeVariantStats (…Genotyper.cpp:470-526)
            for (int i = 0; i < 2; i++)
                if (gt[i] < 0xFFFE) { total_count++; if (gt[i] > 0) alt_counts[gt[i]]++; }
            for (uint16_t a = 0; a < n_all; a++)
                if (saf[a] == 0) acp[a] = std::max(acp[a], app[a]);
        }
        out->an[gv] = total_count;
        for (uint16_t a = 0; a < nA; a++) {
            out->ac[out->valt_off[gv] + a] = alt_counts[a];
            out->af[out->valt_off[gv] + a] = total_count > 0 ? alt_counts[a] / static_cast<float>(total_count) : 0;
            out->acp[out->valt_off[gv] + a] = acp[a];
        }
    }
}

CountTables *asTables(void *p) { return static_cast<CountTables *>(p); }

}  // namespace

extern "C" {

void *bto_count_dist_create(uint32_t S, const double *p, const double *size, float prior_shape, float prior_scale) {
    auto *t = new CountTables();
    t->S = S;
    t->p.assign(p, p + S);
    t->size.assign(size, size + S);
    t->noise_rates.assign(S, 1.0);
    t->prior_shape = prior_shape;
    t->prior_scale = prior_scale;
    t->updateGenomic();
    t->updateNoise();
    return t;
}
void bto_count_dist_set_noise_rates(void *cd, const double *rates) {
    auto *t = asTables(cd);
    t->noise_rates.assign(rates, rates + t->S);
    t->updateNoise();
}
void bto_count_dist_tables(void *cd, double *genomic, double *noise) {
    auto *t = asTables(cd);
    if (genomic) memcpy(genomic, t->genomic.data(), t->genomic.size() * 8);
    if (noise) memcpy(noise, t->noise.data(), t->noise.size() * 8);
}
void bto_count_dist_free(void *cd) { delete asTables(cd); }

// NegativeBinomialDistribution::momentsToParameters (NegativeBinomialDistribution.cpp:68-79) + CountDistribution.cpp:115-116
void bto_nb_moments_to_parameters(double mean, double var, uint32_t multiplicity, double *p_out, double *size_out) {
    const double max_p = 0.99;
    if (max_p < (mean / var)) var = mean / max_p;
    *p_out = mean / var;
    *size_out = std::pow(mean, 2) / (var - mean) / multiplicity;
}

void bto_set_rng_mode(int mode) { g_rng_mode = mode; }
// index of every group of the descriptors passed from now on in the full unit (NULL: group_index_base + position)
void bto_set_group_indices(const uint64_t *idx) { g_group_index = idx; }

// InferenceEngine::estimateGenotypesCallback (InferenceEngine.cpp:278-333) for every group of the unit.
// tally_out (optional): per cluster dense [(H+1)(H+2)/2][S] tallies concatenated (tally_off[c]).
int bto_estimate_genotypes(const btg_unit_desc *d, void *cd, const btg_gibbs_opts *o, btg_genotype_result *out,
                           uint32_t *tally_out, const uint64_t *tally_off) {
    const CountTables &T = *asTables(cd);
    const std::vector<uint64_t> hap_start = haplotypeStarts(d);
    std::vector<uint8_t> shared(sharedRecords(d) * (size_t)d->n_samples, 0);
    for (uint32_t g = 0; g < d->n_groups; g++) {
        const uint8_t *ploidy = d->group_ploidy + (size_t)g * d->n_samples;
        Group grp;
        grp.init(d, o, g, 0, shared.data(), hap_start);  // genotypers are constructed once and persist across chains
        for (uint32_t chain = 0; chain < o->n_chains; chain++) {
            if (!mt() && grp.gts.size() == 1) {
                // stream contract of the default mode for single-cluster groups (csrc/gibbs.cu, k_estimate_genotypes): chains are
                // independent — each chain re-keys the genotyper's two streams with its index and shuffles the original k-mer
                // order.  (The reference keeps one mt19937 running through all chains, InferenceEngine.cpp:292-306.)
                Genotyper &gt = grp.gts[0];
                gt.prng.init(o->random_seed, groupIndex(o, g), d->cluster_idx[gt.c], 0, chain);
                gt.prng_freq.init(o->random_seed, groupIndex(o, g), d->cluster_idx[gt.c], 2, chain);
                gt.uniq.assign(d->uniq_idx + d->cl_uniq_off[gt.c], d->uniq_idx + d->cl_uniq_off[gt.c + 1]);
            }
            grp.reset();
            grp.shuffleBranchOrdering(o, chain);
            for (uint32_t i = 0; i < o->gibbs_burn_in; i++) grp.estimateGenotypes(T, false);
            for (uint32_t i = 0; i < o->gibbs_samples; i++) grp.estimateGenotypes(T, true);
        }
        for (auto &gt : grp.gts) {  // collectGenotypes (…Group.cpp:262-278): every variant is summarised with the CHROMOSOME ploidy
            summarise(gt, ploidy, out);
            if (tally_out) {
                const uint32_t H = gt.H;
                uint32_t *t = tally_out + tally_off[gt.c];
                for (auto &kv : gt.tally) {
                    uint32_t a = kv.first.first == NONE ? H : kv.first.first, b = kv.first.second == NONE ? H : kv.first.second;
                    for (uint32_t s = 0; s < d->n_samples; s++) t[((size_t)b * (b + 1) / 2 + a) * d->n_samples + s] = kv.second[s];
                }
            }
        }
    }
    return 0;
}

// CountDistribution::sampleNoiseParameters (CountDistribution.cpp:173-200): the reference evaluates both parameters in FLOAT
// (pair<float,float> prior combined with ulong)
static void sampleNoiseParameters(CountTables &T, Rng &noise_prng, const std::vector<uint64_t> &hist) {
    for (uint32_t s = 0; s < T.S; s++) {
        uint64_t n_obs = 0, sum = 0;
        for (uint32_t i = 0; i < 256; i++) { n_obs += hist[(size_t)s * 256 + i]; sum += i * hist[(size_t)s * 256 + i]; }
        const float shape_f = (float)T.prior_shape + (float)sum;
        const float scale_f = (float)T.prior_scale / ((float)n_obs * (float)T.prior_scale + 1);
        T.noise_rates[s] = noise_prng.gamma(shape_f, scale_f);
    }
    T.updateNoise();
}
// CountDistribution::resetNoiseRates (CountDistribution.cpp:163-171)
static void resetNoiseRates(CountTables &T, Rng &noise_prng) {
    for (uint32_t s = 0; s < T.S; s++) T.noise_rates[s] = noise_prng.gamma((float)T.prior_shape, (float)T.prior_scale);
    T.updateNoise();
}

// InferenceEngine::estimateNoise (InferenceEngine.cpp:135-276)
int bto_estimate_noise(const btg_unit_desc *d, void *cd, const btg_gibbs_opts *o, double *trace_out) {
    CountTables &T = *asTables(cd);
    const uint32_t S = d->n_samples;
    const uint32_t noise_variants_batch_size = 100000;
    std::vector<uint32_t> noise_groups;
    for (uint32_t g = 0; g < d->n_groups; g++)
        if (d->group_cluster_off[g + 1] - d->group_cluster_off[g] == 1) noise_groups.push_back(g);
    auto groupVariants = [&](uint32_t g) {
        uint32_t n = 0;
        for (uint64_t c = d->group_cluster_off[g]; c < d->group_cluster_off[g + 1]; c++) n += (uint32_t)(d->cl_var_off[c + 1] - d->cl_var_off[c]);
        return n;
    };
    Rng engine, noise_prng;
    engine.init(o->random_seed, (uint64_t)-1, 0, 5);  // key1 = 0: InferenceEngine's own stream (InferenceEngine.cpp:174)
    noise_prng.init(o->random_seed, (uint64_t)-1, 0, 4);  // CountDistribution::prng (CountDistribution.cpp:53)
    // Stream contract (csrc/gibbs.cu, estimate_noise_concurrent): the chains of estimateNoise are independent — chain c draws its noise
    // rates from its own stream (kind 4, chain c + 1), starting with the prior draw.  The reference takes that draw from the running
    // CountDistribution stream: in the ctor for the first chain, at the end of the previous chain afterwards (CountDistribution.cpp:62,
    // InferenceEngine.cpp:253).
    if (mt()) resetNoiseRates(T, noise_prng);
    std::vector<double> mean_rates(S, 0);
    size_t row = 0;
    auto trace = [&](double chain, double it) {
        if (!trace_out) return;
        double *r = trace_out + row * (2 + S);
        r[0] = chain; r[1] = it;
        for (uint32_t s = 0; s < S; s++) r[2 + s] = T.noise_rates[s];
        row++;
    };
    for (uint32_t chain = 0; chain < o->n_chains; chain++) {
        engine.shuffle(noise_groups);
        uint32_t end = 0, nv = 0;
        while (nv < noise_variants_batch_size && end < noise_groups.size()) { nv += groupVariants(noise_groups[end]); end++; }
        std::sort(noise_groups.begin(), noise_groups.begin() + end);
        if (!mt()) {
            noise_prng.init(o->random_seed, (uint64_t)-1, 0, 4, chain + 1);
            resetNoiseRates(T, noise_prng);
        }
        std::vector<Genotyper> gts(end);
        for (uint32_t i = 0; i < end; i++) {  // initGenotypersCallback (InferenceEngine.cpp:60-75): fresh genotypers per chain
            gts[i].init(d, o, (uint32_t)d->group_cluster_off[noise_groups[i]], groupIndex(o, noise_groups[i]), mt() ? chain : chain + 1);
            gts[i].reset();
        }
        trace(chain + 1, 0);
        for (uint32_t it = 1; it <= (uint32_t)o->gibbs_burn_in + o->gibbs_samples; it++) {
            std::vector<uint64_t> hist((size_t)S * 256, 0);
            for (uint32_t i = 0; i < end; i++) {  // sampleGenotypesCallback (InferenceEngine.cpp:77-98)
                const uint8_t *ploidy = d->group_ploidy + (size_t)noise_groups[i] * S;
                gts[i].sampleDiplotypes(T, ploidy, false);
                gts[i].sampleHaplotypeFrequencies();
                gts[i].getNoiseCounts(hist);
                gts[i].clearCache();  // clearGenotyperCache
            }
            sampleNoiseParameters(T, noise_prng, hist);
            trace(chain + 1, it);
            if (o->gibbs_burn_in < it) for (uint32_t s = 0; s < S; s++) mean_rates[s] += T.noise_rates[s];
        }
        if (mt()) resetNoiseRates(T, noise_prng);  // InferenceEngine.cpp:253
    }
    for (uint32_t s = 0; s < S; s++) mean_rates[s] /= (double)o->gibbs_samples * o->n_chains;
    T.noise_rates = mean_rates;
    T.updateNoise();
    trace(0, 0);
    return 0;
}


// InferenceEngine::estimateNoiseAndGenotypes (InferenceEngine.cpp:384-472): every group advances in lock-step, the noise
// rates are redrawn after EVERY iteration (burn-in and sampling alike) and the genotypers persist across chains (only
// chain 0's seed is ever used: VariantClusterGroup::initGenotyper constructs once, InferenceEngine.cpp:70).
int bto_estimate_noise_and_genotypes(const btg_unit_desc *d, void *cd, const btg_gibbs_opts *o, btg_genotype_result *out, double *trace_out) {
    CountTables &T = *asTables(cd);
    const uint32_t S = d->n_samples;
    const std::vector<uint64_t> hap_start = haplotypeStarts(d);
    std::vector<uint8_t> shared(sharedRecords(d) * (size_t)S, 0);
    Rng noise_prng;
    noise_prng.init(o->random_seed, (uint64_t)-1, 0, 4);
    resetNoiseRates(T, noise_prng);  // CountDistribution ctor
    size_t row = 0;
    auto trace = [&](double chain, double it) {
        if (!trace_out) return;
        double *r = trace_out + row * (2 + S);
        r[0] = chain; r[1] = it;
        for (uint32_t s = 0; s < S; s++) r[2 + s] = T.noise_rates[s];
        row++;
    };
    std::vector<Group> grps(d->n_groups);
    for (uint32_t chain = 0; chain < o->n_chains; chain++) {
        for (uint32_t g = 0; g < d->n_groups; g++) {  // initGenotypersCallback
            if (chain == 0) grps[g].init(d, o, g, 0, shared.data(), hap_start);
            grps[g].reset();
            grps[g].shuffleBranchOrdering(o, chain);
        }
        trace(chain + 1, 0);
        for (uint32_t it = 1; it <= (uint32_t)o->gibbs_burn_in + o->gibbs_samples; it++) {
            std::vector<uint64_t> hist((size_t)S * 256, 0);
            for (uint32_t g = 0; g < d->n_groups; g++) {  // sampleGenotypesCallback
                grps[g].estimateGenotypes(T, it > o->gibbs_burn_in);
                for (auto &gt : grps[g].gts) gt.getNoiseCounts(hist);
                for (auto &gt : grps[g].gts) gt.clearCache();
            }
            sampleNoiseParameters(T, noise_prng, hist);
            trace(chain + 1, it);
        }
        resetNoiseRates(T, noise_prng);
    }
    for (uint32_t g = 0; g < d->n_groups; g++)
        for (auto &gt : grps[g].gts) summarise(gt, d->group_ploidy + (size_t)g * S, out);
    return 0;
}

void bto_count_dist_get_noise_rates(void *cd, double *out) {
    auto *t = asTables(cd);
    memcpy(out, t->noise_rates.data(), t->S * 8);
}

}  // extern "C"
