// btgpu.hpp — C++ host mirror of the reference's stage objects over the C ABI (include/btgpu.h).
//
// The reference is C++ whose seams are ordinary classes.  These header-only RAII adaptors keep
// the reference's class and method names and argument meaning, and turn the ABI's error codes into exceptions
// (the reference prints to cerr and calls exit(1), e.g. src/kmerBloom/KmerBloom.cpp:67-71):
//
//   btg::KmerBloom          ≡ KmerBloom<55>          include/kmerBloom/KmerBloom.hpp:48-77
//   btg::CountDistribution  ≡ CountDistribution      include/bayesTyper/CountDistribution.hpp:44-90
//   btg::InferenceUnit      ≡ the VariantClusterHaplotypes of an InferenceUnit (flat descriptors)
//   btg::InferenceEngine    ≡ InferenceEngine        include/bayesTyper/InferenceEngine.hpp:56-98
//   btg::VariantClusterGraphs ≡ the VariantClusterGraph set of an InferenceUnit + findVariantClusterPaths' best paths
//   btg::KmerCounter        ≡ KmerCounter + KmerCountsHash for one unit   include/bayesTyper/KmerCounter.hpp:53-67
//
// There is no CPU implementation behind any of it: every method is one or two libbtgpu calls.
#pragma once
#include <bitset>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "btgpu.h"

namespace btg {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc) {
    if (rc != BTG_OK) throw Error(btg_last_error());
}
template <class T> inline T *check_ptr(T *p) {
    if (!p) throw Error(btg_last_error());
    return p;
}

// one per process and GPU (src/bayesTyper/main.cpp:80 is where the reference would call it)
struct Library {
    explicit Library(int device = 0) { check(btg_init(device)); }
    ~Library() { btg_shutdown(); }
    Library(const Library &) = delete;
    Library &operator=(const Library &) = delete;
};

class KmerBloom {
    btg_bloom *b_;

  public:
    static constexpr int kmer_size = BTG_KMER_SIZE;
    using Kmer = std::bitset<2 * BTG_KMER_SIZE>;  // memory layout == the ABI's packed k-mer (two 64-bit words)
    static_assert(sizeof(Kmer) == 16, "std::bitset<110> must be two 64-bit words");

    KmerBloom(uint64_t num_kmers, float false_positive_rate) : b_(check_ptr(btg_bloom_create(num_kmers, false_positive_rate, kmer_size))) {}
    explicit KmerBloom(const std::string &prefix) : b_(check_ptr(btg_bloom_load(prefix.c_str(), kmer_size))) {}
    ~KmerBloom() { btg_bloom_free(b_); }
    KmerBloom(const KmerBloom &) = delete;
    KmerBloom &operator=(const KmerBloom &) = delete;

    void save(const std::string &prefix) const { check(btg_bloom_save(b_, prefix.c_str())); }
    // batched addKmer(bitset) / lookup(bitset) (KmerBloom.cpp:178-200)
    void addKmers(const std::vector<Kmer> &kmers) { check(btg_bloom_insert(b_, reinterpret_cast<const uint64_t *>(kmers.data()), kmers.size())); }
    std::vector<uint8_t> lookup(const std::vector<Kmer> &kmers) const {
        std::vector<uint8_t> hit(kmers.size());
        check(btg_bloom_lookup(b_, reinterpret_cast<const uint64_t *>(kmers.data()), kmers.size(), hit.data()));
        return hit;
    }
    const btg_bloom *handle() const { return b_; }
};

class CountDistribution {
    btg_count_dist *cd_;
    uint32_t n_samples_;

  public:
    // ctor + setGenomicCountDistributions (CountDistribution.cpp:51-141): per-sample NB (p, size) of one haploid copy
    CountDistribution(const std::vector<double> &nb_p, const std::vector<double> &nb_size, float prior_shape = 1.0f, float prior_scale = 0.01f)
        : cd_(nullptr), n_samples_((uint32_t)nb_p.size()) {
        if (nb_p.size() != nb_size.size()) throw Error("CountDistribution: p / size length mismatch");
        cd_ = check_ptr(btg_count_dist_create(n_samples_, nb_p.data(), nb_size.data(), prior_shape, prior_scale));
    }
    ~CountDistribution() { btg_count_dist_free(cd_); }
    CountDistribution(const CountDistribution &) = delete;
    CountDistribution &operator=(const CountDistribution &) = delete;

    void setNoiseRates(const std::vector<double> &rates) {  // CountDistribution.cpp:153-161
        if (rates.size() != n_samples_) throw Error("setNoiseRates: wrong number of samples");
        check(btg_count_dist_set_noise_rates(cd_, rates.data()));
    }
    std::vector<double> getNoiseRates() const {
        std::vector<double> r(n_samples_);
        check(btg_count_dist_get_noise_rates(cd_, r.data()));
        return r;
    }
    uint32_t numSamples() const { return n_samples_; }
    btg_count_dist *handle() const { return cd_; }
};

// result arrays of one unit: the fields of `Genotypes` (include/bayesTyper/Genotypes.hpp:46-99), flat
struct GenotypeArrays {
    std::vector<uint64_t> allele_off, geno_off, valt_off;
    std::vector<uint16_t> gt, saf, hc;
    std::vector<uint32_t> gq, an, ac;
    std::vector<float> gpp, app, nak, fak, mac, af, acp;
    std::vector<uint8_t> ploidy, anc;
    btg_genotype_result view;

    GenotypeArrays(uint32_t S, const uint16_t *var_nalleles, uint64_t n_variants) {
        allele_off.assign(1, 0); geno_off.assign(1, 0); valt_off.assign(1, 0);
        for (uint64_t v = 0; v < n_variants; v++) {
            const uint64_t nA = var_nalleles[v];
            allele_off.push_back(allele_off.back() + S * nA);
            geno_off.push_back(geno_off.back() + S * (nA * (nA + 1) / 2));
            valt_off.push_back(valt_off.back() + nA);
        }
        const uint64_t nall = allele_off.back(), ngen = geno_off.back(), nalt = valt_off.back();
        gt.resize(n_variants * S * 2); gq.resize(n_variants * S); gpp.resize(ngen); app.resize(nall); nak.resize(nall); fak.resize(nall); mac.resize(nall);
        saf.resize(nall); ploidy.resize(n_variants * S); an.resize(n_variants); ac.resize(nalt); af.resize(nalt); acp.resize(nalt); anc.resize(nalt);
        hc.resize(n_variants);
        bind();
    }
    // `view` points into this object's own vectors: a copy must not keep the source's pointers.  Moves keep the heap buffers, so the
    // view of a moved-to object is rebuilt from its own (now owning) vectors; copies are not offered.
    GenotypeArrays(const GenotypeArrays &) = delete;
    GenotypeArrays &operator=(const GenotypeArrays &) = delete;
    GenotypeArrays(GenotypeArrays &&o) noexcept
        : allele_off(std::move(o.allele_off)), geno_off(std::move(o.geno_off)), valt_off(std::move(o.valt_off)), gt(std::move(o.gt)), saf(std::move(o.saf)),
          hc(std::move(o.hc)), gq(std::move(o.gq)), an(std::move(o.an)), ac(std::move(o.ac)), gpp(std::move(o.gpp)), app(std::move(o.app)), nak(std::move(o.nak)),
          fak(std::move(o.fak)), mac(std::move(o.mac)), af(std::move(o.af)), acp(std::move(o.acp)), ploidy(std::move(o.ploidy)), anc(std::move(o.anc)) { bind(); }
    GenotypeArrays &operator=(GenotypeArrays &&) = delete;

private:
    void bind() {
        const uint64_t n_variants = hc.size();
        view = btg_genotype_result{n_variants, allele_off.data(), geno_off.data(), gt.data(), gq.data(), gpp.data(), app.data(), nak.data(), fak.data(),
                                   mac.data(), saf.data(), ploidy.data(), an.data(), valt_off.data(), ac.data(), af.data(), acp.data(), anc.data(), hc.data()};
    }
};

class InferenceUnit {
    btg_unit *u_;
    uint32_t n_samples_;
    std::vector<uint16_t> var_nalleles_;

  public:
    explicit InferenceUnit(const btg_unit_desc &d) : u_(check_ptr(btg_unit_upload(&d))), n_samples_(d.n_samples) {
        var_nalleles_.assign(d.var_nalleles, d.var_nalleles + d.cl_var_off[d.n_clusters]);
    }
    // adopts a unit that is already resident in HBM (KmerCounter::classifyPathKmers below)
    InferenceUnit(btg_unit *resident, uint32_t n_samples, const uint16_t *var_nalleles, uint64_t n_variants)
        : u_(check_ptr(resident)), n_samples_(n_samples), var_nalleles_(var_nalleles, var_nalleles + n_variants) {}
    ~InferenceUnit() { btg_unit_free(u_); }
    InferenceUnit(const InferenceUnit &) = delete;
    InferenceUnit &operator=(const InferenceUnit &) = delete;
    btg_unit *handle() const { return u_; }
    uint32_t numSamples() const { return n_samples_; }
    GenotypeArrays allocResult() const { return GenotypeArrays(n_samples_, var_nalleles_.data(), var_nalleles_.size()); }
};

// The VariantClusterGraph set of one inference unit (CSR-flattened, btg_graphs_desc) and the best paths that
// KmerCounter::findVariantClusterPaths leaves in it (VariantClusterGraph::best_paths_indices).
class VariantClusterGraphs {
    btg_graphs *g_;
    uint32_t n_clusters_;

  public:
    struct BestPaths {
        std::vector<uint32_t> n_paths;     // [C]
        std::vector<uint64_t> path_off;    // [C+1] byte offset of a cluster's rows in membership
        std::vector<uint8_t> membership;   // one byte per (path, vertex), path-major
    };
    VariantClusterGraphs(const btg_graphs_desc &d, uint32_t num_samples, uint32_t max_sample_haplotypes)
        : g_(check_ptr(btg_graphs_upload(&d, num_samples, max_sample_haplotypes))), n_clusters_(d.n_clusters) {}
    ~VariantClusterGraphs() { btg_graphs_free(g_); }
    VariantClusterGraphs(const VariantClusterGraphs &) = delete;
    VariantClusterGraphs &operator=(const VariantClusterGraphs &) = delete;
    btg_graphs *handle() const { return g_; }
    uint32_t numClusters() const { return n_clusters_; }
    void reset() { check(btg_graphs_reset(g_)); }
    BestPaths bestPaths() const {
        BestPaths b;
        b.n_paths.resize(n_clusters_);
        b.path_off.resize((size_t)n_clusters_ + 1);
        check(btg_get_best_paths(g_, b.n_paths.data(), b.path_off.data(), nullptr, 0));
        b.membership.resize(b.path_off[n_clusters_]);
        check(btg_get_best_paths(g_, b.n_paths.data(), b.path_off.data(), b.membership.data(), b.membership.size()));
        return b;
    }
};

// KmerCounter (include/bayesTyper/KmerCounter.hpp:53-67) with the KmerCountsHash it fills (KmerHash.hpp:73-87), for one inference unit.
// The reference's methods take the unit, the hash and Bloom filters as arguments and spawn threads; here the table lives in HBM
// behind the handle, the sample k-mers and the inter-cluster regions are device buffers (btg_device_alloc / btg_copy_to_device), and
// every method is one launch sequence on the library stream.  Stage order as main.cpp:594-617.
class KmerCounter {
    btg_counter *k_ = nullptr;
    uint32_t n_samples_ = 0;
    std::vector<uint16_t> var_nalleles_;
    uint32_t prng_seed_;

  public:
    explicit KmerCounter(uint32_t prng_seed) : prng_seed_(prng_seed) {}
    ~KmerCounter() { btg_counter_free(k_); }
    KmerCounter(const KmerCounter &) = delete;
    KmerCounter &operator=(const KmerCounter &) = delete;

    // KmerCounter::findVariantClusterPaths (KmerCounter.cpp:70-103) for ONE sample: the reference holds one sample's filter at a time
    // (main.cpp:219-247); samples must come in order
    void findVariantClusterPaths(VariantClusterGraphs *graphs, const KmerBloom &sample_bloom, uint32_t sample_idx, uint16_t max_sample_haplotypes) const {
        check(btg_find_sample_paths(graphs->handle(), sample_bloom.handle(), sample_idx, prng_seed_, max_sample_haplotypes));
    }
    // ... and for the samples first .. first + n - 1 in one launch, when their filters are resident together (same best paths)
    void findVariantClusterPaths(VariantClusterGraphs *graphs, const std::vector<const KmerBloom *> &sample_blooms, uint32_t first_sample_idx,
                                 uint16_t max_sample_haplotypes) const {
        std::vector<const btg_bloom *> h;
        for (const KmerBloom *b : sample_blooms) h.push_back(b->handle());
        check(btg_find_sample_paths_batch(graphs->handle(), h.data(), first_sample_idx, (uint32_t)h.size(), prng_seed_, max_sample_haplotypes));
    }
    // KmerCounter::countPathKmers (KmerCounter.cpp:252-289): `unit` describes the graphs, the best paths and the variant tables of the unit
    // (btg_counter_desc; host arrays, copied); returns the number of distinct path k-mers
    uint64_t countPathKmers(const btg_counter_desc &unit) {
        btg_counter_free(k_);
        k_ = check_ptr(btg_counter_create(&unit));
        n_samples_ = unit.n_samples;
        var_nalleles_.assign(unit.var_nalleles, unit.var_nalleles + unit.cl_var_off[unit.n_clusters]);
        uint64_t n = 0;
        check(btg_counter_count_path_kmers(k_, &n));
        return n;
    }
    // KmerCounter::countInterclusterKmers (KmerCounter.cpp:291-386) for the regions of one ploidy class ('N'-separated text in HBM)
    void countInterclusterKmers(const char *regions_dev, size_t len, bool is_decoy, uint32_t ploidy_female, uint32_t ploidy_male) {
        check(btg_counter_count_intercluster_kmers(need(), regions_dev, len, is_decoy ? 1 : 0, ploidy_female, ploidy_male));
    }
    // KmerCounter::parseSampleKmers (KmerCounter.cpp:431-524) for one sample's KMC records (listing order) in HBM
    void parseSampleKmers(uint32_t sample_idx, const uint64_t *kmers_dev, const uint8_t *counts_dev, size_t n) {
        check(btg_counter_parse_sample_kmers(need(), sample_idx, kmers_dev, counts_dev, n));
    }
    // KmerCounter::classifyPathKmers (KmerCounter.cpp:526-600) + VariantClusterGraph::getHaplotypeCandidates of every cluster:
    // the inference unit the genotypers are built from, resident in HBM (nothing crosses the boundary).
    // multigroup_kmers: <cluster_data>/multigroup_kmers filter of the cluster stage, or nullptr (exact)
    InferenceUnit classifyPathKmers(const KmerBloom *multigroup_kmers, const std::vector<uint8_t> &group_ploidy) {
        btg_unit *u = btg_counter_build_unit(need(), multigroup_kmers ? multigroup_kmers->handle() : nullptr, group_ploidy.data());
        return InferenceUnit(u, n_samples_, var_nalleles_.data(), var_nalleles_.size());
    }
    // countInterclusterParameterKmers' genotype-side half + calculateKmerStats + setGenomicCountDistributions (KmerCounter.cpp:171-250,
    // KmerHash.cpp:257-347, CountDistribution.cpp:66-141): negative-binomial (p, size) per sample
    struct GenomicParameters { std::vector<double> nb_p, nb_size; std::vector<uint32_t> modal_multiplicity; std::vector<uint64_t> n_modal_kmers; };
    GenomicParameters fitGenomicCountDistributions(const char *regions_dev, size_t len, uint32_t ploidy_female, uint32_t ploidy_male,
                                                   const std::vector<const uint64_t *> &sample_kmers_dev, const std::vector<const uint8_t *> &sample_counts_dev,
                                                   const std::vector<size_t> &sample_n, const uint64_t *parameter_kmers, size_t n_parameter_kmers,
                                                   uint64_t max_parameter_kmers) {
        GenomicParameters p;
        p.nb_p.resize(n_samples_); p.nb_size.resize(n_samples_); p.modal_multiplicity.resize(n_samples_); p.n_modal_kmers.resize(n_samples_);
        check(btg_counter_fit_nb(need(), regions_dev, len, ploidy_female, ploidy_male, sample_kmers_dev.data(), sample_counts_dev.data(), sample_n.data(),
                                 parameter_kmers, n_parameter_kmers, prng_seed_, max_parameter_kmers, p.nb_p.data(), p.nb_size.data(),
                                 p.modal_multiplicity.data(), p.n_modal_kmers.data()));
        return p;
    }
    btg_counter *handle() const { return k_; }

  private:
    btg_counter *need() const {
        if (!k_) throw Error("KmerCounter: countPathKmers has not been called");
        return k_;
    }
};

class InferenceEngine {
    btg_gibbs_opts opts_;

  public:
    explicit InferenceEngine(const btg_gibbs_opts &opts) : opts_(opts) {}
    // InferenceEngine::estimateNoise (InferenceEngine.cpp:135-276); returns the rows of <prefix>_noise_parameters.txt
    std::vector<double> estimateNoise(CountDistribution *cd, InferenceUnit *unit) const {
        const size_t rows = (size_t)opts_.n_chains * (opts_.gibbs_burn_in + opts_.gibbs_samples + 1) + 1;
        std::vector<double> trace(rows * (2 + unit->numSamples()));
        check(btg_estimate_noise(unit->handle(), cd->handle(), &opts_, trace.data()));
        return trace;
    }
    // InferenceEngine::estimateGenotypes (InferenceEngine.cpp:278-382)
    void estimateGenotypes(InferenceUnit *unit, const CountDistribution &cd, GenotypeArrays *out) const {
        check(btg_estimate_genotypes(unit->handle(), cd.handle(), &opts_, &out->view));
    }
    // InferenceEngine::estimateNoiseAndGenotypes (InferenceEngine.cpp:384-472, --noise-genotyping)
    std::vector<double> estimateNoiseAndGenotypes(InferenceUnit *unit, CountDistribution *cd, GenotypeArrays *out) const {
        const size_t rows = (size_t)opts_.n_chains * (opts_.gibbs_burn_in + opts_.gibbs_samples + 1);
        std::vector<double> trace(rows * (2 + unit->numSamples()));
        check(btg_estimate_noise_and_genotypes(unit->handle(), cd->handle(), &opts_, &out->view, trace.data()));
        return trace;
    }
};

}  // namespace btg
