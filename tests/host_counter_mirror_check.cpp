#include "btgpu.hpp"
// compile-and-link check of the KmerCounter seam mirror: every method instantiated, nothing run
int use(btg::KmerCounter &kc, btg::VariantClusterGraphs &g, const btg::KmerBloom &b, const btg_counter_desc &d) {
    kc.findVariantClusterPaths(&g, b, 0, 32);
    kc.findVariantClusterPaths(&g, std::vector<const btg::KmerBloom *>{&b}, 0, 32);
    auto bp = g.bestPaths();
    uint64_t n = kc.countPathKmers(d);
    kc.countInterclusterKmers(nullptr, 0, false, 2, 2);
    kc.parseSampleKmers(0, nullptr, nullptr, 0);
    btg::InferenceUnit u = kc.classifyPathKmers(nullptr, std::vector<uint8_t>{2});
    auto p = kc.fitGenomicCountDistributions(nullptr, 0, 2, 2, {nullptr}, {nullptr}, {0}, nullptr, 0, 1000000);
    return (int)(n + bp.n_paths.size() + p.nb_p.size() + u.numSamples());
}
int main(int argc, char **) { return argc > 100 ? 1 : 0; }
