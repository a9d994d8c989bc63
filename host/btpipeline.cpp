// btpipeline — C++ host for BOTH hot paths over the C ABI, with no torch / Python in the process: the stage order of `bayesTyper cluster`'s
// path search and of `bayesTyper genotype` (src/bayesTyper/main.cpp:233-252,594-643) for one inference unit.
//
//   btpipeline <bundle.btd> <out.btd> [--device D] [--random-seed R] [--gibbs-burn-in B] [--gibbs-samples N] [--number-of-gibbs-chains C]
//              [--max-number-of-sample-haplotypes H] [--noise-genotyping]
//
// <bundle.btd> (bayestyper_b200/btd.py; written by tools/make_pipeline_bundle.py) holds what the two commands read for the unit:
//   g.*            the graphs of the unit (VariantClusterGraph ctor output, the arrays of btg_graphs_desc / btg_counter_desc), incl. var_nalleles, var_dep
//   regions        the inter-cluster regions as one 'N'-separated nucleotide buffer
//   s<i>.kmers / s<i>.counts   every sample's KMC records (packed k-mers, counts), meta.genders, meta.ploidy = (female, male) on this contig
//   parameter_kmers (optional) <out>_cluster_data/parameter_kmers.fa.gz as packed k-mers
// Stages: KmerBloom per sample (makeBloom) -> findVariantClusterPaths -> btg_counter {countPathKmers, countInterclusterKmers, parseSampleKmers,
// classifyPathKmers + getHaplotypeCandidates, NB fit} -> CountDistribution -> estimateNoise -> estimateGenotypes (or estimateNoiseAndGenotypes).
// <out.btd>: the fields of btg_genotype_result, the NB parameters and the final noise rates.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "btgpu.hpp"
#include "btd.hpp"

namespace {
using namespace btd;

const Array &get(const std::map<std::string, Array> &m, const std::string &name) {
    auto it = m.find(name);
    if (it == m.end()) throw btg::Error("bundle lacks '" + name + "'");
    return it->second;
}
void check(int rc) { if (rc != BTG_OK) throw btg::Error(btg_last_error()); }
template <class T> T *nonnull(T *p) { if (!p) throw btg::Error(btg_last_error()); return p; }

struct DevBuf {
    void *p = nullptr;
    DevBuf(const void *host, size_t bytes) { p = nonnull(btg_device_alloc(bytes)); check(btg_copy_to_device(p, host, bytes)); }
    ~DevBuf() { btg_device_free(p); }
    DevBuf(const DevBuf &) = delete;
};
}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) { std::cerr << "usage: btpipeline <bundle.btd> <out.btd> [options]\n"; return 2; }
    try {
        int device = 0;
        bool joint = false;
        uint32_t max_hap = 32;
        btg_gibbs_opts o{};
        o.random_seed = 20190401; o.gibbs_burn_in = 100; o.gibbs_samples = 250; o.n_chains = 20;       // main.cpp:389-403
        o.kmer_subsampling_rate = 0.1f; o.max_haplotype_variant_kmers = 500; o.min_genotype_posterior = 0.99f; o.min_number_of_kmers = 1.0f;
        for (int i = 3; i < argc; i++) {
            const std::string a = argv[i];
            auto val = [&]() -> std::string { if (i + 1 >= argc) throw btg::Error("missing value for " + a); return argv[++i]; };
            if (a == "--device") device = std::stoi(val());
            else if (a == "--random-seed" || a == "-r") o.random_seed = (uint32_t)std::stoul(val());
            else if (a == "--gibbs-burn-in") o.gibbs_burn_in = (uint16_t)std::stoul(val());
            else if (a == "--gibbs-samples") o.gibbs_samples = (uint16_t)std::stoul(val());
            else if (a == "--number-of-gibbs-chains") o.n_chains = (uint16_t)std::stoul(val());
            else if (a == "--max-number-of-sample-haplotypes") max_hap = (uint32_t)std::stoul(val());
            else if (a == "--noise-genotyping") joint = true;
            else throw btg::Error("unknown option " + a);
        }
        const auto b = read_btd(argv[1]);
        check(btg_init(device));
        const Array &genders = get(b, "meta.genders"), &pl = get(b, "meta.ploidy");
        const uint32_t S = (uint32_t)genders.count();
        const uint32_t ploidy_f = pl.as<uint32_t>()[0], ploidy_m = pl.as<uint32_t>()[1];
        const Array &gco = get(b, "g.group_cluster_off"), &cvo = get(b, "g.cl_vertex_off");
        const uint32_t G = (uint32_t)gco.count() - 1, C = (uint32_t)cvo.count() - 1;

        // ---- samples: records to HBM, KmerBloom (bayesTyperTools makeBloom) ----
        std::vector<std::unique_ptr<DevBuf>> d_kmers, d_counts;
        std::vector<size_t> n_rec;
        std::vector<btg_bloom *> blooms;
        for (uint32_t s = 0; s < S; s++) {
            const Array &km = get(b, "s" + std::to_string(s) + ".kmers"), &ct = get(b, "s" + std::to_string(s) + ".counts");
            n_rec.push_back(ct.count());
            d_kmers.emplace_back(new DevBuf(km.bytes.data(), km.bytes.size()));
            d_counts.emplace_back(new DevBuf(ct.bytes.data(), ct.bytes.size()));
            btg_bloom *bl = nonnull(btg_bloom_create(ct.count(), 0.001f, BTG_KMER_SIZE));
            check(btg_bloom_insert_dev(bl, (const uint64_t *)d_kmers.back()->p, ct.count(), nullptr));
            blooms.push_back(bl);
        }
        // ---- findVariantClusterPaths ----
        std::vector<uint32_t> cl_group(C);
        for (uint32_t g = 0; g < G; g++) for (uint64_t c = gco.as<uint64_t>()[g]; c < gco.as<uint64_t>()[g + 1]; c++) cl_group[c] = g;
        btg_graphs_desc gd{};
        gd.n_clusters = C;
        gd.cl_vertex_off = cvo.as<uint64_t>(); gd.v_seq_off = get(b, "g.v_seq_off").as<uint64_t>(); gd.seq = get(b, "g.seq").as<uint8_t>();
        gd.v_flags = get(b, "g.v_flags").as<uint8_t>(); gd.v_in_off = get(b, "g.v_in_off").as<uint64_t>(); gd.v_in_src = get(b, "g.v_in_src").as<uint32_t>();
        gd.cl_group = cl_group.data(); gd.cl_idx = get(b, "g.cluster_idx").as<uint32_t>();
        btg_graphs *gr = nonnull(btg_graphs_upload(&gd, S, max_hap));
        check(btg_find_sample_paths_batch(gr, blooms.data(), 0, S, o.random_seed, max_hap));   // all samples' filters are resident: one launch (KmerCounter.cpp:70-103 per sample)
        std::vector<uint32_t> n_paths(C);
        std::vector<uint64_t> path_off((size_t)C + 1);
        check(btg_get_best_paths(gr, n_paths.data(), path_off.data(), nullptr, 0));
        std::vector<uint8_t> path_mem(path_off[C]);
        check(btg_get_best_paths(gr, n_paths.data(), path_off.data(), path_mem.data(), path_mem.size()));
        btg_graphs_free(gr);
        for (auto bl : blooms) btg_bloom_free(bl);
        // ---- KmerCounter stages ----
        btg_counter_desc cd{};
        cd.n_samples = S; cd.n_groups = G; cd.n_clusters = C;
        cd.sample_gender = genders.as<uint8_t>();
        cd.group_cluster_off = gco.as<uint64_t>(); cd.group_src_off = get(b, "g.group_src_off").as<uint64_t>(); cd.group_src = get(b, "g.group_src").as<uint32_t>();
        cd.group_edge_off = get(b, "g.group_edge_off").as<uint64_t>(); cd.group_edge_src = get(b, "g.group_edge_src").as<uint32_t>(); cd.group_edge_dst = get(b, "g.group_edge_dst").as<uint32_t>();
        cd.cluster_idx = get(b, "g.cluster_idx").as<uint32_t>();
        cd.cl_vertex_off = cvo.as<uint64_t>(); cd.v_seq_off = gd.v_seq_off; cd.seq = gd.seq; cd.v_flags = gd.v_flags;
        cd.v_var = get(b, "g.v_var").as<uint16_t>(); cd.v_allele = get(b, "g.v_allele").as<uint16_t>();
        cd.v_refvar_off = get(b, "g.v_refvar_off").as<uint64_t>(); cd.v_refvar = get(b, "g.v_refvar").as<uint16_t>();
        cd.v_nested = b.count("g.v_nested") ? get(b, "g.v_nested").as<uint32_t>() : nullptr;
        cd.n_paths = n_paths.data(); cd.path_mem = path_mem.data();
        cd.cl_var_off = get(b, "g.cl_var_off").as<uint64_t>(); cd.var_nalleles = get(b, "g.var_nalleles").as<uint16_t>(); cd.var_dep = get(b, "g.var_dep").as<uint8_t>();
        btg_counter *kc = nonnull(btg_counter_create(&cd));
        uint64_t n_path_kmers = 0;
        check(btg_counter_count_path_kmers(kc, &n_path_kmers));
        const Array &regions = get(b, "regions");
        DevBuf d_regions(regions.bytes.data(), regions.bytes.size());
        check(btg_counter_count_intercluster_kmers(kc, (const char *)d_regions.p, regions.bytes.size(), 0, ploidy_f, ploidy_m));
        for (uint32_t s = 0; s < S; s++) check(btg_counter_parse_sample_kmers(kc, s, (const uint64_t *)d_kmers[s]->p, (const uint8_t *)d_counts[s]->p, n_rec[s]));
        std::vector<uint8_t> group_ploidy((size_t)G * S);
        for (uint32_t g = 0; g < G; g++) for (uint32_t s = 0; s < S; s++) group_ploidy[(size_t)g * S + s] = (uint8_t)(genders.as<uint8_t>()[s] == 0 ? ploidy_f : ploidy_m);
        btg_unit *unit = nonnull(btg_counter_build_unit(kc, nullptr, group_ploidy.data()));
        std::vector<const uint64_t *> kp(S); std::vector<const uint8_t *> cp(S);
        for (uint32_t s = 0; s < S; s++) { kp[s] = (const uint64_t *)d_kmers[s]->p; cp[s] = (const uint8_t *)d_counts[s]->p; }
        std::vector<double> nb_p(S), nb_size(S);
        const bool has_pk = b.count("parameter_kmers") != 0;
        check(btg_counter_fit_nb(kc, (const char *)d_regions.p, regions.bytes.size(), ploidy_f, ploidy_m, kp.data(), cp.data(), n_rec.data(),
                                 has_pk ? get(b, "parameter_kmers").as<uint64_t>() : nullptr, has_pk ? get(b, "parameter_kmers").count() / 2 : 0, o.random_seed, 1000000, nb_p.data(),
                                 nb_size.data(), nullptr, nullptr));
        btg_counter_free(kc);
        // ---- Gibbs ----
        for (uint32_t s = 0; s < S; s++) {   // Filters ctor (Filters.cpp:42-53): 1 - exp(-0.275 * NB mean) in float
            const double mean = nb_size[s] * (1 - nb_p[s]) / nb_p[s];
            o.min_fraction_observed_kmers[s] = (float)(1 - std::exp(-(0.275f * mean)));
        }
        btg::CountDistribution count_dist(nb_p, nb_size, 1.0f, 0.01f);
        btg::GenotypeArrays res(S, cd.var_nalleles, cd.cl_var_off[C]);
        if (joint) check(btg_estimate_noise_and_genotypes(unit, count_dist.handle(), &o, &res.view, nullptr));
        else { check(btg_estimate_noise(unit, count_dist.handle(), &o, nullptr)); check(btg_estimate_genotypes(unit, count_dist.handle(), &o, &res.view)); }
        std::vector<double> rates(S);
        check(btg_count_dist_get_noise_rates(count_dist.handle(), rates.data()));
        btg_unit_free(unit);
        BtdWriter w(argv[2]);
        w.put("gt", 1, res.gt); w.put("gq", 2, res.gq); w.put("gpp", 5, res.gpp); w.put("app", 5, res.app); w.put("nak", 5, res.nak); w.put("fak", 5, res.fak);
        w.put("mac", 5, res.mac); w.put("saf", 1, res.saf); w.put("an", 2, res.an); w.put("ac", 2, res.ac); w.put("acp", 5, res.acp);
        w.put("nb_p", 6, nb_p); w.put("nb_size", 6, nb_size); w.put("noise_rates", 6, rates);
        std::vector<uint64_t> meta{n_path_kmers, C, btg_launch_count()};
        w.put("meta", 3, meta);
        std::cerr << "btpipeline: " << C << " clusters, " << n_path_kmers << " path k-mers, " << btg_launch_count() << " kernel launches\n";
        btg_shutdown();
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "btpipeline: " << e.what() << "\n";
        return 1;
    }
}
