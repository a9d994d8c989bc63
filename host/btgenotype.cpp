// btgenotype — C++ host for the Gibbs path over the C ABI: the stage order of `bayesTyper genotype` after the k-mer
// stages (src/bayesTyper/main.cpp:618-643) for one inference unit given as flat haplotype-candidate descriptors.
//
//   btgenotype <unit.btd> <out.btd> [--device D] [--random-seed R] [--gibbs-burn-in B] [--gibbs-samples N]
//              [--number-of-gibbs-chains C] [--kmer-subsampling-rate F] [--max-haplotype-variant-kmers M]
//              [--noise-genotyping] [--noise-rates r0,r1,...] [--min-genotype-posterior P] [--min-number-of-kmers K]
//              [--disable-observed-kmers] [--vcf out.vcf] [--output-prefix P]
//
// Option names and defaults are the reference's (main.cpp:378-403).  <unit.btd> is the BTD1 named-array container
// (bayestyper_b200/btd.py) holding the btg_unit_desc arrays as "unit.<field>", "meta.n_samples" and the per-sample
// negative-binomial parameters "tab.nb_p_size" [S][2]; <out.btd> receives the fields of btg_genotype_result, the
// noise trace and the final noise rates.  All compute happens in libbtgpu (there is no CPU path).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "btgpu.hpp"

#include "btgpu_params.hpp"
#include "vcf_desc.hpp"

namespace {

using namespace btd;

template <class T> const T *field(const std::map<std::string, Array> &m, const std::string &name, uint8_t dtype, uint64_t *n = nullptr) {
    auto it = m.find("unit." + name);
    if (it == m.end()) throw btg::Error("unit descriptor lacks '" + name + "'");
    if (it->second.dtype != dtype) throw btg::Error("unit descriptor field '" + name + "' has the wrong type");
    if (n) *n = it->second.count();
    return it->second.as<T>();
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) { std::cerr << "usage: btgenotype <unit.btd> <out.btd> [options]\n"; return 2; }
    try {
        int device = 0;
        bool joint = false, disable_observed = false;
        std::vector<double> fixed_rates;
        std::string vcf_path, out_prefix;
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
            else if (a == "--kmer-subsampling-rate") o.kmer_subsampling_rate = std::stof(val());
            else if (a == "--max-haplotype-variant-kmers") o.max_haplotype_variant_kmers = (uint32_t)std::stoul(val());
            else if (a == "--min-genotype-posterior") o.min_genotype_posterior = std::stof(val());
            else if (a == "--min-number-of-kmers") o.min_number_of_kmers = std::stof(val());
            else if (a == "--noise-genotyping") joint = true;
            else if (a == "--disable-observed-kmers") disable_observed = true;
            else if (a == "--vcf") vcf_path = val();
            else if (a == "--output-prefix" || a == "-o") out_prefix = val();
            else if (a == "--noise-rates") { std::stringstream ss(val()); std::string t; while (std::getline(ss, t, ',')) fixed_rates.push_back(std::stod(t)); }
            else throw btg::Error("unknown option " + a);
        }
        const auto in = read_btd(argv[1]);
        auto ms = in.find("meta.n_samples");
        auto nb = in.find("tab.nb_p_size");
        if (ms == in.end() || nb == in.end()) throw btg::Error("input lacks meta.n_samples / tab.nb_p_size");
        const uint32_t S = ms->second.dtype == 3 ? (uint32_t)ms->second.as<uint64_t>()[0] : (uint32_t)ms->second.as<uint32_t>()[0];
        if (nb->second.dtype != 6 || nb->second.count() != 2ull * S) throw btg::Error("tab.nb_p_size must be f64 [S][2]");
        std::vector<double> p(S), size(S);
        for (uint32_t s = 0; s < S; s++) { p[s] = nb->second.as<double>()[2 * s]; size[s] = nb->second.as<double>()[2 * s + 1]; }
        if (!disable_observed)  // Filters ctor (src/bayesTyper/Filters.cpp:42-53): 1 - exp(-0.275 * NB mean), float
            for (uint32_t s = 0; s < S; s++) o.min_fraction_observed_kmers[s] = (float)(1 - std::exp(-((double)0.275f * (size[s] * (1 - p[s]) / p[s]))));

        btg_unit_desc d{};
        uint64_t n = 0;
        d.n_samples = S;
        d.sample_gender = field<uint8_t>(in, "sample_gender", 0);
        d.group_ploidy = field<uint8_t>(in, "group_ploidy", 0);
        d.group_cluster_off = field<uint64_t>(in, "group_cluster_off", 3, &n); d.n_groups = (uint32_t)(n - 1);
        d.group_src_off = field<uint64_t>(in, "group_src_off", 3); d.group_src = field<uint32_t>(in, "group_src", 2);
        d.group_edge_off = field<uint64_t>(in, "group_edge_off", 3);
        d.group_edge_src = field<uint32_t>(in, "group_edge_src", 2); d.group_edge_dst = field<uint32_t>(in, "group_edge_dst", 2);
        d.cluster_idx = field<uint32_t>(in, "cluster_idx", 2, &n); d.n_clusters = (uint32_t)n;
        d.cl_nhap = field<uint32_t>(in, "cl_nhap", 2);
        d.cl_kmer_off = field<uint64_t>(in, "cl_kmer_off", 3); d.cl_var_off = field<uint64_t>(in, "cl_var_off", 3); d.cl_mult_off = field<uint64_t>(in, "cl_mult_off", 3);
        d.mult = field<uint8_t>(in, "mult", 0); d.k_has_counts = field<uint8_t>(in, "k_has_counts", 0);
        d.k_counts = field<uint8_t>(in, "k_counts", 0); d.k_ic = field<uint8_t>(in, "k_ic", 0); d.k_shared = field<uint32_t>(in, "k_shared", 2);
        d.cl_uniq_off = field<uint64_t>(in, "cl_uniq_off", 3); d.uniq_idx = field<uint32_t>(in, "uniq_idx", 2);
        d.cl_multi_off = field<uint64_t>(in, "cl_multi_off", 3); d.multi_idx = field<uint32_t>(in, "multi_idx", 2);
        d.kmer_vh_off = field<uint64_t>(in, "kmer_vh_off", 3); d.vh_var = field<uint16_t>(in, "vh_var", 1);
        d.vh_bits_off = field<uint64_t>(in, "vh_bits_off", 3); d.vh_bits = field<uint8_t>(in, "vh_bits", 0);
        d.cl_hapvar_off = field<uint64_t>(in, "cl_hapvar_off", 3); d.hap_alleles = field<uint16_t>(in, "hap_alleles", 1);
        d.var_nalleles = field<uint16_t>(in, "var_nalleles", 1); d.var_dep = field<uint8_t>(in, "var_dep", 0);
        d.hap_nested_off = field<uint64_t>(in, "hap_nested_off", 3); d.hap_nested = field<uint32_t>(in, "hap_nested", 2);
        d.cl_dep_off = field<uint64_t>(in, "cl_dep_off", 3); d.dep_cluster = field<uint32_t>(in, "dep_cluster", 2);
        d.dep_var_off = field<uint64_t>(in, "dep_var_off", 3); d.dep_var = field<uint16_t>(in, "dep_var", 1);

        btg::Library lib(device);
        btg::CountDistribution cd(p, size);
        btg::InferenceUnit unit(d);
        btg::InferenceEngine engine(o);
        btg::GenotypeArrays res = unit.allocResult();
        std::vector<double> trace;
        if (joint) {
            trace = engine.estimateNoiseAndGenotypes(&unit, &cd, &res);
        } else {
            if (fixed_rates.empty()) trace = engine.estimateNoise(&cd, &unit); else cd.setNoiseRates(fixed_rates);
            engine.estimateGenotypes(&unit, cd, &res);
        }
        const std::vector<double> rates = cd.getNoiseRates();
        BtdWriter w(argv[2]);
        w.put("gt", 1, res.gt); w.put("gq", 2, res.gq); w.put("gpp", 5, res.gpp); w.put("app", 5, res.app); w.put("nak", 5, res.nak); w.put("fak", 5, res.fak);
        w.put("mac", 5, res.mac); w.put("saf", 1, res.saf); w.put("ploidy", 0, res.ploidy); w.put("an", 2, res.an); w.put("ac", 2, res.ac); w.put("af", 5, res.af);
        w.put("acp", 5, res.acp); w.put("anc", 0, res.anc); w.put("hc", 1, res.hc);
        w.put("noise_rates", 6, rates);
        if (!trace.empty()) w.put("noise_trace", 6, trace.data(), {(uint64_t)(trace.size() / (2 + S)), (uint64_t)(2 + S)});
        if (!out_prefix.empty()) {  // <prefix>_genomic_parameters.txt, <prefix>_noise_parameters.txt (and <prefix>.vcf when the variants are described)
            std::vector<std::string> names;
            if (in.count("vcf.sample_names")) names = vcfdesc::strings(in, "vcf.sample_names");
            else for (uint32_t s = 0; s < S; s++) names.push_back("S" + std::to_string(s + 1));
            std::ofstream gp(out_prefix + "_genomic_parameters.txt");
            btg::writeGenomicParameters(gp, names, p.data(), size.data());
            if (!trace.empty()) { std::ofstream np(out_prefix + "_noise_parameters.txt"); btg::writeNoiseParameters(np, names, trace.data(), trace.size() / (2 + S)); }
            if (vcf_path.empty() && vcfdesc::present(in)) vcf_path = out_prefix + ".vcf";
        }
        if (!vcf_path.empty()) {  // GenotypeWriter (include/btgpu_vcf.hpp): needs the "vcf.*" description of the unit's variants, in unit order
            const vcfdesc::Description vd = vcfdesc::load(in, S);
            if (vd.variants.size() != res.view.n_variants) throw btg::Error("the vcf.* description does not cover the unit's variants");
            std::ofstream vout(vcf_path);
            if (!vout) throw btg::Error("cannot write " + vcf_path);
            btg::writeVcf(vout, vd.header, vd.variants, vd.contigs, res.view, S);
        }
        std::cout << "btgenotype: " << d.n_clusters << " clusters, " << res.view.n_variants << " variants, " << S << " samples; noise rates";
        for (double r : rates) std::cout << ' ' << r;
        std::cout << std::endl;
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "\nERROR: " << e.what() << "\n" << std::endl;
        return 1;
    }
}
