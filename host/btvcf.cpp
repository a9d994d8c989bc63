// btvcf — writes the genotype VCF of `bayesTyper genotype` from the flat result arrays of the C ABI (include/btgpu_vcf.hpp restates
// src/bayesTyper/GenotypeWriter.cpp).  Host-only: no GPU is involved.
//
//   btvcf <in.btd> <out.vcf>
//
// <in.btd> (BTD1 container) holds the fields of btg_genotype_result ("gt", "gq", "gpp", "app", "nak", "fak", "mac", "saf", "ploidy",
// "an", "ac", "af", "acp", "anc", "hc"), "meta.n_samples", and the variant / contig / header description under "vcf.*":
//   strings as byte arrays with uint64 offsets (<name>, <name>_off): sample_names, contig_names, contig_seq, ids, vcr, vcgr,
//   alt_seq, alt_aco; scalars per variant: contig (u32), position (u32, 1-based), has_dependency (u8), vcs (u32), vcgs (u32),
//   alt_off (u64, [n_variants + 1] into the alt arrays), alt_ref_length (u32 per alt allele); contig_decoy (u8 per contig);
//   genome_filename, graph_options_header, genotype_options_header (byte arrays).
#include <iostream>

#include "btgpu_params.hpp"
#include "vcf_desc.hpp"

int main(int argc, char **argv) {
    if (argc != 3) { std::cerr << "usage: btvcf <in.btd> <out.vcf>\n"; return 2; }
    try {
        using namespace vcfdesc;
        const auto in = btd::read_btd(argv[1]);
        const btd::Array &ms = need(in, "meta.n_samples");
        const uint32_t S = ms.dtype == 3 ? (uint32_t)ms.as<uint64_t>()[0] : (uint32_t)ms.as<uint32_t>()[0];
        const Description d = load(in, S);
        const uint64_t nv = d.variants.size();
        btg::GenotypeArrays res(S, d.nalleles.data(), nv);   // offsets (allele_off, geno_off, valt_off) + storage
        auto fill = [&](auto &vec, const char *name, int dtype) {
            const btd::Array &a = need(in, name, dtype);
            if (a.count() != vec.size()) throw btg::Error(std::string("result array '") + name + "' has the wrong length");
            memcpy(vec.data(), a.bytes.data(), a.bytes.size());
        };
        fill(res.gt, "gt", 1); fill(res.gq, "gq", 2); fill(res.gpp, "gpp", 5); fill(res.app, "app", 5); fill(res.nak, "nak", 5); fill(res.fak, "fak", 5);
        fill(res.mac, "mac", 5); fill(res.saf, "saf", 1); fill(res.ploidy, "ploidy", 0); fill(res.an, "an", 2); fill(res.ac, "ac", 2); fill(res.af, "af", 5);
        fill(res.acp, "acp", 5); fill(res.anc, "anc", 0); fill(res.hc, "hc", 1);
        std::ofstream out(argv[2]);
        if (!out) throw btg::Error(std::string("cannot write ") + argv[2]);
        btg::writeVcf(out, d.header, d.variants, d.contigs, res.view, S);
        // optional companions: "noise_trace" [rows][2 + S] and "tab.nb_p_size" [S][2] -> <out minus .vcf>_noise_parameters.txt / _genomic_parameters.txt
        std::string stem = argv[2];
        if (stem.size() > 4 && stem.substr(stem.size() - 4) == ".vcf") stem = stem.substr(0, stem.size() - 4);
        if (in.count("noise_trace")) {
            const btd::Array &t = need(in, "noise_trace", 6);
            std::ofstream np(stem + "_noise_parameters.txt");
            btg::writeNoiseParameters(np, d.sample_names, t.as<double>(), t.count() / (2 + S));
        }
        if (in.count("tab.nb_p_size")) {
            const btd::Array &t = need(in, "tab.nb_p_size", 6);
            std::vector<double> p(S), size(S);
            for (uint32_t s = 0; s < S; s++) { p[s] = t.as<double>()[2 * s]; size[s] = t.as<double>()[2 * s + 1]; }
            std::ofstream gp(stem + "_genomic_parameters.txt");
            btg::writeGenomicParameters(gp, d.sample_names, p.data(), size.data());
        }
        std::cout << "btvcf: wrote " << nv << " variants, " << S << " samples to " << argv[2] << std::endl;
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "\nERROR: " << e.what() << "\n" << std::endl;
        return 1;
    }
}
