// vcf_desc.hpp — the "vcf.*" arrays of a BTD1 container (variant / contig / header description for the VCF writer) -> btgpu_vcf.hpp structures.
#pragma once
#include "btd.hpp"
#include "btgpu_vcf.hpp"

namespace vcfdesc {
using namespace btd;

inline const Array &need(const std::map<std::string, Array> &m, const std::string &name, int dtype = -1) {
    auto it = m.find(name);
    if (it == m.end()) throw btg::Error("input lacks '" + name + "'");
    if (dtype >= 0 && it->second.dtype != dtype) throw btg::Error("'" + name + "' has the wrong type");
    return it->second;
}
inline std::string text(const std::map<std::string, Array> &m, const std::string &name) {
    const Array &a = need(m, name, 0);
    return std::string((const char *)a.bytes.data(), a.bytes.size());
}
inline std::vector<std::string> strings(const std::map<std::string, Array> &m, const std::string &name) {
    const Array &b = need(m, name, 0), &o = need(m, name + "_off", 3);
    std::vector<std::string> out;
    for (uint64_t i = 0; i + 1 < o.count(); i++) out.emplace_back((const char *)b.bytes.data() + o.as<uint64_t>()[i], o.as<uint64_t>()[i + 1] - o.as<uint64_t>()[i]);
    return out;
}

struct Description {
    std::vector<std::string> sample_names;
    std::vector<btg::VcfContig> contigs;
    std::vector<btg::VcfVariant> variants;
    std::vector<uint16_t> nalleles;
    std::string header;
};

inline bool present(const std::map<std::string, Array> &in) { return in.count("vcf.ids") != 0; }

inline Description load(const std::map<std::string, Array> &in, uint32_t S) {
    Description d;
    d.sample_names = strings(in, "vcf.sample_names");
    if (d.sample_names.size() != S) throw btg::Error("vcf.sample_names does not match the number of samples");
    {
        const auto names = strings(in, "vcf.contig_names"), seqs = strings(in, "vcf.contig_seq");
        const Array &dec = need(in, "vcf.contig_decoy", 0);
        if (names.size() != seqs.size() || dec.count() != names.size()) throw btg::Error("contig arrays disagree");
        for (size_t i = 0; i < names.size(); i++) d.contigs.push_back(btg::VcfContig{names[i], seqs[i], dec.as<uint8_t>()[i] != 0});
    }
    const auto ids = strings(in, "vcf.ids"), vcr = strings(in, "vcf.vcr"), vcgr = strings(in, "vcf.vcgr"), alt_seq = strings(in, "vcf.alt_seq"), alt_aco = strings(in, "vcf.alt_aco");
    const uint64_t nv = ids.size();
    const Array &contig = need(in, "vcf.contig", 2), &pos = need(in, "vcf.position", 2), &dep = need(in, "vcf.has_dependency", 0), &vcs = need(in, "vcf.vcs", 2),
                &vcgs = need(in, "vcf.vcgs", 2), &alt_off = need(in, "vcf.alt_off", 3), &alt_len = need(in, "vcf.alt_ref_length", 2);
    if (contig.count() != nv || pos.count() != nv || dep.count() != nv || vcs.count() != nv || vcgs.count() != nv || alt_off.count() != nv + 1 || vcr.size() != nv || vcgr.size() != nv)
        throw btg::Error("per-variant arrays disagree");
    d.variants.resize(nv);
    d.nalleles.resize(nv);
    for (uint64_t i = 0; i < nv; i++) {
        btg::VcfVariant &v = d.variants[i];
        v.contig = contig.as<uint32_t>()[i];
        if (v.contig >= d.contigs.size()) throw btg::Error("variant on an unknown contig");
        v.position = pos.as<uint32_t>()[i]; v.id = ids[i]; v.has_dependency = dep.as<uint8_t>()[i] != 0;
        v.variant_cluster_size = vcs.as<uint32_t>()[i]; v.variant_cluster_group_size = vcgs.as<uint32_t>()[i];
        v.variant_cluster_region = vcr[i]; v.variant_cluster_group_region = vcgr[i];
        for (uint64_t a = alt_off.as<uint64_t>()[i]; a < alt_off.as<uint64_t>()[i + 1]; a++)
            v.alt_alleles.push_back(btg::VcfAltAllele{alt_len.as<uint32_t>()[a], alt_seq.at(a), alt_aco.at(a)});
        if (v.alt_alleles.empty()) throw btg::Error("variant without alternative alleles");
        d.nalleles[i] = (uint16_t)v.numberOfAlleles();
    }
    d.header = btg::vcfHeader(text(in, "vcf.genome_filename"), d.contigs, text(in, "vcf.graph_options_header"), text(in, "vcf.genotype_options_header"), d.sample_names);
    return d;
}

}  // namespace vcfdesc
