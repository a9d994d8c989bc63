// btgpu_vcf.hpp — GenotypeWriter for the flat result arrays of the C ABI.
//
// Restates src/bayesTyper/GenotypeWriter.cpp of the reference for btg_genotype_result: the same header
// (generateHeader :494-545), the same record layout (writeGenotypes :84-127), allele sequences (:145-174), QUAL / FILTER
// (:176-202), INFO (:204-262 + VCS/VCR/VCGS/VCGR/HC), FORMAT GT:GQ:GPP:APP:NAK:FAK:MAC:SAF (:264-352, including the
// reference's ":.:.:.:.:.:." for samples without a genotype), records in contig order and sorted by position (:460-481).
// Numbers go through operator<< of a default-constructed ostream exactly as in the reference (6 significant digits);
// NAK / FAK / MAC are float here and double there, which can differ in the sixth digit in rare rounding cases.
// Host-side only: no device code, no dependency on libbtgpu.so.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include "btgpu.h"

namespace btg {

struct VcfAltAllele {
    uint32_t ref_length = 0;   // VariantInfo::AlleleInfo::ref_length
    std::string sequence;      // ... ::sequence
    std::string aco_att;       // ... ::aco_att ("" -> ".")
};

struct VcfVariant {               // VariantInfo + the cluster fields of Genotypes (include/bayesTyper/Genotypes.hpp:46-99)
    uint32_t contig = 0;          // index into the contig list
    uint32_t position = 0;        // 1-based
    std::string id;
    std::vector<VcfAltAllele> alt_alleles;
    bool has_dependency = false;  // adds the '*' allele
    uint32_t variant_cluster_size = 0, variant_cluster_group_size = 0;
    std::string variant_cluster_region, variant_cluster_group_region;
    uint32_t maxReferenceLength() const {  // VariantInfo::maxReferenceLength
        uint32_t m = 0;
        for (auto &a : alt_alleles) m = std::max(m, a.ref_length);
        return m;
    }
    uint32_t numberOfAlleles() const { return 1 + (uint32_t)alt_alleles.size() + (has_dependency ? 1u : 0u); }
};

struct VcfContig {
    std::string name;
    std::string sequence;
    bool decoy = false;
};

namespace vcf_detail {
inline bool floatCompare(float a, float b) {  // Utils.hpp:89-95
    return (a == b) || (std::fabs(a - b) < std::fabs(a < b ? a : b) * (1.1920929e-07f * 100));
}
// the reference holds the genome in upper case (Chromosomes::convertToUpper, main.cpp:213): REF and the allele suffixes taken from a
// soft-masked contig are written in upper case
inline std::string upperCase(std::string s) { for (auto &c : s) if (c >= 'a' && c <= 'z') c = char(c - 'a' + 'A'); return s; }
template <class T> void field(std::ostream &os, const T *v, size_t n) {  // writeAlleleField
    for (size_t i = 0; i < n; i++) { if (i) os << ","; os << v[i]; }
}
}  // namespace vcf_detail

// GenotypeWriter::generateHeader
inline std::string vcfHeader(const std::string &genome_filename, const std::vector<VcfContig> &contigs, const std::string &graph_options_header,
                             const std::string &genotype_options_header, const std::vector<std::string> &sample_names) {
    std::stringstream h;
    h << "##fileformat=VCFv4.2\n";
    h << "##reference=file:" << genome_filename << "\n";
    for (auto &c : contigs) if (!c.decoy) h << "##contig=<ID=" << c.name << ",length=" << c.sequence.size() << ">\n";
    h << graph_options_header << genotype_options_header;
    h << "##FILTER=<ID=AN0,Description=\"No called genotypes (AN = 0)\">\n";
    h << "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Alternative allele counts in called genotypes\">\n";
    h << "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Alternative allele frequencies in called genotypes\">\n";
    h << "##INFO=<ID=AN,Number=1,Type=Integer,Description=\"Total number of alleles in called genotypes\">\n";
    h << "##INFO=<ID=ACP,Number=R,Type=Float,Description=\"Allele call probabilites (maximum APP across samples)\">\n";
    h << "##INFO=<ID=VCS,Number=1,Type=Integer,Description=\"Variant cluster size\">\n";
    h << "##INFO=<ID=VCR,Number=1,Type=String,Description=\"Variant cluster region (<chromosome>:<start>-<end>)\">\n";
    h << "##INFO=<ID=VCGS,Number=1,Type=Integer,Description=\"Variant cluster group size (number of variant clusters)\">\n";
    h << "##INFO=<ID=VCGR,Number=1,Type=String,Description=\"Variant cluster group region (<chromosome>:<start>-<end>)\">\n";
    h << "##INFO=<ID=HC,Number=1,Type=Integer,Description=\"Number of haplotype candidates used for inference in variant cluster\">\n";
    h << "##INFO=<ID=ANC,Number=.,Type=String,Description=\"Allele(s) not covered by a haplotype candidate ('0': Reference allele)\">\n";
    h << "##INFO=<ID=ACO,Number=A,Type=String,Description=\"Alternative allele call-set origin(s) (<call-set>:...)\">\n";
    h << "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    h << "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype quality (phred-scaled 1 - max(GPP))\">\n";
    h << "##FORMAT=<ID=GPP,Number=G,Type=Float,Description=\"Genotype posterior probabilities\">\n";
    h << "##FORMAT=<ID=APP,Number=R,Type=Float,Description=\"Allele posterior probabilities\">\n";
    h << "##FORMAT=<ID=NAK,Number=R,Type=Float,Description=\"Mean number of allele kmers across gibbs samples ('-1': Not sampled)\">\n";
    h << "##FORMAT=<ID=FAK,Number=R,Type=Float,Description=\"Mean fraction of observed allele kmers across gibbs samples ('-1': Not sampled or NAK = 0)\">\n";
    h << "##FORMAT=<ID=MAC,Number=R,Type=Float,Description=\"Mean allele kmer coverage (mean value) across gibbs samples ('-1': Not sampled or NAK = 0)\">\n";
    h << "##FORMAT=<ID=SAF,Number=R,Type=Integer,Description=\"Sample specific allele filter ('0': PASS, '1': NAK, '2': FAK, '3': NAK and FAK)\">\n";
    h << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
    for (auto &s : sample_names) h << "\t" << s;
    h << "\n";
    return h.str();
}

// one record (everything after CHROM POS ID, as the reference assembles it in writeGenotypes + finalise)
inline void vcfRecord(std::ostream &os, const VcfVariant &v, uint64_t vi, const std::vector<VcfContig> &contigs, const btg_genotype_result &r, uint32_t S) {
    using namespace vcf_detail;
    const std::string &chrom = contigs[v.contig].sequence;
    const uint32_t nA = v.numberOfAlleles(), max_ref = v.maxReferenceLength();
    os << contigs[v.contig].name << "\t" << v.position << "\t" << v.id << "\t" << upperCase(chrom.substr(v.position - 1, max_ref)) << "\t";
    for (size_t a = 0; a < v.alt_alleles.size(); a++) {  // writeAlleleSequences
        if (a) os << ",";
        os << v.alt_alleles[a].sequence << upperCase(chrom.substr(v.position + v.alt_alleles[a].ref_length - 1, max_ref - v.alt_alleles[a].ref_length));
    }
    if (v.has_dependency) os << ",*";
    // writeQualityAndFilter: max_alt_allele_call_probability = max ACP over the alternative alleles, the missing ('*') allele
    // of a dependent variant not among them (VariantClusterGenotyper.cpp:509-513)
    const uint64_t o = r.valt_off[vi];
    float max_acp = 0;
    for (uint32_t a = 1; a <= v.alt_alleles.size(); a++) max_acp = std::max(max_acp, r.acp[o + a]);
    if (floatCompare(max_acp, 1)) os << "\t99";
    else if (floatCompare(max_acp, 0)) os << "\t0";
    else os << "\t" << -10 * std::log10(1 - max_acp);
    os << (r.an[vi] == 0 ? "\tAN0" : "\tPASS");
    os << "\tAC="; field(os, r.ac + o + 1, nA - 1);
    os << ";AF="; field(os, r.af + o + 1, nA - 1);
    os << ";AN=" << r.an[vi];
    os << ";ACP="; field(os, r.acp + o, nA);
    os << ";VCS=" << v.variant_cluster_size << ";VCR=" << v.variant_cluster_region << ";VCGS=" << v.variant_cluster_group_size << ";VCGR=" << v.variant_cluster_group_region
       << ";HC=" << r.hc[vi];
    bool first = true;  // writeAlleleCover: non-covered alleles, ascending
    for (uint32_t a = 0; a < nA; a++)
        if (r.anc[o + a]) { os << (first ? ";ANC=" : ",") << a; first = false; }
    os << ";ACO=";
    for (size_t a = 0; a < v.alt_alleles.size(); a++) { if (a) os << ","; os << (v.alt_alleles[a].aco_att.empty() ? std::string(".") : v.alt_alleles[a].aco_att); }
    if (v.has_dependency) os << ",.";
    os << "\tGT:GQ:GPP:APP:NAK:FAK:MAC:SAF";
    const uint32_t nG = nA * (nA + 1) / 2;
    for (uint32_t s = 0; s < S; s++) {  // writeSamples
        os << "\t";
        const uint8_t pl = r.ploidy[vi * S + s];
        if (pl == 0) { os << ":.:.:.:.:.:."; continue; }
        for (uint32_t i = 0; i < pl; i++) {
            if (i) os << "/";
            const uint16_t g = r.gt[(vi * S + s) * 2 + i];
            if (g != 0xFFFF) os << g; else os << ".";
        }
        os << ":" << r.gq[vi * S + s] << ":";
        field(os, r.gpp + r.geno_off[vi] + (uint64_t)s * nG, pl == 2 ? nG : nA);
        const uint64_t ao = r.allele_off[vi] + (uint64_t)s * nA;
        os << ":"; field(os, r.app + ao, nA);
        os << ":"; field(os, r.nak + ao, nA);
        os << ":"; field(os, r.fak + ao, nA);
        os << ":"; field(os, r.mac + ao, nA);
        os << ":"; field(os, r.saf + ao, nA);
    }
    os << "\n";
}

// GenotypeWriter::finalise: header, then the records contig by contig (the order of `contigs`), sorted by position
inline void writeVcf(std::ostream &os, const std::string &header, const std::vector<VcfVariant> &variants, const std::vector<VcfContig> &contigs,
                     const btg_genotype_result &r, uint32_t S) {
    os << header;
    std::vector<std::vector<uint64_t>> by_contig(contigs.size());
    for (uint64_t i = 0; i < variants.size(); i++) by_contig[variants[i].contig].push_back(i);
    for (size_t c = 0; c < contigs.size(); c++) {
        auto &idx = by_contig[c];
        std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return variants[a].position < variants[b].position; });
        for (uint64_t i : idx) vcfRecord(os, variants[i], i, contigs, r, S);
    }
}

}  // namespace btg
