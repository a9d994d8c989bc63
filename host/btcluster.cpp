// btcluster — the host-side front end of `bayesTyper cluster`: genome FASTA (+ decoy FASTA) and candidate VCF in, the unit's variant
// clusters, groups and graphs out as a BTD1 file with the keys of bayestyper_b200/graph_builder.build_genome_graphs
// (include/btgpu_cluster.hpp does the work; file reading follows Chromosomes::parseFasta, Chromosomes.cpp:72-117, and
// VariantFileParser's line reader, VariantFileParser.cpp:67-167: CHROM POS ID REF ALT and INFO ACO, `.vcf` or `.vcf.gz`).
//
//   btcluster <genome.fa> <candidates.vcf[.gz]> <out.btd> [--decoy <decoy.fa>] [--min-unit-variants N] [--max-allele-length N] [--copy-number-variant-threshold X]
// --regions-prefix P also writes P.txt.gz, the reference's <out>_cluster_data/intercluster_regions.txt.gz.
// With --min-unit-variants the inference units go to <out>_unit_<i>.btd (i = 1..), as the reference's <prefix>_unit_<i>/ directories.
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "btd.hpp"
#include "btgpu_cluster.hpp"

using btg::cluster::Candidate;

static bool readLine(gzFile f, std::string &line) {
    line.clear();
    char buf[1 << 16];
    while (gzgets(f, buf, sizeof buf)) {
        line += buf;
        if (!line.empty() && line.back() == '\n') { line.pop_back(); if (!line.empty() && line.back() == '\r') line.pop_back(); return true; }
    }
    return !line.empty();
}

static void readFasta(const std::string &path, bool decoy, std::vector<std::string> &names, std::vector<std::string> &seqs, std::vector<uint8_t> &flags) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw btg::Error("cannot open " + path);
    std::string line;
    bool any = false;
    while (readLine(f, line)) {
        if (!line.empty() && line[0] == '>') {
            const size_t e = line.find_first_of("\t ");
            std::string name = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
            if (name.empty()) throw btg::Error(path + ": empty contig name");
            for (auto &n : names) if (n == name) throw btg::Error(path + ": contig " + name + " appears twice");
            names.push_back(name); seqs.emplace_back(); flags.push_back(decoy);
            any = true;
        } else {
            if (!any) throw btg::Error(path + " does not start with a '>' line");
            seqs.back() += line;
        }
    }
    gzclose(f);
}

static std::vector<std::string> split(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t a = 0;
    while (true) {
        const size_t b = s.find(sep, a);
        out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}

static std::vector<std::pair<std::string, std::vector<Candidate>>> readCandidates(const std::string &path) {
    const bool gz = path.size() > 7 && path.compare(path.size() - 7, 7, ".vcf.gz") == 0;
    if (!gz && !(path.size() > 4 && path.compare(path.size() - 4, 4, ".vcf") == 0)) throw btg::Error("variant file needs to end in .vcf or .vcf.gz");
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw btg::Error("cannot open " + path);
    std::vector<std::pair<std::string, std::vector<Candidate>>> out;
    std::string line;
    bool header = false;
    while (readLine(f, line)) {
        if (line.empty()) continue;
        if (line[0] == '#') {
            if (line.compare(0, 6, "#CHROM") == 0) {
                if (std::count(line.begin(), line.end(), '\t') + 1 < 8) throw btg::Error("variant file header has fewer than 8 columns");
                header = true;
            }
            continue;
        }
        if (!header) throw btg::Error("variant file has no #CHROM header line");
        auto t = split(line, '\t');
        if (t.size() < 8) throw btg::Error("variant line with fewer than 8 columns");
        if (t[3].find(',') != std::string::npos) throw btg::Error("REF holds several alleles");
        Candidate c;
        if (t[1].empty() || t[1].find_first_not_of("0123456789") != std::string::npos || t[1].size() > 9 || std::stoul(t[1]) == 0)
            throw btg::Error("POS of the variant line \"" + t[0] + "\t" + t[1] + "\t" + t[2] + " ...\" is not a positive integer");
        c.pos = (uint32_t)std::stoul(t[1]) - 1; c.id = t[2]; c.ref = t[3]; c.alts = split(t[4], ',');
        for (auto &kv : split(t[7], ';'))
            if (kv.compare(0, 4, "ACO=") == 0) {
                c.aco = split(kv.substr(4), ',');
                if (c.aco.size() != c.alts.size()) throw btg::Error("ACO lists a different number of origins than there are alternative alleles at " + t[0] + ":" + t[1]);
                break;
            }
        if (out.empty() || out.back().first != t[0]) {
            for (auto &cv : out) if (cv.first == t[0]) throw btg::Error("variants need to be sorted by contig; variants on contig \"" + t[0] + "\" are unordered");
            out.emplace_back(t[0], std::vector<Candidate>());
        }
        out.back().second.push_back(std::move(c));
    }
    gzclose(f);
    return out;
}

int main(int argc, char **argv) {
    try {
        if (argc < 4) { std::fprintf(stderr, "usage: btcluster <genome.fa> <candidates.vcf[.gz]> <out.btd> [--decoy <decoy.fa>] [--min-unit-variants N] [--regions-prefix P] [--max-allele-length N] [--copy-number-variant-threshold X]\n"); return 2; }
        btg::cluster::Options opt;
        std::string decoy, regions_prefix;
        uint32_t min_unit = 0;          // 0: one unit
        for (int i = 4; i + 1 < argc; i += 2) {
            if (!std::strcmp(argv[i], "--decoy")) decoy = argv[i + 1];
            else if (!std::strcmp(argv[i], "--regions-prefix")) regions_prefix = argv[i + 1];
            else if (!std::strcmp(argv[i], "--min-unit-variants")) min_unit = (uint32_t)std::stoul(argv[i + 1]);
            else if (!std::strcmp(argv[i], "--max-allele-length")) opt.max_allele_length = (uint32_t)std::stoul(argv[i + 1]);
            else if (!std::strcmp(argv[i], "--copy-number-variant-threshold")) opt.copy_number_variant_threshold = (float)std::stod(argv[i + 1]);
            else throw btg::Error(std::string("unknown option ") + argv[i]);
        }
        std::vector<std::string> names, seqs;
        std::vector<uint8_t> flags;
        readFasta(argv[1], false, names, seqs, flags);
        if (!decoy.empty()) readFasta(decoy, true, names, seqs, flags);
        auto cand = readCandidates(argv[2]);
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<btg::cluster::Graphs> units;
        if (min_unit) units = btg::cluster::buildGenomeUnits(names, seqs, flags, cand, min_unit, opt);
        else units.push_back(btg::cluster::buildGenomeGraphs(names, seqs, flags, cand, opt));
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::string joined;
        for (size_t i = 0; i < names.size(); i++) { if (i) joined += '\n'; joined += names[i]; }
        std::string stem = argv[3];
        if (stem.size() > 4 && stem.compare(stem.size() - 4, 4, ".btd") == 0) stem.resize(stem.size() - 4);
        size_t n_var = 0, n_cl = 0, n_gr = 0;
        for (size_t u = 0; u < units.size(); u++) {
            const auto &g = units[u];
            btd::BtdWriter w(min_unit ? stem + "_unit_" + std::to_string(u + 1) + ".btd" : std::string(argv[3]));
            w.put("contig_names", 0, (const uint8_t *)joined.data(), {(uint64_t)joined.size()});
            w.put("group_cluster_off", 3, g.group_cluster_off); w.put("group_nvar", 2, g.group_nvar);
            w.put("group_src_off", 3, g.group_src_off); w.put("group_src", 2, g.group_src);
            w.put("group_edge_off", 3, g.group_edge_off); w.put("group_edge_src", 2, g.group_edge_src); w.put("group_edge_dst", 2, g.group_edge_dst);
            w.put("group_start", 2, g.group_start); w.put("group_end", 2, g.group_end); w.put("group_contig", 2, g.group_contig);
            w.put("cluster_idx", 2, g.cluster_idx);
            w.put("cl_vertex_off", 3, g.cl_vertex_off); w.put("cl_var_off", 3, g.cl_var_off);
            w.put("v_seq_off", 3, g.v_seq_off); w.put("seq", 0, g.seq); w.put("v_flags", 0, g.v_flags); w.put("v_var", 1, g.v_var); w.put("v_allele", 1, g.v_allele);
            w.put("v_nested", 2, g.v_nested); w.put("v_refvar_off", 3, g.v_refvar_off); w.put("v_refvar", 1, g.v_refvar);
            w.put("v_in_off", 3, g.v_in_off); w.put("v_in_src", 2, g.v_in_src);
            w.put("var_pos", 2, g.var_pos); w.put("var_dep", 0, g.var_dep); w.put("var_nalt", 1, g.var_nalt); w.put("var_contig", 2, g.var_contig);
            w.put("var_input_idx", 7, g.var_input_idx); w.put("var_alt_off", 3, g.var_alt_off); w.put("alt_reflen", 2, g.alt_reflen);
            w.put("alt_seq_off", 3, g.alt_seq_off); w.put("alt_seq", 0, (const uint8_t *)g.alt_seq.data(), {(uint64_t)g.alt_seq.size()});
            w.put("alt_aco_off", 3, g.alt_aco_off); w.put("alt_aco", 0, (const uint8_t *)g.alt_aco.data(), {(uint64_t)g.alt_aco.size()});
            w.put("var_id_off", 3, g.var_id_off); w.put("var_ids", 0, (const uint8_t *)g.var_ids.data(), {(uint64_t)g.var_ids.size()});
            std::vector<int64_t> regions;      // of the whole genome: with the first unit
            for (auto &r : units.front().regions) { regions.push_back(r.contig); regions.push_back(r.decoy); regions.push_back(r.start); regions.push_back(r.end); }
            w.put("regions", 7, regions.data(), {(uint64_t)units.front().regions.size(), 4});
            n_var += g.var_pos.size(); n_cl += g.cluster_idx.size(); n_gr += g.group_nvar.size();
        }
        if (!regions_prefix.empty()) {
            // <prefix>.txt.gz as VariantFileParser::sortInterclusterRegions + writeInterclusterRegions leave it (VariantFileParser.cpp:59-65,1181-1212):
            // the same std::sort over the same sequence of regions with the same comparison, so regions of equal length come out in the same order
            std::vector<btg::cluster::Region> regs = units.front().regions;
            std::sort(regs.begin(), regs.end(), [](const btg::cluster::Region &a, const btg::cluster::Region &b) { return (a.end - a.start) > (b.end - b.start); });
            gzFile out = gzopen((regions_prefix + ".txt.gz").c_str(), "wb");
            if (!out) throw btg::Error("cannot write " + regions_prefix + ".txt.gz");
            for (auto &r : regs) gzprintf(out, "%s\t%d\t%u\t%u\n", names[r.contig].c_str(), int(r.decoy), r.start, r.end);
            gzclose(out);
        }
        std::fprintf(stderr, "btcluster: %zu variants as %zu clusters in %zu groups and %zu unit(s), %zu intercluster regions (%.3f s)\n", n_var, n_cl, n_gr, units.size(),
                     units.front().regions.size(), dt);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "btcluster: %s\n", e.what());
        return 1;
    }
}
