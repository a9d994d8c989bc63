// btref — oracle-R driver.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.
//
// Links the reference's own translation units (compiled in place from
// /root/reference by ./Makefile, Boost replaced by the header shim in shim/) and
// runs `bayesTyper cluster` + `bayesTyper genotype` stage order in one process
// (src/bayesTyper/main.cpp:110-358 and :360-652), skipping only what needs real
// Boost: program_options, filesystem and the variant_clusters.bin archive (the
// InferenceUnit stays in memory between the two commands).  Sample k-mer counts
// are fed from a flat binary file instead of a KMC database: feedSampleKmers()
// below restates the ~10 lines of KmerCounter::parseSampleKmersCallBack
// (src/bayesTyper/KmerCounter.cpp:388-429) without CKMCFile.
//
// Every stage is timed with steady_clock (timings.json) and selected internal
// state is dumped as BTD1 arrays (btd.hpp) for the parity tests.
#include <algorithm>
#include <atomic>
#include <bitset>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <sys/stat.h>

#include "Eigen/Dense"
#include "boost/graph/adjacency_list.hpp"
#include "boost/functional/hash.hpp"
#include "boost/iostreams/filtering_stream.hpp"
#include "boost/algorithm/string.hpp"

#define private public
#define protected public
#include "Utils.hpp"
#include "KmerBloom.hpp"
#include "Kmer.hpp"
#include "Nucleotide.hpp"
#include "OptionsContainer.hpp"
#include "Sample.hpp"
#include "Chromosomes.hpp"
#include "ChromosomePloidy.hpp"
#include "VariantCluster.hpp"
#include "VariantClusterGraph.hpp"
#include "VariantClusterGroup.hpp"
#include "VariantClusterGenotyper.hpp"
#include "VariantClusterHaplotypes.hpp"
#include "VariantFileParser.hpp"
#include "InferenceUnit.hpp"
#include "KmerHash.hpp"
#include "KmerCounter.hpp"
#include "KmerCounts.hpp"
#include "CountDistribution.hpp"
#include "NegativeBinomialDistribution.hpp"
#include "InferenceEngine.hpp"
#include "GenotypeWriter.hpp"
#include "kmc_file.h"
#include "Filters.hpp"
#undef private
#undef protected

#include "btd.hpp"

using namespace std;
typedef bitset<Utils::kmer_size * 2> KmerBits;

static double now_s() {
    return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

struct Timings {
    vector<pair<string, double> > t;
    void add(const string & k, double v) { t.emplace_back(k, v); }
};

struct Args {
    map<string, string> kv;
    bool has(const string & k) const { return kv.count(k) > 0; }
    string str(const string & k, const string & d) const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
    long num(const string & k, long d) const { auto it = kv.find(k); return it == kv.end() ? d : stol(it->second); }
    double flt(const string & k, double d) const { auto it = kv.find(k); return it == kv.end() ? d : stod(it->second); }
};

static Args parseArgs(int argc, char ** argv, int first) {
    Args a;
    for (int i = first; i < argc; i++) {
        string k = argv[i];
        if (k.substr(0, 2) != "--") { cerr << "bad argument " << k << endl; exit(2); }
        k = k.substr(2);
        if (i + 1 < argc && string(argv[i + 1]).substr(0, 2) != "--") { a.kv[k] = argv[++i]; } else { a.kv[k] = "1"; }
    }
    return a;
}

static KmerBits wordsToBits(uint64_t w0, uint64_t w1) {
    KmerBits b(w1);
    b <<= 64;
    b |= KmerBits(w0);
    return b;
}
static void bitsToWords(const KmerBits & b, uint64_t * w) {
    static const KmerBits mask(~0ULL);
    w[0] = (b & mask).to_ullong();
    w[1] = (b >> 64).to_ullong();
}

// ---------------------------------------------------------------------------------------------
// kat: the known-answer vectors of tests/golden/kmer_kat.json, straight from the reference code
// ---------------------------------------------------------------------------------------------
static int cmdKat() {
    const string seq = "AACGTCCGGCATGTTACACATCTACAAACGTGATGGTTGTACCGCATACCACCCTGGGGT";
    KmerPair<Utils::kmer_size> kp;
    cout << "{\n \"seq60\": \"" << seq << "\",\n \"windows\": [\n";
    int w = 0;
    for (size_t i = 0; i < seq.size(); i++) {
        if (kp.move(Nucleotide::ntToBit<1>(seq[i]))) {
            auto low = kp.getLexicographicalLowestKmer();
            string s = Nucleotide::bitToNt<Utils::kmer_size>(low);
            string fw = seq.substr(w, Utils::kmer_size);
            char buf[512];
            snprintf(buf, sizeof(buf), "  {\"w\": %d, \"canonical\": \"%s\", \"fwd\": %d, \"ntp64\": \"%016lx\", \"root\": %lu, \"F\": \"%016lx\", \"R\": \"%016lx\"}%s\n",
                     w, s.c_str(), int(s == fw), NTP64(s.c_str(), Utils::kmer_size), NTP64(s.c_str(), Utils::kmer_size, 1029283129) % 65536,
                     getFhval(fw.c_str(), Utils::kmer_size), getRhval(fw.c_str(), Utils::kmer_size), (i + 1 < seq.size()) ? "," : "");
            cout << buf;
            w++;
        }
    }
    cout << " ],\n \"sizing\": [\n";
    const uint64_t ns[4] = {1, 1000, 1000000, 3000000000ULL};
    for (int i = 0; i < 4; i++) {
        for (float fpr : {0.001f, 0.0001f}) {
            // sizes only: recompute through the same static helpers to avoid allocating 7 GB
            uint64_t bits = KmerBloom<Utils::kmer_size>::calcOptNumBloomBits(fpr, ns[i]);
            uint h = KmerBloom<Utils::kmer_size>::calcOptNumHashes(bits, ns[i]);
            cout << "  {\"n\": " << ns[i] << ", \"fpr\": " << fpr << ", \"bits\": " << bits << ", \"h\": " << h << "}" << ((i == 3 && fpr < 0.0005f) ? "" : ",") << "\n";
            if (ns[i] > 1000000) break;
        }
    }
    cout << " ]\n}" << endl;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// bloom: build/lookup with the reference KmerBloom on a flat k-mer file (parity checks for a4-a6)
//   kmers file: u64 n, then n x (u64 w0, u64 w1)
// ---------------------------------------------------------------------------------------------
static vector<uint64_t> readKmerFile(const string & path, vector<uint8_t> * counts = nullptr) {
    ifstream f(path, ios::binary);
    if (!f.is_open()) { cerr << "cannot open " << path << endl; exit(1); }
    uint64_t n = 0;
    f.read((char *) &n, 8);
    vector<uint64_t> k(2 * n);
    f.read((char *) k.data(), 16 * n);
    if (counts) {
        counts->resize(n);
        f.read((char *) counts->data(), n);
    }
    return k;
}

static int cmdBloom(const Args & a) {
    // --build <kmers> --fpr f --out <prefix>   |   --load <prefix> --query <kmers> --out <hits file>
    if (a.has("build")) {
        auto k = readKmerFile(a.str("build", ""));
        uint64_t n = k.size() / 2;
        KmerBloom<Utils::kmer_size> bloom(a.num("expected", n), (float) a.flt("fpr", 0.001));
        for (uint64_t i = 0; i < n; i++) bloom.addKmer(wordsToBits(k[2 * i], k[2 * i + 1]));
        bloom.save(a.str("out", "bloom"));
        return 0;
    }
    KmerBloom<Utils::kmer_size> bloom(a.str("load", ""));
    auto k = readKmerFile(a.str("query", ""));
    uint64_t n = k.size() / 2;
    vector<uint8_t> hit(n);
    double t0 = now_s();
    for (uint64_t i = 0; i < n; i++) hit[i] = bloom.lookup(wordsToBits(k[2 * i], k[2 * i + 1]));
    double t1 = now_s();
    ofstream o(a.str("out", "hits.bin"), ios::binary);
    o.write((const char *) hit.data(), n);
    cerr << "lookup " << n << " kmers in " << (t1 - t0) << " s" << endl;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// run: cluster + genotype
// ---------------------------------------------------------------------------------------------

// Restatement of KmerCounter::parseSampleKmersCallBack (KmerCounter.cpp:388-429) on a flat
// (k-mer, count) array instead of CKMCFile batches; thread-strided like the other stages.
static void feedSampleKmers(KmerCountsHash * kmer_hash, ThreadedKmerBloom<Utils::kmer_size> * path_kmer_bloom, const vector<uint64_t> & kmers,
                            const vector<uint8_t> & counts, const ushort sample_idx, const ushort num_threads) {
    const uint64_t n = counts.size();
    vector<thread> threads;
    for (ushort t = 0; t < num_threads; t++) {
        threads.emplace_back([&, t]() {
            const uint64_t lo = n * t / num_threads, hi = n * (t + 1) / num_threads;
            for (uint64_t i = lo; i < hi; i++) {
                KmerBits kmer_bitset = wordsToBits(kmers[2 * i], kmers[2 * i + 1]);
                if (path_kmer_bloom->lookup(kmer_bitset)) {
                    auto hash_lock = kmer_hash->getKmerLock(kmer_bitset);
                    auto kmer_counts = kmer_hash->addKmer(kmer_bitset, false);
                    assert(kmer_counts.first);
                    kmer_counts.first->addSampleCount(sample_idx, counts[i]);
                }
            }
        });
    }
    for (auto & th : threads) th.join();
    kmer_hash->sortKmers();
}

static void dumpGraphs(const string & path, InferenceUnit & unit) {
    btd::Writer w(path);
    vector<uint64_t> group_cluster_off{0}, group_src_off{0}, group_edge_off{0};
    vector<uint32_t> group_src, group_edge_src, group_edge_dst, group_nvar, cluster_idx;
    vector<uint64_t> cl_vertex_off{0}, cl_var_off{0}, cl_path_off{0};
    vector<uint64_t> v_seq_off{0}, v_in_off{0}, v_refvar_off{0};
    vector<uint8_t> seq, v_flags, path_bits;
    vector<uint16_t> v_var, v_allele, v_refvar;
    vector<uint32_t> v_nested, v_in_src;
    vector<uint32_t> var_pos;
    vector<uint8_t> var_dep;
    vector<uint16_t> var_nalt;
    vector<uint64_t> var_alt_off{0};
    vector<uint32_t> alt_reflen;
    vector<uint64_t> alt_seq_off{0};
    string alt_seq, var_ids, chroms;
    vector<uint64_t> var_id_off{0}, chrom_off{0};
    for (auto * g : unit.variant_cluster_groups) {
        chroms += g->chrom_name;
        chrom_off.push_back(chroms.size());
        group_nvar.push_back(g->num_variants);
        for (auto s : g->source_vertices) group_src.push_back(s);
        group_src_off.push_back(group_src.size());
        for (uint u = 0; u < g->out_edges.size(); u++)
            for (auto v : g->out_edges[u]) { group_edge_src.push_back(u); group_edge_dst.push_back(v); }
        group_edge_off.push_back(group_edge_src.size());
        for (auto & vx : g->vertices) {
            cluster_idx.push_back(vx.variant_cluster_idx);
            auto & gr = vx.graph->graph;
            const size_t nv = boost::num_vertices(gr);
            for (size_t v = 0; v < nv; v++) {
                auto & p = gr[v];
                for (size_t i = 0; i < p.sequence.size(); i += 2) seq.push_back((uint8_t) p.sequence[i] | ((uint8_t) p.sequence[i + 1] << 1));
                v_seq_off.push_back(seq.size());
                v_flags.push_back((uint8_t) p.is_first_nucleotides_redundant | ((uint8_t) p.is_disconnected << 1));
                v_var.push_back(p.variant_allele_idx.first);
                v_allele.push_back(p.variant_allele_idx.second);
                v_nested.push_back(p.nested_variant_cluster_index);
                for (auto r : p.reference_variant_indices) v_refvar.push_back(r);
                v_refvar_off.push_back(v_refvar.size());
                for (auto & e : gr.in[v]) v_in_src.push_back(e.src);
                v_in_off.push_back(v_in_src.size());
            }
            cl_vertex_off.push_back(v_flags.size());
            for (auto & vi : vx.graph->variant_cluster_info) {
                var_pos.push_back(vi.position);
                var_dep.push_back(vi.has_dependency);
                var_nalt.push_back(vi.alt_alleles.size());
                var_ids += vi.id;
                var_id_off.push_back(var_ids.size());
                for (auto & al : vi.alt_alleles) {
                    alt_reflen.push_back(al.ref_length);
                    alt_seq += al.sequence;
                    alt_seq_off.push_back(alt_seq.size());
                }
                var_alt_off.push_back(alt_reflen.size());
            }
            cl_var_off.push_back(var_pos.size());
            for (size_t i = 0; i < vx.graph->best_paths_indices.size(); i++) path_bits.push_back(vx.graph->best_paths_indices[i]);
            cl_path_off.push_back(path_bits.size());
        }
        group_cluster_off.push_back(cluster_idx.size());
    }
    w.put_str("chroms", chroms); w.put("chrom_off", chrom_off);
    w.put("group_cluster_off", group_cluster_off); w.put("group_nvar", group_nvar);
    w.put("group_src_off", group_src_off); w.put("group_src", group_src);
    w.put("group_edge_off", group_edge_off); w.put("group_edge_src", group_edge_src); w.put("group_edge_dst", group_edge_dst);
    w.put("cluster_idx", cluster_idx);
    w.put("cl_vertex_off", cl_vertex_off); w.put("cl_var_off", cl_var_off); w.put("cl_path_off", cl_path_off);
    w.put("v_seq_off", v_seq_off); w.put("seq", seq); w.put("v_flags", v_flags); w.put("v_var", v_var); w.put("v_allele", v_allele);
    w.put("v_nested", v_nested); w.put("v_refvar_off", v_refvar_off); w.put("v_refvar", v_refvar);
    w.put("v_in_off", v_in_off); w.put("v_in_src", v_in_src);
    w.put("var_pos", var_pos); w.put("var_dep", var_dep); w.put("var_nalt", var_nalt); w.put("var_alt_off", var_alt_off);
    w.put("alt_reflen", alt_reflen); w.put("alt_seq_off", alt_seq_off); w.put_str("alt_seq", alt_seq);
    w.put("var_id_off", var_id_off); w.put_str("var_ids", var_ids);
    w.put("path_bits", path_bits);
}

// VariantClusterHaplotypes of every cluster (getHaplotypeCandidates, VariantClusterGraph.cpp:941-1135)
static void dumpHaplotypes(const string & path, InferenceUnit & unit, KmerCountsHash * kmer_hash, const vector<Sample> & samples) {
    btd::Writer w(path);
    const ushort S = samples.size();
    vector<uint64_t> cl_kmer_off{0}, cl_mult_off{0}, cl_uniq_off{0}, cl_multi_off{0}, cl_hap_off{0}, cl_hapvar_off{0}, kmer_vh_off{0}, vh_bits_off{0};
    vector<uint8_t> mult, k_has_counts, k_counts, k_ic, k_flags, vh_bits, hap_nested_dummy;
    vector<uint32_t> uniq_idx, multi_idx;
    vector<uint16_t> hap_alleles, vh_var;
    vector<uint64_t> kmer_words;
    vector<uint64_t> hap_nested_off{0};
    vector<uint32_t> hap_nested;
    vector<uint64_t> cl_dep_off{0}, dep_var_off{0};
    vector<uint32_t> dep_cluster;
    vector<uint16_t> dep_var;
    for (auto * g : unit.variant_cluster_groups) {
        for (auto & vx : g->vertices) {
            auto h = vx.graph->getHaplotypeCandidates(kmer_hash, 1);
            const size_t K = h.kmers.size(), H = h.haplotypes.size();
            for (size_t k = 0; k < K; k++) {
                for (size_t j = 0; j < H; j++) mult.push_back(h.haplotype_kmer_multiplicities(k, j));
                auto * c = h.kmers[k].counts;
                k_has_counts.push_back(c != nullptr);
                for (ushort s = 0; s < S; s++) k_counts.push_back(c ? c->getSampleCount(s) : 0);
                k_ic.push_back(c ? c->getInterclusterMultiplicity(Utils::Gender::Female) : 0);
                k_ic.push_back(c ? c->getInterclusterMultiplicity(Utils::Gender::Male) : 0);
                k_flags.push_back(c ? (uint8_t) (c->has_cluster_occ | (c->has_multicluster_occ << 1) | (c->has_multigroup_occ << 2) | (c->has_decoy_occ << 3) | (c->has_max_multiplicity << 4) | (c->is_parameter << 5)) : 0);
                for (auto & vh : h.kmers[k].variant_haplotype_indices) {
                    vh_var.push_back(vh.first);
                    for (size_t j = 0; j < H; j++) vh_bits.push_back(vh.second[j]);
                    vh_bits_off.push_back(vh_bits.size());
                }
                kmer_vh_off.push_back(vh_var.size());
            }
            cl_kmer_off.push_back(k_has_counts.size());
            cl_mult_off.push_back(mult.size());
            for (auto i : h.unique_kmer_indices) uniq_idx.push_back(i);
            for (auto i : h.multicluster_kmer_indices) multi_idx.push_back(i);
            cl_uniq_off.push_back(uniq_idx.size());
            cl_multi_off.push_back(multi_idx.size());
            for (auto & hp : h.haplotypes) {
                for (auto a : hp.variant_allele_indices) hap_alleles.push_back(a);
                for (auto nidx : hp.nested_variant_cluster_indices) hap_nested.push_back(nidx);
                hap_nested_off.push_back(hap_nested.size());
            }
            cl_hap_off.push_back(hap_nested_off.size() - 1);
            cl_hapvar_off.push_back(hap_alleles.size());
            vector<uint32_t> dep_keys;
            for (auto & d : h.nested_variant_cluster_dependency) dep_keys.push_back(d.first);
            sort(dep_keys.begin(), dep_keys.end());
            for (auto key : dep_keys) {
                dep_cluster.push_back(key);
                for (auto v : h.nested_variant_cluster_dependency.at(key)) dep_var.push_back(v);
                dep_var_off.push_back(dep_var.size());
            }
            cl_dep_off.push_back(dep_cluster.size());
        }
    }
    // the k-mer keys in row order need a second pass (KmerInfo does not keep them): recompute rows the way
    // getHaplotypeCandidates assigns them (first-seen order over paths), skipping excluded k-mers
    for (auto * g : unit.variant_cluster_groups) {
        for (auto & vx : g->vertices) {
            auto & gr = vx.graph->graph;
            const uint nv = boost::num_vertices(gr);
            const ushort np = vx.graph->best_paths_indices.size() / nv;
            unordered_set<KmerBits> seen;
            KmerPair<Utils::kmer_size> kp;
            bitset<2> nt;
            for (ushort p = 0; p < np; p++) {
                kp.reset();
                for (uint v = 0; v < nv; v++) {
                    if (!vx.graph->best_paths_indices.at(p * nv + v)) continue;
                    if (gr[v].is_disconnected) kp.reset();
                    for (size_t i = 0; i < gr[v].sequence.size(); i += 2) {
                        nt.set(0, gr[v].sequence[i]);
                        nt.set(1, gr[v].sequence[i + 1]);
                        if (kp.move(make_pair(nt, true))) {
                            auto low = kp.getLexicographicalLowestKmer();
                            auto * c = kmer_hash->findKmer(low);
                            if (c && c->isExcluded()) continue;
                            if (seen.insert(low).second) {
                                uint64_t ww[2];
                                bitsToWords(low, ww);
                                kmer_words.push_back(ww[0]);
                                kmer_words.push_back(ww[1]);
                            }
                        }
                    }
                }
            }
        }
    }
    assert(kmer_words.size() == 2 * k_has_counts.size());
    vector<uint32_t> meta{(uint32_t) S};
    w.put("meta", meta);
    w.put("cl_kmer_off", cl_kmer_off); w.put("cl_mult_off", cl_mult_off); w.put("mult", mult);
    w.put("k_has_counts", k_has_counts); w.put("k_counts", k_counts); w.put("k_ic", k_ic); w.put("k_flags", k_flags);
    w.put("kmer_words", kmer_words);
    w.put("kmer_vh_off", kmer_vh_off); w.put("vh_var", vh_var); w.put("vh_bits_off", vh_bits_off); w.put("vh_bits", vh_bits);
    w.put("cl_uniq_off", cl_uniq_off); w.put("uniq_idx", uniq_idx); w.put("cl_multi_off", cl_multi_off); w.put("multi_idx", multi_idx);
    w.put("cl_hap_off", cl_hap_off); w.put("cl_hapvar_off", cl_hapvar_off); w.put("hap_alleles", hap_alleles);
    w.put("hap_nested_off", hap_nested_off); w.put("hap_nested", hap_nested);
    w.put("cl_dep_off", cl_dep_off); w.put("dep_cluster", dep_cluster); w.put("dep_var_off", dep_var_off); w.put("dep_var", dep_var);
}

static void dumpTables(const string & path, CountDistribution & cd, const vector<Sample> & samples) {
    btd::Writer w(path);
    const uint64_t S = samples.size();
    vector<double> genomic(S * 256 * 256), noise(S * 256), nb(S * 2), rates(S);
    for (uint64_t s = 0; s < S; s++) {
        for (int m = 0; m < 256; m++)
            for (int c = 0; c < 256; c++) genomic[(s * 256 + m) * 256 + c] = cd.genomic_count_log_pmf_cache[s][0][m][c];
        for (int c = 0; c < 256; c++) noise[s * 256 + c] = cd.noise_count_log_pmf_cache[s][c];
        nb[2 * s] = cd.genomic_count_distributions[s][0].p();
        nb[2 * s + 1] = cd.genomic_count_distributions[s][0].size();
        rates[s] = cd.noise_rates[s];
    }
    w.put("genomic_log_pmf", genomic.data(), {S, 256, 256});
    w.put("noise_log_pmf", noise.data(), {S, 256});
    w.put("nb_p_size", nb.data(), {S, 2});
    w.put("noise_rates", rates);
}

static void setCommonOptions(OptionsContainer & oc, const Args & a, const string & wd, const string & out_prefix) {
    oc.parseValue<string>("samples-file", wd + "/samples.tsv");
    oc.parseValue<string>("genome-file", wd + "/genome.fa");
    oc.parseValue<string>("decoy-file", a.str("decoy-file", ""));
    oc.parseValue<string>("output-prefix", out_prefix);
    oc.parseValue<uint>("random-seed", (uint) a.num("seed", 20190401));
    oc.parseValue<ushort>("threads", (ushort) a.num("threads", 1));
}

static int cmdRun(const Args & a) {
    Timings tm;
    const string wd = a.str("workdir", ".");
    const string out_dir = wd + "/" + a.str("out", "ref_out");
    mkdir(out_dir.c_str(), 0755);
    const string out_prefix = out_dir + "/bayestyper";
    const ushort num_threads = a.num("threads", 1);
    const uint max_parameter_kmers = 1000000;

    OptionsContainer copt("cluster", BT_VERSION, "00/00/0000 00:00:00");
    copt.parseValue<string>("variant-file", wd + "/variants.vcf");
    setCommonOptions(copt, a, wd, out_prefix);
    copt.parseValue<uint>("min-number-of-unit-variants", 5000000);
    copt.parseValue<uint>("max-allele-length", 500000);
    copt.parseValue<float>("copy-number-variant-threshold", 0.5);
    copt.parseValue<ushort>("max-number-of-sample-haplotypes", (ushort) a.num("max-sample-haplotypes", 32));

    vector<Sample> samples;
    {
        ifstream sf(wd + "/samples.tsv");
        for (string line; getline(sf, line);) if (!line.empty()) samples.emplace_back(line);
    }
    assert(!samples.empty() && samples.size() <= 30);

    double t0 = now_s();
    Chromosomes chromosomes(copt.getValue<string>("genome-file"), false);
    chromosomes.addFasta(copt.getValue<string>("decoy-file"), true);
    chromosomes.convertToUpper();
    tm.add("load_genome", now_s() - t0);

    // sample Bloom filters (bayesTyperTools makeBloom equivalent): built by the reference KmerBloom if absent
    for (auto & s : samples) {
        ifstream probe(s.file + ".bloomMeta");
        if (!probe.is_open()) {
            auto k = readKmerFile(s.file + ".kmers.bin");
            const uint64_t n = k.size() / 2;
            KmerBloom<Utils::kmer_size> bloom(n, 0.001);
            for (uint64_t i = 0; i < n; i++) bloom.addKmer(wordsToBits(k[2 * i], k[2 * i + 1]));
            bloom.save(s.file);
        }
    }

    KmerCounter kmer_counter(samples, copt);
    t0 = now_s();
    VariantFileParser variant_file_parser(copt);
    const uint num_variants = variant_file_parser.getNumberOfVariants();
    InferenceUnit unit(1);
    unit.cluster_options_header = copt.getHeader();
    bool parsed = variant_file_parser.constructVariantClusterGroups(&unit, num_variants, chromosomes);
    assert(parsed);
    sort(unit.variant_cluster_groups.begin(), unit.variant_cluster_groups.end(), VariantClusterGroupCompare);
    tm.add("construct_clusters", now_s() - t0);
    if (a.has("limit-groups")) {
        // bounded CPU-baseline sample: keep every stride-th group (deterministic)
        const size_t limit = a.num("limit-groups", 0);
        if (limit > 0 && limit < unit.variant_cluster_groups.size()) {
            vector<VariantClusterGroup *> keep;
            const double stride = unit.variant_cluster_groups.size() / (double) limit;
            vector<bool> kept(unit.variant_cluster_groups.size(), false);
            for (size_t i = 0; i < limit; i++) kept[(size_t) (i * stride)] = true;
            uint nv = 0, nc = 0;
            for (size_t i = 0; i < kept.size(); i++) {
                if (kept[i]) { keep.push_back(unit.variant_cluster_groups[i]); nv += keep.back()->numberOfVariants(); nc += keep.back()->numberOfVariantClusters(); }
                else delete unit.variant_cluster_groups[i];
            }
            unit.variant_cluster_groups.swap(keep);
            unit.num_variants = nv;
            unit.num_variant_clusters = nc;
        }
    }
    const size_t num_groups = unit.variant_cluster_groups.size();

    const ulong expected_num_path_kmers = ceil((chromosomes.getTotalLength() - chromosomes.getDecoyLength()) * (1 + (0.05 * 2 * samples.size())));
    ThreadedKmerBloom<Utils::kmer_size> * cluster_path_bloom = new ThreadedKmerBloom<Utils::kmer_size>(expected_num_path_kmers, 0.0001);
    KmerHash<bool> multigroup_kmer_hash(ceil(expected_num_path_kmers * 0.01), num_threads);

    t0 = now_s();
    kmer_counter.findVariantClusterPaths(&unit, copt.getValue<ushort>("max-number-of-sample-haplotypes"));
    tm.add("findVariantClusterPaths", now_s() - t0);
    if (a.has("dump-graphs")) dumpGraphs(out_dir + "/graphs.btd", unit);

    t0 = now_s();
    kmer_counter.countPathMultigroupKmers(&multigroup_kmer_hash, cluster_path_bloom, &unit);
    tm.add("countPathMultigroupKmers", now_s() - t0);

    const string cluster_data_dir = out_prefix + "_cluster_data";
    mkdir(cluster_data_dir.c_str(), 0755);
    variant_file_parser.sortInterclusterRegions();
    variant_file_parser.writeInterclusterRegions(cluster_data_dir + "/intercluster_regions");

    const uint max_intercluster_kmers = 3 * max_parameter_kmers;
    const float parameter_kmer_fraction = min(float(1), static_cast<float>(max_intercluster_kmers) / variant_file_parser.getNumberOfInterclusterRegionKmers());
    {
        KmerHash<bool> parameter_kmer_hash(max_intercluster_kmers + chromosomes.getDecoyLength(), num_threads);
        t0 = now_s();
        kmer_counter.countInterclusterParameterKmers(&parameter_kmer_hash, variant_file_parser.getInterclusterRegions(), chromosomes, *cluster_path_bloom, parameter_kmer_fraction);
        tm.add("countInterclusterParameterKmers", now_s() - t0);
        parameter_kmer_hash.shuffle(copt.getValue<uint>("random-seed"));
        parameter_kmer_hash.writeKmersToFasta(cluster_data_dir + "/parameter_kmers", [](bool value) { return value; }, max_parameter_kmers);
    }
    {
        const ulong num_multigroup_kmers = multigroup_kmer_hash.size();
        KmerBloom<Utils::kmer_size> multigroup_kmer_bloom(num_multigroup_kmers, 0.0001);
        multigroup_kmer_hash.addKmersToBloomFilter(&multigroup_kmer_bloom, [](bool value) { return true; });
        multigroup_kmer_bloom.save(cluster_data_dir + "/multigroup_kmers");
        tm.add("num_multigroup_kmers", num_multigroup_kmers);
    }
    delete cluster_path_bloom;
    tm.add("num_path_kmers", unit.num_path_kmers);
    if (a.has("cluster-only")) goto finish;

    {
        // ------------------------------ genotype (main.cpp:489-652) ------------------------------
        OptionsContainer gopt("genotype", BT_VERSION, "00/00/0000 00:00:00");
        gopt.parseValue<string>("variant-clusters-file", out_prefix + "_unit_1/variant_clusters.bin");
        gopt.parseValue<string>("cluster-data-dir", cluster_data_dir);
        setCommonOptions(gopt, a, wd, out_prefix);
        gopt.parseValue<bool>("gzip-output", false);
        gopt.parseValue<string>("chromosome-ploidy-file", a.str("chromosome-ploidy-file", ""));
        gopt.parseValue<ushort>("gibbs-burn-in", (ushort) a.num("gibbs-burn-in", 100));
        gopt.parseValue<ushort>("gibbs-samples", (ushort) a.num("gibbs-samples", 250));
        gopt.parseValue<ushort>("number-of-gibbs-chains", (ushort) a.num("number-of-gibbs-chains", 20));
        gopt.parseValue<float>("kmer-subsampling-rate", (float) a.flt("kmer-subsampling-rate", 0.1));
        gopt.parseValue<uint>("max-haplotype-variant-kmers", (uint) a.num("max-haplotype-variant-kmers", 500));
        gopt.parseValue<bool>("noise-genotyping", a.has("noise-genotyping"));
        gopt.parseValuePair<float>("noise-rate-prior", a.str("noise-rate-prior", "1,0.01"));
        gopt.parseValue<float>("min-genotype-posterior", (float) a.flt("min-genotype-posterior", 0.99));
        gopt.parseValue<float>("min-number-of-kmers", (float) a.flt("min-number-of-kmers", 1));
        gopt.parseValue<bool>("disable-observed-kmers", a.has("disable-observed-kmers"));

        KmerCounter gkmer_counter(samples, gopt);
        ThreadedKmerBloom<Utils::kmer_size> * path_kmer_bloom = new ThreadedKmerBloom<Utils::kmer_size>(unit.num_path_kmers + max_parameter_kmers, 0.0001);
        KmerCountsHash * kmer_hash;
        if (samples.size() < 4) kmer_hash = new ObservedKmerCountsHash<3>(unit.num_path_kmers + max_parameter_kmers, num_threads);
        else if (samples.size() < 11) kmer_hash = new ObservedKmerCountsHash<10>(unit.num_path_kmers + max_parameter_kmers, num_threads);
        else if (samples.size() < 21) kmer_hash = new ObservedKmerCountsHash<20>(unit.num_path_kmers + max_parameter_kmers, num_threads);
        else kmer_hash = new ObservedKmerCountsHash<30>(unit.num_path_kmers + max_parameter_kmers, num_threads);

        uint num_parameter_kmers = 0;
        {
            ifstream kmers_infile(cluster_data_dir + "/parameter_kmers.fa.gz", std::ios::binary);
            boost::iostreams::filtering_istream in;
            in.push(boost::iostreams::gzip_decompressor());
            in.push(boost::ref(kmers_infile));
            string line;
            getline(in, line);
            assert(line == (">k" + to_string(Utils::kmer_size)));
            while (getline(in, line)) {
                num_parameter_kmers++;
                path_kmer_bloom->addKmer(line);
                auto parameter_kmer = Nucleotide::ntToBit<Utils::kmer_size>(line);
                assert(parameter_kmer.second);
                auto kmer_counts = kmer_hash->addKmer(parameter_kmer.first, false);
                assert(kmer_counts.first);
                assert(kmer_counts.second);
                kmer_counts.first->isParameter(true);
            }
        }
        kmer_hash->sortKmers();
        tm.add("num_parameter_kmers", num_parameter_kmers);

        auto chrom_ploidy = ChromosomePloidy(gopt.getValue<string>("chromosome-ploidy-file"), chromosomes, samples);
        t0 = now_s();
        gkmer_counter.countPathKmers(path_kmer_bloom, &unit);
        tm.add("countPathKmers", now_s() - t0);
        t0 = now_s();
        gkmer_counter.countInterclusterKmers(kmer_hash, path_kmer_bloom, cluster_data_dir + "/intercluster_regions", chromosomes, chrom_ploidy);
        tm.add("countInterclusterKmers", now_s() - t0);

        double t_parse = 0;
        uint64_t n_sample_kmers = 0;
        for (ushort s = 0; s < samples.size(); s++) {
            vector<uint8_t> counts;
            auto kmers = readKmerFile(samples[s].file + ".kmers.bin", &counts);
            n_sample_kmers += counts.size();
            t0 = now_s();
            feedSampleKmers(kmer_hash, path_kmer_bloom, kmers, counts, s, num_threads);
            t_parse += now_s() - t0;
        }
        tm.add("parseSampleKmers", t_parse);
        tm.add("num_sample_kmers", n_sample_kmers);
        delete path_kmer_bloom;

        t0 = now_s();
        gkmer_counter.classifyPathKmers(kmer_hash, &unit, cluster_data_dir + "/multigroup_kmers");
        tm.add("classifyPathKmers", now_s() - t0);

        auto intercluster_kmer_stats = kmer_hash->calculateKmerStats(samples);
        CountDistribution count_distribution(samples, gopt);
        count_distribution.setGenomicCountDistributions(intercluster_kmer_stats, out_prefix + "_genomic_parameters");

        if (a.has("dump-haps")) {
            t0 = now_s();
            dumpHaplotypes(out_dir + "/haps.btd", unit, kmer_hash, samples);
            tm.add("getHaplotypeCandidates_dump", now_s() - t0);
        }
        if (a.has("skip-genotype")) {
            dumpTables(out_dir + "/tables.btd", count_distribution, samples);
            delete kmer_hash;
            goto finish;
        }

        const bool noise_genotyping = gopt.getValue<bool>("noise-genotyping");
        InferenceEngine inference_engine(samples, chrom_ploidy, gopt);
        if (!noise_genotyping) {
            if (a.has("noise-rates")) {
                // fixed noise rates (comma separated) instead of estimateNoise: lets parity tests share T_s[0][.]
                vector<string> f;
                boost::split(f, a.str("noise-rates", ""), boost::is_any_of(","));
                vector<double> r;
                for (auto & x : f) r.push_back(stod(x));
                assert(r.size() == samples.size());
                count_distribution.setNoiseRates(r);
            } else {
                t0 = now_s();
                inference_engine.estimateNoise(&count_distribution, &unit, kmer_hash, out_prefix + "_noise_parameters");
                tm.add("estimateNoise", now_s() - t0);
            }
        }
        dumpTables(out_dir + "/tables.btd", count_distribution, samples);

        Filters filters(gopt, count_distribution.getGenomicCountDistributions());
        GenotypeWriter genotype_writer(out_prefix, num_threads, samples, chromosomes, filters);
        const uint clusters_to_genotype = unit.num_variant_clusters;
        t0 = now_s();
        if (!noise_genotyping) inference_engine.estimateGenotypes(&unit, kmer_hash, count_distribution, filters, &genotype_writer);
        else inference_engine.estimateNoiseAndGenotypes(&unit, &count_distribution, kmer_hash, filters, &genotype_writer, out_prefix + "_noise_parameters");
        tm.add(noise_genotyping ? "estimateNoiseAndGenotypes" : "estimateGenotypes", now_s() - t0);
        tm.add("clusters_genotyped", clusters_to_genotype);
        delete kmer_hash;
        t0 = now_s();
        genotype_writer.finalise(out_prefix, chromosomes, unit.cluster_options_header, gopt, filters);
        tm.add("write_vcf", now_s() - t0);
        unit.variant_cluster_groups.clear();  // deleted by the engine (InferenceEngine.cpp:127,309)
    }

finish:
    {
        ofstream tj(out_dir + "/timings.json");
        tj << "{\"threads\": " << num_threads << ", \"num_groups\": " << num_groups << ", \"num_clusters\": " << unit.num_variant_clusters
           << ", \"num_variants\": " << unit.num_variants << ", \"num_samples\": " << samples.size();
        for (auto & kv : tm.t) tj << ", \"" << kv.first << "\": " << kv.second;
        tj << "}" << endl;
    }
    for (auto * g : unit.variant_cluster_groups) delete g;
    return 0;
}

// btref kmc-list --db <kmc_prefix>: every (k-mer, count) the reference's vendored KMC API lists (CKMCFile::OpenForListing /
// ReadNextKmer, the calls of KmerCounter::parseSampleKmers and MakeBloom::kmc2bloomThreaded), as "<k-mer>\t<count>" lines, preceded
// by one "#info" line.  Pins include/btgpu_kmc.hpp and bayestyper_b200/kmcio.py (tests/test_kmc.py).
// `cluster`'s unit loop (main.cpp:206-247) without the k-mer stages: the variant-cluster groups of every inference unit the
// reference's parser forms for --min-unit-variants, plus the intercluster regions it accumulated over all units
static int cmdUnits(const Args & a) {
    const string wd = a.str("workdir", ".");
    const string out_dir = wd + "/" + a.str("out", "ref_out");
    mkdir(out_dir.c_str(), 0755);
    OptionsContainer copt("cluster", BT_VERSION, "00/00/0000 00:00:00");
    copt.parseValue<string>("variant-file", wd + "/variants.vcf");
    setCommonOptions(copt, a, wd, out_dir + "/bayestyper");
    copt.parseValue<uint>("min-number-of-unit-variants", (uint) a.num("min-unit-variants", 5000000));
    copt.parseValue<uint>("max-allele-length", 500000);
    copt.parseValue<float>("copy-number-variant-threshold", 0.5);
    Chromosomes chromosomes(copt.getValue<string>("genome-file"), false);
    chromosomes.addFasta(copt.getValue<string>("decoy-file"), true);
    chromosomes.convertToUpper();
    VariantFileParser variant_file_parser(copt);
    const uint num_variants = variant_file_parser.getNumberOfVariants();
    const uint num_units = max(uint(1), static_cast<uint>(floor(num_variants / static_cast<float>(copt.getValue<uint>("min-number-of-unit-variants")))));
    bool parsed = false;
    uint unit_idx = 1;
    for (; unit_idx < (num_units + 1); unit_idx++) {
        assert(!parsed);
        InferenceUnit unit(unit_idx);
        parsed = variant_file_parser.constructVariantClusterGroups(&unit, ceil(num_variants / static_cast<float>(num_units)), chromosomes);
        sort(unit.variant_cluster_groups.begin(), unit.variant_cluster_groups.end(), VariantClusterGroupCompare);
        dumpGraphs(out_dir + "/graphs_unit_" + to_string(unit_idx) + ".btd", unit);
        if (parsed) break;
    }
    assert(parsed);
    variant_file_parser.sortInterclusterRegions();
    variant_file_parser.writeInterclusterRegions(out_dir + "/intercluster_regions");
    cout << "units " << unit_idx << " of " << num_units << endl;
    return 0;
}

static int cmdKmcList(const Args & a) {
    CKMCFile db;
    if (!db.OpenForListing(a.str("db", ""))) { cerr << "cannot open KMC database " << a.str("db", "") << endl; return 1; }
    CKMCFileInfo info;
    db.Info(info);
    cout << "#info kmer_length " << info.kmer_length << " mode " << info.mode << " counter_size " << info.counter_size << " lut_prefix_length " << info.lut_prefix_length
         << " min_count " << info.min_count << " max_count " << info.max_count << " total_kmers " << info.total_kmers << " both_strands " << info.both_strands << "\n";
    CKmerAPI kmer(info.kmer_length);
    uint32 count;
    while (db.ReadNextKmer(kmer, count)) cout << kmer.to_string() << "\t" << count << "\n";
    return 0;
}

int main(int argc, char ** argv) {
    if (argc < 2) {
        cerr << "usage: btref <kat|bloom|run|kmc-list> [--key value ...]" << endl;
        return 2;
    }
    const string cmd = argv[1];
    Args a = parseArgs(argc, argv, 2);
    if (cmd == "kat") return cmdKat();
    if (cmd == "bloom") return cmdBloom(a);
    if (cmd == "run") return cmdRun(a);
    if (cmd == "units") return cmdUnits(a);
    if (cmd == "kmc-list") return cmdKmcList(a);
    cerr << "unknown command " << cmd << endl;
    return 2;
}
