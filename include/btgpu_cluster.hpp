// btgpu_cluster.hpp — variant clusters, variant-cluster groups and their graphs from a candidate set.
//
// Native counterpart of bayestyper_b200/graph_builder.py (the two are compared array by array in tests/test_host_cpp.py), i.e. a
// restatement of the reference's host-side `cluster` front end for the flat graph arrays that btg_graphs_upload and the unit take:
//   VariantFileParser::parseVariants / addAlternativeAllele / copyNumberVariantLength / clusterVariants / mergeVariantClusters /
//   getVariantClusterGroupDependencies (src/bayesTyper/VariantFileParser.cpp:241-545,581-733,735-978,1000-1040,1107-1156),
//   the VariantClusterGraph constructor with contained clusters (src/bayesTyper/VariantClusterGraph.cpp:62-377),
//   the VariantClusterGroup constructor and VariantClusterGroupCompare (src/bayesTyper/VariantClusterGroup.cpp:47-105,278-291).
// The cluster order inside a group, the cluster that survives a merge and the order of the dependency edges come from libstdc++'s
// unordered containers in the reference; the same containers are used here with the same sequence of insertions and erasures.
// Host-side only: no device code, no dependency on libbtgpu.so.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace btg {
namespace cluster {

constexpr uint32_t kNone32 = 0xFFFFFFFFu;
constexpr uint16_t kNone16 = 0xFFFFu;

struct Candidate {          // one line of the candidate VCF
    uint32_t pos = 0;       // 0-based position of REF[0]
    std::string ref;
    std::vector<std::string> alts;   // a trailing "*" kept
    std::string id = ".";
    std::vector<std::string> aco;    // per alternative allele, or empty
};

struct Options {
    uint32_t k = 55;
    uint32_t max_allele_length = 500000;
    double copy_number_variant_threshold = 0.5;
};

struct Region { uint32_t contig; bool decoy; uint32_t start, end; };   // inclusive, 0-based

// the unit as flat arrays (names = the keys of graphs.btd)
struct Graphs {
    std::vector<uint64_t> group_cluster_off{0}, group_src_off{0}, group_edge_off{0};
    std::vector<uint32_t> group_nvar, group_src, group_edge_src, group_edge_dst, group_start, group_end, group_contig, cluster_idx;
    std::vector<uint64_t> cl_vertex_off{0}, cl_var_off{0}, v_seq_off{0}, v_in_off{0}, v_refvar_off{0};
    std::vector<uint8_t> seq, v_flags;
    std::vector<uint16_t> v_var, v_allele, v_refvar;
    std::vector<uint32_t> v_nested, v_in_src;
    std::vector<uint32_t> var_pos, var_contig;
    std::vector<uint8_t> var_dep;
    std::vector<uint16_t> var_nalt;
    std::vector<int64_t> var_input_idx;
    std::vector<uint64_t> var_alt_off{0}, alt_seq_off{0}, alt_aco_off{0}, var_id_off{0};
    std::vector<uint32_t> alt_reflen;
    std::string alt_seq, alt_aco, var_ids;
    std::vector<Region> regions;
};

namespace detail {

inline int code(char c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}
inline std::string upper(std::string s) { for (auto &c : s) if (c >= 'a' && c <= 'z') c = char(c - 32); return s; }

inline std::string canonical(const char *p, uint32_t k) {
    std::string f(p, k), r(k, 'N');
    for (uint32_t i = 0; i < k; i++) { const char c = p[k - 1 - i]; r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }
    return f < r ? f : r;
}
inline bool approxEqual(double a, double b) {   // Utils::doubleCompare
    return a == b || std::fabs(a - b) < std::fabs(std::min(a, b)) * std::numeric_limits<double>::epsilon() * 100;
}

// how far downstream of an allele the genome keeps repeating the allele's k-mers: extends the END of the group
inline uint32_t copyNumberVariantLength(const std::string &allele, const std::string &chrom, uint32_t start, const Options &o) {
    const uint32_t k = o.k;
    if (allele.size() < k) return 0;
    std::unordered_set<std::string> kmers;
    uint32_t run = 0;
    for (size_t i = 0; i < allele.size(); i++) {
        run = code(allele[i]) < 4 ? run + 1 : 0;
        if (run >= k) kmers.insert(canonical(allele.data() + i + 1 - k, k));
    }
    if (kmers.empty()) return 0;
    const uint32_t n = (uint32_t)chrom.size();
    uint32_t length = 0;
    uint32_t window_end = (uint32_t)std::min<uint64_t>((uint64_t)start + length + allele.size(), n);
    while (true) {
        uint32_t bases = 0, identical = 0, best_bases = 0;
        double best = 0;
        run = 0;
        for (uint32_t p = start + length; p < window_end; p++) {
            run = code(chrom[p]) < 4 ? run + 1 : 0;
            if (run >= k && kmers.count(canonical(chrom.data() + p + 1 - k, k))) identical++;
            bases++;
            if (identical) {
                const double frac = identical / double(bases - k + 1);
                if (approxEqual(frac, best) || frac > best) { best = frac; best_bases = bases; }
            }
        }
        if (best < o.copy_number_variant_threshold) break;
        length += best_bases;
        if (window_end == n) break;
        window_end = (uint32_t)std::min<uint64_t>((uint64_t)start + length + allele.size(), n);
    }
    return length;
}

struct Allele { uint32_t ref_len; std::string seq, aco; };
struct Variant { int64_t input_idx; bool dep; uint32_t n_red = kNone32; std::vector<Allele> alts; std::string id; };
struct Contained { uint32_t idx, left, right; };
struct Cluster {
    uint32_t idx, left, right;
    std::map<uint32_t, Variant> variants;
    std::vector<Contained> contained;
};
using Group = std::unordered_map<uint32_t, std::unique_ptr<Cluster>>;
using MergeSets = std::list<std::unordered_set<uint32_t>>;

inline int64_t absdiff(uint32_t a, uint32_t b) { return a > b ? int64_t(a) - b : int64_t(b) - a; }

inline void clusterVariant(Variant &&var, uint32_t pos, const std::vector<uint32_t> &ends, std::map<uint32_t, Cluster *> &flanks, Group &group, MergeSets &merge_sets,
                           uint32_t k) {
    while (!flanks.empty() && int64_t(pos) - int64_t(flanks.begin()->first) >= int64_t(k)) flanks.erase(flanks.begin());
    Cluster *first = nullptr;
    std::vector<Cluster *> second;
    auto overlap = [&](Cluster *cl) {
        if (!first) { first = cl; return true; }
        if (first != cl && std::find(second.begin(), second.end(), cl) == second.end()) second.push_back(cl);
        return false;
    };
    for (auto it = flanks.begin(); it != flanks.end();) {
        const uint32_t key = it->first;
        Cluster *cl = it->second;
        if (absdiff(pos, key) + 1 <= int64_t(k)) {
            if (overlap(cl) && pos >= key) { it = flanks.erase(it); continue; }
        }
        for (uint32_t e : ends) {
            if (absdiff(e, key) + 1 <= int64_t(k)) overlap(cl);
            else if (pos < key && key < e) overlap(cl);
        }
        ++it;
    }
    // the reference keeps these in a std::set of POINTERS and feeds the merge set in that order: with two or more extra clusters the
    // surviving index depends on malloc's placement of the clusters; creation order is the deterministic choice made here
    std::sort(second.begin(), second.end(), [](Cluster *a, Cluster *b) { return a->idx < b->idx; });
    const uint32_t last_end = ends.back();
    if (!first) {
        auto cl = std::make_unique<Cluster>();
        cl->idx = (uint32_t)group.size(); cl->left = pos; cl->right = last_end;
        cl->variants.emplace(pos, std::move(var));
        for (uint32_t e : ends) flanks[e] = cl.get();
        if (last_end - pos >= k) flanks.emplace(pos, cl.get());
        const uint32_t idx = cl->idx;
        group.emplace(idx, std::move(cl));
    } else {
        if (!first->variants.emplace(pos, std::move(var)).second)
            throw std::runtime_error("several variants at position " + std::to_string(pos + 1) + ": they need to be one multi-allelic variant");
        first->right = std::max(last_end, first->right);
        for (uint32_t e : ends) flanks.emplace(e, first);
        if (last_end - pos >= k) flanks.emplace(pos, first);
    }
    if (second.empty()) return;
    auto found = merge_sets.end();
    for (auto sit = merge_sets.begin(); sit != merge_sets.end();) {
        if (sit->count(first->idx)) {
            if (found == merge_sets.end()) found = sit;
            else if (sit != found) { for (uint32_t x : *sit) found->insert(x); sit = merge_sets.erase(sit); continue; }
        }
        bool merged = false;
        for (Cluster *cl : second) {
            if (sit->count(cl->idx)) {
                if (found == merge_sets.end()) found = sit;
                else if (sit != found) { for (uint32_t x : *sit) found->insert(x); sit = merge_sets.erase(sit); merged = true; break; }
            }
        }
        if (!merged) ++sit;
    }
    if (found == merge_sets.end()) { merge_sets.emplace_back(); found = std::prev(merge_sets.end()); }
    found->insert(first->idx);
    for (Cluster *cl : second) found->insert(cl->idx);
}

inline void mergeClusters(Group &group, MergeSets &merge_sets) {
    for (auto &s : merge_sets) {
        auto it = s.begin();
        Cluster *keep = group.at(*it).get();
        for (++it; it != s.end(); ++it) {
            Cluster *other = group.at(*it).get();
            keep->left = std::min(keep->left, other->left);
            keep->right = std::max(keep->right, other->right);
            for (auto &pv : other->variants)
                if (!keep->variants.emplace(pv.first, std::move(pv.second)).second) throw std::runtime_error("two clusters to merge hold a variant at the same position");
            group.erase(*it);
        }
    }
}

struct ContigResult { std::vector<Group> groups; std::vector<std::pair<uint32_t, uint32_t>> regions; std::vector<uint8_t> breaks; };

inline ContigResult parseContig(const std::string &reference, const std::vector<Candidate> &variants, const Options &o) {
    const uint32_t k = o.k;
    const std::string chrom = upper(reference);
    const int64_t n = (int64_t)chrom.size();
    ContigResult out;
    Group group;
    MergeSets merge_sets;
    std::map<uint32_t, Cluster *> flanks;
    std::set<uint32_t> dependencies;
    int64_t prev_pos = -1, prev_var_end = -1, group_end = -1;
    auto flush = [&]() {
        if (!group.empty()) { mergeClusters(group, merge_sets); out.groups.push_back(std::move(group)); group = Group(); }
        merge_sets.clear(); flanks.clear();
    };
    auto addRegion = [&](int64_t a, int64_t b) { if (b - a + 1 >= int64_t(k)) out.regions.emplace_back((uint32_t)a, (uint32_t)b); };
    for (size_t vi = 0; vi < variants.size(); vi++) {
        const Candidate &v = variants[vi];
        const int64_t pos = v.pos;
        if (pos < prev_pos) throw std::runtime_error("variants need to be sorted by position: " + std::to_string(prev_pos + 1) + " is before " + std::to_string(pos + 1));
        while (!dependencies.empty() && int64_t(*dependencies.begin()) < pos) dependencies.erase(dependencies.begin());
        out.breaks.push_back(pos - group_end >= int64_t(k));      // an inference unit may end in front of this variant
        prev_pos = pos;
        const std::string ref = upper(v.ref);
        std::vector<std::string> alts;
        for (auto &a : v.alts) alts.push_back(upper(a));
        Variant var;
        var.input_idx = (int64_t)vi; var.dep = !dependencies.empty(); var.id = v.id;
        if (!alts.empty() && alts.back() == "*") {
            if (!var.dep) throw std::runtime_error("'*' allele at position " + std::to_string(pos + 1) + " without an overlapping upstream variant");
            alts.pop_back();
        }
        if (alts.empty()) throw std::runtime_error("variant at position " + std::to_string(pos + 1) + " has no alternative allele");
        for (size_t i = 0; i < alts.size(); i++)
            for (size_t j = i + 1; j < alts.size(); j++)
                if (alts[i] == alts[j]) throw std::runtime_error("duplicate alternative alleles at position " + std::to_string(pos + 1));
        if (pos + (int64_t)ref.size() > n) throw std::runtime_error("variant at position " + std::to_string(pos + 1) + " runs past the end of the contig");
        std::vector<std::pair<std::string, std::string>> pairs;      // right-trimmed (ref, alt)
        for (auto &a : alts) {
            std::string r = ref, b = a;
            while (r.size() > 1 && b.size() > 1 && r.back() == b.back()) { r.pop_back(); b.pop_back(); }
            pairs.emplace_back(std::move(r), std::move(b));
        }
        const bool excluded = chrom.compare((size_t)pos, ref.size(), ref) != 0 || pos < int64_t(k) - 1;
        std::vector<size_t> included;
        if (!excluded)
            for (size_t i = 0; i < pairs.size(); i++) {
                const auto &r = pairs[i].first, &a = pairs[i].second;
                if (pos + (int64_t)r.size() - 1 + k > n || r.size() > o.max_allele_length || a.size() > o.max_allele_length) continue;
                dependencies.insert((uint32_t)(pos + r.size() - 1));
                included.push_back(i);
            }
        if (excluded || included.empty()) continue;
        if (pos - group_end >= int64_t(k)) flush();
        if (pos > prev_var_end + 1) addRegion(prev_var_end + 1, pos - 1);
        std::set<uint32_t> ends;
        for (size_t i : included) {
            const auto &r = pairs[i].first, &a = pairs[i].second;
            uint32_t li = 0;
            while (li < r.size() && li < a.size() && r[li] == a[li]) li++;
            var.n_red = std::min(var.n_red, li);
            var.alts.push_back({(uint32_t)r.size(), a, i < v.aco.size() ? v.aco[i] : std::string()});
            const uint32_t after = (uint32_t)(pos + r.size());
            const uint32_t cnv = std::max(copyNumberVariantLength(r, chrom, after, o), copyNumberVariantLength(a, chrom, after, o));
            ends.insert(after - 1);
            group_end = std::max<int64_t>(group_end, int64_t(after) - 1 + cnv);
        }
        prev_var_end = std::max<int64_t>(prev_var_end, *ends.rbegin());
        clusterVariant(std::move(var), (uint32_t)pos, std::vector<uint32_t>(ends.begin(), ends.end()), flanks, group, merge_sets, k);
    }
    flush();
    if (prev_var_end + 1 <= n - 1) addRegion(prev_var_end + 1, n - 1);
    return out;
}

// cluster -> the tightest cluster that contains it; registers contained clusters on their containers (position order)
inline std::unordered_map<uint32_t, uint32_t> groupDependencies(Group &group) {
    std::unordered_map<uint32_t, uint32_t> deps;
    for (auto &a : group) {
        Cluster *cl = a.second.get(), *container = nullptr;
        for (auto &b : group) {
            if (b.first == a.first) continue;
            Cluster *other = b.second.get();
            if (cl->left > other->left && cl->right < other->right) {
                if (!container || (other->left > container->left && other->right < container->right)) container = other;
            } else if (!(cl->left < other->left && cl->right > other->right)) {
                if (!(cl->right < other->left || other->right < cl->left)) throw std::runtime_error("clusters of a group overlap without one containing the other");
            }
        }
        if (container) deps.emplace(a.first, container->idx);
    }
    for (auto &d : deps) {
        Cluster *cl = group.at(d.first).get();
        group.at(d.second)->contained.push_back({cl->idx, cl->left, cl->right});
    }
    for (auto &a : group)
        std::sort(a.second->contained.begin(), a.second->contained.end(), [](const Contained &x, const Contained &y) { return x.left < y.left; });
    return deps;
}

// one cluster's graph appended to the flat arrays
struct GraphEmitter {
    Graphs &g;
    const std::string &chrom;
    uint32_t k;
    uint32_t base = 0;      // first vertex of the cluster in the flat arrays
    std::vector<std::vector<uint32_t>> in_src;
    std::vector<std::vector<uint8_t>> seq;
    std::vector<std::vector<uint16_t>> refvar;
    std::vector<uint8_t> flags;
    std::vector<uint16_t> var, allele;
    std::vector<uint32_t> nested;

    uint32_t addVertex() {
        in_src.emplace_back(); seq.emplace_back(); refvar.emplace_back(); flags.push_back(0); var.push_back(kNone16); allele.push_back(kNone16); nested.push_back(kNone32);
        return (uint32_t)flags.size() - 1;
    }
    void addEdge(uint32_t u, uint32_t v) { in_src[v].push_back(u); }
    uint32_t initVertex(uint32_t cur, const char *p, size_t len, std::pair<uint16_t, uint16_t> ai, const std::vector<uint16_t> &rv, uint32_t inner, bool redundant) {
        var[cur] = ai.first; allele[cur] = ai.second; refvar[cur] = rv; nested[cur] = inner;
        flags[cur] = uint8_t((redundant ? 1 : 0) | (inner != kNone32 ? 2 : 0));
        bool prev_disc = false;
        for (size_t i = 0; i < len; i++) {
            const int c = code(p[i]);
            if (c > 3) {
                if (!prev_disc) {
                    const uint32_t nxt = addVertex();
                    addEdge(cur, nxt);
                    var[nxt] = ai.first; allele[nxt] = ai.second; refvar[nxt] = rv; flags[nxt] = 2;
                    cur = nxt;
                }
                prev_disc = true;
            } else { seq[cur].push_back((uint8_t)c); prev_disc = false; }
        }
        return cur;
    }
    using Piece = std::pair<const char *, size_t>;
    uint32_t addVertices(uint32_t cur, const std::vector<Piece> &pieces, std::pair<uint16_t, uint16_t> ai, const std::set<uint16_t> &open_ref, const std::vector<uint32_t> &inner,
                         bool redundant) {
        std::vector<uint16_t> rv;
        for (uint16_t r : open_ref) if (r != ai.first) rv.push_back(r);
        cur = initVertex(cur, pieces[0].first, pieces[0].second, ai, rv, kNone32, redundant);
        for (size_t i = 1; i < pieces.size(); i++) {
            const uint32_t nxt = addVertex();
            addEdge(cur, nxt);
            cur = initVertex(nxt, pieces[i].first, pieces[i].second, ai, rv, inner[i - 1], false);
        }
        return cur;
    }
    void build(const Cluster &cl) {
        std::map<uint32_t, std::pair<std::vector<uint32_t>, std::vector<uint16_t>>> added;
        std::set<uint16_t> open_ref;
        size_t next_contained = 0;
        const uint32_t first = cl.variants.begin()->first;
        uint32_t cur = addVertex();
        cur = addVertices(cur, {{chrom.data() + first - (k - 1), k - 1}}, {kNone16, kNone16}, open_ref, {}, false);
        uint32_t prev_vertex = cur;
        added[first].first.push_back(cur);
        uint16_t vi = 0;
        for (auto it = cl.variants.begin(); it != cl.variants.end(); ++it, ++vi) {
            const uint32_t pos = it->first;
            const Variant &v = it->second;
            const bool redundant = v.n_red > 0;
            uint32_t max_ref = 0;
            for (size_t ai = 0; ai < v.alts.size(); ai++) {
                max_ref = std::max(max_ref, v.alts[ai].ref_len);
                uint32_t nxt = addVertex();
                addEdge(cur, nxt);
                nxt = addVertices(nxt, {{v.alts[ai].seq.data(), v.alts[ai].seq.size()}}, {vi, uint16_t(ai + 1)}, open_ref, {}, redundant);
                added[pos + v.alts[ai].ref_len].first.push_back(nxt);
            }
            added[pos + max_ref].second.push_back(vi);
            open_ref.insert(vi);
            auto nit = std::next(it);
            const bool last_variant = nit == cl.variants.end();
            const uint32_t next_pos = last_variant ? 0 : nit->first;
            bool more = true;
            while (more) {
                uint32_t cur_pos = added.begin()->first;
                const std::vector<uint32_t> next_vertices = added.begin()->second.first;
                for (uint16_t r : added.begin()->second.second) open_ref.erase(r);
                added.erase(added.begin());
                uint32_t cur_last;
                if (added.empty()) { more = false; cur_last = last_variant ? cur_pos + k - 1 : next_pos; }
                else {
                    cur_last = added.begin()->first;
                    if (!last_variant && cur_last > next_pos) { more = false; cur_last = next_pos; }
                }
                std::vector<Piece> pieces;
                std::vector<uint32_t> inner;
                while (next_contained < cl.contained.size() && cl.contained[next_contained].left < cur_last) {
                    const Contained &c = cl.contained[next_contained];
                    if (!(cur_pos <= c.left && c.right + k <= cur_last)) throw std::runtime_error("contained cluster does not fit inside one stretch of its container's graph");
                    pieces.emplace_back(chrom.data() + cur_pos, c.left - cur_pos);
                    inner.push_back(c.idx);
                    cur_pos = c.right + 1;
                    next_contained++;
                }
                pieces.emplace_back(chrom.data() + cur_pos, cur_last - cur_pos);
                cur = addVertex();
                bool is_ref = false;
                for (uint32_t u : next_vertices) { if (u == prev_vertex) is_ref = true; addEdge(u, cur); }
                cur = is_ref ? addVertices(cur, pieces, {vi, 0}, open_ref, inner, redundant) : addVertices(cur, pieces, {kNone16, kNone16}, open_ref, inner, false);
                added[cur_last].first.push_back(cur);
            }
            prev_vertex = cur;
        }
        if (next_contained != cl.contained.size()) throw std::runtime_error("contained cluster lies outside its container's graph");
        for (size_t v = 0; v < flags.size(); v++) {
            g.seq.insert(g.seq.end(), seq[v].begin(), seq[v].end()); g.v_seq_off.push_back(g.seq.size());
            g.v_flags.push_back(flags[v]); g.v_var.push_back(var[v]); g.v_allele.push_back(allele[v]); g.v_nested.push_back(nested[v]);
            g.v_refvar.insert(g.v_refvar.end(), refvar[v].begin(), refvar[v].end()); g.v_refvar_off.push_back(g.v_refvar.size());
            g.v_in_src.insert(g.v_in_src.end(), in_src[v].begin(), in_src[v].end()); g.v_in_off.push_back(g.v_in_src.size());
        }
        g.cl_vertex_off.push_back(g.v_flags.size());
    }
};

struct BuiltGroup {
    uint32_t contig;
    std::vector<Cluster *> clusters;       // the order VariantClusterGroup's vertices take
    std::vector<uint32_t> sources;
    std::vector<std::vector<uint32_t>> out_edges;
    uint32_t start, end, n_variants;
    std::string region;
};

}  // namespace detail

namespace detail {

struct GenomeParse {
    std::vector<std::vector<Group>> keep;      // owns the clusters
    std::vector<BuiltGroup> built;             // file order
    std::vector<int64_t> first_variant;        // per built group: smallest input index among its variants
    std::vector<Region> regions;
    std::vector<std::vector<uint8_t>> breaks;  // per candidate contig, per variant line
};

inline GenomeParse parseGenome(const std::vector<std::string> &contig_names, const std::vector<std::string> &contig_seqs, const std::vector<uint8_t> &contig_decoy,
                               const std::vector<std::pair<std::string, std::vector<Candidate>>> &candidates, const Options &o) {
    std::unordered_map<std::string, uint32_t> index;
    for (uint32_t i = 0; i < contig_names.size(); i++) index.emplace(contig_names[i], i);
    GenomeParse p;
    std::vector<bool> visited(contig_names.size(), false);
    for (auto &cv : candidates) {
        auto it = index.find(cv.first);
        if (it == index.end()) throw std::runtime_error("variants on contig \"" + cv.first + "\", which the genome does not hold");
        const uint32_t c = it->second;
        if (contig_decoy[c]) {      // dropped; the running group end is -1 on a contig without clusters
            p.breaks.emplace_back();
            for (auto &v : cv.second) p.breaks.back().push_back(v.pos + 1 >= o.k);
            continue;
        }
        visited[c] = true;
        ContigResult r = parseContig(contig_seqs[c], cv.second, o);
        p.breaks.push_back(std::move(r.breaks));
        for (auto &reg : r.regions) p.regions.push_back({c, false, reg.first, reg.second});
        for (auto &group : r.groups) {
            auto deps = groupDependencies(group);
            BuiltGroup b;
            b.contig = c;
            std::unordered_map<uint32_t, uint32_t> slot;
            for (auto &kv : group) { slot.emplace(kv.first, (uint32_t)b.clusters.size()); b.clusters.push_back(kv.second.get()); }
            for (uint32_t i = 0; i < b.clusters.size(); i++) if (!deps.count(b.clusters[i]->idx)) b.sources.push_back(i);
            b.out_edges.resize(b.clusters.size());
            for (auto &d : deps) b.out_edges[slot.at(d.second)].push_back(slot.at(d.first));
            b.start = kNone32; b.end = 0; b.n_variants = 0;
            int64_t first = std::numeric_limits<int64_t>::max();
            for (Cluster *cl : b.clusters) {
                b.start = std::min(b.start, cl->left + 1); b.end = std::max(b.end, cl->right + 1); b.n_variants += (uint32_t)cl->variants.size();
                for (auto &pv : cl->variants) first = std::min(first, pv.second.input_idx);
            }
            b.region = contig_names[c] + ":" + std::to_string(b.start) + "-" + std::to_string(b.end);
            p.built.push_back(std::move(b));
            p.first_variant.push_back(first);
        }
        p.keep.push_back(std::move(r.groups));
    }
    for (uint32_t c = 0; c < contig_names.size(); c++)
        if (!visited[c] && contig_seqs[c].size() >= o.k) p.regions.push_back({c, contig_decoy[c] != 0, 0, (uint32_t)contig_seqs[c].size() - 1});
    return p;
}

// groups `which` of the parse, in the unit's order, as flat arrays
inline Graphs emitUnit(const GenomeParse &p, std::vector<uint32_t> order, const std::vector<std::string> &contig_seqs, const Options &o) {
    const auto &built = p.built;
    Graphs g;
    // VariantClusterGroupCompare: number of variants desc, then region string desc
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (built[a].n_variants != built[b].n_variants) return built[a].n_variants > built[b].n_variants;
        return built[a].region > built[b].region;
    });
    for (uint32_t gi : order) {
        const BuiltGroup &b = built[gi];
        g.group_nvar.push_back(b.n_variants); g.group_start.push_back(b.start); g.group_end.push_back(b.end); g.group_contig.push_back(b.contig);
        g.group_src.insert(g.group_src.end(), b.sources.begin(), b.sources.end()); g.group_src_off.push_back(g.group_src.size());
        for (uint32_t u = 0; u < b.out_edges.size(); u++)
            for (uint32_t t : b.out_edges[u]) { g.group_edge_src.push_back(u); g.group_edge_dst.push_back(t); }
        g.group_edge_off.push_back(g.group_edge_src.size());
        for (Cluster *cl : b.clusters) {
            GraphEmitter e{g, contig_seqs[b.contig], o.k};
            e.build(*cl);
            g.cluster_idx.push_back(cl->idx);
            for (auto &pv : cl->variants) {
                const Variant &v = pv.second;
                g.var_pos.push_back(pv.first + 1); g.var_dep.push_back(v.dep); g.var_nalt.push_back((uint16_t)v.alts.size());
                g.var_input_idx.push_back(v.input_idx); g.var_contig.push_back(b.contig);
                g.var_ids += v.id; g.var_id_off.push_back(g.var_ids.size());
                for (auto &a : v.alts) {
                    g.alt_reflen.push_back(a.ref_len);
                    g.alt_seq += a.seq; g.alt_seq_off.push_back(g.alt_seq.size());
                    g.alt_aco += a.aco; g.alt_aco_off.push_back(g.alt_aco.size());
                }
                g.var_alt_off.push_back(g.alt_reflen.size());
            }
            g.cl_var_off.push_back(g.var_pos.size());
        }
        g.group_cluster_off.push_back(g.cluster_idx.size());
    }
    return g;
}

}  // namespace detail

// The unit of a genome: contigs in FASTA order (decoys included, flagged), candidates per contig in VCF order.
// Variants on decoy contigs are dropped; a variant on a contig the genome does not hold is an error (Chromosomes.cpp:145).
inline Graphs buildGenomeGraphs(const std::vector<std::string> &contig_names, const std::vector<std::string> &contig_seqs, const std::vector<uint8_t> &contig_decoy,
                                const std::vector<std::pair<std::string, std::vector<Candidate>>> &candidates, const Options &o = Options()) {
    detail::GenomeParse p = detail::parseGenome(contig_names, contig_seqs, contig_decoy, candidates, o);
    std::vector<uint32_t> all(p.built.size());
    for (uint32_t i = 0; i < all.size(); i++) all[i] = i;
    Graphs g = detail::emitUnit(p, all, contig_seqs, o);
    g.regions = p.regions;
    return g;
}

// The inference units `bayesTyper cluster` splits a candidate set into (main.cpp:219,233-247; VariantFileParser.cpp:286-290):
// floor(variants / min_unit_variants) units (at least one) of ceil(variants / units) variant lines each — excluded lines count —
// every unit running on to the next variant that lies at least k past the end of its group.  regions (of the whole genome) go to
// the first unit's `regions`.
inline std::vector<Graphs> buildGenomeUnits(const std::vector<std::string> &contig_names, const std::vector<std::string> &contig_seqs,
                                            const std::vector<uint8_t> &contig_decoy, const std::vector<std::pair<std::string, std::vector<Candidate>>> &candidates,
                                            uint32_t min_unit_variants, const Options &o = Options()) {
    detail::GenomeParse p = detail::parseGenome(contig_names, contig_seqs, contig_decoy, candidates, o);
    uint64_t total = 0;
    for (auto &cv : candidates) total += cv.second.size();
    const uint32_t n_units = std::max<uint32_t>(1, (uint32_t)std::floor((float)total / (float)min_unit_variants));
    const uint32_t per_unit = (uint32_t)std::ceil((float)total / (float)n_units);
    // unit of every variant line, file order
    std::vector<std::vector<uint32_t>> unit_of(candidates.size());
    uint32_t unit = 0, count = 0;
    for (size_t c = 0; c < candidates.size(); c++)
        for (size_t i = 0; i < candidates[c].second.size(); i++) {
            if (count >= per_unit && p.breaks[c][i]) { unit++; count = 0; }
            count++;
            unit_of[c].push_back(unit);
        }
    std::unordered_map<std::string, size_t> cand_index;
    for (size_t c = 0; c < candidates.size(); c++) cand_index.emplace(candidates[c].first, c);
    std::vector<std::vector<uint32_t>> members(unit + 1);
    for (uint32_t gi = 0; gi < p.built.size(); gi++)
        members[unit_of[cand_index.at(contig_names[p.built[gi].contig])][(size_t)p.first_variant[gi]]].push_back(gi);
    std::vector<Graphs> units;
    for (auto &m : members) {
        if (m.empty()) throw std::runtime_error("an inference unit holds no usable variant (only excluded or decoy lines): raise --min-unit-variants");
        units.push_back(detail::emitUnit(p, m, contig_seqs, o));
    }
    units.front().regions = p.regions;
    return units;
}

}  // namespace cluster
}  // namespace btg
