// Stand-in for the subset of boost::adjacency_list<vecS, vecS, bidirectionalS, VP> the
// reference uses (src/bayesTyper/VariantClusterGraph.cpp).  Vertex descriptors are
// size_t indices; out/in edge lists keep insertion order, as boost's vecS storage does.
#pragma once
#include <cstddef>
#include <utility>
#include <vector>
namespace boost {
struct vecS {};
struct bidirectionalS {};
template <class OutS, class VS, class Dir, class VP> class adjacency_list {
   public:
    typedef size_t vertex_descriptor;
    struct edge_descriptor { size_t src, dst; };
    struct vertex_iterator {
        size_t i;
        size_t operator*() const { return i; }
        vertex_iterator &operator++() { ++i; return *this; }
        vertex_iterator operator++(int) { vertex_iterator t = *this; ++i; return t; }
        bool operator==(const vertex_iterator &o) const { return i == o.i; }
        bool operator!=(const vertex_iterator &o) const { return i != o.i; }
    };
    typedef typename std::vector<edge_descriptor>::const_iterator edge_iterator;
    VP &operator[](size_t v) { return props[v]; }
    const VP &operator[](size_t v) const { return props[v]; }
    std::vector<VP> props;
    std::vector<std::vector<edge_descriptor>> out, in;
};
template <class G> struct graph_traits { typedef typename G::vertex_descriptor vertex_descriptor; };
template <class O, class V, class D, class VP> size_t add_vertex(adjacency_list<O, V, D, VP> &g) {
    g.props.emplace_back(); g.out.emplace_back(); g.in.emplace_back();
    return g.props.size() - 1;
}
template <class O, class V, class D, class VP>
std::pair<typename adjacency_list<O, V, D, VP>::edge_descriptor, bool> add_edge(size_t u, size_t v, adjacency_list<O, V, D, VP> &g) {
    typename adjacency_list<O, V, D, VP>::edge_descriptor e{u, v};
    g.out[u].push_back(e); g.in[v].push_back(e);
    return std::make_pair(e, true);
}
template <class O, class V, class D, class VP>
std::pair<typename adjacency_list<O, V, D, VP>::vertex_iterator, typename adjacency_list<O, V, D, VP>::vertex_iterator> vertices(const adjacency_list<O, V, D, VP> &g) {
    typedef typename adjacency_list<O, V, D, VP>::vertex_iterator It;
    return std::make_pair(It{0}, It{g.props.size()});
}
template <class O, class V, class D, class VP>
std::pair<typename adjacency_list<O, V, D, VP>::edge_iterator, typename adjacency_list<O, V, D, VP>::edge_iterator> in_edges(size_t v, const adjacency_list<O, V, D, VP> &g) {
    return std::make_pair(g.in[v].begin(), g.in[v].end());
}
template <class O, class V, class D, class VP>
std::pair<typename adjacency_list<O, V, D, VP>::edge_iterator, typename adjacency_list<O, V, D, VP>::edge_iterator> out_edges(size_t v, const adjacency_list<O, V, D, VP> &g) {
    return std::make_pair(g.out[v].begin(), g.out[v].end());
}
template <class E, class G> size_t source(const E &e, const G &) { return e.src; }
template <class E, class G> size_t target(const E &e, const G &) { return e.dst; }
template <class O, class V, class D, class VP> size_t num_vertices(const adjacency_list<O, V, D, VP> &g) { return g.props.size(); }
}  // namespace boost
