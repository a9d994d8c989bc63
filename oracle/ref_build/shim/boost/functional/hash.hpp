// Minimal stand-in for boost/functional/hash.hpp (test infrastructure, see ../README).
// Only boost::hash<std::pair<A,B>> is used by the reference (lookup-only maps, or
// maps whose iteration order does not influence results — SURVEY.md §8c).
#pragma once
#include <cstddef>
#include <functional>
#include <utility>
namespace boost {
template <class T> struct hash { size_t operator()(const T &v) const { return std::hash<T>()(v); } };
template <class A, class B> struct hash<std::pair<A, B>> {
    size_t operator()(const std::pair<A, B> &p) const {
        size_t seed = 0;
        seed ^= std::hash<A>()(p.first) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        seed ^= std::hash<B>()(p.second) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};
}  // namespace boost
