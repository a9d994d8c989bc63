// Stand-in for boost::split / boost::is_any_of (token_compress_off: empty tokens kept).
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of_pred { std::string chars; bool operator()(char c) const { return chars.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(const std::string &chars) { return is_any_of_pred{chars}; }
template <class Seq> Seq &split(Seq &out, const std::string &in, const is_any_of_pred &p) {
    out.clear();
    std::string cur;
    for (char c : in) {
        if (p(c)) { out.push_back(cur); cur.clear(); } else cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}
}  // namespace boost
