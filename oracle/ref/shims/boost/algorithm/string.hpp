// Minimal stand-in for the two boost::algorithm helpers the reference calls (test infrastructure only).
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of_pred { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline is_any_of_pred is_any_of(const std::string &s) { is_any_of_pred p; p.set = s; return p; }
template <typename Seq, typename Pred>
Seq &split(Seq &out, const std::string &in, Pred pred) {
    out.clear();
    std::string cur;
    for (std::size_t i = 0; i < in.size(); i++) {
        if (pred(in[i])) { out.push_back(cur); cur.clear(); } else cur += in[i];
    }
    out.push_back(cur);
    return out;
}
inline bool ends_with(const std::string &s, const std::string &suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
namespace algorithm { using boost::ends_with; using boost::split; }
}  // namespace boost
