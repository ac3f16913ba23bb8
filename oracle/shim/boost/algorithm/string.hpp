// Oracle build shim (test infrastructure only): the handful of boost::algorithm string helpers the reference calls.
#ifndef ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP_
#define ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP_
#include <cctype>
#include <string>
#include <vector>
namespace boost {
struct shim_char_range { char lo, hi; bool operator()(char c) const { return c >= lo && c <= hi; } };
inline shim_char_range is_from_range(char lo, char hi) { shim_char_range r = {lo, hi}; return r; }
struct shim_any_of { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline shim_any_of is_any_of(const std::string& set) { shim_any_of r = {set}; return r; }
namespace algorithm {
template <class Seq>
inline std::string join(const Seq& parts, const std::string& sep) {
  std::string out;
  bool first = true;
  for (typename Seq::const_iterator it = parts.begin(); it != parts.end(); ++it) {
    if (!first) out += sep;
    out += *it;
    first = false;
  }
  return out;
}
inline bool equals(const std::string& a, const std::string& b) { return a == b; }
inline bool iequals(const std::string& a, const std::string& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (tolower((unsigned char)a[i]) != tolower((unsigned char)b[i])) return false;
  return true;
}
inline bool starts_with(const std::string& s, const std::string& p) {
  return s.size() >= p.size() && s.compare(0, p.size(), p) == 0;
}
inline bool ends_with(const std::string& s, const std::string& p) {
  return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0;
}
inline void trim(std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) ++b;
  while (e > b && isspace((unsigned char)s[e - 1])) --e;
  s = s.substr(b, e - b);
}
template <class Pred>
inline void split(std::vector<std::string>& out, const std::string& text, Pred pred) {
  out.clear();
  std::string cur;
  for (size_t i = 0; i < text.size(); ++i) {
    if (pred(text[i])) { out.push_back(cur); cur.clear(); } else { cur += text[i]; }
  }
  out.push_back(cur);
}
}  // namespace algorithm
}  // namespace boost
#endif
