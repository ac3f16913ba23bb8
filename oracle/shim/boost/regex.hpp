// Oracle build shim (test infrastructure only): boost::regex mapped onto std::regex (ECMAScript covers the
// reference's only pattern, "^(\\d*)=$").
#ifndef ORACLE_SHIM_BOOST_REGEX_HPP_
#define ORACLE_SHIM_BOOST_REGEX_HPP_
#include <regex>
#include <string>
namespace boost {
typedef std::regex regex;
typedef std::smatch smatch;
inline bool regex_match(const std::string& s, const regex& re) { return std::regex_match(s, re); }
inline bool regex_match(const std::string& s, smatch& m, const regex& re) { return std::regex_match(s, m, re); }
}  // namespace boost
#endif
