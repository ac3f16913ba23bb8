// Oracle build shim (test infrastructure only): boost::lexical_cast for the string <-> number uses of the reference.
#ifndef ORACLE_SHIM_BOOST_LEXICAL_CAST_HPP_
#define ORACLE_SHIM_BOOST_LEXICAL_CAST_HPP_
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : public std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <class T, class S>
inline T lexical_cast(const S& v) {
  std::stringstream ss;
  T out;
  if (!(ss << v) || !(ss >> out) || !(ss >> std::ws).eof()) throw bad_lexical_cast();
  return out;
}
}  // namespace boost
#endif
