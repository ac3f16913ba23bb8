// Oracle build shim (test infrastructure only): key=value INI reader.
#ifndef ORACLE_SHIM_BOOST_INI_PARSER_HPP_
#define ORACLE_SHIM_BOOST_INI_PARSER_HPP_
#include <fstream>
#include <stdexcept>
#include <boost/property_tree/ptree.hpp>
namespace boost {
namespace property_tree {
class ini_parser_error : public std::runtime_error {
 public:
  explicit ini_parser_error(const std::string& w) : std::runtime_error(w) {}
};
inline void read_ini(const std::string& file, ptree& pt) {
  std::ifstream in(file.c_str());
  if (!in) throw ini_parser_error("cannot open " + file);
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#' || line[0] == ';' || line[0] == '[') continue;
    size_t k = line.find('=');
    if (k == std::string::npos) continue;
    pt.put(line.substr(0, k), line.substr(k + 1));
  }
}
}  // namespace property_tree
}  // namespace boost
#endif
