// Oracle build shim (test infrastructure only): boost::filesystem::{exists,path::stem} stand-ins.
#ifndef ORACLE_SHIM_BOOST_FILESYSTEM_HPP_
#define ORACLE_SHIM_BOOST_FILESYSTEM_HPP_
#include <fstream>
#include <string>
#include <sys/stat.h>
namespace boost {
namespace filesystem {
class path {
 public:
  path(const std::string& p) : p_(p) {}
  path(const char* p) : p_(p) {}
  path filename() const {
    size_t k = p_.find_last_of('/');
    return path(k == std::string::npos ? p_ : p_.substr(k + 1));
  }
  path stem() const {
    std::string f = filename().p_;
    if (f == "." || f == "..") return path(f);
    size_t k = f.find_last_of('.');
    if (k == std::string::npos || k == 0) return path(f);
    return path(f.substr(0, k));
  }
  std::string string() const { return p_; }
  const char* c_str() const { return p_.c_str(); }
 private:
  std::string p_;
};
inline bool exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
class ofstream : public std::ofstream {
 public:
  ofstream() {}
  explicit ofstream(const std::string& p) : std::ofstream(p.c_str(), std::ios_base::out | std::ios_base::binary) {}
};
class ifstream : public std::ifstream {
 public:
  ifstream() {}
  explicit ifstream(const std::string& p) : std::ifstream(p.c_str(), std::ios_base::in | std::ios_base::binary) {}
};
}  // namespace filesystem
}  // namespace boost
#endif
