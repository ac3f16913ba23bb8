// Oracle build shim (test infrastructure only): a flat string map with the ptree calls the reference makes.
#ifndef ORACLE_SHIM_BOOST_PTREE_HPP_
#define ORACLE_SHIM_BOOST_PTREE_HPP_
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
namespace property_tree {
class shim_node {
 public:
  shim_node() {}
  explicit shim_node(const std::string& v) : v_(v) {}
  const std::string& data() const { return v_; }
 private:
  std::string v_;
};
class ptree {
 public:
  typedef std::map<std::string, shim_node> map_type;
  typedef map_type::const_iterator const_assoc_iterator;
  typedef map_type::const_iterator const_iterator;
  typedef map_type::iterator iterator;
  template <class T>
  void put(const std::string& k, const T& v) { std::ostringstream os; os << v; m_[k] = shim_node(os.str()); }
  void put(const std::string& k, const std::string& v) { m_[k] = shim_node(v); }
  void put(const std::string& k, const char* v) { m_[k] = shim_node(v ? v : ""); }
  void put(const std::string& k, std::nullptr_t) { m_[k] = shim_node(""); }
  void put(const std::string& k, long v) { std::ostringstream os; if (v) os << v; m_[k] = shim_node(os.str()); }
  template <class T>
  T get(const std::string& k) const {
    map_type::const_iterator it = m_.find(k);
    if (it == m_.end()) throw std::runtime_error("ptree: no such key " + k);
    return convert<T>(it->second.data());
  }
  template <class T>
  T get(const std::string& k, const T& def) const {
    map_type::const_iterator it = m_.find(k);
    if (it == m_.end()) return def;
    return convert<T>(it->second.data());
  }
  std::string get(const std::string& k, const char* def) const { return get<std::string>(k, std::string(def)); }
  const_assoc_iterator find(const std::string& k) const { return m_.find(k); }
  const_assoc_iterator not_found() const { return m_.end(); }
  const_iterator begin() const { return m_.begin(); }
  const_iterator end() const { return m_.end(); }
  iterator begin() { return m_.begin(); }
  iterator end() { return m_.end(); }
 private:
  template <class T>
  static T convert(const std::string& s) { std::istringstream is(s); T v = T(); is >> v; return v; }
  map_type m_;
};
template <>
inline std::string ptree::convert<std::string>(const std::string& s) { return s; }
}  // namespace property_tree
}  // namespace boost
#endif
