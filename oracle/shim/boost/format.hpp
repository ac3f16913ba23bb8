// Oracle build shim (test infrastructure only): a tiny printf-style stand-in for boost::format.
// Supports the directives the reference's index/overlap translation units use: %s %d %lu %c %f (+ width/precision).
#ifndef ORACLE_SHIM_BOOST_FORMAT_HPP_
#define ORACLE_SHIM_BOOST_FORMAT_HPP_
#include <functional>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
 public:
  explicit format(const std::string& f) : fmt_(f) {}
  explicit format(const char* f) : fmt_(f) {}
  template <class T>
  format& operator%(const T& v) {
    args_.push_back(Arg());
    Arg& a = args_.back();
    std::ostringstream plain;
    plain << v;
    a.plain = plain.str();
    std::ostringstream fixed;
    fixed << std::fixed;
    a.render = [v](std::ostringstream& os) { os << v; };
    return *this;
  }
  format& operator%(std::nullptr_t) {
    args_.push_back(Arg());
    args_.back().plain = "";
    args_.back().render = [](std::ostringstream&) {};
    return *this;
  }
  std::string str() const {
    std::string out;
    size_t ai = 0;
    for (size_t i = 0; i < fmt_.size(); ++i) {
      char c = fmt_[i];
      if (c != '%') { out += c; continue; }
      if (i + 1 < fmt_.size() && fmt_[i + 1] == '%') { out += '%'; ++i; continue; }
      // parse %[flags][width][.prec][length]conv
      size_t j = i + 1;
      while (j < fmt_.size() && (fmt_[j] == '-' || fmt_[j] == '+' || fmt_[j] == ' ' || fmt_[j] == '0' || fmt_[j] == '#')) ++j;
      int width = 0;
      while (j < fmt_.size() && isdigit((unsigned char)fmt_[j])) { width = width * 10 + (fmt_[j] - '0'); ++j; }
      int prec = -1;
      if (j < fmt_.size() && fmt_[j] == '.') {
        ++j; prec = 0;
        while (j < fmt_.size() && isdigit((unsigned char)fmt_[j])) { prec = prec * 10 + (fmt_[j] - '0'); ++j; }
      }
      while (j < fmt_.size() && (fmt_[j] == 'l' || fmt_[j] == 'h' || fmt_[j] == 'z')) ++j;
      char conv = j < fmt_.size() ? fmt_[j] : 's';
      std::ostringstream os;
      if (width > 0) os << std::setw(width);
      if (conv == 'f' || conv == 'e' || conv == 'g') {
        os << std::fixed;
        os << std::setprecision(prec >= 0 ? prec : 6);
      }
      if (ai < args_.size()) {
        if ((conv == 's') && prec >= 0) {
          os << args_[ai].plain.substr(0, prec);
        } else {
          args_[ai].render(os);
        }
        ++ai;
      }
      out += os.str();
      i = j;
    }
    return out;
  }
 private:
  struct Arg {
    std::string plain;
    std::function<void(std::ostringstream&)> render;
  };
  std::string fmt_;
  std::vector<Arg> args_;
};
inline std::string str(const format& f) { return f.str(); }
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}  // namespace boost
#endif
