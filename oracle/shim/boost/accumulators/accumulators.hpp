// Oracle build shim (test infrastructure only): the count / mean / second-moment accumulator the reference uses once
// (src/bigraph_visitors.cpp:622-628).
#ifndef ORACLE_SHIM_BOOST_ACCUMULATORS_HPP_
#define ORACLE_SHIM_BOOST_ACCUMULATORS_HPP_
#include <cstddef>
namespace boost {
namespace accumulators {
namespace tag {
struct count {};
struct mean {};
template <int N> struct moment {};
}  // namespace tag
template <class A = void, class B = void, class C = void> struct stats {};
template <class T, class S>
struct accumulator_set {
  size_t n;
  T sum, sum2;
  accumulator_set() : n(0), sum(0), sum2(0) {}
  void operator()(T v) { ++n; sum += v; sum2 += v * v; }
};
template <class T, class S> inline size_t count(const accumulator_set<T, S>& a) { return a.n; }
template <class T, class S> inline T mean(const accumulator_set<T, S>& a) { return a.n ? a.sum / (T)a.n : (T)0; }
template <int N, class T, class S> inline T moment(const accumulator_set<T, S>& a) {
  static_assert(N == 2, "only the second moment is shimmed");
  return a.n ? a.sum2 / (T)a.n : (T)0;
}
}  // namespace accumulators
}  // namespace boost
#endif
