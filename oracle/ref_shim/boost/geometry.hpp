/* Stand-in for the slice of Boost.Geometry that Runtimes/Voxel/Spatial/NearestMap.h uses: a 3-D float point and an
 * "rtree" answering nearest(point, 1).  The tree is a linear scan over squared Euclidean distance -- the same answer
 * an R-tree gives for a 1-nearest query (first inserted wins an exact tie). Test infrastructure. */
#pragma once
#include <cstddef>
#include <utility>
#include <vector>
namespace boost { namespace geometry {
namespace cs { struct cartesian {}; }
namespace model {
template <typename T, std::size_t N, typename CS> class point {
  T v_[N];
 public:
  point() : v_{} {}
  point(T a, T b, T c) : v_{a, b, c} {}
  template <std::size_t K> T get() const { return v_[K]; }
};
}  // namespace model
namespace index {
template <std::size_t N> struct quadratic {};
template <typename P> struct nearest_predicate { P p; unsigned k; };
template <typename P> nearest_predicate<P> nearest(const P& p, unsigned k) { return {p, k}; }
template <typename Value, typename Params> class rtree {
  std::vector<Value> v_;
 public:
  void insert(const Value& v) { v_.push_back(v); }
  template <typename P, typename Out> std::size_t query(const nearest_predicate<P>& q, Out out) const {
    if (v_.empty() || q.k == 0) return 0;
    std::size_t best = 0; double bd = 0;
    for (std::size_t i = 0; i < v_.size(); ++i) {
      const auto& p = v_[i].first;
      double dx = (double)p.template get<0>() - (double)q.p.template get<0>();
      double dy = (double)p.template get<1>() - (double)q.p.template get<1>();
      double dz = (double)p.template get<2>() - (double)q.p.template get<2>();
      double d = dx * dx + dy * dy + dz * dz;
      if (i == 0 || d < bd) { bd = d; best = i; }
    }
    *out++ = v_[best];
    return 1;
  }
};
}  // namespace index
}}  // namespace boost::geometry
