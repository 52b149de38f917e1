/* Stand-in for boost::dynamic_bitset<> (Boost 1.83 is un-vendored): bit storage only, which is all
 * Runtimes/Voxel/Occupancy/BinaryOccupancyVolume.h:9-39 uses (sized constructor, operator[]). Test infrastructure. */
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
template <typename Block = unsigned long>
class dynamic_bitset {
  std::vector<bool> bits_;
 public:
  dynamic_bitset() = default;
  explicit dynamic_bitset(std::size_t n) : bits_(n, false) {}
  std::vector<bool>::reference operator[](std::size_t i) { return bits_.at(i); }
  bool operator[](std::size_t i) const { return bits_.at(i); }
  std::size_t size() const { return bits_.size(); }
};
}  // namespace boost
