// ORACLE SHIM (test infrastructure): the container surface of pcl::PointCloud<PointT> (PCL 1.10) the reference uses:
// points / header, push_back, at, operator[], operator+=, size, reserve, clear, iterators, Ptr = boost::shared_ptr.
#pragma once
#include <boost/shared_ptr.hpp>
#include <cstdint>
#include <string>
#include <vector>
namespace pcl {
struct PCLHeader { std::uint32_t seq = 0; std::uint64_t stamp = 0; std::string frame_id; };
template <typename PointT>
class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef typename std::vector<PointT>::iterator iterator;
  typedef typename std::vector<PointT>::const_iterator const_iterator;
  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  PointCloud() {}
  PointCloud& operator+=(const PointCloud& rhs) {
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = static_cast<std::uint32_t>(points.size()); height = 1;
    if (rhs.is_dense && is_dense) is_dense = true; else is_dense = false;
    return *this;
  }
  void push_back(const PointT& p) { points.push_back(p); width = static_cast<std::uint32_t>(points.size()); height = 1; }
  const PointT& at(std::size_t n) const { return points.at(n); }
  PointT& at(std::size_t n) { return points.at(n); }
  const PointT& operator[](std::size_t n) const { return points[n]; }
  PointT& operator[](std::size_t n) { return points[n]; }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void reserve(std::size_t n) { points.reserve(n); }
  void resize(std::size_t n) { points.resize(n); }
  void clear() { points.clear(); width = 0; height = 0; }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
};
}  // namespace pcl
