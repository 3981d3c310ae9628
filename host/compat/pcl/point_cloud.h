// pcl/point_cloud.h — stand-in for pcl::PointCloud<T>: the members the reference's SSC code touches
// (points, width, height, is_dense, Ptr, +=, push_back, size, clear, empty, iteration).
#pragma once
#include <cstdint>
#include <vector>

#include <boost/shared_ptr.hpp>
namespace pcl {
template <typename PointT>
class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef PointT PointType;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;

  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() {
    points.clear();
    width = height = 0;
  }
  void resize(size_t n) {
    points.resize(n);
    width = (uint32_t)n;
    height = 1;
  }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = (uint32_t)points.size();
    height = 1;
  }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
  PointCloud& operator+=(const PointCloud& rhs) {
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = (uint32_t)points.size();
    height = 1;
    is_dense = is_dense && rhs.is_dense;
    return *this;
  }
};
}  // namespace pcl
