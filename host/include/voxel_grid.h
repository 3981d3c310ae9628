// voxel_grid.h — the 0.08 m pcl::VoxelGrid<PointXYZI> downsample of the reference's KITTI loader (src/ssc.cpp:1108-1111),
// restated for the host layer (PCL is not available in this image).
//
// [recollection] of PCL 1.8 filters/impl/voxel_grid.hpp (applyFilter, no filter field, downsample_all_data_ = true,
// min_points_per_voxel_ = 0): bounding box with getMinMax3D, inverse_leaf_size = 1 / leaf in float, voxel coordinate
// static_cast<int>(floor(x * inverse_leaf_size) - (float)min_b), linear index ijk0 + ijk1 * div_b0 + ijk2 * div_b0 * div_b1,
// std::sort of (index, point) pairs on the index alone, one output point per run = float sums of x, y, z, intensity in
// that order divided by the count (CentroidPoint: AccumulatorXYZ + AccumulatorIntensity).  The order of equal keys
// after std::sort is whatever libstdc++'s introsort leaves; the same call on the same vector is used here.
// This is loader code, it runs on the host like the reference's own loader.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <vector>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace ufo {

struct CloudPointIndexIdx {  // pcl::VoxelGrid::cloud_point_index_idx
  unsigned int idx;
  unsigned int cloud_point_index;
  bool operator<(const CloudPointIndexIdx& p) const { return idx < p.idx; }
};

// returns false (and copies the input) when the leaf size is too small for the extent, as PCL does (voxel_grid.hpp:236-242)
inline bool voxelGridXYZI(const pcl::PointCloud<pcl::PointXYZI>& in, float leaf, pcl::PointCloud<pcl::PointXYZI>& out) {
  std::vector<pcl::PointXYZI> result;
  const size_t n = in.points.size();
  if (n == 0) {
    out.points.clear();
    return true;
  }
  const float inv = 1.0f / leaf;
  float min_p[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, max_p[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
  for (const auto& p : in.points) {  // the loader's clouds are dense (is_dense = true): no finiteness filter
    min_p[0] = std::min(min_p[0], p.x);
    min_p[1] = std::min(min_p[1], p.y);
    min_p[2] = std::min(min_p[2], p.z);
    max_p[0] = std::max(max_p[0], p.x);
    max_p[1] = std::max(max_p[1], p.y);
    max_p[2] = std::max(max_p[2], p.z);
  }
  const int64_t dx = (int64_t)((max_p[0] - min_p[0]) * inv) + 1, dy = (int64_t)((max_p[1] - min_p[1]) * inv) + 1,
                dz = (int64_t)((max_p[2] - min_p[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT_MAX) {
    out.points = in.points;
    return false;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int d = 0; d < 3; ++d) {
    min_b[d] = (int)std::floor(min_p[d] * inv);
    max_b[d] = (int)std::floor(max_p[d] * inv);
    div_b[d] = max_b[d] - min_b[d] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<CloudPointIndexIdx> index_vector;
  index_vector.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    const auto& p = in.points[i];
    const int ijk0 = (int)(std::floor(p.x * inv) - (float)min_b[0]);
    const int ijk1 = (int)(std::floor(p.y * inv) - (float)min_b[1]);
    const int ijk2 = (int)(std::floor(p.z * inv) - (float)min_b[2]);
    CloudPointIndexIdx e;
    e.idx = (unsigned int)(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]);
    e.cloud_point_index = (unsigned int)i;
    index_vector.push_back(e);
  }
  std::sort(index_vector.begin(), index_vector.end(), std::less<CloudPointIndexIdx>());
  size_t index = 0;
  while (index < index_vector.size()) {
    size_t i = index + 1;
    while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx) ++i;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (size_t li = index; li < i; ++li) {
      const auto& p = in.points[index_vector[li].cloud_point_index];
      sx += p.x;
      sy += p.y;
      sz += p.z;
      si += p.intensity;
    }
    const float cnt = (float)(i - index);
    pcl::PointXYZI c;
    c.x = sx / cnt;
    c.y = sy / cnt;
    c.z = sz / cnt;
    c.intensity = si / cnt;
    result.push_back(c);
    index = i;
  }
  out.points.swap(result);
  return true;
}

}  // namespace ufo
