// pcl/point_types.h — stand-in for the two PCL point types the SSC class surface uses.
// Same field names and 16-byte-aligned layouts as PCL 1.8 (PointXYZI is 32 bytes, PointXYZRGB is 32 bytes).
#pragma once
#include <cstdint>
namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, data_pad = 1.f;
};
struct alignas(16) PointXYZI {
  float x = 0.f, y = 0.f, z = 0.f, data_pad = 1.f;
  float intensity = 0.f;
  float data_c_pad[3] = {0.f, 0.f, 0.f};
};
struct alignas(16) PointXYZRGB {
  float x = 0.f, y = 0.f, z = 0.f, data_pad = 1.f;
  uint8_t b = 0, g = 0, r = 0, a = 255;
  float data_c_pad[3] = {0.f, 0.f, 0.f};
};
}  // namespace pcl
// point-struct registration macros of PCL used by reference include/utility.h:73-91 (Pose)
#define PCL_ADD_POINT4D float x, y, z, data_pad;
#define PCL_ADD_INTENSITY float intensity
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 __attribute__((aligned(16)))
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)
