// utility.h — host-side mirror of the reference's `Utility` base class and data types
// (reference include/utility.h:73-185 types, :187-327 parameter block, :346-392 polar helpers,
// :440-477 vector helpers, :488-505 Euler extraction).  Same names, same fields, same parameter keys
// and defaults, so code written against the reference's headers keeps compiling; every stage that
// computes anything is forwarded to libscvod_b200.so (include/scvod.h) by SSC / PatchWork.
#pragma once
#ifndef _UTILITY_H_
#define _UTILITY_H_

#include <ros/ros.h>

#include <Eigen/Dense>
#include <opencv2/opencv.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "tictoc.h"

namespace fs = std::filesystem;
using std::ios;
using std::ofstream;

// pose "point" of LIO-SAM style pose clouds (reference include/utility.h:73-93); SSC::tracking reads
// x, y, z, roll, pitch, yaw.
struct PointXYZIRPYT {
  PCL_ADD_POINT4D
  PCL_ADD_INTENSITY;
  float roll;
  float pitch;
  float yaw;
  double time;
} EIGEN_ALIGN16;
typedef PointXYZIRPYT Pose;

// reference include/utility.h:96-106
struct PointAPRI {
  float x, y, z;
  float range;
  float angle;
  float azimuth;
  float intensity = 0.f;
  int range_idx = -1;
  int sector_idx = -1;
  int azimuth_idx = -1;
  int voxel_idx = -1;
};

// reference include/utility.h:109-119
struct Voxel {
  int range_idx;
  int sector_idx;
  int azimuth_idx;
  int label = -1;
  pcl::PointXYZI center;
  std::vector<int> ptIdx;
  std::vector<float> intensity_record;
  float intensity_av = 0.f;
  float intensity_cov = 0.f;
};

struct FeatureValue {
  FeatureValue() = delete;
  FeatureValue(std::string value_name_, double value_) : name(value_name_), value(value_) {}
  std::string name = "";
  double value = 0.0;
};

struct Feature {
  Feature() = delete;
  Feature(std::string feature_name) : name(feature_name) {}
  std::string name = "";
  std::vector<FeatureValue> feature_values;
};

// reference include/utility.h:142-162
struct Cluster {
  Cluster() { allocateMemory(); }
  void allocateMemory() { cloud.reset(new pcl::PointCloud<pcl::PointXYZI>()); }
  int track_id = -1;
  int name = -1;
  int type = -1;   // building, tree, car
  int state = -1;  // dynamic 1, static 0
  int color[3] = {0, 0, 0};
  std::pair<pcl::PointXYZI, pcl::PointXYZI> bounding_box;
  std::vector<int> occupy_pts;
  std::vector<int> occupy_voxels;
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud;
  Eigen::MatrixXd feature_matrix;
};

// reference include/utility.h:165-185 (+ one field: the frame's index inside the CUDA context)
struct Frame {
  Frame() { allocateMemory(); }
  void allocateMemory() {
    vox_cloud.reset(new pcl::PointCloud<pcl::PointXYZI>());
    cloud_use.reset(new pcl::PointCloud<pcl::PointXYZI>());
  }
  int id = 0;
  int max_name = 0;
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_use;
  std::unordered_map<int, Voxel> hash_cloud;
  pcl::PointCloud<pcl::PointXYZI>::Ptr vox_cloud;
  std::unordered_map<int, Cluster> cluster_set;
  std::vector<std::vector<int>> static_pt;
  std::vector<std::vector<int>> dynamic_pt;
  int scvod_frame = -1;  // index of this frame in the scvod context (not in the reference)
};

// The parameter block: one row per ROS parameter the reference's Utility() reads (include/utility.h:260-327),
// (type, key on the parameter server, member name, default).  The rows both declare the public members and
// load them, so the two can never drift apart.
#define UFO_PARAMETER_TABLE(X)                                              \
  X(std::string, "common/out_path_", out_path, std::string(" "))            \
  X(int, "common/kNumOmpCores_", kNumOmpCores, 6)                           \
  X(bool, "common/save_", save, true)                                       \
  X(bool, "common/mapping_init_", mapping_init, false)                      \
  X(bool, "common/is_pcd_", is_pcd, false)                                  \
  X(int, "common/skip_", skip, 2)                                           \
  X(std::string, "session/data_path_", data_path, std::string(" "))         \
  X(std::string, "session/label_path_", label_path, std::string(" "))       \
  X(std::string, "session/pose_path_", pose_path, std::string(" "))         \
  X(int, "session/init_", init, 5)                                          \
  X(int, "session/start_", start, 5)                                        \
  X(int, "session/end_", end, 50)                                           \
  X(std::string, "ssc/calib_path_", calib_path, std::string(" "))           \
  X(std::string, "ssc/seg_path_", seg_path, std::string(" "))               \
  X(std::string, "ssc/pcd_path_", pcd_path, std::string(" "))               \
  X(std::string, "ssc/map_path_", map_path, std::string(" "))               \
  X(std::string, "ssc/evaluate_path_", evaluate_path, std::string(" "))     \
  X(float, "ssc/sensor_height_", sensor_height, 2.0f)                       \
  X(float, "ssc/min_dis_", min_dis, 0.0f)                                   \
  X(float, "ssc/max_dis_", max_dis, 50.0f)                                  \
  X(float, "ssc/min_angle_", min_angle, 0.0f)                               \
  X(float, "ssc/max_angle_", max_angle, 360.0f)                             \
  X(float, "ssc/min_azimuth_", min_azimuth, -30.0f)                         \
  X(float, "ssc/max_azimuth_", max_azimuth, 60.0f)                          \
  X(float, "ssc/range_res_", range_res, 0.2f)                               \
  X(float, "ssc/sector_res_", sector_res, 1.2f)                             \
  X(float, "ssc/azimuth_res_", azimuth_res, 2.0f)                           \
  X(float, "ssc/refine_height_", refine_height, -1.0f)                      \
  X(float, "ssc/max_z_", max_z, 1.0f)                                       \
  X(float, "ssc/min_z_", min_z, -1.0f)                                      \
  X(float, "ssc/car_angle_", car_angle, 120.0f)                             \
  X(float, "ssc/car_height_", car_height, 2.0f)                             \
  X(float, "ssc/car_square_", car_square, 2.0f)                             \
  X(float, "ssc/max_intensity_", max_intensity, 200.0f)                     \
  X(float, "ssc/correct_radius_", correct_radius, 0.5f)                     \
  X(float, "ssc/correct_ratio_", correct_ratio, 0.5f)                       \
  X(int, "ssc/search_num_", search_num, 10)                                 \
  X(int, "ssc/iteration_", iteration, 3)                                    \
  X(int, "ssc/toBeClass_", toBeClass, 1)                                    \
  X(int, "ssc/search_c_", search_c, 2)                                      \
  X(float, "ssc/intensity_diff_", intensity_diff, 50.0f)                    \
  X(float, "ssc/intensity_cov_", intensity_cov, 20.0f)                      \
  X(float, "ssc/occupancy_", occupancy, 0.6f)                               \
  X(int, "ssc/building_", building, 0)                                      \
  X(int, "ssc/tree_", tree, 1)                                              \
  X(int, "ssc/car_", car, 2)                                                \
  X(std::vector<int>, "ssc/dynamic_label_", dynamic_label, std::vector<int>()) \
  X(std::vector<float>, "ssc/tr_", tr_v, std::vector<float>())              \
  X(double, "feature/kOneThird_", kOneThird, 0.333)                         \
  X(double, "feature/kLinearityMax_", kLinearityMax, 740.0)                 \
  X(double, "feature/kPlanarityMax_", kPlanarityMax, 959.0)                 \
  X(double, "feature/kScatteringMax_", kScatteringMax, 1248.0)              \
  X(double, "feature/kOmnivarianceMax_", kOmnivarianceMax, 0.278636)        \
  X(double, "feature/kAnisotropyMax_", kAnisotropyMax, 1248.0)              \
  X(double, "feature/kEigenEntropyMax_", kEigenEntropyMax, 0.956129)        \
  X(double, "feature/kChangeOfCurvatureMax_", kChangeOfCurvatureMax, 0.99702) \
  X(double, "feature/kNPointsMax_", kNPointsMax, 13200.0)

class Utility {
 public:
#define UFO_DECLARE_MEMBER(T, key, member, fallback) T member;
  UFO_PARAMETER_TABLE(UFO_DECLARE_MEMBER)
#undef UFO_DECLARE_MEMBER
  Eigen::Matrix4f tr;  // ssc/tr_ as a row-major 4x4 (velodyne -> camera calibration of KITTI)
  ros::NodeHandle nh;

  virtual ~Utility() {}

  Utility() {
#define UFO_LOAD_MEMBER(T, key, member, fallback) nh.param<T>(key, member, fallback);
    UFO_PARAMETER_TABLE(UFO_LOAD_MEMBER)
#undef UFO_LOAD_MEMBER
    tr = Eigen::Matrix4f::Identity();
    if (tr_v.size() == 16)
      for (int i = 0; i < 16; ++i) tr.d[i] = tr_v[i];
  }

  void fsmkdir(std::string _path) {
    if (!fs::is_directory(_path) || !fs::exists(_path)) fs::create_directories(_path);
  }

  // Scalar helpers kept for source compatibility.  The product path never calls them (binning runs in
  // k_patch_* / k_bin_only with a device port of glibc's atan2f); they are not a CPU fallback.
  template <typename T>
  float rad2deg(const T& radians) {
    return (float)radians * 180.0 / M_PI;
  }
  template <typename T>
  float deg2rad(const T& degrees) {
    return (float)degrees * M_PI / 180.0;
  }
  template <typename PointT>
  float pointDistance2d(const PointT& p1) {
    return (float)sqrt((p1.x) * (p1.x) + (p1.y) * (p1.y));
  }
  template <typename PointT>
  float pointDistance3d(const PointT& p1) {
    return (float)sqrt((p1.x) * (p1.x) + (p1.y) * (p1.y) + (p1.z) * (p1.z));
  }

  template <typename CloudT>
  void getCloudByVec(const CloudT& cloud_, const std::vector<int>& vec_, CloudT& cloud_out_) {
    for (auto& it : vec_) cloud_out_->points.push_back(cloud_->points[it]);
  }
  template <typename T>
  void addVec(std::vector<T>& vec_central_, const std::vector<T>& vec_add_) {
    vec_central_.insert(vec_central_.end(), vec_add_.begin(), vec_add_.end());
  }
  template <typename T>
  void reduceVec(std::vector<T>& vec_central_, const std::vector<T>& vec_reduce_) {
    for (auto it = vec_reduce_.begin(); it != vec_reduce_.end(); it++)
      vec_central_.erase(std::remove(vec_central_.begin(), vec_central_.end(), *it), vec_central_.end());
  }
  template <typename T>
  void sampleVec(std::vector<T>& vec_central_) {
    std::sort(vec_central_.begin(), vec_central_.end());
    vec_central_.erase(std::unique(vec_central_.begin(), vec_central_.end()), vec_central_.end());
  }
  bool findNameInVec(const int& name_, const std::vector<int>& vec_) { return std::count(vec_.begin(), vec_.end(), name_) != 0; }

  Eigen::Vector3f rotationMatrixToEulerAngles(Eigen::Matrix3f& R) {
    float sy = sqrt(R(0, 0) * R(0, 0) + R(1, 0) * R(1, 0));
    bool singular = sy < 1e-6;
    float x, y, z;
    if (!singular) {
      x = atan2(R(2, 1), R(2, 2));
      y = atan2(-R(2, 0), sy);
      z = atan2(R(1, 0), R(0, 0));
    } else {
      x = atan2(-R(1, 2), R(1, 1));
      y = atan2(-R(2, 0), sy);
      z = 0;
    }
    return {x, y, z};
  }
};

#endif
