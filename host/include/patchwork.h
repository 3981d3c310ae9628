// patchwork.h — PatchWork<PointT> with the public interface of the reference's header-only class
// (include/patchwork.h:44 ctor, :105-111 estimate_ground / set_sensor), held by SSC as
// boost::shared_ptr<PatchWork<pcl::PointXYZI>> (include/ssc.h:32).  estimate_ground forwards to
// scvod_ground(): concentric-zone binning, per-patch z-sort, three sequential-order plane fits and the
// uprightness / elevation / flatness gates all run in the CUDA library (k_patch_* kernels); the outputs come
// back in the reference's order (patch-major, [ground...][nonground...] per rejected patch).
#pragma once
#ifndef PATCHWORK_H
#define PATCHWORK_H

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "scvod.h"
#include "tictoc.h"

template <typename PointT>
class PatchWork {
 public:
  typedef std::vector<pcl::PointCloud<PointT>> Ring;
  typedef std::vector<Ring> Zone;

  PatchWork() {}
  ~PatchWork() {
    if (ctx_) scvod_destroy(ctx_);
  }
  PatchWork(const PatchWork&) = delete;
  PatchWork& operator=(const PatchWork&) = delete;

  void set_sensor(const double& height) { sensor_height_ = height; }

  void estimate_ground(const pcl::PointCloud<PointT>& cloudIn, pcl::PointCloud<PointT>& cloudOut, pcl::PointCloud<PointT>& cloudNonground,
                       double& time_taken) {
    const int n = (int)cloudIn.points.size();
    ensure_context(n);
    stage_.resize((size_t)4 * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
      const PointT& p = cloudIn.points[i];
      stage_[4 * i] = p.x;
      stage_[4 * i + 1] = p.y;
      stage_[4 * i + 2] = p.z;
      stage_[4 * i + 3] = p.intensity;
    }
    gidx_.resize(n > 0 ? n : 1);
    nidx_.resize(n > 0 ? n : 1);
    int32_t ng = 0, nn = 0;
    TicToc t;
    if (scvod_ground(ctx_, stage_.data(), n, gidx_.data(), &ng, nidx_.data(), &nn) != SCVOD_OK)
      throw std::runtime_error(std::string("scvod_ground: ") + scvod_last_error());
    time_taken = t.toc() * 1e-3;  // the reference reports seconds
    cloudOut.clear();
    cloudNonground.clear();
    cloudOut.points.reserve(ng);
    cloudNonground.points.reserve(nn);
    for (int i = 0; i < ng; ++i) cloudOut.points.push_back(cloudIn.points[gidx_[i]]);
    for (int i = 0; i < nn; ++i) cloudNonground.points.push_back(cloudIn.points[nidx_[i]]);
    cloudOut.width = ng;
    cloudOut.height = 1;
    cloudNonground.width = nn;
    cloudNonground.height = 1;
  }

 private:
  void ensure_context(int n) {
    if (ctx_ && (float)sensor_height_ == ctx_height_ && n <= ctx_points_) return;
    if (ctx_) scvod_destroy(ctx_);
    ctx_ = nullptr;
    scvod_params p;
    scvod_params_semantickitti(&p);  // only sensor_height matters to the ground stage
    p.sensor_height = (float)sensor_height_;
    int cap = 131072;
    while (cap < n) cap *= 2;
    if (scvod_create(&p, 0, cap, 1, &ctx_) != SCVOD_OK) throw std::runtime_error(std::string("scvod_create: ") + scvod_last_error());
    ctx_height_ = (float)sensor_height_;
    ctx_points_ = cap;
  }

  double sensor_height_ = 1.732;  // include/patchwork.h:123 leaves it unset until set_sensor(); its comment says 1.732
  scvod_ctx* ctx_ = nullptr;
  float ctx_height_ = 0.f;
  int ctx_points_ = 0;
  std::vector<float> stage_;
  std::vector<int32_t> gidx_, nidx_;
};

#endif
