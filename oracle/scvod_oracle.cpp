// scvod_oracle.cpp — CPU restatement of the reference's SCV-OD hot path.
//
// *** TEST INFRASTRUCTURE ONLY ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may load this library.  The product (libscvod_b200.so)
// never links, includes or calls anything in oracle/.
//
// *** PARITY UNPINNED ***  The reference (Yixin-F/DR-Using-SCV-OD @ 8c5538c) ships no tests, golden
// vectors or fixtures for this path, and it cannot be compiled in this image (needs ROS melodic,
// PCL 1.8, Eigen 3.3.4, OpenCV, boost, yaml-cpp: none installed, no network).  So this file restates
// (a) the reference's own sources line by line and (b) the published algorithms of the three
// third-party routines it calls on the path — pcl::computeMeanAndCovarianceMatrix (PCL 1.8
// common/impl/centroid.hpp), Eigen::JacobiSVD on a 3x3 float matrix (Eigen 3.3.4 SVD/JacobiSVD.h,
// misc/RealSvd2x2.h, Jacobi/Jacobi.h) and pcl::getTransformation / Affine3f::inverse — from memory of
// those versions.  "Bit-exact vs the reference" therefore means "bit-exact vs this restatement built
// with g++ -O3 -ffp-contract=off on glibc 2.39".
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Build: see oracle/Makefile (flags mirror the reference's build/compile_commands.json:
// -O3 -DNDEBUG -fopenmp -std=gnu++17, no -march, plus -ffp-contract=off to pin x86-64 baseline
// behaviour = no FMA).

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../include/scvod.h"  // only for the plain-C scvod_params / enum definitions

namespace {

struct Pt {
  float x, y, z, intensity;
  int src;  // index of this point in the scan handed to process(); -1 for carried (tracked) points
};
typedef std::vector<Pt> Cloud;

// ---------------------------------------------------------------------------------------------
// Third-party restatements
// ---------------------------------------------------------------------------------------------

// PCL 1.8 pcl::computeMeanAndCovarianceMatrix(cloud, Matrix3f&, Vector4f&), dense branch
// (common/impl/centroid.hpp).  Call site: include/patchwork.h:218.  Returns the point count; on an
// empty cloud the outputs are left untouched, as in PCL.
static unsigned mean_and_cov_pcl18(const Cloud& c, float cov[3][3], float mean[4]) {
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned n = (unsigned)c.size();
  for (unsigned i = 0; i < n; ++i) {
    const Pt& p = c[i];
    accu[0] += p.x * p.x;
    accu[1] += p.x * p.y;
    accu[2] += p.x * p.z;
    accu[3] += p.y * p.y;
    accu[4] += p.y * p.z;
    accu[5] += p.z * p.z;
    accu[6] += p.x;
    accu[7] += p.y;
    accu[8] += p.z;
  }
  if (n == 0) return 0;
  float fn = (float)n;
  for (int k = 0; k < 9; ++k) accu[k] /= fn;
  mean[0] = accu[6];
  mean[1] = accu[7];
  mean[2] = accu[8];
  mean[3] = 1.f;
  cov[0][0] = accu[0] - accu[6] * accu[6];
  cov[0][1] = accu[1] - accu[6] * accu[7];
  cov[0][2] = accu[2] - accu[6] * accu[8];
  cov[1][1] = accu[3] - accu[7] * accu[7];
  cov[1][2] = accu[4] - accu[7] * accu[8];
  cov[2][2] = accu[5] - accu[8] * accu[8];
  cov[1][0] = cov[0][1];
  cov[2][0] = cov[0][2];
  cov[2][1] = cov[1][2];
  return n;
}

struct Rot {
  float c, s;
};

// Eigen 3.3.4 JacobiRotation<float>::makeJacobi(x, y, z)  (Jacobi/Jacobi.h)
static Rot make_jacobi(float x, float y, float z) {
  Rot r;
  float deno = 2.f * std::fabs(y);
  if (deno < FLT_MIN) {
    r.c = 1.f;
    r.s = 0.f;
  } else {
    float tau = (x - z) / deno;
    float w = std::sqrt(tau * tau + 1.f);
    float t;
    if (tau > 0.f)
      t = 1.f / (tau + w);
    else
      t = 1.f / (tau - w);
    float sign_t = t > 0.f ? 1.f : -1.f;
    float n = 1.f / std::sqrt(t * t + 1.f);
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
  }
  return r;
}

// Eigen 3.3.4 internal::real_2x2_jacobi_svd (misc/RealSvd2x2.h)
static void real_2x2_jacobi_svd(const float W[3][3], int p, int q, Rot* j_left, Rot* j_right) {
  float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
  Rot rot1;
  float t = m00 + m11;
  float d = m10 - m01;
  if (std::fabs(d) < FLT_MIN) {
    rot1.s = 0.f;
    rot1.c = 1.f;
  } else {
    float u = t / d;
    float tmp = std::sqrt(1.f + u * u);
    rot1.s = 1.f / tmp;
    rot1.c = u / tmp;
  }
  // m.applyOnTheLeft(0,1,rot1): x=row0, y=row1 -> x' = c x + s y ; y' = -s x + c y
  if (!(rot1.c == 1.f && rot1.s == 0.f)) {
    float a00 = rot1.c * m00 + rot1.s * m10, a01 = rot1.c * m01 + rot1.s * m11;
    float a10 = -rot1.s * m00 + rot1.c * m10, a11 = -rot1.s * m01 + rot1.c * m11;
    m00 = a00;
    m01 = a01;
    m10 = a10;
    m11 = a11;
  }
  *j_right = make_jacobi(m00, m01, m11);
  // *j_left = rot1 * j_right->transpose();  transpose = (c, -s);  (a*b): c = ac*bc - as*bs ; s = ac*bs + as*bc
  Rot jt = {j_right->c, -j_right->s};
  j_left->c = rot1.c * jt.c - rot1.s * jt.s;
  j_left->s = rot1.c * jt.s + rot1.s * jt.c;
}

// Eigen 3.3.4 JacobiSVD<MatrixXf>(cov, ComputeFullU) on a square 3x3 input (SVD/JacobiSVD.h).
// Call site: include/patchwork.h:220-224.  U is column-major-agnostic here: U[r][c].
static void jacobi_svd_3x3(const float A[3][3], float U[3][3], float sv[3]) {
  const float precision = 2.f * FLT_EPSILON;
  const float considerAsZero = FLT_MIN;
  float scale = 0.f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(A[i][j]));
  if (scale == 0.f) scale = 1.f;
  float W[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      W[i][j] = A[i][j] / scale;
      U[i][j] = (i == j) ? 1.f : 0.f;
    }
  float maxDiag = std::max(std::fabs(W[0][0]), std::max(std::fabs(W[1][1]), std::fabs(W[2][2])));
  bool finished = false;
  while (!finished) {
    finished = true;
    for (int p = 1; p < 3; ++p) {
      for (int q = 0; q < p; ++q) {
        float threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::fabs(W[p][q]) > threshold || std::fabs(W[q][p]) > threshold) {
          finished = false;
          Rot jl, jr;
          real_2x2_jacobi_svd(W, p, q, &jl, &jr);
          // W.applyOnTheLeft(p,q,jl): rows p,q
          if (!(jl.c == 1.f && jl.s == 0.f)) {
            for (int k = 0; k < 3; ++k) {
              float xi = W[p][k], yi = W[q][k];
              W[p][k] = jl.c * xi + jl.s * yi;
              W[q][k] = -jl.s * xi + jl.c * yi;
            }
            // U.applyOnTheRight(p,q,jl.transpose()) == apply_rotation_in_the_plane(col p, col q, jl)
            for (int k = 0; k < 3; ++k) {
              float xi = U[k][p], yi = U[k][q];
              U[k][p] = jl.c * xi + jl.s * yi;
              U[k][q] = -jl.s * xi + jl.c * yi;
            }
          }
          // W.applyOnTheRight(p,q,jr) == apply_rotation_in_the_plane(col p, col q, jr.transpose())
          if (!(jr.c == 1.f && -jr.s == 0.f)) {
            float c = jr.c, s = -jr.s;
            for (int k = 0; k < 3; ++k) {
              float xi = W[k][p], yi = W[k][q];
              W[k][p] = c * xi + s * yi;
              W[k][q] = -s * xi + c * yi;
            }
          }
          maxDiag = std::max(maxDiag, std::max(std::fabs(W[p][p]), std::fabs(W[q][q])));
        }
      }
    }
  }
  for (int i = 0; i < 3; ++i) {
    float a = W[i][i];
    sv[i] = std::fabs(a);
    if (a < 0.f)
      for (int k = 0; k < 3; ++k) U[k][i] = -U[k][i];
  }
  for (int i = 0; i < 3; ++i) sv[i] *= scale;
  for (int i = 0; i < 3; ++i) {
    int pos = 0;
    float best = sv[i];
    for (int k = i + 1; k < 3; ++k)
      if (sv[k] > best) {
        best = sv[k];
        pos = k - i;
      }
    if (best == 0.f) break;
    if (pos) {
      pos += i;
      std::swap(sv[i], sv[pos]);
      for (int k = 0; k < 3; ++k) std::swap(U[k][i], U[k][pos]);
    }
  }
}

// pcl::getTransformation(x,y,z,roll,pitch,yaw) in float (PCL 1.8 common/impl/eigen.hpp) as a 3x4.
// Call sites: src/ssc.cpp:1163,1171,1255-1256.
static void pcl_get_transformation(const float p[6], float T[3][4]) {
  float x = p[0], y = p[1], z = p[2], roll = p[3], pitch = p[4], yaw = p[5];
  float A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch), E = std::cos(roll),
        F = std::sin(roll), DE = D * E, DF = D * F;
  T[0][0] = A * C;
  T[0][1] = A * DF - B * E;
  T[0][2] = B * F + A * DE;
  T[0][3] = x;
  T[1][0] = B * C;
  T[1][1] = A * E + B * DF;
  T[1][2] = B * DE - A * F;
  T[1][3] = y;
  T[2][0] = -D;
  T[2][1] = C * F;
  T[2][2] = C * E;
  T[2][3] = z;
}

// Eigen 3.3.4 Transform<float,3,Affine>::inverse(): linear part by the 3x3 cofactor inverse
// (LU/InverseImpl.h compute_inverse<Matrix3f,3>), translation = -(Rinv * t).
static void affine_inverse(const float T[3][4], float R[3][4]) {
  float m[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m[i][j] = T[i][j];
  auto cof = [&](int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  float c0[3] = {cof(0, 0), cof(1, 0), cof(2, 0)};
  float det = (c0[0] * m[0][0] + c0[1] * m[1][0]) + c0[2] * m[2][0];
  float invdet = 1.f / det;
  float inv[3][3];
  inv[0][0] = c0[0] * invdet;
  inv[0][1] = c0[1] * invdet;
  inv[0][2] = c0[2] * invdet;
  inv[1][0] = cof(0, 1) * invdet;
  inv[1][1] = cof(1, 1) * invdet;
  inv[1][2] = cof(2, 1) * invdet;
  inv[2][0] = cof(0, 2) * invdet;
  inv[2][1] = cof(1, 2) * invdet;
  inv[2][2] = cof(2, 2) * invdet;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = inv[i][j];
    R[i][3] = -((inv[i][0] * T[0][3] + inv[i][1] * T[1][3]) + inv[i][2] * T[2][3]);
  }
}

// Affine3f * Affine3f: linear = L1*L2, translation = L1*t2 + t1 (Eigen Transform product).
static void affine_mul(const float A[3][4], const float B[3][4], float R[3][4]) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = (A[i][0] * B[0][j] + A[i][1] * B[1][j]) + A[i][2] * B[2][j];
    R[i][3] = ((A[i][0] * B[0][3] + A[i][1] * B[1][3]) + A[i][2] * B[2][3]) + A[i][3];
  }
}

// ---------------------------------------------------------------------------------------------
// PatchWork (include/patchwork.h)
// ---------------------------------------------------------------------------------------------
struct PatchRecord {  // debug record of one processed patch (for stage-level parity tests)
  int zone, ring, sector, npts, decision;  // decision: 0 normal split, 1 all-nonground (tilted),
                                           // 2 all-nonground (elevation), 3 flatness recovery
  float normal[3], mean[3], sv[3], d;
};

class PatchWorkOracle {
 public:
  // include/patchwork.h:44-103,115-163
  int num_iter_ = 3, num_lpr_ = 20, num_min_pts_ = 10, num_zones_ = 4, num_rings_of_interest_ = 4;
  double sensor_height_ = 1.723;
  double th_seeds_ = 0.3, th_dist_ = 0.1, max_range_ = 80.0, min_range_ = 2.7, uprightness_thr_ = 0.707,
         adaptive_seed_selection_margin_ = -1.1;
  double min_range_z2_, min_range_z3_, min_range_z4_;
  int num_sectors_each_zone_[4] = {16, 32, 54, 32};
  int num_rings_each_zone_[4] = {2, 4, 4, 4};
  double elevation_thr_[4] = {-1.2, -0.9984, -0.851, -0.605};
  double flatness_thr_[4] = {0.0, 0.000125, 0.000185, 0.000185};
  double sector_sizes[4], ring_sizes[4], min_ranges[4];

  // plane state persists across calls exactly like the reference's members (patchwork.h:136-141)
  float d_ = 0, th_dist_d_ = 0, normal_[3] = {0, 0, 0}, singular_values_[3] = {0, 0, 0};
  float cov_[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, pc_mean_[4] = {0, 0, 0, 0};

  std::vector<std::vector<std::vector<Cloud>>> czm;  // [zone][ring][sector]
  std::vector<PatchRecord> records;
  bool keep_records = false;

  PatchWorkOracle() {
    min_range_z2_ = (7 * min_range_ + max_range_) / 8.0;
    min_range_z3_ = (3 * min_range_ + max_range_) / 4.0;
    min_range_z4_ = (min_range_ + max_range_) / 2.0;
    min_ranges[0] = min_range_;
    min_ranges[1] = min_range_z2_;
    min_ranges[2] = min_range_z3_;
    min_ranges[3] = min_range_z4_;
    ring_sizes[0] = (min_range_z2_ - min_range_) / num_rings_each_zone_[0];
    ring_sizes[1] = (min_range_z3_ - min_range_z2_) / num_rings_each_zone_[1];
    ring_sizes[2] = (min_range_z4_ - min_range_z3_) / num_rings_each_zone_[2];
    ring_sizes[3] = (max_range_ - min_range_z4_) / num_rings_each_zone_[3];
    for (int k = 0; k < 4; ++k) sector_sizes[k] = 2 * M_PI / num_sectors_each_zone_[k];
    czm.resize(4);
    for (int k = 0; k < 4; ++k) {
      czm[k].resize(num_rings_each_zone_[k]);
      for (auto& r : czm[k]) r.resize(num_sectors_each_zone_[k]);
    }
  }
  void set_sensor(const double& h) { sensor_height_ = h; }  // patchwork.h:401-403

  static bool point_z_cmp(Pt a, Pt b) { return a.z < b.z; }  // patchwork.h:33-35

  double xy2theta(const double& x, const double& y) {  // patchwork.h:417-423
    if (y >= 0) return atan2(y, x);
    return 2 * M_PI + atan2(y, x);
  }
  double xy2radius(const double& x, const double& y) { return sqrt(pow(x, 2) + pow(y, 2)); }  // :426-428

  void pc2czm(const Cloud& src) {  // patchwork.h:431-459
    for (auto const& pt : src) {
      int ring_idx, sector_idx;
      double r = xy2radius(pt.x, pt.y);
      if ((r <= max_range_) && (r > min_range_)) {
        double theta = xy2theta(pt.x, pt.y);
        int k = (r < min_range_z2_) ? 0 : (r < min_range_z3_) ? 1 : (r < min_range_z4_) ? 2 : 3;
        ring_idx = std::min(static_cast<int>(((r - min_ranges[k]) / ring_sizes[k])), num_rings_each_zone_[k] - 1);
        sector_idx = std::min(static_cast<int>((theta / sector_sizes[k])), num_sectors_each_zone_[k] - 1);
        czm[k][ring_idx][sector_idx].emplace_back(pt);
      }
    }
  }

  void estimate_plane_(const Cloud& ground) {  // patchwork.h:217-232
    mean_and_cov_pcl18(ground, cov_, pc_mean_);
    float U[3][3];
    jacobi_svd_3x3(cov_, U, singular_values_);
    normal_[0] = U[0][2];
    normal_[1] = U[1][2];
    normal_[2] = U[2][2];
    // (normal_.transpose() * seeds_mean)(0,0): Eigen coefficient-based product, 3-term unrolled
    // redux = a0 + (a1 + a2)  [recollection of Eigen 3.3.4 Redux.h redux_novec_unroller]
    d_ = -(normal_[0] * pc_mean_[0] + (normal_[1] * pc_mean_[1] + normal_[2] * pc_mean_[2]));
    th_dist_d_ = (float)(th_dist_ - (double)d_);
  }

  void extract_initial_seeds_(const int zone_idx, const Cloud& p_sorted, Cloud& init_seeds) {  // :235-268
    init_seeds.clear();
    double sum = 0;
    int cnt = 0;
    int init_idx = 0;
    if (zone_idx == 0) {
      for (size_t i = 0; i < p_sorted.size(); i++) {
        if (p_sorted[i].z < adaptive_seed_selection_margin_ * sensor_height_)
          ++init_idx;
        else
          break;
      }
    }
    for (size_t i = init_idx; i < p_sorted.size() && cnt < num_lpr_; i++) {
      sum += p_sorted[i].z;
      cnt++;
    }
    double lpr_height = cnt != 0 ? sum / cnt : 0;
    for (size_t i = 0; i < p_sorted.size(); i++)
      if (p_sorted[i].z < lpr_height + th_seeds_) init_seeds.push_back(p_sorted[i]);
  }

  void extract_piecewiseground(const int zone_idx, const Cloud& src, Cloud& dst, Cloud& non_ground_dst) {  // :463-504
    Cloud ground_pc_;
    dst.clear();
    non_ground_dst.clear();
    extract_initial_seeds_(zone_idx, src, ground_pc_);
    for (int i = 0; i < num_iter_; i++) {
      estimate_plane_(ground_pc_);
      ground_pc_.clear();
      // result = points * normal_ : Eigen product, per row (x*n0 + y*n1) + z*n2 in float, no FMA
      for (size_t r = 0; r < src.size(); r++) {
        float result = (src[r].x * normal_[0] + src[r].y * normal_[1]) + src[r].z * normal_[2];
        if (i < num_iter_ - 1) {
          if (result < th_dist_d_) ground_pc_.push_back(src[r]);
        } else {
          if (result < th_dist_d_)
            dst.push_back(src[r]);
          else
            non_ground_dst.push_back(src[r]);
        }
      }
    }
  }

  // patchwork.h:278-398.  dropped[] (optional) receives per-src outcome for points that vanish.
  void estimate_ground(const Cloud& cloud_in, Cloud& cloud_out, Cloud& cloud_nonground, std::vector<uint8_t>* cls) {
    Cloud laserCloudIn = cloud_in;
    std::sort(laserCloudIn.begin(), laserCloudIn.end(), point_z_cmp);
    size_t skip = 0;
    for (size_t i = 0; i < laserCloudIn.size(); i++) {
      if (laserCloudIn[i].z < -1.8 * sensor_height_)
        skip++;
      else
        break;
    }
    if (cls)
      for (size_t i = 0; i < skip; ++i) (*cls)[laserCloudIn[i].src] = SCVOD_PT_DROPPED_LOW;
    laserCloudIn.erase(laserCloudIn.begin(), laserCloudIn.begin() + skip);
    for (int k = 0; k < 4; ++k)
      for (auto& r : czm[k])
        for (auto& s : r) s.clear();
    if (cls)
      for (auto& p : laserCloudIn) (*cls)[p.src] = SCVOD_PT_DROPPED_RANGE;  // overwritten below if binned
    pc2czm(laserCloudIn);
    cloud_out.clear();
    cloud_nonground.clear();
    records.clear();
    Cloud regionwise_ground_, regionwise_nonground_;
    int concentric_idx = 0;
    for (int k = 0; k < num_zones_; ++k) {
      auto& zone = czm[k];  // (the reference deep-copies the zone here, :328; no effect on results)
      for (int ring_idx = 0; ring_idx < num_rings_each_zone_[k]; ++ring_idx) {
        for (int sector_idx = 0; sector_idx < num_sectors_each_zone_[k]; ++sector_idx) {
          const Cloud& patch = zone[ring_idx][sector_idx];
          if (patch.size() > (size_t)num_min_pts_) {
            extract_piecewiseground(k, patch, regionwise_ground_, regionwise_nonground_);
            const double ground_z_vec = std::abs(normal_[2]);
            const double ground_z_elevation = pc_mean_[2];
            const float minsv = std::min(singular_values_[0], std::min(singular_values_[1], singular_values_[2]));
            const double surface_variable = minsv / (singular_values_[0] + singular_values_[1] + singular_values_[2]);
            int decision = 0;
            if (ground_z_vec < uprightness_thr_) {
              decision = 1;
            } else if (concentric_idx < num_rings_of_interest_) {
              if (ground_z_elevation > elevation_thr_[ring_idx + 2 * k]) {
                if (flatness_thr_[ring_idx + 2 * k] > surface_variable)
                  decision = 3;
                else
                  decision = 2;
              }
            }
            if (decision == 1 || decision == 2) {
              cloud_nonground.insert(cloud_nonground.end(), regionwise_ground_.begin(), regionwise_ground_.end());
              cloud_nonground.insert(cloud_nonground.end(), regionwise_nonground_.begin(), regionwise_nonground_.end());
            } else {
              cloud_out.insert(cloud_out.end(), regionwise_ground_.begin(), regionwise_ground_.end());
              cloud_nonground.insert(cloud_nonground.end(), regionwise_nonground_.begin(), regionwise_nonground_.end());
            }
            if (keep_records) {
              PatchRecord rec;
              rec.zone = k;
              rec.ring = ring_idx;
              rec.sector = sector_idx;
              rec.npts = (int)patch.size();
              rec.decision = decision;
              for (int t = 0; t < 3; ++t) {
                rec.normal[t] = normal_[t];
                rec.mean[t] = pc_mean_[t];
                rec.sv[t] = singular_values_[t];
              }
              rec.d = d_;
              records.push_back(rec);
            }
          } else if (cls) {
            for (auto& p : patch) (*cls)[p.src] = SCVOD_PT_DROPPED_SPARSE;
          }
        }
        ++concentric_idx;
      }
    }
    if (cls) {
      for (auto& p : cloud_nonground) (*cls)[p.src] = SCVOD_PT_STATIC;
      for (auto& p : cloud_out) (*cls)[p.src] = SCVOD_PT_GROUND;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// SSC (include/utility.h, src/ssc.cpp)
// ---------------------------------------------------------------------------------------------
struct PointAPRI {  // utility.h:96-106
  float x, y, z, range, angle, azimuth, intensity = 0.f;
  int range_idx = -1, sector_idx = -1, azimuth_idx = -1, voxel_idx = -1;
};
struct Voxel {  // utility.h:109-119
  int range_idx, sector_idx, azimuth_idx;
  int label = -1;
  float cx, cy, cz, cintensity;
  std::vector<int> ptIdx;
  std::vector<float> intensity_record;
  float intensity_av = 0.f, intensity_cov = 0.f;
};
struct Cluster {  // utility.h:142-162
  int track_id = -1, name = -1, type = -1, state = -1;
  Pt bb_min = {0.f, 0.f, 0.f, 0.f, -1}, bb_max = {0.f, 0.f, 0.f, 0.f, -1};  // a default pcl::PointXYZI pair: zeros (PCL 1.8 point_types.hpp)
  std::vector<int> occupy_pts, occupy_voxels;
  Cloud cloud;
};
struct Frame {  // utility.h:165-185
  int id = 0, max_name = 0;
  Cloud cloud_use;
  std::unordered_map<int, Voxel> hash_cloud;
  std::unordered_map<int, Cluster> cluster_set;
  // bookkeeping for parity tests (not in the reference)
  int n_in = 0;
  std::vector<int> ground_src, nonground_src, apri_src, apri_vid;
  std::vector<int> pt_cluster[3];
  int n_clusters[3] = {0, 0, 0};
  std::vector<uint8_t> cls;  // per input point
};

static bool sort1(const std::pair<int, Cluster>& a, const std::pair<int, Cluster>& b) {  // ssc.cpp:24-26
  return a.second.occupy_voxels >= b.second.occupy_voxels;
}

class SSCOracle {
 public:
  scvod_params P;
  int range_num, sector_num, azimuth_num, bin_num;
  PatchWorkOracle pw;
  std::vector<PointAPRI> apri_vec;
  std::unordered_map<int, Voxel> hash_cloud;
  Frame frame_ssc;
  Cloud cloud_use;
  std::vector<Frame> frame_set;
  int name = 0;
  int id = 0;

  explicit SSCOracle(const scvod_params& p) : P(p) {  // ssc.cpp:32-39
    range_num = (int)std::ceil((P.max_dis - P.min_dis) / P.range_res);
    sector_num = (int)std::ceil((P.max_angle - P.min_angle) / P.sector_res);
    azimuth_num = (int)std::ceil((P.max_azimuth - P.min_azimuth) / P.azimuth_res);
    bin_num = range_num * sector_num * azimuth_num;
  }

  // utility.h:346-354
  template <typename T>
  float rad2deg(const T& radians) {
    return (float)radians * 180.0 / M_PI;
  }
  template <typename T>
  float deg2rad(const T& degrees) {
    return (float)degrees * M_PI / 180.0;
  }
  float pointDistance2d(const Pt& p1) { return (float)std::sqrt((p1.x) * (p1.x) + (p1.y) * (p1.y)); }  // :371-374
  float getPolarAngle(const Pt& p) {                                                                  // :376-387
    if (p.x == 0 && p.y == 0) {
      return 0.f;
    } else if (p.y >= 0) {
      return (float)rad2deg((float)atan2f(p.y, p.x));
    } else {
      return (float)rad2deg((float)atan2f(p.y, p.x) + 2 * M_PI);
    }
  }
  float getAzimuth(const Pt& p) { return (float)rad2deg((float)atan2f(p.z, (float)pointDistance2d(p))); }  // :389-392

  void binPoint(const Pt& pt, float& dis, float& angle, float& azimuth, int& ri, int& si, int& ei, int& vid) {
    // ssc.cpp:158-160,185-188 (same arithmetic reused at :1280-1286 and :1187-1193)
    dis = pointDistance2d(pt);
    angle = getPolarAngle(pt);
    azimuth = getAzimuth(pt);
    ri = std::ceil((dis - P.min_dis) / P.range_res) - 1;
    si = std::ceil((angle - P.min_angle) / P.sector_res) - 1;
    ei = std::ceil((azimuth - P.min_azimuth) / P.azimuth_res) - 1;
    vid = ei * range_num * sector_num + ri * sector_num + si;
  }

  void reset() {  // ssc.cpp:79-86
    frame_ssc = Frame();
    apri_vec.clear();
    hash_cloud.clear();
    cloud_use.clear();
  }

  Cloud extractGroudByPatchWork(const Cloud& in) {  // ssc.cpp:88-96
    Cloud g, ng;
    pw.set_sensor(P.sensor_height);
    pw.estimate_ground(in, g, ng, &frame_ssc.cls);
    for (auto& p : g) frame_ssc.ground_src.push_back(p.src);
    for (auto& p : ng) frame_ssc.nonground_src.push_back(p.src);
    return ng;
  }

  void makeApriVec(const Cloud& cloud_) {  // ssc.cpp:155-195
    for (size_t i = 0; i < cloud_.size(); i++) {
      Pt pt = cloud_[i];
      float dis, angle, azimuth;
      int ri, si, ei, vid;
      binPoint(pt, dis, angle, azimuth, ri, si, ei, vid);
      if (dis < P.min_dis || dis > P.max_dis || angle < P.min_angle || angle > P.max_angle ||
          azimuth < P.min_azimuth || azimuth > P.max_azimuth) {
        frame_ssc.cls[pt.src] = SCVOD_PT_GATED_OUT;  // cloud_eva_static
        continue;
      }
      cloud_use.push_back(pt);
      frame_ssc.cloud_use.push_back(pt);
      PointAPRI apri;
      apri.x = pt.x;
      apri.y = pt.y;
      apri.z = pt.z;
      apri.range = dis;
      apri.angle = angle;
      apri.azimuth = azimuth;
      apri.intensity = pt.intensity;
      apri.range_idx = ri;
      apri.sector_idx = si;
      apri.azimuth_idx = ei;
      apri.voxel_idx = vid;
      if (apri.voxel_idx > bin_num) continue;  // unreachable (SURVEY hard part 7)
      apri_vec.emplace_back(apri);
      frame_ssc.apri_src.push_back(pt.src);
      frame_ssc.apri_vid.push_back(vid);
    }
  }

  void makeHashCloud(const std::vector<PointAPRI>& apriIn_) {  // ssc.cpp:253-289
    for (size_t i = 0; i < apriIn_.size(); i++) {
      PointAPRI apri = apriIn_[i];
      auto it_find = hash_cloud.find(apri.voxel_idx);
      if (it_find != hash_cloud.end()) {
        it_find->second.ptIdx.emplace_back(i);
        it_find->second.intensity_record.emplace_back(apri.intensity);
        it_find->second.intensity_av += apri.intensity;
      } else {
        Voxel voxel;
        voxel.ptIdx.emplace_back(i);
        voxel.intensity_record.emplace_back(apri.intensity);
        voxel.intensity_av += apri.intensity;
        voxel.range_idx = apri.range_idx;
        voxel.sector_idx = apri.sector_idx;
        voxel.azimuth_idx = apri.azimuth_idx;
        float range_center = (apri.range_idx * 2 + 1) / 2 * P.range_res + P.min_dis;
        float sector_center = deg2rad((apri.sector_idx * 2 + 1) / 2 * P.sector_res) + P.min_angle;
        float azimuth_center = deg2rad((apri.azimuth_idx * 2 + 1) / 2 * P.azimuth_res) + deg2rad(P.min_azimuth);
        voxel.cx = range_center * std::cos(sector_center);
        voxel.cy = range_center * std::sin(sector_center);
        voxel.cz = range_center * std::tan(azimuth_center);
        voxel.cintensity = apri.voxel_idx;
        hash_cloud.insert(std::make_pair(apri.voxel_idx, voxel));
      }
    }
    for (auto& vox : hash_cloud) {
      vox.second.intensity_av /= vox.second.ptIdx.size();
      for (auto& in : vox.second.intensity_record) {
        vox.second.intensity_cov += std::pow((in - vox.second.intensity_av), 2);
      }
      vox.second.intensity_cov /= vox.second.ptIdx.size();
    }
  }

  void process(const Cloud& cloudIn_) {  // ssc.cpp:224-251 (file I/O + dead visualisation omitted)
    frame_ssc.id = id;
    frame_ssc.n_in = (int)cloudIn_.size();
    frame_ssc.cls.assign(cloudIn_.size(), SCVOD_PT_STATIC);
    Cloud ng_cloud = extractGroudByPatchWork(cloudIn_);
    makeApriVec(ng_cloud);
    makeHashCloud(apri_vec);
  }

  std::vector<int> findVoxelNeighbors(const int& range_idx_, const int& sector_idx_, const int& azimuth_idx_, int size_) {
    // ssc.cpp:395-411
    std::vector<int> neighborIdxs;
    if (range_idx_ > range_num * 0.6) size_ = 1;
    for (int x = range_idx_ - size_; x <= range_idx_ + size_; x++) {
      if (x > range_num - 1 || x < 0) continue;
      for (int y = sector_idx_ - size_; y <= sector_idx_ + size_; y++) {
        if (y > sector_num - 1 || y < 0) continue;
        for (int z = azimuth_idx_ - size_; z <= azimuth_idx_ + size_; z++) {
          if (z > azimuth_num - 1 || z < 0) continue;
          neighborIdxs.emplace_back(x * sector_num + y + z * range_num * sector_num);
        }
      }
    }
    return neighborIdxs;
  }

  void mergeClusters(std::vector<int>& clusterIdxs_, const int& idx1_, const int& idx2_) {  // ssc.cpp:413-419
    for (size_t i = 0; i < clusterIdxs_.size(); i++)
      if (clusterIdxs_[i] == idx1_) clusterIdxs_[i] = idx2_;
  }

  static void sampleVec(std::vector<int>& v) {  // utility.h:452-456
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  static void addVec(std::vector<int>& a, const std::vector<int>& b) { a.insert(a.end(), b.begin(), b.end()); }  // :440-443
  static void reduceVec(std::vector<int>& a, const std::vector<int>& b) {                                      // :445-450
    for (auto it = b.begin(); it != b.end(); it++) a.erase(std::remove(a.begin(), a.end(), *it), a.end());
  }
  static bool findNameInVec(const int& n, const std::vector<int>& v) { return std::count(v.begin(), v.end(), n) != 0; }  // :458-465
  void getCloudByVec(const Cloud& c, const std::vector<int>& v, Cloud& out) {  // :432-437
    for (auto& it : v) out.push_back(c[it]);
  }

  void clusterAndCreateFrame(const std::vector<PointAPRI>& apri_vec_, std::unordered_map<int, Voxel>& hash_cloud_) {
    // ssc.cpp:299-393
    int cluster_name = 4;
    std::vector<int> clusterIdxs = std::vector<int>(apri_vec_.size(), -1);
    for (size_t i = 0; i < apri_vec_.size(); i++) {
      PointAPRI apri = apri_vec_[i];
      std::vector<int> neighbors;
      auto it_find1 = hash_cloud_.find(apri.voxel_idx);
      if (it_find1 != hash_cloud_.end()) {
        std::vector<int> neighbor = findVoxelNeighbors(apri.range_idx, apri.sector_idx, apri.azimuth_idx, 1);
        for (size_t k = 0; k < neighbor.size(); k++) {
          auto it_find2 = hash_cloud_.find(neighbor[k]);
          if (it_find2 != hash_cloud_.end()) addVec(neighbors, it_find2->second.ptIdx);
        }
      }
      if (neighbors.size() > 0) {
        for (size_t n = 0; n < neighbors.size(); n++) {
          int oc = clusterIdxs[i];
          int nc = clusterIdxs[neighbors[n]];
          if (oc != -1 && nc != -1) {
            if (oc != nc) mergeClusters(clusterIdxs, oc, nc);
          } else {
            if (nc != -1) {
              clusterIdxs[i] = nc;
            } else {
              if (oc != -1) clusterIdxs[neighbors[n]] = oc;
            }
          }
        }
      }
      if (clusterIdxs[i] == -1) {
        cluster_name++;
        clusterIdxs[i] = cluster_name;
        for (size_t m = 0; m < neighbors.size(); m++) clusterIdxs[neighbors[m]] = cluster_name;
      }
    }
    frame_ssc.max_name = cluster_name++;

    std::unordered_map<int, std::vector<int>> cluster_pt, cluster_vox;
    for (size_t i = 0; i < clusterIdxs.size(); i++) {
      auto it_p = cluster_pt.find(clusterIdxs[i]);
      auto it_v = cluster_vox.find(clusterIdxs[i]);
      if (it_p != cluster_pt.end()) {
        it_p->second.emplace_back(i);
        it_v->second.emplace_back(apri_vec_[i].voxel_idx);
      } else {
        std::vector<int> pt_vec, vox_vec;
        pt_vec.emplace_back(i);
        vox_vec.emplace_back(apri_vec_[i].voxel_idx);
        cluster_pt.insert(std::make_pair(clusterIdxs[i], pt_vec));
        cluster_vox.insert(std::make_pair(clusterIdxs[i], vox_vec));
      }
    }
    for (auto& c : cluster_pt) {
      Cluster cluster;
      cluster.name = c.first;
      cluster.occupy_pts = c.second;
      cluster.occupy_voxels = cluster_vox[c.first];
      getCloudByVec(cloud_use, c.second, cluster.cloud);
      sampleVec(cluster.occupy_voxels);
      frame_ssc.cluster_set.insert(std::make_pair(cluster.name, cluster));
    }
    for (auto& c : frame_ssc.cluster_set)
      for (auto& v : c.second.occupy_voxels) hash_cloud[v].label = c.first;
    frame_ssc.pt_cluster[0] = clusterIdxs;
    frame_ssc.n_clusters[0] = (int)frame_ssc.cluster_set.size();
  }

  void refineClusterByIntensity(Frame& frame) {  // ssc.cpp:571-635
    int iter = P.iteration;
    while (iter) {
      std::vector<std::pair<int, Cluster>> clusters(frame.cluster_set.begin(), frame.cluster_set.end());
      std::sort(clusters.begin(), clusters.end(), sort1);
      std::vector<int> invalid_name;
      std::unordered_map<int, std::vector<int>> fusion_map;
      for (auto& c : clusters) {
        if (findNameInVec(c.first, invalid_name)) continue;
        std::vector<int> neighbor_name, neighbor_vox;
        for (auto& v : c.second.occupy_voxels) {
          std::vector<int> vox = findVoxelNeighbors(hash_cloud[v].range_idx, hash_cloud[v].sector_idx,
                                                    hash_cloud[v].azimuth_idx, P.search_c);
          for (auto& n : vox) {
            auto it_find = hash_cloud.find(n);
            if (it_find != hash_cloud.end() && hash_cloud[n].intensity_cov <= P.intensity_cov &&
                std::fabs(hash_cloud[v].intensity_av - hash_cloud[n].intensity_av) <= P.intensity_diff) {
              neighbor_vox.emplace_back(n);
            }
          }
        }
        sampleVec(neighbor_vox);
        for (auto& n : neighbor_vox)
          if (!findNameInVec(hash_cloud[n].label, invalid_name)) neighbor_name.emplace_back(hash_cloud[n].label);
        sampleVec(neighbor_name);
        if (neighbor_name.size() > 1) {
          addVec(invalid_name, neighbor_name);
          fusion_map.insert(std::make_pair(c.first, neighbor_name));
        }
        sampleVec(invalid_name);
      }
      for (auto& cn : fusion_map) {
        Cluster cluster_fusion;
        for (auto& f : cn.second) {
          cluster_fusion.name = f;
          addVec(cluster_fusion.occupy_pts, frame.cluster_set[f].occupy_pts);
          addVec(cluster_fusion.occupy_voxels, frame.cluster_set[f].occupy_voxels);
          Cloud& fc = frame.cluster_set[f].cloud;
          cluster_fusion.cloud.insert(cluster_fusion.cloud.end(), fc.begin(), fc.end());
          frame.cluster_set.erase(f);
        }
        for (auto v : cluster_fusion.occupy_voxels) hash_cloud[v].label = cluster_fusion.name;
        frame.cluster_set.insert(std::make_pair(cluster_fusion.name, cluster_fusion));
      }
      iter--;
    }
  }

  // pcl::getMinMax3D dense branch (PCL 1.8 common/impl/common.hpp): componentwise min/max.
  static void getBoundingBoxOfCloud(const Cloud& c, Pt& mn, Pt& mx) {  // ssc.cpp:421-425
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (auto& p : c) {
      lo[0] = std::min(lo[0], p.x);
      lo[1] = std::min(lo[1], p.y);
      lo[2] = std::min(lo[2], p.z);
      hi[0] = std::max(hi[0], p.x);
      hi[1] = std::max(hi[1], p.y);
      hi[2] = std::max(hi[2], p.z);
    }
    mn.x = lo[0];
    mn.y = lo[1];
    mn.z = lo[2];
    mx.x = hi[0];
    mx.y = hi[1];
    mx.z = hi[2];
  }

  void refineClusterByBoundingBox(Frame& frame) {  // ssc.cpp:437-467
    std::vector<int> erase_id;
    for (auto& c : frame.cluster_set) {
      getBoundingBoxOfCloud(c.second.cloud, c.second.bb_min, c.second.bb_max);
      Pt point_min = c.second.bb_min, point_max = c.second.bb_max;
      float diff_z = point_max.z - point_min.z;
      if (point_min.z > 0.f || (c.second.occupy_pts.size() < (size_t)P.toBeClass) || diff_z < 0.2) {
        erase_id.emplace_back(c.first);
      }
    }
    for (auto& e : erase_id) {
      for (auto& v : frame.cluster_set[e].occupy_voxels) hash_cloud[v].label = -1;
      frame.cluster_set.erase(e);
    }
  }

  void recordPointClusters(Frame& frame, int stage) {
    frame.pt_cluster[stage].assign(apri_vec.size(), -1);
    for (auto& c : frame.cluster_set)
      for (int p : c.second.occupy_pts) frame.pt_cluster[stage][p] = c.first;
    frame.n_clusters[stage] = (int)frame.cluster_set.size();
  }

  void segment() {  // ssc.cpp:637-656
    clusterAndCreateFrame(apri_vec, hash_cloud);
    refineClusterByIntensity(frame_ssc);
    recordPointClusters(frame_ssc, 1);
    refineClusterByBoundingBox(frame_ssc);
    recordPointClusters(frame_ssc, 2);
    frame_ssc.hash_cloud = hash_cloud;
  }

  void recognize(Frame& frame) {  // ssc.cpp:834-895 with getDescriptorByEigenValue :723-751
    for (auto& c : frame.cluster_set) {
      Pt point_min = c.second.bb_min, point_max = c.second.bb_max;
      double diff_x = point_max.x - point_min.x;
      double diff_y = point_max.y - point_min.y;
      double square = diff_x * diff_y;  // f_11(0,7)
      double f6 = point_max.z, f9 = point_min.z;
      if (square > P.car_square) {
        // regionGrowing (building vs tree, ssc.cpp:797-832) needs PCL normals + RegionGrowing and does
        // not influence dynamic/static labels: both outcomes are "non-car".  Reported as tree.
        c.second.type = P.tree;
      } else {
        if (f9 < P.min_z && square < P.car_square && f6 < P.max_z)
          c.second.type = P.car;
        else
          c.second.type = P.tree;
      }
    }
  }

  void transformCloud(const Cloud& in, const float T[3][4], Cloud& out) {  // utility.h:394-406
    out.resize(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
      out[i].x = T[0][0] * in[i].x + T[0][1] * in[i].y + T[0][2] * in[i].z + T[0][3];
      out[i].y = T[1][0] * in[i].x + T[1][1] * in[i].y + T[1][2] * in[i].z + T[1][3];
      out[i].z = T[2][0] * in[i].x + T[2][1] * in[i].y + T[2][2] * in[i].z + T[2][3];
      out[i].intensity = in[i].intensity;
      out[i].src = -1;
    }
  }

  void tracking(Frame& frame_pre_, Frame& frame_next_, const float pose_pre_[6], const float pose_next_[6]) {
    // ssc.cpp:1250-1426 (colours / cv::RNG omitted: they do not feed labels)
    float trans_next[3][4], trans_pre[3][4], inv_next[3][4], trans_np[3][4];
    pcl_get_transformation(pose_next_, trans_next);
    pcl_get_transformation(pose_pre_, trans_pre);
    affine_inverse(trans_next, inv_next);
    affine_mul(inv_next, trans_pre, trans_np);
    for (auto& c : frame_pre_.cluster_set) {
      if (c.second.type != P.car) continue;
      if (c.second.track_id == -1) {
        c.second.track_id = name;
        name++;
      }
      Cloud cluster;
      transformCloud(c.second.cloud, trans_np, cluster);
      std::unordered_map<int, std::vector<int>> remap_name;
      for (size_t k = 0; k < cluster.size(); k++) {
        Pt pt = cluster[k];
        float dis, angle, azimuth;
        int ri, si, ei, voxel_idx;
        binPoint(pt, dis, angle, azimuth, ri, si, ei, voxel_idx);
        auto it_find = frame_next_.hash_cloud.find(voxel_idx);
        if (it_find != frame_next_.hash_cloud.end() && it_find->second.label != -1) {
          auto l_find = remap_name.find(it_find->second.label);
          if (l_find == remap_name.end()) {
            std::vector<int> vec;
            vec.emplace_back(it_find->first);
            remap_name.insert(std::make_pair(it_find->second.label, vec));
          } else {
            l_find->second.emplace_back(it_find->first);
          }
        }
      }
      for (auto& re : remap_name) sampleVec(re.second);

      if (remap_name.size() == 0) {
        c.second.state = 1;
      } else if (remap_name.size() == 1) {
        auto it = remap_name.begin();
        float ratio = (float)it->second.size() / (float)frame_next_.cluster_set[it->first].occupy_voxels.size();
        if ((ratio) < P.occupancy) {
          if (frame_next_.cluster_set[it->first].type == P.car) {
            c.second.state = 1;
          } else {
            c.second.state = 0;
            c.second.type = frame_next_.cluster_set[it->first].type;
            Cluster cluster_new;
            cluster_new.track_id = c.second.track_id;
            cluster_new.name = frame_next_.max_name++;
            cluster_new.type = frame_next_.cluster_set[it->first].type;
            cluster_new.occupy_voxels = it->second;
            reduceVec(frame_next_.cluster_set[it->first].occupy_voxels, cluster_new.occupy_voxels);
            for (auto& v : it->second) {
              frame_next_.hash_cloud[v].label = cluster_new.name;
              addVec(cluster_new.occupy_pts, frame_next_.hash_cloud[v].ptIdx);
            }
            getCloudByVec(frame_next_.cloud_use, cluster_new.occupy_pts, cluster_new.cloud);
            reduceVec(frame_next_.cluster_set[it->first].occupy_pts, cluster_new.occupy_pts);
            frame_next_.cluster_set.insert(std::make_pair(cluster_new.name, cluster_new));
          }
        } else {
          if (frame_next_.cluster_set[it->first].type == P.car) {
            c.second.state = 0;
            frame_next_.cluster_set[it->first].track_id = c.second.track_id;
            Cloud& nc = frame_next_.cluster_set[it->first].cloud;
            nc.insert(nc.end(), cluster.begin(), cluster.end());
          }
        }
      } else {
        c.second.state = 0;
        Cluster cluster_new;
        cluster_new.track_id = c.second.track_id;
        cluster_new.name = frame_next_.max_name++;
        cluster_new.type = P.car;
        for (auto& re : remap_name) {
          if (frame_next_.cluster_set[re.first].type == P.car &&
              ((float)re.second.size() / (float)frame_next_.cluster_set[re.first].occupy_voxels.size()) >= P.occupancy) {
            addVec(cluster_new.occupy_pts, frame_next_.cluster_set[re.first].occupy_pts);
            addVec(cluster_new.occupy_voxels, frame_next_.cluster_set[re.first].occupy_voxels);
            frame_next_.cluster_set.erase(re.first);
          }
        }
        getCloudByVec(frame_next_.cloud_use, cluster_new.occupy_pts, cluster_new.cloud);
        for (auto& v : cluster_new.occupy_voxels) frame_next_.hash_cloud[v].label = cluster_new.name;
        frame_next_.cluster_set.insert(std::make_pair(cluster_new.name, cluster_new));
      }
    }
  }

  // SSC::intialization (ssc.cpp:1148-1248; dead code in the reference: the call is commented out at :1456-1470).
  // Every frame of the window is diffed against the frame with the fewest clusters; base-frame clusters that one
  // transformed cluster bridges with enough occupancy are fused.  Returns the index of the base frame; the
  // initialised frame is kept in init_frame.
  Frame init_frame;
  int initialization(const float* poses6, int nposes) {
    const int nf = (int)std::min<size_t>(frame_set.size(), (size_t)nposes);
    int max_num = 999999, id_based = 0;
    for (int i = 0; i < nf; i++) {
      if ((int)frame_set[i].cluster_set.size() <= max_num) {
        max_num = (int)frame_set[i].cluster_set.size();
        id_based = i;
      }
    }
    Frame frame_based = frame_set[id_based];
    float trans_based[3][4], inv_based[3][4];
    pcl_get_transformation(poses6 + 6 * id_based, trans_based);
    affine_inverse(trans_based, inv_based);
    for (int i = 0; i < nf; i++) {
      if (i == id_based) continue;
      Frame frame_i = frame_set[i];
      float trans_i[3][4], trans_bi[3][4];
      pcl_get_transformation(poses6 + 6 * i, trans_i);
      affine_mul(inv_based, trans_i, trans_bi);
      for (auto& c : frame_i.cluster_set) {
        Cloud cluster;
        transformCloud(c.second.cloud, trans_bi, cluster);
        std::unordered_map<int, std::vector<int>> remap_name;
        for (size_t k = 0; k < cluster.size(); k++) {
          Pt pt = cluster[k];
          float dis, angle, azimuth;
          int ri, si, ei, voxel_idx;
          binPoint(pt, dis, angle, azimuth, ri, si, ei, voxel_idx);
          auto it_find = frame_based.hash_cloud.find(voxel_idx);
          if (it_find != frame_based.hash_cloud.end() && it_find->second.label != -1) {
            auto l_find = remap_name.find(it_find->second.label);
            if (l_find == remap_name.end()) {
              std::vector<int> vec;
              vec.emplace_back(it_find->first);
              remap_name.insert(std::make_pair(it_find->second.label, vec));
            } else {
              l_find->second.emplace_back(it_find->first);
            }
          }
        }
        if (remap_name.size() > 1) {
          Cluster cluster_fusion;  // name stays -1 when no label passes the occupancy test (utility.h:154)
          std::vector<int> erase_id;
          for (auto& re : remap_name) {
            sampleVec(re.second);
            if (((float)re.second.size() / (float)frame_based.cluster_set[re.first].occupy_voxels.size()) >= P.occupancy) {
              erase_id.emplace_back(re.first);
              cluster_fusion.name = re.first;
              addVec(cluster_fusion.occupy_pts, frame_based.cluster_set[re.first].occupy_pts);
              addVec(cluster_fusion.occupy_voxels, frame_based.cluster_set[re.first].occupy_voxels);
              Cloud& src = frame_based.cluster_set[re.first].cloud;
              cluster_fusion.cloud.insert(cluster_fusion.cloud.end(), src.begin(), src.end());
            }
          }
          for (auto& e : erase_id) frame_based.cluster_set.erase(e);
          frame_based.cluster_set.insert(std::make_pair(cluster_fusion.name, cluster_fusion));
          for (auto& v : cluster_fusion.occupy_voxels) frame_based.hash_cloud[v].label = cluster_fusion.name;
        }
      }
    }
    // recognize(frame_based) (ssc.cpp:1242) reads Cluster::bounding_box as it is: the box refineClusterByBoundingBox stored for
    // the clusters of the base frame, and the zeros of a default-constructed pair for every fused cluster (ssc.cpp:1211-1233
    // never computes one), which the car rule then types as `tree` (min.z = 0 is not < min_z)
    recognize(frame_based);
    init_frame = frame_based;
    return id_based;
  }

  // one iteration of the scan loop of segDF (ssc.cpp:1435-1444) without saveSegCloud
  int pushScan(const float* xyzi, int n) {
    Cloud in(n);
    for (int i = 0; i < n; ++i) {
      in[i].x = xyzi[4 * i];
      in[i].y = xyzi[4 * i + 1];
      in[i].z = xyzi[4 * i + 2];
      in[i].intensity = xyzi[4 * i + 3];
      in[i].src = i;
    }
    process(in);
    segment();
    recognize(frame_ssc);
    frame_set.emplace_back(frame_ssc);
    reset();
    id += 1;
    return (int)frame_set.size() - 1;
  }

  int tracked = 0;  // frames [0, tracked] have been used as frame_pre_
  void trackAll(const float* poses6, int nposes) {  // ssc.cpp:1450-1452
    int nf = (int)std::min<size_t>(frame_set.size(), (size_t)nposes);
    for (int i = tracked; i + 1 < nf; i++) {
      tracking(frame_set[i], frame_set[i + 1], poses6 + 6 * i, poses6 + 6 * (i + 1));
      tracked = i + 1;
    }
  }

  // per-input-point class of frame f
  void frameLabels(int f, uint8_t* out) {
    Frame& fr = frame_set[f];
    std::vector<uint8_t> cls = fr.cls;
    // every apri point starts as UNCLUSTERED, then clusters overwrite
    for (size_t m = 0; m < fr.apri_src.size(); ++m) cls[fr.apri_src[m]] = SCVOD_PT_UNCLUSTERED;
    for (auto& c : fr.cluster_set) {
      uint8_t v = (c.second.state == 1) ? SCVOD_PT_DYNAMIC : SCVOD_PT_STATIC;
      for (int p : c.second.occupy_pts) cls[fr.apri_src[p]] = v;
    }
    std::memcpy(out, cls.data(), cls.size());
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// extern "C" surface for ctypes (tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------------------------
// frame f of the sequence, or the frame produced by orc_initialization when f == -1
static Frame& frame_ref(void* h, int f) { return f < 0 ? ((SSCOracle*)h)->init_frame : ((SSCOracle*)h)->frame_set[f]; }

extern "C" {

void* orc_create(const scvod_params* p) { return new SSCOracle(*p); }
void orc_destroy(void* h) { delete (SSCOracle*)h; }
void orc_grid_dims(void* h, int out[4]) {
  SSCOracle* s = (SSCOracle*)h;
  out[0] = s->range_num;
  out[1] = s->sector_num;
  out[2] = s->azimuth_num;
  out[3] = s->bin_num;
}
int orc_push_scan(void* h, const float* xyzi, int n) { return ((SSCOracle*)h)->pushScan(xyzi, n); }
void orc_track(void* h, const float* poses6, int nposes) { ((SSCOracle*)h)->trackAll(poses6, nposes); }
int orc_num_frames(void* h) { return (int)((SSCOracle*)h)->frame_set.size(); }
void orc_reset_frames(void* h) {
  SSCOracle* s = (SSCOracle*)h;
  s->frame_set.clear();
  s->tracked = 0;
  s->name = 0;
  s->id = 0;
}
void orc_frame_labels(void* h, int f, uint8_t* cls) { ((SSCOracle*)h)->frameLabels(f, cls); }
void orc_frame_counts(void* h, int f, int32_t c[9]) {
  Frame& fr = frame_ref(h, f);
  c[0] = fr.n_in;
  c[1] = (int)fr.ground_src.size();
  c[2] = (int)fr.nonground_src.size();
  c[3] = (int)fr.apri_src.size();
  c[4] = (int)fr.hash_cloud.size();
  c[5] = fr.n_clusters[0];
  c[6] = fr.n_clusters[1];
  c[7] = fr.n_clusters[2];
  c[8] = (int)fr.cluster_set.size();
}
void orc_frame_ground_order(void* h, int f, int32_t* g, int32_t* ng) {
  Frame& fr = ((SSCOracle*)h)->frame_set[f];
  if (g) std::memcpy(g, fr.ground_src.data(), fr.ground_src.size() * 4);
  if (ng) std::memcpy(ng, fr.nonground_src.data(), fr.nonground_src.size() * 4);
}
void orc_frame_apri(void* h, int f, int32_t* src, int32_t* vid) {
  Frame& fr = ((SSCOracle*)h)->frame_set[f];
  if (src) std::memcpy(src, fr.apri_src.data(), fr.apri_src.size() * 4);
  if (vid) std::memcpy(vid, fr.apri_vid.data(), fr.apri_vid.size() * 4);
}
int orc_initialization(void* h, const float* poses6, int nposes) { return ((SSCOracle*)h)->initialization(poses6, nposes); }

void orc_frame_voxels(void* h, int f, int32_t* vid, int32_t* count, float* av, float* cov, float* center, int32_t* tri,
                      int32_t* label) {
  Frame& fr = frame_ref(h, f);
  std::vector<int> keys;
  for (auto& v : fr.hash_cloud) keys.push_back(v.first);
  std::sort(keys.begin(), keys.end());
  for (size_t i = 0; i < keys.size(); ++i) {
    const Voxel& v = fr.hash_cloud[keys[i]];
    if (vid) vid[i] = keys[i];
    if (count) count[i] = (int)v.ptIdx.size();
    if (av) av[i] = v.intensity_av;
    if (cov) cov[i] = v.intensity_cov;
    if (center) {
      center[3 * i] = v.cx;
      center[3 * i + 1] = v.cy;
      center[3 * i + 2] = v.cz;
    }
    if (tri) {
      tri[3 * i] = v.range_idx;
      tri[3 * i + 1] = v.sector_idx;
      tri[3 * i + 2] = v.azimuth_idx;
    }
    if (label) label[i] = v.label;
  }
}
void orc_frame_point_cluster(void* h, int f, int stage, int32_t* name) {
  Frame& fr = ((SSCOracle*)h)->frame_set[f];
  std::memcpy(name, fr.pt_cluster[stage].data(), fr.pt_cluster[stage].size() * 4);
}
int orc_frame_clusters(void* h, int f, int cap, int32_t* name, int32_t* type, int32_t* state, int32_t* npts, int32_t* nvox,
                       float* bbox) {
  Frame& fr = frame_ref(h, f);
  int i = 0;
  for (auto& c : fr.cluster_set) {
    if (i >= cap) break;
    if (name) name[i] = c.first;
    if (type) type[i] = c.second.type;
    if (state) state[i] = c.second.state;
    if (npts) npts[i] = (int)c.second.occupy_pts.size();
    if (nvox) nvox[i] = (int)c.second.occupy_voxels.size();
    if (bbox) {
      bbox[6 * i] = c.second.bb_min.x;
      bbox[6 * i + 1] = c.second.bb_min.y;
      bbox[6 * i + 2] = c.second.bb_min.z;
      bbox[6 * i + 3] = c.second.bb_max.x;
      bbox[6 * i + 4] = c.second.bb_max.y;
      bbox[6 * i + 5] = c.second.bb_max.z;
    }
    ++i;
  }
  return (int)fr.cluster_set.size();
}

// stage-level helpers ---------------------------------------------------------------------------
void orc_bin(const scvod_params* p, const float* xyzi, int n, uint8_t* pass, int32_t* vid, int32_t* ri, int32_t* si,
             int32_t* ei, float* range, float* angle, float* azimuth) {
  SSCOracle s(*p);
  for (int i = 0; i < n; ++i) {
    Pt pt = {xyzi[4 * i], xyzi[4 * i + 1], xyzi[4 * i + 2], xyzi[4 * i + 3], i};
    float d, a, e;
    int r, sc, el, v;
    s.binPoint(pt, d, a, e, r, sc, el, v);
    if (pass)
      pass[i] = !(d < p->min_dis || d > p->max_dis || a < p->min_angle || a > p->max_angle || e < p->min_azimuth ||
                  e > p->max_azimuth);
    if (vid) vid[i] = v;
    if (ri) ri[i] = r;
    if (si) si[i] = sc;
    if (ei) ei[i] = el;
    if (range) range[i] = d;
    if (angle) angle[i] = a;
    if (azimuth) azimuth[i] = e;
  }
}

// returns number of processed patches; rec holds 15 floats per patch:
// zone, ring, sector, npts, decision, normal[3], mean[3], sv[3], d
int orc_ground(const float* xyzi, int n, double sensor_height, int32_t* ground_src, int32_t* n_ground, int32_t* ng_src,
               int32_t* n_ng, uint8_t* cls, float* rec, int rec_cap) {
  PatchWorkOracle pw;
  pw.keep_records = rec != nullptr;
  pw.set_sensor(sensor_height);
  Cloud in(n), g, ng;
  for (int i = 0; i < n; ++i) {
    in[i].x = xyzi[4 * i];
    in[i].y = xyzi[4 * i + 1];
    in[i].z = xyzi[4 * i + 2];
    in[i].intensity = xyzi[4 * i + 3];
    in[i].src = i;
  }
  std::vector<uint8_t> c(n, SCVOD_PT_STATIC);
  pw.estimate_ground(in, g, ng, &c);
  for (size_t i = 0; i < g.size(); ++i) ground_src[i] = g[i].src;
  for (size_t i = 0; i < ng.size(); ++i) ng_src[i] = ng[i].src;
  *n_ground = (int)g.size();
  *n_ng = (int)ng.size();
  if (cls) std::memcpy(cls, c.data(), n);
  int np = (int)pw.records.size();
  if (rec)
    for (int i = 0; i < np && i < rec_cap; ++i) {
      const PatchRecord& r = pw.records[i];
      float* o = rec + 15 * i;
      o[0] = (float)r.zone;
      o[1] = (float)r.ring;
      o[2] = (float)r.sector;
      o[3] = (float)r.npts;
      o[4] = (float)r.decision;
      for (int t = 0; t < 3; ++t) {
        o[5 + t] = r.normal[t];
        o[8 + t] = r.mean[t];
        o[11 + t] = r.sv[t];
      }
      o[14] = r.d;
    }
  return np;
}

void orc_svd3(const float A[9], float U[9], float sv[3]) {
  float a[3][3], u[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = A[3 * i + j];
  jacobi_svd_3x3(a, u, sv);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) U[3 * i + j] = u[i][j];
}

void orc_relative_pose(const float pose_next6[6], const float pose_pre6[6], float T[12]) {
  float tn[3][4], tp[3][4], inv[3][4], r[3][4];
  pcl_get_transformation(pose_next6, tn);
  pcl_get_transformation(pose_pre6, tp);
  affine_inverse(tn, inv);
  affine_mul(inv, tp, r);
  std::memcpy(T, r, sizeof(r));
}

float orc_atan2f(float y, float x) { return atan2f(y, x); }
void orc_atan2f_many(const float* y, const float* x, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = atan2f(y[i], x[i]);
}

// CPU baseline: the per-scan stages of segDF for many scans on `nthreads` host threads (one private
// SSC per thread; the reference itself is single-threaded), then the serial tracking chain.
// Returns wall seconds; labels (optional) receives concatenated per-point classes.
double orc_run_sequence(const scvod_params* p, const float* xyzi, const int64_t* offsets, int nscans, const float* poses6,
                        int nthreads, uint8_t* labels) {
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads < 1) nthreads = 1;
  std::vector<Frame> frames(nscans);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    th.emplace_back([&, t]() {
      SSCOracle w(*p);
      for (int s = t; s < nscans; s += nthreads) {
        w.id = s;
        w.pushScan(xyzi + 4 * offsets[s], (int)(offsets[s + 1] - offsets[s]));
        frames[s] = std::move(w.frame_set.back());
        w.frame_set.clear();
      }
    });
  }
  for (auto& t : th) t.join();
  SSCOracle s(*p);
  s.frame_set = std::move(frames);
  if (poses6) s.trackAll(poses6, nscans);
  if (labels)
    for (int f = 0; f < nscans; ++f) s.frameLabels(f, labels + offsets[f]);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}


// The GPU arm's decomposition on the host cores: `nchunks` independent sequences (chunk c = scans [c*chunk, (c+1)*chunk) of the
// input, which may repeat: chunk c reads input chunk c % in_chunks), each one run like the reference runs a sequence (per-scan
// stages, then tracking(k, k+1) as soon as both frames exist) by ONE thread; threads pull chunks from a queue.  A frame is
// dropped once it has been `pre`, so a thread holds two frames at a time.  labels (may be null) = [nchunks * points of a chunk].
double orc_run_chunks(const scvod_params* p, const float* xyzi, const int64_t* offsets, int in_chunks, int chunk, const float* poses6,
                      int nchunks, int nthreads, uint8_t* labels) {
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    th.emplace_back([&]() {
      for (;;) {
        const int c = next.fetch_add(1);
        if (c >= nchunks) break;
        const int ic = c % in_chunks;
        const int64_t* off = offsets + (size_t)ic * chunk;
        const float* poses = poses6 + 6 * (size_t)ic * chunk;
        SSCOracle w(*p);
        for (int k = 0; k < chunk; ++k) {
          w.id = k;
          w.pushScan(xyzi + 4 * off[k], (int)(off[k + 1] - off[k]));
          if (k > 0) {
            w.tracking(w.frame_set[k - 1], w.frame_set[k], poses + 6 * (k - 1), poses + 6 * k);
            if (labels) w.frameLabels(k - 1, labels + (off[k - 1] - offsets[0]));  // only meaningful when nchunks <= in_chunks
            Frame& done = w.frame_set[k - 1];
            Frame().hash_cloud.swap(done.hash_cloud);
            Frame().cluster_set.swap(done.cluster_set);
            Cloud().swap(done.cloud_use);
          }
        }
        if (labels && chunk > 0) w.frameLabels(chunk - 1, labels + (off[chunk - 1] - offsets[0]));
      }
    });
  }
  for (auto& t : th) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
