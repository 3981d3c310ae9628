// scvod_knn.cu — k-nearest-neighbour normals on the device and the two reference routines built on them:
//
//  * SSC::intensityCalibrationByCurvature (reference src/ssc.cpp:98-153; its call at :234-235 is commented out in the reference):
//    clamp intensity to max_intensity, pcl::NormalEstimationOMP with k = search_num, cos of the angle between the normal and the
//    ray to the point, clamped to >= 0.3, intensity / cos capped at max_intensity.
//  * SSC::regionGrowing (src/ssc.cpp:797-832): pcl::NormalEstimation (k = 10) + pcl::RegionGrowing (10 neighbours, smoothness
//    10 deg, curvature threshold 1.2, segments of >= 20 points); a cluster is a building when its planar segments hold >= 20 % of
//    its points.  recognize() calls it only for clusters whose footprint exceeds car_square (:845-856) and the answer only decides
//    building vs tree, never a per-point class.
//
// k-NN: exact, over the uniform grid of scvod_grid.cuh, shell by shell (Chebyshev rings of cells) until the k-th best distance is
// covered by the rings searched; the neighbour list is sorted by distance and contains the query itself, like a kd-tree k-search
// on the cloud itself.  Normal = eigenvector of the smallest eigenvalue of the neighbours' covariance (cyclic Jacobi in double;
// PCL's float eigen33 is not reproduced bit for bit: direction within ~1e-4 rad, checked by tolerance), curvature = lambda_min /
// trace, flipped towards the origin (flipNormalTowardsViewpoint with the default viewpoint).  The region growing itself is PCL 1.8's
// sequential queue algorithm ([recollection] of segmentation/impl/region_growing.hpp), run on the host over the GPU's neighbour
// lists and normals: its result depends on visiting order by construction.
#include <cmath>
#include <cstring>
#include <queue>

#include "scvod_kernel_common.cuh"

namespace scvod {

namespace {

#include "scvod_grid.cuh"

constexpr int kMaxK = 16;

__device__ void jacobi3_min(double A[3][3], double& lmin, double& trace, double nrm[3]) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  trace = A[0][0] + A[1][1] + A[2][2];
  for (int sweep = 0; sweep < 16; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off == 0.0 || off <= 1e-32 * diag) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int im = 0;
  if (A[1][1] < A[im][im]) im = 1;
  if (A[2][2] < A[im][im]) im = 2;
  lmin = A[im][im];
  nrm[0] = V[0][im];
  nrm[1] = V[1][im];
  nrm[2] = V[2][im];
}

// one thread per query (queries = the cloud itself, in cell-major order so that a warp's queries share cells)
__global__ void __launch_bounds__(128) k_knn_normals(const float4* __restrict__ sorted, const int* __restrict__ sorted_idx, const int* __restrict__ start,
                                                     EGrid g, int n, int k, float4* __restrict__ normal_curv /* by original index */,
                                                     int32_t* __restrict__ nbr /* [n][k] by original index, may be null */) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const float4 q = sorted[s];
    const int qi = sorted_idx[s];
    float bd[kMaxK];
    int bi[kMaxK], bs[kMaxK];
    int cnt = 0;
    int cx, cy, cz;
    cell_coords(g, q.x, q.y, q.z, cx, cy, cz);
    const int rmax = max(max(g.nx, g.ny), g.nz);
    auto scan_run = [&](int lo, int hi) {
      for (int t = lo; t < hi; ++t) {
        const float4 p = __ldg(&sorted[t]);
        const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const int id = __ldg(&sorted_idx[t]);
        if (cnt == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && id < bi[k - 1]))) continue;
        int pos = cnt < k ? cnt : k - 1;  // insertion (ascending distance, ties by original index)
        while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && bi[pos - 1] > id))) {
          bd[pos] = bd[pos - 1];
          bi[pos] = bi[pos - 1];
          bs[pos] = bs[pos - 1];
          --pos;
        }
        bd[pos] = d2;
        bi[pos] = id;
        bs[pos] = t;
        if (cnt < k) ++cnt;
      }
    };
    for (int r = 0; r <= rmax; ++r) {
      for (int dx = -r; dx <= r; ++dx) {
        const int x = cx + dx;
        if (x < 0 || x >= g.nx) continue;
        for (int dy = -r; dy <= r; ++dy) {
          const int y = cy + dy;
          if (y < 0 || y >= g.ny) continue;
          const int row = (x * g.ny + y) * g.nz;
          if (dx == -r || dx == r || dy == -r || dy == r) {  // a whole column of the shell: contiguous cells
            const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
            if (z0 <= z1) scan_run(start[row + z0], start[row + z1 + 1]);
          } else {  // only the two caps
            const int za = cz - r, zb = cz + r;
            if (za >= 0 && za < g.nz) scan_run(start[row + za], start[row + za + 1]);
            if (zb >= 0 && zb < g.nz && zb != za) scan_run(start[row + zb], start[row + zb + 1]);
          }
        }
      }
      // every point closer than r * h has been seen (the query lies inside its cell)
      const float cover = __fmul_rn((float)r, g.h);
      if (cnt == k && bd[k - 1] <= __fmul_rn(cover, cover)) break;
    }
    // covariance of the neighbours (two-pass, double) -> normal, curvature
    double m[3] = {0, 0, 0};
    for (int j = 0; j < cnt; ++j) {
      const float4 p = sorted[bs[j]];
      m[0] += p.x;
      m[1] += p.y;
      m[2] += p.z;
    }
    const double inv = cnt > 0 ? 1.0 / cnt : 0.0;
    m[0] *= inv;
    m[1] *= inv;
    m[2] *= inv;
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int j = 0; j < cnt; ++j) {
      const float4 p = sorted[bs[j]];
      const double d[3] = {p.x - m[0], p.y - m[1], p.z - m[2]};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) C[a][b] += d[a] * d[b];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) C[a][b] *= inv;
    double lmin, trace, nr[3];
    jacobi3_min(C, lmin, trace, nr);
    float4 out;
    if (cnt < 3 || !(trace > 0.0)) {  // PCL: fewer than 3 neighbours / degenerate -> NaN normal
      const float qn = __int_as_float(0x7fc00000);
      out = make_float4(qn, qn, qn, qn);
    } else {
      // flipNormalTowardsViewpoint(point, 0, 0, 0, n): the normal looks at the sensor
      const double dot = -(q.x * nr[0] + q.y * nr[1] + q.z * nr[2]);
      const double sg = dot < 0 ? -1.0 : 1.0;
      out = make_float4((float)(sg * nr[0]), (float)(sg * nr[1]), (float)(sg * nr[2]), (float)fabs(lmin / trace));
    }
    normal_curv[qi] = out;
    if (nbr)
      for (int j = 0; j < k; ++j) nbr[(size_t)qi * k + j] = j < cnt ? bi[j] : -1;
  }
}

// ssc.cpp:101-105 and :137-151
__global__ void __launch_bounds__(256) k_calibrate_intensity(float4* __restrict__ pts, const float4* __restrict__ normal_curv, int n, float max_intensity) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    if (p.w > max_intensity) p.w = max_intensity;
    const float4 nc = normal_curv[i];
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(nc.x, p.x), __fmul_rn(nc.y, p.y)), __fmul_rn(nc.z, p.z));
    const float nn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(nc.x, nc.x), __fmul_rn(nc.y, nc.y)), __fmul_rn(nc.z, nc.z)));
    const float pn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z)));
    float angleCos = fabsf(__fdiv_rn(dot, __fmul_rn(nn, pn)));
    if (angleCos < 0.3f) angleCos = 0.3f;  // NaN (no normal) compares false: the division below then yields NaN, as in the reference
    const float v = __fdiv_rn(p.w, angleCos);
    p.w = (v > max_intensity) ? max_intensity : v;
    pts[i] = p;
  }
}

int knn_normals_dev(scvod_ctx* c, const float4* pts_dev, int n, int k, float4* normal_curv_dev, int32_t* nbr_dev) {
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  if (n <= 0) return SCVOD_OK;
  // cell edge from the mean density of the occupied volume would need a pass; a fixed 0.25 m edge suits LiDAR clouds (the ring
  // search adapts to sparse regions), doubled by build_grid while the grid would exceed its cell budget
  BuiltGrid bg;
  int rc = build_grid(c, pts_dev, n, 0.25f, bg);
  if (rc) return rc;
  {
    void* stream_ = (void*)st;
    TIMED("k_knn_normals", TSTREAM);
    k_knn_normals<<<std::max(1, std::min((n + 127) / 128, num_sms() * 16)), 128, 0, st>>>(bg.sorted.as<float4>(), bg.sorted_idx.as<int>(), bg.start.as<int>(), bg.g, n,
                                                                                        k, normal_curv_dev, nbr_dev);
  }
  ctx_add_launches(c, 1);
  ECU(cudaGetLastError());
  ECU(cudaStreamSynchronize(st));  // the grid goes out of scope
  return SCVOD_OK;
}

}  // namespace
}  // namespace scvod

using namespace scvod;

extern "C" int scvod_knn_normals(scvod_ctx* c, const float* xyzi, int n, int k, float* normals3, float* curvature, int32_t* neighbors) {
  if (!c || n < 0 || (n > 0 && !xyzi) || k < 1 || k > kMaxK) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_knn_normals (1 <= k <= 16)");
  if (n == 0) return SCVOD_OK;
  if (cudaSetDevice(ctx_device(c)) != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  DTmp d_p, d_n, d_nb;
  ECU(d_p.alloc(sizeof(float4) * (size_t)n));
  ECU(d_n.alloc(sizeof(float4) * (size_t)n));
  if (neighbors) ECU(d_nb.alloc(sizeof(int32_t) * (size_t)n * k));
  ECU(cudaMemcpyAsync(d_p.p, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  int rc = knn_normals_dev(c, d_p.as<float4>(), n, k, d_n.as<float4>(), neighbors ? d_nb.as<int32_t>() : nullptr);
  if (rc) return rc;
  std::vector<float> nc((size_t)n * 4);
  ECU(cudaMemcpyAsync(nc.data(), d_n.p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (neighbors) ECU(cudaMemcpyAsync(neighbors, d_nb.p, sizeof(int32_t) * (size_t)n * k, cudaMemcpyDeviceToHost, st));
  ECU(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i) {
    if (normals3) {
      normals3[3 * i] = nc[4 * (size_t)i];
      normals3[3 * i + 1] = nc[4 * (size_t)i + 1];
      normals3[3 * i + 2] = nc[4 * (size_t)i + 2];
    }
    if (curvature) curvature[i] = nc[4 * (size_t)i + 3];
  }
  return SCVOD_OK;
}

extern "C" int scvod_calibrate_intensity(scvod_ctx* c, float* xyzi, int n, int search_num, float max_intensity) {
  if (!c || n < 0 || (n > 0 && !xyzi) || search_num < 1 || search_num > kMaxK) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_calibrate_intensity");
  if (n == 0) return SCVOD_OK;
  if (cudaSetDevice(ctx_device(c)) != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  DTmp d_p, d_n;
  ECU(d_p.alloc(sizeof(float4) * (size_t)n));
  ECU(d_n.alloc(sizeof(float4) * (size_t)n));
  ECU(cudaMemcpyAsync(d_p.p, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  int rc = knn_normals_dev(c, d_p.as<float4>(), n, search_num, d_n.as<float4>(), nullptr);
  if (rc) return rc;
  {
    void* stream_ = (void*)st;
    TIMED("k_calibrate_intensity", TSTREAM);
    k_calibrate_intensity<<<std::max(1, std::min((n + 255) / 256, num_sms() * 8)), 256, 0, st>>>(d_p.as<float4>(), d_n.as<float4>(), n, max_intensity);
  }
  ctx_add_launches(c, 1);
  ECU(cudaGetLastError());
  ECU(cudaMemcpyAsync(xyzi, d_p.p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, st));
  ECU(cudaStreamSynchronize(st));
  return SCVOD_OK;
}

// pcl::RegionGrowing::extract with the parameters of SSC::regionGrowing (ssc.cpp:797-832); segment_of[i] = segment of point i in
// growth order (every point gets one), planar_points = points in segments of >= 20 points
extern "C" int scvod_region_growing(scvod_ctx* c, const float* xyzi, int n, int32_t* is_building, int32_t* segment_of, int32_t* planar_points) {
  if (!c || n < 0 || (n > 0 && !xyzi) || !is_building) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_region_growing");
  *is_building = 0;
  if (planar_points) *planar_points = 0;
  if (n == 0) return SCVOD_OK;  // 0 >= 0 * 0.2 is true in the reference, but recognize never passes an empty cluster
  const int k = 10;
  std::vector<float> nrm((size_t)n * 3), curv((size_t)n);
  std::vector<int32_t> nbr((size_t)n * k);
  int rc = scvod_knn_normals(c, xyzi, n, k, nrm.data(), curv.data(), nbr.data());
  if (rc) return rc;
  const float cosine_threshold = cosf((float)(10.0 / 180.0 * M_PI));  // setSmoothnessThreshold
  const float curvature_threshold = 1.2f;
  std::vector<std::pair<float, int>> residual((size_t)n);
  for (int i = 0; i < n; ++i) residual[i] = std::make_pair(curv[i], i);
  std::sort(residual.begin(), residual.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });  // comparePair
  std::vector<int> label((size_t)n, -1), seg_size;
  int seed_counter = 0, seed = residual[0].second, segmented = 0, nseg = 0;
  while (segmented < n) {
    // growRegion
    std::queue<int> seeds;
    seeds.push(seed);
    label[seed] = nseg;
    int in_seg = 1;
    while (!seeds.empty()) {
      const int cur = seeds.front();
      seeds.pop();
      for (int j = 0; j < k; ++j) {
        const int idx = nbr[(size_t)cur * k + j];
        if (idx < 0) break;
        if (label[idx] != -1) continue;
        // validatePoint, smooth mode: angle between the normals of the current seed and the neighbour
        const float dot = std::fabs(nrm[3 * (size_t)cur] * nrm[3 * (size_t)idx] + nrm[3 * (size_t)cur + 1] * nrm[3 * (size_t)idx + 1] +
                                    nrm[3 * (size_t)cur + 2] * nrm[3 * (size_t)idx + 2]);
        if (dot < cosine_threshold) continue;  // also rejects NaN normals? no: NaN < x is false -> accepted, as in PCL
        label[idx] = nseg;
        ++in_seg;
        if (!(curv[idx] > curvature_threshold)) seeds.push(idx);
      }
    }
    segmented += in_seg;
    seg_size.push_back(in_seg);
    ++nseg;
    for (int i_seed = seed_counter + 1; i_seed < n; ++i_seed) {
      const int idx = residual[i_seed].second;
      if (label[idx] == -1) {
        seed = idx;
        seed_counter = i_seed;
        break;
      }
    }
  }
  int plane_pts = 0;
  for (int sz : seg_size)
    if (sz >= 20 && sz <= 1000000) plane_pts += sz;
  if (segment_of) std::memcpy(segment_of, label.data(), sizeof(int32_t) * (size_t)n);
  if (planar_points) *planar_points = plane_pts;
  *is_building = ((double)plane_pts >= (double)n * 0.2) ? 1 : 0;  // plane_pts.size() >= points.size() * 0.2 (:825)
  return SCVOD_OK;
}
