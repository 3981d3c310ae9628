// scvod_eval.cu — the reference's two quality measures on the device, over a uniform search grid:
//
//  * static-map preservation / rejection (tool/analysis.py:124-194, the ERASOR metric the reference reports in doc/note.txt):
//    every ground-truth map point looks up its nearest neighbour in the estimated static map (sklearn NearestNeighbors, k = 1);
//    it is "preserved" when that neighbour is closer than voxelsize * sqrt(3) / 2 (:132,141), statically preserved when both
//    labels are static, dynamically preserved when both are dynamic (:146-151);
//    PR = static preserved / gt static, RR = (gt dynamic - dynamic preserved) / gt dynamic, F1 of the two (:186-188).
//  * TP / FN / TN / FN colouring (src/evaluate.cpp:79-145): a point predicted static is TP when a ground-truth static point lies
//    within 0.15 m, else FN(orange) when a ground-truth dynamic point lies within 0.1 m; a point predicted dynamic is TN when a
//    ground-truth dynamic point lies within 0.15 m, else FN(pink) when a ground-truth static point lies within 0.1 m.
//
// Only neighbours inside the threshold matter to either measure, so a grid whose cells are at least as large as the threshold and a
// 27-cell search is exact: a nearest neighbour beyond the threshold and no neighbour at all give the same answer.  Distances: the
// preservation test squares in double like sklearn's kd-tree on float64 data; the radius tests compare float squared distances
// like FLANN's L2_Simple (dist < radius^2, [recollection] of flann/util/result_set.h RadiusResultSet::addPoint).
// Ties between equidistant neighbours go to the smaller index (kd-tree tie order is implementation defined).
#include <cmath>
#include <cstring>

#include "scvod_kernel_common.cuh"

namespace scvod {

namespace {

#include "scvod_grid.cuh"

struct DynSet {
  int n;
  uint32_t cls[16];
};
__device__ __forceinline__ bool is_dynamic(const DynSet& d, uint32_t sem) {
  bool f = false;
#pragma unroll 4
  for (int k = 0; k < d.n; ++k) f = f || (d.cls[k] == sem);
  return f;
}

// nearest neighbour (within thr) of every ground-truth point in the estimate's grid; counters: {preserved, static preserved,
// dynamic preserved, gt dynamic, [8..24) gt per class}
__global__ void __launch_bounds__(256) k_eval_preservation(const float4* __restrict__ gt, long long n_gt, EGrid g, const float4* __restrict__ est_sorted,
                                                           const int* __restrict__ est_idx, const int* __restrict__ start, long long n_est, double thr2,
                                                           DynSet dyn, unsigned long long* __restrict__ counters, int* __restrict__ nn_out) {
  unsigned long long pres = 0, spres = 0, dpres = 0, gdyn = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_gt; i += (long long)gridDim.x * blockDim.x) {
    const float4 q = __ldg(&gt[i]);
    const uint32_t gsem = ((uint32_t)q.w) & 0xFFFFu;
    const bool gd = is_dynamic(dyn, gsem);
    if (gd) {
      ++gdyn;
      for (int k = 0; k < dyn.n; ++k)
        if (dyn.cls[k] == gsem) atomicAdd(&counters[8 + k], 1ULL);
    }
    int best = -1;
    double best_d2 = thr2;
    float best_w = 0.f;
    if (n_est > 0) {
      int cx, cy, cz;
      cell_coords(g, q.x, q.y, q.z, cx, cy, cz);
      for (int x = cx - 1; x <= cx + 1; ++x) {
        if (x < 0 || x >= g.nx) continue;
        for (int y = cy - 1; y <= cy + 1; ++y) {
          if (y < 0 || y >= g.ny) continue;
          const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.nz - 1);
          if (z0 > z1) continue;
          const int lo = start[(x * g.ny + y) * g.nz + z0], hi = start[(x * g.ny + y) * g.nz + z1 + 1];
          for (int t = lo; t < hi; ++t) {
            const float4 p = __ldg(&est_sorted[t]);
            const double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
            const double d2 = dx * dx + dy * dy + dz * dz;
            const int id = est_idx[t];
            if (d2 < best_d2 || (d2 == best_d2 && best >= 0 && id < best)) {
              best_d2 = d2;
              best = id;
              best_w = p.w;
            }
          }
        }
      }
    }
    if (nn_out) nn_out[i] = best;
    if (best >= 0) {  // sqrt(d2) < thr  <=>  d2 < thr^2 (both non-negative doubles)
      ++pres;
      const bool ed = is_dynamic(dyn, ((uint32_t)best_w) & 0xFFFFu);
      if (!gd && !ed) ++spres;
      if (gd && ed) ++dpres;
    }
  }
  // warp reduction, one atomic per warp and counter
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pres += __shfl_xor_sync(0xffffffffu, pres, o);
    spres += __shfl_xor_sync(0xffffffffu, spres, o);
    dpres += __shfl_xor_sync(0xffffffffu, dpres, o);
    gdyn += __shfl_xor_sync(0xffffffffu, gdyn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (pres) atomicAdd(&counters[0], pres);
    if (spres) atomicAdd(&counters[1], spres);
    if (dpres) atomicAdd(&counters[2], dpres);
    if (gdyn) atomicAdd(&counters[3], gdyn);
  }
}

// dynamic-class population of a cloud: counters[0] = dynamic points, counters[8 + k] = points of class k
__global__ void __launch_bounds__(256) k_eval_class_count(const float4* __restrict__ pts, long long n, DynSet dyn, unsigned long long* __restrict__ counters) {
  unsigned long long nd = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t sem = ((uint32_t)__ldg(&pts[i]).w) & 0xFFFFu;
    for (int k = 0; k < dyn.n; ++k)
      if (dyn.cls[k] == sem) {
        ++nd;
        atomicAdd(&counters[8 + k], 1ULL);
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, o);
  if ((threadIdx.x & 31) == 0 && nd) atomicAdd(&counters[0], nd);
}

__device__ __forceinline__ bool any_within(const EGrid& g, const float4* __restrict__ sorted, const int* __restrict__ start, long long n, float4 q, float r2) {
  if (n <= 0) return false;
  int cx, cy, cz;
  cell_coords(g, q.x, q.y, q.z, cx, cy, cz);
  for (int x = cx - 1; x <= cx + 1; ++x) {
    if (x < 0 || x >= g.nx) continue;
    for (int y = cy - 1; y <= cy + 1; ++y) {
      if (y < 0 || y >= g.ny) continue;
      const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.nz - 1);
      if (z0 > z1) continue;
      const int lo = start[(x * g.ny + y) * g.nz + z0], hi = start[(x * g.ny + y) * g.nz + z1 + 1];
      for (int t = lo; t < hi; ++t) {
        const float4 p = __ldg(&sorted[t]);
        const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < r2) return true;
      }
    }
  }
  return false;
}

// src/evaluate.cpp:79-145; classes: 0 TP, 1 FN (predicted static, a dynamic point within r_miss), 2 TN, 3 FN (predicted dynamic, a
// static point within r_miss), 4 not shown
__global__ void __launch_bounds__(256) k_eval_confusion(const float4* __restrict__ pred, long long n, EGrid gs, const float4* __restrict__ s_sorted,
                                                        const int* __restrict__ s_start, long long ns, EGrid gd, const float4* __restrict__ d_sorted,
                                                        const int* __restrict__ d_start, long long nd, float r_hit2, float r_miss2,
                                                        unsigned long long* __restrict__ counters, uint8_t* __restrict__ per_point) {
  unsigned long long c[5] = {0, 0, 0, 0, 0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 q = __ldg(&pred[i]);
    int k = 4;
    if (q.w != 0.f) {  // predicted static (ori.g != 0)
      if (any_within(gs, s_sorted, s_start, ns, q, r_hit2)) k = 0;
      else if (any_within(gd, d_sorted, d_start, nd, q, r_miss2)) k = 1;
    } else {
      if (any_within(gd, d_sorted, d_start, nd, q, r_hit2)) k = 2;
      else if (any_within(gs, s_sorted, s_start, ns, q, r_miss2)) k = 3;
    }
    ++c[k];
    if (per_point) per_point[i] = (uint8_t)k;
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c[j] += __shfl_xor_sync(0xffffffffu, c[j], o);
    if ((threadIdx.x & 31) == 0 && c[j]) atomicAdd(&counters[j], c[j]);
  }
}

int upload(DTmp& d, const float* host, long long n, cudaStream_t st) {
  ECU(d.alloc(sizeof(float) * 4 * (size_t)std::max<long long>(n, 1)));
  if (n > 0) ECU(cudaMemcpyAsync(d.p, host, sizeof(float) * 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  return SCVOD_OK;
}

}  // namespace
}  // namespace scvod

using namespace scvod;

extern "C" int scvod_evaluate_map(scvod_ctx* c, const float* gt_xyzl, int64_t n_gt, const float* est_xyzl, int64_t n_est, float voxelsize,
                                  const int32_t* dynamic_classes, int n_classes, scvod_eval_result* out, int32_t* nn_index) {
  if (!c || !out || n_gt < 0 || n_est < 0 || (n_gt > 0 && !gt_xyzl) || (n_est > 0 && !est_xyzl) || !(voxelsize > 0.f) || n_classes < 0 || n_classes > 8 ||
      (n_classes > 0 && !dynamic_classes) || n_gt > 0x7fffffffLL || n_est > 0x7fffffffLL)
    return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_evaluate_map");
  if (cudaSetDevice(ctx_device(c)) != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  DynSet dyn;
  dyn.n = n_classes;
  for (int k = 0; k < 16; ++k) dyn.cls[k] = k < n_classes ? (uint32_t)dynamic_classes[k] : 0xffffffffu;
  DTmp d_gt, d_est, d_cnt, d_nn;
  int rc = upload(d_gt, gt_xyzl, n_gt, st);
  if (rc) return rc;
  rc = upload(d_est, est_xyzl, n_est, st);
  if (rc) return rc;
  ECU(d_cnt.alloc(sizeof(unsigned long long) * 48));
  ECU(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long) * 48, st));
  if (nn_index) ECU(d_nn.alloc(sizeof(int) * (size_t)std::max<int64_t>(n_gt, 1)));
  const double thr = (double)voxelsize * std::sqrt(3.0) / 2.0;  // DETERMINISTIC_INLIER_THR (tool/analysis.py:132)
  BuiltGrid bg;
  rc = build_grid(c, d_est.as<float4>(), n_est, (float)thr * 1.0001f, bg);
  if (rc) return rc;
  unsigned long long* cg = d_cnt.as<unsigned long long>();
  if (n_gt > 0) {
    void* stream_ = (void*)st;
    TIMED("k_eval_preservation", TSTREAM);
    k_eval_preservation<<<grid_blocks(n_gt), 256, 0, st>>>(d_gt.as<float4>(), n_gt, bg.g, bg.sorted.as<float4>(), bg.sorted_idx.as<int>(), bg.start.as<int>(), n_est,
                                                          thr * thr, dyn, cg, nn_index ? d_nn.as<int>() : nullptr);
    ctx_add_launches(c, 1);
  }
  if (n_est > 0) {
    void* stream_ = (void*)st;
    TIMED("k_eval_class_count", TSTREAM);
    k_eval_class_count<<<grid_blocks(n_est), 256, 0, st>>>(d_est.as<float4>(), n_est, dyn, cg + 24);
    ctx_add_launches(c, 1);
  }
  ECU(cudaGetLastError());
  unsigned long long h[48];
  ECU(cudaMemcpyAsync(h, d_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  if (nn_index && n_gt > 0) ECU(cudaMemcpyAsync(nn_index, d_nn.p, sizeof(int) * (size_t)n_gt, cudaMemcpyDeviceToHost, st));
  ECU(cudaStreamSynchronize(st));
  std::memset(out, 0, sizeof(*out));
  out->gt_dynamic = (int64_t)h[3];
  out->gt_static = n_gt - out->gt_dynamic;
  out->est_dynamic = (int64_t)h[24];
  out->est_static = n_est - out->est_dynamic;
  out->preserved = (int64_t)h[0];
  out->static_preserved = (int64_t)h[1];
  out->dynamic_preserved = (int64_t)h[2];
  for (int k = 0; k < n_classes; ++k) {
    out->gt_per_class[k] = (int64_t)h[8 + k];
    out->est_per_class[k] = (int64_t)h[32 + k];
  }
  // tool/analysis.py:186-188 (a map without dynamic or without static points divides by zero there; reported as NaN here)
  const double nan = std::nan("");
  out->preservation_rate = out->gt_static > 0 ? (double)out->static_preserved / (double)out->gt_static * 100.0 : nan;
  out->rejection_rate = out->gt_dynamic > 0 ? (double)(out->gt_dynamic - out->dynamic_preserved) / (double)out->gt_dynamic * 100.0 : nan;
  const double pr = out->preservation_rate / 100.0, rr = out->rejection_rate / 100.0;
  out->f1 = (pr + rr) > 0 ? 2.0 * pr * rr / (pr + rr) : nan;
  return SCVOD_OK;
}

extern "C" int scvod_evaluate_confusion(scvod_ctx* c, const float* pred_xyzs, int64_t n, const float* static_gt_xyz, int64_t ns, const float* dynamic_gt_xyz,
                                        int64_t nd, float r_hit, float r_miss, int64_t counts5[5], uint8_t* per_point) {
  if (!c || !counts5 || n < 0 || ns < 0 || nd < 0 || (n > 0 && !pred_xyzs) || (ns > 0 && !static_gt_xyz) || (nd > 0 && !dynamic_gt_xyz) || !(r_hit > 0.f) ||
      !(r_miss > 0.f) || n > 0x7fffffffLL || ns > 0x7fffffffLL || nd > 0x7fffffffLL)
    return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_evaluate_confusion");
  if (cudaSetDevice(ctx_device(c)) != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  DTmp d_p, d_s, d_d, d_cnt, d_pp;
  int rc = upload(d_p, pred_xyzs, n, st);
  if (rc) return rc;
  rc = upload(d_s, static_gt_xyz, ns, st);
  if (rc) return rc;
  rc = upload(d_d, dynamic_gt_xyz, nd, st);
  if (rc) return rc;
  ECU(d_cnt.alloc(sizeof(unsigned long long) * 8));
  ECU(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long) * 8, st));
  if (per_point) ECU(d_pp.alloc((size_t)std::max<int64_t>(n, 1)));
  const float rmax = std::max(r_hit, r_miss) * 1.0001f;
  BuiltGrid gs, gd;
  rc = build_grid(c, d_s.as<float4>(), ns, rmax, gs);
  if (rc) return rc;
  rc = build_grid(c, d_d.as<float4>(), nd, rmax, gd);
  if (rc) return rc;
  if (n > 0) {
    void* stream_ = (void*)st;
    TIMED("k_eval_confusion", TSTREAM);
    k_eval_confusion<<<grid_blocks(n), 256, 0, st>>>(d_p.as<float4>(), n, gs.g, gs.sorted.as<float4>(), gs.start.as<int>(), ns, gd.g, gd.sorted.as<float4>(),
                                                    gd.start.as<int>(), nd, r_hit * r_hit, r_miss * r_miss, d_cnt.as<unsigned long long>(),
                                                    per_point ? d_pp.as<uint8_t>() : nullptr);
    ctx_add_launches(c, 1);
  }
  ECU(cudaGetLastError());
  unsigned long long h[8];
  ECU(cudaMemcpyAsync(h, d_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  if (per_point && n > 0) ECU(cudaMemcpyAsync(per_point, d_pp.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  ECU(cudaStreamSynchronize(st));
  for (int k = 0; k < 5; ++k) counts5[k] = (int64_t)h[k];
  return SCVOD_OK;
}
