// scvod_kernel_common.cuh — what the kernel translation units share: warp / block scans, launch geometry, the per-kernel
// CUDA-event timer macro.  Every .cu of the library is compiled with -fmad=false; index- and threshold-determining float
// expressions also go through explicit _rn intrinsics so that no FMA contraction can change a voxel index or a label.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "scvod_device_math.cuh"
#include "scvod_internal.h"

namespace scvod {

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one int per thread across the block; returns exclusive prefix, total in *total.
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* s_warp /* THREADS/32 + 1 ints */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = (lane < THREADS / 32) ? s_warp[lane] : 0;
    int xi = warp_incl_scan(x);
    if (lane < THREADS / 32) s_warp[lane] = xi - x;
    if (lane == THREADS / 32 - 1) s_warp[THREADS / 32] = xi;
  }
  __syncthreads();
  int res = inc - v + s_warp[w];
  *total = s_warp[THREADS / 32];
  __syncthreads();
  return res;
}

// voxel_idx -> compact voxel id through the occupancy bitmap and its popcount ranks (replaces hash_cloud.find), -1 if empty
__device__ __forceinline__ int vox_lookup(const uint32_t* __restrict__ bm, const int32_t* __restrict__ wr, const GridSpec& g, int vid) {
  int key = vid + g.key_off;
  if (key < 0 || key >= g.key_count) return -1;
  uint32_t w = bm[key >> 5];
  uint32_t bit = 1u << (key & 31);
  if (!(w & bit)) return -1;
  return wr[key >> 5] + __popc(w & (bit - 1));
}

// RAII CUDA-event timer around a launch (LaunchTimer, scvod_timing.cu): active only while scvod_kernel_timing is enabled
#define TIMED(name, st) LaunchTimer timer__(name, (void*)(st))
#define TSTREAM ((cudaStream_t)stream_)

inline BinParams make_bin_params(const HostParams& hp) {
  BinParams bp;
  bp.min_dis = hp.p.min_dis;
  bp.max_dis = hp.p.max_dis;
  bp.min_angle = hp.p.min_angle;
  bp.max_angle = hp.p.max_angle;
  bp.min_azimuth = hp.p.min_azimuth;
  bp.max_azimuth = hp.p.max_azimuth;
  bp.range_res = hp.p.range_res;
  bp.sector_res = hp.p.sector_res;
  bp.azimuth_res = hp.p.azimuth_res;
  bp.range_num = hp.g.range_num;
  bp.sector_num = hp.g.sector_num;
  // dev_bin_filtered: guard band of 1e-3 degrees around every bin edge and gate (>= 5x the provable distance between the
  // approximate and the exact chain), plus the rounding of the bin coordinate itself (a few ulps of its largest magnitude)
  bp.eps_deg = 1.0e-3f;
  bp.inv_sector_res = 1.0f / bp.sector_res;
  bp.inv_azimuth_res = 1.0f / bp.azimuth_res;
  bp.eps_qs = bp.eps_deg * fabsf(bp.inv_sector_res) + 1.0e-6f * (360.0f + fabsf(bp.min_angle)) * fabsf(bp.inv_sector_res);
  bp.eps_qe = bp.eps_deg * fabsf(bp.inv_azimuth_res) + 1.0e-6f * (90.0f + fabsf(bp.min_azimuth)) * fabsf(bp.inv_azimuth_res);
  return bp;
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// grid.x for "per point of a scan" kernels: enough CTAs per scan that nscans * gx covers the GPU a
// few times over, in multiples of the SM count.
inline int grid_x_for(int nscans, int per_scan_items, int threads) {
  int want = (per_scan_items + threads - 1) / threads;
  int cap = (num_sms() * 8 + nscans - 1) / nscans;
  if (cap < 1) cap = 1;
  return want < cap ? (want < 1 ? 1 : want) : cap;
}

}  // namespace scvod
