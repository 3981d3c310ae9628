// scvod_kernels.cu — hand-written sm_100a kernels for the SCV-OD hot path.
//
// Stages (reference file:line each one replaces):
//   ground     PatchWork::estimate_ground            include/patchwork.h:278-504
//   binning    SSC::makeApriVec + Utility polar math  src/ssc.cpp:155-195, include/utility.h:346-392
//   descriptor SSC::makeHashCloud                     src/ssc.cpp:253-289
//   cluster    findVoxelNeighbors / CVC adjacency / intensity-similarity edges  src/ssc.cpp:395-411,299-351,587-595
//   diff       transformCloud + re-binning + next-frame lookup of SSC::tracking  include/utility.h:394-406, src/ssc.cpp:1275-1315
//
// Everything here is HBM/latency-bound integer and float work: no tensor-core path exists for it.
// The file is compiled with -fmad=false; index- and threshold-determining float expressions also go
// through explicit _rn intrinsics so that no FMA contraction can change a voxel index or a label.
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

__device__ __forceinline__ uint32_t float_sort_key(float z) {
  if (z == 0.f) z = 0.f;  // -0 and +0 compare equal in point_z_cmp (patchwork.h:33-35)
  uint32_t u = __float_as_uint(z);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// PatchWork constants (patchwork.h:48-51,83-94,115-129)
__constant__ int c_zone_sectors[4] = {16, 32, 54, 32};
__constant__ int c_zone_rings[4] = {2, 4, 4, 4};
__constant__ int c_zone_base[4] = {0, 32, 160, 376};
__constant__ int c_zone_ring0[4] = {0, 2, 6, 10};  // concentric_idx of the zone's first ring
__constant__ double c_elev_thr[4] = {-1.2, -0.9984, -0.851, -0.605};
__constant__ double c_flat_thr[4] = {0.0, 0.000125, 0.000185, 0.000185};


struct Mat34 {
  float m[12];
};


struct GroundConst {
  double low_thr;   // -1.8 * sensor_height_  (patchwork.h:304)
  double seed_thr;  // adaptive_seed_selection_margin_ * sensor_height_ (patchwork.h:247)
  double min_range, max_range, z2, z3, z4;
  double ring_size[4], sector_size[4], rmin[4];
};

// ------------------------------------------------------------------------------------------------
// G1: per-point patch assignment (pc2czm, patchwork.h:431-459) + per-patch histogram
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_patch_assign(const float4* __restrict__ pts, const int64_t* __restrict__ off,
                                                      GroundConst gc, int16_t* __restrict__ patch_of,
                                                      int32_t* __restrict__ patch_cnt, uint8_t* __restrict__ cls) {
  __shared__ int s_hist[kNumPatches];
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int n = (int)(off[b + 1] - base);
  for (int i = threadIdx.x; i < kNumPatches; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&pts[base + i]);
    int pid;
    if ((double)p.z < gc.low_thr) {
      pid = -1;
    } else {
      double x = (double)p.x, y = (double)p.y;
      double r = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
      if ((r <= gc.max_range) && (r > gc.min_range)) {
        double theta = (y >= 0) ? atan2(y, x) : __dadd_rn(2.0 * 3.14159265358979323846, atan2(y, x));
        int k = (r < gc.z2) ? 0 : (r < gc.z3) ? 1 : (r < gc.z4) ? 2 : 3;
        int ring = min((int)__ddiv_rn(__dsub_rn(r, gc.rmin[k]), gc.ring_size[k]), c_zone_rings[k] - 1);
        int sector = min((int)__ddiv_rn(theta, gc.sector_size[k]), c_zone_sectors[k] - 1);
        pid = c_zone_base[k] + ring * c_zone_sectors[k] + sector;
      } else {
        pid = -2;
      }
    }
    patch_of[base + i] = (int16_t)pid;
    if (pid >= 0)
      atomicAdd(&s_hist[pid], 1);
    else
      cls[base + i] = (pid == -1) ? SCVOD_PT_DROPPED_LOW : SCVOD_PT_DROPPED_RANGE;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumPatches; i += blockDim.x) {
    int c = s_hist[i];
    if (c) atomicAdd(&patch_cnt[b * kNumPatches + i], c);
  }
}

// G2: exclusive scan over the 504 patch counts of a scan; patches above 1024 points go on the worklist of their sort tier
__global__ void __launch_bounds__(512) k_patch_scan(const int32_t* __restrict__ patch_cnt, int32_t* __restrict__ patch_off,
                                                    int32_t* __restrict__ patch_cur, int32_t* __restrict__ sort_ctr /* [3][2] */,
                                                    int32_t* __restrict__ sort_list /* [3][list_cap] */, int list_cap) {
  __shared__ int s_w[17];
  const int b = blockIdx.x;
  int v = (threadIdx.x < kNumPatches) ? patch_cnt[b * kNumPatches + threadIdx.x] : 0;
  int total;
  int ex = block_excl_scan<512>(v, &total, s_w);
  if (threadIdx.x < kNumPatches) {
    patch_off[b * (kNumPatches + 1) + threadIdx.x] = ex;
    patch_cur[b * kNumPatches + threadIdx.x] = 0;
    if (v > 1024) {
      const int tier = (v <= 4096) ? 0 : (v <= 16384) ? 1 : 2;
      const int slot = atomicAdd(&sort_ctr[2 * tier], 1);
      sort_list[(size_t)tier * list_cap + slot] = b * kNumPatches + threadIdx.x;
    }
  }
  if (threadIdx.x == 0) patch_off[b * (kNumPatches + 1) + kNumPatches] = total;
}

// G3: scatter (z key, local index) into the patch buckets.  A CTA owns a contiguous chunk of the scan: it counts its
// points per patch in shared memory, reserves one range per (CTA, patch) with a single global atomic, and hands out
// the slots inside the range with shared-memory atomics (the order inside a bucket is irrelevant: it is sorted next).
__global__ void __launch_bounds__(256) k_patch_scatter(const float4* __restrict__ pts, const int64_t* __restrict__ off,
                                                       const int16_t* __restrict__ patch_of,
                                                       const int32_t* __restrict__ patch_off, int32_t* __restrict__ patch_cur,
                                                       uint64_t* __restrict__ bucket_kv) {
  __shared__ int s_cnt[kNumPatches];
  __shared__ int s_base[kNumPatches];
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int n = (int)(off[b + 1] - base);
  const int chunk = (n + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * chunk, i1 = min(n, i0 + chunk);
  for (int i = threadIdx.x; i < kNumPatches; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int pid = patch_of[base + i];
    if (pid >= 0) atomicAdd(&s_cnt[pid], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumPatches; i += blockDim.x) {
    const int c = s_cnt[i];
    s_base[i] = c ? patch_off[b * (kNumPatches + 1) + i] + atomicAdd(&patch_cur[b * kNumPatches + i], c) : 0;
    s_cnt[i] = 0;
  }
  __syncthreads();
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int pid = patch_of[base + i];
    if (pid < 0) continue;
    const float z = __ldg(&pts[base + i]).z;
    const int slot = s_base[pid] + atomicAdd(&s_cnt[pid], 1);
    bucket_kv[base + slot] = ((uint64_t)float_sort_key(z) << 32) | (uint32_t)i;
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 one-sided... no: two-sided Jacobi SVD, float, in the evaluation order of Eigen 3.3.4's
// JacobiSVD<MatrixXf> for a square input (called at patchwork.h:220).  U columns = left vectors.
// ------------------------------------------------------------------------------------------------
struct Rot2 {
  float c, s;
};

__device__ __forceinline__ Rot2 dev_make_jacobi(float x, float y, float z) {
  Rot2 r;
  float deno = dm(2.f, fabsf(y));
  if (deno < 1.17549435e-38f) {
    r.c = 1.f;
    r.s = 0.f;
  } else {
    float tau = dd(ds(x, z), deno);
    float w = __fsqrt_rn(da(dm(tau, tau), 1.f));
    float t = (tau > 0.f) ? dd(1.f, da(tau, w)) : dd(1.f, ds(tau, w));
    float sign_t = t > 0.f ? 1.f : -1.f;
    float n = dd(1.f, __fsqrt_rn(da(dm(t, t), 1.f)));
    r.s = dm(dm(dm(-sign_t, dd(y, fabsf(y))), fabsf(t)), n);
    r.c = n;
  }
  return r;
}

__device__ void dev_svd3(const float A[3][3], float U[3][3], float sv[3]) {
  const float precision = dm(2.f, 1.1920929e-07f);
  const float tinyf = 1.17549435e-38f;
  float scale = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) scale = fmaxf(scale, fabsf(A[i][j]));
  if (scale == 0.f) scale = 1.f;
  float W[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      W[i][j] = dd(A[i][j], scale);
      U[i][j] = (i == j) ? 1.f : 0.f;
    }
  float maxDiag = fmaxf(fabsf(W[0][0]), fmaxf(fabsf(W[1][1]), fabsf(W[2][2])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 64) {
    finished = true;
#pragma unroll
    for (int p = 1; p < 3; ++p) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q >= p) continue;
        float threshold = fmaxf(tinyf, dm(precision, maxDiag));
        if (fabsf(W[p][q]) > threshold || fabsf(W[q][p]) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd
          float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
          Rot2 rot1;
          float t = da(m00, m11);
          float d = ds(m10, m01);
          if (fabsf(d) < tinyf) {
            rot1.s = 0.f;
            rot1.c = 1.f;
          } else {
            float u = dd(t, d);
            float tmp = __fsqrt_rn(da(1.f, dm(u, u)));
            rot1.s = dd(1.f, tmp);
            rot1.c = dd(u, tmp);
          }
          if (!(rot1.c == 1.f && rot1.s == 0.f)) {
            float a00 = da(dm(rot1.c, m00), dm(rot1.s, m10)), a01 = da(dm(rot1.c, m01), dm(rot1.s, m11));
            float a10 = da(dm(-rot1.s, m00), dm(rot1.c, m10)), a11 = da(dm(-rot1.s, m01), dm(rot1.c, m11));
            m00 = a00;
            m01 = a01;
            m10 = a10;
            m11 = a11;
          }
          Rot2 jr = dev_make_jacobi(m00, m01, m11);
          Rot2 jl;
          jl.c = ds(dm(rot1.c, jr.c), dm(rot1.s, -jr.s));
          jl.s = da(dm(rot1.c, -jr.s), dm(rot1.s, jr.c));
          if (!(jl.c == 1.f && jl.s == 0.f)) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              float xi = W[p][k], yi = W[q][k];
              W[p][k] = da(dm(jl.c, xi), dm(jl.s, yi));
              W[q][k] = da(dm(-jl.s, xi), dm(jl.c, yi));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              float xi = U[k][p], yi = U[k][q];
              U[k][p] = da(dm(jl.c, xi), dm(jl.s, yi));
              U[k][q] = da(dm(-jl.s, xi), dm(jl.c, yi));
            }
          }
          {
            float c = jr.c, s = -jr.s;
            if (!(c == 1.f && s == 0.f)) {
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                float xi = W[k][p], yi = W[k][q];
                W[k][p] = da(dm(c, xi), dm(s, yi));
                W[k][q] = da(dm(-s, xi), dm(c, yi));
              }
            }
          }
          maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[p][p]), fabsf(W[q][q])));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = W[i][i];
    sv[i] = fabsf(a);
    if (a < 0.f) {
#pragma unroll
      for (int k = 0; k < 3; ++k) U[k][i] = -U[k][i];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) sv[i] = dm(sv[i], scale);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    int pos = i;
    float best = sv[i];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (k > i && sv[k] > best) {
        best = sv[k];
        pos = k;
      }
    if (best == 0.f) break;
    if (pos != i) {
      float t = sv[i];
      sv[i] = sv[pos];
      sv[pos] = t;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float u = U[k][i];
        U[k][i] = U[k][pos];
        U[k][pos] = u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// G4: R-GPF of one patch, split in three kernels so that each one maps to what bounds it:
//   G4a k_patch_sort   CTA per (patch, scan): z-sort of the patch in shared memory (8 B/point), points
//                      written back in sorted order                                       (patchwork.h:289-295)
//   G4b k_patch_chain  WARP per (patch, scan): seeds + the three plane fits.  PCL's single-pass float
//                      covariance is order dependent (SURVEY.md hard part 3), so the sums are accumulated
//                      strictly sequentially, one lane per accumulator; the chain is latency bound, hence one
//                      warp per patch, many patches per SM, points streamed through a small cp.async ring
//                      instead of a whole-patch shared-memory tile   (patchwork.h:235-268, 217-232, 463-504)
//   G4c k_patch_rank_* warp (<= 1024 points) or CTA (worklists) per (patch, scan): final ground test, gating outcome, curved-voxel binning of the
//                      nonground points (ssc.cpp:158-172,185-188) and ordered ranks     (patchwork.h:331-384)
// ------------------------------------------------------------------------------------------------
struct FitArgs {
  const float4* pts;
  const int64_t* off;
  const int32_t* patch_cnt;
  const int32_t* patch_off;
  uint64_t* bucket_kv;
  float4* sorted_xyz;  // per bucket slot: the point, in z-sorted order inside its patch
  int32_t* sorted_idx;
  int32_t* slot_pos;
  int32_t* slot_apos;
  int32_t* slot_vid;
  int16_t* slot_patch;
  int32_t* patch_out;  // [scans][504][8]
  float* patch_plane;  // [scans][504][12]: normal, mean, singular values, d, decision, npts
  uint8_t* cls;
  int32_t* err;
  GroundConst gc;
  BinParams bp;
};

constexpr int kPatchOutStride = 8;  // n_ground_out, n_nonground_out, n_apri, n_quirk, nG, nGP, rejected, -

// tiers of k_patch_sort by patch size: <= 1024 points (direct grid, almost every patch, bitonic network in shared memory),
// <= 4096 (radix sort in a double-buffered 64 KB shared-memory tile) and above (radix sort through L2).  The patches
// above 1024 points are put on per-tier worklists by k_patch_scan and sorted by persistent CTAs, instead of launching
// 504 x scans CTAs per tier that mostly exit at once.
constexpr int kSortT0 = 1024, kSortT1 = 4096, kSortT2 = 16384;

template <int THREADS, bool GLOBAL>
__device__ __forceinline__ void sort_one_patch(const FitArgs& a, int p, int b, int n, uint64_t* kv_smem) {
  const int64_t base = a.off[b];
  const int slot0 = a.patch_off[b * (kNumPatches + 1) + p];
  const int tid = threadIdx.x;
  uint64_t* kv = GLOBAL ? (a.bucket_kv + base + slot0) : kv_smem;
  // Bitonic network with a "flip" first step per merge, so every compare-exchange is ascending and the
  // virtual +inf padding above n never moves: indices >= n are simply skipped.  Keys are (z key, index).
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  if (!GLOBAL) {
    for (int j = tid; j < n; j += THREADS) kv[j] = a.bucket_kv[base + slot0 + j];
  }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    const int hk = k >> 1;
    for (int t = tid; t < (np2 >> 1); t += THREADS) {
      int blk = t / hk, j = t - blk * hk;
      int i = blk * k + j, l = blk * k + k - 1 - j;
      if (l < n) {
        uint64_t x = kv[i], y = kv[l];
        if (x > y) {
          kv[i] = y;
          kv[l] = x;
        }
      }
    }
    __syncthreads();
    for (int s2 = hk >> 1; s2 > 0; s2 >>= 1) {
      for (int t = tid; t < (np2 >> 1); t += THREADS) {
        int i = ((t & ~(s2 - 1)) << 1) | (t & (s2 - 1));
        int l = i | s2;
        if (l < n) {
          uint64_t x = kv[i], y = kv[l];
          if (x > y) {
            kv[i] = y;
            kv[l] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < n; j += THREADS) {
    const int idx = (int)(uint32_t)kv[j];
    float4 q = __ldg(&a.pts[base + idx]);
    q.w = 0.f;
    a.sorted_xyz[base + slot0 + j] = q;
    a.sorted_idx[base + slot0 + j] = idx;
    a.slot_patch[base + slot0 + j] = (int16_t)p;
  }
}

// lowest tier: one CTA per (patch, scan)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_patch_sort(FitArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = blockIdx.x, b = blockIdx.y;
  const int n = a.patch_cnt[b * kNumPatches + p];
  if (n > kSortT0) return;
  if (n <= kMinPatchPts) {  // patchwork.h:331: the patch vanishes from both outputs
    const int64_t base = a.off[b];
    const int slot0 = a.patch_off[b * (kNumPatches + 1) + p];
    const int tid = threadIdx.x;
    for (int j = tid; j < n; j += THREADS) {
      int idx = (int)(uint32_t)a.bucket_kv[base + slot0 + j];
      a.cls[base + idx] = SCVOD_PT_DROPPED_SPARSE;
      a.slot_pos[base + slot0 + j] = (3 << 30);
      a.slot_patch[base + slot0 + j] = (int16_t)p;
    }
    if (tid < kPatchOutStride) a.patch_out[(b * kNumPatches + p) * kPatchOutStride + tid] = 0;
    return;
  }
  sort_one_patch<THREADS, false>(a, p, b, n, reinterpret_cast<uint64_t*>(smem_raw));
}

// Upper tiers: persistent CTAs pull (scan, patch) items from the tier's worklist and sort them with a stable LSD radix
// sort on the 32-bit z key (4 passes of 8 bits; 16 B of shared-memory or L2 traffic per element and pass instead of the
// log^2 n passes of the bitonic network).  Every warp owns a contiguous slice: pass = count digits per (warp, digit),
// prefix over (digit, warp), then scatter with match_any ranks, which keeps equal digits in slice order.
// Equal z keys would keep the (arbitrary) bucket order, so a patch that contains a tie is re-sorted by the bitonic
// network on the full (z key, index) pair: the result is always the order of that pair.
template <int THREADS>
__device__ __forceinline__ void radix_sort_kv(uint64_t* bufA, uint64_t* bufB, int n, int* s_cnt /* [THREADS/32][256] */,
                                              int* s_scan /* THREADS/32 + 1 */) {
  constexpr int W = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slice = (((n + W - 1) / W) + 31) & ~31;
  const int j_lo = min(n, wid * slice), j_hi = min(n, j_lo + slice);
  uint64_t* src = bufA;
  uint64_t* dst = bufB;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 32 + 8 * pass;
    for (int i = tid; i < W * 256; i += THREADS) s_cnt[i] = 0;
    __syncthreads();
    for (int j = j_lo + lane; j < j_hi; j += 32) atomicAdd(&s_cnt[wid * 256 + (int)((src[j] >> shift) & 255u)], 1);
    __syncthreads();
    // exclusive prefix in (digit, warp) order: THREADS >= 256, thread d < 256 owns digit d
    int tot = 0;
    if (tid < 256) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const int c = s_cnt[w * 256 + tid];
        s_cnt[w * 256 + tid] = tot;
        tot += c;
      }
    }
    int all;
    const int ex = block_excl_scan<THREADS>(tid < 256 ? tot : 0, &all, s_scan);
    if (tid < 256) {
#pragma unroll
      for (int w = 0; w < W; ++w) s_cnt[w * 256 + tid] += ex;
    }
    __syncthreads();
    for (int j0 = j_lo; j0 < j_hi; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = j < j_hi;
      const uint64_t kv = valid ? src[j] : 0ull;
      const int d = valid ? (int)((kv >> shift) & 255u) : 256 + lane;
      const unsigned same = __match_any_sync(0xffffffffu, d);
      const int rank = __popc(same & ((1u << lane) - 1u));
      int basepos = 0;
      if (valid) basepos = s_cnt[wid * 256 + d];
      __syncwarp();
      if (valid) {
        dst[basepos + rank] = kv;
        if (rank == 0) s_cnt[wid * 256 + d] = basepos + __popc(same);
      }
      __syncwarp();
    }
    __syncthreads();
    uint64_t* t = src;
    src = dst;
    dst = t;
  }
  // four passes: the sorted data is back in bufA
}

template <int THREADS, bool GLOBAL>
__device__ __forceinline__ void radix_sort_one_patch(const FitArgs& a, int p, int b, int n, unsigned char* smem_raw, int smem_elems) {
  __shared__ int s_cnt[(THREADS / 32) * 256];
  __shared__ int s_scan[THREADS / 32 + 1];
  const int64_t base = a.off[b];
  const int slot0 = a.patch_off[b * (kNumPatches + 1) + p];
  const int tid = threadIdx.x;
  uint64_t *bufA, *bufB;
  if (GLOBAL) {
    bufA = a.bucket_kv + base + slot0;
    bufB = reinterpret_cast<uint64_t*>(a.sorted_xyz + base + slot0);  // scratch until the sorted points are written below
  } else {
    bufA = reinterpret_cast<uint64_t*>(smem_raw);
    bufB = bufA + smem_elems;
    for (int j = tid; j < n; j += THREADS) bufA[j] = a.bucket_kv[base + slot0 + j];
  }
  __syncthreads();
  radix_sort_kv<THREADS>(bufA, bufB, n, s_cnt, s_scan);
  int tie = 0;
  for (int j = tid; j + 1 < n; j += THREADS) tie |= ((bufA[j] >> 32) == (bufA[j + 1] >> 32)) ? 1 : 0;
  if (__syncthreads_or(tie)) {  // rare: equal z inside the patch -> order by (z key, index) with the bitonic network
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = 2; k <= np2; k <<= 1) {
      const int hk = k >> 1;
      for (int t = tid; t < (np2 >> 1); t += THREADS) {
        int blk = t / hk, j = t - blk * hk;
        int i = blk * k + j, l = blk * k + k - 1 - j;
        if (l < n) {
          uint64_t x = bufA[i], y = bufA[l];
          if (x > y) {
            bufA[i] = y;
            bufA[l] = x;
          }
        }
      }
      __syncthreads();
      for (int s2 = hk >> 1; s2 > 0; s2 >>= 1) {
        for (int t = tid; t < (np2 >> 1); t += THREADS) {
          int i = ((t & ~(s2 - 1)) << 1) | (t & (s2 - 1));
          int l = i | s2;
          if (l < n) {
            uint64_t x = bufA[i], y = bufA[l];
            if (x > y) {
              bufA[i] = y;
              bufA[l] = x;
            }
          }
        }
        __syncthreads();
      }
    }
  }
  // GLOBAL: the scratch half (bufB) aliases sorted_xyz; after four passes the data is in bufA, so it is free to be written
  for (int j = tid; j < n; j += THREADS) {
    const int idx = (int)(uint32_t)bufA[j];
    float4 q = __ldg(&a.pts[base + idx]);
    q.w = 0.f;
    a.sorted_xyz[base + slot0 + j] = q;
    a.sorted_idx[base + slot0 + j] = idx;
    a.slot_patch[base + slot0 + j] = (int16_t)p;
  }
}

template <int THREADS, bool GLOBAL>
__global__ void __launch_bounds__(THREADS) k_patch_sort_list(FitArgs a, const int32_t* __restrict__ list, int32_t* __restrict__ ctr /* [0] count, [1] cursor */,
                                                             int smem_elems) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_item;
  const int count = ctr[0];
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(&ctr[1], 1);
    __syncthreads();
    const int it = s_item;
    __syncthreads();
    if (it >= count) break;
    const int g = list[it];
    const int b = g / kNumPatches, p = g - b * kNumPatches;
    radix_sort_one_patch<THREADS, GLOBAL>(a, p, b, a.patch_cnt[g], smem_raw, smem_elems);
    __syncthreads();
  }
}

// plane of one R-GPF iteration from the nine sequential sums (estimate_plane_, patchwork.h:217-232) and the gating
// decision of the patch (patchwork.h:339-384): shared by the warp-per-patch and the thread-per-patch chain kernels
struct PlaneState {
  float n0 = 0.f, n1 = 0.f, n2 = 0.f, th = 0.f;
  float meanx = 0.f, meany = 0.f, meanz = 0.f, sv0 = 0.f, sv1 = 0.f, sv2 = 0.f, d = 0.f;
};

__device__ __forceinline__ void solve_plane(float accu[9], int cnt, PlaneState& pl) {
  const float fn = (float)cnt;
#pragma unroll
  for (int k = 0; k < 9; ++k) accu[k] = dd(accu[k], fn);
  float C[3][3];
  C[0][0] = ds(accu[0], dm(accu[6], accu[6]));
  C[0][1] = ds(accu[1], dm(accu[6], accu[7]));
  C[0][2] = ds(accu[2], dm(accu[6], accu[8]));
  C[1][1] = ds(accu[3], dm(accu[7], accu[7]));
  C[1][2] = ds(accu[4], dm(accu[7], accu[8]));
  C[2][2] = ds(accu[5], dm(accu[8], accu[8]));
  C[1][0] = C[0][1];
  C[2][0] = C[0][2];
  C[2][1] = C[1][2];
  float U[3][3], sv[3];
  dev_svd3(C, U, sv);
  pl.n0 = U[0][2];
  pl.n1 = U[1][2];
  pl.n2 = U[2][2];
  // d_ = -(normal^T * mean): Eigen 3-term unrolled redux a0 + (a1 + a2)
  pl.d = -da(dm(pl.n0, accu[6]), da(dm(pl.n1, accu[7]), dm(pl.n2, accu[8])));
  pl.th = (float)__dsub_rn(0.1, (double)pl.d);  // th_dist_d_ = th_dist_ - d_
  pl.meanx = accu[6];
  pl.meany = accu[7];
  pl.meanz = accu[8];
  pl.sv0 = sv[0];
  pl.sv1 = sv[1];
  pl.sv2 = sv[2];
}

__device__ __forceinline__ void write_gating(const FitArgs& a, int p, int b, int n, int zone, const PlaneState& pl) {
  const double ground_z_vec = (double)fabsf(pl.n2);
  const double ground_z_elevation = (double)pl.meanz;
  const float minsv = fminf(pl.sv0, fminf(pl.sv1, pl.sv2));
  const double surface_variable = (double)dd(minsv, da(da(pl.sv0, pl.sv1), pl.sv2));
  const int ring_i = (p - c_zone_base[zone]) / c_zone_sectors[zone];
  const int concentric_idx = c_zone_ring0[zone] + ring_i;
  int decision = 0;
  if (ground_z_vec < 0.707) {
    decision = 1;
  } else if (concentric_idx < 4) {
    if (ground_z_elevation > c_elev_thr[ring_i + 2 * zone]) decision = (c_flat_thr[ring_i + 2 * zone] > surface_variable) ? 3 : 2;
  }
  float* rec = a.patch_plane + (size_t)(b * kNumPatches + p) * 12;
  rec[0] = pl.n0;
  rec[1] = pl.n1;
  rec[2] = pl.n2;
  rec[3] = pl.meanx;
  rec[4] = pl.meany;
  rec[5] = pl.meanz;
  rec[6] = pl.sv0;
  rec[7] = pl.sv1;
  rec[8] = pl.sv2;
  rec[9] = pl.d;
  rec[10] = (float)decision;
  rec[11] = (float)n;
}

// smallest float >= d: for a float z, ((double)z < d) == (z < float_at_or_above(d))
__device__ __forceinline__ float float_at_or_above(double d) { return __double2float_ru(d); }

constexpr int kChainWarps = 4;    // warps (= patches) per CTA of k_patch_chain
constexpr int kChainStages = 4;   // ring depth, 32 points per stage
constexpr int kChainCtasPerSm = 3;  // residency cap: the chain is latency bound, so few warps per scheduler keep the
                                    // long (zone-0) patches fast while the many short ones fill the remaining slots

// ---- mbarrier + 1-D bulk async copy (TMA, UBLKCP in SASS): one elected lane moves a whole 512-byte stage ----------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// TMA = false (default): the ring is filled with per-lane 16-byte cp.async (LDGSTS).  TMA = true: one elected lane issues one
// 512-byte bulk copy per stage (cp.async.bulk, UBLKCP) that completes on an mbarrier.  Both were measured on B200 with the
// same results bit for bit; the bulk-copy variant costs more issue slots per stage (elected-lane branch, arrive.expect_tx,
// try_wait loop, one more __syncwarp) in a kernel that is issue bound: 0.31 ms vs 0.25 ms per 64 scans, so it is opt-in
// (scvod_set_option "chain_tma").
template <bool TMA>
__global__ void __launch_bounds__(kChainWarps * 32) k_patch_chain(FitArgs a, int nscans, const int32_t* __restrict__ sort_list,
                                                                  int32_t* __restrict__ sort_ctr, int list_cap) {
  __shared__ __align__(128) float4 s_ring[kChainWarps][kChainStages][32];
  __shared__ float s_prod[kChainWarps][2][32 * 9];
  __shared__ __align__(8) uint64_t s_bar[kChainWarps][kChainStages];  // one "stage has landed" mbarrier per ring slot
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (TMA) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kChainStages; ++i) mbar_init(&s_bar[wid][i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
  }
  unsigned phase = 0;  // bit s: parity of the next completion of ring slot s (warp-uniform)
  // Persistent warps pull patches from one queue, longest chains first: the worklists of the sort tiers (above 16384,
  // above 4096, above 1024 points), then every remaining (patch, scan) in patch-major order (zone 0 = the largest first).
  const int c2 = sort_ctr[4], c1 = sort_ctr[2], c0 = sort_ctr[0];
  const int n_listed = c2 + c1 + c0;
  const int n_items = n_listed + kNumPatches * nscans;
  for (;;) {
  int item = 0;
  if (lane == 0) item = atomicAdd(&sort_ctr[6], 1);
  item = __shfl_sync(0xffffffffu, item, 0);
  if (item >= n_items) break;
  int p, b;
  if (item < n_listed) {
    const int g = (item < c2) ? sort_list[2 * (size_t)list_cap + item] : (item < c2 + c1) ? sort_list[(size_t)list_cap + item - c2] : sort_list[item - c2 - c1];
    b = g / kNumPatches;
    p = g - b * kNumPatches;
  } else {
    const int g = item - n_listed;
    p = g / nscans;
    b = g - p * nscans;
  }
  const int n = a.patch_cnt[b * kNumPatches + p];
  if (n <= kMinPatchPts || (item >= n_listed && n > kSortT0)) continue;
  const int64_t base = a.off[b];
  const int slot0 = a.patch_off[b * (kNumPatches + 1) + p];
  const float4* __restrict__ S = a.sorted_xyz + base + slot0;
  float4(*ring)[32] = s_ring[wid];
  const int zone = (p < 32) ? 0 : (p < 160) ? 1 : (p < 376) ? 2 : 3;

  // ---- seeds (patchwork.h:235-268): sorted ascending => the points below the margin form a prefix ----
  int init_idx = 0;
  if (zone == 0) {
    const float margin = float_at_or_above(a.gc.seed_thr);
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      const bool below = (j < n) && (__ldg(&S[j]).z < margin);
      const unsigned m = __ballot_sync(0xffffffffu, below);
      init_idx += __popc(m);
      if (m != 0xffffffffu) break;
    }
  }
  double lpr;
  {
    const int j = init_idx + lane;
    const float zv = (lane < 20 && j < n) ? __ldg(&S[j]).z : 0.f;
    const int cnt = min(20, n - init_idx);
    double sum = 0;
    for (int i = 0; i < cnt; ++i) sum = __dadd_rn(sum, (double)__shfl_sync(0xffffffffu, zv, i));
    lpr = cnt > 0 ? __ddiv_rn(sum, (double)cnt) : 0.0;
  }
  const float seed_cut = float_at_or_above(__dadd_rn(lpr, 0.3));  // z < lpr + th_seeds_ (patchwork.h:262)

  PlaneState pl;
  float &n0 = pl.n0, &n1 = pl.n1, &n2 = pl.n2, &th = pl.th;
  const int nstages = (n + 31) >> 5;
  // TMA staging: one lane arms the slot's mbarrier with the byte count and issues ONE bulk copy of the whole stage
  // (32 points = 512 B, contiguous in the z-sorted copy); the warp waits on the mbarrier phase before it reads the slot.
  // cp.async staging: every lane copies its own 16 bytes; past the end of the patch the copy degenerates to a zero fill.
  uint64_t* bars = s_bar[wid];
  auto issue = [&](int st) {
    if (TMA) {
      if (st < nstages && lane == 0) {
        const unsigned bytes = 16u * (unsigned)min(32, n - st * 32);
        mbar_expect_tx(&bars[st % kChainStages], bytes);
        bulk_g2s(&ring[st % kChainStages][0], S + st * 32, bytes, &bars[st % kChainStages]);
      }
    } else {
      const int j = st * 32 + lane;
      const bool live = j < n;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&ring[st % kChainStages][lane]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(S + (live ? j : 0)), "r"(live ? 16 : 0));
      asm volatile("cp.async.commit_group;\n" ::);
    }
  };
  auto wait_stage = [&](int st) {  // stage st has landed in its ring slot
    if (TMA) {
      const int slot = st % kChainStages;
      mbar_wait(&bars[slot], (phase >> slot) & 1u);
      phase ^= 1u << slot;
    } else {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(kChainStages - 1));  // one group per stage, kChainStages in flight
    }
  };
  // Lane L < 9 accumulates accu[L] of pcl::computeMeanAndCovarianceMatrix (xx xy xz yy yz zz x y z) STRICTLY in
  // z-sorted order — PCL's single-pass float sums are order dependent.  Per stage of 32 points the work is split:
  //   parallel part   lane t takes point t: ground-set test, its nine terms (a point outside the set, or past the
  //                   end of the patch, contributes -0.0f: an exact identity of float addition), written to
  //                   shared memory as a 32 x 9 tile;
  //   sequential part lane L walks column L of the tile: one LDS + one dependent FADD per point.
  const int col = lane < 9 ? lane : 8;
  for (int it = 0; it < 3; ++it) {
    float acc = 0.f;
    int cnt = 0;
    // Software pipeline inside the warp: the products of stage st+1 are computed (and stored to the other half of
    // s_prod) in the same straight-line block as the column walk of stage st, so their issue slots and shared-memory
    // latencies hide in the 4-cycle bubbles of the dependent FADD chain.  One __syncwarp per stage.
    auto produce = [&](int st) {  // stage st -> s_prod[st & 1]; a stage past the end contributes -0.0f everywhere
      const bool live = st * 32 + lane < n;
      const float4 q = live ? ring[st % kChainStages][lane] : make_float4(0.f, 0.f, 0.f, 0.f);
      bool in;
      if (it == 0) {
        in = live && (q.z < seed_cut);
      } else {
        // result = points * normal_ : (x*n0 + y*n1) + z*n2, three rounded products (patchwork.h:486)
        in = live && (da(da(dm(q.x, n0), dm(q.y, n1)), dm(q.z, n2)) < th);
      }
      cnt += __popc(__ballot_sync(0xffffffffu, in));
      float* pr = s_prod[wid][st & 1] + lane * 9;
      pr[0] = in ? dm(q.x, q.x) : -0.0f;
      pr[1] = in ? dm(q.x, q.y) : -0.0f;
      pr[2] = in ? dm(q.x, q.z) : -0.0f;
      pr[3] = in ? dm(q.y, q.y) : -0.0f;
      pr[4] = in ? dm(q.y, q.z) : -0.0f;
      pr[5] = in ? dm(q.z, q.z) : -0.0f;
      pr[6] = in ? q.x : -0.0f;
      pr[7] = in ? q.y : -0.0f;
      pr[8] = in ? q.z : -0.0f;
    };
    for (int st = 0; st < kChainStages; ++st) issue(st);
    wait_stage(0);
    produce(0);
    if (TMA) __syncwarp();  // every lane has read slot 0 before the async proxy overwrites it (cp.async: a lane refills its own element)
    issue(kChainStages);
    if (!TMA) __syncwarp();
    for (int st = 0; st < nstages; ++st) {
      if (!TMA || st + 1 < nstages) wait_stage(st + 1);  // warp-uniform
      produce(st + 1);
      if (!TMA) issue(st + 1 + kChainStages);
      const float* pc = s_prod[wid][st & 1] + col;
#pragma unroll
      for (int t = 0; t < 32; ++t) acc = da(acc, pc[t * 9]);
      __syncwarp();  // products of stage st + 1 visible; slot (st + 1) % kChainStages has been read by every lane
      if (TMA) issue(st + 1 + kChainStages);
    }
    if (!TMA) asm volatile("cp.async.wait_group 0;\n" ::);
    float accu[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) accu[k] = __shfl_sync(0xffffffffu, acc, k);
    if (cnt == 0) {
      if (lane == 0) atomicOr(a.err, 2);  // cannot happen for finite input; plane kept
    } else {  // every lane evaluates the (tiny) plane solve redundantly: no broadcast, no divergence
      solve_plane(accu, cnt, pl);
    }
  }
  if (lane == 0) write_gating(a, p, b, n, zone, pl);
  __syncwarp();
  }  // queue loop
}

constexpr uint32_t F_G = 1u;      // final ground set
constexpr uint32_t F_PASS = 2u;   // survives the SSC gates
constexpr uint32_t F_QUIRK = 4u;  // some index is -1 (aliasing quirk, SURVEY.md hard part 7)

// slot_pos encoding: role << 30 | in-ground-set << 29 | rank inside its class (ground set / complement);
// slot_apos: rank among the gate-passing points of the same class, or -1.  k_emit turns them into positions with
// the per-patch totals of patch_out (a rejected patch emits [ground set][complement] into cloud_nonground).
// One patch by a group of THREADS threads: a warp (THREADS == 32, the many patches up to 1024 points, no block barrier)
// or a whole CTA (the long patches on the sort worklists).
template <int THREADS>
__device__ __forceinline__ void rank_one_patch(const FitArgs& a, int p, int b, int n, int tid, int* s_scan) {
  const int64_t base = a.off[b];
  const int slot0 = a.patch_off[b * (kNumPatches + 1) + p];
  const float4* __restrict__ S = a.sorted_xyz + base + slot0;
  const float* rec = a.patch_plane + (size_t)(b * kNumPatches + p) * 12;
  const float n0 = rec[0], n1 = rec[1], n2 = rec[2];
  const float th = (float)__dsub_rn(0.1, (double)rec[9]);
  const int decision = (int)rec[10];
  const bool rejected = (decision == 1 || decision == 2);
  int cG = 0, cGP = 0, cNP = 0, cQ = 0;  // running totals (group uniform)
  for (int j0 = 0; j0 < n; j0 += THREADS) {
    const int j = j0 + tid;
    const bool valid = j < n;
    uint32_t f = 0;
    int vid = 0;
    if (valid) {
      const float4 q = __ldg(&S[j]);
      const float res = da(da(dm(q.x, n0), dm(q.y, n1)), dm(q.z, n2));
      if (res < th) f |= F_G;
      if (rejected || !(f & F_G)) {
        BinResult r = dev_bin_point(q.x, q.y, q.z, a.bp);
        if (r.pass) {
          f |= F_PASS;
          if (r.ri < 0 || r.si < 0 || r.ei < 0) f |= F_QUIRK;
        }
        vid = r.vid;
      }
    }
    const bool g = f & F_G, ps = f & F_PASS;
    // chunk counts are <= THREADS <= 512: three 10-bit fields in one scan
    const int packed = (g ? 1 : 0) | ((g && ps) ? (1 << 10) : 0) | ((!g && ps) ? (1 << 20) : 0);
    int total, ex, nq;
    if (THREADS == 32) {
      const int inc = warp_incl_scan(packed);
      total = __shfl_sync(0xffffffffu, inc, 31);
      ex = inc - packed;
      nq = __popc(__ballot_sync(0xffffffffu, (f & F_QUIRK) != 0));
    } else {
      ex = block_excl_scan<(THREADS == 32 ? 64 : THREADS)>(packed, &total, s_scan);
      nq = __syncthreads_count((f & F_QUIRK) ? 1 : 0);
    }
    if (valid) {
      const int rG = cG + (ex & 1023), rGP = cGP + ((ex >> 10) & 1023), rNP = cNP + ((ex >> 20) & 1023);
      const int rN = j - rG;  // complement points before j
      int role, apos = -1;
      if (!rejected && g) {
        role = 0;
      } else {
        role = ps ? 2 : 1;
        if (ps) apos = g ? rGP : rNP;
      }
      a.slot_pos[base + slot0 + j] = (role << 30) | (g ? (1 << 29) : 0) | (g ? rG : rN);
      a.slot_apos[base + slot0 + j] = apos;
      a.slot_vid[base + slot0 + j] = vid;
    }
    cG += total & 1023;
    cGP += (total >> 10) & 1023;
    cNP += (total >> 20) & 1023;
    cQ += nq;
  }
  if (tid == 0) {
    int32_t* pout = a.patch_out + (b * kNumPatches + p) * kPatchOutStride;
    pout[0] = rejected ? 0 : cG;
    pout[1] = rejected ? n : n - cG;
    pout[2] = rejected ? (cGP + cNP) : cNP;
    pout[3] = cQ;
    pout[4] = cG;
    pout[5] = cGP;
    pout[6] = rejected ? 1 : 0;
    pout[7] = 0;
  }
}

constexpr int kRankWarps = 4;  // patches per CTA of k_patch_rank_small

// patches up to 1024 points: one warp each
__global__ void __launch_bounds__(kRankWarps * 32) k_patch_rank_small(FitArgs a) {
  const int p = blockIdx.x * kRankWarps + (threadIdx.x >> 5), b = blockIdx.y;
  if (p >= kNumPatches) return;
  const int n = a.patch_cnt[b * kNumPatches + p];
  if (n <= kMinPatchPts || n > kSortT0) return;
  rank_one_patch<32>(a, p, b, n, threadIdx.x & 31, nullptr);
}

// patches above 1024 points (the three sort worklists): persistent CTAs
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_patch_rank_list(FitArgs a, const int32_t* __restrict__ sort_list, int32_t* __restrict__ sort_ctr,
                                                             int list_cap) {
  __shared__ int s_scan[THREADS / 32 + 1];
  __shared__ int s_item;
  const int c2 = sort_ctr[4], c1 = sort_ctr[2], c0 = sort_ctr[0];
  const int n_listed = c2 + c1 + c0;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(&sort_ctr[7], 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if (item >= n_listed) break;
    const int g = (item < c2) ? sort_list[2 * (size_t)list_cap + item] : (item < c2 + c1) ? sort_list[(size_t)list_cap + item - c2] : sort_list[item - c2 - c1];
    const int b = g / kNumPatches, p = g - b * kNumPatches;
    rank_one_patch<THREADS>(a, p, b, a.patch_cnt[g], threadIdx.x, s_scan);
    __syncthreads();
  }
}

// G5: per-scan exclusive scans of the per-patch output counts (patch-major output order, :327-391)
__global__ void __launch_bounds__(512) k_patch_out_scan(const int32_t* __restrict__ patch_cnt, const int32_t* __restrict__ patch_out,
                                                        int32_t* __restrict__ patch_out_off, int32_t* __restrict__ scan_counts) {
  __shared__ int s_w[17];
  const int b = blockIdx.x;
  const int t = threadIdx.x;
  const bool live = t < kNumPatches && patch_cnt[b * kNumPatches + t] > 0;
  int v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = live ? patch_out[(b * kNumPatches + t) * kPatchOutStride + k] : 0;
  int tot[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int ex = block_excl_scan<512>(v[k], &tot[k], s_w);
    if (k < 3 && t < kNumPatches) patch_out_off[(b * (kNumPatches + 1) + t) * 3 + k] = ex;
  }
  if (t == 0) {
    scan_counts[b * 8 + 0] = tot[0];
    scan_counts[b * 8 + 1] = tot[1];
    scan_counts[b * 8 + 2] = tot[2];
    scan_counts[b * 8 + 4] = tot[3];
  }
}

// G6: emit cloud_out / cloud_nonground order and the apri arrays
__global__ void __launch_bounds__(256) k_emit(const float4* __restrict__ pts, const int64_t* __restrict__ off,
                                              const int32_t* __restrict__ patch_off, const int32_t* __restrict__ patch_out,
                                              const int32_t* __restrict__ patch_out_off, const int32_t* __restrict__ sorted_idx,
                                              const int32_t* __restrict__ slot_pos, const int32_t* __restrict__ slot_apos,
                                              const int32_t* __restrict__ slot_vid, const int16_t* __restrict__ slot_patch,
                                              int32_t* __restrict__ ground_src, int32_t* __restrict__ ng_src,
                                              int32_t* __restrict__ apri_src, int32_t* __restrict__ apri_vid,
                                              float4* __restrict__ apri_xyzi, uint8_t* __restrict__ cls) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int nslots = patch_off[b * (kNumPatches + 1) + kNumPatches];
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nslots; q += gridDim.x * blockDim.x) {
    const int sp = slot_pos[base + q];
    const int role = (sp >> 30) & 3;
    if (role == 3) continue;
    const bool g = (sp >> 29) & 1;
    const int rank = sp & 0x1fffffff;
    const int p = slot_patch[base + q];
    const int idx = sorted_idx[base + q];
    const int32_t* o = patch_out_off + (b * (kNumPatches + 1) + p) * 3;
    const int32_t* po = patch_out + (b * kNumPatches + p) * kPatchOutStride;
    if (role == 0) {
      ground_src[base + o[0] + rank] = idx;
      cls[base + idx] = SCVOD_PT_GROUND;
    } else {
      // cloud_nonground: the complement of the ground set, or [ground set][complement] for a rejected patch
      // (patchwork.h:348-349,373-374)
      const bool rejected = po[6] != 0;
      const int pos = (rejected && !g) ? po[4] + rank : rank;
      ng_src[base + o[1] + pos] = idx;
      if (role == 1) {
        cls[base + idx] = SCVOD_PT_GATED_OUT;
      } else {
        const int ar = slot_apos[base + q];
        const int m = o[2] + ((rejected && !g) ? po[5] + ar : ar);
        apri_src[base + m] = idx;
        apri_vid[base + m] = slot_vid[base + q];
        apri_xyzi[base + m] = __ldg(&pts[base + idx]);
        cls[base + idx] = SCVOD_PT_UNCLUSTERED;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Descriptor stage (SSC::makeHashCloud, ssc.cpp:253-289).  The unordered_map<int,Voxel> becomes an
// occupancy bitmap + popcount rank per scan (324 KB at the KITTI grid), which gives an O(1)
// voxel_idx -> compact id lookup that the neighbour searches and the tracking diff reuse.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vox_mark(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                  const int32_t* __restrict__ apri_vid, GridSpec g, uint32_t* __restrict__ bitmap) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int m_total = scan_counts[b * 8 + 2];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < m_total; m += gridDim.x * blockDim.x) {
    int key = apri_vid[base + m] + g.key_off;
    if (key >= 0 && key < g.key_count) atomicOr(&bitmap[(size_t)b * g.words + (key >> 5)], 1u << (key & 31));
  }
}

__global__ void __launch_bounds__(1024) k_vox_rank(const int64_t* __restrict__ off, GridSpec g, const uint32_t* __restrict__ bitmap,
                                                   int32_t* __restrict__ word_rank, int32_t* __restrict__ vox_vid,
                                                   int32_t* __restrict__ vox_cnt, int32_t* __restrict__ scan_counts) {
  __shared__ int s_w[33];
  const int b = blockIdx.x;
  const int64_t base = off[b];
  const uint32_t* bm = bitmap + (size_t)b * g.words;
  int32_t* wr = word_rank + (size_t)b * g.words;
  const int chunk = (g.words + 1023) / 1024;
  const int w0 = min(g.words, (int)threadIdx.x * chunk), w1 = min(g.words, w0 + chunk);
  int c = 0;
  for (int w = w0; w < w1; ++w) c += __popc(bm[w]);
  int total;
  int ex = block_excl_scan<1024>(c, &total, s_w);
  for (int w = w0; w < w1; ++w) {
    uint32_t bits = bm[w];
    wr[w] = ex;
    while (bits) {
      int bit = __ffs(bits) - 1;
      bits &= bits - 1;
      vox_vid[base + ex] = (w << 5) + bit - g.key_off;
      vox_cnt[base + ex] = 0;
      ++ex;
    }
  }
  if (threadIdx.x == 0) scan_counts[b * 8 + 3] = total;
}

__device__ __forceinline__ int vox_lookup(const uint32_t* __restrict__ bm, const int32_t* __restrict__ wr, const GridSpec& g, int vid) {
  int key = vid + g.key_off;
  if (key < 0 || key >= g.key_count) return -1;
  uint32_t w = bm[key >> 5];
  uint32_t bit = 1u << (key & 31);
  if (!(w & bit)) return -1;
  return wr[key >> 5] + __popc(w & (bit - 1));
}

__global__ void __launch_bounds__(256) k_vox_count(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                   const int32_t* __restrict__ apri_vid, GridSpec g, const uint32_t* __restrict__ bitmap,
                                                   const int32_t* __restrict__ word_rank, int32_t* __restrict__ apri_cid,
                                                   int32_t* __restrict__ vox_cnt) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int m_total = scan_counts[b * 8 + 2];
  const uint32_t* bm = bitmap + (size_t)b * g.words;
  const int32_t* wr = word_rank + (size_t)b * g.words;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < m_total; m += gridDim.x * blockDim.x) {
    int cid = vox_lookup(bm, wr, g, apri_vid[base + m]);
    apri_cid[base + m] = cid;
    if (cid >= 0) atomicAdd(&vox_cnt[base + cid], 1);
  }
}

__global__ void __launch_bounds__(1024) k_vox_offsets(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                      const int32_t* __restrict__ vox_cnt, int32_t* __restrict__ vox_off,
                                                      int32_t* __restrict__ vox_cur) {
  __shared__ int s_w[33];
  const int b = blockIdx.x;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  int carry = 0;
  for (int v0 = 0; v0 < V; v0 += 1024) {
    int v = v0 + threadIdx.x;
    int c = (v < V) ? vox_cnt[base + v] : 0;
    int total;
    int ex = block_excl_scan<1024>(c, &total, s_w);
    if (v < V) {
      vox_off[base + v] = carry + ex;
      vox_cur[base + v] = 0;
    }
    carry += total;
  }
}

__global__ void __launch_bounds__(256) k_vox_fill(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                  const int32_t* __restrict__ apri_cid, const int32_t* __restrict__ vox_off,
                                                  int32_t* __restrict__ vox_cur, int32_t* __restrict__ vox_pts_tmp) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int m_total = scan_counts[b * 8 + 2];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < m_total; m += gridDim.x * blockDim.x) {
    int cid = apri_cid[base + m];
    if (cid < 0) continue;
    int slot = vox_off[base + cid] + atomicAdd(&vox_cur[base + cid], 1);
    vox_pts_tmp[base + slot] = m;
  }
}

// One warp per voxel: order the voxel's points by m (rank by counting), then the strictly sequential
// float intensity mean / population variance of ssc.cpp:261-287, voxel "centre" (:271-277), the index
// triple of the first inserted point (:268-270) and the voxel's bounding box.
// The sums are order dependent, so they stay a chain of dependent adds — but only the adds: the intensities are
// staged in m order (shared memory, or the freshly written CSR for very full voxels), read 32 at a time by the
// whole warp and fed to the chain with shuffles, so no memory latency sits on the dependent path.
constexpr int kVoxStage = 512;  // intensities staged in shared memory per warp

__global__ void __launch_bounds__(256) k_vox_stats(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                   const int32_t* __restrict__ vox_cnt, const int32_t* __restrict__ vox_off,
                                                   const int32_t* __restrict__ vox_pts_tmp, const float4* __restrict__ apri_xyzi,
                                                   int32_t* __restrict__ vox_pts, int32_t* __restrict__ apri_rank,
                                                   float* __restrict__ vox_av, float* __restrict__ vox_cov,
                                                   float* __restrict__ vox_bbox) {
  __shared__ float s_val[8][kVoxStage];
  __shared__ int s_seg[8][kVoxStage];
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  float* sv = s_val[wid];
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < V; v += warps) {
    const int k = vox_cnt[base + v];
    const int o = vox_off[base + v];
    const int32_t* seg = vox_pts_tmp + base + o;
    const bool staged = k <= kVoxStage;
    if (staged) {
      for (int e = lane; e < k; e += 32) s_seg[wid][e] = seg[e];
      __syncwarp();
      seg = s_seg[wid];
    }
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int e = lane; e < k; e += 32) {
      int me = seg[e];
      int r = 0;
      for (int t = 0; t < k; ++t) r += (seg[t] < me) ? 1 : 0;
      vox_pts[base + o + r] = me;
      apri_rank[base + me] = r;
      float4 q = __ldg(&apri_xyzi[base + me]);
      if (staged) sv[r] = q.w;
      lo[0] = fminf(lo[0], q.x);
      lo[1] = fminf(lo[1], q.y);
      lo[2] = fminf(lo[2], q.z);
      hi[0] = fmaxf(hi[0], q.x);
      hi[1] = fmaxf(hi[1], q.y);
      hi[2] = fmaxf(hi[2], q.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
        hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
      }
    __threadfence_block();
    __syncwarp();
    const int32_t* srt = vox_pts + base + o;
    auto value_at = [&](int e) { return staged ? sv[e] : apri_xyzi[base + srt[e]].w; };
    float sum = 0.f;  // every lane runs the same chain: no broadcast at the end
    for (int e0 = 0; e0 < k; e0 += 32) {
      const float mine = (e0 + lane < k) ? value_at(e0 + lane) : 0.f;
      const int m = min(32, k - e0);
      for (int t = 0; t < m; ++t) sum = da(sum, __shfl_sync(0xffffffffu, mine, t));
    }
    const float av = dd(sum, (float)k);
    float cov = 0.f;
    for (int e0 = 0; e0 < k; e0 += 32) {
      const float mine = (e0 + lane < k) ? value_at(e0 + lane) : 0.f;
      const float dlt = ds(mine, av);
      const double sq = __dmul_rn((double)dlt, (double)dlt);  // std::pow(in - av, 2) in double (ssc.cpp:285)
      const int m = min(32, k - e0);
      for (int t = 0; t < m; ++t) cov = (float)__dadd_rn((double)cov, __shfl_sync(0xffffffffu, sq, t));
    }
    cov = dd(cov, (float)k);
    if (lane == 0) {
      vox_av[base + v] = av;
      vox_cov[base + v] = cov;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        vox_bbox[6 * (base + v) + d] = lo[d];
        vox_bbox[6 * (base + v) + 3 + d] = hi[d];
      }
    }
    __syncwarp();
  }
}

// index triple of the first inserted point of every voxel (ssc.cpp:268-270) and the voxel "centre" (:271-277): one
// thread per voxel (the binning and the three libm calls are scalar work; a warp per voxel would waste 31 lanes on them)
__global__ void __launch_bounds__(256) k_vox_center(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                    const int32_t* __restrict__ vox_off, const int32_t* __restrict__ vox_pts,
                                                    const float4* __restrict__ apri_xyzi, BinParams bp, scvod_params sp,
                                                    float* __restrict__ vox_center, int32_t* __restrict__ vox_tri) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const float4 q0 = __ldg(&apri_xyzi[base + vox_pts[base + vox_off[base + v]]]);
    BinResult r = dev_bin_point(q0.x, q0.y, q0.z, bp);
    vox_tri[3 * (base + v) + 0] = r.ri;
    vox_tri[3 * (base + v) + 1] = r.si;
    vox_tri[3 * (base + v) + 2] = r.ei;
    float range_center = da(dm((float)((r.ri * 2 + 1) / 2), sp.range_res), sp.min_dis);
    float sector_center = da(dev_deg2rad_f(dm((float)((r.si * 2 + 1) / 2), sp.sector_res)), sp.min_angle);
    float azimuth_center = da(dev_deg2rad_f(dm((float)((r.ei * 2 + 1) / 2), sp.azimuth_res)), dev_deg2rad_f(sp.min_azimuth));
    vox_center[3 * (base + v) + 0] = dm(range_center, cosf(sector_center));
    vox_center[3 * (base + v) + 1] = dm(range_center, sinf(sector_center));
    vox_center[3 * (base + v) + 2] = dm(range_center, tanf(azimuth_center));
  }
}

// ------------------------------------------------------------------------------------------------
// Cluster preparation: 27-neighbour adjacency in findVoxelNeighbors order (ssc.cpp:395-411), GPU
// connected components, intensity-similarity edges between components (ssc.cpp:587-595), and the
// ordered list of "clustering events" the host needs to reproduce the sequential cluster names
// (ssc.cpp:304-352; see host_cluster.cpp).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vox_nbr(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts, GridSpec g,
                                                 const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ word_rank,
                                                 const int32_t* __restrict__ vox_tri, int32_t* __restrict__ vox_nbr,
                                                 int32_t* __restrict__ vox_root) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  const uint32_t* bm = bitmap + (size_t)b * g.words;
  const int32_t* wr = word_rank + (size_t)b * g.words;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    int ri = vox_tri[3 * (base + v)], si = vox_tri[3 * (base + v) + 1], ei = vox_tri[3 * (base + v) + 2];
    int32_t* out = vox_nbr + 27 * (base + v);
    // the three sector neighbours of a (range, azimuth) row are consecutive bits of the occupancy bitmap: one or two
    // word loads per row instead of three lookups; output order = findVoxelNeighbors order (x, y, z nested, :400-407)
    const int y_lo = max(0, si - 1), y_hi = min(g.sector_num - 1, si + 1);
    for (int x = ri - 1; x <= ri + 1; ++x)
      for (int z = ei - 1; z <= ei + 1; ++z) {
        uint32_t occ = 0, w_lo = 0, w_hi = 0;
        int key_lo = 0;
        if (!(x > g.range_num - 1 || x < 0 || z > g.azimuth_num - 1 || z < 0) && y_lo <= y_hi) {
          key_lo = x * g.sector_num + y_lo + z * g.range_num * g.sector_num + g.key_off;
          const int key_hi = key_lo + (y_hi - y_lo);
          if (key_lo >= 0 && key_hi < g.key_count) {
            w_lo = bm[key_lo >> 5];
            w_hi = ((key_hi >> 5) != (key_lo >> 5)) ? bm[key_hi >> 5] : 0u;
            occ = (uint32_t)((((unsigned long long)w_hi << 32) | w_lo) >> (key_lo & 31)) & ((1u << (y_hi - y_lo + 1)) - 1u);
          }
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int y = si + dy, j = y - y_lo;
          int cid = -1;
          if (y >= y_lo && y <= y_hi && ((occ >> j) & 1u)) {
            const int key = key_lo + j;
            const uint32_t w = ((key >> 5) == (key_lo >> 5)) ? w_lo : w_hi;
            cid = wr[key >> 5] + __popc(w & ((1u << (key & 31)) - 1u));
          }
          out[((x - ri + 1) * 3 + (dy + 1)) * 3 + (z - ei + 1)] = cid;
        }
      }
    vox_root[base + v] = v;
  }
}

__device__ __forceinline__ int uf_find(int32_t* parent, int v) {
  int r = v;
  while (true) {
    int pr = parent[r];
    if (pr == r) break;
    int gp = parent[pr];
    if (gp != pr) parent[r] = gp;  // path halving (benign race: always points to an ancestor)
    r = pr;
  }
  return r;
}

__global__ void __launch_bounds__(256) k_ccl_union(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                   const int32_t* __restrict__ vox_nbr, int32_t* __restrict__ vox_root) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  int32_t* parent = vox_root + base;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const int32_t* nb = vox_nbr + 27 * (base + v);
    for (int t = 0; t < 27; ++t) {
      int u = nb[t];
      if (u < 0 || u >= v) continue;  // each undirected edge once
      int ra = uf_find(parent, v), rb = uf_find(parent, u);
      while (ra != rb) {
        if (ra < rb) {
          int tmp = ra;
          ra = rb;
          rb = tmp;
        }
        int old = atomicCAS(&parent[ra], ra, rb);  // hook the larger root under the smaller
        if (old == ra) break;
        ra = uf_find(parent, old);
        rb = uf_find(parent, rb);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_ccl_flatten(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                     int32_t* __restrict__ vox_root) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    int r = v;
    while (vox_root[base + r] != r) r = vox_root[base + r];
    vox_root[base + v] = r;
  }
}

// directed component edges (root(v) -> root(n)) for every voxel pair that satisfies the intensity
// similarity test of refineClusterByIntensity (ssc.cpp:588-594); self pairs included.  One warp per voxel, one lane
// per (range, azimuth) row of the search cube: the <= 5 sector neighbours of a row are consecutive bits of the
// occupancy bitmap, so a row costs one or two word loads instead of five lookups; only occupied neighbours go on
// to the rank / descriptor loads.  Duplicates inside a round are dropped with match_any, the rest by the hash set.
__global__ void __launch_bounds__(256) k_similar_edges(const int64_t* __restrict__ off, int32_t* __restrict__ scan_counts, GridSpec g,
                                                       const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ word_rank,
                                                       const int32_t* __restrict__ vox_tri, const float* __restrict__ vox_av,
                                                       const float* __restrict__ vox_cov, const int32_t* __restrict__ vox_root,
                                                       int search_c, float intensity_cov, float intensity_diff,
                                                       unsigned long long* __restrict__ edge_hash, int hash_cap,
                                                       int32_t* __restrict__ edge_buf, int edge_cap) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  const uint32_t* bm = bitmap + (size_t)b * g.words;
  const int32_t* wr = word_rank + (size_t)b * g.words;
  unsigned long long* table = edge_hash + (size_t)b * hash_cap;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < V; v += warps) {
    const int ri = vox_tri[3 * (base + v)], si = vox_tri[3 * (base + v) + 1], ei = vox_tri[3 * (base + v) + 2];
    const int size = ((double)ri > (double)g.range_num * 0.6) ? 1 : search_c;  // ssc.cpp:397-399
    const int side = 2 * size + 1, rows = side * side;
    const float avv = vox_av[base + v];
    const int rv = vox_root[base + v];
    const int y_lo = max(0, si - size), y_hi = min(g.sector_num - 1, si + size);  // clipped, no wrap (ssc.cpp:400-407)
    for (int r0 = 0; r0 < rows; r0 += 32) {  // search_c <= 2: a single round
      const int row = r0 + lane;
      uint32_t occ = 0;   // occupied sectors y_lo + j of this lane's row
      uint32_t w_lo = 0, w_hi = 0;
      int key_lo = 0;
      if (row < rows && y_lo <= y_hi) {
        const int x = ri - size + row / side, z = ei - size + row % side;
        if (!(x > g.range_num - 1 || x < 0 || z > g.azimuth_num - 1 || z < 0)) {
          key_lo = x * g.sector_num + y_lo + z * g.range_num * g.sector_num + g.key_off;
          const int key_hi = key_lo + (y_hi - y_lo);
          if (key_lo >= 0 && key_hi < g.key_count) {
            w_lo = bm[key_lo >> 5];
            w_hi = ((key_hi >> 5) != (key_lo >> 5)) ? bm[key_hi >> 5] : 0u;
            const unsigned long long both = ((unsigned long long)w_hi << 32) | w_lo;
            occ = (uint32_t)(both >> (key_lo & 31)) & ((1u << (y_hi - y_lo + 1)) - 1u);
          }
        }
      }
      // Nearly every similar neighbour is in the voxel's own component: the self edge (rv, rv) is only remembered here and
      // inserted once per voxel after the loop; the rounds below only deal with edges to OTHER components.
      bool self_edge = false;
      auto insert_edge = [&](int ru) {
        unsigned long long key = ((unsigned long long)(uint32_t)rv << 32) | (uint32_t)ru;
        unsigned long long h = key * 0x9E3779B97F4A7C15ULL;
        int slot = (int)(h >> 40) % hash_cap;
        bool inserted = false, done = false;
        for (int probe = 0; probe < hash_cap && !done; ++probe) {
          unsigned long long old = atomicCAS(&table[slot], ~0ull, key);
          if (old == ~0ull) {
            inserted = true;
            done = true;
          } else if (old == key) {
            done = true;
          } else {
            slot = (slot + 1 == hash_cap) ? 0 : slot + 1;
          }
        }
        if (!done) atomicExch(&scan_counts[b * 8 + 7], -1 << 20);  // table full
        if (inserted) {
          int e = atomicAdd(&scan_counts[b * 8 + 7], 1);
          if (e >= 0 && e < edge_cap) {
            edge_buf[((size_t)b * edge_cap + e) * 2] = rv;
            edge_buf[((size_t)b * edge_cap + e) * 2 + 1] = ru;
          }
        }
      };
      for (int j = 0; j < side; ++j) {  // warp-uniform trip count; lanes without a j-th sector idle
        int ru = -1;
        if ((occ >> j) & 1u) {
          const int key = key_lo + j;
          const uint32_t w = ((key >> 5) == (key_lo >> 5)) ? w_lo : w_hi;
          const int u = wr[key >> 5] + __popc(w & ((1u << (key & 31)) - 1u));
          if (vox_cov[base + u] <= intensity_cov && fabsf(ds(avv, vox_av[base + u])) <= intensity_diff) ru = vox_root[base + u];
        }
        if (ru == rv) {
          self_edge = true;
          ru = -1;
        }
        if (__ballot_sync(0xffffffffu, ru >= 0) == 0u) continue;
        const unsigned same = __match_any_sync(0xffffffffu, ru);
        if (ru < 0 || (same & ((1u << lane) - 1u))) continue;  // nothing, or a lower lane inserts this root
        insert_edge(ru);
      }
      const unsigned self_mask = __ballot_sync(0xffffffffu, self_edge);
      if (self_mask && lane == __ffs(self_mask) - 1) insert_edge(rv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cluster names of SSC::clusterAndCreateFrame (ssc.cpp:299-354) replayed on the device.
//
// The reference walks the apri points in order and propagates names through the <=27 voxels around each
// point (oc/nc rules of :323-351, mergeClusters :413-419 renames the current point's cluster to the
// neighbour's), so names depend on the visiting sequence.  At voxel granularity (no index is -1):
//   * a voxel is unlabelled / only its first point labelled / fully labelled, and its labelled points are in
//     one set;  an unlabelled visitor skips unlabelled voxels until the first labelled one (position p),
//     adopts that set, and from then on labels or merges everything it meets;
//   * the union keeps the name of the LAST set met for the first time (every merge renames the current
//     set to the neighbour's), i.e. of the highest lane whose root occurs for the first time;
//   * after an event whose point was already labelled, or that skipped nothing, the voxel is "stable":
//     all 27 neighbours share its set for good and later points of the voxel are no-ops.  Only the
//     first three points of a voxel can find it unstable: those are the events (k_events).
// One event = one warp step: lane k owns neighbour k (findVoxelNeighbors order), union-find with path
// halving lives in shared memory, the order-dependent part is resolved with ballot / match_any.
//
// Parallelism.  An event only touches voxels of its own 26-connected component (k_ccl_*), so components
// replay independently; the single coupling is the name counter (:345-346), and the k-th "new class" event
// of the scan in event order simply gets name 5 + k.  One CTA per scan therefore
//   A. counts the events of every component, deals the components to its NW warps (largest first to the
//      least loaded warp, the many small ones by water-filling) ...
//   B. ... and splits the ordered event list into one ordered list per warp (stable partition);
//   C. every warp replays its own list; a new class records its creating event instead of a number;
//   D. names = 5 + rank of the creating event among all creating events (popcount prefix over a bit per event).
// ------------------------------------------------------------------------------------------------
constexpr int kReplayWarps = 8;  // warps per scan: components are dealt to them by event count
constexpr int kReplayRows = 32;   // events per chunk; neighbour rows of the next chunk are prefetched while one is replayed

template <int NW, bool GLOBAL>
__global__ void __launch_bounds__(NW * 32) k_name_replay(const int64_t* __restrict__ off, int32_t* __restrict__ scan_counts,
                                                         const int32_t* __restrict__ ev_cid, const int32_t* __restrict__ vox_root,
                                                         const int32_t* __restrict__ vox_nbr, int2* __restrict__ ev_list,
                                                         int32_t* __restrict__ g_parent, int32_t* __restrict__ g_setname,
                                                         int32_t* __restrict__ g_first, int32_t* __restrict__ g_state,
                                                         int32_t* __restrict__ g_flags, int32_t* __restrict__ vox_name,
                                                         int32_t* __restrict__ name_first, int name_cap) {
  constexpr int T = NW * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[T / 32 + 1];
  __shared__ int s_load[NW], s_lbase[NW + 1], s_cursor[NW], s_free[NW + 1];
  __shared__ int s_wcnt[NW][NW];
  __shared__ int s_big[32], s_nbig, s_target;
  const int b = blockIdx.x;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  const int E = scan_counts[b * 8 + 5];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nfw = (E + 31) >> 5;  // one "new class" bit per event
  // dynamic shared memory: [neighbour-row ring NW x 2 x 32 x 32 ints][union-find state 6 B / voxel][flags + prefix]
  int(*ring)[kReplayRows][32] = reinterpret_cast<int(*)[kReplayRows][32]>(smem_raw) + 2 * wid;
  unsigned char* sm_state = smem_raw + sizeof(int) * NW * 2 * kReplayRows * 32;
  int32_t *parent, *setname, *first_ev;
  uint8_t *state, *stable;
  uint32_t* flags;
  int32_t* fprefix;
  if (GLOBAL) {
    parent = g_parent + base;
    setname = g_setname + base;
    first_ev = g_first + base;
    state = reinterpret_cast<uint8_t*>(g_state + base);
    stable = state + V;  // g_state has 4 bytes per voxel
    flags = reinterpret_cast<uint32_t*>(g_flags + base);  // E <= M <= N ints available: nfw flags, then nfw prefixes
    fprefix = g_flags + base + nfw;
  } else {
    // shared memory holds what sits on the dependent path of every event (parent, state, stable: 6 B / voxel, so scans of
    // up to ~25k voxels fit); set names and first events are written once and read at the end: global scratch
    parent = reinterpret_cast<int32_t*>(sm_state);
    flags = reinterpret_cast<uint32_t*>(parent + V);
    fprefix = reinterpret_cast<int32_t*>(flags + nfw);
    state = reinterpret_cast<uint8_t*>(fprefix + nfw);
    stable = state + V;
    setname = g_setname + base;
    first_ev = g_first + base;
  }
  const int32_t* ev = ev_cid + base;
  const int32_t* root = vox_root + base;
  const int32_t* nbr = vox_nbr + 27 * base;
  int2* lst = ev_list + base;

  // ---- A. events per component (cnt lives in parent[], the owner warp of a root in state[]) ---------------
  int32_t* cnt = parent;
  uint8_t* owner = state;
  for (int v = tid; v < V; v += T) cnt[v] = 0;
  if (tid < NW) {
    s_load[tid] = 0;
    s_cursor[tid] = 0;
  }
  if (tid == 0) s_nbig = 0;
  __syncthreads();
  for (int e = tid; e < E; e += T) atomicAdd(&cnt[root[ev[e]]], 1);
  __syncthreads();
  const int thr = E / (4 * NW) + 1;  // fewer than 4 * NW <= 32 components can be this large
  for (int v = tid; v < V; v += T)
    if (cnt[v] >= thr) s_big[atomicAdd(&s_nbig, 1)] = v;
  __syncthreads();
  if (wid == 0) {  // largest first to the least loaded warp; lane w < NW keeps the load of warp w
    const int nbig = s_nbig;
    const int my_root = lane < nbig ? s_big[lane] : -1;
    int my_cnt = lane < nbig ? cnt[my_root] : -1;
    int load = 0;
    for (int it = 0; it < nbig; ++it) {
      int best = (my_cnt << 5) | (31 - lane);  // max count, ties to the lowest lane
      if (my_cnt < 0) best = -1;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, s));
      const int src = 31 - (best & 31), c = best >> 5;
      int least = lane < NW ? ((load << 5) | lane) : 0x7fffffff;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) least = min(least, __shfl_xor_sync(0xffffffffu, least, s));
      const int w = least & 31;
      if (lane == w) load += c;
      if (lane == src) {
        owner[my_root] = (uint8_t)w;
        my_cnt = -1;
      }
    }
    // water level for the small components: nobody above max(largest load, ceil(E / NW))
    int mx = lane < NW ? load : 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    const int target = max(mx, (E + NW - 1) / NW);
    const int fr = lane < NW ? target - load : 0;
    const int inc = warp_incl_scan(fr);
    if (lane < NW) s_free[lane + 1] = inc;
    if (lane == 0) {
      s_free[0] = 0;
      s_target = target;
    }
  }
  __syncthreads();
  {
    int carry = 0;
    for (int v0 = 0; v0 < V; v0 += T) {
      const int v = v0 + tid;
      const int c = (v < V) ? cnt[v] : 0;
      const bool small = c > 0 && c < thr;
      int total;
      const int ex = block_excl_scan<T>(small ? c : 0, &total, s_scan);
      if (small) {
        const int pos = carry + ex;
        int w = 0;
#pragma unroll
        for (int k = 1; k < NW; ++k) w += (s_free[k] <= pos) ? 1 : 0;  // last warp whose free range starts at or before pos
        owner[v] = (uint8_t)w;
      }
      carry += total;
    }
  }
  __syncthreads();
  for (int v = tid; v < V; v += T) {
    const int c = cnt[v];
    if (c > 0) atomicAdd(&s_load[owner[v]], c);
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < NW; ++w) {
      s_lbase[w] = run;
      run += s_load[w];
    }
    s_lbase[NW] = run;
  }
  __syncthreads();
  // ---- B. stable partition of the event list by owner warp ---------------------------------------------------
  for (int e0 = 0; e0 < E; e0 += T) {
    const int e = e0 + tid;
    int cid = -1, own = -1 - lane;  // idle lanes match nobody
    if (e < E) {
      cid = ev[e];
      own = owner[root[cid]];
    }
    if (lane < NW) s_wcnt[wid][lane] = 0;
    __syncwarp();
    const unsigned same = __match_any_sync(0xffffffffu, own);
    const int rank = __popc(same & ((1u << lane) - 1u));
    if (own >= 0 && rank == 0) s_wcnt[wid][own] = __popc(same);
    __syncthreads();
    if (tid < NW) {
      int run = s_cursor[tid];
      for (int w = 0; w < NW; ++w) {
        const int c = s_wcnt[w][tid];
        s_wcnt[w][tid] = run;
        run += c;
      }
      s_cursor[tid] = run;
    }
    __syncthreads();
    if (own >= 0) lst[s_lbase[own] + s_wcnt[wid][own] + rank] = make_int2(e, cid);
    __syncwarp();
  }
  __syncthreads();
  // ---- C. replay ------------------------------------------------------------------------------------------------
  for (int v = tid; v < V; v += T) {
    parent[v] = v;
    setname[v] = -1;
    first_ev[v] = 0x7fffffff;
    state[v] = 0;
    stable[v] = 0;
  }
  for (int w = tid; w < nfw; w += T) flags[w] = 0u;
  __threadfence_block();
  __syncthreads();
  auto find = [&](int v) {
    int r = v;
    while (true) {
      int pr = parent[r];
      if (pr == r) break;
      int gp = parent[pr];
      if (gp != pr) parent[r] = gp;
      r = pr;
    }
    return r;
  };
  {
    const int L = s_load[wid];
    const int2* my = lst + s_lbase[wid];
    // Neighbour rows (27 ints, L2 resident) are fetched one chunk of 32 events ahead with cp.async, and only for
    // voxels that are not yet stable (stable never resets): the common no-op events never touch global memory.
    // The <= 3 events of a voxel usually sit in the same chunk: the row is fetched once per distinct voxel (by the first
    // lane of its group, `same` = lanes with the same voxel) and the later events read it from that lane's ring slot.
    auto prefetch = [&](const int2 evn, unsigned same, int buf) {
      const bool need = evn.y >= 0 && !stable[evn.y] && (__ffs(same) - 1 == lane);
      unsigned pm = __ballot_sync(0xffffffffu, need);
      while (pm) {
        const int j = __ffs(pm) - 1;
        pm &= pm - 1;
        const int Wj = __shfl_sync(0xffffffffu, evn.y, j);
        if (lane < 27) {
          unsigned dst = (unsigned)__cvta_generic_to_shared(&ring[buf][j][lane]);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(nbr + 27 * (size_t)Wj + lane));
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    int2 nxt = (lane < L) ? my[lane] : make_int2(0, -1);
    unsigned nxt_same = __match_any_sync(0xffffffffu, nxt.y >= 0 ? nxt.y : -1 - lane);
    prefetch(nxt, nxt_same, 0);
    for (int c = 0; c * 32 < L; ++c) {
      const int2 cur = nxt;
      const unsigned cur_same = nxt_same;
      const int j1 = (c + 1) * 32 + lane;
      nxt = (j1 < L) ? my[j1] : make_int2(0, -1);
      nxt_same = __match_any_sync(0xffffffffu, nxt.y >= 0 ? nxt.y : -1 - lane);
      prefetch(nxt, nxt_same, (c + 1) & 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
      __syncwarp();
      unsigned pend = __ballot_sync(0xffffffffu, cur.y >= 0 && !stable[cur.y]);
      while (pend) {
        const int i = __ffs(pend) - 1;
        pend &= pend - 1;
        const int W = __shfl_sync(0xffffffffu, cur.y, i);
        const int e = __shfl_sync(0xffffffffu, cur.x, i);
        const unsigned group = __shfl_sync(0xffffffffu, cur_same, i);  // events of this chunk on the same voxel
        if (stable[W]) {  // became stable inside this chunk (warp-uniform): its other events here are no-ops too
          pend &= ~group;
          continue;
        }
        const int slot = __ffs(group) - 1;  // the row was fetched by the first event of the voxel in this chunk
        const int Vn = (lane < 27) ? ring[c & 1][slot][lane] : -1;
        if (lane == 0) atomicMin(&first_ev[W], e);  // fire-and-forget reduction: no load on the event path
        const bool exist = Vn >= 0;
        const int st = exist ? state[Vn] : 0;
        const bool lab = exist && st != 0;
        const int r = lab ? find(Vn) : -1;
        const bool labelled = (state[W] == 2);
        const int oc0 = labelled ? find(W) : -1;
        __syncwarp();
        const unsigned lab_mask = __ballot_sync(0xffffffffu, lab);
        const unsigned unl_mask = __ballot_sync(0xffffffffu, exist && !lab);
        if (!labelled && lab_mask == 0u) {  // a new class (:345-351): named after the creating event, numbered in D
          if (lane == 0) {
            atomicOr(&flags[e >> 5], 1u << (e & 31));
            parent[W] = W;
            setname[W] = e;
            state[W] = 2;
            stable[W] = 1;
          }
          __syncwarp();
          if (exist && Vn != W) {
            parent[Vn] = W;
            state[Vn] = 2;
          }
          __syncwarp();
          continue;
        }
        // roots met for the first time, in visit order (a root equal to oc0 was "met" before the loop)
        bool first_occ;
        unsigned fo_mask;
        {
          const int lane0 = __ffs(lab_mask) - 1;  // lab_mask != 0 here: a labelled W is its own (labelled) neighbour
          const int r0 = __shfl_sync(0xffffffffu, r, lane0);
          if (__ballot_sync(0xffffffffu, lab && r != r0) == 0u) {  // the usual case: every labelled neighbour is in one set
            first_occ = (lane == lane0) && (r0 != oc0);
            fo_mask = (r0 != oc0) ? (1u << lane0) : 0u;
          } else {
            const unsigned same = __match_any_sync(0xffffffffu, lab ? r : (-2 - lane));
            first_occ = lab && (r != oc0) && ((same & ((1u << lane) - 1u)) == 0u);
            fo_mask = __ballot_sync(0xffffffffu, first_occ);
          }
        }
        int f;  // surviving root: the last first-met set (mergeClusters renames oc to nc each time)
        if (fo_mask) {
          f = __shfl_sync(0xffffffffu, r, 31 - __clz(fo_mask));
        } else {
          f = oc0;
        }
        const int p = labelled ? -1 : (__ffs(lab_mask) - 1);  // position where the visitor becomes labelled
        if (first_occ && r != f) parent[r] = f;
        if (lane == 0 && labelled && oc0 != f) parent[oc0] = f;
        __syncwarp();
        if (lab) state[Vn] = 2;
        const bool take = exist && !lab && lane > p;  // unlabelled voxels met after the visitor got its label (:338)
        if (take) {
          parent[Vn] = f;
          state[Vn] = 2;
        }
        const unsigned skipped = unl_mask & ((p >= 0) ? ((1u << p) - 1u) : 0u);
        __syncwarp();
        if (lane == 0) {
          if (state[W] == 0) {  // only this (first) point of W got the label
            state[W] = 1;
            parent[W] = f;
          }
          stable[W] = skipped ? 0 : 1;
        }
        __syncwarp();
      }
      __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
  }
  __threadfence_block();
  __syncthreads();
  // ---- D. numbers: the k-th creating event (in event order) is name 5 + k (cluster_name starts at 4, :300) ----
  int n_names;
  {
    int carry = 0;
    for (int w0 = 0; w0 < nfw; w0 += T) {
      const int w = w0 + tid;
      const int c = (w < nfw) ? __popc(flags[w]) : 0;
      int total;
      const int ex = block_excl_scan<T>(c, &total, s_scan);
      if (w < nfw) fprefix[w] = carry + ex;
      carry += total;
    }
    n_names = carry;
  }
  __syncthreads();
  const int cluster_name = 4 + n_names;
  // final names + first point (event) of every name, which fixes the insertion order of cluster_pt (:360-375)
  int32_t* nf = name_first + (size_t)b * name_cap;
  for (int i = tid; i <= cluster_name && i < name_cap; i += T) nf[i] = 0x7fffffff;
  __syncthreads();
  for (int v = tid; v < V; v += T) {
    int r = v;  // read-only walk: other threads resolve voxels of the same component at the same time
    while (parent[r] != r) r = parent[r];
    const int ec = setname[r];
    int nm = -1;
    if (ec >= 0) nm = 5 + fprefix[ec >> 5] + __popc(flags[ec >> 5] & ((1u << (ec & 31)) - 1u));
    vox_name[base + v] = nm;
    if (nm >= 0 && nm < name_cap) atomicMin(&nf[nm], first_ev[v]);
  }
  if (tid == 0) scan_counts[b * 8 + 6] = cluster_name;
}

// ordered compaction of the apri points whose rank inside their voxel is < 3 ("clustering events").  One CTA per scan;
// every warp owns a contiguous slice: count (ballot / popcount), ONE block scan over the 32 warp totals, then the same
// walk again writing at the warp's offset — two barriers per scan instead of three per 1024 points.
__global__ void __launch_bounds__(1024) k_events(const int64_t* __restrict__ off, int32_t* __restrict__ scan_counts,
                                                 const int32_t* __restrict__ apri_cid, const int32_t* __restrict__ apri_rank,
                                                 int32_t* __restrict__ ev_cid) {
  __shared__ int s_w[33];
  const int b = blockIdx.x;
  const int64_t base = off[b];
  const int M = scan_counts[b * 8 + 2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int slice = (((M + 31) / 32) + 31) & ~31;  // per warp, multiple of 32
  const int m0 = min(M, wid * slice), m1 = min(M, m0 + slice);
  int cnt = 0;
  for (int m = m0 + lane; m < m1; m += 32) cnt += (apri_cid[base + m] >= 0 && apri_rank[base + m] < 3) ? 1 : 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
  int total;
  const int ex = block_excl_scan<1024>(lane == 0 ? cnt : 0, &total, s_w);  // lane 0 of warp w carries the warp total
  int pos = __shfl_sync(0xffffffffu, ex, 0);
  for (int j = m0; j < m1; j += 32) {
    const int m = j + lane;
    const int cid = (m < m1) ? apri_cid[base + m] : -1;
    const bool f = cid >= 0 && apri_rank[base + m] < 3;
    const unsigned mask = __ballot_sync(0xffffffffu, f);
    if (f) ev_cid[base + pos + __popc(mask & ((1u << lane) - 1u))] = cid;
    pos += __popc(mask);
  }
  if (threadIdx.x == 0) scan_counts[b * 8 + 5] = total;
}

// ------------------------------------------------------------------------------------------------
// stand-alone binning of an arbitrary cloud (scvod_bin) and the tracking diff kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bin_only(const float4* __restrict__ pts, int n, BinParams bp, uint8_t* pass, int32_t* vid,
                                                  int32_t* ri, int32_t* si, int32_t* ei, float* range, float* angle, float* azimuth) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&pts[i]);
    BinResult r = dev_bin_point(p.x, p.y, p.z, bp);
    if (pass) pass[i] = r.pass ? 1 : 0;
    if (vid) vid[i] = r.vid;
    if (ri) ri[i] = r.ri;
    if (si) si[i] = r.si;
    if (ei) ei[i] = r.ei;
    if (range) range[i] = r.dis;
    if (angle) angle[i] = r.angle;
    if (azimuth) azimuth[i] = r.azimuth;
  }
}

// ------------------------------------------------------------------------------------------------
// Tracking diff (SSC::tracking, ssc.cpp:1274-1321): all car clusters of frame_pre_ in one launch.
// The clouds are described by segments (own points gathered by apri index, or carried points that
// already live on the device from the previous pair).  Per point: transformCloud (utility.h:394-406,
// left-to-right float, no FMA) -> ungated re-binning (ssc.cpp:1280-1286) -> next.hash_cloud.find
// (ssc.cpp:1304) via the bitmap rank.  Instead of shipping one hit per point to the host, the kernel
// keeps, per (cluster, hit voxel), the smallest point position: that is exactly the information the
// reference's remap_name needs (set of hit voxels per label + order of first occurrence).
// ------------------------------------------------------------------------------------------------
// Two ways to tell the kernel which points to take:
//   RUNS = true   the usual one: a car cluster is a few runs of the frame's device-resident car CSR (csr_ptoff / csr_vox /
//                 csr_part, uploaded once per batch) plus carried ranges; the run table (<= kTrackMaxRuns entries) travels in
//                 the kernel arguments, so a frame pair costs NO host->device copy (on the copy engine it would queue
//                 behind the bulk scan uploads of the other contexts);
//   RUNS = false  one uploaded segment per voxel + per-block index (scvod_initialization, or more runs than fit).
template <bool RUNS>
__global__ void __launch_bounds__(256) k_track(const float4* __restrict__ own, const int32_t* __restrict__ vox_off,
                                               const int32_t* __restrict__ vox_pts, const float4* __restrict__ carried,
                                               const int4* __restrict__ segs,
                                               const int32_t* __restrict__ first_seg /* per block of 256 points */,
                                               int nseg, const __grid_constant__ TrackRuns runs, const int32_t* __restrict__ csr_ptoff,
                                               const int32_t* __restrict__ csr_vox, const int32_t* __restrict__ csr_part,
                                               int k, Mat34 T, BinParams bp, GridSpec g,
                                               const uint32_t* __restrict__ bm, const int32_t* __restrict__ wr, int vn,
                                               float4* __restrict__ out_xyzi, unsigned long long* __restrict__ first,
                                               int32_t* __restrict__ ctr /* [0] distinct hits, [1] finished CTAs */,
                                               int32_t* __restrict__ hit_list, int32_t* __restrict__ out_quads, int cap_quads) {
  __shared__ int s_last;
  // A block of 256 points only needs the <= 257 segments that overlap it, found through the per-block index the host
  // wrote next to the table: they are staged in shared memory, the per-point binary search never leaves the SM.
  __shared__ int4 s_seg[257];
  __shared__ int s_range[2];
  const int nblk = (k + 255) >> 8;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    int4 sg;  // x = dst_off, y = source (>= 0: voxel of frame_pre_, its points come from the voxel CSR; < 0: carried range
              // at -1-y), z = cluster, w = order of the segment inside the cluster's cloud (part index / carried ordinal)
    const int i = (blk << 8) + threadIdx.x;
    if (RUNS) {
      if (i >= k) continue;
      int lo = 0, hi = runs.n - 1;  // last run with dst_off <= i (kernel-argument table: constant bank)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (runs.dst_off[mid] <= i)
          lo = mid;
        else
          hi = mid - 1;
      }
      const int src = runs.src[lo];
      sg.z = runs.cluster[lo];
      if (src >= 0) {  // own voxels: find the voxel inside the run by its point offsets (a short, L1-resident slice)
        const int32_t* po = csr_ptoff + src;
        const int jr = i - runs.dst_off[lo] + po[0];
        int a = 0, b = runs.len[lo] - 1;  // last CSR position with ptoff <= jr
        while (a < b) {
          const int mid = (a + b + 1) >> 1;
          if (po[mid] <= jr)
            a = mid;
          else
            b = mid - 1;
        }
        sg.x = i - (jr - po[a]);
        sg.y = csr_vox[src + a];
        sg.w = runs.order[lo] + csr_part[src + a];
      } else {
        sg.x = runs.dst_off[lo];
        sg.y = src;
        sg.w = runs.order[lo];
      }
    } else {
      __syncthreads();
      if (threadIdx.x < 2) s_range[threadIdx.x] = (blk + (int)threadIdx.x < nblk) ? first_seg[blk + threadIdx.x] : nseg - 1;
      __syncthreads();
      const int s0 = s_range[0], cnt = s_range[1] - s0 + 1;
      for (int t = threadIdx.x; t < cnt; t += 256) s_seg[t] = segs[s0 + t];
      __syncthreads();
      if (i >= k) continue;
      int lo = 0, hi = cnt - 1;  // last segment with dst_off <= i
      while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (s_seg[mid].x <= i)
          lo = mid;
        else
          hi = mid - 1;
      }
      sg = s_seg[lo];
    }
    const int j = i - sg.x;
    float4 p;
    unsigned low;
    if (sg.y >= 0) {
      int m = vox_pts[vox_off[sg.y] + j];
      p = __ldg(&own[m]);
      low = (unsigned)m;  // inside a part the reference's cloud is in ascending apri index (ssc.cpp:360-380)
    } else {
      p = __ldg(&carried[(-1 - sg.y) + j]);
      low = (unsigned)j;
    }
    float4 q;
    q.x = da(da(da(dm(T.m[0], p.x), dm(T.m[1], p.y)), dm(T.m[2], p.z)), T.m[3]);
    q.y = da(da(da(dm(T.m[4], p.x), dm(T.m[5], p.y)), dm(T.m[6], p.z)), T.m[7]);
    q.z = da(da(da(dm(T.m[8], p.x), dm(T.m[9], p.y)), dm(T.m[10], p.z)), T.m[11]);
    q.w = p.w;
    out_xyzi[i] = q;
    BinResult r = dev_bin_point(q.x, q.y, q.z, bp);
    int hit = vox_lookup(bm, wr, g, r.vid);
    if (hit >= 0) {
      const int e = sg.z * vn + hit;
      const unsigned long long old = atomicMin(&first[e], ((unsigned long long)(unsigned)sg.w << 32) | low);
      if (old == ~0ull) {  // first point to touch this (cluster, voxel): remember the entry for the epilogue
        const int slot = atomicAdd(&ctr[0], 1);
        if (slot < cap_quads) hit_list[slot] = e;
      }
    }
  }
  // ---- epilogue by the last CTA to finish: (cluster, voxel, first-occurrence key) quads straight into host-mapped
  // pinned memory; every consumed table entry goes back to "empty", so the table needs no memset between pairs ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctr[1], 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nh = *reinterpret_cast<volatile int32_t*>(&ctr[0]);
  const int take = min(nh, cap_quads);
  // several independent entries per thread and round: the three dependent accesses (list -> table -> host) overlap
  constexpr int kEpi = 8;
  for (int t0 = 0; t0 < take; t0 += 256 * kEpi) {
    int e[kEpi];
    unsigned long long f[kEpi];
#pragma unroll
    for (int u = 0; u < kEpi; ++u) {
      const int t = t0 + u * 256 + threadIdx.x;
      e[u] = (t < take) ? hit_list[t] : -1;
    }
#pragma unroll
    for (int u = 0; u < kEpi; ++u) f[u] = (e[u] >= 0) ? first[e[u]] : 0ull;
#pragma unroll
    for (int u = 0; u < kEpi; ++u) {
      const int t = t0 + u * 256 + threadIdx.x;
      if (e[u] >= 0) {
        first[e[u]] = ~0ull;
        reinterpret_cast<int4*>(out_quads + 4)[t] =
            make_int4(e[u] / vn, e[u] % vn, (int)(unsigned)(f[u] >> 32), (int)(unsigned)(f[u] & 0xffffffffu));
      }
    }
  }
  __threadfence_system();  // the quads must have reached host memory before the host sees the count
  __syncthreads();
  if (threadIdx.x == 0) {
    ctr[0] = 0;
    ctr[1] = 0;
    out_quads[1] = 0;
    // the host polls this word (it stores -1 before the launch); > cap_quads tells it that the table overflowed
    *reinterpret_cast<volatile int32_t*>(out_quads) = nh;
  }
}

// per-point classes of every frame of a batch from the per-voxel classes decided on the host
__global__ void __launch_bounds__(256) k_final_labels(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                      const int32_t* __restrict__ apri_src, const int32_t* __restrict__ apri_cid,
                                                      const int32_t* __restrict__ vcls_off, const uint8_t* __restrict__ vcls,
                                                      uint8_t* __restrict__ cls) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int m_total = scan_counts[b * 8 + 2];
  const uint8_t* vc = vcls + vcls_off[b];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < m_total; m += gridDim.x * blockDim.x) {
    int cid = apri_cid[base + m];
    if (cid >= 0) cls[base + apri_src[base + m]] = vc[cid];
  }
}

// static submap: every non-dynamic point of the frames of a batch moved to the map frame (transformCloud arithmetic
// with the frame's pose).  A CTA owns a contiguous chunk of a scan and every warp a contiguous part of it: the static
// points are counted first (1 B / point), the CTA reserves its output range with ONE atomic on the global counter,
// and the second pass writes with ballot / popcount ranks (no per-warp atomics on a single address).
__global__ void __launch_bounds__(256) k_submap(const float4* __restrict__ pts, const uint8_t* __restrict__ cls,
                                                const int64_t* __restrict__ off, const float* __restrict__ Ts, int first_scan,
                                                float4* __restrict__ out, unsigned long long* __restrict__ counter, long long cap) {
  __shared__ int s_wcnt[8];
  __shared__ unsigned long long s_base;
  const int b = first_scan + blockIdx.y;
  const int64_t base = off[b];
  const int n = (int)(off[b + 1] - base);
  float t[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) t[i] = Ts[blockIdx.y * 12 + i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per_warp = (((n + gridDim.x * 8 - 1) / (gridDim.x * 8)) + 31) & ~31;  // multiple of 32: aligned 1-byte loads
  const int i0 = min(n, (blockIdx.x * 8 + wid) * per_warp), i1 = min(n, i0 + per_warp);
  int cnt = 0;
  for (int i = i0 + lane; i < i1; i += 32) cnt += (cls[base + i] != SCVOD_PT_DYNAMIC) ? 1 : 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
  if (lane == 0) s_wcnt[wid] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < 8; ++w) {
      const int c = s_wcnt[w];
      s_wcnt[w] = run;
      run += c;
    }
    s_base = run ? atomicAdd(counter, (unsigned long long)run) : 0ull;
  }
  __syncthreads();
  long long pos0 = (long long)s_base + s_wcnt[wid];
  for (int j = i0; j < i1; j += 32) {
    const int i = j + lane;
    const bool keep = (i < i1) && cls[base + i] != SCVOD_PT_DYNAMIC;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const long long pos = pos0 + __popc(mask & ((1u << lane) - 1));
      if (pos < cap) {
        float4 p = __ldg(&pts[base + i]);
        float4 q;
        q.x = da(da(da(dm(t[0], p.x), dm(t[1], p.y)), dm(t[2], p.z)), t[3]);
        q.y = da(da(da(dm(t[4], p.x), dm(t[5], p.y)), dm(t[6], p.z)), t[7]);
        q.z = da(da(da(dm(t[8], p.x), dm(t[9], p.y)), dm(t[10], p.z)), t[11]);
        q.w = p.w;
        out[pos] = q;
      }
    }
    pos0 += __popc(mask);
  }
}

// gather of the per-scan voxel tables into one packed buffer (one D2H instead of hundreds)
__global__ void __launch_bounds__(256) k_pack(const PackDesc* __restrict__ descs, int32_t* __restrict__ out) {
  const PackDesc d = descs[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x) out[d.dst + i] = d.src[i];
}

__global__ void k_atan2f_probe(const float* y, const float* x, float* out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = dev_atan2f(y[i], x[i]);
}

// ------------------------------------------------------------------------------------------------
// optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline block)
// ------------------------------------------------------------------------------------------------
namespace {
struct TimedLaunch {
  int name_id;
  cudaEvent_t e0, e1;
};
bool g_timing = false;
std::vector<std::string> g_names;
std::vector<double> g_ms;
std::vector<long long> g_cnt;
std::vector<TimedLaunch> g_pending;
std::vector<cudaEvent_t> g_event_pool;
std::mutex g_timing_mu;  // contexts on different host threads share the table

cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
int name_id(const char* name) {
  for (size_t i = 0; i < g_names.size(); ++i)
    if (g_names[i] == name) return (int)i;
  g_names.push_back(name);
  g_ms.push_back(0.0);
  g_cnt.push_back(0);
  return (int)g_names.size() - 1;
}
struct ScopedTimer {
  cudaStream_t st;
  TimedLaunch t;
  bool on;
  ScopedTimer(const char* name, cudaStream_t s) : st(s), on(g_timing) {
    if (on) {
      {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        t.name_id = name_id(name);
        t.e0 = get_event();
        t.e1 = get_event();
      }
      cudaEventRecord(t.e0, st);
    }
  }
  ~ScopedTimer() {
    if (on) {
      cudaEventRecord(t.e1, st);
      std::lock_guard<std::mutex> lk(g_timing_mu);
      g_pending.push_back(t);
    }
  }
};
}  // namespace
#define TIMED(name, st) ScopedTimer timer__(name, st)

LaunchTimer::LaunchTimer(const char* name, void* stream) : st(stream), id(0), e0(nullptr), e1(nullptr), on(g_timing) {
  if (on) {
    {
      std::lock_guard<std::mutex> lk(g_timing_mu);
      id = name_id(name);
      e0 = (void*)get_event();
      e1 = (void*)get_event();
    }
    cudaEventRecord((cudaEvent_t)e0, (cudaStream_t)st);
  }
}
LaunchTimer::~LaunchTimer() {
  if (on) {
    cudaEventRecord((cudaEvent_t)e1, (cudaStream_t)st);
    TimedLaunch t;
    t.name_id = id;
    t.e0 = (cudaEvent_t)e0;
    t.e1 = (cudaEvent_t)e1;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_pending.push_back(t);
  }
}
#define TSTREAM ((cudaStream_t)stream_)

void timing_enable(bool on) { g_timing = on; }
void timing_reset() {
  timing_collect();
  for (auto& v : g_ms) v = 0;
  for (auto& v : g_cnt) v = 0;
}
void timing_collect() {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& t : g_pending) {
    cudaEventSynchronize(t.e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t.e0, t.e1);
    g_ms[t.name_id] += ms;
    g_cnt[t.name_id] += 1;
    g_event_pool.push_back(t.e0);
    g_event_pool.push_back(t.e1);
  }
  g_pending.clear();
}
std::string timing_report() {
  timing_collect();
  std::string out;
  char line[256];
  for (size_t i = 0; i < g_names.size(); ++i) {
    snprintf(line, sizeof(line), "%s %.6f %lld\n", g_names[i].c_str(), g_ms[i], g_cnt[i]);
    out += line;
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// host-side launch wrappers
// ------------------------------------------------------------------------------------------------
static BinParams make_bin_params(const HostParams& hp) {
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
  return bp;
}

static GroundConst make_ground_const(const HostParams& hp) {
  GroundConst gc;
  const double h = (double)hp.p.sensor_height;  // set_sensor(const double&) receives the float param
  const double min_range = 2.7, max_range = 80.0;
  gc.low_thr = -1.8 * h;
  gc.seed_thr = -1.1 * h;
  gc.min_range = min_range;
  gc.max_range = max_range;
  gc.z2 = (7 * min_range + max_range) / 8.0;
  gc.z3 = (3 * min_range + max_range) / 4.0;
  gc.z4 = (min_range + max_range) / 2.0;
  gc.rmin[0] = min_range;
  gc.rmin[1] = gc.z2;
  gc.rmin[2] = gc.z3;
  gc.rmin[3] = gc.z4;
  gc.ring_size[0] = (gc.z2 - min_range) / 2;
  gc.ring_size[1] = (gc.z3 - gc.z2) / 4;
  gc.ring_size[2] = (gc.z4 - gc.z3) / 4;
  gc.ring_size[3] = (max_range - gc.z4) / 4;
  const int sectors[4] = {16, 32, 54, 32};
  for (int k = 0; k < 4; ++k) gc.sector_size[k] = 2 * M_PI / sectors[k];
  return gc;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// grid.x for "per point of a scan" kernels: enough CTAs per scan that nscans * gx covers the GPU a
// few times over, in multiples of the SM count.
static int grid_x_for(int nscans, int per_scan_items, int threads) {
  int want = (per_scan_items + threads - 1) / threads;
  int cap = (num_sms() * 8 + nscans - 1) / nscans;
  if (cap < 1) cap = 1;
  return want < cap ? (want < 1 ? 1 : want) : cap;
}

int launch_ground(const HostParams& hp, BatchDev& d, int nscans, int max_scan_points, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  GroundConst gc = make_ground_const(hp);
  BinParams bp = make_bin_params(hp);
  int launches = 0;
  cudaMemsetAsync(d.patch_cnt, 0, sizeof(int32_t) * (size_t)nscans * kNumPatches, st);
  cudaMemsetAsync(d.sort_ctr, 0, sizeof(int32_t) * 8, st);
  dim3 gpt(grid_x_for(nscans, max_scan_points, 256), nscans);
  { TIMED("k_patch_assign", TSTREAM); k_patch_assign<<<gpt, 256, 0, st>>>(d.pts, d.off, gc, d.patch_of, d.patch_cnt, d.cls); }
  { TIMED("k_patch_scan", TSTREAM); k_patch_scan<<<nscans, 512, 0, st>>>(d.patch_cnt, d.patch_off, d.patch_cur, d.sort_ctr, d.sort_list, d.cap_scans * kNumPatches); }
  { TIMED("k_patch_scatter", TSTREAM); k_patch_scatter<<<gpt, 256, 0, st>>>(d.pts, d.off, d.patch_of, d.patch_off, d.patch_cur, d.bucket_kv); }
  FitArgs fa;
  fa.pts = d.pts;
  fa.off = d.off;
  fa.patch_cnt = d.patch_cnt;
  fa.patch_off = d.patch_off;
  fa.bucket_kv = d.bucket_kv;
  fa.sorted_xyz = d.sorted_xyz;
  fa.sorted_idx = d.sorted_idx;
  fa.slot_pos = d.slot_pos;
  fa.slot_apos = d.slot_apos;
  fa.slot_vid = d.slot_vid;
  fa.slot_patch = d.slot_patch;
  fa.patch_out = d.patch_out;
  fa.patch_plane = d.patch_dbg;
  fa.cls = d.cls;
  fa.err = d.scan_counts + (size_t)d.cap_scans * 8;  // one extra int past the per-scan counters
  fa.gc = gc;
  fa.bp = bp;
  static std::once_flag sort_once;
  std::call_once(sort_once, [] {
    cudaFuncSetAttribute(k_patch_sort_list<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kSortT1 * 8);
  });
  dim3 gfit(kNumPatches, nscans);
  { TIMED("k_patch_sort_1k", TSTREAM); k_patch_sort<128><<<gfit, 128, kSortT0 * 8, st>>>(fa); }
  launches += 1;
  const int list_cap = d.cap_scans * kNumPatches;
  if (max_scan_points > kSortT0) {  // persistent CTAs over the worklists: three 72 KB CTAs per SM
    { TIMED("k_patch_sort_4k", TSTREAM); k_patch_sort_list<256, false><<<num_sms() * 3, 256, 2 * kSortT1 * 8, st>>>(fa, d.sort_list, d.sort_ctr, kSortT1); }
    launches += 1;
  }
  if (max_scan_points > kSortT1) {
    { TIMED("k_patch_sort_16k", TSTREAM); k_patch_sort_list<512, true><<<num_sms() * 2, 512, 0, st>>>(fa, d.sort_list + list_cap, d.sort_ctr + 2, 0); }
    launches += 1;
  }
  if (max_scan_points > kSortT2) {
    { TIMED("k_patch_sort_overflow", TSTREAM); k_patch_sort_list<512, true><<<num_sms() * 2, 512, 0, st>>>(fa, d.sort_list + 2 * (size_t)list_cap, d.sort_ctr + 4, 0); }
    launches += 1;
  }
  {
    // persistent: kChainCtasPerSm CTAs of kChainWarps warps per SM (the chain is issue/latency bound, see the kernel)
    static const int ctas_per_sm = getenv("SCVOD_CHAIN_CTAS") ? std::max(1, atoi(getenv("SCVOD_CHAIN_CTAS"))) : kChainCtasPerSm;  // tuning hook
    TIMED("k_patch_chain", TSTREAM);
    if (hp.chain_tma)
      k_patch_chain<true><<<num_sms() * ctas_per_sm, kChainWarps * 32, 0, st>>>(fa, nscans, d.sort_list, d.sort_ctr, d.cap_scans * kNumPatches);
    else
      k_patch_chain<false><<<num_sms() * ctas_per_sm, kChainWarps * 32, 0, st>>>(fa, nscans, d.sort_list, d.sort_ctr, d.cap_scans * kNumPatches);
  }
  { TIMED("k_patch_rank_small", TSTREAM); k_patch_rank_small<<<dim3((kNumPatches + kRankWarps - 1) / kRankWarps, nscans), kRankWarps * 32, 0, st>>>(fa); }
  if (max_scan_points > kSortT0) {
    { TIMED("k_patch_rank_list", TSTREAM); k_patch_rank_list<256><<<num_sms() * 4, 256, 0, st>>>(fa, d.sort_list, d.sort_ctr, d.cap_scans * kNumPatches); }
    launches += 1;
  }
  launches += 2;
  { TIMED("k_patch_out_scan", TSTREAM); k_patch_out_scan<<<nscans, 512, 0, st>>>(d.patch_cnt, d.patch_out, d.patch_out_off, d.scan_counts); }
  { TIMED("k_emit", TSTREAM); k_emit<<<gpt, 256, 0, st>>>(d.pts, d.off, d.patch_off, d.patch_out, d.patch_out_off, d.sorted_idx, d.slot_pos, d.slot_apos, d.slot_vid,
                             d.slot_patch, d.ground_src, d.ng_src, d.apri_src, d.apri_vid, d.apri_xyzi, d.cls); }
  launches += 5;
  return launches;
}

int launch_descriptor(const HostParams& hp, BatchDev& d, int nscans, int max_scan_points, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  BinParams bp = make_bin_params(hp);
  cudaMemsetAsync(d.bitmap, 0, sizeof(uint32_t) * (size_t)nscans * hp.g.words, st);
  dim3 gpt(grid_x_for(nscans, max_scan_points, 256), nscans);
  { TIMED("k_vox_mark", TSTREAM); k_vox_mark<<<gpt, 256, 0, st>>>(d.off, d.scan_counts, d.apri_vid, hp.g, d.bitmap); }
  { TIMED("k_vox_rank", TSTREAM); k_vox_rank<<<nscans, 1024, 0, st>>>(d.off, hp.g, d.bitmap, d.word_rank, d.vox_vid, d.vox_cnt, d.scan_counts); }
  { TIMED("k_vox_count", TSTREAM); k_vox_count<<<gpt, 256, 0, st>>>(d.off, d.scan_counts, d.apri_vid, hp.g, d.bitmap, d.word_rank, d.apri_cid, d.vox_cnt); }
  { TIMED("k_vox_offsets", TSTREAM); k_vox_offsets<<<nscans, 1024, 0, st>>>(d.off, d.scan_counts, d.vox_cnt, d.vox_off, d.vox_cur); }
  { TIMED("k_vox_fill", TSTREAM); k_vox_fill<<<gpt, 256, 0, st>>>(d.off, d.scan_counts, d.apri_cid, d.vox_off, d.vox_cur, d.vox_pts_tmp); }
  dim3 gv(grid_x_for(nscans, max_scan_points / 4 + 1, 8), nscans);  // one warp per voxel, 8 warps per CTA
  { TIMED("k_vox_stats", TSTREAM); k_vox_stats<<<gv, 256, 0, st>>>(d.off, d.scan_counts, d.vox_cnt, d.vox_off, d.vox_pts_tmp, d.apri_xyzi, d.vox_pts,
                                  d.apri_rank, d.vox_av, d.vox_cov, d.vox_bbox); }
  dim3 gc(grid_x_for(nscans, max_scan_points / 4 + 1, 256), nscans);
  { TIMED("k_vox_center", TSTREAM); k_vox_center<<<gc, 256, 0, st>>>(d.off, d.scan_counts, d.vox_off, d.vox_pts, d.apri_xyzi, bp, hp.p, d.vox_center, d.vox_tri); }
  return 7;
}

int launch_cluster_prep(const HostParams& hp, BatchDev& d, int nscans, int max_scan_points, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  cudaMemsetAsync(d.edge_hash, 0xff, sizeof(unsigned long long) * (size_t)nscans * d.hash_cap, st);
  dim3 gv(grid_x_for(nscans, max_scan_points / 4 + 1, 256), nscans);
  { TIMED("k_vox_nbr", TSTREAM); k_vox_nbr<<<gv, 256, 0, st>>>(d.off, d.scan_counts, hp.g, d.bitmap, d.word_rank, d.vox_tri, d.vox_nbr, d.vox_root); }
  { TIMED("k_ccl_union", TSTREAM); k_ccl_union<<<gv, 256, 0, st>>>(d.off, d.scan_counts, d.vox_nbr, d.vox_root); }
  { TIMED("k_ccl_flatten", TSTREAM); k_ccl_flatten<<<gv, 256, 0, st>>>(d.off, d.scan_counts, d.vox_root); }
  dim3 gw(grid_x_for(nscans, max_scan_points / 4 + 1, 8), nscans);  // one warp per voxel, 8 warps per CTA
  { TIMED("k_similar_edges", TSTREAM); k_similar_edges<<<gw, 256, 0, st>>>(d.off, d.scan_counts, hp.g, d.bitmap, d.word_rank, d.vox_tri, d.vox_av, d.vox_cov, d.vox_root,
                                      hp.p.search_c, hp.p.intensity_cov, hp.p.intensity_diff,
                                      reinterpret_cast<unsigned long long*>(d.edge_hash), d.hash_cap, d.edge_buf, d.edge_cap); }
  { TIMED("k_events", TSTREAM); k_events<<<nscans, 1024, 0, st>>>(d.off, d.scan_counts, d.apri_cid, d.apri_rank, d.ev_cid); }
  return 5;
}

int launch_bin_only(const HostParams& hp, const float4* pts_dev, int n, uint8_t* pass, int32_t* vid, int32_t* ri, int32_t* si,
                    int32_t* ei, float* range, float* angle, float* azimuth, void* stream_) {
  if (n <= 0) return 0;
  int blocks = (n + 255) / 256;
  int cap = num_sms() * 16;
  if (blocks > cap) blocks = cap;
  { TIMED("k_bin_only", TSTREAM); k_bin_only<<<blocks, 256, 0, (cudaStream_t)stream_>>>(pts_dev, n, make_bin_params(hp), pass, vid, ri, si, ei, range, angle, azimuth); }
  return 1;
}

int launch_track(const HostParams& hp, const float4* own_xyzi, const int32_t* vox_off, const int32_t* vox_pts, const float4* carried,
                 const int4* segs, const int32_t* first_seg, int nseg, const TrackRuns* runs, const int32_t* csr_ptoff, const int32_t* csr_vox,
                 const int32_t* csr_part, int k, const float T12[12], const uint32_t* next_bitmap, const int32_t* next_word_rank, int ncl,
                 int vn, float4* out_xyzi, unsigned long long* first, int32_t* ctr_dev, int32_t* hit_list_dev, int32_t* out_quads_mapped,
                 int cap_quads, void* stream_) {
  if (k <= 0 || ncl <= 0 || vn <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  Mat34 T;
  for (int i = 0; i < 12; ++i) T.m[i] = T12[i];
  int blocks = (k + 255) / 256;
  static const int forced = getenv("SCVOD_TRACK_CTAS") ? std::max(1, atoi(getenv("SCVOD_TRACK_CTAS"))) : 0;  // tuning hook
  const int ctas_per_sm = forced ? forced : std::max(1, hp.track_ctas_per_sm);
  int cap = num_sms() * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (runs) {
    TIMED("k_track", TSTREAM);
    k_track<true><<<blocks, 256, 0, st>>>(own_xyzi, vox_off, vox_pts, carried, nullptr, nullptr, 0, *runs, csr_ptoff, csr_vox, csr_part, k, T,
                                           make_bin_params(hp), hp.g, next_bitmap, next_word_rank, vn, out_xyzi, first, ctr_dev, hit_list_dev,
                                           out_quads_mapped, cap_quads);
  } else {
    static const TrackRuns none = {};
    TIMED("k_track_segs", TSTREAM);
    k_track<false><<<blocks, 256, 0, st>>>(own_xyzi, vox_off, vox_pts, carried, segs, first_seg, nseg, none, nullptr, nullptr, nullptr, k, T,
                                            make_bin_params(hp), hp.g, next_bitmap, next_word_rank, vn, out_xyzi, first, ctr_dev, hit_list_dev,
                                            out_quads_mapped, cap_quads);
  }
  return 1;
}

int launch_final_labels(const int64_t* off, const int32_t* scan_counts, int nscans, int max_scan_points, const int32_t* apri_src,
                        const int32_t* apri_cid, const int32_t* vcls_off, const uint8_t* vcls, uint8_t* cls, void* stream_) {
  if (nscans <= 0) return 0;
  dim3 grid(grid_x_for(nscans, max_scan_points, 256), nscans);
  { TIMED("k_final_labels", TSTREAM); k_final_labels<<<grid, 256, 0, (cudaStream_t)stream_>>>(off, scan_counts, apri_src, apri_cid, vcls_off, vcls, cls); }
  return 1;
}

int launch_submap(const float4* pts, const uint8_t* cls, const int64_t* off, const float* Ts_dev, int first_scan, int nscans,
                  int max_scan_points, float4* out, unsigned long long* counter, long long cap, void* stream_) {
  if (nscans <= 0) return 0;
  dim3 grid(grid_x_for(nscans, max_scan_points, 256), nscans);
  { TIMED("k_submap", TSTREAM); k_submap<<<grid, 256, 0, (cudaStream_t)stream_>>>(pts, cls, off, Ts_dev, first_scan, out, counter, cap); }
  return 1;
}

int launch_name_replay(BatchDev& d, int nscans, int max_vox, int max_events, bool force_global, int32_t* vox_name, int32_t* name_first, int name_cap,
                       void* stream_) {
  if (nscans <= 0) return 0;
  constexpr int NW = kReplayWarps;
  const size_t ring = sizeof(int) * NW * 2 * kReplayRows * 32;
  const size_t nfw = ((size_t)max_events + 31) / 32;
  const size_t smem = ring + (size_t)max_vox * 6 + nfw * 8 + 16;
  static std::once_flag replay_once;
  std::call_once(replay_once, [ring] {
    cudaFuncSetAttribute(k_name_replay<NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_name_replay<NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
  });
  int2* ev_list = reinterpret_cast<int2*>(d.bucket_kv);  // the ground stage is done with its (key, index) buckets
  if (smem <= 220 * 1024 && !force_global) {
    { TIMED("k_name_replay", TSTREAM); k_name_replay<NW, false><<<nscans, NW * 32, smem, (cudaStream_t)stream_>>>(d.off, d.scan_counts, d.ev_cid, d.vox_root, d.vox_nbr, ev_list, nullptr, d.vox_pts_tmp, d.apri_rank, nullptr, nullptr, vox_name, name_first, name_cap); }
  } else {  // very dense scans: union-find state in (L2-resident) global scratch that the earlier stages are done with
    { TIMED("k_name_replay_global", TSTREAM); k_name_replay<NW, true><<<nscans, NW * 32, ring, (cudaStream_t)stream_>>>(d.off, d.scan_counts, d.ev_cid, d.vox_root, d.vox_nbr, ev_list, d.vox_cur, d.vox_pts_tmp, d.apri_rank, d.sorted_idx, d.slot_pos, vox_name, name_first, name_cap); }
  }
  return 1;
}

int launch_pack(const PackDesc* descs_dev, int ndesc, int max_n, int32_t* out, void* stream_) {
  if (ndesc <= 0) return 0;
  int gx = (max_n + 255) / 256;
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  { TIMED("k_pack", TSTREAM); k_pack<<<dim3(gx, ndesc), 256, 0, (cudaStream_t)stream_>>>(descs_dev, out); }
  return 1;
}

int launch_atan2f_probe(const float* y, const float* x, float* out, long long n, void* stream_) {
  { TIMED("k_atan2f_probe", TSTREAM); k_atan2f_probe<<<num_sms() * 8, 256, 0, (cudaStream_t)stream_>>>(y, x, out, n); }
  return 1;
}

}  // namespace scvod
