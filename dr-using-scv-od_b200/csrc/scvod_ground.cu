// scvod_ground.cu — ground stage: PatchWork::estimate_ground (reference include/patchwork.h:278-504) with the curved-voxel
// binning of the non-ground points (SSC::makeApriVec, src/ssc.cpp:155-195) fused into its last kernels.
// HBM-, shared-memory- and issue-bound integer / float work: no tensor-core path exists for it.
#include "scvod_kernel_common.cuh"

namespace scvod {

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

struct GroundConst {
  double low_thr;   // -1.8 * sensor_height_  (patchwork.h:304)
  double seed_thr;  // adaptive_seed_selection_margin_ * sensor_height_ (patchwork.h:247)
  double min_range, max_range, z2, z3, z4;
  double ring_size[4], sector_size[4], rmin[4];
  // float filter of k_patch_assign (same idea as dev_bin_filtered): approximate ring / sector coordinates decide unless they
  // are within a guard band of an integer, the exact double chain decides the rest
  float low_thr_f;  // smallest float >= low_thr: ((double)z < low_thr) == (z < low_thr_f)
  float max_range_f, z2f, z3f, z4f, rmin_f[4], inv_ring_f[4], inv_sector_f[4];
};

// pc2czm (patchwork.h:431-459) for one point: patch id, -1 below the height threshold (patchwork.h:304), -2 outside (min_range, max_range]
__device__ __forceinline__ int dev_patch_exact(float px, float py, float pz, const GroundConst& gc) {
  if ((double)pz < gc.low_thr) return -1;
  double x = (double)px, y = (double)py;
  double r = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
  if ((r <= gc.max_range) && (r > gc.min_range)) {
    double theta = (y >= 0) ? atan2(y, x) : __dadd_rn(2.0 * 3.14159265358979323846, atan2(y, x));
    int k = (r < gc.z2) ? 0 : (r < gc.z3) ? 1 : (r < gc.z4) ? 2 : 3;
    int ring = min((int)__ddiv_rn(__dsub_rn(r, gc.rmin[k]), gc.ring_size[k]), c_zone_rings[k] - 1);
    int sector = min((int)__ddiv_rn(theta, gc.sector_size[k]), c_zone_sectors[k] - 1);
    return c_zone_base[k] + ring * c_zone_sectors[k] + sector;
  }
  return -2;
}
static __device__ __noinline__ int dev_patch_exact_call(float px, float py, float pz, const GroundConst& gc) { return dev_patch_exact(px, py, pz, gc); }

// The same result through a float filter (the idea of dev_bin_filtered): the float radius and CUDA's atan2f are within ~1e-5 of the
// exact double ring / sector coordinates (float rounding of a radius <= 80 m over a ring >= 4.8 m; 3 ulp of pi over a sector
// >= 0.116 rad), so coordinates farther than the guard bands (2e-4 / 5e-4: a 20x margin) from every integer decide at once; the
// double chain (dsqrt, double atan2, two double divisions: ~4x the instructions) is evaluated only for the rest.
__device__ __forceinline__ int dev_patch_filtered(float px, float py, float pz, const GroundConst& gc, bool* took_exact = nullptr) {
  if (took_exact) *took_exact = false;
  if (pz < gc.low_thr_f) return -1;
  const float rf = sqrtf(px * px + py * py);
  const int k = (rf < gc.z2f) ? 0 : (rf < gc.z3f) ? 1 : (rf < gc.z4f) ? 2 : 3;
  const float qr = (rf - gc.rmin_f[k]) * gc.inv_ring_f[k];
  float th = atan2f(py, px);
  if (!(py >= 0.f)) th += 6.283185307179586f;
  const float qt = th * gc.inv_sector_f[k];
  const float fr = qr - floorf(qr), ft = qt - floorf(qt);
  if (qr > 2.0e-4f && qr < (float)c_zone_rings[k] - 2.0e-4f && fr > 2.0e-4f && fr < 1.f - 2.0e-4f && qt > 5.0e-4f && ft > 5.0e-4f &&
      ft < 1.f - 5.0e-4f && qt < (float)c_zone_sectors[k] + 0.5f)
    return c_zone_base[k] + (int)qr * c_zone_sectors[k] + min((int)qt, c_zone_sectors[k] - 1);
  if (rf < gc.rmin_f[0] - 1.0e-3f || rf > gc.max_range_f + 1.0e-3f) return -2;  // clearly outside (min_range, max_range]
  if (took_exact) *took_exact = true;
  return dev_patch_exact_call(px, py, pz, gc);
}

// G1 filter soundness probe (scvod_bin_filter_check): stats[5] points that took the double chain, stats[6] mismatches
__global__ void __launch_bounds__(256) k_patch_filter_check(long long n, uint32_t seed, GroundConst gc, float extent,
                                                            unsigned long long* __restrict__ stats) {
  unsigned long long n_exact = 0, n_bad = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t h[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      uint32_t v = (uint32_t)i * 3u + (uint32_t)j + seed + (uint32_t)(i >> 32) * 0x9e3779b9u;
      v ^= v >> 16; v *= 0x7feb352du; v ^= v >> 15; v *= 0x846ca68bu; v ^= v >> 16;
      h[j] = v;
    }
    float x = ((float)h[0] * 2.3283064e-10f - 0.5f) * 2.f * extent;
    float y = ((float)h[1] * 2.3283064e-10f - 0.5f) * 2.f * extent;
    const float z = ((float)h[2] * 2.3283064e-10f - 0.5f) * 8.f;
    const int kind = (int)(i & 7), sub = (int)((i >> 3) & 7);
    if (kind == 0) {
      if (sub == 0) y = 0.f;
      else if (sub == 1) x = 0.f;
      else if (sub == 2) y = -0.f;
      else if (sub == 3) { x = 0.f; y = 0.f; }
      else if (sub == 4) {  // on a sector edge of some zone
        const int zn = (int)(h[2] & 3);
        const float r = sqrtf(x * x + y * y) + 0.1f, a = (float)(h[1] % (uint32_t)c_zone_sectors[zn]) * (float)gc.sector_size[zn];
        x = r * cosf(a); y = r * sinf(a);
      } else if (sub == 5) {  // on a ring edge
        const int zn = (int)(h[2] & 3);
        const float r = fmaxf(sqrtf(x * x + y * y), 1e-6f), rr = (float)(gc.rmin[zn] + (double)(h[1] % 5u) * gc.ring_size[zn]);
        x *= rr / r; y *= rr / r;
      } else if (sub == 6) y = -fabsf(y) * 1e-7f;
    }
    bool slow;
    const int f = dev_patch_filtered(x, y, z, gc, &slow);
    const int e = dev_patch_exact(x, y, z, gc);
    n_exact += slow ? 1 : 0;
    n_bad += (f != e) ? 1 : 0;
  }
  atomicAdd(&stats[5], n_exact);
  atomicAdd(&stats[6], n_bad);
}

// ------------------------------------------------------------------------------------------------
// G1: per-point patch assignment (pc2czm, patchwork.h:431-459) + per-patch histogram
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_patch_assign(const float4* __restrict__ pts, const int64_t* __restrict__ off,
                                                      GroundConst gc, int16_t* __restrict__ patch_of, uint32_t* __restrict__ zkey,
                                                      int32_t* __restrict__ patch_cnt, uint8_t* __restrict__ cls) {
  __shared__ int s_hist[kNumPatches];
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int n = (int)(off[b + 1] - base);
  for (int i = threadIdx.x; i < kNumPatches; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&pts[base + i]);
    const int pid = dev_patch_filtered(p.x, p.y, p.z, gc);
    patch_of[base + i] = (int16_t)pid;
    zkey[base + i] = float_sort_key(p.z);  // the scatter pass needs nothing else of the point: 4 B instead of a 16-byte re-read
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
__global__ void __launch_bounds__(256) k_patch_scatter(const uint32_t* __restrict__ zkey, const int64_t* __restrict__ off,
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
  // Latency bound: every thread keeps kU independent loads in flight (0.057 -> 0.046 ms per 64 scans).  Warp-aggregated counters
  // were measured too (one atomic per warp and key): slower, 0.053 ms - same-address shared-memory atomics are cheap on sm_100.
  constexpr int kU = 4;
  for (int ib = i0 + threadIdx.x; ib < i1; ib += kU * blockDim.x) {
    int pid[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = ib + u * blockDim.x;
      pid[u] = (i < i1) ? (int)patch_of[base + i] : -1;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (pid[u] >= 0) atomicAdd(&s_cnt[pid[u]], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kNumPatches; i += blockDim.x) {
    const int c = s_cnt[i];
    s_base[i] = c ? patch_off[b * (kNumPatches + 1) + i] + atomicAdd(&patch_cur[b * kNumPatches + i], c) : 0;
    s_cnt[i] = 0;
  }
  __syncthreads();
  for (int ib = i0 + threadIdx.x; ib < i1; ib += kU * blockDim.x) {
    int pid[kU];
    uint32_t zk[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = ib + u * blockDim.x;
      pid[u] = (i < i1) ? (int)patch_of[base + i] : -1;
      zk[u] = (i < i1) ? __ldg(&zkey[base + i]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (pid[u] < 0) continue;
      const int i = ib + u * blockDim.x;
      const int slot = s_base[pid[u]] + atomicAdd(&s_cnt[pid[u]], 1);
      bucket_kv[base + slot] = ((uint64_t)zk[u] << 32) | (uint32_t)i;
    }
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
  // gather of the points in sorted order (xyz for the fit, the intensity rides along to k_emit): four independent 16-byte gathers
  // in flight per thread
  for (int j0 = tid; j0 < n; j0 += 4 * THREADS) {
    int idx[4];
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) idx[u] = (j0 + u * THREADS < n) ? (int)(uint32_t)kv[j0 + u * THREADS] : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(&a.pts[base + idx[u]]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * THREADS;
      if (j >= n) break;
      a.sorted_xyz[base + slot0 + j] = q[u];
      a.sorted_idx[base + slot0 + j] = idx[u];
      a.slot_patch[base + slot0 + j] = (int16_t)p;
    }
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
      const int d = valid ? (int)((kv >> shift) & 255u) : 0;
      // lanes holding the same digit: eight ballots (one per digit bit) instead of __match_any_sync, whose result was the
      // hottest stall of this kernel in round 1 (profiles/r01_source_hotspots.txt)
      unsigned same = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int bit = 0; bit < 8; ++bit) {
        const unsigned bal = __ballot_sync(0xffffffffu, (d >> bit) & 1);
        same &= ((d >> bit) & 1) ? bal : ~bal;
      }
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
  for (int j0 = tid; j0 < n; j0 += 4 * THREADS) {
    int idx[4];
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) idx[u] = (j0 + u * THREADS < n) ? (int)(uint32_t)bufA[j0 + u * THREADS] : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(&a.pts[base + idx[u]]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * THREADS;
      if (j >= n) break;
      a.sorted_xyz[base + slot0 + j] = q[u];
      a.sorted_idx[base + slot0 + j] = idx[u];
      a.slot_patch[base + slot0 + j] = (int16_t)p;
    }
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
constexpr int kAposQuirk = 1 << 30;

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
        const BinIdx r = dev_bin_filtered(q.x, q.y, q.z, a.bp);
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
        if (ps) apos = (g ? rGP : rNP) | ((f & F_QUIRK) ? kAposQuirk : 0);  // the -1-index flag rides along to k_emit
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
__global__ void __launch_bounds__(256) k_emit(const float4* __restrict__ sorted_xyzi, const int64_t* __restrict__ off,
                                              const int32_t* __restrict__ patch_off, const int32_t* __restrict__ patch_out,
                                              const int32_t* __restrict__ patch_out_off, const int32_t* __restrict__ sorted_idx,
                                              const int32_t* __restrict__ slot_pos, const int32_t* __restrict__ slot_apos,
                                              const int32_t* __restrict__ slot_vid, const int16_t* __restrict__ slot_patch,
                                              int32_t* __restrict__ ground_src, int32_t* __restrict__ ng_src,
                                              int32_t* __restrict__ apri_src, int32_t* __restrict__ apri_vid,
                                              float4* __restrict__ apri_xyzi, uint8_t* __restrict__ cls,
                                              int32_t* __restrict__ taint_cnt, int32_t* __restrict__ q_list) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int nslots = patch_off[b * (kNumPatches + 1) + kNumPatches];
  // Latency bound (slot record -> patch tables -> scattered writes): every thread carries kU slots through the levels together,
  // all first-level loads (five small arrays) issued before anything is consumed (0.098 -> 0.084 ms per 64 scans).
  constexpr int kU = 2;
  const int stride = gridDim.x * blockDim.x;
  for (int q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < nslots; q0 += kU * stride) {
    int sp[kU], pp[kU], idx[kU], ar_raw[kU], vid[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = q0 + u * stride;
      const bool live = q < nslots;
      sp[u] = live ? __ldg(&slot_pos[base + q]) : (3 << 30);
      pp[u] = live ? (int)__ldg(&slot_patch[base + q]) : 0;
      idx[u] = live ? __ldg(&sorted_idx[base + q]) : 0;
      ar_raw[u] = live ? __ldg(&slot_apos[base + q]) : -1;
      vid[u] = live ? __ldg(&slot_vid[base + q]) : 0;
    }
    float4 xyzi[kU];
    int o0[kU], o1[kU], o2[kU], po4[kU], po5[kU], po6[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int role = (sp[u] >> 30) & 3;
      const int32_t* o = patch_out_off + (b * (kNumPatches + 1) + pp[u]) * 3;
      const int32_t* po = patch_out + (b * kNumPatches + pp[u]) * kPatchOutStride;
      o0[u] = __ldg(&o[0]);
      o1[u] = __ldg(&o[1]);
      o2[u] = __ldg(&o[2]);
      po4[u] = __ldg(&po[4]);
      po5[u] = __ldg(&po[5]);
      po6[u] = __ldg(&po[6]);
      // the z-sorted copy of the point (coalesced), not a gather of the input
      xyzi[u] = (role == 2) ? __ldg(&sorted_xyzi[base + q0 + u * stride]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int role = (sp[u] >> 30) & 3;
      if (role == 3) continue;
      const bool g = (sp[u] >> 29) & 1;
      const int rank = sp[u] & 0x1fffffff;
      if (role == 0) {
        ground_src[base + o0[u] + rank] = idx[u];
        cls[base + idx[u]] = SCVOD_PT_GROUND;
      } else {
        // cloud_nonground: the complement of the ground set, or [ground set][complement] for a rejected patch
        // (patchwork.h:348-349,373-374)
        const bool rejected = po6[u] != 0;
        const int pos = (rejected && !g) ? po4[u] + rank : rank;
        ng_src[base + o1[u] + pos] = idx[u];
        if (role == 1) {
          cls[base + idx[u]] = SCVOD_PT_GATED_OUT;
        } else {
          const int ar = ar_raw[u] & (kAposQuirk - 1);
          const int m = o2[u] + ((rejected && !g) ? po5[u] + ar : ar);
          if (ar_raw[u] & kAposQuirk) {  // a point with a -1 index: its voxel will be named point by point (k_taint_*)
            const int k = atomicAdd(&taint_cnt[b * kTaintCntStride], 1);
            if (k < kQuirkCap) q_list[(size_t)b * kQuirkCap + k] = m;
          }
          apri_src[base + m] = idx[u];
          apri_vid[base + m] = vid[u];
          apri_xyzi[base + m] = xyzi[u];
          cls[base + idx[u]] = SCVOD_PT_UNCLUSTERED;
        }
      }
    }
  }
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
  gc.low_thr_f = nextafterf((float)gc.low_thr, INFINITY);
  while ((double)nextafterf(gc.low_thr_f, -INFINITY) >= gc.low_thr) gc.low_thr_f = nextafterf(gc.low_thr_f, -INFINITY);
  gc.max_range_f = (float)gc.max_range;
  gc.z2f = (float)gc.z2;
  gc.z3f = (float)gc.z3;
  gc.z4f = (float)gc.z4;
  for (int k = 0; k < 4; ++k) {
    gc.rmin_f[k] = (float)gc.rmin[k];
    gc.inv_ring_f[k] = (float)(1.0 / gc.ring_size[k]);
    gc.inv_sector_f[k] = (float)(1.0 / gc.sector_size[k]);
  }
  return gc;
}

int launch_patch_filter_check(const HostParams& hp, long long n, uint32_t seed, float extent, unsigned long long* stats_dev, void* stream_) {
  { TIMED("k_patch_filter_check", TSTREAM); k_patch_filter_check<<<num_sms() * 8, 256, 0, (cudaStream_t)stream_>>>(n, seed, make_ground_const(hp), extent, stats_dev); }
  return 1;
}

int launch_ground(const HostParams& hp, BatchDev& d, int nscans, int max_scan_points, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  GroundConst gc = make_ground_const(hp);
  BinParams bp = make_bin_params(hp);
  int launches = 0;
  cudaMemsetAsync(d.patch_cnt, 0, sizeof(int32_t) * (size_t)nscans * kNumPatches, st);
  cudaMemsetAsync(d.sort_ctr, 0, sizeof(int32_t) * 8, st);
  cudaMemsetAsync(d.taint_cnt, 0, sizeof(int32_t) * (size_t)nscans * kTaintCntStride, st);
  dim3 gpt(grid_x_for(nscans, max_scan_points, 256), nscans);
  { TIMED("k_patch_assign", TSTREAM); k_patch_assign<<<gpt, 256, 0, st>>>(d.pts, d.off, gc, d.patch_of, reinterpret_cast<uint32_t*>(d.slot_vid), d.patch_cnt, d.cls); }  // z keys live in slot_vid until the rank kernels overwrite it
  { TIMED("k_patch_scan", TSTREAM); k_patch_scan<<<nscans, 512, 0, st>>>(d.patch_cnt, d.patch_off, d.patch_cur, d.sort_ctr, d.sort_list, d.cap_scans * kNumPatches); }
  { TIMED("k_patch_scatter", TSTREAM); k_patch_scatter<<<gpt, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(d.slot_vid), d.off, d.patch_of, d.patch_off, d.patch_cur, d.bucket_kv); }
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
  { TIMED("k_emit", TSTREAM); k_emit<<<gpt, 256, 0, st>>>(d.sorted_xyz, d.off, d.patch_off, d.patch_out, d.patch_out_off, d.sorted_idx, d.slot_pos, d.slot_apos, d.slot_vid,
                             d.slot_patch, d.ground_src, d.ng_src, d.apri_src, d.apri_vid, d.apri_xyzi, d.cls, d.taint_cnt, d.q_list); }
  launches += 5;
  return launches;
}

}  // namespace scvod
