// scvod_gicp.cu — GICP scan-to-map stage (docs/gicp_spec.md): uniform-grid build, shared-memory
// fixed-radius search (per-point normals), per-iteration correspondence search + Gauss-Newton
// accumulation with warp-shuffle reduction, and the extern "C" entry points declared in include/scvod.h.
//
// The reference names this stage but holds no code for it (src/gicp.cpp:1-57 is a PCD merge tool;
// src/ssc.cpp:1458,1467 are commented-out TODOs), so the spec file is the definition and the CPU oracle
// (oracle/gicp_oracle.cpp) is the checker.  Section numbers in comments refer to the spec.
//
// Kernel design (B200): points are stored cell-major, so the 27-cell neighbourhood of a query cell is 9
// contiguous runs.  One warp owns a "group" = up to 32 queries of one cell; it streams the 9 runs through
// a private shared-memory tile (coalesced float4 loads, then broadcast LDS.128 in the inner loop, one
// candidate per iteration for all 32 queries), so every candidate is fetched from L2 once per group
// instead of once per query.  The Gauss-Newton sums are reduced in 64-bit fixed point: warp shuffles,
// shared-memory atomics per CTA, 29 global atomics per CTA — exact, hence order independent.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "scvod_internal.h"

namespace scvod {
namespace {

#define GCU(call)                                                                                                \
  do {                                                                                                           \
    cudaError_t e__ = (call);                                                                                    \
    if (e__ != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

struct GGrid {
  float ox, oy, oz, h;
  int nx, ny, nz, ncells;
  int rings;  // ceil(max_corr_dist / h): cell rings a correspondence search may have to visit (spec §2)
};

constexpr int kWarpsPerCta = 8;
constexpr int kTile = 128;        // candidates staged per warp and pass
constexpr int kNumAcc = 29;       // 21 H (upper triangle) + 6 g + cost + count
constexpr double kScaleH = 65536.0;        // 2^16
constexpr double kScaleG = 16777216.0;     // 2^24
constexpr int kMaxCells = 1 << 22;

template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    count += count / 4 + 256;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(float v, float o, float h) {  // spec §2, float IEEE sub/div/floor
  return (int)floorf(__fdiv_rn(__fsub_rn(v, o), h));
}
__device__ __forceinline__ int cell_linear(const GGrid& g, float x, float y, float z) {
  int cx = cell_coord(x, g.ox, g.h), cy = cell_coord(y, g.oy, g.h), cz = cell_coord(z, g.oz, g.h);
  if (cx < 0 || cy < 0 || cz < 0 || cx >= g.nx || cy >= g.ny || cz >= g.nz) return -1;
  return (cx * g.ny + cy) * g.nz + cz;
}

__device__ __forceinline__ int warp_incl_scan_i(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan (1024 threads)
__device__ __forceinline__ int block_excl_scan_1024(int v, int* total, int* s_w /* 33 ints */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_incl_scan_i(v);
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = s_w[lane];
    int xi = warp_incl_scan_i(x);
    s_w[lane] = xi - x;
    if (lane == 31) s_w[32] = xi;
  }
  __syncthreads();
  int res = inc - v + s_w[w];
  *total = s_w[32];
  __syncthreads();
  return res;
}

// ------------------------------------------------------------------------------------------------
// bounding box (ordered-int atomics)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__global__ void __launch_bounds__(256) k_gicp_bbox(const float4* __restrict__ pts, int n, int* __restrict__ box /* lo[3], hi[3] */) {
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&pts[i]);
    int a = f2ord(p.x), b = f2ord(p.y), c = f2ord(p.z);
    lo[0] = min(lo[0], a); hi[0] = max(hi[0], a);
    lo[1] = min(lo[1], b); hi[1] = max(hi[1], b);
    lo[2] = min(lo[2], c); hi[2] = max(hi[2], c);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&box[d], lo[d]);
      atomicMax(&box[3 + d], hi[d]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// grid build: count (the atomic's return value is the point's slot inside its cell), scan, place
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gicp_count(const float4* __restrict__ pts, int n, GGrid g, int* __restrict__ cell_of,
                                                    int* __restrict__ slot_in_cell, int* __restrict__ cell_cnt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&pts[i]);
    int c = cell_linear(g, p.x, p.y, p.z);
    cell_of[i] = c;
    slot_in_cell[i] = (c >= 0) ? atomicAdd(&cell_cnt[c], 1) : -1;
  }
}

// exclusive scan of cell_cnt[0..L) into start[0..L]; three passes (L <= 2^22 -> <= 4096 blocks)
__global__ void __launch_bounds__(1024) k_gicp_scan_blocks(const int* __restrict__ cnt, int L, int* __restrict__ start,
                                                           int* __restrict__ block_sum) {
  __shared__ int s_w[33];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  int v = (i < L) ? cnt[i] : 0;
  int total;
  int ex = block_excl_scan_1024(v, &total, s_w);
  if (i < L) start[i] = ex;
  if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_gicp_scan_sums(int* __restrict__ block_sum, int nblocks, int* __restrict__ start, int L,
                                                         int* __restrict__ counters /* zeroed here: group counter */) {
  __shared__ int s_w[33];
  int carry = 0;
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    int b = b0 + threadIdx.x;
    int v = (b < nblocks) ? block_sum[b] : 0;
    int total;
    int ex = block_excl_scan_1024(v, &total, s_w);
    if (b < nblocks) block_sum[b] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    start[L] = carry;
    counters[0] = 0;
  }
}
__global__ void __launch_bounds__(1024) k_gicp_scan_add(int* __restrict__ start, int L, const int* __restrict__ block_sum) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i < L) start[i] += block_sum[blockIdx.x];
}

// Deterministic placement (target / covariance side): the position of a point inside its cell is its rank by
// original index (spec §2), found by counting the cell's smaller indices in the unordered fill.
__global__ void __launch_bounds__(256) k_gicp_fill_tmp(int n, const int* __restrict__ cell_of, const int* __restrict__ slot_in_cell,
                                                       const int* __restrict__ start, int* __restrict__ tmp) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int c = cell_of[i];
    if (c >= 0) tmp[start[c] + slot_in_cell[i]] = i;
  }
}
__global__ void __launch_bounds__(256) k_gicp_place_ranked(const float4* __restrict__ pts, int n, const int* __restrict__ cell_of,
                                                           const int* __restrict__ start, const int* __restrict__ tmp,
                                                           float4* __restrict__ sorted, int* __restrict__ sorted_cell,
                                                           int* __restrict__ groups, int* __restrict__ counters) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int c = cell_of[i];
    if (c < 0) continue;
    const int s0 = start[c], s1 = start[c + 1];
    int rank = 0;
    for (int t = s0; t < s1; ++t) rank += (tmp[t] < i) ? 1 : 0;
    float4 p = __ldg(&pts[i]);
    p.w = __int_as_float(i);
    sorted[s0 + rank] = p;
    sorted_cell[s0 + rank] = c;
    if ((rank & 31) == 0) groups[atomicAdd(&counters[0], 1)] = s0 + rank;
  }
}

// ------------------------------------------------------------------------------------------------
// §3: radius search, normals.  One warp per group (<= 32 queries of one cell).
// ------------------------------------------------------------------------------------------------
__device__ void jacobi3_smallest(double A[3][3], double& l0, double& l1, double& l2, double nrm[3]) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-30 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  double w[3] = {A[0][0], A[1][1], A[2][2]};
  int imin = 0;
  if (w[1] < w[imin]) imin = 1;
  if (w[2] < w[imin]) imin = 2;
  int ia = (imin + 1) % 3, ib = (imin + 2) % 3;
  l2 = w[imin];
  l0 = fmax(w[ia], w[ib]);
  l1 = fmin(w[ia], w[ib]);
  nrm[0] = V[0][imin];
  nrm[1] = V[1][imin];
  nrm[2] = V[2][imin];
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) k_gicp_normals(const float4* __restrict__ sorted, const int* __restrict__ sorted_cell,
                                                                    const int* __restrict__ start, const int* __restrict__ groups,
                                                                    const int* __restrict__ counters, GGrid g, float r2, int min_neighbors,
                                                                    float planarity, float4* __restrict__ normal /* sorted order */,
                                                                    int* __restrict__ count /* sorted order */) {
  __shared__ float4 s_tile[kWarpsPerCta][kTile];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ngroups = counters[0];
  float4* tile = s_tile[wid];
  for (int grp = blockIdx.x * kWarpsPerCta + wid; grp < ngroups; grp += gridDim.x * kWarpsPerCta) {
    const int s0 = groups[grp];
    const int c = sorted_cell[s0];
    const int nq = min(32, start[c + 1] - s0);
    const bool active = lane < nq;
    const float4 q = sorted[s0 + (active ? lane : 0)];
    const int cz = c % g.nz, cy = (c / g.nz) % g.ny, cx = c / (g.nz * g.ny);
    int k = 0;
    float s1x = 0.f, s1y = 0.f, s1z = 0.f, sxx = 0.f, sxy = 0.f, sxz = 0.f, syy = 0.f, syz = 0.f, szz = 0.f;
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = cx + dx;
      if (x < 0 || x >= g.nx) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = cy + dy;
        if (y < 0 || y >= g.ny) continue;
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.nz - 1);
        const int lo = start[(x * g.ny + y) * g.nz + z0], hi = start[(x * g.ny + y) * g.nz + z1 + 1];
        for (int base = lo; base < hi; base += kTile) {
          const int m = min(kTile, hi - base);
          for (int t = lane; t < m; t += 32) tile[t] = __ldg(&sorted[base + t]);
          __syncwarp();
#pragma unroll 4
          for (int t = 0; t < m; ++t) {
            const float4 cnd = tile[t];
            const float d0 = cnd.x - q.x, d1 = cnd.y - q.y, d2 = cnd.z - q.z;
            const float dd = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
            if (dd <= r2) {
              ++k;
              s1x += d0; s1y += d1; s1z += d2;
              sxx += d0 * d0; sxy += d0 * d1; sxz += d0 * d2; syy += d1 * d1; syz += d1 * d2; szz += d2 * d2;
            }
          }
          __syncwarp();
        }
      }
    }
    if (active) {
      float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k >= min_neighbors) {
        const double inv = 1.0 / (double)k, m0 = s1x * inv, m1 = s1y * inv, m2 = s1z * inv;
        double C[3][3];
        C[0][0] = sxx * inv - m0 * m0;
        C[0][1] = C[1][0] = sxy * inv - m0 * m1;
        C[0][2] = C[2][0] = sxz * inv - m0 * m2;
        C[1][1] = syy * inv - m1 * m1;
        C[1][2] = C[2][1] = syz * inv - m1 * m2;
        C[2][2] = szz * inv - m2 * m2;
        double l0, l1, l2, nr[3];
        jacobi3_smallest(C, l0, l1, l2, nr);
        if (l1 > 1e-10 && l2 <= (double)planarity * l1) out = make_float4((float)nr[0], (float)nr[1], (float)nr[2], 1.f);
      }
      normal[s0 + lane] = out;
      count[s0 + lane] = k;
    }
  }
}

// scatter of the sorted-order results back to input order (scvod_gicp_normals)
__global__ void __launch_bounds__(256) k_gicp_unsort(const float4* __restrict__ sorted, const float4* __restrict__ normal,
                                                     const int* __restrict__ count, int n, float* __restrict__ normals3,
                                                     uint8_t* __restrict__ valid, int* __restrict__ count_out) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const int i = __float_as_int(sorted[s].w);
    const float4 nm = normal[s];
    normals3[3 * i] = nm.x;
    normals3[3 * i + 1] = nm.y;
    normals3[3 * i + 2] = nm.z;
    valid[i] = nm.w != 0.f ? 1 : 0;
    count_out[i] = count[s];
  }
}

// ------------------------------------------------------------------------------------------------
// §4 per iteration: transform + bin the valid source points into the TARGET grid, place them cell-major
// (order inside a cell is irrelevant: the reduction is exact), then search + accumulate.
// ------------------------------------------------------------------------------------------------
struct Rt {
  float r[9], t[3];
};

__global__ void __launch_bounds__(256) k_gicp_src_count(const float4* __restrict__ src_sorted, const float4* __restrict__ src_normal, int n,
                                                        Rt T, GGrid g, float4* __restrict__ moved, int* __restrict__ cell_of,
                                                        int* __restrict__ slot_in_cell, int* __restrict__ cell_cnt,
                                                        unsigned long long* __restrict__ acc) {
  if (blockIdx.x == 0 && threadIdx.x < kNumAcc) acc[threadIdx.x] = 0ull;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    int c = -1;
    if (src_normal[s].w != 0.f) {
      const float4 a = __ldg(&src_sorted[s]);
      float4 p;  // p = R a + t, float, left to right (spec §4)
      p.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.r[0], a.x), __fmul_rn(T.r[1], a.y)), __fmul_rn(T.r[2], a.z)), T.t[0]);
      p.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.r[3], a.x), __fmul_rn(T.r[4], a.y)), __fmul_rn(T.r[5], a.z)), T.t[1]);
      p.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.r[6], a.x), __fmul_rn(T.r[7], a.y)), __fmul_rn(T.r[8], a.z)), T.t[2]);
      p.w = __int_as_float(s);
      moved[s] = p;
      c = cell_linear(g, p.x, p.y, p.z);
    }
    cell_of[s] = c;
    slot_in_cell[s] = (c >= 0) ? atomicAdd(&cell_cnt[c], 1) : -1;
  }
}

__global__ void __launch_bounds__(256) k_gicp_src_place(const float4* __restrict__ moved, int n, const int* __restrict__ cell_of,
                                                        const int* __restrict__ slot_in_cell, const int* __restrict__ start,
                                                        float4* __restrict__ q_sorted, int* __restrict__ q_cell, int* __restrict__ groups,
                                                        int* __restrict__ counters) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const int c = cell_of[s];
    if (c < 0) continue;
    const int k = slot_in_cell[s];
    const int pos = start[c] + k;
    q_sorted[pos] = moved[s];
    q_cell[pos] = c;
    if ((k & 31) == 0) groups[atomicAdd(&counters[0], 1)] = pos;
  }
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) k_gicp_corr(const float4* __restrict__ q_sorted, const int* __restrict__ q_cell,
                                                                 const int* __restrict__ q_start, const int* __restrict__ groups,
                                                                 const int* __restrict__ counters, const float4* __restrict__ src_normal,
                                                                 const float4* __restrict__ tgt_sorted, const float4* __restrict__ tgt_normal,
                                                                 const int* __restrict__ tgt_start, GGrid g, Rt T, float dmax2, float k1,
                                                                 unsigned long long* __restrict__ acc) {
  __shared__ float4 s_tile[kWarpsPerCta][kTile];
  __shared__ unsigned long long s_acc[kNumAcc];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < kNumAcc) s_acc[threadIdx.x] = 0ull;
  __syncthreads();
  const int ngroups = counters[0];
  float4* tile = s_tile[wid];
  long long my_acc = 0;  // lane v of the warp accumulates value v of the 29 sums
  for (int grp = blockIdx.x * kWarpsPerCta + wid; grp < ngroups; grp += gridDim.x * kWarpsPerCta) {
    const int s0 = groups[grp];
    const int c = q_cell[s0];
    const int nq = min(32, q_start[c + 1] - s0);
    const bool active = lane < nq;
    const float4 q = q_sorted[s0 + (active ? lane : 0)];
    const int cz = c % g.nz, cy = (c / g.nz) % g.ny, cx = c / (g.nz * g.ny);
    float best = 3.0e38f;
    int best_pos = -1, best_idx = 0x7fffffff;
    // one contiguous run of cells (x, y, z0..z1) streamed through the warp's tile
    auto scan_run = [&](int x, int y, int z0, int z1) {
      z0 = max(z0, 0);
      z1 = min(z1, g.nz - 1);
      if (x < 0 || x >= g.nx || y < 0 || y >= g.ny || z0 > z1) return;
      const int lo = tgt_start[(x * g.ny + y) * g.nz + z0], hi = tgt_start[(x * g.ny + y) * g.nz + z1 + 1];
      for (int base = lo; base < hi; base += kTile) {
        const int m = min(kTile, hi - base);
        for (int t = lane; t < m; t += 32) {
          float4 cnd = __ldg(&tgt_sorted[base + t]);
          // an invalid target point can never be a correspondence: poison its coordinates once, at staging time
          if (__ldg(&tgt_normal[base + t]).w == 0.f) cnd.x = 3.0e18f;
          tile[t] = cnd;
        }
        __syncwarp();
#pragma unroll 4
        for (int t = 0; t < m; ++t) {
          const float4 cnd = tile[t];
          const float d0 = q.x - cnd.x, d1 = q.y - cnd.y, d2 = q.z - cnd.z;
          const float dd = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
          const int idx = __float_as_int(cnd.w);
          if (dd < best || (dd == best && idx < best_idx)) {
            best = dd;
            best_pos = base + t;
            best_idx = idx;
          }
        }
        __syncwarp();
      }
    };
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy) scan_run(cx + dx, cy + dy, cz - 1, cz + 1);
    // Every unvisited point is at least one cell edge away from every query of this cell, so the first ring is final
    // for a query whose best distance is already <= h; only groups with a worse query look further (rare once the
    // alignment has started to converge).
    const int K = g.rings;
    if (K > 1 && __any_sync(0xffffffffu, active && best > g.h * g.h)) {
      for (int dx = -K; dx <= K; ++dx)
        for (int dy = -K; dy <= K; ++dy) {
          if (dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1) {
            scan_run(cx + dx, cy + dy, cz - K, cz - 2);
            scan_run(cx + dx, cy + dy, cz + 2, cz + K);
          } else {
            scan_run(cx + dx, cy + dy, cz - K, cz + K);
          }
        }
    }
    // per-correspondence terms in double (once per query and iteration: negligible next to the search)
    double v[kNumAcc];
#pragma unroll
    for (int i = 0; i < kNumAcc; ++i) v[i] = 0.0;
    if (active && best_pos >= 0 && best <= dmax2) {
      const float4 b = tgt_sorted[best_pos];
      const float4 nbf = tgt_normal[best_pos];
      const float4 naf = src_normal[__float_as_int(q.w)];
      const double p0 = q.x, p1 = q.y, p2 = q.z;
      const double e0 = p0 - (double)b.x, e1 = p1 - (double)b.y, e2 = p2 - (double)b.z;
      const double nb0 = nbf.x, nb1 = nbf.y, nb2 = nbf.z;
      const double m0 = (double)T.r[0] * naf.x + (double)T.r[1] * naf.y + (double)T.r[2] * naf.z;
      const double m1 = (double)T.r[3] * naf.x + (double)T.r[4] * naf.y + (double)T.r[5] * naf.z;
      const double m2 = (double)T.r[6] * naf.x + (double)T.r[7] * naf.y + (double)T.r[8] * naf.z;
      const double kk = (double)k1;
      const double S00 = 2.0 - kk * (nb0 * nb0 + m0 * m0), S01 = -kk * (nb0 * nb1 + m0 * m1), S02 = -kk * (nb0 * nb2 + m0 * m2);
      const double S11 = 2.0 - kk * (nb1 * nb1 + m1 * m1), S12 = -kk * (nb1 * nb2 + m1 * m2), S22 = 2.0 - kk * (nb2 * nb2 + m2 * m2);
      const double c00 = S11 * S22 - S12 * S12, c01 = S12 * S02 - S01 * S22, c02 = S01 * S12 - S11 * S02;
      const double det = S00 * c00 + S01 * c01 + S02 * c02;
      if (fabs(det) > 1e-300) {
        const double id = 1.0 / det;
        const double M00 = c00 * id, M01 = c01 * id, M02 = c02 * id;
        const double M11 = (S00 * S22 - S02 * S02) * id, M12 = (S02 * S01 - S00 * S12) * id, M22 = (S00 * S11 - S01 * S01) * id;
        // B = P M with P = [p]x
        const double B00 = -p2 * M01 + p1 * M02, B01 = -p2 * M11 + p1 * M12, B02 = -p2 * M12 + p1 * M22;
        const double B10 = p2 * M00 - p0 * M02, B11 = p2 * M01 - p0 * M12, B12 = p2 * M02 - p0 * M22;
        const double B20 = -p1 * M00 + p0 * M01, B21 = -p1 * M01 + p0 * M11, B22 = -p1 * M02 + p0 * M12;
        // Hww = B P^T : row r of B dotted with rows of P
        //   P rows: [0,-p2,p1], [p2,0,-p0], [-p1,p0,0]
        const double W00 = -B01 * p2 + B02 * p1, W01 = B00 * p2 - B02 * p0, W02 = -B00 * p1 + B01 * p0;
        const double W11 = B10 * p2 - B12 * p0, W12 = -B10 * p1 + B11 * p0;
        const double W22 = -B20 * p1 + B21 * p0;
        const double Me0 = M00 * e0 + M01 * e1 + M02 * e2, Me1 = M01 * e0 + M11 * e1 + M12 * e2, Me2 = M02 * e0 + M12 * e1 + M22 * e2;
        // upper triangle of H, row-major: (0,0..5), (1,1..5), (2,2..5), (3,3..5), (4,4..5), (5,5)
        v[0] = W00; v[1] = W01; v[2] = W02; v[3] = B00; v[4] = B01; v[5] = B02;
        v[6] = W11; v[7] = W12; v[8] = B10; v[9] = B11; v[10] = B12;
        v[11] = W22; v[12] = B20; v[13] = B21; v[14] = B22;
        v[15] = M00; v[16] = M01; v[17] = M02;
        v[18] = M11; v[19] = M12;
        v[20] = M22;
        v[21] = p1 * Me2 - p2 * Me1; v[22] = p2 * Me0 - p0 * Me2; v[23] = p0 * Me1 - p1 * Me0;
        v[24] = Me0; v[25] = Me1; v[26] = Me2;
        v[27] = e0 * Me0 + e1 * Me1 + e2 * Me2;
        v[28] = 1.0 / kScaleG;  // count: one unit after scaling
      }
    }
#pragma unroll
    for (int i = 0; i < kNumAcc; ++i) {
      long long f = __double2ll_rn(v[i] * (i < 21 ? kScaleH : kScaleG));
      f = warp_sum_ll(f);
      if (lane == i) my_acc += f;
    }
  }
  if (lane < kNumAcc && my_acc != 0) atomicAdd(&s_acc[lane], (unsigned long long)my_acc);
  __syncthreads();
  if (threadIdx.x < kNumAcc && s_acc[threadIdx.x] != 0ull) atomicAdd(&acc[threadIdx.x], s_acc[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct GCloud {
  int n = 0;
  GGrid g{};
  DBuf<float4> pts, sorted, normal;
  DBuf<int> cell_of, slot_in_cell, cell_cnt, start, block_sum, tmp, sorted_cell, groups, count;
  void release() {
    pts.release(); sorted.release(); normal.release(); cell_of.release(); slot_in_cell.release(); cell_cnt.release();
    start.release(); block_sum.release(); tmp.release(); sorted_cell.release(); groups.release(); count.release();
  }
};

struct GicpState {
  scvod_gicp_params P{};
  bool have_target = false;
  GCloud tgt, src, probe;
  // per-iteration query structures (source points binned into the target grid)
  DBuf<float4> moved, q_sorted;
  DBuf<int> q_cell_of, q_slot, q_cnt, q_start, q_block_sum, q_cell, q_groups;
  DBuf<int> counters;            // [0] group counter, [1..6] bbox scratch
  DBuf<int> box;
  DBuf<unsigned long long> acc;  // 29 fixed-point sums
  unsigned long long* h_acc = nullptr;  // pinned
  int* h_box = nullptr;                 // pinned
  int sms = 148;
  void release() {
    tgt.release(); src.release(); probe.release(); moved.release(); q_sorted.release(); q_cell_of.release(); q_slot.release();
    q_cnt.release(); q_start.release(); q_block_sum.release(); q_cell.release(); q_groups.release(); counters.release(); box.release();
    acc.release();
    if (h_acc) cudaFreeHost(h_acc);
    if (h_box) cudaFreeHost(h_box);
    h_acc = nullptr;
    h_box = nullptr;
  }
};

void gicp_free(void* p) {
  GicpState* s = (GicpState*)p;
  s->release();
  delete s;
}

int get_state(scvod_ctx* c, GicpState** out) {
  void** slot = ctx_gicp_slot(c);
  if (!*slot) {
    GicpState* s = new GicpState();
    scvod_gicp_default_params(&s->P);
    cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, ctx_device(c));
    if (s->sms <= 0) s->sms = 148;
    GCU(cudaMallocHost((void**)&s->h_acc, sizeof(unsigned long long) * 32));
    GCU(cudaMallocHost((void**)&s->h_box, sizeof(int) * 8));
    GCU(s->counters.alloc(8));
    GCU(s->box.alloc(8));
    GCU(s->acc.alloc(32));
    *slot = s;
    ctx_set_gicp_free(c, gicp_free);
  }
  *out = (GicpState*)*slot;
  return SCVOD_OK;
}

inline float ord2f(int o) {
  int i = o >= 0 ? o : o ^ 0x7fffffff;
  float f;
  std::memcpy(&f, &i, 4);
  return f;
}

int grid_blocks(const GicpState& S, int n, int threads) {
  int want = (n + threads - 1) / threads;
  int cap = S.sms * 8;
  return std::max(1, std::min(want, cap));
}

#define GT(name) LaunchTimer timer__(name, (void*)st)

// exclusive scan cnt[0..L) -> start[0..L]; also zeroes counters[0]
int run_scan(scvod_ctx* c, GicpState& S, const int* cnt, int L, int* start, DBuf<int>& block_sum, cudaStream_t st) {
  const int nb = (L + 1023) / 1024;
  GCU(block_sum.alloc(nb + 1));
  { GT("k_gicp_scan"); k_gicp_scan_blocks<<<nb, 1024, 0, st>>>(cnt, L, start, block_sum.p);
    k_gicp_scan_sums<<<1, 1024, 0, st>>>(block_sum.p, nb, start, L, S.counters.p);
    k_gicp_scan_add<<<nb, 1024, 0, st>>>(start, L, block_sum.p); }
  ctx_add_launches(c, 3);
  return SCVOD_OK;
}

// spec §2 + §3 for one cloud whose points are already in cl.pts
int build_cloud(scvod_ctx* c, GicpState& S, GCloud& cl, int n, const scvod_gicp_params& P) {
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  cl.n = n;
  // bounding box -> grid geometry (host decides the dimensions)
  int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  GCU(cudaMemcpyAsync(S.box.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  if (n > 0) {
    { GT("k_gicp_bbox"); k_gicp_bbox<<<grid_blocks(S, n, 256), 256, 0, st>>>(cl.pts.p, n, S.box.p); }
    ctx_add_launches(c, 1);
  }
  GCU(cudaMemcpyAsync(S.h_box, S.box.p, sizeof(int) * 6, cudaMemcpyDeviceToHost, st));
  GCU(cudaStreamSynchronize(st));
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = n > 0 ? ord2f(S.h_box[a]) : 0.f;
    hi[a] = n > 0 ? ord2f(S.h_box[3 + a]) : 0.f;
    if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return api_fail(SCVOD_ERR_ARG, "GICP cloud holds non-finite coordinates");
  }
  float h = P.cov_radius;
  GGrid g;
  for (;;) {  // spec §2: cell edge = cov_radius (doubled while the grid is too large), margin = `rings` empty layers
    g.h = h;
    g.rings = std::max(1, (int)std::ceil(P.max_corr_dist / h));
    const float margin = (float)g.rings * h;
    g.ox = lo[0] - margin;
    g.oy = lo[1] - margin;
    g.oz = lo[2] - margin;
    g.nx = (int)floorf((hi[0] - g.ox) / h) + 1 + g.rings;
    g.ny = (int)floorf((hi[1] - g.oy) / h) + 1 + g.rings;
    g.nz = (int)floorf((hi[2] - g.oz) / h) + 1 + g.rings;
    if ((double)g.nx * g.ny * g.nz <= (double)kMaxCells) break;
    h = h * 2.f;
  }
  g.ncells = g.nx * g.ny * g.nz;
  cl.g = g;
  const size_t N = (size_t)std::max(n, 1);
  GCU(cl.sorted.alloc(N));
  GCU(cl.normal.alloc(N));
  GCU(cl.cell_of.alloc(N));
  GCU(cl.slot_in_cell.alloc(N));
  GCU(cl.tmp.alloc(N));
  GCU(cl.sorted_cell.alloc(N));
  GCU(cl.groups.alloc(N));
  GCU(cl.count.alloc(N));
  GCU(cl.cell_cnt.alloc((size_t)g.ncells + 1));
  GCU(cl.start.alloc((size_t)g.ncells + 2));
  GCU(cudaMemsetAsync(cl.cell_cnt.p, 0, sizeof(int) * ((size_t)g.ncells + 1), st));
  const int gb = grid_blocks(S, n, 256);
  if (n > 0) {
    { GT("k_gicp_count"); k_gicp_count<<<gb, 256, 0, st>>>(cl.pts.p, n, g, cl.cell_of.p, cl.slot_in_cell.p, cl.cell_cnt.p); }
    ctx_add_launches(c, 1);
  }
  int rc = run_scan(c, S, cl.cell_cnt.p, g.ncells, cl.start.p, cl.block_sum, st);
  if (rc) return rc;
  if (n > 0) {
    { GT("k_gicp_fill_tmp"); k_gicp_fill_tmp<<<gb, 256, 0, st>>>(n, cl.cell_of.p, cl.slot_in_cell.p, cl.start.p, cl.tmp.p); }
    { GT("k_gicp_place_ranked"); k_gicp_place_ranked<<<gb, 256, 0, st>>>(cl.pts.p, n, cl.cell_of.p, cl.start.p, cl.tmp.p, cl.sorted.p, cl.sorted_cell.p, cl.groups.p, S.counters.p); }
    const float r2 = P.cov_radius * P.cov_radius;
    { GT("k_gicp_normals"); k_gicp_normals<<<S.sms * 4, kWarpsPerCta * 32, 0, st>>>(cl.sorted.p, cl.sorted_cell.p, cl.start.p, cl.groups.p, S.counters.p, g, r2,
                                                                                   P.min_neighbors, P.planarity, cl.normal.p, cl.count.p); }
    ctx_add_launches(c, 3);
  }
  GCU(cudaGetLastError());
  return SCVOD_OK;
}

int upload(scvod_ctx* c, GCloud& cl, const void* xyzi, int n, bool on_device) {
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  GCU(cl.pts.alloc((size_t)std::max(n, 1)));
  if (n > 0) GCU(cudaMemcpyAsync(cl.pts.p, xyzi, sizeof(float4) * (size_t)n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  return SCVOD_OK;
}

bool cholesky_solve6(const double H[36], const double g[6], double x[6]) {
  double L[36] = {0};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = H[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i * 6 + i] = std::sqrt(s);
      } else {
        L[i * 6 + j] = s / L[j * 6 + j];
      }
    }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = -g[i];
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
    y[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s / L[i * 6 + i];
  }
  return true;
}

void rodrigues(const double w[3], double E[3][3]) {
  const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double a = 1.0, b = 0.5;
  if (th >= 1e-12) {
    a = std::sin(th) / th;
    b = (1.0 - std::cos(th)) / (th * th);
  }
  const double K[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double k2 = 0;
      for (int k = 0; k < 3; ++k) k2 += K[i][k] * K[k][j];
      E[i][j] = (i == j ? 1.0 : 0.0) + a * K[i][j] + b * k2;
    }
}

int set_target_impl(scvod_ctx* c, const void* xyzi, int n, const scvod_gicp_params* p, bool on_device) {
  if (!c || (!xyzi && n > 0) || n < 0) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_gicp_set_target");
  GCU(cudaSetDevice(ctx_device(c)));
  GicpState* S;
  int rc = get_state(c, &S);
  if (rc) return rc;
  if (p) S->P = *p;
  if (!(S->P.cov_radius > 0.f) || !(S->P.max_corr_dist > 0.f)) return api_fail(SCVOD_ERR_ARG, "GICP radii must be positive");
  rc = upload(c, S->tgt, xyzi, n, on_device);
  if (rc) return rc;
  rc = build_cloud(c, *S, S->tgt, n, S->P);
  if (rc) return rc;
  S->have_target = true;
  return SCVOD_OK;
}

int align_impl(scvod_ctx* c, const void* xyzi, int n, const float T0[12], scvod_gicp_result* out, bool on_device) {
  if (!c || (!xyzi && n > 0) || n < 0 || !T0 || !out) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_gicp_align");
  GCU(cudaSetDevice(ctx_device(c)));
  GicpState* Sp;
  int rc = get_state(c, &Sp);
  if (rc) return rc;
  GicpState& S = *Sp;
  if (!S.have_target) return api_fail(SCVOD_ERR_STATE, "scvod_gicp_align called before scvod_gicp_set_target");
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  const scvod_gicp_params& P = S.P;
  rc = upload(c, S.src, xyzi, n, on_device);
  if (rc) return rc;
  rc = build_cloud(c, S, S.src, n, P);
  if (rc) return rc;
  std::memset(out, 0, sizeof(*out));
  // valid counts (inspection): one small reduction on the host is not worth a kernel; count on the device lazily
  const GGrid tg = S.tgt.g;
  const size_t N = (size_t)std::max(n, 1);
  GCU(S.moved.alloc(N));
  GCU(S.q_sorted.alloc(N));
  GCU(S.q_cell_of.alloc(N));
  GCU(S.q_slot.alloc(N));
  GCU(S.q_cell.alloc(N));
  GCU(S.q_groups.alloc(N));
  GCU(S.q_cnt.alloc((size_t)tg.ncells + 1));
  GCU(S.q_start.alloc((size_t)tg.ncells + 2));
  double R[3][3], t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = T0[4 * i + j];
    t[i] = T0[4 * i + 3];
  }
  double H[36], g[6], cost = 0;
  long long ncorr = 0;
  const float dmax2 = P.max_corr_dist * P.max_corr_dist;
  const float k1 = 1.0f - P.cov_eps;
  const int gb = grid_blocks(S, n, 256);
  for (int it = 0; it < P.max_iter && n > 0 && S.tgt.n > 0; ++it) {
    Rt T;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T.r[3 * i + j] = (float)R[i][j];
      T.t[i] = (float)t[i];
    }
    GCU(cudaMemsetAsync(S.q_cnt.p, 0, sizeof(int) * ((size_t)tg.ncells + 1), st));
    { GT("k_gicp_src_count"); k_gicp_src_count<<<gb, 256, 0, st>>>(S.src.sorted.p, S.src.normal.p, n, T, tg, S.moved.p, S.q_cell_of.p, S.q_slot.p, S.q_cnt.p, S.acc.p); }
    rc = run_scan(c, S, S.q_cnt.p, tg.ncells, S.q_start.p, S.q_block_sum, st);
    if (rc) return rc;
    { GT("k_gicp_src_place"); k_gicp_src_place<<<gb, 256, 0, st>>>(S.moved.p, n, S.q_cell_of.p, S.q_slot.p, S.q_start.p, S.q_sorted.p, S.q_cell.p, S.q_groups.p, S.counters.p); }
    { GT("k_gicp_corr"); k_gicp_corr<<<S.sms * 4, kWarpsPerCta * 32, 0, st>>>(S.q_sorted.p, S.q_cell.p, S.q_start.p, S.q_groups.p, S.counters.p, S.src.normal.p, S.tgt.sorted.p,
                                                                             S.tgt.normal.p, S.tgt.start.p, tg, T, dmax2, k1, S.acc.p); }
    ctx_add_launches(c, 3);
    GCU(cudaGetLastError());
    GCU(cudaMemcpyAsync(S.h_acc, S.acc.p, sizeof(unsigned long long) * kNumAcc, cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    double v[kNumAcc];
    for (int i = 0; i < kNumAcc; ++i) v[i] = (double)(long long)S.h_acc[i] / (i < 21 ? kScaleH : kScaleG);
    int q = 0;
    for (int r = 0; r < 6; ++r)
      for (int cc = r; cc < 6; ++cc) {
        H[r * 6 + cc] = v[q];
        H[cc * 6 + r] = v[q];
        ++q;
      }
    for (int r = 0; r < 6; ++r) g[r] = v[21 + r];
    cost = v[27];
    ncorr = (long long)S.h_acc[28];
    out->iterations = it + 1;
    if (ncorr < 10) break;
    double delta[6];
    if (!cholesky_solve6(H, g, delta)) break;
    double E[3][3];
    rodrigues(delta, E);
    double Rn[3][3], tn[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Rn[i][j] = E[i][0] * R[0][j] + E[i][1] * R[1][j] + E[i][2] * R[2][j];
      tn[i] = E[i][0] * t[0] + E[i][1] * t[1] + E[i][2] * t[2] + delta[3 + i];
    }
    std::memcpy(R, Rn, sizeof(R));
    std::memcpy(t, tn, sizeof(t));
    const double wn = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    const double vn = std::sqrt(delta[3] * delta[3] + delta[4] * delta[4] + delta[5] * delta[5]);
    if (wn < (double)P.rot_eps && vn < (double)P.trans_eps) {
      out->converged = 1;
      break;
    }
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) out->T[4 * i + j] = (float)R[i][j];
    out->T[4 * i + 3] = (float)t[i];
  }
  // Utility::rotationMatrixToEulerAngles (reference include/utility.h:488-505)
  const double sy = std::sqrt(R[0][0] * R[0][0] + R[1][0] * R[1][0]);
  double rx, ry, rz;
  if (!(sy < 1e-6)) {
    rx = std::atan2(R[2][1], R[2][2]);
    ry = std::atan2(-R[2][0], sy);
    rz = std::atan2(R[1][0], R[0][0]);
  } else {
    rx = std::atan2(-R[1][2], R[1][1]);
    ry = std::atan2(-R[2][0], sy);
    rz = 0;
  }
  out->pose6[0] = (float)t[0];
  out->pose6[1] = (float)t[1];
  out->pose6[2] = (float)t[2];
  out->pose6[3] = (float)rx;
  out->pose6[4] = (float)ry;
  out->pose6[5] = (float)rz;
  std::memcpy(out->H, H, sizeof(H));
  std::memcpy(out->b, g, sizeof(g));
  out->cost = cost;
  out->n_corr = (int32_t)ncorr;
  out->n_src_valid = -1;  // filled by scvod_gicp_normals on request; not needed by the solver
  out->n_tgt_valid = -1;
  return SCVOD_OK;
}

}  // namespace
}  // namespace scvod

using namespace scvod;

extern "C" void scvod_gicp_default_params(scvod_gicp_params* p) {
  if (!p) return;
  p->cov_radius = 0.6f;
  p->max_corr_dist = 1.0f;
  p->cov_eps = 1e-3f;
  p->planarity = 0.25f;
  p->min_neighbors = 6;
  p->max_iter = 32;
  p->rot_eps = 5e-5f;
  p->trans_eps = 5e-5f;
}

extern "C" int scvod_gicp_set_target(scvod_ctx* c, const float* tgt_xyzi, int n, const scvod_gicp_params* p) {
  return set_target_impl(c, tgt_xyzi, n, p, false);
}
extern "C" int scvod_gicp_set_target_dev(scvod_ctx* c, const void* tgt_xyzi_dev, int n, const scvod_gicp_params* p) {
  return set_target_impl(c, tgt_xyzi_dev, n, p, true);
}
extern "C" int scvod_gicp_align(scvod_ctx* c, const float* src_xyzi, int n, const float T0[12], scvod_gicp_result* out) {
  return align_impl(c, src_xyzi, n, T0, out, false);
}
extern "C" int scvod_gicp_align_dev(scvod_ctx* c, const void* src_xyzi_dev, int n, const float T0[12], scvod_gicp_result* out) {
  return align_impl(c, src_xyzi_dev, n, T0, out, true);
}

extern "C" int scvod_gicp_normals(scvod_ctx* c, const float* xyzi, int n, const scvod_gicp_params* p, float* normals3, uint8_t* valid,
                                  int32_t* count) {
  if (!c || (!xyzi && n > 0) || n < 0) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_gicp_normals");
  GCU(cudaSetDevice(ctx_device(c)));
  GicpState* S;
  int rc = get_state(c, &S);
  if (rc) return rc;
  scvod_gicp_params P = S->P;
  if (p) P = *p;
  if (!(P.cov_radius > 0.f) || !(P.max_corr_dist > 0.f)) return api_fail(SCVOD_ERR_ARG, "GICP radii must be positive");
  if (n == 0) return SCVOD_OK;
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  rc = upload(c, S->probe, xyzi, n, false);
  if (rc) return rc;
  rc = build_cloud(c, *S, S->probe, n, P);
  if (rc) return rc;
  DBuf<float> dn;
  DBuf<uint8_t> dv;
  DBuf<int> dc;
  GCU(dn.alloc((size_t)3 * n));
  GCU(dv.alloc(n));
  GCU(dc.alloc(n));
  { LaunchTimer timer__("k_gicp_unsort", (void*)st);
    k_gicp_unsort<<<grid_blocks(*S, n, 256), 256, 0, st>>>(S->probe.sorted.p, S->probe.normal.p, S->probe.count.p, n, dn.p, dv.p, dc.p); }
  ctx_add_launches(c, 1);
  GCU(cudaGetLastError());
  if (normals3) GCU(cudaMemcpyAsync(normals3, dn.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
  if (valid) GCU(cudaMemcpyAsync(valid, dv.p, n, cudaMemcpyDeviceToHost, st));
  if (count) GCU(cudaMemcpyAsync(count, dc.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  GCU(cudaStreamSynchronize(st));
  dn.release();
  dv.release();
  dc.release();
  return SCVOD_OK;
}
