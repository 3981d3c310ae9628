// scvod_track.cu — stand-alone binning (scvod_bin), the tracking diff of SSC::tracking (transformCloud + re-binning + next-frame
// lookup, reference include/utility.h:394-406, src/ssc.cpp:1275-1315), per-point classes, the static submap and small helpers.
#include "scvod_kernel_common.cuh"

namespace scvod {

struct Mat34 {
  float m[12];
};

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
                                               const int32_t* __restrict__ tv_pts /* apri indices of the subgroups of tainted voxels */,
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
  if (RUNS) {
    // The launch is a chain of dependent memory accesses per point (run table -> CSR offsets -> voxel point list -> point ->
    // bitmap word -> table atomic), i.e. latency bound, and it shares the GPU with the kernels of the other contexts: what it
    // costs them is the time its warps stay resident.  Every thread therefore carries kU points through the chain TOGETHER:
    // the binary searches advance in lock step (fixed trip counts), so each level of the chain is kU independent loads in
    // flight instead of one, and the latency of a level is paid once per kU points.
    constexpr int kU = 4;
    const int nsuper = (k + 256 * kU - 1) / (256 * kU);
    int run_iters = 0;
    while ((1 << run_iters) < runs.n) ++run_iters;
    for (int blk = blockIdx.x; blk < nsuper; blk += gridDim.x) {
      int i[kU], lo[kU];
      bool live[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        i[u] = (blk * kU + u) * 256 + threadIdx.x;
        live[u] = i[u] < k;
        lo[u] = 0;
      }
      // last run with dst_off <= i (kernel-argument table: constant bank): branch-free bisection on the bit positions
      for (int bit = run_iters - 1; bit >= 0; --bit) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int cand = lo[u] | (1 << bit);
          if (cand < runs.n && runs.dst_off[cand] <= (live[u] ? i[u] : 0)) lo[u] = cand;
        }
      }
      int src[kU], a[kU], jr[kU], blen[kU];
      const int32_t* po[kU];
      int len_max = 1;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        src[u] = runs.src[lo[u]];
        a[u] = 0;
        blen[u] = (live[u] && src[u] >= 0) ? runs.len[lo[u]] : 1;
        po[u] = csr_ptoff + max(src[u], 0);
        len_max = max(len_max, blen[u]);
      }
      int first_pt[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) first_pt[u] = (live[u] && src[u] >= 0) ? __ldg(&po[u][0]) : 0;
#pragma unroll
      for (int u = 0; u < kU; ++u) jr[u] = i[u] - runs.dst_off[lo[u]] + first_pt[u];
      // own voxels: last CSR position of the run with ptoff <= jr (same bisection, trip count = bits of the longest run among
      // the warp's points; the slices are short and L1 resident)
      len_max = __reduce_max_sync(0xffffffffu, len_max);
      int csr_iters = 0;
      while ((1 << csr_iters) < len_max) ++csr_iters;
      for (int bit = csr_iters - 1; bit >= 0; --bit) {
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int cand = a[u] | (1 << bit);
          if (cand < blen[u] && __ldg(&po[u][cand]) <= jr[u]) a[u] = cand;
        }
      }
      int sgx[kU], sgy[kU], sgz[kU], sgw[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        sgz[u] = runs.cluster[lo[u]];
        if (live[u] && src[u] >= 0) {
          sgx[u] = i[u] - (jr[u] - __ldg(&po[u][a[u]]));
          sgy[u] = __ldg(&csr_vox[src[u] + a[u]]);
          sgw[u] = runs.order[lo[u]] + __ldg(&csr_part[src[u] + a[u]]);
        } else {
          sgx[u] = runs.dst_off[lo[u]];
          sgy[u] = live[u] ? src[u] : -1;
          sgw[u] = runs.order[lo[u]];
        }
      }
      // the point: own voxels through the voxel CSR of the frame (two more levels), carried ranges directly
      int voff[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        voff[u] = (live[u] && sgy[u] >= 0 && !(sgy[u] & kVirtualVox)) ? __ldg(&vox_off[sgy[u]]) : 0;
      int m[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int j = i[u] - sgx[u];
        m[u] = j;
        if (live[u] && sgy[u] >= 0) m[u] = (sgy[u] & kVirtualVox) ? __ldg(&tv_pts[(sgy[u] & (kVirtualVox - 1)) + j]) : __ldg(&vox_pts[voff[u] + j]);
      }
      float4 p[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        p[u] = make_float4(1.f, 1.f, 0.f, 0.f);
        if (live[u]) p[u] = (sgy[u] >= 0) ? __ldg(&own[m[u]]) : __ldg(&carried[(-1 - sgy[u]) + m[u]]);
      }
      int key[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        float4 q;
        q.x = da(da(da(dm(T.m[0], p[u].x), dm(T.m[1], p[u].y)), dm(T.m[2], p[u].z)), T.m[3]);
        q.y = da(da(da(dm(T.m[4], p[u].x), dm(T.m[5], p[u].y)), dm(T.m[6], p[u].z)), T.m[7]);
        q.z = da(da(da(dm(T.m[8], p[u].x), dm(T.m[9], p[u].y)), dm(T.m[10], p[u].z)), T.m[11]);
        q.w = p[u].w;
        if (live[u]) out_xyzi[i[u]] = q;
        const BinIdx r = dev_bin_filtered(q.x, q.y, q.z, bp);
        key[u] = r.vid + g.key_off;
        if (!live[u] || key[u] < 0 || key[u] >= g.key_count) key[u] = -1;
      }
      // next.hash_cloud.find: occupancy word and its rank are fetched together (the rank is wasted on a miss; misses are rare)
      uint32_t w[kU];
      int wrk[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        w[u] = key[u] >= 0 ? __ldg(&bm[key[u] >> 5]) : 0u;
        wrk[u] = key[u] >= 0 ? __ldg(&wr[key[u] >> 5]) : 0;
      }
      unsigned long long old[kU];
      int e[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        e[u] = -1;
        old[u] = 0ull;
        if (key[u] >= 0) {
          const uint32_t bit = 1u << (key[u] & 31);
          if (w[u] & bit) {
            e[u] = sgz[u] * vn + wrk[u] + __popc(w[u] & (bit - 1));
            old[u] = atomicMin(&first[e[u]], ((unsigned long long)(unsigned)sgw[u] << 32) | (unsigned)m[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (e[u] >= 0 && old[u] == ~0ull) {  // first point to touch this (cluster, voxel): remember the entry for the epilogue
          const int slot = atomicAdd(&ctr[0], 1);
          if (slot < cap_quads) hit_list[slot] = e[u];
        }
      }
    }
  } else {
  const int nblk = (k + 255) >> 8;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    int4 sg;  // x = dst_off, y = source (>= 0: voxel of frame_pre_, its points come from the voxel CSR; < 0: carried range
              // at -1-y), z = cluster, w = order of the segment inside the cluster's cloud (part index / carried ordinal)
    const int i = (blk << 8) + threadIdx.x;
    {
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
      // a whole voxel of the frame's CSR, or (kVirtualVox) the points one cluster owns of a tainted voxel
      const int m = (sg.y & kVirtualVox) ? tv_pts[(sg.y & (kVirtualVox - 1)) + j] : vox_pts[vox_off[sg.y] + j];
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
    const BinIdx r = dev_bin_filtered(q.x, q.y, q.z, bp);
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

// points of tainted voxels: their class follows the cluster that owns the point, not the voxel (items = batch-global apri position, class)
__global__ void __launch_bounds__(256) k_label_override(const int2* __restrict__ items, int n, const int32_t* __restrict__ apri_src,
                                                        const int64_t* __restrict__ off, const int32_t* __restrict__ item_scan,
                                                        uint8_t* __restrict__ cls) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int2 it = items[i];
    cls[off[item_scan[i]] + apri_src[it.x]] = (uint8_t)it.y;
  }
}

// static submap: every non-dynamic point of the frames of a batch moved to the map frame (transformCloud arithmetic
// with the frame's pose).  A CTA owns a contiguous chunk of a scan and every warp a contiguous part of it: the static
// points are counted first (1 B / point), the CTA reserves its output range with ONE atomic on the global counter,
// and the second pass writes with ballot / popcount ranks (no per-warp atomics on a single address).
__global__ void __launch_bounds__(256) k_submap(const float4* __restrict__ pts, const uint8_t* __restrict__ cls,
                                                const int64_t* __restrict__ off, const float* __restrict__ Ts, int first_scan,
                                                float4* __restrict__ out, unsigned long long* __restrict__ counter, long long cap,
                                                unsigned keep_mask /* bit c set: points of class c go into the submap */) {
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
  for (int i = i0 + lane; i < i1; i += 32) cnt += (int)((keep_mask >> cls[base + i]) & 1u);
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
    const bool keep = (i < i1) && ((keep_mask >> cls[base + i]) & 1u);
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

// Filter soundness probe (scvod_bin_filter_check): n pseudo-random points (counter-based generator; every 8th point is snapped
// onto a structured edge case: an axis, the origin, a gate radius, a bin edge angle) through dev_bin_filtered AND dev_bin_point.
// stats: [0] points, [1] points that took the exact path, [2] mismatches (must be 0), [3] max |q_approx - q_exact| of the sector
// coordinate among filter-decided points in units of 1e-9, [4] same for the azimuth coordinate.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__global__ void __launch_bounds__(256) k_bin_filter_check(long long n, uint32_t seed, BinParams bp, float extent,
                                                          const float4* __restrict__ pts, unsigned long long* __restrict__ stats) {
  unsigned long long n_exact = 0, n_bad = 0;
  unsigned int dqs = 0, dqe = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float x, y, z;
    if (pts) {
      const float4 p = pts[i];
      x = p.x; y = p.y; z = p.z;
    } else {
      const uint32_t h0 = mix32((uint32_t)i * 3u + seed), h1 = mix32((uint32_t)i * 3u + 1u + seed), h2 = mix32((uint32_t)i * 3u + 2u + seed + (uint32_t)(i >> 32));
      x = ((float)h0 * 2.3283064e-10f - 0.5f) * 2.f * extent;
      y = ((float)h1 * 2.3283064e-10f - 0.5f) * 2.f * extent;
      z = ((float)h2 * 2.3283064e-10f - 0.5f) * 0.25f * extent;
      const int kind = (int)(i & 7);
      const int sub = (int)((i >> 3) & 7);
      if (kind == 0) {
        if (sub == 0) y = 0.f;
        else if (sub == 1) x = 0.f;
        else if (sub == 2) { x = 0.f; y = (h0 & 1) ? 0.f : -0.f; }
        else if (sub == 3) y = -0.f;
        else if (sub == 4) {  // on a sector edge, up to rounding
          const float r = fabsf(x) + 1.f;
          const float a = (float)((h1 % 300u)) * bp.sector_res * 0.017453292f;
          x = r * cosf(a); y = r * sinf(a);
        } else if (sub == 5) {  // on an azimuth edge
          const float d = sqrtf(x * x + y * y);
          const float a = (bp.min_azimuth + (float)(h1 % 60u) * bp.azimuth_res) * 0.017453292f;
          z = d * tanf(a);
        } else if (sub == 6) {  // on a gate radius
          const float d = sqrtf(x * x + y * y);
          const float s = ((h1 & 1) ? bp.min_dis : bp.max_dis) / fmaxf(d, 1e-6f);
          x *= s; y *= s;
        } else {
          y = -fabsf(y) * 1e-7f;  // just below the positive x axis: angle rounds to 360
        }
      }
    }
    bool slow;
    const BinIdx f = dev_bin_filtered(x, y, z, bp, &slow);
    const BinResult e = dev_bin_point(x, y, z, bp);
    n_exact += slow ? 1 : 0;
    if (f.ri != e.ri || f.si != e.si || f.ei != e.ei || f.vid != e.vid || f.pass != e.pass) ++n_bad;
    if (!slow) {
      float a = atan2f(y, x);
      if (!(y >= 0.f)) a += 6.283185307179586f;
      const float qs = (a * 57.29577951308232f - bp.min_angle) * bp.inv_sector_res;
      const float qe = (atan2f(z, e.dis) * 57.29577951308232f - bp.min_azimuth) * bp.inv_azimuth_res;
      const float es = dd(ds(e.angle, bp.min_angle), bp.sector_res), ee = dd(ds(e.azimuth, bp.min_azimuth), bp.azimuth_res);
      dqs = max(dqs, (unsigned int)fminf(4.0e9f, fabsf(qs - es) * 1e9f));
      dqe = max(dqe, (unsigned int)fminf(4.0e9f, fabsf(qe - ee) * 1e9f));
    }
  }
  atomicAdd(&stats[1], n_exact);
  atomicAdd(&stats[2], n_bad);
  atomicMax(&stats[3], (unsigned long long)dqs);
  atomicMax(&stats[4], (unsigned long long)dqe);
  if (blockIdx.x == 0 && threadIdx.x == 0) stats[0] = (unsigned long long)n;
}

__global__ void k_atan2f_probe(const float* y, const float* x, float* out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = dev_atan2f(y[i], x[i]);
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
                 const int32_t* csr_part, const int32_t* tv_pts, int k, const float T12[12], const uint32_t* next_bitmap, const int32_t* next_word_rank, int ncl,
                 int vn, float4* out_xyzi, unsigned long long* first, int32_t* ctr_dev, int32_t* hit_list_dev, int32_t* out_quads_mapped,
                 int cap_quads, void* stream_) {
  if (k <= 0 || ncl <= 0 || vn <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  Mat34 T;
  for (int i = 0; i < 12; ++i) T.m[i] = T12[i];
  int blocks = runs ? (k + 1023) / 1024 : (k + 255) / 256;  // the run-table variant carries 4 points per thread
  static const int forced = getenv("SCVOD_TRACK_CTAS") ? std::max(1, atoi(getenv("SCVOD_TRACK_CTAS"))) : 0;  // tuning hook
  const int ctas_per_sm = forced ? forced : std::max(1, hp.track_ctas_per_sm);
  int cap = num_sms() * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (runs) {
    TIMED("k_track", TSTREAM);
    k_track<true><<<blocks, 256, 0, st>>>(own_xyzi, vox_off, vox_pts, carried, nullptr, nullptr, 0, *runs, csr_ptoff, csr_vox, csr_part, tv_pts, k, T,
                                           make_bin_params(hp), hp.g, next_bitmap, next_word_rank, vn, out_xyzi, first, ctr_dev, hit_list_dev,
                                           out_quads_mapped, cap_quads);
  } else {
    static const TrackRuns none = {};
    TIMED("k_track_segs", TSTREAM);
    k_track<false><<<blocks, 256, 0, st>>>(own_xyzi, vox_off, vox_pts, carried, segs, first_seg, nseg, none, nullptr, nullptr, nullptr, tv_pts, k, T,
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

int launch_label_override(const int32_t* items_dev /* [n][2] */, const int32_t* item_scan_dev, int n, const int32_t* apri_src, const int64_t* off,
                          uint8_t* cls, void* stream_) {
  if (n <= 0) return 0;
  { TIMED("k_label_override", TSTREAM); k_label_override<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(reinterpret_cast<const int2*>(items_dev), n, apri_src, off, item_scan_dev, cls); }
  return 1;
}

int launch_submap(const float4* pts, const uint8_t* cls, const int64_t* off, const float* Ts_dev, int first_scan, int nscans,
                  int max_scan_points, float4* out, unsigned long long* counter, long long cap, unsigned keep_mask, void* stream_) {
  if (nscans <= 0) return 0;
  dim3 grid(grid_x_for(nscans, max_scan_points, 256), nscans);
  { TIMED("k_submap", TSTREAM); k_submap<<<grid, 256, 0, (cudaStream_t)stream_>>>(pts, cls, off, Ts_dev, first_scan, out, counter, cap, keep_mask); }
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

int launch_bin_filter_check(const HostParams& hp, long long n, uint32_t seed, float extent, const float4* pts_dev, unsigned long long* stats_dev,
                            void* stream_) {
  { TIMED("k_bin_filter_check", TSTREAM); k_bin_filter_check<<<num_sms() * 8, 256, 0, (cudaStream_t)stream_>>>(n, seed, make_bin_params(hp), extent, pts_dev, stats_dev); }
  return 1;
}

int launch_atan2f_probe(const float* y, const float* x, float* out, long long n, void* stream_) {
  { TIMED("k_atan2f_probe", TSTREAM); k_atan2f_probe<<<num_sms() * 8, 256, 0, (cudaStream_t)stream_>>>(y, x, out, n); }
  return 1;
}

}  // namespace scvod
