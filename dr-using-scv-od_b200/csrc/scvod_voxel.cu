// scvod_voxel.cu — occupancy descriptor (SSC::makeHashCloud, reference src/ssc.cpp:253-289), voxel adjacency / connected
// components / intensity-similarity edges (src/ssc.cpp:395-411, 299-351, 587-595) and the replay of the sequential cluster
// names of SSC::clusterAndCreateFrame (src/ssc.cpp:299-354).
#include "scvod_kernel_common.cuh"

namespace scvod {

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
                                                   int32_t* __restrict__ vox_cnt, int32_t* __restrict__ vox_tnt,
                                                   int32_t* __restrict__ scan_counts) {
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
      vox_tnt[base + ex] = -1;
      ++ex;
    }
  }
  if (threadIdx.x == 0) scan_counts[b * 8 + 3] = total;
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
// Tainted voxels.  A point whose range / sector / azimuth index is -1 (dis == min_dis, angle == min_angle i.e. y == 0 with
// x > 0, azimuth == min_azimuth; ssc.cpp:185-188) hashes into a voxel that is not its own cell: the voxel's points then
// no longer share one neighbour list, and clusterAndCreateFrame (ssc.cpp:299-352) has to be replayed point by point for
// them.  These kernels build the (tiny) side tables: the tainted voxels of a scan in ascending compact id and their
// points in ptIdx order.  Scans without such points leave after one load.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_taint_mark(const int64_t* __restrict__ off, const int32_t* __restrict__ taint_cnt,
                                                    const int32_t* __restrict__ q_list, const int32_t* __restrict__ apri_cid,
                                                    int32_t* __restrict__ vox_tnt) {
  const int b = blockIdx.x;
  const int nq = min(taint_cnt[b * kTaintCntStride], kQuirkCap);
  if (nq == 0) return;
  const int64_t base = off[b];
  for (int k = threadIdx.x; k < nq; k += blockDim.x) {
    const int cid = apri_cid[base + q_list[(size_t)b * kQuirkCap + k]];
    if (cid >= 0) vox_tnt[base + cid] = -2;
  }
}

__global__ void __launch_bounds__(1024) k_taint_build(const int64_t* __restrict__ off, int32_t* __restrict__ taint_cnt,
                                                      const int32_t* __restrict__ q_list, const int32_t* __restrict__ apri_cid,
                                                      const int32_t* __restrict__ vox_cnt, const int32_t* __restrict__ vox_off,
                                                      const int32_t* __restrict__ vox_pts, const float4* __restrict__ apri_xyzi,
                                                      int32_t* __restrict__ vox_tnt, int32_t* __restrict__ tv_cid,
                                                      int32_t* __restrict__ tv_base, int32_t* __restrict__ tp_m,
                                                      int32_t* __restrict__ tp_cid, float4* __restrict__ tp_xyz) {
  __shared__ int s_cid[kTvCap];
  __shared__ int s_base[kTvCap + 1];
  __shared__ int s_w[33];
  __shared__ int s_n;
  const int b = blockIdx.x;
  const int nq_all = taint_cnt[b * kTaintCntStride];
  if (nq_all == 0) return;
  const int64_t base = off[b];
  const int tid = threadIdx.x;
  const int nq = min(nq_all, kQuirkCap);
  if (tid == 0) s_n = 0;
  for (int i = tid; i < kTvCap; i += 1024) s_cid[i] = 0x7fffffff;
  __syncthreads();
  bool overflow = nq_all > kQuirkCap;
  for (int k = tid; k < nq; k += 1024) {
    const int cid = apri_cid[base + q_list[(size_t)b * kQuirkCap + k]];
    if (cid >= 0 && atomicCAS(&vox_tnt[base + cid], -2, -4) == -2) {  // first quirk point of this voxel
      const int slot = atomicAdd(&s_n, 1);
      if (slot < kTvCap) s_cid[slot] = cid;
    }
  }
  __syncthreads();
  const int ntv = min(s_n, kTvCap);
  if (s_n > kTvCap) overflow = true;
  // ascending compact id (bitonic network over the padded table)
  for (int k = 2; k <= kTvCap; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < kTvCap; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const int x = s_cid[i], y = s_cid[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            s_cid[i] = y;
            s_cid[l] = x;
          }
        }
      }
      __syncthreads();
    }
  int carry = 0;
  for (int i0 = 0; i0 < ntv; i0 += 1024) {
    const int i = i0 + tid;
    const int c = (i < ntv) ? vox_cnt[base + s_cid[i]] : 0;
    int total;
    const int ex = block_excl_scan<1024>(c, &total, s_w);
    if (i < ntv) s_base[i] = carry + ex;
    carry += total;
  }
  if (tid == 0) s_base[ntv] = carry;
  __syncthreads();
  const int T = carry;
  if (T > kTpCap) overflow = true;
  if (!overflow) {
    for (int i = tid; i < ntv; i += 1024) {
      tv_cid[(size_t)b * kTvCap + i] = s_cid[i];
      tv_base[(size_t)b * (kTvCap + 1) + i] = s_base[i];
      vox_tnt[base + s_cid[i]] = s_base[i];
    }
    if (tid == 0) tv_base[(size_t)b * (kTvCap + 1) + ntv] = T;
    // points, voxel-major in ptIdx (ascending apri index) order: a warp per voxel
    const int lane = tid & 31, wid = tid >> 5;
    for (int i = wid; i < ntv; i += 32) {
      const int cid = s_cid[i], o = vox_off[base + cid], n = s_base[i + 1] - s_base[i];
      for (int j = lane; j < n; j += 32) {
        const int m = vox_pts[base + o + j];
        const size_t dst = (size_t)b * kTpCap + s_base[i] + j;
        tp_m[dst] = m;
        tp_cid[dst] = cid;
        tp_xyz[dst] = apri_xyzi[base + m];
      }
    }
  }
  if (tid == 0) {
    taint_cnt[b * kTaintCntStride + 1] = overflow ? 0 : ntv;
    taint_cnt[b * kTaintCntStride + 2] = overflow ? 0 : T;
    taint_cnt[b * kTaintCntStride + 3] = overflow ? 1 : 0;
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
                                                 int32_t* __restrict__ vox_root, const int32_t* __restrict__ taint_cnt,
                                                 int32_t* __restrict__ vox_tnt) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  const uint32_t* bm = bitmap + (size_t)b * g.words;
  const int32_t* wr = word_rank + (size_t)b * g.words;
  const bool has_taint = taint_cnt[b * kTaintCntStride + 2] > 0;
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
    if (has_taint && vox_tnt[base + v] == -1) {  // an ordinary voxel that sees a tainted one: its events take the general replay path
      bool near = false;
      for (int k = 0; k < 27; ++k) {
        const int u = out[k];
        if (u >= 0 && vox_tnt[base + u] >= 0) near = true;  // only ever changes between -1 and -3 concurrently: the test is stable
      }
      if (near) vox_tnt[base + v] = -3;
    }
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
                                                   const int32_t* __restrict__ vox_nbr, int32_t* __restrict__ vox_root,
                                                   const int32_t* __restrict__ taint_cnt, const int32_t* __restrict__ vox_tnt) {
  const int b = blockIdx.y;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  int32_t* parent = vox_root + base;
  // Tainted voxels stay singletons here: two adjacent ORDINARY voxels always end up in one cluster (each sees the other,
  // ssc.cpp:323-351), a tainted voxel's points need not.  Components of ordinary voxels therefore lie inside one cluster.
  const bool has_taint = taint_cnt[b * kTaintCntStride + 2] > 0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    if (has_taint && vox_tnt[base + v] >= 0) continue;
    const int32_t* nb = vox_nbr + 27 * (base + v);
    for (int t = 0; t < 27; ++t) {
      int u = nb[t];
      if (u < 0 || u >= v) continue;  // each undirected edge once
      if (has_taint && vox_tnt[base + u] >= 0) continue;
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

// Components of the voxel graph including every link a tainted voxel takes part in: the ordinary voxels that list it among
// their 27 neighbours, and the neighbour lists of each of its points.  The name replay deals whole groups to its warps.
__global__ void __launch_bounds__(1024) k_taint_group(const int64_t* __restrict__ off, const int32_t* __restrict__ scan_counts,
                                                      const int32_t* __restrict__ taint_cnt, GridSpec g, BinParams bp,
                                                      const uint32_t* __restrict__ bitmap, const int32_t* __restrict__ word_rank,
                                                      const int32_t* __restrict__ vox_nbr, const int32_t* __restrict__ vox_root,
                                                      const int32_t* __restrict__ vox_tnt, const int32_t* __restrict__ tp_cid,
                                                      const float4* __restrict__ tp_xyz, int32_t* __restrict__ vox_group) {
  const int b = blockIdx.x;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  const int T = taint_cnt[b * kTaintCntStride + 2];
  int32_t* parent = vox_group + base;
  for (int v = threadIdx.x; v < V; v += blockDim.x) parent[v] = vox_root[base + v];
  if (T == 0) return;
  __syncthreads();
  auto unite = [&](int a, int c) {
    int ra = uf_find(parent, a), rb = uf_find(parent, c);
    while (ra != rb) {
      if (ra < rb) {
        const int tmp = ra;
        ra = rb;
        rb = tmp;
      }
      const int old = atomicCAS(&parent[ra], ra, rb);
      if (old == ra) break;
      ra = uf_find(parent, old);
      rb = uf_find(parent, rb);
    }
  };
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    if (vox_tnt[base + v] != -3) continue;
    const int32_t* nb = vox_nbr + 27 * (base + v);
    for (int k = 0; k < 27; ++k) {
      const int u = nb[k];
      if (u >= 0 && vox_tnt[base + u] >= 0) unite(v, u);
    }
  }
  const uint32_t* bm = bitmap + (size_t)b * g.words;
  const int32_t* wr = word_rank + (size_t)b * g.words;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float4 q = tp_xyz[(size_t)b * kTpCap + t];
    const int X = tp_cid[(size_t)b * kTpCap + t];
    const BinResult r = dev_bin_point(q.x, q.y, q.z, bp);
    for (int x = r.ri - 1; x <= r.ri + 1; ++x)
      for (int y = r.si - 1; y <= r.si + 1; ++y)
        for (int z = r.ei - 1; z <= r.ei + 1; ++z) {
          if (x > g.range_num - 1 || x < 0 || y > g.sector_num - 1 || y < 0 || z > g.azimuth_num - 1 || z < 0) continue;
          const int c = vox_lookup(bm, wr, g, x * g.sector_num + y + z * g.range_num * g.sector_num);
          if (c >= 0) unite(X, c);
        }
  }
  __syncthreads();
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    int r = v;
    while (parent[r] != r) r = parent[r];
    parent[v] = r;  // concurrent walks stay correct: an entry only ever moves to an ancestor
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
struct TaintArgs {  // side tables of the tainted voxels (k_taint_build) + what a point needs to compute its own neighbour list
  const int32_t* taint_cnt;
  const int32_t* vox_tnt;
  const int32_t* vox_cnt;
  const int32_t* tp_cid;
  const float4* tp_xyz;
  int32_t* tp_name;
  const uint32_t* bitmap;
  const int32_t* word_rank;
  GridSpec g;
  BinParams bp;
};

constexpr int kReplayWarps = 8;  // warps per scan: components are dealt to them by event count
constexpr int kReplayRows = 32;   // events per chunk; neighbour rows of the next chunk are prefetched while one is replayed

template <int NW, bool GLOBAL>
__global__ void __launch_bounds__(NW * 32) k_name_replay(const int64_t* __restrict__ off, int32_t* __restrict__ scan_counts,
                                                         const int32_t* __restrict__ ev_cid, const int32_t* __restrict__ vox_root,
                                                         const int32_t* __restrict__ vox_nbr, int2* __restrict__ ev_list,
                                                         int32_t* __restrict__ g_parent, int32_t* __restrict__ g_setname,
                                                         int32_t* __restrict__ g_first, int32_t* __restrict__ g_state,
                                                         int32_t* __restrict__ g_flags, int32_t* __restrict__ vox_name,
                                                         int32_t* __restrict__ name_first, int name_cap, TaintArgs ta) {
  constexpr int T = NW * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[T / 32 + 1];
  __shared__ int s_load[NW], s_lbase[NW + 1], s_cursor[NW], s_free[NW + 1];
  __shared__ int s_wcnt[NW][NW];
  __shared__ int s_big[32], s_nbig, s_target;
  __shared__ int s_row[NW][32];
  const int b = blockIdx.x;
  const int64_t base = off[b];
  const int V = scan_counts[b * 8 + 3];
  const int E = scan_counts[b * 8 + 5];
  // tainted voxels (see k_taint_build): their NP points are nodes V .. V + NP - 1 of the union-find
  const int NP = ta.taint_cnt[b * kTaintCntStride + 2];
  const bool has_taint = NP > 0;
  const int NV = V + NP;
  const int32_t* tnt = ta.vox_tnt + base;
  const int32_t* tpc = ta.tp_cid + (size_t)b * kTpCap;
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
    stable = state + NV;  // g_state has 4 bytes per voxel (NV <= 2 * scan points is checked by the host)
    flags = reinterpret_cast<uint32_t*>(g_flags + base);  // E <= M <= N ints available: nfw flags, then nfw prefixes
    fprefix = g_flags + base + nfw;
  } else {
    // shared memory holds what sits on the dependent path of every event (parent, state, stable: 6 B / voxel, so scans of
    // up to ~25k voxels fit); set names and first events are written once and read at the end: global scratch
    parent = reinterpret_cast<int32_t*>(sm_state);
    flags = reinterpret_cast<uint32_t*>(parent + NV);
    fprefix = reinterpret_cast<int32_t*>(flags + nfw);
    state = reinterpret_cast<uint8_t*>(fprefix + nfw);
    stable = state + NV;
    setname = g_setname + base;
    first_ev = g_first + base;
  }
  const int32_t* ev = ev_cid + base;
  const int32_t* root = vox_root + base;  // work partition: the groups of k_taint_group when the batch has tainted voxels
  const int32_t* nbr = vox_nbr + 27 * base;
  int2* lst = ev_list + base;
  auto ev_voxel = [&](int node) { return node >= V ? tpc[node - V] : node; };  // the voxel an event's point hashes into

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
  for (int e = tid; e < E; e += T) atomicAdd(&cnt[root[ev_voxel(ev[e])]], 1);
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
      own = owner[root[ev_voxel(cid)]];
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
  for (int v = tid; v < NV; v += T) {
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
      const bool need = evn.y >= 0 && evn.y < V && !stable[evn.y] && (__ffs(same) - 1 == lane);
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
        if (has_taint && (W >= V || tnt[W] == -3)) {
          // ---- general path: the visitor is a point of a tainted voxel, or an ordinary voxel with a tainted neighbour.
          // A tainted voxel is visited point by point (ptIdx order); a point of one visits with its own neighbour list
          // (findVoxelNeighbors on its own index triple, ssc.cpp:311), computed here.  One lane walks the sequence. ----
          int rowv = -1;
          if (W < V) {
            rowv = (lane < 27) ? ring[c & 1][slot][lane] : -1;
          } else {
            const float4 q = ta.tp_xyz[(size_t)b * kTpCap + (W - V)];
            const BinResult br = dev_bin_point(q.x, q.y, q.z, ta.bp);
            if (lane < 27) {
              const int x = br.ri - 1 + lane / 9, y = br.si - 1 + (lane / 3) % 3, z = br.ei - 1 + lane % 3;
              if (!(x > ta.g.range_num - 1 || x < 0 || y > ta.g.sector_num - 1 || y < 0 || z > ta.g.azimuth_num - 1 || z < 0))
                rowv = vox_lookup(ta.bitmap + (size_t)b * ta.g.words, ta.word_rank + (size_t)b * ta.g.words, ta.g,
                                  x * ta.g.sector_num + y + z * ta.g.range_num * ta.g.sector_num);
            }
          }
          s_row[wid][lane] = rowv;
          __syncwarp();
          if (lane == 0) {
            atomicMin(&first_ev[W], e);
            const bool is_pt = W >= V;
            int oc = (state[W] == 2) ? find(W) : -1;
            bool skipped = false;
            auto visit = [&](int node) {
              if (state[node] == 0) {
                if (oc >= 0) {
                  parent[node] = oc;  // clusterIdxs[neighbor] = oc (:338)
                  state[node] = 2;
                } else {
                  skipped = true;
                }
              } else {
                const int r = find(node);
                if (oc < 0) {
                  oc = r;  // clusterIdxs[i] = nc (:334)
                } else if (r != oc) {
                  parent[oc] = r;  // mergeClusters(oc -> nc) (:329)
                  oc = r;
                }
                state[node] = 2;
              }
            };
            for (int k = 0; k < 27; ++k) {
              const int cN = s_row[wid][k];
              if (cN < 0) continue;
              const int tb = tnt[cN];
              if (tb >= 0) {
                const int n = ta.vox_cnt[base + cN];
                for (int j = 0; j < n; ++j) visit(V + tb + j);
              } else {
                visit(cN);
              }
            }
            if (oc < 0) {  // a new class (:345-351)
              atomicOr(&flags[e >> 5], 1u << (e & 31));
              parent[W] = W;
              setname[W] = e;
              state[W] = 2;
              stable[W] = 1;
              for (int k = 0; k < 27; ++k) {
                const int cN = s_row[wid][k];
                if (cN < 0) continue;
                const int tb = tnt[cN];
                if (tb >= 0) {
                  const int n = ta.vox_cnt[base + cN];
                  for (int j = 0; j < n; ++j)
                    if (V + tb + j != W) {
                      parent[V + tb + j] = W;
                      state[V + tb + j] = 2;
                    }
                } else if (cN != W) {
                  parent[cN] = W;
                  state[cN] = 2;
                }
              }
            } else if (is_pt) {
              if (state[W] == 0) {
                state[W] = 2;
                parent[W] = oc;
              }
            } else {
              if (state[W] == 0) {  // only this (first) point of W got the label
                state[W] = 1;
                parent[W] = oc;
              }
              stable[W] = skipped ? 0 : 1;
            }
          }
          __syncwarp();
          continue;
        }
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
  for (int v = tid; v < NV; v += T) {
    int r = v;  // read-only walk: other threads resolve voxels of the same component at the same time
    while (parent[r] != r) r = parent[r];
    const int ec = setname[r];
    int nm = -1;
    if (ec >= 0) nm = 5 + fprefix[ec >> 5] + __popc(flags[ec >> 5] & ((1u << (ec & 31)) - 1u));
    if (v < V)
      vox_name[base + v] = nm;  // -1 for a tainted voxel: its points carry the names
    else
      ta.tp_name[(size_t)b * kTpCap + (v - V)] = nm;
    if (nm >= 0 && nm < name_cap) atomicMin(&nf[nm], first_ev[v]);
  }
  if (tid == 0) scan_counts[b * 8 + 6] = cluster_name;
}

// ordered compaction of the apri points whose rank inside their voxel is < 3 ("clustering events").  One CTA per scan;
// every warp owns a contiguous slice: count (ballot / popcount), ONE block scan over the 32 warp totals, then the same
// walk again writing at the warp's offset — two barriers per scan instead of three per 1024 points.
// Every point of a tainted voxel is an event of its own: payload V + (offset of the voxel's points in tp_*) + rank.
__global__ void __launch_bounds__(1024) k_events(const int64_t* __restrict__ off, int32_t* __restrict__ scan_counts,
                                                 const int32_t* __restrict__ apri_cid, const int32_t* __restrict__ apri_rank,
                                                 const int32_t* __restrict__ taint_cnt, const int32_t* __restrict__ vox_tnt,
                                                 int32_t* __restrict__ ev_cid) {
  __shared__ int s_w[33];
  const int b = blockIdx.x;
  const int64_t base = off[b];
  const int M = scan_counts[b * 8 + 2];
  const int V = scan_counts[b * 8 + 3];
  const bool has_taint = taint_cnt[b * kTaintCntStride + 2] > 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int slice = (((M + 31) / 32) + 31) & ~31;  // per warp, multiple of 32
  const int m0 = min(M, wid * slice), m1 = min(M, m0 + slice);
  auto payload = [&](int m) {  // -1: no event
    const int cid = apri_cid[base + m];
    if (cid < 0) return -1;
    const int rk = apri_rank[base + m];
    if (has_taint) {
      const int tb = vox_tnt[base + cid];
      if (tb >= 0) return V + tb + rk;
    }
    return rk < 3 ? cid : -1;
  };
  int cnt = 0;
  for (int m = m0 + lane; m < m1; m += 32) cnt += (payload(m) >= 0) ? 1 : 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
  int total;
  const int ex = block_excl_scan<1024>(lane == 0 ? cnt : 0, &total, s_w);  // lane 0 of warp w carries the warp total
  int pos = __shfl_sync(0xffffffffu, ex, 0);
  for (int j = m0; j < m1; j += 32) {
    const int m = j + lane;
    const int cid = (m < m1) ? payload(m) : -1;
    const bool f = cid >= 0;
    const unsigned mask = __ballot_sync(0xffffffffu, f);
    if (f) ev_cid[base + pos + __popc(mask & ((1u << lane) - 1u))] = cid;
    pos += __popc(mask);
  }
  if (threadIdx.x == 0) scan_counts[b * 8 + 5] = total;
}

int launch_descriptor(const HostParams& hp, BatchDev& d, int nscans, int max_scan_points, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  BinParams bp = make_bin_params(hp);
  cudaMemsetAsync(d.bitmap, 0, sizeof(uint32_t) * (size_t)nscans * hp.g.words, st);
  dim3 gpt(grid_x_for(nscans, max_scan_points, 256), nscans);
  { TIMED("k_vox_mark", TSTREAM); k_vox_mark<<<gpt, 256, 0, st>>>(d.off, d.scan_counts, d.apri_vid, hp.g, d.bitmap); }
  { TIMED("k_vox_rank", TSTREAM); k_vox_rank<<<nscans, 1024, 0, st>>>(d.off, hp.g, d.bitmap, d.word_rank, d.vox_vid, d.vox_cnt, d.vox_tnt, d.scan_counts); }
  { TIMED("k_vox_count", TSTREAM); k_vox_count<<<gpt, 256, 0, st>>>(d.off, d.scan_counts, d.apri_vid, hp.g, d.bitmap, d.word_rank, d.apri_cid, d.vox_cnt); }
  { TIMED("k_taint_mark", TSTREAM); k_taint_mark<<<nscans, 256, 0, st>>>(d.off, d.taint_cnt, d.q_list, d.apri_cid, d.vox_tnt); }
  { TIMED("k_vox_offsets", TSTREAM); k_vox_offsets<<<nscans, 1024, 0, st>>>(d.off, d.scan_counts, d.vox_cnt, d.vox_off, d.vox_cur); }
  { TIMED("k_vox_fill", TSTREAM); k_vox_fill<<<gpt, 256, 0, st>>>(d.off, d.scan_counts, d.apri_cid, d.vox_off, d.vox_cur, d.vox_pts_tmp); }
  dim3 gv(grid_x_for(nscans, max_scan_points / 4 + 1, 8), nscans);  // one warp per voxel, 8 warps per CTA
  { TIMED("k_vox_stats", TSTREAM); k_vox_stats<<<gv, 256, 0, st>>>(d.off, d.scan_counts, d.vox_cnt, d.vox_off, d.vox_pts_tmp, d.apri_xyzi, d.vox_pts,
                                  d.apri_rank, d.vox_av, d.vox_cov, d.vox_bbox); }
  dim3 gc(grid_x_for(nscans, max_scan_points / 4 + 1, 256), nscans);
  { TIMED("k_vox_center", TSTREAM); k_vox_center<<<gc, 256, 0, st>>>(d.off, d.scan_counts, d.vox_off, d.vox_pts, d.apri_xyzi, bp, hp.p, d.vox_center, d.vox_tri); }
  { TIMED("k_taint_build", TSTREAM); k_taint_build<<<nscans, 1024, 0, st>>>(d.off, d.taint_cnt, d.q_list, d.apri_cid, d.vox_cnt, d.vox_off, d.vox_pts, d.apri_xyzi, d.vox_tnt,
                                       d.tv_cid, d.tv_base, d.tp_m, d.tp_cid, d.tp_xyz); }
  return 9;
}

int launch_cluster_prep(const HostParams& hp, BatchDev& d, int nscans, int max_scan_points, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  cudaMemsetAsync(d.edge_hash, 0xff, sizeof(unsigned long long) * (size_t)nscans * d.hash_cap, st);
  dim3 gv(grid_x_for(nscans, max_scan_points / 4 + 1, 256), nscans);
  { TIMED("k_vox_nbr", TSTREAM); k_vox_nbr<<<gv, 256, 0, st>>>(d.off, d.scan_counts, hp.g, d.bitmap, d.word_rank, d.vox_tri, d.vox_nbr, d.vox_root, d.taint_cnt, d.vox_tnt); }
  { TIMED("k_ccl_union", TSTREAM); k_ccl_union<<<gv, 256, 0, st>>>(d.off, d.scan_counts, d.vox_nbr, d.vox_root, d.taint_cnt, d.vox_tnt); }
  { TIMED("k_ccl_flatten", TSTREAM); k_ccl_flatten<<<gv, 256, 0, st>>>(d.off, d.scan_counts, d.vox_root); }
  dim3 gw(grid_x_for(nscans, max_scan_points / 4 + 1, 8), nscans);  // one warp per voxel, 8 warps per CTA
  { TIMED("k_similar_edges", TSTREAM); k_similar_edges<<<gw, 256, 0, st>>>(d.off, d.scan_counts, hp.g, d.bitmap, d.word_rank, d.vox_tri, d.vox_av, d.vox_cov, d.vox_root,
                                      hp.p.search_c, hp.p.intensity_cov, hp.p.intensity_diff,
                                      reinterpret_cast<unsigned long long*>(d.edge_hash), d.hash_cap, d.edge_buf, d.edge_cap); }
  { TIMED("k_events", TSTREAM); k_events<<<nscans, 1024, 0, st>>>(d.off, d.scan_counts, d.apri_cid, d.apri_rank, d.taint_cnt, d.vox_tnt, d.ev_cid); }
  return 5;
}

int launch_name_replay(const HostParams& hp, BatchDev& d, int nscans, int max_nodes, int max_events, bool force_global, bool any_taint,
                       int32_t* vox_name, int32_t* name_first, int name_cap, void* stream_) {
  if (nscans <= 0) return 0;
  constexpr int NW = kReplayWarps;
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t ring = sizeof(int) * NW * 2 * kReplayRows * 32;
  const size_t nfw = ((size_t)max_events + 31) / 32;
  const size_t smem = ring + (size_t)max_nodes * 6 + nfw * 8 + 16;
  static std::once_flag replay_once;
  std::call_once(replay_once, [ring] {
    cudaFuncSetAttribute(k_name_replay<NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(k_name_replay<NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
  });
  int launches = 1;
  const int32_t* group = d.vox_root;
  if (any_taint) {  // some scan of the batch has tainted voxels: events interact across ordinary components through them
    TIMED("k_taint_group", TSTREAM);
    k_taint_group<<<nscans, 1024, 0, st>>>(d.off, d.scan_counts, d.taint_cnt, hp.g, make_bin_params(hp), d.bitmap, d.word_rank, d.vox_nbr, d.vox_root,
                                           d.vox_tnt, d.tp_cid, d.tp_xyz, d.vox_group);
    group = d.vox_group;
    ++launches;
  }
  TaintArgs ta;
  ta.taint_cnt = d.taint_cnt;
  ta.vox_tnt = d.vox_tnt;
  ta.vox_cnt = d.vox_cnt;
  ta.tp_cid = d.tp_cid;
  ta.tp_xyz = d.tp_xyz;
  ta.tp_name = d.tp_name;
  ta.bitmap = d.bitmap;
  ta.word_rank = d.word_rank;
  ta.g = hp.g;
  ta.bp = make_bin_params(hp);
  int2* ev_list = reinterpret_cast<int2*>(d.bucket_kv);  // the ground stage is done with its (key, index) buckets
  if (smem <= 220 * 1024 && !force_global) {
    { TIMED("k_name_replay", TSTREAM); k_name_replay<NW, false><<<nscans, NW * 32, smem, st>>>(d.off, d.scan_counts, d.ev_cid, group, d.vox_nbr, ev_list, nullptr, d.vox_pts_tmp, d.apri_rank, nullptr, nullptr, vox_name, name_first, name_cap, ta); }
  } else {  // very dense scans: union-find state in (L2-resident) global scratch that the earlier stages are done with
    { TIMED("k_name_replay_global", TSTREAM); k_name_replay<NW, true><<<nscans, NW * 32, ring, st>>>(d.off, d.scan_counts, d.ev_cid, group, d.vox_nbr, ev_list, d.vox_cur, d.vox_pts_tmp, d.apri_rank, d.sorted_idx, d.slot_pos, vox_name, name_first, name_cap, ta); }
  }
  return launches;
}

}  // namespace scvod
