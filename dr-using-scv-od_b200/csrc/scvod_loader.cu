// scvod_loader.cu — the front end of the reference's KITTI loader on the device: SSC::getCloud
// (reference src/ssc.cpp:1060-1111) drops the points whose SemanticKITTI label is 0 or 1 (:1063), scales the intensity by
// max_intensity (:1071), keeps every point (the distance test `dis >= min_dis || dis <= max_dis` at :1096 is always true) and
// downsamples with pcl::VoxelGrid<PointXYZI>, leaf 0.08 m (:1108-1111).
//
// PCL 1.8 filters/impl/voxel_grid.hpp [recollection; the same restatement as host/include/voxel_grid.h]: float bounding box,
// inverse_leaf = 1 / leaf in float, min_b = floor(min * inverse_leaf), leaf coordinate (int)(floor(x * inverse_leaf) - (float)min_b),
// linear index ijk0 + ijk1 * div_b0 + ijk2 * div_b0 * div_b1, points sorted by that index, ONE output point per occupied leaf in
// ascending index = float sums of x, y, z, intensity divided by the count.
//
// What is exact and what is not: the mask, the scaled intensity, the leaf index of every point, the set and order of output leaves
// and their point counts are exact.  The centroid is a float sum whose order the reference leaves to std::sort (introsort is
// unstable: the order of a leaf's points is implementation defined); here a leaf's points are summed in ascending input index
// (a stable LSD radix sort), so a centroid of a leaf with >= 3 points may differ from the host restatement in the last bits
// (leaves with 1 or 2 points are bit-identical).  tests/test_gpu_loader.py measures both.
//
// One CTA per scan: ordered compaction of the kept points + bounding box, key computation, 8-bit LSD radix sort of (key, index)
// through global scratch with per-tile warp-match ranking, ordered compaction of the run heads + sequential sums.
#include "scvod_kernel_common.cuh"

namespace scvod {

namespace {

constexpr int kLoadThreads = 1024;

struct LoaderScratch {
  float4* kept;       // masked points, intensity scaled, input order
  uint32_t* key[2];   // leaf index of every kept point (ping-pong)
  int32_t* idx[2];    // index into kept (ping-pong)
};

__device__ __forceinline__ float block_reduce_minmax(float v, bool is_min, float* s_red /* 32 floats */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_min ? fminf(v, t) : fmaxf(v, t);
  }
  if (lane == 0) s_red[w] = v;
  __syncthreads();
  if (w == 0) {
    float x = s_red[lane];  // 32 warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, x, o);
      x = is_min ? fminf(x, t) : fmaxf(x, t);
    }
    if (lane == 0) s_red[0] = x;
  }
  __syncthreads();
  const float r = s_red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kLoadThreads) k_loader_voxelgrid(const float4* __restrict__ raw, const uint32_t* __restrict__ labels,
                                                                   const int64_t* __restrict__ off, float inv_leaf, float max_intensity,
                                                                   LoaderScratch sc, float4* __restrict__ out, int32_t* __restrict__ out_cnt /* [scans][4] */) {
  __shared__ int s_w[33];
  __shared__ float s_red[32];
  __shared__ int s_tab[32][256];
  __shared__ int s_base[256];
  __shared__ int s_geo[8];  // min_b[3], mul[3], nbits, passthrough
  const int b = blockIdx.x;
  const int64_t base = off[b] - off[0];
  const int n = (int)(off[b + 1] - off[b]);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float4* kept = sc.kept + base;

  // ---- phase 0: label mask (ssc.cpp:1063), intensity scale (:1071), ordered compaction, bounding box ----
  float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, hi[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
  int m = 0;
  for (int t0 = 0; t0 < n; t0 += kLoadThreads) {
    const int i = t0 + tid;
    bool keep = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      const uint32_t sem = labels ? (labels[base + i] & 0xFFFFu) : 2u;
      keep = !(sem == 0u || sem == 1u);
      if (keep) {
        p = raw[base + i];
        p.w = __fmul_rn(p.w, max_intensity);
      }
    }
    int total;
    const int ex = block_excl_scan<kLoadThreads>(keep ? 1 : 0, &total, s_w);
    if (keep) {
      kept[m + ex] = p;
      lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
      hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    m += total;
  }
  if (m == 0) {
    if (tid == 0) {
      out_cnt[4 * b] = 0;
      out_cnt[4 * b + 1] = 0;
      out_cnt[4 * b + 2] = 0;
    }
    return;
  }
  float mn[3], mx[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    mn[d] = block_reduce_minmax(lo[d], true, s_red);
    mx[d] = block_reduce_minmax(hi[d], false, s_red);
  }
  // ---- phase 1: grid geometry (voxel_grid.hpp: min_b_, div_b_, divb_mul_; leaf-size-too-small check) ----
  if (tid == 0) {
    long long dd[3];
    int min_b[3], div_b[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      dd[d] = (long long)__fmul_rn(__fsub_rn(mx[d], mn[d]), inv_leaf) + 1;
      min_b[d] = (int)floorf(__fmul_rn(mn[d], inv_leaf));
      div_b[d] = (int)floorf(__fmul_rn(mx[d], inv_leaf)) - min_b[d] + 1;
    }
    const bool pass = dd[0] * dd[1] * dd[2] > 2147483647LL;  // PCL warns and returns the input unchanged
    s_geo[0] = min_b[0];
    s_geo[1] = min_b[1];
    s_geo[2] = min_b[2];
    s_geo[3] = 1;
    s_geo[4] = div_b[0];
    s_geo[5] = div_b[0] * div_b[1];
    const long long cells = (long long)div_b[0] * div_b[1] * div_b[2];
    int nbits = 1;
    while (nbits < 32 && (1LL << nbits) < cells) ++nbits;
    s_geo[6] = nbits;
    s_geo[7] = pass ? 1 : 0;
  }
  __syncthreads();
  if (s_geo[7]) {
    for (int i = tid; i < m; i += kLoadThreads) out[base + i] = kept[i];
    if (tid == 0) {
      out_cnt[4 * b] = m;
      out_cnt[4 * b + 1] = m;
      out_cnt[4 * b + 2] = 1;
    }
    return;
  }
  uint32_t* kA = sc.key[0] + base;
  uint32_t* kB = sc.key[1] + base;
  int32_t* iA = sc.idx[0] + base;
  int32_t* iB = sc.idx[1] + base;
  {
    const float fb0 = (float)s_geo[0], fb1 = (float)s_geo[1], fb2 = (float)s_geo[2];
    const int mul1 = s_geo[4], mul2 = s_geo[5];
    for (int i = tid; i < m; i += kLoadThreads) {
      const float4 p = kept[i];
      const int ijk0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv_leaf)), fb0);
      const int ijk1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv_leaf)), fb1);
      const int ijk2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv_leaf)), fb2);
      kA[i] = (uint32_t)(ijk0 + ijk1 * mul1 + ijk2 * mul2);
      iA[i] = i;
    }
  }
  __syncthreads();
  // ---- phase 2: stable LSD radix sort of (key, index), 8 bits per pass ----
  const int npass = (s_geo[6] + 7) / 8;
  for (int pass = 0; pass < npass; ++pass) {
    const int shift = 8 * pass;
    if (tid < 256) s_base[tid] = 0;
    __syncthreads();
    for (int i = tid; i < m; i += kLoadThreads) atomicAdd(&s_base[(kA[i] >> shift) & 255u], 1);
    __syncthreads();
    {
      int total;
      const int v = tid < 256 ? s_base[tid] : 0;
      const int ex = block_excl_scan<kLoadThreads>(v, &total, s_w);
      if (tid < 256) s_base[tid] = ex;
    }
    __syncthreads();
    for (int t0 = 0; t0 < m; t0 += kLoadThreads) {
      int* tab = &s_tab[0][0];
#pragma unroll
      for (int k = 0; k < 8; ++k) tab[tid + k * kLoadThreads] = 0;
      __syncthreads();
      const int i = t0 + tid;
      const bool valid = i < m;
      uint32_t key = 0;
      int id = 0, d = 0x100 + lane;  // idle lanes match nobody
      if (valid) {
        key = kA[i];
        id = iA[i];
        d = (int)((key >> shift) & 255u);
      }
      const unsigned same = __match_any_sync(0xffffffffu, d);
      const int rank = __popc(same & ((1u << lane) - 1u));
      if (valid && rank == 0) s_tab[wid][d] = __popc(same);
      __syncthreads();
      if (tid < 256) {  // digit tid: exclusive prefix over the warps (tile order), then advance the digit's base
        int run = s_base[tid];
#pragma unroll 8
        for (int w = 0; w < 32; ++w) {
          const int c = s_tab[w][tid];
          s_tab[w][tid] = run;
          run += c;
        }
        s_base[tid] = run;
      }
      __syncthreads();
      if (valid) {
        const int pos = s_tab[wid][d] + rank;
        kB[pos] = key;
        iB[pos] = id;
      }
      __syncthreads();
    }
    uint32_t* tk = kA; kA = kB; kB = tk;
    int32_t* ti = iA; iA = iB; iB = ti;
  }
  // ---- phase 3: one output per run of equal keys, in ascending key; sums in ascending input index ----
  int nout = 0;
  for (int t0 = 0; t0 < m; t0 += kLoadThreads) {
    const int i = t0 + tid;
    const bool head = i < m && (i == 0 || kA[i] != kA[i - 1]);
    int total;
    const int ex = block_excl_scan<kLoadThreads>(head ? 1 : 0, &total, s_w);
    if (head) {
      const uint32_t key = kA[i];
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int j = i;
      for (; j < m && kA[j] == key; ++j) {
        const float4 p = kept[iA[j]];
        sx = __fadd_rn(sx, p.x);
        sy = __fadd_rn(sy, p.y);
        sz = __fadd_rn(sz, p.z);
        si = __fadd_rn(si, p.w);
      }
      const float cnt = (float)(j - i);
      out[base + nout + ex] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
    }
    nout += total;
  }
  if (tid == 0) {
    out_cnt[4 * b] = nout;
    out_cnt[4 * b + 1] = m;
    out_cnt[4 * b + 2] = 0;
  }
}

struct DevTmp {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  ~DevTmp() {
    if (p) cudaFree(p);
  }
};

}  // namespace

// raw / labels / out are device pointers; out has room for every input point (scan b is written at offsets[b] - offsets[0]);
// counts_host[4 * b] = output points, [4 * b + 1] = points that passed the label mask, [4 * b + 2] = 1 when the leaf grid
// overflowed an int and the masked scan was passed through (PCL's "leaf size too small" path)
int loader_run(scvod_ctx* c, const void* raw_dev, const uint32_t* labels_dev, const int64_t* offsets, int nscans, float leaf, float max_intensity,
               void* out_dev, int32_t* counts_host) {
  if (nscans <= 0) return SCVOD_OK;
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  const int64_t total = offsets[nscans] - offsets[0];
  DevTmp kept, k0, k1, i0, i1, off, cnt;
  if (kept.alloc(sizeof(float4) * (size_t)total) != cudaSuccess || k0.alloc(4 * (size_t)total) != cudaSuccess || k1.alloc(4 * (size_t)total) != cudaSuccess ||
      i0.alloc(4 * (size_t)total) != cudaSuccess || i1.alloc(4 * (size_t)total) != cudaSuccess || off.alloc(sizeof(int64_t) * (nscans + 1)) != cudaSuccess ||
      cnt.alloc(sizeof(int32_t) * 4 * (size_t)nscans) != cudaSuccess)
    return api_fail(SCVOD_ERR_CUDA, "loader: out of device memory");
  if (cudaMemcpyAsync(off.p, offsets, sizeof(int64_t) * (nscans + 1), cudaMemcpyHostToDevice, st) != cudaSuccess)
    return api_fail(SCVOD_ERR_CUDA, "loader: offsets upload failed");
  LoaderScratch sc;
  sc.kept = (float4*)kept.p;
  sc.key[0] = (uint32_t*)k0.p;
  sc.key[1] = (uint32_t*)k1.p;
  sc.idx[0] = (int32_t*)i0.p;
  sc.idx[1] = (int32_t*)i1.p;
  const float inv_leaf = 1.0f / leaf;  // Eigen::Array4f::Ones() / leaf_size_ (float)
  {
    void* stream_ = (void*)st;
    TIMED("k_loader_voxelgrid", TSTREAM);
    k_loader_voxelgrid<<<nscans, kLoadThreads, 0, st>>>((const float4*)raw_dev, labels_dev, (const int64_t*)off.p, inv_leaf, max_intensity, sc, (float4*)out_dev,
                                                       (int32_t*)cnt.p);
  }
  ctx_add_launches(c, 1);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts_host, cnt.p, sizeof(int32_t) * 4 * (size_t)nscans, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, std::string("loader: ") + cudaGetErrorString(e));
  return SCVOD_OK;
}

}  // namespace scvod

using namespace scvod;

extern "C" int scvod_load_kitti_dev(scvod_ctx* c, const void* raw_xyzi_dev, const uint32_t* labels_dev, const int64_t* offsets, int nscans, float leaf,
                                    float max_intensity, void* out_xyzi_dev, int32_t* counts4) {
  if (!c || !offsets || nscans < 0 || !counts4 || !(leaf > 0.f)) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_load_kitti_dev");
  if (nscans > 0 && offsets[nscans] > offsets[0] && (!raw_xyzi_dev || !out_xyzi_dev)) return api_fail(SCVOD_ERR_ARG, "null cloud pointer");
  if (cudaSetDevice(ctx_device(c)) != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, "cudaSetDevice failed");
  return loader_run(c, raw_xyzi_dev, labels_dev, offsets, nscans, leaf, max_intensity, out_xyzi_dev, counts4);
}

extern "C" int scvod_load_kitti(scvod_ctx* c, const float* raw_xyzi, const uint32_t* labels, const int64_t* offsets, int nscans, float leaf,
                                float max_intensity, float* out_xyzi, int64_t* out_offsets) {
  if (!c || !offsets || nscans < 0 || !out_offsets || !(leaf > 0.f)) return api_fail(SCVOD_ERR_ARG, "bad arguments to scvod_load_kitti");
  out_offsets[0] = 0;
  if (nscans == 0) return SCVOD_OK;
  const int64_t total = offsets[nscans] - offsets[0];
  if (total > 0 && (!raw_xyzi || !out_xyzi)) return api_fail(SCVOD_ERR_ARG, "null cloud pointer");
  if (cudaSetDevice(ctx_device(c)) != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  void *d_raw = nullptr, *d_lab = nullptr, *d_out = nullptr;
  auto cleanup = [&]() {
    if (d_raw) cudaFree(d_raw);
    if (d_lab) cudaFree(d_lab);
    if (d_out) cudaFree(d_out);
  };
  const size_t bytes = sizeof(float) * 4 * (size_t)std::max<int64_t>(total, 1);
  if (cudaMalloc(&d_raw, bytes) != cudaSuccess || cudaMalloc(&d_out, bytes) != cudaSuccess ||
      (labels && cudaMalloc(&d_lab, sizeof(uint32_t) * (size_t)std::max<int64_t>(total, 1)) != cudaSuccess)) {
    cleanup();
    return api_fail(SCVOD_ERR_CUDA, "loader: out of device memory");
  }
  cudaMemcpyAsync(d_raw, raw_xyzi + 4 * offsets[0], sizeof(float) * 4 * (size_t)total, cudaMemcpyHostToDevice, st);
  if (labels) cudaMemcpyAsync(d_lab, labels + offsets[0], sizeof(uint32_t) * (size_t)total, cudaMemcpyHostToDevice, st);
  std::vector<int32_t> cnt(4 * (size_t)nscans, 0);
  int rc = loader_run(c, d_raw, (const uint32_t*)d_lab, offsets, nscans, leaf, max_intensity, d_out, cnt.data());
  if (rc == SCVOD_OK) {
    for (int b = 0; b < nscans; ++b) {
      out_offsets[b + 1] = out_offsets[b] + cnt[4 * (size_t)b];
      if (cnt[4 * (size_t)b] > 0 &&
          cudaMemcpyAsync(out_xyzi + 4 * out_offsets[b], (const float*)d_out + 4 * (offsets[b] - offsets[0]), sizeof(float) * 4 * (size_t)cnt[4 * (size_t)b],
                          cudaMemcpyDeviceToHost, st) != cudaSuccess)
        rc = api_fail(SCVOD_ERR_CUDA, "loader: download failed");
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = api_fail(SCVOD_ERR_CUDA, "loader: download failed");
  }
  cleanup();
  return rc;
}
