// scvod_grid.cuh — a uniform search grid over a device cloud (cell-major copy of the points + cell start offsets), shared by the
// evaluation kernels (scvod_eval.cu) and the k-nearest-neighbour kernels (scvod_knn.cu).  Included inside namespace scvod { namespace {.
#pragma once

#define ECU(call)                                                                                                \
  do {                                                                                                           \
    cudaError_t e__ = (call);                                                                                    \
    if (e__ != cudaSuccess) return api_fail(SCVOD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

constexpr int kEvalMaxCells = 1 << 24;

struct EGrid {
  float ox, oy, oz, h;
  int nx, ny, nz, ncells;
};

struct DTmp {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  ~DTmp() {
    if (p) cudaFree(p);
  }
  template <typename T>
  T* as() {
    return (T*)p;
  }
};

__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
inline float ord2f_host(int i) {
  int j = i >= 0 ? i : i ^ 0x7fffffff;
  float f;
  std::memcpy(&f, &j, 4);
  return f;
}

__global__ void __launch_bounds__(256) k_eval_bbox(const float4* __restrict__ pts, long long n, int* __restrict__ box) {
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(&pts[i]);
    const int a = f2ord(p.x), b = f2ord(p.y), c = f2ord(p.z);
    lo[0] = min(lo[0], a); lo[1] = min(lo[1], b); lo[2] = min(lo[2], c);
    hi[0] = max(hi[0], a); hi[1] = max(hi[1], b); hi[2] = max(hi[2], c);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&box[d], lo[d]);
      atomicMax(&box[3 + d], hi[d]);
    }
  }
}

__device__ __forceinline__ void cell_coords(const EGrid& g, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = (int)floorf(__fdiv_rn(__fsub_rn(x, g.ox), g.h));
  cy = (int)floorf(__fdiv_rn(__fsub_rn(y, g.oy), g.h));
  cz = (int)floorf(__fdiv_rn(__fsub_rn(z, g.oz), g.h));
}

__global__ void __launch_bounds__(256) k_eval_count(const float4* __restrict__ pts, long long n, EGrid g, int* __restrict__ cell_of,
                                                    int* __restrict__ slot, int* __restrict__ cnt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(&pts[i]);
    int cx, cy, cz;
    cell_coords(g, p.x, p.y, p.z, cx, cy, cz);
    const int c = (cx * g.ny + cy) * g.nz + cz;  // every target point lies inside its own bounding box
    cell_of[i] = c;
    slot[i] = atomicAdd(&cnt[c], 1);
  }
}

__global__ void __launch_bounds__(1024) k_eval_scan_blocks(const int* __restrict__ cnt, int L, int* __restrict__ start, int* __restrict__ block_sum) {
  __shared__ int s_w[33];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int v = (i < L) ? cnt[i] : 0;
  int total;
  const int ex = block_excl_scan<1024>(v, &total, s_w);
  if (i < L) start[i] = ex;
  if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_eval_scan_sums(int* __restrict__ block_sum, int nblocks, int* __restrict__ start, int L) {
  __shared__ int s_w[33];
  int carry = 0;
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int v = (b < nblocks) ? block_sum[b] : 0;
    int total;
    const int ex = block_excl_scan<1024>(v, &total, s_w);
    if (b < nblocks) block_sum[b] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) start[L] = carry;
}
__global__ void __launch_bounds__(1024) k_eval_scan_add(int* __restrict__ start, int L, const int* __restrict__ block_sum) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i < L) start[i] += block_sum[blockIdx.x];
}

__global__ void __launch_bounds__(256) k_eval_fill(const float4* __restrict__ pts, long long n, const int* __restrict__ cell_of,
                                                   const int* __restrict__ slot, const int* __restrict__ start, float4* __restrict__ sorted,
                                                   int* __restrict__ sorted_idx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int pos = start[cell_of[i]] + slot[i];
    sorted[pos] = __ldg(&pts[i]);
    sorted_idx[pos] = (int)i;
  }
}

struct BuiltGrid {
  EGrid g;
  DTmp sorted, sorted_idx, start;
  long long n = 0;
};

int grid_blocks(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)num_sms() * 16)); }

// uniform grid over the bounding box of `pts_dev` with cells of edge >= r (doubled while the grid would be too large)
int build_grid(scvod_ctx* c, const float4* pts_dev, long long n, float r, BuiltGrid& out) {
  cudaStream_t st = (cudaStream_t)ctx_stream(c);
  out.n = n;
  out.g = EGrid{0.f, 0.f, 0.f, r, 1, 1, 1, 1};
  if (n <= 0) {
    ECU(out.start.alloc(sizeof(int) * 2));
    ECU(cudaMemsetAsync(out.start.p, 0, sizeof(int) * 2, st));
    return SCVOD_OK;
  }
  DTmp box, cell_of, slot, cnt, bsum;
  ECU(box.alloc(sizeof(int) * 6));
  const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  ECU(cudaMemcpyAsync(box.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  { void* stream_ = (void*)st; TIMED("k_eval_bbox", TSTREAM); k_eval_bbox<<<grid_blocks(n), 256, 0, st>>>(pts_dev, n, box.as<int>()); }
  int hb[6];
  ECU(cudaMemcpyAsync(hb, box.p, sizeof(hb), cudaMemcpyDeviceToHost, st));
  ECU(cudaStreamSynchronize(st));
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = ord2f_host(hb[a]);
    hi[a] = ord2f_host(hb[3 + a]);
    if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return api_fail(SCVOD_ERR_ARG, "evaluation cloud holds non-finite coordinates");
  }
  EGrid g;
  float h = r;
  for (;;) {
    g.h = h;
    g.ox = lo[0];
    g.oy = lo[1];
    g.oz = lo[2];
    g.nx = (int)floorf((hi[0] - g.ox) / h) + 1;
    g.ny = (int)floorf((hi[1] - g.oy) / h) + 1;
    g.nz = (int)floorf((hi[2] - g.oz) / h) + 1;
    if ((double)g.nx * g.ny * g.nz <= (double)kEvalMaxCells) break;
    h *= 2.f;
  }
  g.ncells = g.nx * g.ny * g.nz;
  out.g = g;
  ECU(cell_of.alloc(sizeof(int) * (size_t)n));
  ECU(slot.alloc(sizeof(int) * (size_t)n));
  ECU(cnt.alloc(sizeof(int) * ((size_t)g.ncells + 1)));
  ECU(out.start.alloc(sizeof(int) * ((size_t)g.ncells + 2)));
  ECU(out.sorted.alloc(sizeof(float4) * (size_t)n));
  ECU(out.sorted_idx.alloc(sizeof(int) * (size_t)n));
  const int nb = (g.ncells + 1023) / 1024;
  ECU(bsum.alloc(sizeof(int) * ((size_t)nb + 1)));
  ECU(cudaMemsetAsync(cnt.p, 0, sizeof(int) * ((size_t)g.ncells + 1), st));
  {
    void* stream_ = (void*)st;
    TIMED("k_eval_grid_build", TSTREAM);
    k_eval_count<<<grid_blocks(n), 256, 0, st>>>(pts_dev, n, g, cell_of.as<int>(), slot.as<int>(), cnt.as<int>());
    k_eval_scan_blocks<<<nb, 1024, 0, st>>>(cnt.as<int>(), g.ncells, out.start.as<int>(), bsum.as<int>());
    k_eval_scan_sums<<<1, 1024, 0, st>>>(bsum.as<int>(), nb, out.start.as<int>(), g.ncells);
    k_eval_scan_add<<<nb, 1024, 0, st>>>(out.start.as<int>(), g.ncells, bsum.as<int>());
    k_eval_fill<<<grid_blocks(n), 256, 0, st>>>(pts_dev, n, cell_of.as<int>(), slot.as<int>(), out.start.as<int>(), out.sorted.as<float4>(),
                                                out.sorted_idx.as<int>());
  }
  ctx_add_launches(c, 6);
  ECU(cudaGetLastError());
  ECU(cudaStreamSynchronize(st));  // the temporaries go out of scope
  return SCVOD_OK;
}

