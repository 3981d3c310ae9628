// Probe: host<->device round-trip latency of a small kernel, stream synchronisation vs polling a word that the kernel
// writes into pinned host memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o poll_probe poll_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <immintrin.h>

__global__ void k_work(volatile int* flag, int* ctr, int value, int spin, int mode) {
  // emulate ~spin cycles of work in every CTA, last CTA publishes
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  __threadfence();
  __shared__ int last;
  if (threadIdx.x == 0) last = atomicAdd(ctr, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) {
    *ctr = 0;
    if (mode == 1) __threadfence_system();
    *flag = value;
    if (mode == 2) __threadfence_system();
  }
}

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  int *flag_h, *ctr, *buf_h, *buf_d;
  cudaMallocHost(&flag_h, 4096);
  cudaMallocHost(&buf_h, 65536);
  cudaMalloc(&buf_d, 65536);
  cudaMalloc(&ctr, 4);
  cudaMemset(ctr, 0, 4);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const int iters = 300;
  for (int grid : {1, 200}) {
    for (int spin : {2000, 40000}) {
      for (int with_copy = 0; with_copy < 2; ++with_copy) {
        // (a) stream synchronize
        double t0 = now_us();
        for (int i = 0; i < iters; ++i) {
          if (with_copy) cudaMemcpyAsync(buf_d, buf_h, 16384, cudaMemcpyHostToDevice, st);
          k_work<<<grid, 256, 0, st>>>(flag_h, ctr, i, spin, 0);
          cudaStreamSynchronize(st);
        }
        double a = (now_us() - t0) / iters;
        for (int mode = 0; mode < 3; ++mode) {
          t0 = now_us();
          long long spins = 0;
          for (int i = 0; i < iters; ++i) {
            *(volatile int*)flag_h = -1;
            if (with_copy) cudaMemcpyAsync(buf_d, buf_h, 16384, cudaMemcpyHostToDevice, st);
            k_work<<<grid, 256, 0, st>>>(flag_h, ctr, i, spin, mode);
            while (*(volatile int*)flag_h < 0) { _mm_pause(); ++spins; }
          }
          double b = (now_us() - t0) / iters;
          printf("grid %3d spin %5d copy %d : sync %.1f us | poll(mode %d) %.1f us (%.0f spins/iter)\n", grid, spin, with_copy, a, mode, b, (double)spins / iters);
        }
        // (c) poll with cudaStreamQuery
        t0 = now_us();
        for (int i = 0; i < iters; ++i) {
          if (with_copy) cudaMemcpyAsync(buf_d, buf_h, 16384, cudaMemcpyHostToDevice, st);
          k_work<<<grid, 256, 0, st>>>(flag_h, ctr, i, spin, 0);
          while (cudaStreamQuery(st) == cudaErrorNotReady) {}
        }
        printf("   streamQuery poll %.1f us\n", (now_us() - t0) / iters);
      }
    }
  }
  cudaStreamSynchronize(st);
  return 0;
}
