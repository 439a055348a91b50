// microbenchmark: throughput of same-address global atomics issued by one thread per CTA (the per-tile output
// reservation pattern), vs atomics spread over many addresses, vs none.   nvcc -arch=sm_100a -O3 atomics.cu -o atomics
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned long long* cur, int iters, int n_addr, int mode, unsigned long long* sink) {
  __shared__ unsigned long long s_base;
  unsigned long long acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (mode == 1) {
      if (threadIdx.x == 0) s_base = atomicAdd(cur + ((blockIdx.x + i) % n_addr) * 16, 1000ull);
      __syncthreads();
      acc += s_base;
      __syncthreads();
    } else if (mode == 2) {  // every warp its own atomic
      unsigned long long b = 0;
      if ((threadIdx.x & 31) == 0) b = atomicAdd(cur + ((blockIdx.x + i) % n_addr) * 16, 128ull);
      acc += __shfl_sync(0xffffffffu, b, 0);
    } else {
      __syncthreads();
      acc += i;
      __syncthreads();
    }
  }
  if (acc == 0x1234567) sink[0] = acc;
}
int main() {
  unsigned long long *cur, *sink;
  cudaMalloc(&cur, 1 << 20); cudaMemset(cur, 0, 1 << 20); cudaMalloc(&sink, 64);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 592, iters = 200;
  for (int mode = 0; mode <= 2; ++mode)
    for (int n_addr : {1, 8, 64, 1024}) {
      if (mode == 0 && n_addr > 1) continue;
      k<<<grid, 256>>>(cur, iters, n_addr, mode, sink);
      cudaEventRecord(a);
      k<<<grid, 256>>>(cur, iters, n_addr, mode, sink);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double n_at = (double)grid * iters * (mode == 2 ? 8 : 1);
      printf("mode %d (0 none, 1 per-CTA, 2 per-warp) addrs %4d: %.3f ms, %.1f ns per atomic (serial view), %.2f us per CTA iteration\n", mode, n_addr, ms,
             ms * 1e6 / n_at, ms * 1e3 / iters);
    }
  return 0;
}
