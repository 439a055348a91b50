// microbenchmark: random 16-byte loads and 64-bit CAS into a table of a given size (L2-resident or not), R independent
// accesses in flight per thread.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 l2_random.cu -o l2_random
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
template <int R, int MODE>
__global__ void k(ulonglong2* table, uint32_t slots, int iters, unsigned long long* sink) {
  unsigned long long acc = 0;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    uint32_t s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = __umulhi((uint32_t)mix64(t * 131 + it * R + r + 1), slots);
    if (MODE == 0) {  // 16-byte loads
      ulonglong2 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(v[r].x), "=l"(v[r].y) : "l"(table + s[r]));
#pragma unroll
      for (int r = 0; r < R; ++r) acc += v[r].x + v[r].y;
    } else if (MODE == 1) {  // 64-bit CAS (always fails after the first touch: compare with a value never stored)
      unsigned long long o[R];
#pragma unroll
      for (int r = 0; r < R; ++r) o[r] = atomicCAS(&table[s[r]].x, 0x123456789ull, t);
#pragma unroll
      for (int r = 0; r < R; ++r) acc += o[r];
    } else if (MODE == 2) {  // 8-byte stores (fire and forget)
#pragma unroll
      for (int r = 0; r < R; ++r) table[s[r]].y = t;
    } else {  // 32-bit CAS
      unsigned o[R];
#pragma unroll
      for (int r = 0; r < R; ++r) o[r] = atomicCAS((unsigned*)&table[s[r]].x, 0x12345u, (unsigned)t);
#pragma unroll
      for (int r = 0; r < R; ++r) acc += o[r];
    }
  }
  if (acc == 0x1234567) sink[0] = acc;
}
template <int R, int MODE>
void run(ulonglong2* table, uint32_t slots, unsigned long long* sink, int ctas_per_sm, const char* what) {
  const int iters = 64 / R * 4, grid = 148 * ctas_per_sm, block = 256;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<R, MODE><<<grid, block>>>(table, slots, iters, sink);
  cudaEventRecord(a);
  k<R, MODE><<<grid, block>>>(table, slots, iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double n = (double)grid * block * iters * R;
  printf("%-10s table %5.0f MB  R=%d  %d CTAs/SM: %7.1f G accesses/s  (%.3f ms for %.0f M)\n", what, slots * 16.0 / 1e6, R, ctas_per_sm, n / ms / 1e6, ms, n / 1e6);
}
int main() {
  ulonglong2* table; unsigned long long* sink;
  const size_t max_bytes = 1ull << 30;
  cudaMalloc(&table, max_bytes); cudaMemset(table, 0xff, max_bytes); cudaMalloc(&sink, 64);
  for (size_t mb : {8, 32, 96, 512}) {
    const uint32_t slots = (uint32_t)(mb * 1000000 / 16);
    run<4, 0>(table, slots, sink, 4, "ld16");
    run<4, 0>(table, slots, sink, 8, "ld16");
    run<8, 0>(table, slots, sink, 4, "ld16");
    run<8, 0>(table, slots, sink, 8, "ld16");
    run<4, 1>(table, slots, sink, 8, "cas64");
    run<8, 1>(table, slots, sink, 8, "cas64");
    run<8, 3>(table, slots, sink, 8, "cas32");
    run<8, 2>(table, slots, sink, 8, "st8");
  }
  return 0;
}
