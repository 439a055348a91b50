// smem_atomics.cu — throughput of shared-memory atomics on random slots (what the partitioned GROUP BY's table costs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int S = 1024;
__device__ __forceinline__ uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 10; }

template <int KIND>
__global__ void __launch_bounds__(256) k(unsigned long long* out, int iters) {
  __shared__ unsigned long long tab[S * 2];
  for (int i = threadIdx.x; i < S * 2; i += 256) tab[i] = KIND == 1 ? ~0ull : 0ull;
  __syncthreads();
  uint32_t s = blockIdx.x * 256 + threadIdx.x + 1;
  unsigned long long acc = 0;
  for (int i = 0; i < iters; ++i) {
    const uint32_t a = rnd(s) & (S - 1);
    if (KIND == 0) acc += atomicAdd((unsigned int*)&tab[a], 1u);                       // native 32-bit add
    if (KIND == 1) acc += atomicCAS(&tab[a], ~0ull, (unsigned long long)a);              // 64-bit CAS
    if (KIND == 2) atomicAdd((double*)&tab[a], 1.0);                                    // f64 add (CAS loop)
    if (KIND == 3) atomicMin((long long*)&tab[a], (long long)s);                        // s64 min (CAS loop)
    if (KIND == 4) atomicAdd(&tab[a], 1ull);                                            // u64 add (CAS loop)
    if (KIND == 5) { tab[a] += 1; }                                                     // plain RMW (racy; the no-atomic floor)
    if (KIND == 6) acc += tab[a];                                                       // plain 64-bit load
  }
  if (acc == 0x1234567) out[0] = acc;
  if (threadIdx.x == 0) out[1 + blockIdx.x % 8] = tab[3];
}

template <int KIND>
void run(const char* name) {
  unsigned long long* out;
  cudaMalloc(&out, 1024);
  const int iters = 20000, grid = 148 * 3;
  k<KIND><<<grid, 256>>>(out, 100);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<KIND><<<grid, 256>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)grid * 256 * iters;
  printf("%-28s %8.3f ms  %8.1f G ops/s  (%.2f ops/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}

int main() {
  run<0>("atomicAdd u32 (ATOMS.ADD)");
  run<1>("atomicCAS u64");
  run<2>("atomicAdd f64 (CAS loop)");
  run<3>("atomicMin s64 (CAS loop)");
  run<4>("atomicAdd u64 (CAS loop)");
  run<5>("plain RMW u64");
  run<6>("plain load u64");
  return 0;
}
