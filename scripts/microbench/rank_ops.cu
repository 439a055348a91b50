// rank_ops.cu — cost of ranking rows inside a warp by region: MATCH.ANY vs a shared-memory atomicAdd that returns the rank.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rank_ops rank_ops.cu && ./rank_ops
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 9; }

template <int KIND>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, uint32_t parts) {
  __shared__ uint32_t hist[8][257];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 257; i += 32) hist[warp][i] = 0;
  __syncwarp();
  uint32_t s = blockIdx.x * 256 + threadIdx.x + 1, acc = 0;
  for (int i = 0; i < iters; ++i) {
    const uint32_t p = rnd(s) % parts;
    if (KIND == 0) {  // MATCH.ANY + leader RMW
      const uint32_t peers = __match_any_sync(0xffffffffu, p);
      const uint32_t old = hist[warp][p];
      __syncwarp();
      if (lane == __ffs(peers) - 1) hist[warp][p] = old + __popc(peers);
      acc += old + __popc(peers & ((1u << lane) - 1));
      __syncwarp();
    } else if (KIND == 1) {  // warp-private histogram, atomicAdd returns the rank
      acc += atomicAdd(&hist[warp][p], 1u);
    } else if (KIND == 2) {  // MATCH.ANY alone
      acc += __match_any_sync(0xffffffffu, p);
    } else {  // RED (no return)
      atomicAdd(&hist[warp][p], 1u);
    }
  }
  if (acc == 0x12345) out[0] = acc;
  if (lane == 0) out[1 + blockIdx.x % 8] = hist[warp][3];
}

template <int KIND>
void run(const char* name, uint32_t parts) {
  unsigned* out;
  cudaMalloc(&out, 1024);
  const int iters = 20000, grid = 148 * 3;
  k<KIND><<<grid, 256>>>(out, 100, parts);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<KIND><<<grid, 256>>>(out, iters, parts);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)grid * 256 * iters;
  printf("%-34s parts %3u %8.3f ms %8.1f G rows/s (%.2f cycles per warp-op per SM)\n", name, parts, ms, ops / ms / 1e6,
         1.9e6 * ms / (ops / 32 / 148));
  cudaFree(out);
}

int main() {
  for (uint32_t parts : {29u, 58u, 128u, 256u}) {
    run<0>("match.any + leader RMW", parts);
    run<1>("smem atomicAdd returning rank", parts);
    run<2>("match.any alone", parts);
    run<3>("smem RED (no return)", parts);
  }
  return 0;
}
