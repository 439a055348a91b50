// sort_kernels.cu — device lexsort: the permutation that arrow's `lexsort_to_indices` gives Merge::sorted_merge
// (crates/query-distributed/src/operators.rs:141-194), as a least-significant-digit radix sort.
//
// One sort column at a time, last column first (every pass is stable, so earlier columns dominate): the column's values
// become order-preserving 64-bit keys (integers: sign bit flipped; floats: IEEE totalOrder key, like arrow's total_cmp;
// descending: complemented), the keys are sorted by ten 7-bit digit passes, and one more 1-bit pass puts the NULLs first
// or last.  A digit pass = per-tile digit histogram -> exclusive scan (partition-major) -> stable placement with
// warp-match ranking (k_part_place of exchange_kernels.cu, which here writes the carried row index).
#include "device_utils.cuh"
#include "exchange_kernels.cuh"
#include "sort_kernels.cuh"
#include "util_kernels.cuh"

namespace qgpu {
namespace {

constexpr int ST = 256;
constexpr int SR = 4;
constexpr int STILE = ST * SR;  // == the tile of k_part_hist / k_part_place

// order key of every row of one column (in ORIGINAL row order); NULL rows get key 0 and a 0 in `notnull`
__global__ void k_sort_keys(const __grid_constant__ ColRef col, int64_t n, int descending, uint64_t* keys, uint8_t* notnull) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool ok = row_valid(col, i);
    uint64_t k = 0;
    if (ok) {
      k = (uint64_t)cmp_key(load_value(col, i), (uint8_t)col.dtype) ^ 0x8000000000000000ull;
      if (descending) k = ~k;
    }
    keys[i] = k;
    notnull[i] = ok ? 1 : 0;
  }
}

// digit of every position of the current order + per-tile histogram (partition-major: hist[d * n_tiles + tile])
__global__ void __launch_bounds__(ST) k_digit_hist(const uint64_t* keys, const uint8_t* flags, const uint32_t* order, int64_t n, int shift,
                                                   int n_digits, int flag_pass, int flag_invert, uint8_t* digit, uint32_t* hist,
                                                   int64_t n_tiles) {
  __shared__ uint32_t s_hist[MAX_PARTS];
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (tid < n_digits) s_hist[tid] = 0;
    __syncthreads();
    const int64_t base = tile * STILE;
#pragma unroll
    for (int j = 0; j < SR; ++j) {
      const int64_t pos = base + j * ST + tid;
      if (pos < n) {
        const uint32_t row = order[pos];
        uint32_t d;
        if (flag_pass) d = (uint32_t)(flags[row] ^ (uint8_t)flag_invert) & 1u;
        else d = (uint32_t)(keys[row] >> shift) & (uint32_t)(n_digits - 1);
        digit[pos] = (uint8_t)d;
        atomicAdd(&s_hist[d], 1u);
      }
    }
    __syncthreads();
    if (tid < n_digits) hist[(int64_t)tid * n_tiles + tile] = s_hist[tid];
    __syncthreads();
  }
}

}  // namespace

void sort_permutation(Ctx* c, const std::vector<ColRef>& cols, const std::vector<int>& descending, const std::vector<int>& nulls_first,
                      int64_t n, uint32_t* perm) {
  launch_iota_u32(c, perm, n);
  if (n <= 1 || cols.empty()) return;
  constexpr int DIGIT_BITS = 7, N_DIGITS = 1 << DIGIT_BITS;
  static_assert(N_DIGITS <= MAX_PARTS, "digit passes reuse the partition kernels");
  const int64_t n_tiles = part_tiles(n);
  DevBuf keys(c, pad_bytes((size_t)n * 8)), notnull(c, pad_bytes((size_t)n)), digit(c, pad_bytes((size_t)n));
  DevBuf other(c, pad_bytes((size_t)n * 4));
  DevBuf hist(c, (size_t)N_DIGITS * n_tiles * 4 + 256), hist_off(c, ((size_t)N_DIGITS * n_tiles + 1) * 8 + 256);
  uint32_t* cur = perm;
  uint32_t* nxt = other.as<uint32_t>();
  auto pass = [&](int shift, int n_digits, int flag_pass, int flag_invert) {
    {
      ProfScope ps(c, "k_digit_hist");
      k_digit_hist<<<persistent_grid(c, k_digit_hist, ST, n_tiles), ST, 0, c->stream>>>(keys.as<uint64_t>(), notnull.as<uint8_t>(), cur, n, shift,
                                                                                       n_digits, flag_pass, flag_invert, digit.as<uint8_t>(),
                                                                                       hist.as<uint32_t>(), n_tiles);
      QG_CUDA(cudaGetLastError());
    }
    scan_u32_to_i64(c, hist.as<uint32_t>(), (int64_t)n_digits * n_tiles, hist_off.as<int64_t>());
    launch_part_perm(c, n, n_digits, digit.as<uint8_t>(), hist_off.as<int64_t>(), nxt, cur);
    std::swap(cur, nxt);
  };
  for (int k = (int)cols.size() - 1; k >= 0; --k) {
    {
      ProfScope ps(c, "k_sort_keys");
      k_sort_keys<<<persistent_grid(c, k_sort_keys, 256, (n + 255) / 256), 256, 0, c->stream>>>(cols[(size_t)k], n, descending[(size_t)k],
                                                                                             keys.as<uint64_t>(), notnull.as<uint8_t>());
      QG_CUDA(cudaGetLastError());
    }
    for (int shift = 0; shift < 64; shift += DIGIT_BITS) pass(shift, N_DIGITS, 0, 0);
    if (cols[(size_t)k].valid != nullptr || cols[(size_t)k].dtype == DT_NULL)
      pass(0, 2, 1, nulls_first[(size_t)k] ? 0 : 1);  // digit = notnull (nulls first) or its complement (nulls last)
  }
  if (cur != perm) QG_CUDA(cudaMemcpyAsync(perm, cur, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));  // (the scratch buffers above go back to the pool in stream order anyway)
}

}  // namespace qgpu
