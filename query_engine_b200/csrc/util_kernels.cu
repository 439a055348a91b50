// util_kernels.cu — prefix scans, order-preserving stream compaction (warp-ballot / popc), gathers.
#include "device_utils.cuh"
#include "util_kernels.cuh"

namespace qgpu {
namespace {

constexpr int SCAN_T = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_CHUNK = SCAN_T * SCAN_ITEMS;

// block-wide exclusive scan of one int64 per thread (1024 threads); returns exclusive prefix, total via smem
__device__ __forceinline__ int64_t block_excl_scan_i64(int64_t v, int64_t* s_warp /*[33]*/, int64_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int64_t w = s_warp[lane];
    int64_t xs = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t y = __shfl_up_sync(0xffffffffu, xs, d);
      if (lane >= d) xs += y;
    }
    s_warp[lane] = xs - w;
    if (lane == 31) s_warp[32] = xs;
  }
  __syncthreads();
  const int64_t res = s_warp[warp] + x - v;
  total = s_warp[32];
  __syncthreads();
  return res;
}

template <class TIn>
__global__ void __launch_bounds__(SCAN_T) k_scan_block_sums(const TIn* in, int64_t n, int64_t* sums) {
  __shared__ int64_t s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) v += (int64_t)in[base + i];
  int64_t total;
  block_excl_scan_i64(v, s_warp, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// in-place exclusive scan of nb block sums by one CTA; sums[nb] = grand total
__global__ void __launch_bounds__(SCAN_T) k_scan_sums(int64_t* sums, int64_t nb) {
  __shared__ int64_t s_warp[33];
  int64_t carry = 0;
  for (int64_t b0 = 0; b0 < nb; b0 += SCAN_T) {
    const int64_t i = b0 + threadIdx.x;
    const int64_t v = i < nb ? sums[i] : 0;
    int64_t total;
    const int64_t ex = block_excl_scan_i64(v, s_warp, total);
    if (i < nb) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) sums[nb] = carry;
}

template <class TIn>
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const TIn* in, int64_t n, const int64_t* sums, int64_t* out) {
  __shared__ int64_t s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t item[SCAN_ITEMS];
  int64_t v = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    item[i] = base + i < n ? (int64_t)in[base + i] : 0;
    v += item[i];
  }
  int64_t total;
  int64_t run = sums[blockIdx.x] + block_excl_scan_i64(v, s_warp, total);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += item[i];
    if (base + i == n - 1) out[n] = run;
  }
}

template <class TIn>
void scan_impl(Ctx* c, const TIn* in, int64_t n, int64_t* out) {
  if (n <= 0) {
    QG_CUDA(cudaMemsetAsync(out, 0, 8, c->stream));
    return;
  }
  const int64_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  DevBuf sums(c, (size_t)(nb + 1) * 8);
  {
    ProfScope ps(c, "k_scan_block_sums");
    k_scan_block_sums<TIn><<<(unsigned)nb, SCAN_T, 0, c->stream>>>(in, n, sums.as<int64_t>());
    QG_CUDA(cudaGetLastError());
  }
  {
    ProfScope ps(c, "k_scan_sums");
    k_scan_sums<<<1, SCAN_T, 0, c->stream>>>(sums.as<int64_t>(), nb);
    QG_CUDA(cudaGetLastError());
  }
  {
    ProfScope ps(c, "k_scan_apply");
    k_scan_apply<TIn><<<(unsigned)nb, SCAN_T, 0, c->stream>>>(in, n, sums.as<int64_t>(), out);
    QG_CUDA(cudaGetLastError());
  }
}

// ------------------------------- compaction -------------------------------------------------------
struct CompactParams {
  int64_t n_rows;
  const uint32_t* mask;
  const int64_t* tile_off;
  int32_t n_cols;
  int32_t tile_rows;
  CompactCol cols[MAX_COMPACT_COLS];
};

// One CTA per tile.  Lane l of the warp that owns bitmap word w handles row 32*w + l, so the output
// position is tile_off + (popc of earlier words in the tile) + popc(word & lanemask_lt): the classic
// ballot/popc stream compaction, order preserving.
template <int J>
__global__ void __launch_bounds__(256) k_compact(const __grid_constant__ CompactParams P) {
  __shared__ uint32_t s_word[J * 8];
  __shared__ uint32_t s_pref[J * 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile = blockIdx.x;
  const int64_t base = tile * P.tile_rows;
  constexpr int NWORDS = J * 8;
  if (tid < NWORDS) s_word[tid] = P.mask[(base >> 5) + tid];
  __syncthreads();
  if (warp == 0) {
    // exclusive prefix of popc over the tile's words (NWORDS is 32 or 64)
    uint32_t run = 0;
    for (int w0 = 0; w0 < NWORDS; w0 += 32) {
      const uint32_t cnt = __popc(s_word[w0 + lane]);
      uint32_t tot;
      const uint32_t ex = warp_excl_scan_u32(cnt, tot);
      s_pref[w0 + lane] = run + ex;
      run += tot;
    }
  }
  __syncthreads();
  const int64_t out0 = P.tile_off[tile];
  int64_t pos[J];
  bool sel[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int w = j * 8 + warp;
    const uint32_t word = s_word[w];
    sel[j] = (word >> lane) & 1u;
    pos[j] = out0 + s_pref[w] + __popc(word & lanemask_lt());
  }
  // Columns go through in groups of CG: all loads of a group are issued before its first store (the compiler cannot hoist a
  // column's loads above the previous column's stores - they might alias), so a tile costs one memory round trip per
  // group instead of one per column.
  constexpr int CG = 4;
  for (int c0 = 0; c0 < P.n_cols; c0 += CG) {
    uint64_t v[CG][J];
#pragma unroll
    for (int ci = 0; ci < CG; ++ci) {
      if (c0 + ci >= P.n_cols) continue;
      const CompactCol& c = P.cols[c0 + ci];
      if (c.kind == CK_W8) {
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (sel[j]) v[ci][j] = ld_stream_u64((const uint64_t*)c.in + base + j * 256 + tid);
      } else if (c.kind == CK_W4) {
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (sel[j]) v[ci][j] = ld_stream_u32((const uint32_t*)c.in + base + j * 256 + tid);
      }
    }
#pragma unroll
    for (int ci = 0; ci < CG; ++ci) {
      if (c0 + ci >= P.n_cols) continue;
      const CompactCol& c = P.cols[c0 + ci];
      if (c.kind == CK_W8) {
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (sel[j]) ((uint64_t*)c.out)[pos[j]] = v[ci][j];
      } else if (c.kind == CK_W4) {
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (sel[j]) ((uint32_t*)c.out)[pos[j]] = (uint32_t)v[ci][j];
      } else if (c.kind == CK_BITS) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
          if (sel[j]) {
            const int64_t row = base + j * 256 + tid;
            const uint32_t bit = (((const uint32_t*)c.in)[row >> 5] >> (row & 31)) & 1u;
            if (bit) atomicOr((uint32_t*)c.out + (pos[j] >> 5), 1u << (pos[j] & 31));
          }
        }
      } else {  // CK_ROWIDX
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (sel[j]) ((uint32_t*)c.out)[pos[j]] = (uint32_t)(base + j * 256 + tid);
      }
      if (c.out_valid != nullptr) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
          if (sel[j]) {
            const int64_t row = base + j * 256 + tid;
            const uint32_t bit = c.in_valid ? ((c.in_valid[row >> 5] >> (row & 31)) & 1u) : 1u;
            if (bit) atomicOr(c.out_valid + (pos[j] >> 5), 1u << (pos[j] & 31));
          }
        }
      }
    }
  }
}

__global__ void k_prefix_at(const int64_t* bounds, int nb, int tile_rows, const uint32_t* mask, const int64_t* tile_off,
                            int64_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const int64_t rb = bounds[i];
  const int64_t tile = rb / tile_rows;
  int64_t pos = tile_off[tile];
  const int64_t w0 = (tile * tile_rows) >> 5, w1 = rb >> 5;
  for (int64_t w = w0; w < w1; ++w) pos += __popc(mask[w]);
  const int rem = (int)(rb & 31);
  if (rem) pos += __popc(mask[w1] & ((1u << rem) - 1u));
  out[i] = pos;
}

// ------------------------------- gathers -----------------------------------------------------------
struct GatherParams {
  const uint32_t* idx;
  int64_t n_out;
  int32_t n_cols;
  int32_t pad;
  GatherCol cols[MAX_GATHER_COLS];
};

__global__ void __launch_bounds__(256) k_gather(const __grid_constant__ GatherParams P) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = ((P.n_out + 31) / 32) * 32;  // keep whole warps together for the ballots
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += nthreads) {
    const bool inb = i < P.n_out;
    const uint32_t ix = inb ? (P.idx ? P.idx[i] : (uint32_t)i) : NULL_IDX;
    const bool has = ix != NULL_IDX;
    for (int cidx = 0; cidx < P.n_cols; ++cidx) {
      const GatherCol c = P.cols[cidx];
      if (c.kind == CK_W8) {
        if (inb) ((uint64_t*)c.out)[i] = has ? ((const uint64_t*)c.in)[ix] : 0ull;
      } else if (c.kind == CK_W4) {
        if (inb) ((uint32_t*)c.out)[i] = has ? ((const uint32_t*)c.in)[ix] : 0u;
      } else if (c.kind == CK_BITS) {
        const bool bit = has && ((((const uint32_t*)c.in)[ix >> 5] >> (ix & 31)) & 1u);
        const uint32_t w = __ballot_sync(0xffffffffu, bit);
        if ((threadIdx.x & 31) == 0 && inb) ((uint32_t*)c.out)[i >> 5] = w;
      }
      if (c.out_valid != nullptr) {
        const bool v = has && (c.in_valid == nullptr || ((c.in_valid[ix >> 5] >> (ix & 31)) & 1u));
        const uint32_t w = __ballot_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0 && inb) c.out_valid[i >> 5] = w;
      }
    }
  }
}

__global__ void k_gather_utf8_len(const int32_t* off, const uint32_t* idx, int64_t n, int64_t* len) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t ix = idx ? idx[i] : (uint32_t)i;
    len[i] = ix == NULL_IDX ? 0 : (int64_t)(off[ix + 1] - off[ix]);
  }
}

__global__ void k_i64_to_i32(const int64_t* in, int64_t n, int32_t* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int32_t)in[i];
}

__global__ void k_gather_utf8_copy(const int32_t* in_off, const uint8_t* in_chars, const uint32_t* idx, int64_t n,
                                   const int32_t* out_off, uint8_t* out_chars) {
  // one warp per output string: lanes stride over its bytes
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t i = warp_global; i < n; i += nwarps) {
    const uint32_t ix = idx ? idx[i] : (uint32_t)i;
    if (ix == NULL_IDX) continue;
    const int32_t sb = in_off[ix], len = in_off[ix + 1] - sb, db = out_off[i];
    for (int32_t k = lane; k < len; k += 32) out_chars[db + k] = in_chars[sb + k];
  }
}

struct FoldParams {
  const uint64_t* all;
  uint64_t* finals;
  int world, n_acc;
  uint8_t kinds[MAX_ACC];
};
__global__ void k_fold_finals(const __grid_constant__ FoldParams F) {
  const int s = threadIdx.x;
  if (s > F.n_acc) return;
  if (s == F.n_acc) {
    uint64_t e = 0;
    for (int r = 0; r < F.world; ++r) e |= F.all[(size_t)r * (F.n_acc + 1) + s];
    F.finals[s] = e;
    return;
  }
  uint64_t v = acc_identity(F.kinds[s]);
  for (int r = 0; r < F.world; ++r) v = acc_combine(F.kinds[s], v, F.all[(size_t)r * (F.n_acc + 1) + s]);
  F.finals[s] = v;
}

__global__ void k_gather_strided_i64(const int64_t* in, int64_t stride, int count, int64_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = in[(int64_t)i * stride];
}

__global__ void k_iota_u32(uint32_t* out, int64_t n, uint32_t first) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = first + (uint32_t)i;
}

__global__ void k_bits_to_bytes(const uint32_t* bits, int64_t n, uint8_t* bytes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bytes[i] = bits ? (uint8_t)((bits[i >> 5] >> (i & 31)) & 1u) : (uint8_t)1;
}

__global__ void k_bytes_to_bits(const uint8_t* bytes, int64_t n, uint32_t* bits) {
  const int64_t n_round = ((n + 31) / 32) * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = __ballot_sync(0xffffffffu, i < n && bytes[i] != 0);
    if ((threadIdx.x & 31) == 0 && i < n) bits[i >> 5] = w;
  }
}

// Sparse selections (Q6's predicate keeps 1.8 %): one WARP per 1024-row tile, no CTA barrier, work proportional to the
// survivors.  Lane l holds bitmap word l; survivor number s of the tile (row order) is found by a 5-step search over the
// lanes' exclusive popcounts plus __fns inside the word, so lane (s mod 32) moves survivor s: all its column loads are in
// flight together, the stores of a warp go to consecutive output rows.  k_compact (one CTA per tile, every row
// inspected by a thread) spends ~3.7 us of launch + barrier + round-trip latency per tile however few rows survive:
// 0.44 ms against 0.1x ms here for 5 columns of lineitem SF10.
__global__ void __launch_bounds__(256) k_compact_sparse(const __grid_constant__ CompactParams P, int64_t n_tiles) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t tile = warp; tile < n_tiles; tile += n_warps) {
    const uint32_t word = P.mask[tile * 32 + lane];
    uint32_t total;
    const uint32_t ex = warp_excl_scan_u32((uint32_t)__popc(word), total);
    if (total == 0) continue;
    const int64_t out0 = P.tile_off[tile];
    const int64_t base = tile * 1024;
    for (uint32_t s0 = 0; s0 < total; s0 += 32) {
      const uint32_t s = s0 + lane;
      const bool active = s < total;
      const uint32_t sq = active ? s : total - 1;
      uint32_t w = 0;  // the last word whose exclusive count is <= sq (words without survivors share their successor's count)
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const uint32_t cand = w + step;
        const uint32_t e = __shfl_sync(0xffffffffu, ex, cand & 31);
        if (cand < 32 && e <= sq) w = cand;
      }
      const uint32_t wword = __shfl_sync(0xffffffffu, word, w);
      const uint32_t wex = __shfl_sync(0xffffffffu, ex, w);
      const uint32_t bit = __fns(wword, 0, (int)(sq - wex) + 1);
      const int64_t row = base + w * 32 + bit;
      const int64_t pos = out0 + s;
      constexpr int CG = 8;
      for (int c0 = 0; c0 < P.n_cols; c0 += CG) {
        uint64_t v[CG];
#pragma unroll
        for (int ci = 0; ci < CG; ++ci) {
          if (c0 + ci >= P.n_cols || !active) continue;
          const CompactCol& c = P.cols[c0 + ci];
          if (c.kind == CK_W8) v[ci] = ld_stream_u64((const uint64_t*)c.in + row);
          else if (c.kind == CK_W4) v[ci] = ld_stream_u32((const uint32_t*)c.in + row);
        }
#pragma unroll
        for (int ci = 0; ci < CG; ++ci) {
          if (c0 + ci >= P.n_cols || !active) continue;
          const CompactCol& c = P.cols[c0 + ci];
          if (c.kind == CK_W8) {
            ((uint64_t*)c.out)[pos] = v[ci];
          } else if (c.kind == CK_W4) {
            ((uint32_t*)c.out)[pos] = (uint32_t)v[ci];
          } else if (c.kind == CK_BITS) {
            const uint32_t b = (((const uint32_t*)c.in)[row >> 5] >> (row & 31)) & 1u;
            if (b) atomicOr((uint32_t*)c.out + (pos >> 5), 1u << (pos & 31));
          } else {  // CK_ROWIDX
            ((uint32_t*)c.out)[pos] = (uint32_t)row;
          }
          if (c.out_valid != nullptr) {
            const uint32_t b = c.in_valid ? ((c.in_valid[row >> 5] >> (row & 31)) & 1u) : 1u;
            if (b) atomicOr(c.out_valid + (pos >> 5), 1u << (pos & 31));
          }
        }
      }
    }
  }
}

// selected rows per 1024-row tile of a selection bitmap (one warp per tile, one 32-bit word per lane)
__global__ void k_mask_tile_counts(const uint32_t* mask, int64_t n_tiles, uint32_t* counts) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_tiles; t += n_warps) {
    const uint32_t cnt = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mask[t * 32 + (threadIdx.x & 31)]));
    if ((threadIdx.x & 31) == 0) counts[t] = cnt;
  }
}

template <class K>
unsigned grid_for(Ctx* c, K kernel, int64_t n, int threads) {
  return persistent_grid(c, kernel, threads, (n + threads - 1) / threads);
}

}  // namespace

void scan_u32_to_i64(Ctx* c, const uint32_t* in, int64_t n, int64_t* out) { scan_impl<uint32_t>(c, in, n, out); }
void scan_i64_to_i64(Ctx* c, const int64_t* in, int64_t n, int64_t* out) { scan_impl<int64_t>(c, in, n, out); }

void launch_compact(Ctx* c, int64_t n_rows, int tile_rows, const uint32_t* mask, const int64_t* tile_off,
                    const CompactCol* cols, int n_cols, int64_t n_selected) {
  if (n_rows <= 0) return;
  const int64_t n_tiles = (n_rows + tile_rows - 1) / tile_rows;
  // at most one row in "compact.sparse_den" (default 8) survives: the warp-per-tile kernel
  const int64_t den = c->opt("compact.sparse_den", 8);
  const bool sparse = tile_rows == 1024 && n_selected >= 0 && den > 0 && n_selected * den <= n_rows;
  for (int c0 = 0; c0 < n_cols; c0 += MAX_COMPACT_COLS) {
    CompactParams P;
    P.n_rows = n_rows;
    P.mask = mask;
    P.tile_off = tile_off;
    P.tile_rows = tile_rows;
    P.n_cols = std::min(MAX_COMPACT_COLS, n_cols - c0);
    for (int i = 0; i < P.n_cols; ++i) P.cols[i] = cols[c0 + i];
    if (sparse) {
      ProfScope ps(c, "k_compact_sparse");
      k_compact_sparse<<<persistent_grid(c, k_compact_sparse, 256, (n_tiles + 7) / 8), 256, 0, c->stream>>>(P, n_tiles);
      QG_CUDA(cudaGetLastError());
      continue;
    }
    ProfScope ps(c, "k_compact");
    if (tile_rows == 1024) k_compact<4><<<(unsigned)n_tiles, 256, 0, c->stream>>>(P);
    else if (tile_rows == 2048) k_compact<8><<<(unsigned)n_tiles, 256, 0, c->stream>>>(P);
    else throw_error(QGPU_ERR_INVALID, "bad compaction tile %d", tile_rows);
    QG_CUDA(cudaGetLastError());
  }
}

void launch_prefix_at(Ctx* c, const int64_t* d_boundaries, int nb, int tile_rows, const uint32_t* mask,
                      const int64_t* tile_off, int64_t* d_out) {
  ProfScope ps(c, "k_prefix_at");
  k_prefix_at<<<(nb + 127) / 128, 128, 0, c->stream>>>(d_boundaries, nb, tile_rows, mask, tile_off, d_out);
  QG_CUDA(cudaGetLastError());
}

void launch_gather(Ctx* c, const uint32_t* idx, int64_t n_out, const GatherCol* cols, int n_cols) {
  if (n_out <= 0) return;
  for (int c0 = 0; c0 < n_cols; c0 += MAX_GATHER_COLS) {
    GatherParams P;
    P.idx = idx;
    P.n_out = n_out;
    P.n_cols = std::min(MAX_GATHER_COLS, n_cols - c0);
    for (int i = 0; i < P.n_cols; ++i) P.cols[i] = cols[c0 + i];
    ProfScope ps(c, "k_gather");
    k_gather<<<grid_for(c, k_gather, n_out, 256), 256, 0, c->stream>>>(P);
    QG_CUDA(cudaGetLastError());
  }
}

void launch_gather_utf8_len(Ctx* c, const int32_t* in_off, const uint32_t* idx, int64_t n_out, int64_t* out_len) {
  if (n_out <= 0) return;
  ProfScope ps(c, "k_gather_utf8_len");
  k_gather_utf8_len<<<grid_for(c, k_gather_utf8_len, n_out, 256), 256, 0, c->stream>>>(in_off, idx, n_out, out_len);
  QG_CUDA(cudaGetLastError());
}

void launch_offsets_i64_to_i32(Ctx* c, const int64_t* in, int64_t n, int32_t* out) {
  if (n <= 0) return;
  ProfScope ps(c, "k_i64_to_i32");
  k_i64_to_i32<<<grid_for(c, k_i64_to_i32, n, 256), 256, 0, c->stream>>>(in, n, out);
  QG_CUDA(cudaGetLastError());
}

void launch_gather_utf8_copy(Ctx* c, const int32_t* in_off, const uint8_t* in_chars, const uint32_t* idx, int64_t n_out,
                             const int32_t* out_off, uint8_t* out_chars) {
  if (n_out <= 0) return;
  ProfScope ps(c, "k_gather_utf8_copy");
  k_gather_utf8_copy<<<grid_for(c, k_gather_utf8_copy, n_out * 32, 256), 256, 0, c->stream>>>(in_off, in_chars, idx, n_out, out_off, out_chars);
  QG_CUDA(cudaGetLastError());
}

void launch_bits_to_bytes(Ctx* c, const uint32_t* bits, int64_t n, uint8_t* bytes) {
  if (n <= 0) return;
  ProfScope ps(c, "k_bits_to_bytes");
  k_bits_to_bytes<<<grid_for(c, k_bits_to_bytes, n, 256), 256, 0, c->stream>>>(bits, n, bytes);
  QG_CUDA(cudaGetLastError());
}

void launch_bytes_to_bits(Ctx* c, const uint8_t* bytes, int64_t n, uint32_t* bits) {
  if (n <= 0) return;
  ProfScope ps(c, "k_bytes_to_bits");
  k_bytes_to_bits<<<grid_for(c, k_bytes_to_bits, n, 256), 256, 0, c->stream>>>(bytes, n, bits);
  QG_CUDA(cudaGetLastError());
}

void launch_mask_tile_counts(Ctx* c, const uint32_t* mask, int64_t n_tiles, uint32_t* counts) {
  if (n_tiles <= 0) return;
  ProfScope ps(c, "k_mask_tile_counts");
  k_mask_tile_counts<<<grid_for(c, k_mask_tile_counts, n_tiles * 32, 256), 256, 0, c->stream>>>(mask, n_tiles, counts);
  QG_CUDA(cudaGetLastError());
}

void launch_fold_finals(Ctx* c, const uint64_t* all, int world, int n_acc, const uint8_t* kinds_host, uint64_t* finals) {
  FoldParams F;
  F.all = all;
  F.finals = finals;
  F.world = world;
  F.n_acc = n_acc;
  for (int i = 0; i < n_acc && i < MAX_ACC; ++i) F.kinds[i] = kinds_host[i];
  ProfScope ps(c, "k_fold_finals");
  k_fold_finals<<<1, 64, 0, c->stream>>>(F);
  QG_CUDA(cudaGetLastError());
}

void launch_gather_strided_i64(Ctx* c, const int64_t* in, int64_t stride, int count, int64_t* out) {
  if (count <= 0) return;
  ProfScope ps(c, "k_gather_strided_i64");
  k_gather_strided_i64<<<(count + 127) / 128, 128, 0, c->stream>>>(in, stride, count, out);
  QG_CUDA(cudaGetLastError());
}

void launch_iota_u32(Ctx* c, uint32_t* out, int64_t n, uint32_t first) {
  if (n <= 0) return;
  ProfScope ps(c, "k_iota_u32");
  k_iota_u32<<<grid_for(c, k_iota_u32, n, 256), 256, 0, c->stream>>>(out, n, first);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
