// exchange_kernels.cu — hash partitioning of rows for Exchange / the NCCL shuffle.
//
// Row -> partition is the reference's: Rust `DefaultHasher` (SipHash-1-3, keys 0,0) over the non-null
// Int64 / Utf8 / Int32 key values of the row, `as usize % num_partitions`
// (crates/query-distributed/src/partition.rs:183-186, 292-316).  Rows keep their input order inside a
// partition (the reference pushes row indices in order and `take`s them, :189-200).
//
// Pass 1 hashes every row once, stores its partition id (1 byte) and builds per-tile histograms;
// a scan over the partition-major histogram gives every (partition, tile) its output offset; pass 2
// ranks rows inside the tile with __match_any_sync and writes the permutation consumed by the gathers.
#include "device_utils.cuh"
#include "exchange_kernels.cuh"

namespace qgpu {
namespace {

struct Sip {
  uint64_t v0, v1, v2, v3, tail;
  uint32_t ntail, total;
};
QG_D uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
QG_D void sip_round(Sip& s) {
  s.v0 += s.v1; s.v1 = rotl64(s.v1, 13); s.v1 ^= s.v0; s.v0 = rotl64(s.v0, 32);
  s.v2 += s.v3; s.v3 = rotl64(s.v3, 16); s.v3 ^= s.v2;
  s.v0 += s.v3; s.v3 = rotl64(s.v3, 21); s.v3 ^= s.v0;
  s.v2 += s.v1; s.v1 = rotl64(s.v1, 17); s.v1 ^= s.v2; s.v2 = rotl64(s.v2, 32);
}
QG_D void sip_init(Sip& s) {
  s.v0 = 0x736f6d6570736575ull; s.v1 = 0x646f72616e646f6dull; s.v2 = 0x6c7967656e657261ull; s.v3 = 0x7465646279746573ull;
  s.tail = 0; s.ntail = 0; s.total = 0;
}
QG_D void sip_block(Sip& s, uint64_t m) {
  s.v3 ^= m;
  sip_round(s);  // c = 1
  s.v0 ^= m;
}
// append the low `nb` bytes (1..8) of w, little endian
QG_D void sip_write(Sip& s, uint64_t w, uint32_t nb) {
  if (nb < 8) w &= (1ull << (8 * nb)) - 1ull;
  s.total += nb;
  s.tail |= s.ntail < 8 ? (w << (8 * s.ntail)) : 0ull;
  if (s.ntail + nb >= 8) {
    sip_block(s, s.tail);
    const uint32_t used = 8 - s.ntail;
    const uint32_t rest = nb - used;
    s.tail = (rest > 0 && used < 8) ? (w >> (8 * used)) : 0ull;
    s.ntail = rest;
  } else {
    s.ntail += nb;
  }
}
QG_D uint64_t sip_finish(Sip& s) {
  const uint64_t b = s.tail | ((uint64_t)(s.total & 0xff) << 56);
  sip_block(s, b);
  s.v2 ^= 0xff;
  sip_round(s); sip_round(s); sip_round(s);  // d = 3
  return s.v0 ^ s.v1 ^ s.v2 ^ s.v3;
}

QG_D uint64_t row_hash_sip(const PartKeys& K, int64_t row) {
  Sip s;
  sip_init(s);
  for (int k = 0; k < K.n_keys; ++k) {
    const ColRef& c = K.keys[k];
    if (!row_valid(c, row)) continue;  // nulls contribute nothing (:300-311)
    if (c.dtype == DT_I64) sip_write(s, ((const uint64_t*)c.data)[row], 8);
    else if (c.dtype == DT_I32) sip_write(s, (uint64_t)((const uint32_t*)c.data)[row], 4);
    else if (c.dtype == DT_UTF8) {
      const int32_t* off = (const int32_t*)c.data;
      for (int32_t i = off[row]; i < off[row + 1]; ++i) sip_write(s, c.chars[i], 1);
      sip_write(s, 0xff, 1);  // `impl Hash for str`
    }
    // other column types are not hashed by the reference
  }
  return sip_finish(s);
}

QG_D uint64_t row_hash_mix(const PartKeys& K, int64_t row) {
  // one key: exactly part_hash1(value) (the partitioned join kernels recompute it from the key alone)
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (int k = 0; k < K.n_keys; ++k) {
    const ColRef& c = K.keys[k];
    const uint64_t v = row_valid(c, row) ? load_value(c, row) : PART_NULL_MARK;
    h = mix64(h ^ v) + (uint64_t)k;
  }
  return h;
}

constexpr int PT = 256;
constexpr int PR = 4;
constexpr int PTILE = PT * PR;

// pass 1: partition id per row + per-tile histogram, partition-major: hist[p * n_tiles + tile]
__global__ void __launch_bounds__(PT) k_part_hist(const __grid_constant__ PartKeys K, int64_t n, int num_parts, int use_sip,
                                                   uint8_t* part_id, uint32_t* hist, int64_t n_tiles) {
  __shared__ uint32_t s_hist[MAX_PARTS];
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (tid < num_parts) s_hist[tid] = 0;
    __syncthreads();
    const int64_t base = tile * PTILE;
#pragma unroll
    for (int j = 0; j < PR; ++j) {
      const int64_t row = base + j * PT + tid;
      if (row < n) {
        uint32_t p;
        if (use_sip) {
          p = (uint32_t)(row_hash_sip(K, row) % (uint64_t)num_parts);  // `hash as usize % n`, partition.rs:186
        } else {
          // internal placement: low bits for a power-of-two count (what the partitioned join kernels recompute),
          // else a 32-bit modulo of the high word (a 64-bit `%` costs ~100 instructions per row)
          const uint64_t h = row_hash_mix(K, row);
          p = (num_parts & (num_parts - 1)) == 0 ? ((uint32_t)h & (uint32_t)(num_parts - 1)) : ((uint32_t)(h >> 32) % (uint32_t)num_parts);
        }
        part_id[row] = (uint8_t)p;
        atomicAdd(&s_hist[p], 1u);
      }
    }
    __syncthreads();
    if (tid < num_parts) hist[(int64_t)tid * n_tiles + tile] = s_hist[tid];
    __syncthreads();
  }
}

// pass 2: stable placement: position = offset(p, tile) + rank of the row among the tile's p-rows; the row index is
// written there (a permutation consumed by the column gathers).
__global__ void __launch_bounds__(PT) k_part_place(int64_t n, int num_parts, const uint8_t* part_id, const int64_t* hist_off,
                                                    int64_t n_tiles, uint32_t* perm, const uint32_t* carry) {
  __shared__ uint32_t s_cnt[PR * 8 * MAX_PARTS];  // [j][warp][p]
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int i = tid; i < PR * 8 * num_parts; i += PT) s_cnt[i] = 0;
    __syncthreads();
    const int64_t base = tile * PTILE;
    uint32_t p[PR], rank[PR];
#pragma unroll
    for (int j = 0; j < PR; ++j) {
      const int64_t row = base + j * PT + tid;
      const bool inb = row < n;
      p[j] = inb ? (uint32_t)part_id[row] : 0xffffu;  // out-of-range rows group together and are ignored
      const uint32_t peers = __match_any_sync(0xffffffffu, p[j]);
      rank[j] = __popc(peers & lanemask_lt());
      if (inb && rank[j] == 0) s_cnt[(j * 8 + warp) * num_parts + p[j]] = __popc(peers);
    }
    __syncthreads();
    // exclusive prefix over (j, warp) for every partition
    if (tid < num_parts) {
      uint32_t run = 0;
      for (int g = 0; g < PR * 8; ++g) {
        const uint32_t v = s_cnt[g * num_parts + tid];
        s_cnt[g * num_parts + tid] = run;
        run += v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PR; ++j) {
      const int64_t row = base + j * PT + tid;
      if (row < n) {
        const int64_t pos = hist_off[(int64_t)p[j] * n_tiles + tile] + s_cnt[(j * 8 + warp) * num_parts + p[j]] + rank[j];
        perm[pos] = carry ? carry[row] : (uint32_t)row;  // (the radix sort carries the original row index along)
      }
    }
    __syncthreads();
  }
}

}  // namespace

int64_t part_tiles(int64_t n) { return (n + PTILE - 1) / PTILE; }

void launch_part_hist(Ctx* c, const PartKeys& K, int64_t n, int num_parts, bool use_sip, uint8_t* part_id, uint32_t* hist) {
  const int64_t n_tiles = part_tiles(n);
  if (n_tiles == 0) return;
  const unsigned grid = persistent_grid(c, k_part_hist, PT, n_tiles);
  ProfScope ps(c, "k_part_hist");
  k_part_hist<<<grid, PT, 0, c->stream>>>(K, n, num_parts, use_sip ? 1 : 0, part_id, hist, n_tiles);
  QG_CUDA(cudaGetLastError());
}

void launch_part_perm(Ctx* c, int64_t n, int num_parts, const uint8_t* part_id, const int64_t* hist_off, uint32_t* perm,
                      const uint32_t* carry) {
  const int64_t n_tiles = part_tiles(n);
  if (n_tiles == 0) return;
  const unsigned grid = persistent_grid(c, k_part_place, PT, n_tiles);
  ProfScope ps(c, "k_part_perm");
  k_part_place<<<grid, PT, 0, c->stream>>>(n, num_parts, part_id, hist_off, n_tiles, perm, carry);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
