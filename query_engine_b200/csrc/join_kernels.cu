// join_kernels.cu — hash join build / probe (INNER, LEFT) on one equality key.
//
// Replaces the body of QueryExecutor::execute_join (crates/query-executor/src/executor.rs:343-399), whose
// shipped form is a Cartesian product that ignores `on`; the semantics implemented are the documented
// equi-join (docs/SQL_REFERENCE.md:215-229): NULL keys never match, LEFT keeps unmatched
// probe rows with NULL right columns, output rows come in probe (left) row order.
//
// Table: open addressing, linear probing, 16-byte slots {key, build row + 1}, capacity a power of two
// >= 2 x build rows, in HBM.  Duplicate build keys occupy separate slots of the same probe chain.
#include "device_utils.cuh"
#include "join_kernels.cuh"

namespace qgpu {
namespace {

QG_D uint32_t slot_of(int64_t key, uint32_t mask) { return (uint32_t)(mix64((uint64_t)key) >> 20) & mask; }

__global__ void __launch_bounds__(256) k_join_build(ColRef key, int64_t n, JoinSlot* tab, uint32_t mask) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
    if (!row_valid(key, row)) continue;  // NULL keys never match
    const int64_t k = (int64_t)load_value(key, row);
    uint32_t s = slot_of(k, mask);
    for (uint32_t probes = 0; probes <= mask; ++probes) {
      const uint32_t prev = atomicCAS(&tab[s].row1, 0u, (uint32_t)row + 1u);
      if (prev == 0) {
        tab[s].key = k;
        break;
      }
      s = (s + 1) & mask;
    }
  }
}

QG_D void load_slot(const JoinSlot* p, int64_t& key, uint32_t& row1) {
  uint64_t a, b;
  asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
  key = (int64_t)a;
  row1 = (uint32_t)b;
}

constexpr int JT = 256;
constexpr int JR = 4;
constexpr int JTILE = JT * JR;

// block-wide exclusive scan of one uint32 per thread, 256 threads
QG_D uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_w /*[9]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t wt;
  const uint32_t ex = warp_excl_scan_u32(v, wt);
  if (lane == 0) s_w[warp] = wt;
  __syncthreads();
  if (warp == 0) {
    const uint32_t x = lane < 8 ? s_w[lane] : 0;
    uint32_t t;
    const uint32_t e = warp_excl_scan_u32(x, t);
    if (lane < 8) s_w[lane] = e;
    if (lane == 0) s_w[8] = t;
  }
  __syncthreads();
  const uint32_t res = s_w[warp] + ex;
  total = s_w[8];
  __syncthreads();
  return res;
}

// pass 1: per probe row, number of output rows and the first matching build row
__global__ void __launch_bounds__(JT) k_join_count(ColRef key, int64_t n, const JoinSlot* tab, uint32_t mask, int left_join,
                                                    uint32_t* first_match, uint32_t* tile_counts, uint32_t* dup_flag) {
  __shared__ uint32_t s_w[9];
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile * JTILE < n; tile += gridDim.x) {
    const int64_t base = tile * JTILE;
    uint32_t mine = 0;
    bool dup = false;
#pragma unroll
    for (int j = 0; j < JR; ++j) {
      const int64_t row = base + j * JT + tid;
      if (row >= n) continue;
      uint32_t cnt = 0, first = 0;
      if (row_valid(key, row)) {
        const int64_t k = (int64_t)load_value(key, row);
        uint32_t s = slot_of(k, mask);
        for (uint32_t probes = 0; probes <= mask; ++probes) {
          int64_t sk;
          uint32_t r1;
          load_slot(tab + s, sk, r1);
          if (r1 == 0) break;
          if (sk == k) {
            if (cnt == 0) first = r1;
            ++cnt;
          }
          s = (s + 1) & mask;
        }
      }
      first_match[row] = first;
      if (cnt > 1) dup = true;
      mine += (left_join && cnt == 0) ? 1u : cnt;
    }
    if (dup) atomicOr(dup_flag, 1u);
    uint32_t total;
    block_excl_scan_256(mine, s_w, total);
    if (tid == 0) tile_counts[tile] = total;
  }
}

// pass 2: (left row, right row) index pairs in probe-row order
__global__ void __launch_bounds__(JT) k_join_emit(ColRef key, int64_t n, const JoinSlot* tab, uint32_t mask, int left_join,
                                                   int has_dups, const uint32_t* first_match, const int64_t* tile_off,
                                                   uint32_t* out_left, uint32_t* out_right) {
  __shared__ uint32_t s_w[9];
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile * JTILE < n; tile += gridDim.x) {
    const int64_t base = tile * JTILE;
    int64_t run = tile_off[tile];
#pragma unroll 1
    for (int j = 0; j < JR; ++j) {
      const int64_t row = base + j * JT + tid;
      const bool inb = row < n;
      uint32_t cnt = 0;
      uint32_t first = inb ? first_match[row] : 0;
      if (inb) {
        if (!has_dups) {
          cnt = first ? 1u : 0u;
        } else if (first) {
          const int64_t k = (int64_t)load_value(key, row);
          uint32_t s = slot_of(k, mask);
          for (uint32_t probes = 0; probes <= mask; ++probes) {
            int64_t sk;
            uint32_t r1;
            load_slot(tab + s, sk, r1);
            if (r1 == 0) break;
            if (sk == k) ++cnt;
            s = (s + 1) & mask;
          }
        }
      }
      const uint32_t outc = (inb && left_join && cnt == 0) ? 1u : cnt;
      uint32_t total;
      const uint32_t ex = block_excl_scan_256(outc, s_w, total);
      int64_t pos = run + ex;
      if (inb) {
        if (cnt == 0) {
          if (left_join) {
            out_left[pos] = (uint32_t)row;
            out_right[pos] = 0xffffffffu;
          }
        } else if (!has_dups) {
          out_left[pos] = (uint32_t)row;
          out_right[pos] = first - 1;
        } else {
          const int64_t k = (int64_t)load_value(key, row);
          uint32_t s = slot_of(k, mask);
          for (uint32_t probes = 0; probes <= mask; ++probes) {
            int64_t sk;
            uint32_t r1;
            load_slot(tab + s, sk, r1);
            if (r1 == 0) break;
            if (sk == k) {
              out_left[pos] = (uint32_t)row;
              out_right[pos] = r1 - 1;
              ++pos;
            }
            s = (s + 1) & mask;
          }
        }
      }
      run += total;
    }
  }
}

}  // namespace

void launch_join_build(Ctx* c, const ColRef& key, int64_t n, JoinSlot* tab, uint32_t cap) {
  QG_CUDA(cudaMemsetAsync(tab, 0, (size_t)cap * sizeof(JoinSlot), c->stream));
  if (n <= 0) return;
  int64_t grid = (n + 255) / 256;
  const int64_t cap_g = (int64_t)c->sm_count * 16;
  if (grid > cap_g) grid = cap_g;
  ProfScope ps(c, "k_join_build");
  k_join_build<<<(unsigned)grid, 256, 0, c->stream>>>(key, n, tab, cap - 1);
  QG_CUDA(cudaGetLastError());
}

int64_t join_tiles(int64_t n) { return (n + JTILE - 1) / JTILE; }

void launch_join_count(Ctx* c, const ColRef& key, int64_t n, const JoinSlot* tab, uint32_t cap, bool left_join,
                       uint32_t* first_match, uint32_t* tile_counts, uint32_t* dup_flag) {
  if (n <= 0) return;
  int64_t grid = join_tiles(n);
  const int64_t cap_g = (int64_t)c->sm_count * 8;
  if (grid > cap_g) grid = cap_g;
  ProfScope ps(c, "k_join_count");
  k_join_count<<<(unsigned)grid, JT, 0, c->stream>>>(key, n, tab, cap - 1, left_join ? 1 : 0, first_match, tile_counts, dup_flag);
  QG_CUDA(cudaGetLastError());
}

void launch_join_emit(Ctx* c, const ColRef& key, int64_t n, const JoinSlot* tab, uint32_t cap, bool left_join, bool has_dups,
                      const uint32_t* first_match, const int64_t* tile_off, uint32_t* out_left, uint32_t* out_right) {
  if (n <= 0) return;
  int64_t grid = join_tiles(n);
  const int64_t cap_g = (int64_t)c->sm_count * 8;
  if (grid > cap_g) grid = cap_g;
  ProfScope ps(c, "k_join_emit");
  k_join_emit<<<(unsigned)grid, JT, 0, c->stream>>>(key, n, tab, cap - 1, left_join ? 1 : 0, has_dups ? 1 : 0, first_match,
                                                    tile_off, out_left, out_right);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
