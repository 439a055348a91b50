// rscatter_kernels.cu — radix scatter of 8-byte columns, second generation (see rscatter_kernels.cuh), and the region
// histogram / offsets kernels that give a scatter exact (gap-free) placement.
//
// Used by: the hash join (ops_join.cu: build side over-allocated, probe side straight into the output columns at exact
// positions - crates/query-executor/src/executor.rs:343-399 is the operator it serves), the high-cardinality GROUP BY
// (ops_groupby.cu: two radix levels so that every partition's groups fit a shared-memory table) and their multi-GPU
// forms, where the scatter's stores go to the owning GPU over NVLink (the exchange of
// crates/query-distributed/src/planner.rs:200-249).
#include <mutex>
#include <set>

#include "pj_device.cuh"
#include "rscatter_kernels.cuh"

namespace qgpu {
namespace {

constexpr int RT = 256;             // threads
constexpr int RR = RS_TILE / RT;    // rows per thread
constexpr int RWARPS = RT / 32;

// shared memory: [STAGES][NC][RS_TILE] staged columns | mbarriers | order_src u16[RS_TILE] | order_region u8[RS_TILE] |
// gdelta u32[parts + 1] | owner u32[parts + 1] (MULTI) | scan_w u32[16] | warp histograms u32[RWARPS][parts + 1]
template <int NC, int STAGES>
constexpr size_t rs_fixed_smem() {
  return (size_t)STAGES * NC * RS_TILE * 8 + 64 + (size_t)RS_TILE * 2 + (size_t)RS_TILE + 64;
}
size_t rs_smem_bytes(int nc, int stages, uint32_t parts, bool multi) {
  return (size_t)stages * nc * RS_TILE * 8 + 64 + (size_t)RS_TILE * 2 + (size_t)RS_TILE + 64 +
         (size_t)(parts + 1) * 4 * (multi ? 2 : 1) + (size_t)RWARPS * (parts + 1) * 4 + 64;
}

template <int NC, int STAGES, bool MULTI, int KIND, bool RANKA>
__global__ void __launch_bounds__(RT, STAGES == 1 ? (NC <= 3 ? 4 : 3) : (NC <= 2 ? 3 : 2))
    k_rscatter(const __grid_constant__ RSArgs A, const __grid_constant__ RSPeers PE) {
  extern __shared__ __align__(128) unsigned char rs_smem[];
  uint64_t* stage_base = reinterpret_cast<uint64_t*>(rs_smem);
  unsigned char* sp = rs_smem + (size_t)STAGES * NC * RS_TILE * 8;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sp);
  sp += 64;
  uint16_t* order_src = reinterpret_cast<uint16_t*>(sp);
  sp += (size_t)RS_TILE * 2;
  uint8_t* order_region = sp;
  sp += RS_TILE;
  uint32_t* scan_w = reinterpret_cast<uint32_t*>(sp);
  sp += 64;
  const uint32_t parts = A.M.parts;
  const uint32_t hstride = parts + 1;
  uint32_t* gdelta = reinterpret_cast<uint32_t*>(sp);
  sp += (size_t)hstride * 4;
  uint32_t* sowner = gdelta;
  if (MULTI) {
    sowner = reinterpret_cast<uint32_t*>(sp);
    sp += (size_t)hstride * 4;
  }
  uint32_t* whist = reinterpret_cast<uint32_t*>(sp);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* wh = whist + warp * hstride;
  const int64_t n_tiles = (A.n + RS_TILE - 1) / RS_TILE;
  const uint64_t pol_stream = l2_policy_evict_first();

  // live rows of a tile (region-major input: rows past the region's count are gaps; a tile never straddles regions)
  auto tile_rows = [&](int64_t tile, uint32_t& in_region) -> uint32_t {
    const int64_t row0 = tile * RS_TILE;
    if (A.in_cap == 0) {
      in_region = 0;
      return (uint32_t)min((int64_t)RS_TILE, A.n - row0);
    }
    in_region = (uint32_t)(row0 / A.in_cap);
    const int64_t live_end = (int64_t)in_region * A.in_cap + min(A.in_count[(size_t)in_region * A.in_ctr_stride], A.in_cap);
    return row0 >= live_end ? 0u : (uint32_t)min((int64_t)RS_TILE, live_end - row0);
  };
  auto next_live = [&](int64_t tile) -> int64_t {
    uint32_t g;
    while (tile < n_tiles && tile_rows(tile, g) == 0) tile += gridDim.x;
    return tile;
  };
  auto stage_col = [&](uint32_t st, int ci) -> uint64_t* { return stage_base + ((size_t)st * NC + ci) * RS_TILE; };
  auto issue = [&](uint32_t st, int64_t tile) {  // one thread
    uint32_t g;
    const uint32_t rows = tile_rows(tile, g);
    const uint32_t bytes = ((rows * 8u) + 15u) & ~15u;  // (columns are padded: reading 8 bytes past the end is safe)
    mbar_expect_tx(&bar[st], (uint32_t)NC * bytes);
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) bulk_g2s(stage_col(st, ci), A.in[ci] + tile * RS_TILE, bytes, &bar[st], pol_stream);
  };

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int64_t cur = next_live(blockIdx.x);
  if (tid == 0 && cur < n_tiles) issue(0, cur);
  uint32_t phase_bits = 0, st = 0;
  while (cur < n_tiles) {
    const int64_t nxt = next_live(cur + gridDim.x);
    // (STAGES == 2: the other stage was released by the closing barrier of the previous tile)
    if (STAGES > 1 && tid == 0 && nxt < n_tiles) issue(st ^ 1u, nxt);
    uint32_t in_region;
    const uint32_t rows = tile_rows(cur, in_region);
    const uint32_t part_base = (A.in_cap && !A.merge_in) ? in_region * parts : 0u;  // first output region of this tile's input region
    mbar_wait(&bar[st], (phase_bits >> st) & 1u);
    phase_bits ^= 1u << st;
    const uint64_t* skey = stage_col(st, 0);

    // ---- regions; rank inside the warp: rows of one region keep (r, lane) order ----
    uint32_t p[RR], rank[RR];
#pragma unroll
    for (int r = 0; r < RR; ++r) {
      const uint32_t i = r * RT + tid;
      p[r] = i < rows ? pj_region_of_k<KIND>(A.M, skey[i]) : parts;  // bucket `parts` = past the end / dropped
    }
    for (uint32_t j = lane; j < hstride; j += 32) wh[j] = 0;
    if (RANKA) {
      // the warp-private histogram's atomicAdd hands every row its rank (lanes of one region are serialised by the
      // hardware; the order inside a region is arbitrary, which nobody needs): one instruction per row
      __syncwarp();
#pragma unroll
      for (int r = 0; r < RR; ++r) rank[r] = atomicAdd(&wh[p[r]], 1u);
    } else {
      // the eight matches are independent and issue back to back; only the histogram updates form a chain
      uint32_t peers[RR];
#pragma unroll
      for (int r = 0; r < RR; ++r) peers[r] = __match_any_sync(0xffffffffu, p[r]);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < RR; ++r) {
        // every peer reads the count so far (one broadcast), the leader then adds the group: no shuffle
        const uint32_t old = wh[p[r]];
        __syncwarp();
        if (lane == __ffs(peers[r]) - 1) wh[p[r]] = old + __popc(peers[r]);
        rank[r] = old + __popc(peers[r] & lanemask_lt());
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- per region (one thread each; parts <= RT): offsets of the warps inside the tile, total, reservation ----
    uint32_t total = 0, gbase = 0;
    const uint32_t q_out = part_base + (uint32_t)tid;  // output region of thread tid
    if ((uint32_t)tid < parts) {
#pragma unroll
      for (int w = 0; w < RWARPS; ++w) {
        const uint32_t cnt = whist[w * hstride + tid];
        whist[w * hstride + tid] = total;
        total += cnt;
      }
      if (total) {  // in flight during the scan
        if (MULTI) {
          const uint32_t owner = q_out / PE.regions_per_rank;
          gbase = atomicAdd(PE.cursor[owner] + (size_t)(q_out - owner * PE.regions_per_rank) * A.ctr_stride, total);
        } else {
          gbase = atomicAdd(A.cursor + (size_t)q_out * A.ctr_stride, total);
        }
      }
    }
    {  // block exclusive scan of `total` over the regions (threads >= parts contribute 0)
      uint32_t wt;
      const uint32_t ex = warp_excl_scan_u32(total, wt);
      if (lane == 0) scan_w[warp] = wt;
      __syncthreads();
      if (warp == 0) {
        const uint32_t x = lane < RWARPS ? scan_w[lane] : 0;
        uint32_t t;
        const uint32_t e = warp_excl_scan_u32(x, t);
        if (lane < RWARPS) scan_w[lane] = e;
        if (lane == 0) scan_w[RWARPS] = t;
      }
      __syncthreads();
      if ((uint32_t)tid < parts) {
        const uint32_t ts = scan_w[warp] + ex;  // start of the region's run in the tile's region-major order
        // fold the run start into the warp offsets: position of a row = whist[warp][region] + rank
#pragma unroll
        for (int w = 0; w < RWARPS; ++w) whist[w * hstride + tid] += ts;
        // first output row of the region (MULTI: inside the owner's copy); positions are mod 2^32, checked by the host
        const uint32_t ql = MULTI ? q_out % PE.regions_per_rank : q_out;
        uint32_t first = (A.start ? A.start[ql] : ql * A.cap) + A.first_base;
        if (A.start == nullptr && total && gbase + total > A.cap) {
          // an overflowing run goes to the trash rows: the flag is up and the caller takes another path
          atomicOr(A.flags, (uint32_t)PJ_OVERFLOW);
          first = A.trash_row;
          gbase = ts;
        }
        gdelta[tid] = first + gbase - ts;
        if (MULTI) sowner[tid] = q_out / PE.regions_per_rank;
      }
    }
    __syncthreads();
    const uint32_t tcount = scan_w[RWARPS];  // rows of the tile that go to a region
#pragma unroll
    for (int r = 0; r < RR; ++r)
      if (p[r] < parts) {
        const uint32_t pos = wh[p[r]] + rank[r];
        order_src[pos] = (uint16_t)(r * RT + tid);
        order_region[pos] = (uint8_t)p[r];
      }
    __syncthreads();
    // ---- write-out, position-major: consecutive threads write consecutive rows of a run ----
#pragma unroll
    for (int k = 0; k < RR; ++k) {
      const uint32_t i = k * RT + tid;
      if (i < tcount) {
        const uint32_t q = order_region[i], src = order_src[i];
        const uint32_t d = gdelta[q] + i;
        if (MULTI) {
          const uint32_t owner = sowner[q];
#pragma unroll
          for (int ci = 0; ci < NC; ++ci) PE.out[owner][ci][d] = stage_col(st, ci)[src];
        } else {
          if (A.plain_stores) {
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) A.out[ci][d] = stage_col(st, ci)[src];
          } else {
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) st_evict_first_u64(A.out[ci] + d, stage_col(st, ci)[src], pol_stream);
          }
        }
      }
    }
    __syncthreads();  // order / histograms / this stage are rewritten from here on
    if (STAGES == 1 && tid == 0 && nxt < n_tiles) issue(0, nxt);
    cur = nxt;
    if (STAGES > 1) st ^= 1u;
  }
}

// ---- region histogram: one shared-memory counter per (region, lane) - lane l only ever touches bank l, so the native
// 32-bit RED is conflict-free whatever the regions of a warp's rows are; one global atomic per (CTA, region) at the end --
constexpr int HT = 256;
constexpr int HR = 8;
template <int KIND>
__global__ void __launch_bounds__(HT) k_region_hist(const int64_t* __restrict__ key, int64_t n, const __grid_constant__ PJMap M,
                                                    uint32_t* counts, uint32_t ctr_stride) {
  extern __shared__ uint32_t hist[];  // [parts + 1][32]
  const int tid = threadIdx.x, lane = tid & 31;
  for (uint32_t j = tid; j < (M.parts + 1) * 32; j += HT) hist[j] = 0;
  __syncthreads();
  const int64_t tile_rows = (int64_t)HT * HR;
  const int64_t n_tiles = (n + tile_rows - 1) / tile_rows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * tile_rows + tid;
    uint64_t k[HR];
#pragma unroll
    for (int r = 0; r < HR; ++r) k[r] = base + (int64_t)r * HT < n ? ld_stream_u64((const uint64_t*)key + base + (int64_t)r * HT) : 0ull;
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      const uint32_t p = base + (int64_t)r * HT < n ? pj_region_of_k<KIND>(M, k[r]) : M.parts;
      atomicAdd(&hist[p * 32 + lane], 1u);
    }
  }
  __syncthreads();
  for (uint32_t j = tid; j < M.parts; j += HT) {
    uint32_t t = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) t += hist[j * 32 + ((l + j) & 31)];  // (rotated: the threads of a warp start on different banks)
    if (t) atomicAdd(counts + (size_t)j * ctr_stride, t);
  }
}

__global__ void __launch_bounds__(PJ_MAX_PARTS) k_region_offsets(const uint32_t* counts, uint32_t ctr_stride, uint32_t parts, uint32_t max_count,
                                                                 uint32_t* start, uint32_t* flags) {
  __shared__ uint32_t s_w[PJ_MAX_PARTS / 32 + 1];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t v = tid < parts ? counts[(size_t)tid * ctr_stride] : 0u;
  if (v > max_count) atomicOr(flags, (uint32_t)PJ_OVERFLOW);
  uint32_t wt;
  const uint32_t ex = warp_excl_scan_u32(v, wt);
  if (lane == 0) s_w[warp] = wt;
  __syncthreads();
  if (warp == 0) {
    const uint32_t x = lane < PJ_MAX_PARTS / 32 ? s_w[lane] : 0u;
    uint32_t t;
    const uint32_t e = warp_excl_scan_u32(x, t);
    if (lane < PJ_MAX_PARTS / 32) s_w[lane] = e;
    if (lane == 0) s_w[PJ_MAX_PARTS / 32] = t;
  }
  __syncthreads();
  if (tid < parts) start[tid] = s_w[warp] + ex;
  if (tid == 0) start[parts] = s_w[PJ_MAX_PARTS / 32];
}

// function attributes are per device: run `f` the first time device c->device reaches call site `site` (thread-safe)
template <class F>
void once_per_device(Ctx* c, int site, F f) {
  static std::mutex mu;
  static std::set<std::pair<int, int>> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.insert({c->device, site}).second) f();
}

template <int NC, int STAGES, bool MULTI, int KIND, bool RANKA>
void launch_rs_kr(Ctx* c, const RSArgs& a, const RSPeers& pe, const char* name) {
  auto kern = k_rscatter<NC, STAGES, MULTI, KIND, RANKA>;
  once_per_device(c, ((((NC * 8 + STAGES * 2 + (MULTI ? 1 : 0)) << 2) + KIND) << 1) + (RANKA ? 1 : 0), [&] {
    QG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)rs_smem_bytes(NC, STAGES, PJ_MAX_PARTS, MULTI)));
  });
  const size_t smem = rs_smem_bytes(NC, STAGES, a.M.parts, MULTI);
  const int64_t tiles = (a.n + RS_TILE - 1) / RS_TILE;
  ProfScope ps(c, name);
  kern<<<persistent_grid(c, kern, RT, tiles, smem), RT, smem, c->stream>>>(a, pe);
  QG_CUDA(cudaGetLastError());
}

template <int NC, int STAGES, bool MULTI, int KIND>
void launch_rs_k(Ctx* c, const RSArgs& a, const RSPeers& pe, const char* name) {
  if (c->opt("scatter.rank_atomic", 1) != 0) return launch_rs_kr<NC, STAGES, MULTI, KIND, true>(c, a, pe, name);
  return launch_rs_kr<NC, STAGES, MULTI, KIND, false>(c, a, pe, name);
}

template <int NC, int STAGES, bool MULTI>
void launch_rs(Ctx* c, const RSArgs& a, const RSPeers& pe, const char* name) {
  switch (a.M.dense) {
    case 0: return launch_rs_k<NC, STAGES, MULTI, 0>(c, a, pe, name);
    case 1: return launch_rs_k<NC, STAGES, MULTI, 1>(c, a, pe, name);
    case 2: return launch_rs_k<NC, STAGES, MULTI, 2>(c, a, pe, name);
    case 3: return launch_rs_k<NC, STAGES, MULTI, 3>(c, a, pe, name);
    default: throw_error(QGPU_ERR_INVALID, "radix scatter: unknown key map %d", a.M.dense);
  }
}

}  // namespace

void launch_rscatter(Ctx* c, const RSArgs& a_in, int n_cols, const RSPeers* peers, const char* name) {
  if (a_in.n <= 0) return;
  if (a_in.M.parts < 1 || a_in.M.parts > PJ_MAX_PARTS) throw_error(QGPU_ERR_INVALID, "radix scatter into %u regions", a_in.M.parts);
  RSPeers none;
  memset(&none, 0, sizeof none);
  const RSPeers& pe = peers ? *peers : none;
  RSArgs a = a_in;
  a.plain_stores = c->opt("scatter.plain_stores", 1) != 0 ? 1u : 0u;
  const bool multi = peers != nullptr;
  // "scatter.stages": 2 = the next tile is staged while this one is processed (fewer, fatter CTAs per SM), 1 = single
  // buffer, more CTAs per SM hide each other's staging; a fourth column is always staged single-buffered
  const bool two = c->opt("scatter.stages", 2) != 1 && n_cols < 4;
#define QG_RS(NC)                                                                             \
  if (n_cols == NC) {                                                                         \
    if (two) {                                                                                \
      if (multi) return launch_rs<NC, (NC < 4 ? 2 : 1), true>(c, a, pe, name);                \
      return launch_rs<NC, (NC < 4 ? 2 : 1), false>(c, a, pe, name);                          \
    }                                                                                         \
    if (multi) return launch_rs<NC, 1, true>(c, a, pe, name);                                 \
    return launch_rs<NC, 1, false>(c, a, pe, name);                                           \
  }
  QG_RS(1) QG_RS(2) QG_RS(3) QG_RS(4)
#undef QG_RS
  throw_error(QGPU_ERR_INVALID, "radix scatter of %d columns", n_cols);
}

void launch_region_hist(Ctx* c, const int64_t* key, int64_t n, const PJMap& map, uint32_t* counts, uint32_t ctr_stride) {
  if (n <= 0) return;
  const int64_t tiles = (n + HT * HR - 1) / (HT * HR);
  ProfScope ps(c, "k_region_hist");
  const size_t smem = (size_t)(map.parts + 1) * 32 * 4;  // <= 33 KB
  switch (map.dense) {
    case 0: k_region_hist<0><<<persistent_grid(c, k_region_hist<0>, HT, tiles, smem), HT, smem, c->stream>>>(key, n, map, counts, ctr_stride); break;
    case 1: k_region_hist<1><<<persistent_grid(c, k_region_hist<1>, HT, tiles, smem), HT, smem, c->stream>>>(key, n, map, counts, ctr_stride); break;
    case 2: k_region_hist<2><<<persistent_grid(c, k_region_hist<2>, HT, tiles, smem), HT, smem, c->stream>>>(key, n, map, counts, ctr_stride); break;
    default: k_region_hist<3><<<persistent_grid(c, k_region_hist<3>, HT, tiles, smem), HT, smem, c->stream>>>(key, n, map, counts, ctr_stride); break;
  }
  QG_CUDA(cudaGetLastError());
}

void launch_region_offsets(Ctx* c, const uint32_t* counts, uint32_t ctr_stride, uint32_t parts, uint32_t max_count, uint32_t* start,
                           uint32_t* flags) {
  ProfScope ps(c, "k_region_offsets");
  k_region_offsets<<<1, PJ_MAX_PARTS, 0, c->stream>>>(counts, ctr_stride, parts, max_count, start, flags);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
