// pjoin_kernels.cu — radix-partitioned hash join kernels (see pjoin_kernels.cuh for the design).
//
// Replaces the body of QueryExecutor::execute_join (crates/query-executor/src/executor.rs:343-399) for the common
// shape `probe.key = build.key` on Int64 keys with a duplicate-free build side (a primary-key / foreign-key join):
// NULL-free keys, INNER keeps matched pairs, LEFT also keeps unmatched probe rows with NULL right columns
// (docs/SQL_REFERENCE.md:215-229); output schema = left fields ++ right fields (executor.rs:532-537).  Output row
// order is the order tiles reserve output space in (unspecified by the reference; parity is on row sets).
#include "device_utils.cuh"
#include "pjoin_kernels.cuh"

namespace qgpu {
namespace {

constexpr unsigned long long PJ_EMPTY = ~0ull;

// one 64-bit mix per key: the TOP bits choose the region, the LOW 32 bits the bucket inside the region table
QG_D uint32_t pj_region(uint64_t h, uint32_t log_parts) { return log_parts ? (uint32_t)(h >> (64 - log_parts)) : 0u; }
// region of a key; M.parts = "no region" (the row is dropped)
QG_D uint32_t pj_region_of(const PJMap& M, uint64_t key) {
  if (!M.dense) return pj_region(mix64(key), M.log_parts);
  const uint64_t idx = key - (uint64_t)M.kmin;
  if (idx < M.range) return (uint32_t)(idx >> M.shift);
  return M.drop_outside ? M.parts : M.parts - 1;
}

// ------------------------------------------------------------------------------------------------------------
// scatter: rows -> regions
// ------------------------------------------------------------------------------------------------------------
constexpr int ST = 512;            // threads
constexpr int SR = 8;              // rows per thread
constexpr int STILE = ST * SR;     // 4096 rows per tile
constexpr int SWARPS = ST / 32;

struct ScatterSmem {
  uint64_t stage[STILE];    // one column of the tile, region-major
  uint32_t dest[STILE];     // global row position of every staged row (0xffffffff: dropped, region overflow)
  uint32_t tstart[PJ_MAX_PARTS + 2];
  uint32_t gbase[PJ_MAX_PARTS + 1];  // first free row of the region reserved for this tile
  uint32_t scan_w[SWARPS + 1];
  // followed by warp histograms: SWARPS x (parts + 1) uint32
};

QG_D uint64_t load_col(const PJCol& c, int64_t row) {
  return c.width == 8 ? ld_stream_u64((const uint64_t*)c.in + row) : (uint64_t)ld_stream_u32((const uint32_t*)c.in + row);
}

__global__ void __launch_bounds__(ST, 2) k_pj_scatter(const int64_t* __restrict__ key, int64_t n, const __grid_constant__ PJMap M,
                                                       uint32_t cap, uint32_t* cursor, uint32_t* flags,
                                                       const __grid_constant__ PJCols C) {
  const uint32_t parts = M.parts;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem& S = *reinterpret_cast<ScatterSmem*>(smem_raw);
  uint32_t* whist = reinterpret_cast<uint32_t*>(smem_raw + sizeof(ScatterSmem));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t hstride = parts + 1;
  uint32_t* wh = whist + warp * hstride;
  const int64_t n_tiles = (n + STILE - 1) / STILE;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * STILE + tid;
    // ---- keys, regions ----
    uint64_t k[SR];
    uint32_t p[SR];
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      const int64_t row = row0 + (int64_t)r * ST;
      const bool live = row < n;
      k[r] = live ? ld_stream_u64((const uint64_t*)key + row) : 0ull;
      p[r] = live ? pj_region_of(M, k[r]) : parts;  // bucket `parts` = past the end / dropped
    }
    for (uint32_t j = lane; j < hstride; j += 32) wh[j] = 0;
    __syncwarp();
    // ---- rank inside the warp: rows of one region keep (r, lane) order ----
    uint32_t rank[SR];
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      const uint32_t peers = __match_any_sync(0xffffffffu, p[r]);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = wh[p[r]];
        wh[p[r]] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[r] = old + __popc(peers & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();
    // ---- per region: offsets of the warps inside the tile, total, global reservation ----
    uint32_t total = 0;
    if ((uint32_t)tid < parts) {
#pragma unroll 4
      for (int w = 0; w < SWARPS; ++w) {
        const uint32_t c = whist[w * hstride + tid];
        whist[w * hstride + tid] = total;
        total += c;
      }
    }
    {  // block exclusive scan of `total` over the regions (threads >= parts contribute 0)
      uint32_t wt;
      const uint32_t ex = warp_excl_scan_u32(total, wt);
      if (lane == 0) S.scan_w[warp] = wt;
      __syncthreads();
      if (warp == 0) {
        const uint32_t x = lane < SWARPS ? S.scan_w[lane] : 0;
        uint32_t t;
        const uint32_t e = warp_excl_scan_u32(x, t);
        if (lane < SWARPS) S.scan_w[lane] = e;
        if (lane == 0) S.scan_w[SWARPS] = t;
      }
      __syncthreads();
      if ((uint32_t)tid < parts) {
        const uint32_t ts = S.scan_w[warp] + ex;
        S.tstart[tid] = ts;
        uint32_t gbase = 0;
        if (total) {
          gbase = atomicAdd(cursor + tid, total);
          if (gbase + total > cap) atomicOr(flags, (uint32_t)PJ_OVERFLOW);
        }
        S.gbase[tid] = gbase;
      }
    }
    __syncthreads();
    const uint32_t tcount = S.scan_w[SWARPS];  // live rows of the tile
    // ---- staged positions ----
    uint32_t pos[SR];
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      if (p[r] < parts) {
        const uint32_t in_run = wh[p[r]] + rank[r];       // index inside this tile's run of the region
        const uint32_t j = S.gbase[p[r]] + in_run;         // index inside the region (< 2^32: cursors count rows)
        pos[r] = S.tstart[p[r]] + in_run;
        S.dest[pos[r]] = j < cap ? p[r] * cap + j : 0xffffffffu;  // parts * cap < 2^32 (checked by the host)
      } else {
        pos[r] = 0xffffffffu;
      }
    }
    // ---- columns, one at a time through the staging buffer ----
    uint64_t v[SR];
#pragma unroll
    for (int r = 0; r < SR; ++r) v[r] = C.key_col == 0 ? k[r] : (pos[r] != 0xffffffffu ? load_col(C.c[0], row0 + (int64_t)r * ST) : 0ull);
    for (int ci = 0; ci < C.n; ++ci) {
#pragma unroll
      for (int r = 0; r < SR; ++r)
        if (pos[r] != 0xffffffffu) S.stage[pos[r]] = v[r];
      __syncthreads();
      if (ci + 1 < C.n) {  // next column's loads are in flight while this one is written out
        if (ci + 1 == C.key_col) {
#pragma unroll
          for (int r = 0; r < SR; ++r) v[r] = k[r];
        } else {
          const PJCol nc = C.c[ci + 1];
#pragma unroll
          for (int r = 0; r < SR; ++r) v[r] = pos[r] != 0xffffffffu ? load_col(nc, row0 + (int64_t)r * ST) : 0ull;
        }
      }
      const PJCol oc = C.c[ci];
      if (oc.width == 8) {
        for (uint32_t i = tid; i < tcount; i += ST) {
          const uint32_t d = S.dest[i];
          if (d != 0xffffffffu) ((uint64_t*)oc.out)[d] = S.stage[i];
        }
      } else {
        for (uint32_t i = tid; i < tcount; i += ST) {
          const uint32_t d = S.dest[i];
          if (d != 0xffffffffu) ((uint32_t*)oc.out)[d] = (uint32_t)S.stage[i];
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// region tables: 32-byte buckets of two slots {k0, k1, v0, v1}; ~0 = empty key
// ------------------------------------------------------------------------------------------------------------
// One bucket = one 32-byte sector = what a probe step costs at the L2 whatever it reads, so a step looks at two
// slots.  Inserts fill slot 0 before slot 1 and move to the next bucket only when both are taken, so a lookup ends
// at the first empty slot.  At <= 0.45 load almost every chain is one bucket long, which matters more than it
// seems: a warp repeats a dependent L2 round trip until its LONGEST chain is done (measured with one-slot linear
// probing: 4-5 round trips per warp instead of the average 1.4).
QG_D uint32_t pj_bucket(uint64_t h, uint32_t n_buckets) { return __umulhi((uint32_t)h, n_buckets); }

// A memset of the whole table followed by the inserts costs two extra trips to HBM (the table is several times the
// L2: every CAS would first pull its sector back in).  Instead ONE kernel clears and fills region after region:
// work items are handed out in order by a ticket and an insert item waits until all clear items of its region have
// finished.  Clear items run one region AHEAD of the insert items - [clear 0][clear 1][insert 0][clear 2][insert 1]
// ... - so an insert item practically never waits, and when it does, it waits for lower tickets that are already
// running and never wait themselves: no deadlock.  The regions in flight stay in L2: the CAS hit L2 and the table
// reaches HBM once, as full lines.
constexpr int BT = 256;
constexpr uint32_t B_CLEAR_W16 = 4096;      // 64 KB of table per clear item (16-byte words)
constexpr uint32_t B_INS_ROWS = 1024;       // build rows per insert item
constexpr int BR = B_INS_ROWS / BT;

QG_D uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool DENSE>
__global__ void __launch_bounds__(BT) k_pj_build(const int64_t* __restrict__ key, const uint64_t* __restrict__ val,
                                                  const __grid_constant__ PJMap M, uint32_t cap, const uint32_t* __restrict__ count,
                                                  void* table_raw, uint32_t n_entries, uint32_t* flags, uint32_t* ticket,
                                                  uint32_t* cleared) {
  __shared__ uint32_t s_item;
  const int tid = threadIdx.x;
  const uint32_t parts = M.parts;
  // a region's table in 16-byte words: n_entries buckets of 32 B, or n_entries values of 8 B
  const uint32_t region_w16 = DENSE ? (n_entries + 1) / 2 : n_entries * 2;
  const uint32_t n_buckets = n_entries;
  PJBucket* table = reinterpret_cast<PJBucket*>(table_raw);
  const uint32_t n_clear = (region_w16 + B_CLEAR_W16 - 1) / B_CLEAR_W16;
  const uint32_t n_ins = (cap + B_INS_ROWS - 1) / B_INS_ROWS;
  const uint32_t per_step = n_clear + n_ins;
  const uint32_t n_items = (parts + 1) * per_step;  // step s: clear region s, insert region s - 1
  for (;;) {
    if (tid == 0) s_item = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    __syncthreads();
    if (item >= n_items) break;
    const uint32_t step = item / per_step;
    const uint32_t j = item - step * per_step;
    if (j < n_clear) {
      if (step >= parts) continue;
      ulonglong2* w = reinterpret_cast<ulonglong2*>(table_raw) + (size_t)step * region_w16;
      const uint32_t w0 = j * B_CLEAR_W16, w1 = min(w0 + B_CLEAR_W16, region_w16);
      for (uint32_t i = w0 + tid; i < w1; i += BT) w[i] = make_ulonglong2(PJ_EMPTY, PJ_EMPTY);
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(cleared + step, 1u);
      }
      continue;
    }
    if (step == 0) continue;
    const uint32_t p = step - 1;
    PJBucket* region = table + (size_t)p * n_buckets;
    const uint32_t i0 = (j - n_clear) * B_INS_ROWS;
    const uint32_t cnt = min(count[p], cap);
    if (i0 >= cnt) continue;
    if (tid == 0)
      while (ld_acquire_u32(cleared + p) < n_clear) __nanosleep(32);
    __syncthreads();
    const uint64_t base = (uint64_t)p * cap;
    // BR independent rows per thread: all loads first, then the CAS of every pending row back to back (one L2
    // round trip per step for all of them), until no row is pending
    unsigned long long k[BR], v[BR];
    uint32_t b[BR], sub[BR];
    uint32_t pend = 0;
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const uint32_t i = i0 + r * BT + tid;
      const bool live = i < cnt;
      k[r] = live ? (unsigned long long)ld_stream_u64((const uint64_t*)key + base + i) : PJ_EMPTY;
      v[r] = !live ? 0ull : (val ? (unsigned long long)ld_stream_u64(val + base + i) : (unsigned long long)(base + i));
      if (live) pend |= 1u << r;
    }
    if (DENSE) {
      // direct-addressed: entry (key - kmin) of the table; the region's slice is the one this item waited for.
      // (even-sized regions: region p starts at 16-byte word p * region_w16 = value index 2 * p * region_w16)
      unsigned long long* vals = reinterpret_cast<unsigned long long*>(table_raw) + (size_t)p * region_w16 * 2;
      const uint64_t first = (uint64_t)p << M.shift;
      unsigned long long old[BR];
#pragma unroll
      for (int r = 0; r < BR; ++r) {
        if (!((pend >> r) & 1u)) continue;
        if (v[r] == PJ_EMPTY) atomicOr(flags, (uint32_t)PJ_SENTINEL);
        old[r] = atomicExch(vals + ((k[r] - (uint64_t)M.kmin) - first), v[r]);
      }
#pragma unroll
      for (int r = 0; r < BR; ++r)
        if (((pend >> r) & 1u) && old[r] != PJ_EMPTY) atomicOr(flags, (uint32_t)PJ_DUPLICATE);
      continue;
    }
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      if (((pend >> r) & 1u) && k[r] == PJ_EMPTY) {
        atomicOr(flags, (uint32_t)PJ_SENTINEL);
        pend &= ~(1u << r);
      }
      b[r] = pj_bucket(mix64(k[r]), n_buckets);
      sub[r] = 0;
    }
    for (uint32_t steps = 0; pend != 0 && steps < 2 * n_buckets; ++steps) {
      unsigned long long old[BR];
#pragma unroll
      for (int r = 0; r < BR; ++r)
        if ((pend >> r) & 1u) old[r] = atomicCAS(&region[b[r]].k[sub[r]], PJ_EMPTY, k[r]);
#pragma unroll
      for (int r = 0; r < BR; ++r) {
        if (!((pend >> r) & 1u)) continue;
        if (old[r] == PJ_EMPTY) {
          region[b[r]].v[sub[r]] = v[r];  // read by the probe only
          pend &= ~(1u << r);
        } else if (old[r] == k[r]) {
          atomicOr(flags, (uint32_t)PJ_DUPLICATE);
          pend &= ~(1u << r);
        } else if (sub[r] == 0) {
          sub[r] = 1;
        } else {
          sub[r] = 0;
          b[r] = b[r] + 1 == n_buckets ? 0 : b[r] + 1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// probe + emit
// ------------------------------------------------------------------------------------------------------------
constexpr int PT = 256;
constexpr int PR = PJ_TILE / PT;  // 4 rows per thread

QG_D void ld_bucket(const PJBucket* p, unsigned long long& k0, unsigned long long& k1, unsigned long long& v0, unsigned long long& v1) {
  asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "l"(p));
  asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "l"((const char*)p + 16));
}

// ---- bulk-async staging (TMA engine, 1-D): global -> shared, completion on an mbarrier -------------------------
QG_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
QG_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
QG_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
QG_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
QG_D void bulk_g2s(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// The probe is latency-bound, not bandwidth-bound, unless the streaming inputs are taken off its critical path
// (measured at SF10: 1.20 ms; 0.53 ms with the table lookup removed; barriers + cursor atomics cost 0.06 ms).  So
// the key tile and the first 8-byte payload tile of the NEXT tile are fetched by the bulk-copy engine into shared
// memory (two stages, mbarrier completion, no registers) while the CTA works on the current one: what is left on
// the critical path of a tile is the L2 round trip for the buckets.
constexpr int PSTAGES = 2;
struct ProbeSmem {
  uint64_t keys[PSTAGES][PJ_TILE];
  uint64_t pay[PSTAGES][PJ_TILE];
  uint64_t bar[PSTAGES];
  uint32_t cnt[PR * (PT / 32)];
  unsigned long long base;
};

template <bool DENSE>
__global__ void __launch_bounds__(PT, 4) k_pj_probe(const int64_t* __restrict__ key, const __grid_constant__ PJMap M, uint32_t cap,
                                                  const uint32_t* __restrict__ count, const void* __restrict__ table_raw,
                                                  uint32_t n_entries, unsigned long long* out_cursor,
                                                  const __grid_constant__ PJEmit E) {
  __shared__ __align__(128) ProbeSmem S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t parts = M.parts;
  const uint32_t n_buckets = n_entries;
  const PJBucket* table = reinterpret_cast<const PJBucket*>(table_raw);
  const uint32_t region_w16 = DENSE ? (n_entries + 1) / 2 : n_entries * 2;
  const uint32_t tiles_per_part = cap / PJ_TILE;
  const uint64_t n_tiles = (uint64_t)parts * tiles_per_part;
  const uint64_t* pay_col = E.prefetch_left >= 0 ? (const uint64_t*)E.left[E.prefetch_left].in : nullptr;

  // this CTA's next tile (tiles b, b + grid, ...) that holds rows, or n_tiles
  auto next_tile = [&](uint64_t t) {
    for (; t < n_tiles; t += gridDim.x) {
      const uint32_t p = (uint32_t)(t / tiles_per_part);
      if ((uint32_t)(t - (uint64_t)p * tiles_per_part) * PJ_TILE < min(count[p], cap)) break;
    }
    return t;
  };
  auto issue = [&](int stage, uint64_t t) {  // one thread
    const uint32_t p = (uint32_t)(t / tiles_per_part);
    const uint32_t j = (uint32_t)(t - (uint64_t)p * tiles_per_part);
    const uint64_t row = (uint64_t)p * cap + (uint64_t)j * PJ_TILE;
    mbar_expect_tx(&S.bar[stage], (pay_col ? 2u : 1u) * PJ_TILE * 8u);
    bulk_g2s(S.keys[stage], key + row, PJ_TILE * 8u, &S.bar[stage]);
    if (pay_col) bulk_g2s(S.pay[stage], pay_col + row, PJ_TILE * 8u, &S.bar[stage]);
  };

  if (tid == 0) {
    for (int i = 0; i < PSTAGES; ++i) mbar_init(&S.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint64_t tile = next_tile(blockIdx.x);
  if (tid == 0 && tile < n_tiles) issue(0, tile);
  uint32_t phase_bits = 0;
  for (int stage = 0; tile < n_tiles; stage ^= 1) {
    const uint64_t tile_next = next_tile(tile + gridDim.x);
    // stage^1 was released by the barriers of the previous iteration
    if (tid == 0 && tile_next < n_tiles) issue(stage ^ 1, tile_next);
    const uint32_t p = (uint32_t)(tile / tiles_per_part);
    const uint32_t i0 = (uint32_t)(tile - (uint64_t)p * tiles_per_part) * PJ_TILE;
    const uint32_t cnt = min(count[p], cap);
    const PJBucket* region = table + (size_t)p * n_buckets;  // (hashed layout)
    const uint64_t row0 = (uint64_t)p * cap + i0 + tid;
    mbar_wait(&S.bar[stage], (phase_bits >> stage) & 1u);
    phase_bits ^= 1u << stage;

    unsigned long long k[PR], val[PR];
    bool live[PR];
    uint32_t b[PR];
    uint32_t pend = 0, found_bits = 0;
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      live[r] = i0 + r * PT + tid < cnt;
      k[r] = live[r] ? S.keys[stage][r * PT + tid] : PJ_EMPTY;
      val[r] = 0;
    }
    if (DENSE) {
      // direct-addressed: one 8-byte load per row, absent = ~0.  Rows of this region have (key - kmin) >> shift == p,
      // except keys outside the build range that a LEFT join parks in the last region.
      const unsigned long long* vals = reinterpret_cast<const unsigned long long*>(table_raw) + (size_t)p * region_w16 * 2;
      const uint64_t first = (uint64_t)p << M.shift;
#pragma unroll
      for (int r = 0; r < PR; ++r) {
        const uint64_t idx = k[r] - (uint64_t)M.kmin;
        val[r] = (live[r] && idx < M.range) ? ld_stream_u64((const uint64_t*)vals + (idx - first)) : PJ_EMPTY;
      }
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (val[r] != PJ_EMPTY) found_bits |= 1u << r;
    } else {
#pragma unroll
      for (int r = 0; r < PR; ++r) {
        b[r] = pj_bucket(mix64(k[r]), n_buckets);
        if (live[r] && k[r] != PJ_EMPTY) pend |= 1u << r;
      }
    }
    // one bucket per pending row per step, all rows of the thread in flight together; ends when no row is pending
    for (uint32_t steps = 0; pend != 0 && steps < n_buckets; ++steps) {
      unsigned long long k0[PR], k1[PR], v0[PR], v1[PR];
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if ((pend >> r) & 1u) ld_bucket(region + b[r], k0[r], k1[r], v0[r], v1[r]);
#pragma unroll
      for (int r = 0; r < PR; ++r) {
        if (!((pend >> r) & 1u)) continue;
        if (k0[r] == k[r] || k1[r] == k[r]) {
          val[r] = k0[r] == k[r] ? v0[r] : v1[r];
          found_bits |= 1u << r;
          pend &= ~(1u << r);
        } else if (k0[r] == PJ_EMPTY || k1[r] == PJ_EMPTY) {
          pend &= ~(1u << r);
        } else {
          b[r] = b[r] + 1 == n_buckets ? 0 : b[r] + 1;
        }
      }
    }
    bool found[PR];
#pragma unroll
    for (int r = 0; r < PR; ++r) found[r] = (found_bits >> r) & 1u;
    uint64_t pre[PR];
    if (pay_col) {
#pragma unroll
      for (int r = 0; r < PR; ++r) pre[r] = S.pay[stage][r * PT + tid];
    }
    // ---- output positions: round-major so that a fully matching warp writes 32 consecutive rows per store ----
    uint32_t lrank[PR];
    bool emit[PR];
    uint64_t pos[PR];
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      emit[r] = live[r] && (found[r] || E.left_join);
      const uint32_t b = __ballot_sync(0xffffffffu, emit[r]);
      lrank[r] = __popc(b & lanemask_lt());
      if (lane == 0) S.cnt[r * (PT / 32) + warp] = __popc(b);
    }
    __syncthreads();  // also: every thread has taken its keys / payload out of this stage
    if (warp == 0) {
      uint32_t t;
      const uint32_t e = warp_excl_scan_u32(S.cnt[lane], t);
      S.cnt[lane] = e;
      if (lane == 0) S.base = t ? atomicAdd(out_cursor, (unsigned long long)t) : 0ull;
    }
    __syncthreads();
    const unsigned long long base = S.base;
#pragma unroll
    for (int r = 0; r < PR; ++r) pos[r] = base + S.cnt[r * (PT / 32) + warp] + lrank[r];
    __syncthreads();  // S.cnt / S.base are rewritten by the next tile
    tile = tile_next;

    // ---- left columns: straight copy of the probe row ----
    for (int ci = 0; ci < E.n_left; ++ci) {
      const PJCol c = E.left[ci];
      if (ci == E.left_key) {
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint64_t*)c.out)[pos[r]] = k[r];
      } else if (ci == E.prefetch_left) {
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint64_t*)c.out)[pos[r]] = pre[r];
      } else if (c.width == 8) {
        uint64_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = emit[r] ? ld_stream_u64((const uint64_t*)c.in + row0 + r * PT) : 0ull;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint64_t*)c.out)[pos[r]] = v[r];
      } else {
        uint32_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = emit[r] ? ld_stream_u32((const uint32_t*)c.in + row0 + r * PT) : 0u;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint32_t*)c.out)[pos[r]] = v[r];
      }
    }
    // ---- right columns: the key is the probe key, the inlined payload rides in the slot, others are gathered ----
    for (int ci = 0; ci < E.n_right; ++ci) {
      const PJCol c = E.right[ci];
      if (ci == E.right_key) {
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint64_t*)c.out)[pos[r]] = found[r] ? k[r] : 0ull;
      } else if (ci == E.inline_right) {
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint64_t*)c.out)[pos[r]] = found[r] ? val[r] : 0ull;
      } else if (c.width == 8) {
        uint64_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = (emit[r] && found[r]) ? __ldg((const unsigned long long*)c.in + val[r]) : 0ull;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint64_t*)c.out)[pos[r]] = v[r];
      } else {
        uint32_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = (emit[r] && found[r]) ? __ldg((const uint32_t*)c.in + val[r]) : 0u;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (emit[r]) ((uint32_t*)c.out)[pos[r]] = v[r];
      }
    }
    if (E.right_valid_bytes != nullptr) {
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (emit[r]) E.right_valid_bytes[pos[r]] = found[r] ? 1 : 0;
    }
  }
}

}  // namespace

void launch_pj_scatter(Ctx* c, const int64_t* key, int64_t n, const PJMap& map, uint32_t cap, uint32_t* cursor, uint32_t* flags,
                       const PJCols& cols, const char* name) {
  if (n <= 0) return;
  const size_t smem = sizeof(ScatterSmem) + (size_t)SWARPS * (map.parts + 1) * sizeof(uint32_t);
  static bool attr_set = false;
  if (!attr_set) {
    QG_CUDA(cudaFuncSetAttribute(k_pj_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ScatterSmem) + SWARPS * (PJ_MAX_PARTS + 1) * 4)));
    attr_set = true;
  }
  ProfScope ps(c, name);
  const int64_t tiles = (n + STILE - 1) / STILE;
  k_pj_scatter<<<persistent_grid(c, k_pj_scatter, ST, tiles, smem), ST, smem, c->stream>>>(key, n, map, cap, cursor, flags, cols);
  QG_CUDA(cudaGetLastError());
}

void launch_pj_build(Ctx* c, const int64_t* key_part, const uint64_t* val_part, const PJMap& map, uint32_t cap, const uint32_t* count,
                     void* table, uint32_t n_entries, uint32_t* flags, uint32_t* ticket, uint32_t* cleared) {
  ProfScope ps(c, map.dense ? "k_pj_build_dense" : "k_pj_build");
  const uint32_t region_w16 = map.dense ? (n_entries + 1) / 2 : n_entries * 2;
  const int64_t items = (int64_t)(map.parts + 1) * ((region_w16 + B_CLEAR_W16 - 1) / B_CLEAR_W16 + (cap + B_INS_ROWS - 1) / B_INS_ROWS);
  if (map.dense)
    k_pj_build<true><<<persistent_grid(c, k_pj_build<true>, BT, items), BT, 0, c->stream>>>(key_part, val_part, map, cap, count, table,
                                                                                         n_entries, flags, ticket, cleared);
  else
    k_pj_build<false><<<persistent_grid(c, k_pj_build<false>, BT, items), BT, 0, c->stream>>>(key_part, val_part, map, cap, count, table,
                                                                                           n_entries, flags, ticket, cleared);
  QG_CUDA(cudaGetLastError());
}

void launch_pj_probe(Ctx* c, const int64_t* key_part, const PJMap& map, uint32_t cap, const uint32_t* count, const void* table,
                     uint32_t n_entries, unsigned long long* out_cursor, const PJEmit& e) {
  ProfScope ps(c, map.dense ? "k_pj_probe_dense" : "k_pj_probe");
  const int64_t tiles = (int64_t)map.parts * (cap / PJ_TILE);
  if (map.dense)
    k_pj_probe<true><<<persistent_grid(c, k_pj_probe<true>, PT, tiles), PT, 0, c->stream>>>(key_part, map, cap, count, table, n_entries,
                                                                                         out_cursor, e);
  else
    k_pj_probe<false><<<persistent_grid(c, k_pj_probe<false>, PT, tiles), PT, 0, c->stream>>>(key_part, map, cap, count, table, n_entries,
                                                                                           out_cursor, e);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
