// pjoin_kernels.cu — radix-partitioned hash join kernels (see pjoin_kernels.cuh for the design).
//
// Replaces the body of QueryExecutor::execute_join (crates/query-executor/src/executor.rs:343-399) for the common
// shape `probe.key = build.key` on Int64 keys with a duplicate-free build side (a primary-key / foreign-key join):
// NULL-free keys, INNER keeps matched pairs, LEFT also keeps unmatched probe rows with NULL right columns
// (docs/SQL_REFERENCE.md:215-229); output schema = left fields ++ right fields (executor.rs:532-537).  Output row
// order is the order tiles reserve output space in (unspecified by the reference; parity is on row sets).
#include "pj_device.cuh"
#include "pjoin_kernels.cuh"

namespace qgpu {
namespace {

// ------------------------------------------------------------------------------------------------------------
// scatter: rows -> regions
// ------------------------------------------------------------------------------------------------------------
// Per tile of 2048 rows: (1) the key tile and the first 8-byte payload tile were fetched into shared memory by the
// bulk-copy engine while the previous tile was processed (two stages, mbarrier completion); (2) every warp ranks its
// rows inside their regions with MATCH.ANY and a warp-private histogram (no shared-memory atomics: measured 2 cycles
// per lane, slower than the global REDs they would replace); (3) a prefix over the warps gives every row its place in
// the tile's region-major order, and that order is recorded as (region, source row) - the values themselves are
// not moved inside shared memory; (4) one atomicAdd per (tile, region) reserves the run in the region; (5) every
// column is written out with consecutive threads on consecutive rows of a run, reading its value through the
// recorded source row: from the staged tile for the prefetched columns, from global memory (the tile's 16 KB
// window) for the others.
constexpr int ST = 256;             // threads
constexpr int SR = 8;               // rows per thread
constexpr int STILE = ST * SR;      // 2048 rows per tile
constexpr int SWARPS = ST / 32;
constexpr int SPRE = 2;             // columns staged by the bulk-copy engine: the key and the first 8-byte payload
constexpr int SSTAGES = 2;

struct ScatterSmem {
  uint64_t in[SSTAGES][SPRE][STILE];  // staged input tiles
  uint64_t bar[SSTAGES];
  uint16_t order_src[STILE];          // region-major order of the tile: source row ...
  uint8_t order_region[STILE];        // ... and region of every position
  uint32_t gdelta[PJ_MAX_PARTS + 1];  // region-major position i of region q goes to global row gdelta[q] + i ...
  uint32_t glimit[PJ_MAX_PARTS + 1];  // ... if that is below glimit[q] (the end of the region)
  uint32_t scan_w[SWARPS + 1];
  // followed by warp histograms: SWARPS x (parts + 1) uint32
};

// MULTI: the regions are spread over the GPUs of the box (region q belongs to GPU q / regions_per_rank): its rows,
// and the atomicAdd that reserves their run, go straight to the owner's exchange arena over NVLink - the radix
// scatter IS the exchange (no separate partition + materialise + send/recv).
template <bool MULTI>
__global__ void __launch_bounds__(ST, 3) k_pj_scatter(const int64_t* __restrict__ key, int64_t n, const __grid_constant__ PJMap M,
                                                       uint32_t cap, uint32_t* cursor, uint32_t* flags,
                                                       const __grid_constant__ PJCols C, const __grid_constant__ PJPeers PE) {
  const uint32_t parts = M.parts;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ScatterSmem& S = *reinterpret_cast<ScatterSmem*>(smem_raw);
  uint32_t* whist = reinterpret_cast<uint32_t*>(smem_raw + sizeof(ScatterSmem));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t hstride = parts + 1;
  uint32_t* wh = whist + warp * hstride;
  const int64_t n_tiles = (n + STILE - 1) / STILE;
  const uint64_t pol_stream = l2_policy_evict_first();
  // the staged payload column: the first 8-byte column that is not the key
  int pre_col = -1;
  for (int ci = 0; ci < C.n && pre_col < 0; ++ci)
    if (ci != C.key_col && C.c[ci].width == 8) pre_col = ci;

  auto issue = [&](uint32_t stage, int64_t tile) {  // one thread
    const int64_t row = tile * STILE;
    const uint32_t rows = (uint32_t)min((int64_t)STILE, n - row);
    const uint32_t bytes = ((rows * 8u) + 15u) & ~15u;  // (columns are padded: reading 8 bytes past the end is safe)
    mbar_expect_tx(&S.bar[stage], (pre_col >= 0 ? 2u : 1u) * bytes);
    bulk_g2s(S.in[stage][0], key + row, bytes, &S.bar[stage], pol_stream);
    if (pre_col >= 0) bulk_g2s(S.in[stage][1], (const uint64_t*)C.c[pre_col].in + row, bytes, &S.bar[stage], pol_stream);
  };

  if (tid == 0) {
    for (int i = 0; i < SSTAGES; ++i) mbar_init(&S.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && (int64_t)blockIdx.x < n_tiles) issue(0, blockIdx.x);
  uint32_t phase_bits = 0, stage = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, stage ^= 1u) {
    // the other stage was released by the closing barrier of the previous tile
    if (tid == 0 && tile + gridDim.x < n_tiles) issue(stage ^ 1u, tile + gridDim.x);
    const int64_t row0 = tile * STILE;
    const uint32_t tile_rows = (uint32_t)min((int64_t)STILE, n - row0);
    mbar_wait(&S.bar[stage], (phase_bits >> stage) & 1u);
    phase_bits ^= 1u << stage;
    const uint64_t* skey = S.in[stage][0];

    // ---- regions; rank inside the warp: rows of one region keep (r, lane) order ----
    uint32_t p[SR], rank[SR];
#pragma unroll
    for (int r = 0; r < SR; ++r) {
      const uint32_t i = r * ST + tid;
      p[r] = i < tile_rows ? pj_region_of(M, skey[i]) : parts;  // bucket `parts` = past the end / dropped
    }
    for (uint32_t j = lane; j < hstride; j += 32) wh[j] = 0;
    __syncwarp();
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
    for (uint32_t q = tid; q < parts; q += ST) {  // (parts <= ST)
#pragma unroll
      for (int w = 0; w < SWARPS; ++w) {
        const uint32_t c = whist[w * hstride + q];
        whist[w * hstride + q] = total;
        total += c;
      }
    }
    uint32_t gbase = 0;
    if ((uint32_t)tid < parts && total) {  // in flight during the scan
      if (MULTI) {
        const uint32_t owner = (uint32_t)tid / PE.regions_per_rank;
        gbase = atomicAdd(PE.cursor[owner] + ((uint32_t)tid - owner * PE.regions_per_rank) * PJ_CTR_STRIDE, total);
      } else {
        gbase = atomicAdd(cursor + tid * PJ_CTR_STRIDE, total);
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
        const uint32_t ts = S.scan_w[warp] + ex;  // start of the region's run in the tile's region-major order
        // fold the run start into the warp offsets: position of a row = whist[warp][region] + rank
#pragma unroll
        for (int w = 0; w < SWARPS; ++w) whist[w * hstride + tid] += ts;
        if (total && gbase + total > cap) atomicOr(flags, (uint32_t)PJ_OVERFLOW);
        // (mod 2^32; parts * cap < 2^32 is checked by the host.)  MULTI: positions are relative to the owner's copy
        const uint32_t ql = MULTI ? (uint32_t)tid % PE.regions_per_rank : (uint32_t)tid;
        S.gdelta[tid] = ql * cap + gbase - ts;
        S.glimit[tid] = (ql + 1) * cap;
      }
    }
    __syncthreads();
    const uint32_t tcount = S.scan_w[SWARPS];  // rows of the tile that go to a region
#pragma unroll
    for (int r = 0; r < SR; ++r)
      if (p[r] < parts) {
        const uint32_t pos = wh[p[r]] + rank[r];
        S.order_src[pos] = (uint16_t)(r * ST + tid);
        S.order_region[pos] = (uint8_t)p[r];
      }
    __syncthreads();
    // ---- columns: position-major, so that a row's destination is worked out once for all its columns ----
    for (uint32_t i = tid; i < tcount; i += ST) {
      const uint32_t q = S.order_region[i], src = S.order_src[i];
      const uint32_t d = S.gdelta[q] + i;
      // an overflowing region drops its rows (the flag is up and the join is redone on another path); a wrapped d
      // is wrong but still inside the buffers
      const uint32_t owner = MULTI ? q / PE.regions_per_rank : 0u;
      if (d >= S.glimit[q] || d < (MULTI ? q - owner * PE.regions_per_rank : q) * cap) continue;
      for (int ci = 0; ci < C.n; ++ci) {
        PJCol oc = C.c[ci];
        if (MULTI) oc.out = PE.out[owner][ci];
        if (ci == C.key_col) st_evict_first_u64((uint64_t*)oc.out + d, skey[src], pol_stream);
        else if (ci == pre_col) st_evict_first_u64((uint64_t*)oc.out + d, S.in[stage][1][src], pol_stream);
        else if (oc.width == 8) st_evict_first_u64((uint64_t*)oc.out + d, ld_stream_u64((const uint64_t*)oc.in + row0 + src), pol_stream);
        else ((uint32_t*)oc.out)[d] = ld_stream_u32((const uint32_t*)oc.in + row0 + src);
      }
    }
    __syncthreads();  // order / histograms / this stage are rewritten from here on
  }
}

// ------------------------------------------------------------------------------------------------------------
// region tables + probe, one kernel
// ------------------------------------------------------------------------------------------------------------
// hashed layout: 32-byte buckets of two slots {k0, k1, v0, v1}; ~0 = empty key.  One bucket = one 32-byte sector =
// what a probe step costs at the L2 whatever it reads, so a step looks at two slots.  Inserts fill slot 0 before
// slot 1 and move to the next bucket only when both are taken, so a lookup ends at the first empty slot.
// dense layout: one 8-byte value per key of the region's key range; ~0 = absent.
//
// Building the whole table and then probing it costs three extra trips to HBM (clear, fill - every CAS first pulls
// its sector back in - and, worst, the probe's first touch of every table sector is a RANDOM HBM read: measured
// 50 G/s against 283 G/s from L2).  So ONE kernel walks the regions as a software pipeline of work items
//     step s:  [clear table of region s]  [insert build rows of region s-1]  [probe tiles of region s-2]
// Items are handed out in that order.  An insert item waits until every clear item of its region has
// finished, a probe tile until every insert item of its region has; both are a whole step (> grid items) behind,
// so the waits practically never spin.  A CTA only ever waits for lower-numbered items, and with all CTAs
// co-resident (cooperative launch) the lowest unfinished item is always running: no deadlock.  The three regions
// in flight stay in L2: clears, CAS / exchanges and lookups all hit L2 and a table line reaches HBM once, when it
// is evicted after its region is done.
constexpr int JT = 256;                 // threads of the join kernel
constexpr uint32_t J_CLEAR_W16 = 4096;  // 64 KB of table per clear item (16-byte words)
constexpr uint32_t J_INS_ROWS = 1024;   // build rows per insert item
constexpr int BR = J_INS_ROWS / JT;
constexpr int PR = PJ_TILE / JT;        // probe rows per thread

// The key tile and the first 8-byte payload tile of a CTA's NEXT probe item are fetched by the bulk-copy engine
// into shared memory (two stages, mbarrier completion, no registers) while the CTA works on the current item: the
// streaming reads are off the critical path of a tile, what is left on it is the L2 round trip of the lookup.
constexpr int PSTAGES = 2;
struct JoinSmem {
  uint64_t keys[PSTAGES][PJ_TILE];
  uint64_t pay[PSTAGES][PJ_TILE];
  uint64_t bar[PSTAGES];
  uint32_t cnt[PR * (JT / 32)];
  uint32_t ticket;
  unsigned long long base;
};

enum ItemKind : uint32_t { IT_CLEAR = 0, IT_INSERT = 1, IT_PROBE = 2, IT_NONE = 3 };

struct ItemPlan {
  uint32_t n_clear, n_ins, n_tiles, per_step;
  uint64_t n_items;
};
QG_HD ItemPlan item_plan(bool dense, uint32_t parts, uint32_t n_entries, uint32_t cap_r, uint32_t cap_l) {
  ItemPlan P;
  const uint32_t region_w16 = dense ? n_entries / 2 : n_entries * 2;  // a region's table in 16-byte words
  P.n_clear = (region_w16 + J_CLEAR_W16 - 1) / J_CLEAR_W16;
  P.n_ins = (cap_r + J_INS_ROWS - 1) / J_INS_ROWS;
  P.n_tiles = cap_l / PJ_TILE;
  P.per_step = P.n_clear + P.n_ins + P.n_tiles;
  P.n_items = (uint64_t)(parts + 2) * P.per_step;
  return P;
}

template <bool DENSE>
__global__ void __launch_bounds__(JT, DENSE ? 4 : 3) k_pj_join(const __grid_constant__ PJMap M, const __grid_constant__ PJJoinArgs A,
                                                  const __grid_constant__ PJEmit E) {
  extern __shared__ __align__(128) unsigned char join_smem_raw[];
  JoinSmem& S = *reinterpret_cast<JoinSmem*>(join_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t parts = M.parts;
  const uint32_t region_w16 = DENSE ? A.n_entries / 2 : A.n_entries * 2;
  const ItemPlan IP = item_plan(DENSE, parts, A.n_entries, A.cap_r, A.cap_l);
  const uint64_t* pay_col = E.prefetch_left >= 0 ? (const uint64_t*)E.left[E.prefetch_left].in : nullptr;
  const uint64_t pol_stream = l2_policy_evict_first();  // read / written once
  const uint64_t pol_table = l2_policy_evict_last();     // a region's table: cleared, filled and probed within ~3 steps

  // item -> (kind, region, index inside the region); IT_NONE outside the pipeline's ends or when it holds no rows
  auto decode = [&](uint64_t item, uint32_t& p, uint32_t& j) -> uint32_t {
    if (item >= IP.n_items) return IT_NONE;
    const uint32_t step = (uint32_t)(item / IP.per_step);
    j = (uint32_t)(item - (uint64_t)step * IP.per_step);
    if (j < IP.n_clear) {
      p = step;
      return step < parts ? IT_CLEAR : IT_NONE;
    }
    j -= IP.n_clear;
    if (j < IP.n_ins) {
      p = step - 1;
      return (step >= 1 && step <= parts) ? IT_INSERT : IT_NONE;
    }
    j -= IP.n_ins;
    p = step - 2;
    if (step < 2) return IT_NONE;
    return j * PJ_TILE < min(A.count_l[p * PJ_CTR_STRIDE], A.cap_l) ? IT_PROBE : IT_NONE;
  };
  auto issue = [&](uint32_t stage, uint32_t p, uint32_t j) {  // one thread: stage the probe tile's keys (+ payload)
    const uint64_t row = (uint64_t)p * A.cap_l + (uint64_t)j * PJ_TILE;
    mbar_expect_tx(&S.bar[stage], (pay_col ? 2u : 1u) * PJ_TILE * 8u);
    bulk_g2s(S.keys[stage], A.probe_key + row, PJ_TILE * 8u, &S.bar[stage], pol_stream);
    if (pay_col) bulk_g2s(S.pay[stage], pay_col + row, PJ_TILE * 8u, &S.bar[stage], pol_stream);
  };
  auto finish = [&](uint32_t* counter) {  // the item's global writes happen-before the counter's increment
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
    }
  };
  unsigned long long* t_wait_ptr = nullptr;
  auto wait_for = [&](const uint32_t* counter, uint32_t target) {
    if (tid == 0) {
      const unsigned long long t0 = A.timing ? globaltimer_ns() : 0ull;
      while (ld_acquire_u32(counter) < target) __nanosleep(32);
      if (A.timing && t_wait_ptr) *t_wait_ptr += globaltimer_ns() - t0;
    }
    __syncthreads();
  };

  if (tid == 0) {
    for (int i = 0; i < PSTAGES; ++i) mbar_init(&S.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // Items are handed out in order by a ticket (work-conserving: a CTA that drew cheap items simply draws more, so no
  // region's inserts sit behind a slow CTA).  A CTA holds the current item, the next one (whose probe tile, if any,
  // is already being staged) and has the ticket after that in flight: the atomic's round trip overlaps the current
  // item's work.  Items are processed in increasing order, so the deadlock argument above still holds.
  // Staged probe tiles are numbered as they are issued; tile number t uses buffer t & 1.
  auto take = [&]() -> uint64_t {  // blocking draw (prologue only)
    if (tid == 0) S.ticket = atomicAdd(A.ticket, 1u);
    __syncthreads();
    const uint32_t t = S.ticket;
    __syncthreads();
    return t;
  };
  uint32_t issued = 0, consumed = 0, phase_bits = 0, ready_region = 0xffffffffu;
  uint32_t p = 0, j = 0, p_next = 0, j_next = 0;
  uint64_t item = take(), item_next = take();
  uint32_t kind = decode(item, p, j), kind_next = decode(item_next, p_next, j_next);
  if (kind == IT_PROBE) {
    if (tid == 0) issue(issued & 1u, p, j);
    ++issued;
  }
  if (kind_next == IT_PROBE) {
    if (tid == 0) issue(issued & 1u, p_next, j_next);
    ++issued;
  }
  for (; item < IP.n_items;) {
    uint32_t ticket_after = 0;
    if (tid == 0) ticket_after = atomicAdd(A.ticket, 1u);  // consumed at the end of this iteration
    const uint32_t kind_cur = kind, pc = p, jc = j;
    unsigned long long t_begin = 0, t_wait = 0;
    t_wait_ptr = &t_wait;
    if (A.timing && tid == 0) t_begin = globaltimer_ns();
    // ---- rotate: (item, next) <- (next, after); runs at the end of the iteration, see `advance` ----
    auto advance = [&]() {
      if (A.timing && tid == 0) {  // debug: [kind*4+0] items, [+1] ns in the item, [+2] ns of that spent waiting
        const unsigned long long t_end = globaltimer_ns();
        atomicAdd(A.timing + kind_cur * 4 + 0, 1ull);
        atomicAdd(A.timing + kind_cur * 4 + 1, t_end - t_begin);
        atomicAdd(A.timing + kind_cur * 4 + 2, t_wait);
      }
      if (tid == 0) S.ticket = ticket_after;
      __syncthreads();
      item = item_next;
      kind = kind_next;
      p = p_next;
      j = j_next;
      item_next = S.ticket;
      kind_next = decode(item_next, p_next, j_next);
      if (kind_next == IT_PROBE) {  // its buffer was released by the barriers of the staged tile before `item`
        if (tid == 0) issue(issued & 1u, p_next, j_next);
        ++issued;
      }
      __syncthreads();  // S.ticket is rewritten by the next iteration
    };
    if (kind_cur == IT_NONE) {
      // outside the pipeline's ends, or a probe tile past the rows of its region: nothing to do.  (Insert items
      // without rows are IT_INSERT: the region's counter must still reach its target.)
      advance();
      continue;
    }

    if (kind_cur == IT_CLEAR) {
      ulonglong2* w = reinterpret_cast<ulonglong2*>(A.table) + (size_t)pc * region_w16;
      const uint32_t w0 = jc * J_CLEAR_W16, w1 = min(w0 + J_CLEAR_W16, region_w16);
      for (uint32_t i = w0 + tid; i < w1; i += JT) st_hint_v2u64(w + i, PJ_EMPTY, PJ_EMPTY, pol_table);
      finish(A.cleared + pc * PJ_CTR_STRIDE);
      advance();
      continue;
    }

    if (kind_cur == IT_INSERT) {
      const uint32_t i0 = jc * J_INS_ROWS;
      const uint32_t cnt = min(A.count_r[pc * PJ_CTR_STRIDE], A.cap_r);
      if (i0 < cnt) {
        wait_for(A.cleared + pc * PJ_CTR_STRIDE, IP.n_clear);
        const uint64_t base = (uint64_t)pc * A.cap_r;
        // BR independent rows per thread: all loads first, then the atomics of every pending row back to back (one
        // L2 round trip per step for all of them), until no row is pending
        unsigned long long k[BR], v[BR];
        uint32_t pend = 0;
#pragma unroll
        for (int r = 0; r < BR; ++r) {
          const uint32_t i = i0 + r * JT + tid;
          const bool live = i < cnt;
          k[r] = live ? (unsigned long long)ld_hint_u64((const uint64_t*)A.build_key + base + i, pol_stream) : PJ_EMPTY;
          v[r] = !live ? 0ull : (A.build_val ? (unsigned long long)ld_hint_u64(A.build_val + base + i, pol_stream) : (unsigned long long)(base + i));
          if (live) pend |= 1u << r;
        }
        if (DENSE) {
          // direct-addressed: entry (key - kmin); region pc owns entries [pc << shift, (pc + 1) << shift)
          unsigned long long* vals = reinterpret_cast<unsigned long long*>(A.table) + (size_t)pc * region_w16 * 2;
          const uint64_t first = (uint64_t)pc << M.shift;
          unsigned long long old[BR];
#pragma unroll
          for (int r = 0; r < BR; ++r) {
            if (!((pend >> r) & 1u)) continue;
            if (v[r] == PJ_EMPTY) atomicOr(A.flags, (uint32_t)PJ_SENTINEL);
            old[r] = atomicExch(vals + ((k[r] - (uint64_t)M.kmin) - first), v[r]);
          }
#pragma unroll
          for (int r = 0; r < BR; ++r)
            if (((pend >> r) & 1u) && old[r] != PJ_EMPTY) atomicOr(A.flags, (uint32_t)PJ_DUPLICATE);
        } else {
          PJBucket* region = reinterpret_cast<PJBucket*>(A.table) + (size_t)pc * A.n_entries;
          uint32_t b[BR], sub[BR];
#pragma unroll
          for (int r = 0; r < BR; ++r) {
            if (((pend >> r) & 1u) && k[r] == PJ_EMPTY) {
              atomicOr(A.flags, (uint32_t)PJ_SENTINEL);
              pend &= ~(1u << r);
            }
            b[r] = pj_bucket(mix64(k[r]), A.n_entries);
            sub[r] = 0;
          }
          for (uint32_t steps = 0; pend != 0 && steps < 2 * A.n_entries; ++steps) {
            unsigned long long old[BR];
#pragma unroll
            for (int r = 0; r < BR; ++r)
              if ((pend >> r) & 1u) old[r] = atomicCAS(&region[b[r]].k[sub[r]], PJ_EMPTY, k[r]);
#pragma unroll
            for (int r = 0; r < BR; ++r) {
              if (!((pend >> r) & 1u)) continue;
              if (old[r] == PJ_EMPTY) {
                region[b[r]].v[sub[r]] = v[r];
                pend &= ~(1u << r);
              } else if (old[r] == k[r]) {
                atomicOr(A.flags, (uint32_t)PJ_DUPLICATE);
                pend &= ~(1u << r);
              } else if (sub[r] == 0) {
                sub[r] = 1;
              } else {
                sub[r] = 0;
                b[r] = b[r] + 1 == A.n_entries ? 0 : b[r] + 1;
              }
            }
          }
        }
      }
      finish(A.inserted + pc * PJ_CTR_STRIDE);
      advance();
      continue;
    }

    // ---- probe tile jc of region pc ----
    const uint32_t stage = consumed & 1u;
    ++consumed;
    const uint32_t i0 = jc * PJ_TILE;
    const uint32_t cnt = min(A.count_l[pc * PJ_CTR_STRIDE], A.cap_l);
    const uint64_t row0 = (uint64_t)pc * A.cap_l + i0 + tid;
    if (ready_region != pc) {  // the CTA's consecutive tiles are mostly of one region: one acquire per region
      wait_for(A.inserted + pc * PJ_CTR_STRIDE, IP.n_ins);
      ready_region = pc;
    }
    {
      const unsigned long long t0 = (A.timing && tid == 0) ? globaltimer_ns() : 0ull;
      mbar_wait(&S.bar[stage], (phase_bits >> stage) & 1u);
      if (A.timing && tid == 0) atomicAdd(A.timing + 20, globaltimer_ns() - t0);  // [20]: ns waiting for staged tiles
    }
    phase_bits ^= 1u << stage;
    const uint64_t* skeys = S.keys[stage] + tid;  // row r of this thread: skeys[r * JT]; valid until advance()
    const uint64_t* spay = S.pay[stage] + tid;

    unsigned long long val[PR];
    uint32_t live_bits = 0, found_bits = 0;
#pragma unroll
    for (int r = 0; r < PR; ++r)
      if (i0 + r * JT + tid < cnt) live_bits |= 1u << r;
    if (DENSE) {
      // one 8-byte load per row, absent = ~0.  Rows of this region have (key - kmin) >> shift == pc, except keys
      // outside the build range that a LEFT join parks in the last region.
      const unsigned long long* vals = reinterpret_cast<const unsigned long long*>(A.table) + (size_t)pc * region_w16 * 2;
      const uint64_t first = (uint64_t)pc << M.shift;
#pragma unroll
      for (int r = 0; r < PR; ++r) {
        const uint64_t idx = skeys[r * JT] - (uint64_t)M.kmin;
        val[r] = (((live_bits >> r) & 1u) && idx < M.range) ? ld_entry(vals + (idx - first), pol_table) : PJ_EMPTY;
      }
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (val[r] != PJ_EMPTY) found_bits |= 1u << r;
    } else {
      const PJBucket* region = reinterpret_cast<const PJBucket*>(A.table) + (size_t)pc * A.n_entries;
      // two rounds of PR / 2 rows (register budget); one bucket per pending row per step, the rows of a round in
      // flight together
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        constexpr int HR = PR / 2;
        unsigned long long k[HR];
        uint32_t b[HR];
        uint32_t pend = 0;
#pragma unroll
        for (int q = 0; q < HR; ++q) {
          const int r = h * HR + q;
          k[q] = skeys[r * JT];
          b[q] = pj_bucket(mix64(k[q]), A.n_entries);
          val[r] = 0;
          if (((live_bits >> r) & 1u) && k[q] != PJ_EMPTY) pend |= 1u << q;
        }
        for (uint32_t steps = 0; pend != 0 && steps < A.n_entries; ++steps) {
          unsigned long long k0[HR], k1[HR], v0[HR], v1[HR];
#pragma unroll
          for (int q = 0; q < HR; ++q)
            if ((pend >> q) & 1u) ld_bucket(region + b[q], k0[q], k1[q], v0[q], v1[q], pol_table);
#pragma unroll
          for (int q = 0; q < HR; ++q) {
            if (!((pend >> q) & 1u)) continue;
            if (k0[q] == k[q] || k1[q] == k[q]) {
              val[h * HR + q] = k0[q] == k[q] ? v0[q] : v1[q];
              found_bits |= 1u << (h * HR + q);
              pend &= ~(1u << q);
            } else if (k0[q] == PJ_EMPTY || k1[q] == PJ_EMPTY) {
              pend &= ~(1u << q);
            } else {
              b[q] = b[q] + 1 == A.n_entries ? 0 : b[q] + 1;
            }
          }
        }
      }
    }
    if (A.timing && tid == 0) atomicAdd(A.timing + 16, globaltimer_ns() - t_begin);  // [16] begin -> lookups done (thread 0)
    // ---- output positions: round-major so that a fully matching warp writes 32 consecutive rows per store ----
    const uint32_t emit_bits = E.left_join ? live_bits : found_bits;
    uint32_t lrank[PR];
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (emit_bits >> r) & 1u);
      lrank[r] = __popc(bal & lanemask_lt());
      if (lane == 0) S.cnt[r * (JT / 32) + warp] = __popc(bal);
    }
    __syncthreads();
    if (A.timing && tid == 0) atomicAdd(A.timing + 17, globaltimer_ns() - t_begin);  // [17] -> all warps' lookups done
    if (warp == 0) {  // exclusive scan of the PR x warps (round, warp) counts: up to two per lane
      constexpr int NC = PR * (JT / 32);
      static_assert(NC <= 64, "the scan below handles up to 64 counts");
      const uint32_t c0 = lane < NC ? S.cnt[lane] : 0u, c1 = lane + 32 < NC ? S.cnt[lane + 32] : 0u;
      uint32_t t0, t1;
      const uint32_t e0 = warp_excl_scan_u32(c0, t0);
      const uint32_t e1 = warp_excl_scan_u32(c1, t1);
      if (lane < NC) S.cnt[lane] = e0;
      if (lane + 32 < NC) S.cnt[lane + 32] = t0 + e1;
      if (lane == 0) S.base = (t0 + t1) ? atomicAdd(A.out_cursor, (unsigned long long)(t0 + t1)) : 0ull;
    }
    __syncthreads();
    if (A.timing && tid == 0) atomicAdd(A.timing + 18, globaltimer_ns() - t_begin);  // [18] -> output space reserved
    const unsigned long long base = S.base;
    uint32_t off[PR];  // output row = base + off[r]
#pragma unroll
    for (int r = 0; r < PR; ++r) off[r] = S.cnt[r * (JT / 32) + warp] + lrank[r];
#define PJ_EMIT(r) ((emit_bits >> (r)) & 1u)
#define PJ_FOUND(r) ((found_bits >> (r)) & 1u)

    __syncthreads();  // S.cnt / S.base are rewritten by the next tile
    // ---- left columns: straight copy of the probe row (key and prefetched payload from the staged tile) ----
    // (Tried and dropped: compacting the tile in shared memory and writing each column with one cp.async.bulk store.
    // It costs 32 KB more shared memory per CTA = one CTA per SM less, and the tile got slower, 1.17 -> 1.40 ms.)
    for (int ci = 0; ci < E.n_left; ++ci) {
      const PJCol c = E.left[ci];
      if (ci == E.left_key || ci == E.prefetch_left) {
        const uint64_t* src = ci == E.left_key ? skeys : spay;
        uint64_t* o = (uint64_t*)c.out + base;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (PJ_EMIT(r)) st_evict_first_u64(o + off[r], src[r * JT], pol_stream);
      } else if (c.width == 8) {
        uint64_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = PJ_EMIT(r) ? ld_stream_u64((const uint64_t*)c.in + row0 + r * JT) : 0ull;
        uint64_t* o = (uint64_t*)c.out + base;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (PJ_EMIT(r)) st_evict_first_u64(o + off[r], v[r], pol_stream);
      } else {
        uint32_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = PJ_EMIT(r) ? ld_stream_u32((const uint32_t*)c.in + row0 + r * JT) : 0u;
        uint32_t* o = (uint32_t*)c.out + base;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (PJ_EMIT(r)) o[off[r]] = v[r];
      }
    }
    // ---- right columns: the key is the probe key, the inlined payload rides in the table, others are gathered ----
    for (int ci = 0; ci < E.n_right; ++ci) {
      const PJCol c = E.right[ci];
      if (ci == E.right_key) {
        uint64_t* o = (uint64_t*)c.out + base;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (PJ_EMIT(r)) st_evict_first_u64(o + off[r], PJ_FOUND(r) ? skeys[r * JT] : 0ull, pol_stream);
      } else if (ci == E.inline_right) {
        uint64_t* o = (uint64_t*)c.out + base;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (PJ_EMIT(r)) st_evict_first_u64(o + off[r], PJ_FOUND(r) ? val[r] : 0ull, pol_stream);
      } else if (c.width == 8) {
        uint64_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = PJ_FOUND(r) ? __ldg((const unsigned long long*)c.in + val[r]) : 0ull;
        uint64_t* o = (uint64_t*)c.out + base;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (PJ_EMIT(r)) st_evict_first_u64(o + off[r], v[r], pol_stream);
      } else {
        uint32_t v[PR];
#pragma unroll
        for (int r = 0; r < PR; ++r) v[r] = PJ_FOUND(r) ? __ldg((const uint32_t*)c.in + val[r]) : 0u;
        uint32_t* o = (uint32_t*)c.out + base;
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if (PJ_EMIT(r)) o[off[r]] = v[r];
      }
    }
    if (E.right_valid_bytes != nullptr) {
      uint8_t* o = E.right_valid_bytes + base;
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (PJ_EMIT(r)) o[off[r]] = PJ_FOUND(r) ? 1 : 0;
    }
#undef PJ_EMIT
#undef PJ_FOUND
    if (A.timing && tid == 0) atomicAdd(A.timing + 19, globaltimer_ns() - t_begin);  // [19] -> stores issued
    advance();
  }
}

}  // namespace

// function attributes are per device: run `f` the first time device c->device reaches call site `site` (thread-safe)
template <class F>
static void once_per_device(Ctx* c, int site, F f) {
  static std::mutex mu;
  static std::set<std::pair<int, int>> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.insert({c->device, site}).second) f();
}

void launch_pj_scatter(Ctx* c, const int64_t* key, int64_t n, const PJMap& map, uint32_t cap, uint32_t* cursor, uint32_t* flags,
                       const PJCols& cols, const PJPeers* peers, const char* name) {
  if (n <= 0) return;
  const size_t smem = sizeof(ScatterSmem) + (size_t)SWARPS * (map.parts + 1) * sizeof(uint32_t);
  once_per_device(c, 0, [] {
    const int max_smem = (int)(sizeof(ScatterSmem) + SWARPS * (PJ_MAX_PARTS + 1) * 4);
    QG_CUDA(cudaFuncSetAttribute(k_pj_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    QG_CUDA(cudaFuncSetAttribute(k_pj_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  });
  ProfScope ps(c, name);
  const int64_t tiles = (n + STILE - 1) / STILE;
  if (peers) {
    k_pj_scatter<true><<<persistent_grid(c, k_pj_scatter<true>, ST, tiles, smem), ST, smem, c->stream>>>(key, n, map, cap, cursor, flags, cols,
                                                                                                      *peers);
  } else {
    PJPeers none;
    memset(&none, 0, sizeof none);
    k_pj_scatter<false><<<persistent_grid(c, k_pj_scatter<false>, ST, tiles, smem), ST, smem, c->stream>>>(key, n, map, cap, cursor, flags,
                                                                                                        cols, none);
  }
  QG_CUDA(cudaGetLastError());
}

void launch_pj_join(Ctx* c, const PJMap& map, const PJJoinArgs& args, const PJEmit& emit) {
  ProfScope ps(c, map.dense ? "k_pj_join_dense" : "k_pj_join_hashed");
  const ItemPlan ip = item_plan(map.dense != 0, map.parts, args.n_entries, args.cap_r, args.cap_l);
  const void* kernel = map.dense ? (const void*)k_pj_join<true> : (const void*)k_pj_join<false>;
  const size_t smem = sizeof(JoinSmem);
  once_per_device(c, 1, [smem] {
    QG_CUDA(cudaFuncSetAttribute(k_pj_join<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_CUDA(cudaFuncSetAttribute(k_pj_join<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  const unsigned grid = map.dense ? persistent_grid(c, k_pj_join<true>, JT, (int64_t)ip.n_items, smem)
                                  : persistent_grid(c, k_pj_join<false>, JT, (int64_t)ip.n_items, smem);
  void* params[] = {(void*)&map, (void*)&args, (void*)&emit};
  // cooperative: every CTA of the grid must be resident for the item pipeline's waits to be deadlock-free
  QG_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(JT), params, smem, c->stream));
}

}  // namespace qgpu
