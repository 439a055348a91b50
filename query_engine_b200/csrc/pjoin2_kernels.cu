// pjoin2_kernels.cu — build + probe of the radix-partitioned hash join, second generation.
//
// What changed against k_pj_join (pjoin_kernels.cu), and why (profiles/r01h_ncu_full.csv: 0.89 ms, 3.5 GB of DRAM
// traffic for a kernel whose inputs and outputs are 3.1 GB, every probe tile a chain lookup -> ballot -> barrier ->
// cursor atomic -> barrier -> stores):
//   * The probe side is not copied region-major and then compacted into the output: the radix scatter
//     (rscatter_kernels.cu, exact placement from a region histogram) has already written the probe-side OUTPUT
//     columns, region after region.  This kernel reads the probe keys back from the output key column, looks them up
//     and writes the build-side output columns at the SAME row: no output cursor, no compaction, no barrier inside a
//     probe tile.  A LEFT join is complete after that; an INNER join is complete when every probe row matched (the
//     foreign-key case) - otherwise the host compacts the output once by the match flags.
//   * Region tables live in a RING of `ring` slots that is reused every `ring` regions instead of one table per
//     region: the ring (<= 64 MB) stays in L2, a table line is cleared, filled, probed and cleared again without ever
//     being written back to HBM (the 0.4 GB write-back of the first version), and table memory is bounded.
// (Tried and dropped in round 2, each measured at SF10 against 0.695 ms for this form - profiles/r02k_join2_stalls.md holds
// the stall profile that motivated them: a warp-specialised form, one producer warp staging items into a 3-4 stage
// shared-memory ring for eight consumer warps, completion counted per warp, no CTA barrier: 1.14-1.18 ms, the warps of
// a CTA drift over several items and every item completes later, which the next step's dependency waits amplify; the
// clear items at the END of a step instead of its start: +0.11 ms, the inserts that open the next step then wait; 256 KB
// or 32 KB clear items instead of 64 KB: +0.17 / +0.10 ms; a few more live registers in this loop - it sits exactly at
// the 64-register limit of four CTAs per SM - spill and cost 0.17 ms.)
// (Tried and dropped: every WARP an independent worker drawing its own items - no CTA barrier at all.  With ~2 800 items
// per region step and 4 736 resident warps an item's latency, not throughput, sets the pace: 1.88 ms against 0.70 ms.)
// The item pipeline is the one of k_pj_join: ticket-ordered items [clear region s] [insert s-1] [probe s-2], waits only
// ever target lower tickets, all CTAs co-resident (cooperative launch).  A clear item additionally waits until the
// ring slot's previous tenant (region s - ring) has been probed completely.
//
// Replaces the body of QueryExecutor::execute_join (crates/query-executor/src/executor.rs:343-399) for `probe.key =
// build.key` on Int64 keys, duplicate-free build side, INNER / LEFT; output schema left ++ right (executor.rs:532-537).
#include <mutex>
#include <set>

#include "pj_device.cuh"
#include "pjoin2_kernels.cuh"

namespace qgpu {
namespace {

constexpr int JT = 256;                 // threads
constexpr uint32_t J_CLEAR_W16 = 4096;  // 64 KB of table per clear item (16-byte words)
constexpr uint32_t J_INS_ROWS = 1024;   // build rows per insert item
constexpr int BR = J_INS_ROWS / JT;
constexpr int PR = PJ2_TILE / JT;       // probe rows per thread
constexpr int PSTAGES = 2;

struct Join2Smem {
  uint64_t keys[PSTAGES][PJ2_TILE];
  uint64_t bar[PSTAGES];
  uint32_t ticket;
  uint32_t pad;
  uint32_t start[PJ_MAX_PARTS + 2];  // start_l, read once: item decoding stays off the L2
};

enum ItemKind : uint32_t { IT_CLEAR = 0, IT_INSERT = 1, IT_PROBE = 2, IT_NONE = 3 };

struct ItemPlan2 {
  uint32_t n_clear, n_ins, n_tiles, per_step;
  uint64_t n_items;
};
QG_HD ItemPlan2 item_plan2(bool dense, uint32_t parts, uint32_t n_entries, uint32_t cap_r, uint32_t tiles_per_region) {
  ItemPlan2 P;
  const uint32_t region_w16 = dense ? n_entries / 2 : n_entries * 2;  // a region's table in 16-byte words
  P.n_clear = (region_w16 + J_CLEAR_W16 - 1) / J_CLEAR_W16;
  P.n_ins = (cap_r + J_INS_ROWS - 1) / J_INS_ROWS;
  P.n_tiles = tiles_per_region;
  P.per_step = P.n_clear + P.n_ins + P.n_tiles;
  P.n_items = (uint64_t)(parts + 2) * P.per_step;
  return P;
}

template <bool DENSE>
__global__ void __launch_bounds__(JT, DENSE ? 4 : 3) k_pj_join2(const __grid_constant__ PJMap M, const __grid_constant__ PJ2Args A) {
  extern __shared__ __align__(128) unsigned char join2_smem_raw[];
  Join2Smem& S = *reinterpret_cast<Join2Smem*>(join2_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t parts = M.parts;
  const uint32_t region_w16 = DENSE ? A.n_entries / 2 : A.n_entries * 2;
  const ItemPlan2 IP = item_plan2(DENSE, parts, A.n_entries, A.cap_r, A.tiles_per_region);
  const uint64_t pol_stream = l2_policy_evict_first();  // read / written once
  const uint64_t pol_table = l2_policy_evict_last();     // the ring of region tables: reused for the whole kernel

  // probe windows are PJ2_TILE-aligned row ranges of the output; region p owns rows [start_l[p], start_l[p + 1])
  for (uint32_t i = tid; i <= parts; i += JT) S.start[i] = A.start_l[i];
  auto live_tiles = [&](uint32_t p) -> uint32_t {
    const uint32_t s = S.start[p], e = S.start[p + 1];
    return e > s ? (e - 1) / PJ2_TILE - s / PJ2_TILE + 1 : 0u;
  };
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
    return j < live_tiles(p) ? IT_PROBE : IT_NONE;
  };
  auto issue = [&](uint32_t stage, uint32_t p, uint32_t j) {  // one thread: stage the window's keys
    const uint64_t row = ((uint64_t)(S.start[p] / PJ2_TILE) + j) * PJ2_TILE;
    const uint64_t left = A.probe_rows_alloc - row;
    const uint32_t bytes = (uint32_t)(left < PJ2_TILE ? ((left * 8 + 15) & ~15ull) : PJ2_TILE * 8ull);
    mbar_expect_tx(&S.bar[stage], bytes);
    bulk_g2s(S.keys[stage], A.probe_key + row, bytes, &S.bar[stage], pol_stream);
  };
  auto finish = [&](uint32_t* counter) {  // the item's global writes happen-before the counter's increment
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
    }
  };
  auto wait_for = [&](const uint32_t* counter, uint32_t target) {
    if (tid == 0) {
      // bounded (~2 s): a pipeline bug must surface as a flag (the host then takes another path), never as a hung GPU
      uint32_t spins = 0;
      while (ld_acquire_u32(counter) < target) {
        __nanosleep(64);
        if (++spins > (1u << 25)) {
          atomicOr(A.flags, (uint32_t)PJ_TIMEOUT);
          break;
        }
      }
    }
    __syncthreads();
  };

  if (tid == 0) {
    for (int i = 0; i < PSTAGES; ++i) mbar_init(&S.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
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
  uint32_t n_found = 0;  // this thread's matched probe rows
  for (; item < IP.n_items;) {
    uint32_t ticket_after = 0;
    if (tid == 0) ticket_after = atomicAdd(A.ticket, 1u);  // consumed at the end of this iteration
    const uint32_t kind_cur = kind, pc = p, jc = j;
    auto advance = [&]() {  // (item, next) <- (next, after)
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
      advance();
      continue;
    }
    const uint32_t slot = pc % A.ring;  // the region's table in the ring

    if (kind_cur == IT_CLEAR) {
      // the slot's previous tenant must have been probed completely
      if (pc >= A.ring) wait_for(A.probed + (pc - A.ring) * PJ_CTR_STRIDE, live_tiles(pc - A.ring));
      ulonglong2* w = reinterpret_cast<ulonglong2*>(A.table) + (size_t)slot * region_w16;
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
          unsigned long long* vals = reinterpret_cast<unsigned long long*>(A.table) + (size_t)slot * region_w16 * 2;
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
          PJBucket* region = reinterpret_cast<PJBucket*>(A.table) + (size_t)slot * A.n_entries;
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

    // ---- probe window jc of region pc: rows [row_base, row_base + PJ_TILE) that belong to the region ----
    const uint32_t stage = consumed & 1u;
    ++consumed;
    const uint32_t lo = S.start[pc], hi = S.start[pc + 1];
    const uint32_t row_base = (lo / PJ2_TILE + jc) * PJ2_TILE;
    if (ready_region != pc) {  // the CTA's consecutive windows are mostly of one region: one acquire per region
      wait_for(A.inserted + pc * PJ_CTR_STRIDE, IP.n_ins);
      ready_region = pc;
    }
    mbar_wait(&S.bar[stage], (phase_bits >> stage) & 1u);
    phase_bits ^= 1u << stage;
    const uint64_t* skeys = S.keys[stage] + tid;  // row r of this thread: skeys[r * JT]; valid until advance()

    unsigned long long val[PR];
    uint32_t live_bits = 0, found_bits = 0;
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const uint32_t row = row_base + r * JT + tid;
      if (row >= lo && row < hi) live_bits |= 1u << r;
    }
    if (DENSE) {
      // one 8-byte load per row, absent = ~0.  Rows of this region have (key - kmin) >> shift == pc, except keys
      // outside the build range that a LEFT join parks in the last region.
      const unsigned long long* vals = reinterpret_cast<const unsigned long long*>(A.table) + (size_t)slot * region_w16 * 2;
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
      const PJBucket* region = reinterpret_cast<const PJBucket*>(A.table) + (size_t)slot * A.n_entries;
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
    n_found += __popc(found_bits);
    // ---- build-side output columns, at the probe row's own position ----
#define PJ_LIVE(r) ((live_bits >> (r)) & 1u)
#define PJ_FOUND(r) ((found_bits >> (r)) & 1u)
    if (A.out_right_key) {
      uint64_t* o = A.out_right_key + row_base + tid;
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (PJ_LIVE(r)) st_evict_first_u64(o + r * JT, PJ_FOUND(r) ? skeys[r * JT] : 0ull, pol_stream);
    }
    if (A.out_inline) {
      uint64_t* o = A.out_inline + row_base + tid;
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (PJ_LIVE(r)) st_evict_first_u64(o + r * JT, PJ_FOUND(r) ? val[r] : 0ull, pol_stream);
    }
    for (int ci = 0; ci < A.n_gather; ++ci) {  // other build columns: gathered from the (L2-resident) region copy
      const PJCol c = A.gather[ci];
      uint64_t v[PR];
#pragma unroll
      for (int r = 0; r < PR; ++r) v[r] = PJ_FOUND(r) ? __ldg((const unsigned long long*)c.in + val[r]) : 0ull;
      uint64_t* o = (uint64_t*)c.out + row_base + tid;
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (PJ_LIVE(r)) st_evict_first_u64(o + r * JT, v[r], pol_stream);
    }
    if (A.vbytes) {
      uint8_t* o = A.vbytes + row_base + tid;
#pragma unroll
      for (int r = 0; r < PR; ++r)
        if (PJ_LIVE(r)) o[r * JT] = PJ_FOUND(r) ? 1 : 0;
    }
#undef PJ_LIVE
#undef PJ_FOUND
    // (no fence: the ring slot's next clear must come after this window's table READS, and those have completed - their
    // values were consumed above; the output stores need not be visible to anybody inside this kernel)
    __syncthreads();
    if (tid == 0) atomicAdd(A.probed + pc * PJ_CTR_STRIDE, 1u);
    advance();
  }
  // matched rows of this CTA -> one atomic
  n_found = __reduce_add_sync(0xffffffffu, n_found);
  if (lane == 0 && n_found) atomicAdd(A.found, (unsigned long long)n_found);
}

template <class F>
void once_per_device(Ctx* c, int site, F f) {
  static std::mutex mu;
  static std::set<std::pair<int, int>> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.insert({c->device, site}).second) f();
}

}  // namespace

uint32_t pj2_tiles_per_region(int64_t probe_rows, uint32_t parts) {
  uint64_t cap = (uint64_t)probe_rows / parts;
  cap += cap / 4 + 2048;  // 25 % slack over an even spread; a region that needs more raises PJ_OVERFLOW (launch_region_offsets)
  return (uint32_t)((cap + PJ2_TILE - 1) / PJ2_TILE) + 1;
}

void launch_pj_join2(Ctx* c, const PJMap& map, const PJ2Args& args) {
  ProfScope ps(c, map.dense ? "k_pj_join2_dense" : "k_pj_join2_hashed");
  const ItemPlan2 ip = item_plan2(map.dense != 0, map.parts, args.n_entries, args.cap_r, args.tiles_per_region);
  const void* kernel = map.dense ? (const void*)k_pj_join2<true> : (const void*)k_pj_join2<false>;
  const size_t smem = sizeof(Join2Smem);
  once_per_device(c, 0, [smem] {
    QG_CUDA(cudaFuncSetAttribute(k_pj_join2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QG_CUDA(cudaFuncSetAttribute(k_pj_join2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  const unsigned grid = map.dense ? persistent_grid(c, k_pj_join2<true>, JT, (int64_t)ip.n_items, smem)
                                  : persistent_grid(c, k_pj_join2<false>, JT, (int64_t)ip.n_items, smem);
  void* params[] = {(void*)&map, (void*)&args};
  // cooperative: every CTA of the grid must be resident for the item pipeline's waits to be deadlock-free
  QG_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(JT), params, smem, c->stream));
}

}  // namespace qgpu
