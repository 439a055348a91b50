// scan_body.cuh — the body of the fused column-scan kernels (sm_100a), shared by the ahead-of-time interpreter
// kernels (scan_kernels.cu) and the NVRTC-specialised ones (jit.cu).  Device-safe: no standard-library includes.
//
//   SK_MASK          predicate -> selection bitmap + per-tile survivor counts   (Filter, executor.rs:131-155)
//   SK_SCALAR        predicate + ungrouped aggregates in one pass               (HashAggregate scalar path, :167-187)
//   SK_GROUP_SMEM    predicate + low-cardinality GROUP BY, per-thread private accumulators in shared memory
//   SK_GROUP_GLOBAL  predicate + GROUP BY through an open-addressing table in HBM
//   SK_STORE         expression materialisation (Projection, :93-129)
//
// Execution model: persistent grid (sm_count x CTAs/SM), 256 threads, each thread evaluates R rows per tile; a tile
// is 256*R consecutive rows (row(r) = base + r*256 + tid: every load is a fully coalesced 256 B per warp, and every
// warp-level ballot is one aligned 32-row bitmap word).  Column values are read exactly once with streaming loads;
// columns referenced only after the predicate are fetched for surviving rows only (sectors without a survivor are
// never requested from HBM); the next tile of every referenced column is bulk-prefetched into L2 while the current one
// is evaluated.  Everything is HBM-bound integer / f64 work: no tensor cores.
//
// `Prog` is the program policy:
//   InterpProg  (scan_kernels.cu)  instruction fields are run-time values read from shared memory
//   JitProg     (generated)        one exec_insn call per instruction with literal fields; column types, validity
//                                  presence and accumulator counts are constexpr, so loads are typed at compile time,
//                                  accumulators and the spill stack live in registers.
#pragma once
#include "device_utils.cuh"
#include "scan_params.cuh"

// SK_GROUP_PART, specialised kernels: request every referenced column of a tile up front ("group.preload")
#ifndef QGPU_GP_PRELOAD
#define QGPU_GP_PRELOAD 0
#endif

namespace qgpu {

constexpr int SCAN_NW = SCAN_THREADS / WARP;

template <bool B>
struct BoolTag {
  static constexpr bool value = B;
};

// ---------------- loader: row(r) = base + r*T + tid -------------------------------------------
template <class Prog, int RR>
struct DevLoader {
  const ColRef* cols;
  int64_t base;
  uint32_t want;  // rows to fetch: the in-bounds mask before OP_GATE, the surviving rows after it
  // specialised kernels keep every column they have fetched for this tile in registers: `have` and the slot indices
  // are compile-time values along the straight-line program, so a second reference to a column costs nothing
  uint32_t have;
  Lanes<RR> cache[Prog::STATIC ? (Prog::N_COLS > 0 ? Prog::N_COLS : 1) : 1];

  QG_D void gate(uint32_t pass) { want = pass; }
  // specialised kernels, when (nearly) every row survives the predicate: fetch every referenced column up front, all
  // loads in flight together, instead of one dependent round trip for the predicate's columns and a second one for the
  // rest (Q1 keeps 98 % of its rows: gating the loads on the survivors saves nothing there)
  QG_D void preload_all() {
    if constexpr (Prog::STATIC) {
#pragma unroll
      for (int s = 0; s < Prog::N_COLS; ++s) load_fresh<RR>(s, cache[s]);
      have = Prog::N_COLS >= 32 ? 0xffffffffu : ((1u << Prog::N_COLS) - 1u);
    }
  }
  QG_D void new_tile(int64_t b, uint32_t inb) {
    base = b;
    want = inb;
    have = 0;
  }

  template <int R>
  QG_D void load(int slot, Lanes<R>& x) {
    if constexpr (Prog::STATIC) {
      if (have & (1u << slot)) {  // rows fetched earlier are a superset of the rows wanted now
        x = cache[slot];
        return;
      }
    }
    load_fresh<R>(slot, x);
    if constexpr (Prog::STATIC) {
      cache[slot] = x;
      have |= 1u << slot;
    }
  }

  template <int R>
  QG_D void load_fresh(int slot, Lanes<R>& x) const {
    constexpr int T = SCAN_THREADS;
    const int dt = Prog::col_dtype(slot, cols);
    const bool has_valid = Prog::col_has_valid(slot, cols);
    const int64_t row0 = base + threadIdx.x;
    uint32_t ok = want;
    switch (dt) {
      case DT_I64:
      case DT_F64: {
        const uint64_t* p = (const uint64_t*)cols[slot].data + row0;
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = ((want >> r) & 1u) ? ld_stream_u64(p + (int64_t)r * T) : 0ull;
        break;
      }
      case DT_I32: {
        const uint32_t* p = (const uint32_t*)cols[slot].data + row0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          x.v[r] = ((want >> r) & 1u) ? (uint64_t)(int64_t)(int32_t)ld_stream_u32(p + (int64_t)r * T) : 0ull;
        break;
      }
      case DT_F32: {
        const uint32_t* p = (const uint32_t*)cols[slot].data + row0;
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = ((want >> r) & 1u) ? (uint64_t)ld_stream_u32(p + (int64_t)r * T) : 0ull;
        break;
      }
      case DT_BOOL: {
        const uint32_t* p = (const uint32_t*)cols[slot].data;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int64_t row = row0 + (int64_t)r * T;
          x.v[r] = ((want >> r) & 1u) ? ((__ldg(p + (row >> 5)) >> (row & 31)) & 1u) : 0ull;
        }
        break;
      }
      case DT_NULL:
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = 0;
        ok = 0;
        break;
      default:  // Utf8: only validity is consumed (COUNT)
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = 0;
        break;
    }
    if (has_valid) {
      const uint32_t* vp = cols[slot].valid;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if ((want >> r) & 1u) {
          const int64_t row = row0 + (int64_t)r * T;
          const uint32_t w = __ldg(vp + (row >> 5));
          if (!((w >> (row & 31)) & 1u)) ok &= ~(1u << r);
        }
      }
    }
    x.ok = ok;
  }
};

// ---------------- spill stacks -----------------------------------------------------------------
struct SmemStack {  // interpreter: [depth][R][T] words + [depth][T] masks in shared memory
  uint64_t* v;
  uint32_t* ok;
  int sp;
  template <int R>
  QG_D void push(const Lanes<R>& a) {
    constexpr int T = SCAN_THREADS;
#pragma unroll
    for (int r = 0; r < R; ++r) v[(sp * R + r) * T + threadIdx.x] = a.v[r];
    ok[sp * T + threadIdx.x] = a.ok;
    ++sp;
  }
  template <int R>
  QG_D void pop(Lanes<R>& a) {
    constexpr int T = SCAN_THREADS;
    --sp;
#pragma unroll
    for (int r = 0; r < R; ++r) a.v[r] = v[(sp * R + r) * T + threadIdx.x];
    a.ok = ok[sp * T + threadIdx.x];
  }
};

template <int RR, int DEPTH>
struct RegStack {  // specialised: the program is straight-line, so `sp` is a compile-time value at every use
  Lanes<RR> s[DEPTH > 0 ? DEPTH : 1];
  int sp;
  template <int R>
  QG_D void push(const Lanes<R>& a) { s[sp++] = a; }
  template <int R>
  QG_D void pop(Lanes<R>& a) { a = s[--sp]; }
};

// value fed to an accumulator of the given kind
QG_D uint64_t agg_value(const Insn& in, uint64_t v) {
  switch (in.mode) {
    case ACC_ADD_F64: return in.ty == DT_F32 ? f64_bits((double)bits_f32(v)) : v;
    case ACC_MIN_I64:
    case ACC_MAX_I64: return (uint64_t)cmp_key(v, in.ty);
    default: return v;
  }
}

struct SinkNone {
  QG_D void fault(const Insn&, int, uint64_t, uint64_t, uint32_t) {}
  template <int R> QG_D void agg(const Insn&, const Lanes<R>&, uint32_t) {}
  template <int R> QG_D void store(const Insn&, const Lanes<R>&, uint32_t) {}
  template <int R> QG_D void key(const Insn&, const Lanes<R>&, uint32_t) {}
};

// ---------------- SK_SCALAR: per-thread private accumulators ------------------------------------
// interpreter: in shared memory [n_acc][T];  specialised: in registers (slot indices are literals)
template <class Prog>
struct SinkScalar : SinkNone {
  uint64_t* sacc;
  uint64_t racc[(Prog::STATIC && Prog::N_ACC > 0) ? Prog::N_ACC : 1];

  template <int R>
  QG_D void agg(const Insn& in, const Lanes<R>& a, uint32_t pass) {
    const uint32_t m = a.ok & pass;
    if constexpr (Prog::STATIC) {
      if (in.mode == ACC_COUNT) {
        racc[in.arg] += (uint64_t)__popc(m);
        return;
      }
      uint64_t cur = racc[in.arg];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if ((m >> r) & 1u) cur = acc_combine(in.mode, cur, agg_value(in, a.v[r]));
      racc[in.arg] = cur;
    } else {
      uint64_t* p = sacc + in.arg * SCAN_THREADS + threadIdx.x;
      if (in.mode == ACC_COUNT) {
        *p += (uint64_t)__popc(m);
        return;
      }
      if (m == 0) return;
      uint64_t cur = *p;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if ((m >> r) & 1u) cur = acc_combine(in.mode, cur, agg_value(in, a.v[r]));
      *p = cur;
    }
  }
};

// ---------------- key hashing / equality on table rows (group-by in HBM, also the smem-tier flush) ---
QG_D uint64_t hash_utf8(const ColRef& c, int64_t row) {
  const int32_t* off = (const int32_t*)c.data;
  const int32_t b = off[row], e = off[row + 1];
  uint64_t h = 0xcbf29ce484222325ull;
  for (int32_t i = b; i < e; ++i) h = (h ^ c.chars[i]) * 0x100000001b3ull;
  return h ^ (uint64_t)(e - b);
}

QG_D uint64_t hash_row(const GroupTable& g, int64_t row) {
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (int k = 0; k < g.n_keys; ++k) {
    const ColRef& c = g.keys[k];
    uint64_t v;
    if (!row_valid(c, row)) v = 0x6e756c6c6e756c6cull;
    else if (c.dtype == DT_UTF8) v = hash_utf8(c, row);
    else v = load_value(c, row);
    h = mix64(h ^ v) + (uint64_t)k;
  }
  return h;
}

QG_D bool rows_equal(const GroupTable& g, int64_t a, int64_t b) {
  if (a == b) return true;
  for (int k = 0; k < g.n_keys; ++k) {
    const ColRef& c = g.keys[k];
    const bool va = row_valid(c, a), vb = row_valid(c, b);
    if (va != vb) return false;
    if (!va) continue;
    if (c.dtype == DT_UTF8) {
      const int32_t* off = (const int32_t*)c.data;
      const int32_t ab = off[a], ae = off[a + 1], bb = off[b], be = off[b + 1];
      if (ae - ab != be - bb) return false;
      for (int32_t i = 0; i < ae - ab; ++i)
        if (c.chars[ab + i] != c.chars[bb + i]) return false;
    } else if (load_value(c, a) != load_value(c, b)) {
      return false;
    }
  }
  return true;
}

constexpr uint32_t GROUP_MAX_PROBES = 128;

QG_D uint32_t find_or_insert(const GroupTable& g, int64_t row, uint32_t& err) {
  const uint32_t mask = g.cap - 1;
  uint32_t s = group_home_slot(hash_row(g, row), g);
  // a chain this long means the table is (nearly) full: give up early - the host retries with a larger table -
  // instead of walking a full table once per row
  const uint32_t max_probes = mask < GROUP_MAX_PROBES ? mask : GROUP_MAX_PROBES;
  for (uint32_t probes = 0; probes <= max_probes; ++probes) {
    uint32_t cur = *((volatile uint32_t*)(g.rep + s));
    if (cur == 0) {
      const uint32_t prev = atomicCAS(g.rep + s, 0u, (uint32_t)row + 1u);
      if (prev == 0) return s;
      cur = prev;
    }
    if (rows_equal(g, row, (int64_t)cur - 1)) return s;
    s = (s + 1) & mask;
  }
  err |= ERR_CAPACITY;
  return 0;
}

QG_D void atomic_acc(uint8_t kind, uint64_t* p, uint64_t v) {
  switch (kind) {
    case ACC_ADD_F64: atomicAdd((double*)p, bits_f64(v)); break;
    case ACC_MIN_I64: atomicMin((long long*)p, (long long)v); break;
    case ACC_MAX_I64: atomicMax((long long*)p, (long long)v); break;
    default: atomicAdd((unsigned long long*)p, (unsigned long long)v); break;
  }
}

// ---------------- SK_GROUP_GLOBAL: every row goes through the HBM table ------------------------
template <int R>
struct SinkGroupGlobal : SinkNone {
  const GroupTable* g;
  uint32_t slot[R];
  uint32_t* err;
  template <class Loader>
  QG_D void group(Loader& ld, uint32_t pass) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      slot[r] = 0;
      if ((pass >> r) & 1u) slot[r] = find_or_insert(*g, ld.base + (int64_t)r * SCAN_THREADS + threadIdx.x, *err);
    }
  }
  template <int RR>
  QG_D void agg(const Insn& in, const Lanes<RR>& a, uint32_t pass) {
    const uint32_t m = a.ok & pass;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((m >> r) & 1u) {
        uint64_t* p = g->acc + (size_t)slot[r] * g->n_acc + in.arg;
        atomic_acc(in.mode, p, in.mode == ACC_COUNT ? 1ull : agg_value(in, a.v[r]));
      }
    }
  }
};

// ---------------- SK_GROUP_SMEM: <= gp groups per CTA, private accumulators, no atomics ---------
struct GroupDir {
  uint64_t* keys;      // [gp][n_keys]  register representation of the key value
  uint32_t* nullmask;  // [gp] bit k = key k is NULL
  uint32_t* row;       // [gp] a representative global row
  int* cnt;
  int* elect;
  int* overflow;
};

template <int R, class Prog>
struct SinkGroupSmem : SinkNone {
  GroupDir d;
  uint64_t* sacc;  // [(g*n_acc + s)][T]
  int n_acc, n_keys_rt, gp;
  const int32_t* key_slots;
  int gid[R];

  QG_D int nkeys() const { return Prog::STATIC ? Prog::N_KEYS : n_keys_rt; }
  QG_D int kslot(int k) const { return Prog::STATIC ? Prog::key_slot(k) : key_slots[k]; }

  template <class Loader>
  QG_D uint32_t lookup(Loader& ld, uint32_t want) {
    // returns the mask of wanted rows still without a group
    uint32_t cand[R];
#pragma unroll
    for (int r = 0; r < R; ++r) cand[r] = 0xffffffffu;
    const int cnt = *((volatile int*)d.cnt);
    const int n_keys = nkeys();
#pragma unroll
    for (int k = 0; k < n_keys; ++k) {
      Lanes<R> kw;
      ld.template load<R>(kslot(k), kw);
      for (int g = 0; g < cnt; ++g) {
        const uint64_t dk = d.keys[g * n_keys + k];
        const bool dn = (d.nullmask[g] >> k) & 1u;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool rn = !((kw.ok >> r) & 1u);
          const bool eq = dn ? rn : (!rn && kw.v[r] == dk);
          if (!eq) cand[r] &= ~(1u << g);
        }
      }
    }
    const uint32_t live = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
    uint32_t missing = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((want >> r) & 1u) {
        const uint32_t m = cand[r] & live;
        if (m) gid[r] = __ffs(m) - 1;
        else { gid[r] = -1; missing |= 1u << r; }
      }
    }
    return missing;
  }

  // returns false when this CTA overflowed its directory (uniform across the CTA)
  template <class Loader>
  QG_D bool group(Loader& ld, uint32_t pass) {
    const int n_keys = nkeys();
#pragma unroll
    for (int r = 0; r < R; ++r) gid[r] = -1;
    uint32_t missing = lookup(ld, pass);
    while (true) {
      if (!__syncthreads_or(missing != 0)) break;
      if (missing) atomicMin(d.elect, (int)threadIdx.x);
      __syncthreads();
      if ((int)threadIdx.x == *d.elect) {
        const int r0 = __ffs(missing) - 1;
        const int g = *d.cnt;
        if (g >= gp) {
          *d.overflow = 1;
        } else {
          uint32_t nm = 0;
#pragma unroll
          for (int k = 0; k < n_keys; ++k) {
            Lanes<R> kw;
            ld.template load<R>(kslot(k), kw);
            uint64_t val = 0;
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (r == r0) val = kw.v[r];
            if (!((kw.ok >> r0) & 1u)) nm |= 1u << k;
            d.keys[g * n_keys + k] = val;
          }
          d.nullmask[g] = nm;
          d.row[g] = (uint32_t)(ld.base + (int64_t)r0 * SCAN_THREADS + threadIdx.x);
          __threadfence_block();
          *d.cnt = g + 1;
        }
      }
      __syncthreads();
      if (*d.overflow) return false;
      if (threadIdx.x == 0) *d.elect = 0x7fffffff;
      if (missing) missing = lookup(ld, missing);
    }
    return true;
  }

  // (Tried and dropped, round 2: no CTA barrier in group() - the directory append-only under a spin lock (test-and-
  // test-and-set with back-off), everybody else running on.  0.644 against 0.658 ms for Q1 on one GPU, all parity tests
  // green, but 0.99 ms on the slowest of 8 ranks in the only 8-GPU run (profiles/r02y_bench_n8.json) with no GPU time
  // left to find out why: the three barriers per tile stay.  Also untested and therefore not kept: comparing one
  // 32-bit tag per (row, group) and verifying the first match only - lookup() is ~37 % of Q1's 200 instructions per row.)
  // (Tried and dropped: combining the rows of a thread that share a group in registers before one shared-memory update -
  // 40 % fewer read-modify-writes, but the data-dependent member masks cost more than they save: 0.67 -> 1.05 ms on Q1.)
  template <int RR>
  QG_D void agg(const Insn& in, const Lanes<RR>& a, uint32_t pass) {
    const uint32_t m = a.ok & pass;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((m >> r) & 1u) {
        uint64_t* p = sacc + (size_t)(gid[r] * n_acc + in.arg) * SCAN_THREADS + threadIdx.x;
        *p = acc_combine(in.mode, *p, in.mode == ACC_COUNT ? 1ull : agg_value(in, a.v[r]));
      }
    }
  }
};

// ---------------- SK_GROUP_PART: one partition's groups in a shared-memory table --------------------
// Open addressing on the 8-byte key word itself (~0 marks an empty slot; a key equal to the marker owns the extra slot
// `slots`); accumulators [slot][n_acc] are updated with shared-memory atomics.  All rows of a key are in one partition
// (the input was radix-partitioned on the key hash), so a partition's table holds final groups: flushed to the output
// arrays, never merged.  A table that fills up raises ERR_CAPACITY (the host takes the HBM-table path instead).
constexpr uint64_t GP_EMPTY = ~0ull;

// Shared memory has native 32-bit integer atomics and a native 64-bit compare-and-swap; the CUDA 64-bit add / min / max
// and f64 add compile to lock-and-spin loops that retry ~3x under this access pattern (measured: 17 ATOMS per row for 6
// updates; scripts/microbench/smem_atomics.cu: 1.7 ops/clk/SM against 8.4 for the native 32-bit add).  So:
//   COUNT     one native 32-bit add on the low word (a partition holds < 2^32 rows)
//   SUM i64   two native 32-bit adds: the low words, then the high words plus the carry out of THIS add - the carries of
//             all adds sum to the number of wrap-arounds of the low word whatever the interleaving, so the result is the
//             exact wrapping 64-bit sum
//   SUM f64   the CUDA atomicAdd(double) (a hand-rolled read + CAS loop measured slower: 4.9 ms against 2.9 ms per pass)
//   MIN / MAX read first - most rows do not improve the extreme - then the CUDA atomicMin / atomicMax
// (Tried and dropped, round 2: no accumulator atomics at all - every row chains itself to its slot with one native
// exchange, the aggregate inputs are parked in shared memory and after a barrier the slot's owner thread folds its
// chain with plain loads / stores.  2.74 ms against 2.40 ms for this form on config 5: two CTA barriers per tile, twice
// the shared memory (two CTAs per SM instead of four) and the uneven chain lengths cost more than the atomics.)
QG_D void smem_atomic_acc(uint8_t kind, uint64_t* p, uint64_t v) {
  switch (kind) {
    case ACC_ADD_F64: atomicAdd((double*)p, bits_f64(v)); break;
    case ACC_MIN_I64:
      if ((long long)v < *(volatile long long*)p) atomicMin((long long*)p, (long long)v);
      break;
    case ACC_MAX_I64:
      if ((long long)v > *(volatile long long*)p) atomicMax((long long*)p, (long long)v);
      break;
    case ACC_COUNT: atomicAdd((unsigned int*)p, (unsigned int)v); break;
    default: {  // ACC_ADD_I64
      unsigned int* w = (unsigned int*)p;
      const unsigned int lo = (unsigned int)v;
      const unsigned int old = atomicAdd(w, lo);
      const unsigned int hi = (unsigned int)(v >> 32) + ((unsigned int)(old + lo) < lo ? 1u : 0u);
      if (hi) atomicAdd(w + 1, hi);
      break;
    }
  }
}

template <int R, class Prog>
struct SinkGroupPart : SinkNone {
  unsigned long long* keys;  // [slots + 1]
  uint64_t* acc;             // [slots + 1][n_acc]
  uint32_t* has_marker;      // the marker-valued key occurred in this partition
  uint32_t slots;
  int n_acc, key_slot_rt;
  uint32_t slot[R];
  uint32_t* err;

  QG_D int kslot() const { return Prog::STATIC ? Prog::key_slot(0) : key_slot_rt; }

  template <class Loader>
  QG_D void group(Loader& ld, uint32_t pass) {
    Lanes<R> kw;
    ld.template load<R>(kslot(), kw);
    const uint32_t mask = slots - 1;
    // The R rows of a thread probe TOGETHER: every round issues one (independent) shared-memory load per pending row, so
    // a thread pays max_r(chain length) dependent round trips instead of their sum - and a warp runs until its longest
    // chain is done, which at 30 % load is 4-5 slots.  (Measured on config 5: the partition pass is 30 % faster at 15 %
    // load than at 30 % with the rows probing one after the other, scripts/gpu_r2s.sh.)
    uint32_t pend = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      slot[r] = 0;
      if (!((pass >> r) & 1u)) continue;
      if (kw.v[r] == GP_EMPTY) {
        *has_marker = 1u;
        slot[r] = slots;
        continue;
      }
      // (bits 20.. of the hash: the partitioning consumed its top bits)
      slot[r] = (uint32_t)(mix64(0x9e3779b97f4a7c15ull ^ kw.v[r]) >> 20) & mask;
      pend |= 1u << r;
    }
    for (uint32_t probes = 0; pend != 0; ++probes) {
      if (probes > mask) {
        *err |= ERR_CAPACITY;
        break;
      }
      // most rows find their key already in place (a key has many rows): a plain load first, the CAS only on an
      // empty slot (shared-memory CAS runs at half the rate of a load, scripts/microbench/smem_atomics.cu)
      unsigned long long old[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if ((pend >> r) & 1u) old[r] = *(volatile unsigned long long*)(keys + slot[r]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!((pend >> r) & 1u)) continue;
        const unsigned long long k = kw.v[r];
        if (old[r] == GP_EMPTY) old[r] = atomicCAS(keys + slot[r], GP_EMPTY, k);
        if (old[r] == GP_EMPTY || old[r] == k) pend &= ~(1u << r);
        else slot[r] = (slot[r] + 1) & mask;
      }
    }
    __syncwarp();  // (see the tile loop: fragments must not carry on separately)
  }
  // (Tried and dropped: issuing the accumulator loads / CAS of a thread's R rows together, like the probes above: 112
  // registers instead of 64, two CTAs per SM, 3.1-3.2 ms against 1.8-2.1 ms.)
  template <int RR>
  QG_D void agg(const Insn& in, const Lanes<RR>& a, uint32_t pass) {
    const uint32_t m = a.ok & pass;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if ((m >> r) & 1u) smem_atomic_acc(in.mode, acc + (size_t)slot[r] * n_acc + in.arg, in.mode == ACC_COUNT ? 1ull : agg_value(in, a.v[r]));
    __syncwarp();
  }
};

// ---------------- SK_DIAG: which checked operation failed first? ----------------------------------
// order key: section (6 bits) | batch (24) | instruction (8) | row inside the batch (26)
struct SinkDiag : SinkNone {
  const ScanParams* P;
  int64_t base;
  int pc;
  QG_D void fault(const Insn& in, int r, uint64_t lhs, uint64_t rhs, uint32_t e) {
    const int64_t row = base + (int64_t)r * SCAN_THREADS + threadIdx.x;
    int lo = 0, hi = P->diag_n_batches;  // the batch that holds `row`: offsets[lo] <= row < offsets[lo + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (P->diag_batch_offsets[mid] <= row) lo = mid;
      else hi = mid;
    }
    const uint64_t in_batch = (uint64_t)(row - P->diag_batch_offsets[lo]);
    const unsigned long long key = ((unsigned long long)(P->diag_section_of_pc[pc] & 63u) << 58) | ((unsigned long long)(lo & 0xffffff) << 34) |
                                   ((unsigned long long)(pc & 255) << 26) | (unsigned long long)(in_batch < 0x3ffffffull ? in_batch : 0x3ffffffull);
    if (P->diag_mode == 0) {
      atomicMin(P->diag_key, key);
    } else if (key == *P->diag_key) {
      DiagRecord rec;
      rec.lhs = lhs;
      rec.rhs = rhs;
      rec.op = in.op;
      rec.ty = in.ty;
      rec.err = e;
      rec.pad = 0;
      *P->diag_rec = rec;
    }
  }
};

// ---------------- SK_STORE: write expression results as columns --------------------------------
struct SinkStore : SinkNone {
  const ScanParams* P;
  int64_t base;
  template <int R>
  QG_D void store(const Insn& in, const Lanes<R>& a, uint32_t pass) {
    constexpr int T = SCAN_THREADS;
    const int o = in.arg;
    const uint8_t ty = in.ty;
    const int64_t row0 = base + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + (int64_t)r * T;
      const bool in_b = (pass >> r) & 1u;
      if (ty == DT_I64 || ty == DT_F64) {
        if (in_b) st_stream_u64((uint64_t*)P->out_data[o] + row, a.v[r]);
      } else if (ty == DT_I32 || ty == DT_F32) {
        if (in_b) st_stream_u32((uint32_t*)P->out_data[o] + row, (uint32_t)a.v[r]);
      } else if (ty == DT_BOOL) {
        const uint32_t w = __ballot_sync(0xffffffffu, in_b && (a.v[r] & 1u));
        if (lane_id() == 0 && (row - lane_id()) < P->n_rows) ((uint32_t*)P->out_data[o])[row >> 5] = w;
      }
      if (P->out_valid[o] != nullptr) {
        const uint32_t w = __ballot_sync(0xffffffffu, in_b && ((a.ok >> r) & 1u));
        if (lane_id() == 0 && row < P->n_rows) P->out_valid[o][row >> 5] = w;
      }
    }
  }
};

// ---------------- interpreter program policy ------------------------------------------------------
struct InterpProg {
  static constexpr bool STATIC = false;
  static constexpr bool EAGER = false;
  static constexpr int N_ACC = 0;
  static constexpr int N_COLS = 0;
  static constexpr int N_KEYS = 0;
  static constexpr int STACK = 0;
  static QG_D int key_slot(int) { return 0; }
  static QG_D uint8_t acc_kind(int) { return 0; }
  const Insn* code;  // in shared memory
  int n_pred, n_insn;
  static QG_D int col_dtype(int slot, const ColRef* cols) { return cols[slot].dtype; }
  static QG_D bool col_has_valid(int slot, const ColRef* cols) { return cols[slot].valid != nullptr; }
  template <int R, class L, class S, class K>
  QG_D void run_pred(Machine<R>& m, L& ld, S& st, K& sk) const { run_program<R>(code, n_pred, m, ld, st, sk); }
  template <int R, class L, class S, class K>
  QG_D void run_rest(Machine<R>& m, L& ld, S& st, K& sk) const { run_program<R>(code + n_pred, n_insn - n_pred, m, ld, st, sk); }
};

// ------------------------------------------------------------------------------------------------
// The kernel body.  Shared memory layout (dynamic): [program][stack][sink state]; a specialised kernel has neither
// program nor stack there, and no accumulators for SK_SCALAR.
// ------------------------------------------------------------------------------------------------
template <int R, int SINK, class Prog>
QG_D void scan_body(const ScanParams& P, Prog prog, unsigned char* smem_raw) {
  constexpr int T = SCAN_THREADS;
  constexpr int NW = SCAN_NW;
  constexpr int64_t TILE = (int64_t)T * R;
  constexpr uint32_t ALL = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
  unsigned char* sp = smem_raw;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  // ---- program + stack ----
  SmemStack sstk;
  RegStack<R, Prog::STACK> rstk;
  rstk.sp = 0;
  if constexpr (!Prog::STATIC) {
    Insn* sprog = (Insn*)sp;
    sp += (size_t)P.n_insn * sizeof(Insn);
    sstk.v = (uint64_t*)sp;
    sp += (size_t)P.stack_depth * R * T * 8;
    sstk.ok = (uint32_t*)sp;
    sp += (size_t)P.stack_depth * T * 4;
    sstk.sp = 0;
    for (int i = tid; i < P.n_insn; i += T) sprog[i] = P.prog[i];
    prog.code = sprog;
    prog.n_pred = P.n_pred;
    prog.n_insn = P.n_insn;
  }
  const int n_acc = Prog::STATIC ? Prog::N_ACC : P.n_acc;

  // ---- sink set-up ----
  [[maybe_unused]] uint64_t* sacc = nullptr;
  [[maybe_unused]] uint32_t* s_cnt = nullptr;
  [[maybe_unused]] SinkScalar<Prog> sk_scalar;
  [[maybe_unused]] SinkGroupSmem<R, Prog> sk_gs;
  [[maybe_unused]] SinkGroupGlobal<R> sk_gg;
  [[maybe_unused]] SinkGroupPart<R, Prog> sk_gp;
  [[maybe_unused]] uint16_t* gp_list = nullptr;   // used slots of the partition being flushed
  [[maybe_unused]] uint32_t* gp_ctl = nullptr;    // [0] used-slot count, [1] marker key seen, [2] output base
  [[maybe_unused]] SinkStore sk_store;
  [[maybe_unused]] SinkDiag sk_diag;
  [[maybe_unused]] SinkNone sk_none;
  [[maybe_unused]] uint64_t* swarp = nullptr;
  [[maybe_unused]] int* s_last = nullptr;
  uint32_t err = 0;

  if constexpr (SINK == SK_MASK) {
    s_cnt = (uint32_t*)sp;
    if (tid == 0) *s_cnt = 0;
  } else if constexpr (SINK == SK_SCALAR) {
    if constexpr (Prog::STATIC) {
#pragma unroll
      for (int s = 0; s < Prog::N_ACC; ++s) sk_scalar.racc[s] = acc_identity(Prog::acc_kind(s));
    } else {
      sacc = (uint64_t*)sp;
      sp += (size_t)n_acc * T * 8;
      for (int s = 0; s < n_acc; ++s) sacc[s * T + tid] = acc_identity(P.acc_kinds[s]);
      sk_scalar.sacc = sacc;
    }
    swarp = (uint64_t*)sp;
    sp += (size_t)n_acc * NW * 8;
    s_last = (int*)sp;
  } else if constexpr (SINK == SK_GROUP_SMEM) {
    sacc = (uint64_t*)sp;
    sp += (size_t)P.gp * n_acc * T * 8;
    sk_gs.d.keys = (uint64_t*)sp;
    sp += (size_t)P.gp * P.n_keys * 8;
    sk_gs.d.nullmask = (uint32_t*)sp;
    sp += (size_t)P.gp * 4;
    sk_gs.d.row = (uint32_t*)sp;
    sp += (size_t)P.gp * 4;
    sk_gs.d.cnt = (int*)sp;
    sk_gs.d.elect = (int*)sp + 1;
    sk_gs.d.overflow = (int*)sp + 2;
    sp += 16;
    // (the flush epilogue uses the gp words at sp as its group -> table-slot map)
    for (int g = 0; g < P.gp; ++g)
      for (int s = 0; s < n_acc; ++s) sacc[(size_t)(g * n_acc + s) * T + tid] = acc_identity(P.acc_kinds[s]);
    if (tid == 0) {
      *sk_gs.d.cnt = 0;
      *sk_gs.d.elect = 0x7fffffff;
      *sk_gs.d.overflow = 0;
    }
    sk_gs.sacc = sacc;
    sk_gs.n_acc = n_acc;
    sk_gs.n_keys_rt = P.n_keys;
    sk_gs.gp = P.gp;
    sk_gs.key_slots = P.key_slots;
  } else if constexpr (SINK == SK_GROUP_GLOBAL) {
    sk_gg.g = &P.gt;
    sk_gg.err = &err;
  } else if constexpr (SINK == SK_GROUP_PART) {
    const uint32_t S = P.part_slots;
    sk_gp.keys = (unsigned long long*)sp;
    sp += (size_t)(S + 1) * 8;
    sk_gp.acc = (uint64_t*)sp;
    sp += (size_t)(S + 1) * n_acc * 8;
    gp_ctl = (uint32_t*)sp;
    sp += 16;
    gp_list = (uint16_t*)sp;
    sk_gp.has_marker = gp_ctl + 1;
    sk_gp.slots = S;
    sk_gp.n_acc = n_acc;
    sk_gp.key_slot_rt = P.key_slots[0];
    sk_gp.err = &err;
    for (uint32_t i = tid; i <= S; i += T) sk_gp.keys[i] = GP_EMPTY;
    for (uint32_t i = tid; i < (S + 1) * (uint32_t)n_acc; i += T) sk_gp.acc[i] = acc_identity(P.acc_kinds[i % (uint32_t)n_acc]);
    if (tid < 4) gp_ctl[tid] = 0;
  } else if constexpr (SINK == SK_STORE) {
    sk_store.P = &P;
  } else if constexpr (SINK == SK_DIAG) {
    sk_diag.P = &P;
  }
  __syncthreads();

  DevLoader<Prog, R> ld;
  ld.cols = P.cols;
  Machine<R> m;
  m.err = 0;

  bool cta_dead = false;
  [[maybe_unused]] bool eager = false;  // (the first tile gates its loads; see the tile loop)
  [[maybe_unused]] uint32_t n_kept = 0;  // rows of this thread that passed the predicate (feedback for the host)
  if constexpr (SINK == SK_GROUP_PART) {
    // ---- partition-major input: this CTA owns whole partitions; a partition = part_regions consecutive input regions ----
    const uint32_t S = P.part_slots;
    const int64_t part_rows = (int64_t)P.region_cap * P.part_regions;
    const int64_t n_parts = P.n_rows / part_rows;
    for (int64_t part = blockIdx.x; part < n_parts; part += gridDim.x) {
      for (int g = 0; g < P.part_regions; ++g) {
        const int64_t region = part * P.part_regions + g;
        const int64_t rbase = region * (int64_t)P.region_cap;
        const uint32_t cnt = min(P.region_count[(size_t)region * P.region_count_stride], P.region_cap);
        for (uint32_t t0 = 0; t0 < cnt; t0 += (uint32_t)TILE) {
          const int64_t base = rbase + t0;
          if (P.l2_prefetch && tid < P.n_cols && t0 + (uint32_t)TILE < cnt && P.col_width[tid] > 0)
            prefetch_l2_bulk((const char*)P.cols[tid].data + (base + TILE) * P.col_width[tid], (uint32_t)(TILE * P.col_width[tid]));
          uint32_t inb = 0;
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (t0 + (uint32_t)(r * T + tid) < cnt) inb |= 1u << r;
          // Reconverge first.  The probe / CAS / min-max branches below diverge per row and nothing in this loop is a
          // CTA barrier, so without it the fragments of a warp run on through tile after tile on their own until the
          // flush: ncu counted 12 active lanes per executed instruction over the whole tile body - the loads included -
          // i.e. 2.6x the instructions (profiles/r02u_group_part_divergence.md).
          __syncwarp();
          ld.new_tile(base, inb);
          // every referenced column of the tile is requested up front: one HBM round trip per tile instead of one per
          // column (key -> probe loop -> first aggregate input -> ... are dependent phases otherwise); 2.88 -> 2.40 ms
          // on BASELINE config 5 (gpurun_out r2m)
          if constexpr (QGPU_GP_PRELOAD != 0) ld.preload_all();
          machine_reset(m, inb);
          if constexpr (Prog::STATIC) {
            rstk.sp = 0;
            prog.template run_pred<R>(m, ld, rstk, sk_none);
            sk_gp.group(ld, m.pass);
            prog.template run_rest<R>(m, ld, rstk, sk_gp);
          } else {
            sstk.sp = 0;
            prog.template run_pred<R>(m, ld, sstk, sk_none);
            sk_gp.group(ld, m.pass);
            prog.template run_rest<R>(m, ld, sstk, sk_gp);
          }
        }
      }
      __syncthreads();
      // ---- flush: list the used slots, reserve output rows, write (key, accumulator row), reset the used slots ----
      for (uint32_t s0 = 0; s0 < S; s0 += T) {
        const uint32_t sl = s0 + tid;
        const bool used = sl < S && sk_gp.keys[sl] != GP_EMPTY;
        const uint32_t bal = __ballot_sync(0xffffffffu, used);
        uint32_t wbase = 0;
        if (lane_id() == 0 && bal) wbase = atomicAdd(gp_ctl, (uint32_t)__popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (used) gp_list[wbase + __popc(bal & lanemask_lt())] = (uint16_t)sl;
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t ng = gp_ctl[0];
        if (gp_ctl[1]) gp_list[ng++] = (uint16_t)S;
        gp_ctl[0] = ng;
        gp_ctl[2] = ng ? atomicAdd(P.gp_cursor, ng) : 0u;
      }
      __syncthreads();
      const uint32_t ng = gp_ctl[0], gbase = gp_ctl[2];
      if (gbase + ng > P.gp_out_cap || gbase + ng < gbase) {
        err |= ERR_OUTCAP;  // the output arrays are too small: the host retries with room for every row
      } else {
        for (uint32_t i = tid; i < ng; i += T) P.gp_out_keys[gbase + i] = sk_gp.keys[gp_list[i]];
        const uint32_t words = ng * (uint32_t)n_acc;
        uint64_t* orow = P.gp_out_acc + (size_t)gbase * n_acc;
        for (uint32_t w = tid; w < words; w += T) {
          const uint32_t i = w / (uint32_t)n_acc, a = w - i * (uint32_t)n_acc;
          orow[w] = sk_gp.acc[(size_t)gp_list[i] * n_acc + a];
        }
      }
      __syncthreads();
      {
        const uint32_t words = ng * (uint32_t)n_acc;
        for (uint32_t w = tid; w < words; w += T) {
          const uint32_t i = w / (uint32_t)n_acc, a = w - i * (uint32_t)n_acc;
          sk_gp.acc[(size_t)gp_list[i] * n_acc + a] = acc_identity(P.acc_kinds[a]);
        }
        for (uint32_t i = tid; i < ng; i += T) sk_gp.keys[gp_list[i]] = GP_EMPTY;
        if (tid < 3) gp_ctl[tid] = 0;
      }
      __syncthreads();
    }
  }
  for (int64_t tile = blockIdx.x; SINK != SK_GROUP_PART && tile < P.n_tiles && !cta_dead; tile += gridDim.x) {
    const int64_t base = tile * TILE;
    // bulk L2 prefetch of this CTA's next tile, one thread per referenced column
    if (P.l2_prefetch && tid < P.n_cols) {
      const int64_t nb = base + (int64_t)gridDim.x * TILE;
      if (nb + TILE <= P.n_rows && P.col_width[tid] > 0)
        prefetch_l2_bulk((const char*)P.cols[tid].data + nb * P.col_width[tid], (uint32_t)(TILE * P.col_width[tid]));
    }
    uint32_t inb = ALL;
    if (base + TILE > P.n_rows) {
      inb = 0;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (base + (int64_t)r * T + tid < P.n_rows) inb |= 1u << r;
    }
    if (P.region_cap) {  // region-major input: rows past the region's count are gaps (a tile never straddles regions)
      const uint32_t region = (uint32_t)(base / P.region_cap);
      const int64_t live_end = (int64_t)region * P.region_cap + min(P.region_count[(size_t)region * P.region_count_stride], P.region_cap);
      if (base >= live_end) continue;
      if (base + TILE > live_end) {
        uint32_t m = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (base + (int64_t)r * T + tid < live_end) m |= 1u << r;
        inb &= m;
      }
    }
    if constexpr (Prog::STATIC && Prog::EAGER && (SINK == SK_SCALAR || SINK == SK_GROUP_SMEM)) {
      // two copies of the tile's straight-line code: loads gated on the predicate's survivors (Q6: 2 % survive, the
      // other columns' sectors mostly stay in HBM), or every column fetched up front (Q1: 98 % survive).  The previous
      // tile's survivor rate picks the copy, warp by warp.
      auto tile_body = [&](auto eager_tag) {
        ld.new_tile(base, inb);
        if constexpr (decltype(eager_tag)::value) ld.preload_all();
        machine_reset(m, inb);
        rstk.sp = 0;
        if constexpr (SINK == SK_SCALAR) {
          prog.template run_pred<R>(m, ld, rstk, sk_scalar);
          prog.template run_rest<R>(m, ld, rstk, sk_scalar);
        } else {
          prog.template run_pred<R>(m, ld, rstk, sk_none);
          if (!sk_gs.group(ld, m.pass)) cta_dead = true;
          else prog.template run_rest<R>(m, ld, rstk, sk_gs);
        }
      };
      if (eager) tile_body(BoolTag<true>{});
      else tile_body(BoolTag<false>{});
      const uint32_t kept = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m.pass));
      const uint32_t seen = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(inb));
      eager = kept * 2u >= seen;
      n_kept += (uint32_t)__popc(m.pass);
      continue;
    }
    ld.new_tile(base, inb);
    machine_reset(m, inb);
    auto run_all = [&](auto& stk, auto& sink) {
      prog.template run_pred<R>(m, ld, stk, sink);
      prog.template run_rest<R>(m, ld, stk, sink);
    };
    auto with_stack = [&](auto&& f) {
      if constexpr (Prog::STATIC) {
        rstk.sp = 0;
        f(rstk);
      } else {
        sstk.sp = 0;
        f(sstk);
      }
    };

    if constexpr (SINK == SK_MASK) {
      with_stack([&](auto& stk) { run_all(stk, sk_none); });
      uint32_t wsum = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t w = __ballot_sync(0xffffffffu, (m.pass >> r) & 1u);
        if (lane_id() == 0) {
          P.mask[(base + (int64_t)r * T + warp * 32) >> 5] = w;
          wsum += __popc(w);
        }
      }
      if (lane_id() == 0 && wsum) atomicAdd(s_cnt, wsum);
      __syncthreads();
      if (tid == 0) {
        P.tile_counts[tile] = *s_cnt;
        *s_cnt = 0;
      }
      __syncthreads();
    } else if constexpr (SINK == SK_SCALAR) {
      with_stack([&](auto& stk) { run_all(stk, sk_scalar); });
      if constexpr (Prog::STATIC) n_kept += (uint32_t)__popc(m.pass);
    } else if constexpr (SINK == SK_GROUP_SMEM) {
      with_stack([&](auto& stk) {
        prog.template run_pred<R>(m, ld, stk, sk_none);
        if (!sk_gs.group(ld, m.pass)) cta_dead = true;
        else prog.template run_rest<R>(m, ld, stk, sk_gs);
      });
      if constexpr (Prog::STATIC) n_kept += (uint32_t)__popc(m.pass);
    } else if constexpr (SINK == SK_GROUP_GLOBAL) {
      with_stack([&](auto& stk) {
        prog.template run_pred<R>(m, ld, stk, sk_none);
        sk_gg.group(ld, m.pass);
        prog.template run_rest<R>(m, ld, stk, sk_gg);
      });
      // a full table: tell everybody at once and stop (the pass is thrown away and redone with a larger table)
      if (err & ERR_CAPACITY) atomicOr(P.err, (uint32_t)ERR_CAPACITY);
      if (__syncthreads_or((err & ERR_CAPACITY) || (*((volatile uint32_t*)P.err) & ERR_CAPACITY))) cta_dead = true;
    } else if constexpr (SINK == SK_STORE) {
      sk_store.base = base;
      with_stack([&](auto& stk) { run_all(stk, sk_store); });
    } else if constexpr (SINK == SK_DIAG) {
      // interpreter only: one instruction at a time so that the sink knows which one is running
      if constexpr (!Prog::STATIC) {
        sk_diag.base = base;
        sstk.sp = 0;
        for (int pc = 0; pc < prog.n_insn; ++pc) {
          sk_diag.pc = pc;
          exec_insn<R>(prog.code[pc], m, ld, sstk, sk_diag);
        }
      }
    }
  }
  err |= m.err;
  if constexpr (Prog::STATIC && (SINK == SK_SCALAR || SINK == SK_GROUP_SMEM)) {
    // survivors of the predicate, for the host's choice between the gated and the eager form of this kernel next time
    const uint32_t k = __reduce_add_sync(0xffffffffu, n_kept);
    if (lane_id() == 0 && k) atomicAdd(P.err + 2, k);
  }

  // ---- epilogues ----
  if constexpr (SINK == SK_SCALAR) {
    __syncthreads();
    if constexpr (Prog::STATIC) {
#pragma unroll
      for (int s = 0; s < Prog::N_ACC; ++s) {
        const uint64_t v = warp_reduce_acc(Prog::acc_kind(s), sk_scalar.racc[s]);
        if (lane_id() == 0) swarp[s * NW + warp] = v;
      }
    } else {
      for (int s = 0; s < n_acc; ++s) {
        const uint64_t v = warp_reduce_acc(P.acc_kinds[s], sacc[s * T + tid]);
        if (lane_id() == 0) swarp[s * NW + warp] = v;
      }
    }
    __syncthreads();
    if (tid < n_acc) {
      uint64_t v = acc_identity(P.acc_kinds[tid]);
      for (int w = 0; w < NW; ++w) v = acc_combine(P.acc_kinds[tid], v, swarp[tid * NW + w]);
      P.partials[(size_t)blockIdx.x * n_acc + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const uint32_t t = atomicAdd(P.ticket, 1u);
      *s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (*s_last) {
      // last CTA folds the per-CTA partials in a fixed order -> run-to-run deterministic f64 sums
      __threadfence();
      for (int s = warp; s < n_acc; s += NW) {
        uint64_t v = acc_identity(P.acc_kinds[s]);
        for (int b = lane_id(); b < (int)gridDim.x; b += 32)
          v = acc_combine(P.acc_kinds[s], v, __ldcg(P.partials + (size_t)b * n_acc + s));
        v = warp_reduce_acc(P.acc_kinds[s], v);
        if (lane_id() == 0) P.finals[s] = v;
      }
      if (tid == 0) *P.ticket = 0;
    }
  } else if constexpr (SINK == SK_GROUP_SMEM) {
    __syncthreads();
    if (*sk_gs.d.overflow) {
      if (tid == 0) atomicExch(P.overflow, 1u);
    } else {
      // flush this CTA's groups into the HBM table
      uint32_t* s_gslot = (uint32_t*)sp;
      const int cnt = *sk_gs.d.cnt;
      if (tid < cnt) s_gslot[tid] = find_or_insert(P.gt, (int64_t)sk_gs.d.row[tid], err);
      __syncthreads();
      const int pairs = cnt * n_acc;
      for (int pr = warp; pr < pairs; pr += NW) {
        const int s = pr % n_acc;
        const int g = pr / n_acc;
        uint64_t v = acc_identity(P.acc_kinds[s]);
        for (int t = lane_id(); t < T; t += 32) v = acc_combine(P.acc_kinds[s], v, sacc[(size_t)pr * T + t]);
        v = warp_reduce_acc(P.acc_kinds[s], v);
        if (lane_id() == 0) atomic_acc(P.acc_kinds[s], P.gt.acc + (size_t)s_gslot[g] * n_acc + s, v);
      }
    }
  }
  if (err) atomicOr(P.err, err);
}

}  // namespace qgpu
