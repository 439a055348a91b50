// scan_kernels.cu — the fused column-scan kernels of the hot path (sm_100a).
//
//   k_scan<R, SK_MASK>          predicate -> selection bitmap + per-tile survivor counts   (Filter, executor.rs:131-155)
//   k_scan<R, SK_SCALAR>        predicate + ungrouped aggregates in one pass               (HashAggregate scalar path, :167-187)
//   k_scan<R, SK_GROUP_SMEM>    predicate + low-cardinality GROUP BY, per-thread private accumulators in shared memory
//   k_scan<R, SK_GROUP_GLOBAL>  predicate + GROUP BY through an open-addressing table in HBM
//   k_scan<R, SK_STORE>         expression materialisation (Projection, :93-129)
//
// Execution model: persistent grid (sm_count x CTAs/SM), 256 threads, each thread evaluates R rows per
// tile; a tile is 256*R consecutive rows so every warp-level ballot is one aligned 32-row bitmap word.
// Column values are read exactly once with streaming loads; the next tile of every referenced column is
// bulk-prefetched into L2 while the current one is evaluated.  Everything is HBM-bound integer / f64 work:
// no tensor cores.
#include "device_utils.cuh"
#include "scan_kernels.cuh"

namespace qgpu {

namespace {

constexpr int T = SCAN_THREADS;
constexpr int NW = T / WARP;

// ---------------- loader: row(r) = base + r*T + tid -------------------------------------------
struct DevLoader {
  const ColRef* cols;
  int64_t base;
  int64_t n;
  uint32_t inb;

  template <int R>
  QG_D void load(int slot, Lanes<R>& x) const {
    const ColRef c = cols[slot];
    const int64_t row0 = base + threadIdx.x;
    uint32_t ok = inb;
    switch (c.dtype) {
      case DT_I64:
      case DT_F64: {
        const uint64_t* p = (const uint64_t*)c.data + row0;
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = ((inb >> r) & 1u) ? ld_stream_u64(p + (int64_t)r * T) : 0ull;
        break;
      }
      case DT_I32: {
        const uint32_t* p = (const uint32_t*)c.data + row0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          x.v[r] = ((inb >> r) & 1u) ? (uint64_t)(int64_t)(int32_t)ld_stream_u32(p + (int64_t)r * T) : 0ull;
        break;
      }
      case DT_F32: {
        const uint32_t* p = (const uint32_t*)c.data + row0;
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = ((inb >> r) & 1u) ? (uint64_t)ld_stream_u32(p + (int64_t)r * T) : 0ull;
        break;
      }
      case DT_BOOL: {
        const uint32_t* p = (const uint32_t*)c.data;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int64_t row = row0 + (int64_t)r * T;
          x.v[r] = ((inb >> r) & 1u) ? ((__ldg(p + (row >> 5)) >> (row & 31)) & 1u) : 0ull;
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
    if (c.valid != nullptr) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if ((inb >> r) & 1u) {
          const int64_t row = row0 + (int64_t)r * T;
          const uint32_t w = __ldg(c.valid + (row >> 5));
          if (!((w >> (row & 31)) & 1u)) ok &= ~(1u << r);
        }
      }
    }
    x.ok = ok;
  }
};

// ---------------- spill stack in shared memory -------------------------------------------------
struct DevStack {
  uint64_t* v;   // [depth][R][T]
  uint32_t* ok;  // [depth][T]
  int sp;
  template <int R>
  QG_D void push(const Lanes<R>& a) {
#pragma unroll
    for (int r = 0; r < R; ++r) v[(sp * R + r) * T + threadIdx.x] = a.v[r];
    ok[sp * T + threadIdx.x] = a.ok;
    ++sp;
  }
  template <int R>
  QG_D void pop(Lanes<R>& a) {
    --sp;
#pragma unroll
    for (int r = 0; r < R; ++r) a.v[r] = v[(sp * R + r) * T + threadIdx.x];
    a.ok = ok[sp * T + threadIdx.x];
  }
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
  template <int R> QG_D void agg(const Insn&, const Lanes<R>&, uint32_t) {}
  template <int R> QG_D void store(const Insn&, const Lanes<R>&, uint32_t) {}
  template <int R> QG_D void key(const Insn&, const Lanes<R>&, uint32_t) {}
};

// ---------------- SK_SCALAR: per-thread private accumulators in shared memory ----------------
struct SinkScalar : SinkNone {
  uint64_t* sacc;  // [n_acc][T]
  template <int R>
  QG_D void agg(const Insn& in, const Lanes<R>& a, uint32_t pass) {
    const uint32_t m = a.ok & pass;
    uint64_t* p = sacc + in.arg * T + threadIdx.x;
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
};

// ---------------- key hashing / equality on table rows (group-by in HBM, also the T1 flush) ---
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

QG_D uint32_t find_or_insert(const GroupTable& g, int64_t row, uint32_t& err) {
  const uint32_t mask = g.cap - 1;
  uint32_t s = (uint32_t)(hash_row(g, row) >> 20) & mask;
  for (uint32_t probes = 0; probes <= mask; ++probes) {
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
  QG_D void group(const DevLoader& ld, uint32_t pass) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      slot[r] = 0;
      if ((pass >> r) & 1u) slot[r] = find_or_insert(*g, ld.base + (int64_t)r * T + threadIdx.x, *err);
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

template <int R>
struct SinkGroupSmem : SinkNone {
  GroupDir d;
  uint64_t* sacc;  // [(g*n_acc + s)][T]
  int n_acc, n_keys, gp;
  const int32_t* key_slots;
  int gid[R];

  QG_D uint32_t lookup(const DevLoader& ld, uint32_t want) {
    // returns the mask of wanted rows still without a group
    uint32_t cand[R];
#pragma unroll
    for (int r = 0; r < R; ++r) cand[r] = 0xffffffffu;
    const int cnt = *((volatile int*)d.cnt);
    for (int k = 0; k < n_keys; ++k) {
      Lanes<R> kw;
      ld.template load<R>(key_slots[k], kw);
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
  QG_D bool group(const DevLoader& ld, uint32_t pass) {
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
          for (int k = 0; k < n_keys; ++k) {
            Lanes<R> kw;
            ld.template load<R>(key_slots[k], kw);
            uint64_t val = 0;
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (r == r0) val = kw.v[r];
            if (!((kw.ok >> r0) & 1u)) nm |= 1u << k;
            d.keys[g * n_keys + k] = val;
          }
          d.nullmask[g] = nm;
          d.row[g] = (uint32_t)(ld.base + (int64_t)r0 * T + threadIdx.x);
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

  template <int RR>
  QG_D void agg(const Insn& in, const Lanes<RR>& a, uint32_t pass) {
    const uint32_t m = a.ok & pass;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((m >> r) & 1u) {
        uint64_t* p = sacc + (size_t)(gid[r] * n_acc + in.arg) * T + threadIdx.x;
        *p = acc_combine(in.mode, *p, in.mode == ACC_COUNT ? 1ull : agg_value(in, a.v[r]));
      }
    }
  }
};

// ---------------- SK_STORE: write expression results as columns --------------------------------
struct SinkStore : SinkNone {
  const ScanParams* P;
  int64_t base;
  template <int R>
  QG_D void store(const Insn& in, const Lanes<R>& a, uint32_t pass) {
    const int o = in.arg;
    const uint8_t ty = P->out_types[o];
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

// ------------------------------------------------------------------------------------------------
template <int R, int SINK>
__global__ void __launch_bounds__(SCAN_THREADS, 2) k_scan(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* sp = smem_raw;
  Insn* sprog = (Insn*)sp;
  sp += (size_t)P.n_insn * sizeof(Insn);
  DevStack stk;
  stk.v = (uint64_t*)sp;
  sp += (size_t)P.stack_depth * R * T * 8;
  stk.ok = (uint32_t*)sp;
  sp += (size_t)P.stack_depth * T * 4;
  stk.sp = 0;

  for (int i = threadIdx.x; i < P.n_insn; i += T) sprog[i] = P.prog[i];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  uint32_t err = 0;
  constexpr int64_t TILE = (int64_t)T * R;

  // ---- sink set-up ----
  [[maybe_unused]] uint64_t* sacc = nullptr;
  [[maybe_unused]] uint32_t* s_cnt = nullptr;
  [[maybe_unused]] SinkScalar sk_scalar;
  [[maybe_unused]] SinkGroupSmem<R> sk_gs;
  [[maybe_unused]] SinkGroupGlobal<R> sk_gg;
  [[maybe_unused]] SinkStore sk_store;
  [[maybe_unused]] SinkNone sk_none;
  [[maybe_unused]] uint64_t* swarp = nullptr;
  [[maybe_unused]] int* s_last = nullptr;

  if constexpr (SINK == SK_MASK) {
    s_cnt = (uint32_t*)sp;
    if (tid == 0) *s_cnt = 0;
  } else if constexpr (SINK == SK_SCALAR) {
    sacc = (uint64_t*)sp;
    sp += (size_t)P.n_acc * T * 8;
    swarp = (uint64_t*)sp;
    sp += (size_t)P.n_acc * NW * 8;
    s_last = (int*)sp;
    for (int s = 0; s < P.n_acc; ++s) sacc[s * T + tid] = acc_identity(P.acc_kinds[s]);
    sk_scalar.sacc = sacc;
  } else if constexpr (SINK == SK_GROUP_SMEM) {
    sacc = (uint64_t*)sp;
    sp += (size_t)P.gp * P.n_acc * T * 8;
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
      for (int s = 0; s < P.n_acc; ++s) sacc[(size_t)(g * P.n_acc + s) * T + tid] = acc_identity(P.acc_kinds[s]);
    if (tid == 0) {
      *sk_gs.d.cnt = 0;
      *sk_gs.d.elect = 0x7fffffff;
      *sk_gs.d.overflow = 0;
    }
    sk_gs.sacc = sacc;
    sk_gs.n_acc = P.n_acc;
    sk_gs.n_keys = P.n_keys;
    sk_gs.gp = P.gp;
    sk_gs.key_slots = P.key_slots;
  } else if constexpr (SINK == SK_GROUP_GLOBAL) {
    sk_gg.g = &P.gt;
    sk_gg.err = &err;
  } else if constexpr (SINK == SK_STORE) {
    sk_store.P = &P;
  }
  __syncthreads();

  DevLoader ld;
  ld.cols = P.cols;
  ld.n = P.n_rows;

  bool cta_dead = false;
  for (int64_t tile = blockIdx.x; tile < P.n_tiles && !cta_dead; tile += gridDim.x) {
    const int64_t base = tile * TILE;
    // bulk L2 prefetch of this CTA's next tile, one thread per referenced column
    if (P.l2_prefetch && tid < P.n_cols) {
      const int64_t nb = base + (int64_t)gridDim.x * TILE;
      if (nb + TILE <= P.n_rows && P.col_width[tid] > 0)
        prefetch_l2_bulk((const char*)P.cols[tid].data + nb * P.col_width[tid], (uint32_t)(TILE * P.col_width[tid]));
    }
    uint32_t inb = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (base + (int64_t)r * T + tid < P.n_rows) inb |= 1u << r;
    ld.base = base;
    ld.inb = inb;
    uint32_t pass = inb;
    stk.sp = 0;

    if constexpr (SINK == SK_MASK) {
      run_program<R>(sprog, P.n_insn, ld, stk, sk_none, pass, err);
      uint32_t wsum = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t w = __ballot_sync(0xffffffffu, (pass >> r) & 1u);
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
      run_program<R>(sprog, P.n_insn, ld, stk, sk_scalar, pass, err);
    } else if constexpr (SINK == SK_GROUP_SMEM) {
      run_program<R>(sprog, P.n_pred, ld, stk, sk_none, pass, err);
      if (!sk_gs.group(ld, pass)) {
        cta_dead = true;
      } else {
        run_program<R>(sprog + P.n_pred, P.n_insn - P.n_pred, ld, stk, sk_gs, pass, err);
      }
    } else if constexpr (SINK == SK_GROUP_GLOBAL) {
      run_program<R>(sprog, P.n_pred, ld, stk, sk_none, pass, err);
      sk_gg.group(ld, pass);
      run_program<R>(sprog + P.n_pred, P.n_insn - P.n_pred, ld, stk, sk_gg, pass, err);
    } else if constexpr (SINK == SK_STORE) {
      sk_store.base = base;
      run_program<R>(sprog, P.n_insn, ld, stk, sk_store, pass, err);
    }
  }

  // ---- epilogues ----
  if constexpr (SINK == SK_SCALAR) {
    __syncthreads();
    for (int s = 0; s < P.n_acc; ++s) {
      const uint64_t v = warp_reduce_acc(P.acc_kinds[s], sacc[s * T + tid]);
      if (lane_id() == 0) swarp[s * NW + warp] = v;
    }
    __syncthreads();
    if (tid < P.n_acc) {
      uint64_t v = acc_identity(P.acc_kinds[tid]);
      for (int w = 0; w < NW; ++w) v = acc_combine(P.acc_kinds[tid], v, swarp[tid * NW + w]);
      P.partials[(size_t)blockIdx.x * P.n_acc + tid] = v;
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
      for (int s = warp; s < P.n_acc; s += NW) {
        uint64_t v = acc_identity(P.acc_kinds[s]);
        for (int b = lane_id(); b < (int)gridDim.x; b += 32)
          v = acc_combine(P.acc_kinds[s], v, __ldcg(P.partials + (size_t)b * P.n_acc + s));
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
      const int pairs = cnt * P.n_acc;
      for (int pr = warp; pr < pairs; pr += NW) {
        const int s = pr % P.n_acc;
        const int g = pr / P.n_acc;
        uint64_t v = acc_identity(P.acc_kinds[s]);
        for (int t = lane_id(); t < T; t += 32) v = acc_combine(P.acc_kinds[s], v, sacc[(size_t)pr * T + t]);
        v = warp_reduce_acc(P.acc_kinds[s], v);
        if (lane_id() == 0) atomic_acc(P.acc_kinds[s], P.gt.acc + (size_t)s_gslot[g] * P.n_acc + s, v);
      }
    }
  }
  if (err) atomicOr(P.err, err);
}

// ---------------- group table helpers ------------------------------------------------------------
__global__ void k_group_init(uint64_t* acc, size_t n, int n_acc, const uint64_t* ident) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[i] = ident[i % n_acc];
}

__global__ void k_group_collect(const uint32_t* rep, uint32_t cap, uint32_t* out_rep, uint32_t* out_slot, uint32_t* count) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x) {
    const uint32_t r = rep[s];
    if (r != 0) {
      const uint32_t i = atomicAdd(count, 1u);
      out_rep[i] = r - 1;
      out_slot[i] = s;
    }
  }
}

struct FinalizeParams {
  const uint64_t* acc;
  const uint32_t* slots;
  int n_acc;
  int n_desc;
  int64_t n_groups;
  AggFinalizeDesc d[MAX_ACC];
};

// evaluate_aggregate (operators.rs:745-848) on the reduced accumulators of each group
__global__ void k_agg_finalize(const __grid_constant__ FinalizeParams F) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool inb = i < F.n_groups;
  const uint64_t* a = F.acc + (size_t)(inb ? (F.slots ? F.slots[i] : (uint32_t)i) : 0) * F.n_acc;
  for (int j = 0; j < F.n_desc; ++j) {
    const AggFinalizeDesc& d = F.d[j];
    bool valid = false;
    uint64_t out = 0;
    if (inb) {
      const uint64_t cnt = a[d.cnt_slot];
      switch (d.func) {
        case QGPU_COUNT: out = cnt; valid = true; break;
        case QGPU_SUM:
          valid = cnt > 0;
          out = a[d.val_slot];
          if (d.in_type == DT_I32) out = (uint64_t)(int64_t)(int32_t)(uint32_t)out;  // i32 sum wraps, then widens (:755-757)
          // Float32: arrow sums in f32 and the reference widens the f32 result (:761-763); we accumulate in f64 and
          // round once to f32, so the value is an f32 like the reference's (within f32 accumulation error of it)
          if (d.in_type == DT_F32) out = f64_bits((double)(float)bits_f64(out));
          break;
        case QGPU_AVG: {
          valid = cnt > 0;
          double s;
          if (d.in_type == DT_I64) s = (double)(int64_t)a[d.val_slot];
          else if (d.in_type == DT_I32) s = (double)(int32_t)(uint32_t)a[d.val_slot];
          else if (d.in_type == DT_F32) s = (double)(float)bits_f64(a[d.val_slot]);  // f32 sum `as f64` (:797-798)
          else s = bits_f64(a[d.val_slot]);
          out = f64_bits(valid ? s / (double)cnt : 0.0);
          break;
        }
        default: {  // MIN / MAX: decode the order key back to the input type
          valid = cnt > 0;
          const int64_t k = (int64_t)a[d.val_slot];
          if (d.in_type == DT_F64) out = f64_from_order_key(k);
          else if (d.in_type == DT_F32) out = f32_from_order_key(k);
          else out = (uint64_t)k;
          break;
        }
      }
      if (!valid) out = 0;
      if (d.out_type == DT_I64 || d.out_type == DT_F64) ((uint64_t*)d.out_data)[i] = out;
      else ((uint32_t*)d.out_data)[i] = (uint32_t)out;
    }
    const uint32_t w = __ballot_sync(0xffffffffu, inb && valid);
    if ((threadIdx.x & 31) == 0 && inb) d.out_valid[i >> 5] = w;
  }
}

template <int R, int SINK>
void launch_one(Ctx* c, ScanParams& p, const char* name) {
  const size_t smem = scan_smem_bytes(SINK, R, p);
  auto kern = k_scan<R, SINK>;
  if (smem > 48 * 1024) QG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SCAN_THREADS, smem));
  if (per_sm < 1) throw_error(QGPU_ERR_CUDA, "scan kernel %s does not fit on an SM (smem %zu)", name, smem);
  int64_t want = c->opt("scan.ctas_per_sm", 0);
  if (want > 0 && want < per_sm) per_sm = (int)want;
  int64_t grid = (int64_t)c->sm_count * per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  if (grid < 1) grid = 1;
  ProfScope ps(c, name);
  kern<<<(unsigned)grid, SCAN_THREADS, smem, c->stream>>>(p);
  QG_CUDA(cudaGetLastError());
}

}  // namespace

size_t scan_smem_bytes(int sink, int R, const ScanParams& p) {
  size_t b = (size_t)p.n_insn * sizeof(Insn);
  b += (size_t)p.stack_depth * R * T * 8 + (size_t)p.stack_depth * T * 4;
  switch (sink) {
    case SK_MASK: b += 16; break;
    case SK_SCALAR: b += (size_t)p.n_acc * T * 8 + (size_t)p.n_acc * NW * 8 + 16; break;
    case SK_GROUP_SMEM:
      b += (size_t)p.gp * p.n_acc * T * 8 + (size_t)p.gp * p.n_keys * 8 + (size_t)p.gp * 8 + 16 + (size_t)p.gp * 4;
      break;
    default: break;
  }
  return (b + 15) & ~(size_t)15;
}

void launch_scan(Ctx* c, int sink, int R, ScanParams& p, const char* name) {
  const int64_t tile = (int64_t)SCAN_THREADS * R;
  p.n_tiles = (p.n_rows + tile - 1) / tile;
  if (p.n_tiles == 0) p.n_tiles = 1;
#define QG_CASE(RR, SK) \
  if (R == RR && sink == SK) return launch_one<RR, SK>(c, p, name);
  QG_CASE(4, SK_MASK) QG_CASE(8, SK_MASK)
  QG_CASE(4, SK_SCALAR) QG_CASE(8, SK_SCALAR)
  QG_CASE(4, SK_GROUP_SMEM) QG_CASE(8, SK_GROUP_SMEM)
  QG_CASE(4, SK_GROUP_GLOBAL) QG_CASE(8, SK_GROUP_GLOBAL)
  QG_CASE(4, SK_STORE) QG_CASE(8, SK_STORE)
#undef QG_CASE
  throw_error(QGPU_ERR_INVALID, "no scan kernel for rows_per_thread=%d sink=%d", R, sink);
}

void launch_group_init(Ctx* c, const GroupTable& gt, const uint8_t* kinds) {
  uint64_t ident[MAX_ACC];
  for (int i = 0; i < gt.n_acc; ++i) ident[i] = acc_identity(kinds[i]);
  DevBuf d_ident(c, sizeof ident);
  void* h = c->scratch(sizeof ident);
  memcpy(h, ident, sizeof ident);
  QG_CUDA(cudaMemcpyAsync(d_ident.p, h, sizeof ident, cudaMemcpyHostToDevice, c->stream));
  QG_CUDA(cudaMemsetAsync(gt.rep, 0, (size_t)gt.cap * 4, c->stream));
  const size_t n = (size_t)gt.cap * gt.n_acc;
  int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 8);
  if (grid < 1) grid = 1;
  ProfScope ps(c, "k_group_init");
  k_group_init<<<grid, 256, 0, c->stream>>>(gt.acc, n, gt.n_acc, d_ident.as<uint64_t>());
  QG_CUDA(cudaGetLastError());
  // the scratch buffer is reused by later calls: make sure the copy has been issued from it
  QG_CUDA(cudaStreamSynchronize(c->stream));
}

void launch_group_collect(Ctx* c, const GroupTable& gt, uint32_t* out_rep, uint32_t* out_slot, uint32_t* d_count) {
  QG_CUDA(cudaMemsetAsync(d_count, 0, 4, c->stream));
  int grid = (int)std::min<size_t>(((size_t)gt.cap + 255) / 256, (size_t)c->sm_count * 8);
  ProfScope ps(c, "k_group_collect");
  k_group_collect<<<grid, 256, 0, c->stream>>>(gt.rep, gt.cap, out_rep, out_slot, d_count);
  QG_CUDA(cudaGetLastError());
}

void launch_agg_finalize(Ctx* c, const uint64_t* acc, const uint32_t* slots, int n_acc, int64_t n_groups,
                         const AggFinalizeDesc* descs, int n_desc) {
  if (n_desc > MAX_ACC) throw_error(QGPU_ERR_UNSUPPORTED, "too many aggregates (%d)", n_desc);
  FinalizeParams F;
  F.acc = acc;
  F.slots = slots;
  F.n_acc = n_acc;
  F.n_desc = n_desc;
  F.n_groups = n_groups;
  for (int i = 0; i < n_desc; ++i) F.d[i] = descs[i];
  const int64_t blocks = (n_groups + 255) / 256;
  ProfScope ps(c, "k_agg_finalize");
  k_agg_finalize<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, c->stream>>>(F);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
