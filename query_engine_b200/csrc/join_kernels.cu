// join_kernels.cu — hash join build / probe (INNER, LEFT) on one equality key.
//
// Replaces the body of QueryExecutor::execute_join (crates/query-executor/src/executor.rs:343-399), whose
// shipped form is a Cartesian product that ignores `on`; the semantics implemented are the documented
// equi-join (docs/SQL_REFERENCE.md:215-229): NULL keys never match, LEFT keeps unmatched
// probe rows with NULL right columns, output rows come in probe (left) row order of the table the kernels see.
//
// This is the GENERIC path (any key type, duplicate build keys, validity bitmaps, Utf8 payloads): one flat table in HBM
// (join_kernels.cuh) of 32-byte buckets, linear probing over buckets, a count pass and an emit pass.  Int64-key joins
// with a duplicate-free build side take the radix-partitioned path instead (pjoin_kernels.cu).
//
// Row layout of the probe-side kernels: a tile is 1024 consecutive rows, thread t owns rows base + 4t .. 4t+3, so the
// per-row side arrays (first match, output positions) and the fused emit's column reads/writes are 16/32-byte vectors
// and the output keeps probe-row order.
#include "device_utils.cuh"
#include "join_kernels.cuh"
#include "scan_body.cuh"

namespace qgpu {
namespace {

struct TableRef {
  void* slots;
  uint32_t bmask;    // buckets - 1
  int64_t kmin, kmax;
};

QG_D uint32_t bucket_of(const TableRef& t, int64_t key) { return (uint32_t)(mix64((uint64_t)key) >> 20) & t.bmask; }

// ---------------- slot formats ----------------------------------------------------------------------
template <bool COMPACT>
struct Fmt;

template <>
struct Fmt<true> {  // 4 x 8-byte slots per bucket
  static constexpr int SLOTS = 4;
  QG_D static bool in_range(const TableRef& t, int64_t k) { return k >= t.kmin && k <= t.kmax; }
  QG_D static uint64_t* bucket(const TableRef& t, uint32_t b) {
    return (uint64_t*)t.slots + (size_t)b * SLOTS;
  }
  QG_D static bool insert(const TableRef& t, uint32_t b, int64_t k, uint32_t row1) {
    uint64_t* p = bucket(t, b);
    const uint64_t v = ((uint64_t)(uint32_t)(k - t.kmin) << 32) | row1;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if (atomicCAS((unsigned long long*)(p + s), 0ull, (unsigned long long)v) == 0ull) return true;
    return false;
  }
  // scans one bucket: adds matches, remembers the first; returns true when the chain ends here (an empty slot)
  QG_D static bool scan(const TableRef& t, uint32_t b, int64_t k, uint32_t& cnt, uint32_t& first) {
    const uint64_t* p = bucket(t, b);
    uint64_t w[4];
    asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p));
    asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(w[2]), "=l"(w[3]) : "l"(p + 2));
    const uint32_t ko = (uint32_t)(k - t.kmin);
    bool end = false;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (w[s] == 0) end = true;
      else if ((uint32_t)(w[s] >> 32) == ko) {
        if (cnt == 0) first = (uint32_t)w[s];
        ++cnt;
      }
    }
    return end;
  }
  template <class F>
  QG_D static bool for_matches(const TableRef& t, uint32_t b, int64_t k, F f) {
    const uint64_t* p = bucket(t, b);
    const uint32_t ko = (uint32_t)(k - t.kmin);
    bool end = false;
    for (int s = 0; s < SLOTS; ++s) {
      const uint64_t w = __ldg((const unsigned long long*)(p + s));
      if (w == 0) end = true;
      else if ((uint32_t)(w >> 32) == ko) f((uint32_t)w);
    }
    return end;
  }
};

struct __align__(16) WideSlot {
  int64_t key;
  uint32_t row1;
  uint32_t pad;
};

template <>
struct Fmt<false> {  // 2 x 16-byte slots per bucket
  static constexpr int SLOTS = 2;
  QG_D static bool in_range(const TableRef&, int64_t) { return true; }
  QG_D static WideSlot* bucket(const TableRef& t, uint32_t b) {
    return (WideSlot*)t.slots + (size_t)b * SLOTS;
  }
  QG_D static bool insert(const TableRef& t, uint32_t b, int64_t k, uint32_t row1) {
    WideSlot* p = bucket(t, b);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if (atomicCAS(&p[s].row1, 0u, row1) == 0u) {
        p[s].key = k;  // probed only after the build kernel has finished
        return true;
      }
    return false;
  }
  QG_D static bool scan(const TableRef& t, uint32_t b, int64_t k, uint32_t& cnt, uint32_t& first) {
    const WideSlot* p = bucket(t, b);
    uint64_t a[2], r[2];
    asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a[0]), "=l"(r[0]) : "l"(p));
    asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a[1]), "=l"(r[1]) : "l"(p + 1));
    bool end = false;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if ((uint32_t)r[s] == 0) end = true;
      else if ((int64_t)a[s] == k) {
        if (cnt == 0) first = (uint32_t)r[s];
        ++cnt;
      }
    }
    return end;
  }
  template <class F>
  QG_D static bool for_matches(const TableRef& t, uint32_t b, int64_t k, F f) {
    const WideSlot* p = bucket(t, b);
    bool end = false;
    for (int s = 0; s < SLOTS; ++s) {
      uint64_t a, r;
      asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(r) : "l"(p + s));
      if ((uint32_t)r == 0) end = true;
      else if ((int64_t)a == k) f((uint32_t)r);
    }
    return end;
  }
};

// ---------------- min / max of the build keys ---------------------------------------------------------
__global__ void __launch_bounds__(256) k_key_minmax(ColRef key, int64_t n, int64_t* out) {
  int64_t lo = INT64_MAX, hi = INT64_MIN;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
    if (!row_valid(key, row)) continue;
    const int64_t k = (int64_t)load_value(key, row);
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const int64_t l2 = __shfl_xor_sync(0xffffffffu, lo, m), h2 = __shfl_xor_sync(0xffffffffu, hi, m);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin((long long*)out, (long long)lo);
    atomicMax((long long*)out + 1, (long long)hi);
  }
}

// ---------------- build -----------------------------------------------------------------------------------
template <bool COMPACT>
__global__ void __launch_bounds__(256) k_join_build(ColRef key, int64_t n, TableRef t) {
  using F = Fmt<COMPACT>;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
    if (!row_valid(key, row)) continue;  // NULL keys never match
    const int64_t k = (int64_t)load_value(key, row);
    uint32_t b = bucket_of(t, k);
    for (uint32_t probes = 0; probes <= t.bmask; ++probes) {
      if (F::insert(t, b, k, (uint32_t)row + 1u)) break;
      b = (b + 1) & t.bmask;
    }
  }
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

// the thread's 4 consecutive keys (vector loads on the aligned 8-byte fast path)
QG_D void load_keys4(const ColRef& key, int64_t row0, int64_t n, int64_t* k, bool* live) {
  if (row0 + JR <= n && key.valid == nullptr && (key.dtype == DT_I64 || key.dtype == DT_F64)) {
    uint64_t a, b, c2, d;
    ld_nc_v2u64((const uint64_t*)key.data + row0, a, b);
    ld_nc_v2u64((const uint64_t*)key.data + row0 + 2, c2, d);
    k[0] = (int64_t)a; k[1] = (int64_t)b; k[2] = (int64_t)c2; k[3] = (int64_t)d;
#pragma unroll
    for (int j = 0; j < JR; ++j) live[j] = true;
    return;
  }
#pragma unroll
  for (int j = 0; j < JR; ++j) {
    const int64_t row = row0 + j;
    live[j] = row < n && row_valid(key, row);
    k[j] = live[j] ? (int64_t)load_value(key, row) : 0;
  }
}

// pass 1: per probe row, number of output rows and the first matching build row
template <bool COMPACT>
__global__ void __launch_bounds__(JT) k_join_count(ColRef key, int64_t n, TableRef t, int left_join, uint32_t* first_match,
                                                    uint32_t* tile_counts, uint32_t* dup_flag) {
  using F = Fmt<COMPACT>;
  __shared__ uint32_t s_w[9];
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile * JTILE < n; tile += gridDim.x) {
    const int64_t row0 = tile * JTILE + (int64_t)tid * JR;
    int64_t k[JR];
    bool live[JR];
    load_keys4(key, row0, n, k, live);
    uint32_t cnt[JR], first[JR], b[JR];
    bool more[JR];
    // round 1 for all four rows together (four independent sector reads in flight), rare continuations after
#pragma unroll
    for (int j = 0; j < JR; ++j) {
      cnt[j] = 0;
      first[j] = 0;
      live[j] = live[j] && F::in_range(t, k[j]);
      b[j] = bucket_of(t, k[j]);
    }
#pragma unroll
    for (int j = 0; j < JR; ++j) more[j] = live[j] && !F::scan(t, b[j], k[j], cnt[j], first[j]);
#pragma unroll
    for (int j = 0; j < JR; ++j) {
      for (uint32_t probes = 1; more[j] && probes <= t.bmask; ++probes) {
        b[j] = (b[j] + 1) & t.bmask;
        more[j] = !F::scan(t, b[j], k[j], cnt[j], first[j]);
      }
    }
    uint32_t mine = 0;
    bool dup = false;
#pragma unroll
    for (int j = 0; j < JR; ++j) {
      if (row0 + j >= n) continue;
      if (cnt[j] > 1) dup = true;
      mine += (left_join && cnt[j] == 0) ? 1u : cnt[j];
    }
    if (row0 + JR <= n) {
      *reinterpret_cast<uint4*>(first_match + row0) = make_uint4(first[0], first[1], first[2], first[3]);
    } else {
#pragma unroll
      for (int j = 0; j < JR; ++j)
        if (row0 + j < n) first_match[row0 + j] = first[j];
    }
    if (dup) atomicOr(dup_flag, 1u);
    uint32_t total;
    block_excl_scan_256(mine, s_w, total);
    if (tid == 0) tile_counts[tile] = total;
  }
}

// pass 2 (generic): (left row, right row) index pairs in probe-row order
template <bool COMPACT>
__global__ void __launch_bounds__(JT) k_join_emit(ColRef key, int64_t n, TableRef t, int left_join, int has_dups,
                                                   const uint32_t* first_match, const int64_t* tile_off, uint32_t* out_left,
                                                   uint32_t* out_right) {
  using F = Fmt<COMPACT>;
  __shared__ uint32_t s_w[9];
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile * JTILE < n; tile += gridDim.x) {
    const int64_t row0 = tile * JTILE + (int64_t)tid * JR;
    uint32_t cnt[JR], first[JR];
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < JR; ++j) {
      const int64_t row = row0 + j;
      const bool inb = row < n;
      first[j] = inb ? first_match[row] : 0;
      cnt[j] = first[j] ? 1u : 0u;
      if (inb && has_dups && first[j]) {
        const int64_t k = (int64_t)load_value(key, row);
            uint32_t b = bucket_of(t, k), c2 = 0, f2 = 0;
        for (uint32_t probes = 0; probes <= t.bmask; ++probes) {
          if (F::scan(t, b, k, c2, f2)) break;
          b = (b + 1) & t.bmask;
        }
        cnt[j] = c2;
      }
      mine += (inb && left_join && cnt[j] == 0) ? 1u : (inb ? cnt[j] : 0u);
    }
    uint32_t total;
    int64_t pos = tile_off[tile] + block_excl_scan_256(mine, s_w, total);
#pragma unroll 1
    for (int j = 0; j < JR; ++j) {
      const int64_t row = row0 + j;
      if (row >= n) continue;
      if (cnt[j] == 0) {
        if (left_join) {
          out_left[pos] = (uint32_t)row;
          out_right[pos] = 0xffffffffu;
          ++pos;
        }
      } else if (!has_dups) {
        out_left[pos] = (uint32_t)row;
        out_right[pos] = first[j] - 1;
        ++pos;
      } else {
        const int64_t k = (int64_t)load_value(key, row);
            uint32_t b = bucket_of(t, k);
        for (uint32_t probes = 0; probes <= t.bmask; ++probes) {
          const bool end = F::for_matches(t, b, k, [&](uint32_t r1) {
            out_left[pos] = (uint32_t)row;
            out_right[pos] = r1 - 1;
            ++pos;
          });
          if (end) break;
          b = (b + 1) & t.bmask;
        }
      }
    }
  }
}

// pass 2 (fused, duplicate-free build side): every probe row yields 0 or 1 output rows; write the output columns
__global__ void __launch_bounds__(JT) k_join_emit_fused(int64_t n, int left_join, const uint32_t* first_match, const int64_t* tile_off,
                                                         const __grid_constant__ JoinEmitCols C) {
  __shared__ uint32_t s_w[9];
  const int tid = threadIdx.x;
  for (int64_t tile = blockIdx.x; tile * JTILE < n; tile += gridDim.x) {
    const int64_t row0 = tile * JTILE + (int64_t)tid * JR;
    const bool full = row0 + JR <= n;
    uint32_t first[JR];
    if (full) {
      const uint4 f = *reinterpret_cast<const uint4*>(first_match + row0);
      first[0] = f.x; first[1] = f.y; first[2] = f.z; first[3] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < JR; ++j) first[j] = row0 + j < n ? first_match[row0 + j] : 0;
    }
    bool emit[JR];
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < JR; ++j) {
      emit[j] = row0 + j < n && (first[j] != 0 || left_join);
      mine += emit[j] ? 1u : 0u;
    }
    uint32_t total;
    const int64_t pos0 = tile_off[tile] + block_excl_scan_256(mine, s_w, total);
    const bool dense = full && mine == JR && (pos0 & 1) == 0;  // 4 consecutive outputs, 16-byte aligned for 8-byte columns
    for (int ci = 0; ci < C.n_left; ++ci) {
      const JoinOutCol c = C.left[ci];
      if (c.width == 8) {
        uint64_t v[JR];
        if (full) {
          ld_nc_v2u64((const uint64_t*)c.in + row0, v[0], v[1]);
          ld_nc_v2u64((const uint64_t*)c.in + row0 + 2, v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < JR; ++j) v[j] = emit[j] ? ld_stream_u64((const uint64_t*)c.in + row0 + j) : 0ull;
        }
        if (dense) {
          uint64_t* o = (uint64_t*)c.out + pos0;
          *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(v[0], v[1]);
          *reinterpret_cast<ulonglong2*>(o + 2) = make_ulonglong2(v[2], v[3]);
        } else {
          int64_t pos = pos0;
#pragma unroll
          for (int j = 0; j < JR; ++j)
            if (emit[j]) ((uint64_t*)c.out)[pos++] = v[j];
        }
      } else {
        int64_t pos = pos0;
#pragma unroll
        for (int j = 0; j < JR; ++j)
          if (emit[j]) ((uint32_t*)c.out)[pos++] = ld_stream_u32((const uint32_t*)c.in + row0 + j);
      }
    }
    for (int ci = 0; ci < C.n_right; ++ci) {
      const JoinOutCol c = C.right[ci];
      if (c.width == 8) {
        uint64_t v[JR];
#pragma unroll
        for (int j = 0; j < JR; ++j) v[j] = (emit[j] && first[j]) ? __ldg((const unsigned long long*)c.in + (first[j] - 1)) : 0ull;
        if (dense) {
          uint64_t* o = (uint64_t*)c.out + pos0;
          *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(v[0], v[1]);
          *reinterpret_cast<ulonglong2*>(o + 2) = make_ulonglong2(v[2], v[3]);
        } else {
          int64_t pos = pos0;
#pragma unroll
          for (int j = 0; j < JR; ++j)
            if (emit[j]) ((uint64_t*)c.out)[pos++] = v[j];
        }
      } else {
        uint32_t v[JR];
#pragma unroll
        for (int j = 0; j < JR; ++j) v[j] = (emit[j] && first[j]) ? __ldg((const uint32_t*)c.in + (first[j] - 1)) : 0u;
        int64_t pos = pos0;
#pragma unroll
        for (int j = 0; j < JR; ++j)
          if (emit[j]) ((uint32_t*)c.out)[pos++] = v[j];
      }
    }
    if (C.right_valid_bytes != nullptr) {
      int64_t pos = pos0;
#pragma unroll
      for (int j = 0; j < JR; ++j)
        if (emit[j]) C.right_valid_bytes[pos++] = first[j] ? 1 : 0;
    }
  }
}

TableRef table_ref(const JoinTable& t) {
  TableRef r;
  r.slots = t.slots;
  r.bmask = t.cap / (t.compact ? 4 : 2) - 1;
  r.kmin = t.key_min;
  r.kmax = t.key_max;
  return r;
}

template <class K>
unsigned tile_grid(Ctx* c, K kernel, int64_t n) {
  return persistent_grid(c, kernel, JT, (n + JTILE - 1) / JTILE);
}

// ---- composite / Utf8 keys: a 64-bit surrogate per row, verified after the join ---------------------------------
// A key tuple (any mix of Int32/Int64 - compared as Int64 like coerce_numeric_types -, Float, Bool, Utf8 columns) is
// hashed to one Int64 surrogate per row; a row with a NULL key component gets an invalid surrogate (NULL keys never
// match).  The join then runs on the surrogates, and k_verify_pairs counts emitted pairs whose REAL keys differ - a
// 64-bit collision - in which case the host retries with another seed.
__global__ void k_hash_rows(const __grid_constant__ GroupTable g, int64_t n, uint64_t seed, int64_t* out, uint32_t* valid) {
  const int64_t n_round = ((n + 31) / 32) * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * blockDim.x) {
    bool ok = i < n;
    if (ok)
      for (int k = 0; k < g.n_keys; ++k) ok &= row_valid(g.keys[k], i);
    if (i < n) out[i] = ok ? (int64_t)mix64(hash_row(g, i) ^ seed) : 0;
    const uint32_t w = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && i < n) valid[i >> 5] = w;
  }
}

QG_D bool keys_equal2(const GroupTable& a, int64_t ra, const GroupTable& b, int64_t rb) {
  for (int k = 0; k < a.n_keys; ++k) {
    const ColRef& ca = a.keys[k];
    const ColRef& cb = b.keys[k];
    if (ca.dtype == DT_UTF8) {
      const int32_t* oa = (const int32_t*)ca.data;
      const int32_t* ob = (const int32_t*)cb.data;
      const int32_t ab = oa[ra], ae = oa[ra + 1], bb = ob[rb], be = ob[rb + 1];
      if (ae - ab != be - bb) return false;
      for (int32_t i = 0; i < ae - ab; ++i)
        if (ca.chars[ab + i] != cb.chars[bb + i]) return false;
    } else if (load_value(ca, ra) != load_value(cb, rb)) {
      return false;
    }
  }
  return true;
}

__global__ void k_verify_pairs(const __grid_constant__ GroupTable gl, const __grid_constant__ GroupTable gr, const uint32_t* out_left,
                               const uint32_t* out_right, int64_t n, uint32_t* mismatches) {
  uint32_t bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t l = out_left[i], r = out_right[i];
    if (l == 0xffffffffu || r == 0xffffffffu) continue;
    if (!keys_equal2(gl, l, gr, r)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

// ---- RIGHT / FULL: the build rows no probe row matched -------------------------------------------------------------
__global__ void k_mark_rows(const uint32_t* idx, int64_t n, uint32_t* bits) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = idx[i];
    if (r != 0xffffffffu) atomicOr(bits + (r >> 5), 1u << (r & 31));
  }
}
__global__ void k_list_unmarked(const uint32_t* bits, int64_t n, uint32_t* list, uint32_t* count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!((bits[i >> 5] >> (i & 31)) & 1u)) list[atomicAdd(count, 1u)] = (uint32_t)i;
}
__global__ void k_fill_u32(uint32_t* p, int64_t n, uint32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

void launch_key_minmax(Ctx* c, const ColRef& key, int64_t n, int64_t* d_minmax) {
  const int64_t init[2] = {INT64_MAX, INT64_MIN};
  QG_CUDA(cudaMemcpyAsync(d_minmax, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
  if (n <= 0) return;
  ProfScope ps(c, "k_key_minmax");
  k_key_minmax<<<persistent_grid(c, k_key_minmax, 256, (n + 255) / 256), 256, 0, c->stream>>>(key, n, d_minmax);
  QG_CUDA(cudaGetLastError());
}

void launch_join_build(Ctx* c, const ColRef& key, int64_t n, const JoinTable& t) {
  QG_CUDA(cudaMemsetAsync(t.slots, 0, t.bytes(), c->stream));
  if (n <= 0) return;
  ProfScope ps(c, "k_join_build");
  const int64_t blocks = (n + 255) / 256;
  if (t.compact) k_join_build<true><<<persistent_grid(c, k_join_build<true>, 256, blocks), 256, 0, c->stream>>>(key, n, table_ref(t));
  else k_join_build<false><<<persistent_grid(c, k_join_build<false>, 256, blocks), 256, 0, c->stream>>>(key, n, table_ref(t));
  QG_CUDA(cudaGetLastError());
}

int64_t join_tiles(int64_t n) { return (n + JTILE - 1) / JTILE; }

void launch_join_count(Ctx* c, const ColRef& key, int64_t n, const JoinTable& t, bool left_join, uint32_t* first_match,
                       uint32_t* tile_counts, uint32_t* dup_flag) {
  if (n <= 0) return;
  ProfScope ps(c, "k_join_count");
  if (t.compact)
    k_join_count<true><<<tile_grid(c, k_join_count<true>, n), JT, 0, c->stream>>>(key, n, table_ref(t), left_join ? 1 : 0, first_match, tile_counts, dup_flag);
  else
    k_join_count<false><<<tile_grid(c, k_join_count<false>, n), JT, 0, c->stream>>>(key, n, table_ref(t), left_join ? 1 : 0, first_match, tile_counts, dup_flag);
  QG_CUDA(cudaGetLastError());
}

void launch_join_emit(Ctx* c, const ColRef& key, int64_t n, const JoinTable& t, bool left_join, bool has_dups,
                      const uint32_t* first_match, const int64_t* tile_off, uint32_t* out_left, uint32_t* out_right) {
  if (n <= 0) return;
  ProfScope ps(c, "k_join_emit");
  if (t.compact)
    k_join_emit<true><<<tile_grid(c, k_join_emit<true>, n), JT, 0, c->stream>>>(key, n, table_ref(t), left_join ? 1 : 0, has_dups ? 1 : 0, first_match,
                                                                 tile_off, out_left, out_right);
  else
    k_join_emit<false><<<tile_grid(c, k_join_emit<false>, n), JT, 0, c->stream>>>(key, n, table_ref(t), left_join ? 1 : 0, has_dups ? 1 : 0, first_match,
                                                                  tile_off, out_left, out_right);
  QG_CUDA(cudaGetLastError());
}

void launch_join_emit_fused(Ctx* c, int64_t n, bool left_join, const uint32_t* first_match, const int64_t* tile_off,
                            const JoinEmitCols& cols) {
  if (n <= 0) return;
  ProfScope ps(c, "k_join_emit_fused");
  k_join_emit_fused<<<tile_grid(c, k_join_emit_fused, n), JT, 0, c->stream>>>(n, left_join ? 1 : 0, first_match, tile_off, cols);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu

namespace qgpu {

static GroupTable key_table(const std::vector<ColRef>& keys) {
  GroupTable g;
  memset(&g, 0, sizeof g);
  g.n_keys = (int32_t)keys.size();
  for (size_t i = 0; i < keys.size(); ++i) g.keys[i] = keys[i];
  return g;
}

void launch_hash_rows(Ctx* c, const std::vector<ColRef>& keys, int64_t n, uint64_t seed, int64_t* out, uint32_t* valid) {
  if (n <= 0) return;
  ProfScope ps(c, "k_hash_rows");
  k_hash_rows<<<persistent_grid(c, k_hash_rows, 256, (n + 255) / 256), 256, 0, c->stream>>>(key_table(keys), n, seed, out, valid);
  QG_CUDA(cudaGetLastError());
}

void launch_verify_pairs(Ctx* c, const std::vector<ColRef>& lkeys, const std::vector<ColRef>& rkeys, const uint32_t* out_left,
                         const uint32_t* out_right, int64_t n, uint32_t* d_mismatches) {
  if (n <= 0) return;
  ProfScope ps(c, "k_verify_pairs");
  k_verify_pairs<<<persistent_grid(c, k_verify_pairs, 256, (n + 255) / 256), 256, 0, c->stream>>>(key_table(lkeys), key_table(rkeys), out_left,
                                                                                                 out_right, n, d_mismatches);
  QG_CUDA(cudaGetLastError());
}

void launch_unmatched_rows(Ctx* c, const uint32_t* matched_idx, int64_t n_idx, int64_t n_rows, uint32_t* bits, uint32_t* list,
                           uint32_t* d_count) {
  QG_CUDA(cudaMemsetAsync(bits, 0, (size_t)((n_rows + 31) / 32) * 4, c->stream));
  QG_CUDA(cudaMemsetAsync(d_count, 0, 4, c->stream));
  if (n_idx > 0) {
    ProfScope ps(c, "k_mark_rows");
    k_mark_rows<<<persistent_grid(c, k_mark_rows, 256, (n_idx + 255) / 256), 256, 0, c->stream>>>(matched_idx, n_idx, bits);
    QG_CUDA(cudaGetLastError());
  }
  if (n_rows > 0) {
    ProfScope ps(c, "k_list_unmarked");
    k_list_unmarked<<<persistent_grid(c, k_list_unmarked, 256, (n_rows + 255) / 256), 256, 0, c->stream>>>(bits, n_rows, list, d_count);
    QG_CUDA(cudaGetLastError());
  }
}

void launch_fill_u32(Ctx* c, uint32_t* p, int64_t n, uint32_t v) {
  if (n <= 0) return;
  ProfScope ps(c, "k_fill_u32");
  k_fill_u32<<<persistent_grid(c, k_fill_u32, 256, (n + 255) / 256), 256, 0, c->stream>>>(p, n, v);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
