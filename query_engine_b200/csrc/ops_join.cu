// ops_join.cu — the join operators: the body of QueryExecutor::execute_join (crates/query-executor/src/executor.rs:
// 343-399) on device tables - radix-partitioned fast path (pjoin_kernels.cu), generic path (join_kernels.cu) - and the
// multi-GPU join whose radix scatter is the exchange (crates/query-distributed/src/planner.rs:228-249).
#include <algorithm>
#include <cstring>

#include "join_kernels.cuh"
#include "pjoin2_kernels.cuh"
#include "pjoin_kernels.cuh"
#include "qgpu_internal.h"
#include "rscatter_kernels.cuh"
#include "util_kernels.cuh"

namespace qgpu {

// ----------------------------------------------------------------------------------------------
// HashJoin INNER / LEFT on `left.col = right.col`
// ----------------------------------------------------------------------------------------------
namespace {
// min / max of an integer key column, cached on the column (one small kernel + one sync the first time)
void column_range(Ctx* c, const Column& col, int64_t& lo, int64_t& hi) {
  if (!col.has_range) {
    DevBuf mm(c, 256);
    launch_key_minmax(c, colref_of(col), col.n, mm.as<int64_t>());
    int64_t* h = (int64_t*)c->scratch(16);
    QG_CUDA(cudaMemcpyAsync(h, mm.p, 16, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    col.vmin = h[0];
    col.vmax = h[1];
    col.has_range = true;
  }
  lo = col.vmin;
  hi = col.vmax;
}

// Radix-partitioned join (pjoin_kernels.cuh): Int64 keys without validity, fixed-width columns without validity on
// both sides, a duplicate-free build side.  Returns false when a device-side check (region overflow, duplicate or
// marker-valued build key) asks for another path instead; `out` is then untouched.  `allow_dense`: try the
// direct-addressed layout when the build key's range permits.
bool pjoin(Ctx* c, const Table& left, const Table& right, int lki, int rki, bool left_join, bool allow_dense, HostMarks& hm,
           std::shared_ptr<Table>& out, bool& empty, bool& was_dense) {
  const int64_t nl = left.n_rows, nr = right.n_rows;
  const int ncl = (int)left.cols.size(), ncr = (int)right.cols.size();
  const uint64_t region_bytes = (uint64_t)c->opt("join.region_bytes", 16 << 20);
  const uint32_t bucket_factor = (uint32_t)std::max<int64_t>(1, c->opt("join.bucket_factor", 2));

  PJMap map;
  memset(&map, 0, sizeof map);
  uint32_t n_entries = 0;  // table entries per region
  was_dense = false;
  if (allow_dense) {
    int64_t kmin, kmax;
    column_range(c, right.cols[(size_t)rki], kmin, kmax);
    hm.mark("build key range");
    const uint64_t range = kmax >= kmin ? (uint64_t)kmax - (uint64_t)kmin + 1 : 0;
    if (range != 0 && range <= (uint64_t)nr * (uint64_t)c->opt("join.dense_max_sparsity", 8) && range < (1ull << 36)) {
      uint32_t shift = 1;
      while ((8ull << shift) < region_bytes) ++shift;                       // keys per region from the region size ...
      while (((range - 1) >> shift) + 1 > PJ_MAX_PARTS) ++shift;             // ... unless that needs too many regions
      if (c->opt("join.parts", 0) > 0) {
        shift = 1;
        while (((range - 1) >> shift) + 1 > (uint64_t)c->opt("join.parts", 0)) ++shift;
      }
      map.dense = 1;
      map.shift = shift;
      map.kmin = kmin;
      map.range = range;
      map.parts = (uint32_t)(((range - 1) >> shift) + 1);
      map.drop_outside = 0;
      n_entries = 1u << shift;
      was_dense = true;
    }
  }
  if (!map.dense) {
    // hashed: two 32-byte buckets (four slots) per build row - at load <= 0.23 nearly every insert and lookup ends in
    // its first bucket, and a warp repeats a dependent L2 round trip until its LONGEST chain is done (measured at SF10:
    // 2.9 ms with one bucket per row, 1.7 ms with two); regions of <= 16 MB of table
    const uint64_t region_bytes_h = (uint64_t)c->opt("join.region_bytes_hashed", 16 << 20);
    uint32_t parts = 1;
    while (parts < PJ_MAX_PARTS && (uint64_t)nr * 34 * bucket_factor / parts > region_bytes_h) parts <<= 1;
    if (c->opt("join.parts", 0) > 0) {
      parts = 1;
      while (parts < (uint32_t)c->opt("join.parts", 0) && parts < PJ_MAX_PARTS) parts <<= 1;
    }
    map.parts = parts;
    while ((1u << map.log_parts) < parts) ++map.log_parts;
  }
  const uint32_t parts = map.parts;
  auto region_cap = [&](int64_t n, uint32_t mult) {
    uint64_t cap = (uint64_t)n / parts;
    cap += cap / 8 + 1024;  // 12.5 % slack: hash-partitioned distinct keys / evenly dense keys stay far inside it
    return (uint32_t)((cap + mult - 1) / mult * mult);
  };
  const uint32_t cap_l = region_cap(nl, PJ_TILE), cap_r = region_cap(nr, 32);
  if ((uint64_t)parts * cap_l >= 0xfff00000ull || (uint64_t)parts * cap_r >= 0xfff00000ull) return false;
  if (!map.dense) n_entries = cap_r * bucket_factor;
  const size_t table_bytes = (size_t)parts * n_entries * (map.dense ? 8 : sizeof(PJBucket));

  // control block (zeroed): four arrays of per-region counters, one counter per 128-byte line, then the single words,
  // each on a line of its own (hot words sharing a sector would serialise their atomics on one L2 slice)
  constexpr size_t CTR_WORDS = (size_t)PJ_MAX_PARTS * PJ_CTR_STRIDE;
  DevBuf ctl(c, (4 * CTR_WORDS + 8 * 64) * 4, true);
  uint32_t* cnt_l = ctl.as<uint32_t>();            // probe-side rows per region
  uint32_t* cnt_r = cnt_l + CTR_WORDS;              // build-side rows per region
  uint32_t* cleared = cnt_l + 2 * CTR_WORDS;        // finished clear items per region
  uint32_t* inserted = cnt_l + 3 * CTR_WORDS;       // finished insert items per region
  uint32_t* singles = cnt_l + 4 * CTR_WORDS;
  uint32_t* flags = singles;                        // [0]: PJFlags, [2..3]: output row count
  unsigned long long* out_cursor = reinterpret_cast<unsigned long long*>(singles + 64);
  uint32_t* ticket = singles + 128;
  unsigned long long* timing_words = reinterpret_cast<unsigned long long*>(singles + 192);

  // ---- region-major copies of both sides ----
  auto scatter = [&](const Table& t, int kc, uint32_t cap, uint32_t* cnt, std::vector<DevBuf>& bufs, const PJMap& m, const char* name) {
    PJCols pc;
    memset(&pc, 0, sizeof pc);
    pc.n = (int)t.cols.size();
    pc.key_col = kc;
    for (int i = 0; i < pc.n; ++i) {
      const int w = dtype_width(t.cols[i].dtype);
      bufs.emplace_back(c, pad_bytes((size_t)parts * cap * w));
      pc.c[i] = PJCol{t.cols[i].data, bufs.back().p, w, 0};
    }
    launch_pj_scatter(c, (const int64_t*)t.cols[kc].data, t.n_rows, m, cap, cnt, flags, pc, nullptr, name);
  };
  std::vector<DevBuf> lp, rp;
  scatter(right, rki, cap_r, cnt_r, rp, map, "k_pj_scatter_build");
  PJMap lmap = map;
  lmap.drop_outside = left_join ? 0 : 1;  // INNER: probe keys outside the build range cannot match
  scatter(left, lki, cap_l, cnt_l, lp, lmap, "k_pj_scatter_probe");
  hm.mark("pjoin scatter launches");

  // ---- region tables + probe + emit: a duplicate-free build side bounds the output by the probe side ----
  int inline_right = -1;
  if (ncr == 2 && dtype_width(right.cols[1 - rki].dtype) == 8) inline_right = 1 - rki;
  DevBuf tab(c, table_bytes + 256);

  out = new_table(c);
  PJEmit E;
  memset(&E, 0, sizeof E);
  E.n_left = ncl;
  E.n_right = ncr;
  E.left_key = lki;
  E.right_key = rki;
  E.inline_right = inline_right;
  E.left_join = left_join ? 1 : 0;
  E.prefetch_left = -1;
  for (int i = 0; i < ncl && E.prefetch_left < 0; ++i)
    if (i != lki && dtype_width(left.cols[i].dtype) == 8) E.prefetch_left = i;
  for (int i = 0; i < ncl; ++i) {  // schema = left fields ++ right fields (executor.rs:532-537)
    Column d = alloc_column(c, left.cols[i].dtype, nl, false);
    E.left[i] = PJCol{lp[(size_t)i].p, d.data, dtype_width(d.dtype), 0};
    out->fields.push_back(left.fields[i]);
    out->cols.push_back(d);
  }
  DevBuf vbytes;
  if (left_join) {
    vbytes = DevBuf(c, pad_bytes((size_t)nl));
    E.right_valid_bytes = vbytes.as<uint8_t>();
  }
  for (int i = 0; i < ncr; ++i) {
    Column d = alloc_column(c, right.cols[i].dtype, nl, left_join);
    E.right[i] = PJCol{rp[(size_t)i].p, d.data, dtype_width(d.dtype), 0};
    Field f = right.fields[i];
    if (left_join) f.nullable = true;
    out->fields.push_back(f);
    out->cols.push_back(d);
  }
  PJJoinArgs A;
  memset(&A, 0, sizeof A);
  A.build_key = rp[(size_t)rki].as<int64_t>();
  A.build_val = inline_right >= 0 ? rp[(size_t)inline_right].as<uint64_t>() : nullptr;
  A.probe_key = lp[(size_t)lki].as<int64_t>();
  A.count_r = cnt_r;
  A.count_l = cnt_l;
  A.table = tab.p;
  A.cap_r = cap_r;
  A.cap_l = cap_l;
  A.n_entries = n_entries;
  A.flags = flags;
  A.cleared = cleared;
  A.inserted = inserted;
  A.out_cursor = out_cursor;
  A.ticket = ticket;
  const bool timing = c->opt("debug.join_timing", 0) != 0;
  A.timing = timing ? timing_words : nullptr;
  launch_pj_join(c, lmap, A, E);
  hm.mark("pjoin build/probe launches");

  uint32_t* h = (uint32_t*)c->scratch(16);
  QG_CUDA(cudaMemcpyAsync(h, flags, 4, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaMemcpyAsync(h + 2, out_cursor, 8, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));
  hm.mark("pjoin sync");
  const uint32_t fl = h[0];
  const int64_t n_out = (int64_t)(((uint64_t)h[3] << 32) | h[2]);
  if (timing) {
    unsigned long long t[24];
    QG_CUDA(cudaMemcpy(t, A.timing, sizeof t, cudaMemcpyDeviceToHost));
    const char* names[4] = {"clear", "insert", "probe", "none"};
    for (int k = 0; k < 3; ++k)
      fprintf(stderr, "[qgpu join] %-6s items %8llu  avg %8.2f us  of which waiting on counters %8.2f us\n", names[k], t[k * 4],
              t[k * 4] ? t[k * 4 + 1] * 1e-3 / t[k * 4] : 0.0, t[k * 4] ? t[k * 4 + 2] * 1e-3 / t[k * 4] : 0.0);
    fprintf(stderr, "[qgpu join] staged-tile wait %8.2f us per probe tile\n", t[8] ? t[20] * 1e-3 / t[8] : 0.0);
    if (t[8])
      fprintf(stderr, "[qgpu join] probe tile, us from its start (thread 0): lookups done %.2f, all warps done %.2f, output reserved %.2f, stores issued %.2f\n",
              t[16] * 1e-3 / t[8], t[17] * 1e-3 / t[8], t[18] * 1e-3 / t[8], t[19] * 1e-3 / t[8]);
  }
  if (fl != 0) {
    out.reset();
    return false;
  }
  empty = n_out == 0;
  if (empty) {
    out.reset();
    return true;
  }
  out->n_rows = n_out;
  out->batch_offsets = {0, n_out};
  for (Column& col : out->cols) col.n = n_out;
  if (left_join && ncr > 0) {
    Column& c0 = out->cols[(size_t)ncl];
    launch_bytes_to_bits(c, vbytes.as<uint8_t>(), n_out, c0.valid);
    for (int i = 1; i < ncr; ++i)
      QG_CUDA(cudaMemcpyAsync(out->cols[(size_t)ncl + i].valid, c0.valid, (size_t)((n_out + 31) / 32) * 4, cudaMemcpyDeviceToDevice,
                              c->stream));
  }
  return true;
}

// Second-generation radix join (rscatter_kernels.cu + pjoin2_kernels.cu): 8-byte columns only, at most RS_MAX_COLS per
// side.  The probe side is scattered straight into the output columns at exact positions (region histogram first),
// the build + probe kernel writes the build-side output columns at the probe row's position; region tables live in a
// ring that stays in L2.  Same contract as pjoin(): false = a device-side check asks for another path.
bool pjoin2(Ctx* c, const Table& left, const Table& right, int lki, int rki, bool left_join, bool allow_dense, HostMarks& hm,
            std::shared_ptr<Table>& out, bool& empty, bool& was_dense) {
  const int64_t nl = left.n_rows, nr = right.n_rows;
  const int ncl = (int)left.cols.size(), ncr = (int)right.cols.size();
  const uint64_t region_bytes = (uint64_t)c->opt("join.region_bytes2", 16 << 20);
  const uint32_t bucket_factor = (uint32_t)std::max<int64_t>(1, c->opt("join.bucket_factor", 2));
  const uint32_t ring = (uint32_t)std::max<int64_t>(3, c->opt("join.ring", 3));

  PJMap map;
  memset(&map, 0, sizeof map);
  uint32_t n_entries = 0;
  was_dense = false;
  if (allow_dense) {
    int64_t kmin, kmax;
    column_range(c, right.cols[(size_t)rki], kmin, kmax);
    const uint64_t range = kmax >= kmin ? (uint64_t)kmax - (uint64_t)kmin + 1 : 0;
    if (range != 0 && range <= (uint64_t)nr * (uint64_t)c->opt("join.dense_max_sparsity", 8) && range < (1ull << 36)) {
      uint32_t shift = 1;
      while ((8ull << shift) < region_bytes) ++shift;
      while (((range - 1) >> shift) + 1 > PJ_MAX_PARTS) ++shift;
      if (c->opt("join.parts", 0) > 0) {
        shift = 1;
        while (((range - 1) >> shift) + 1 > (uint64_t)c->opt("join.parts", 0)) ++shift;
      }
      map.dense = 1;
      map.shift = shift;
      map.kmin = kmin;
      map.range = range;
      map.parts = (uint32_t)(((range - 1) >> shift) + 1);
      n_entries = 1u << shift;
      was_dense = true;
    }
  }
  if (!map.dense) {
    const uint64_t region_bytes_h = (uint64_t)c->opt("join.region_bytes_hashed2", 8 << 20);
    uint32_t parts = 1;
    while (parts < PJ_MAX_PARTS && (uint64_t)nr * 34 * bucket_factor / parts > region_bytes_h) parts <<= 1;
    if (c->opt("join.parts", 0) > 0) {
      parts = 1;
      while (parts < (uint32_t)c->opt("join.parts", 0) && parts < PJ_MAX_PARTS) parts <<= 1;
    }
    map.parts = parts;
    while ((1u << map.log_parts) < parts) ++map.log_parts;
  }
  const uint32_t parts = map.parts;
  uint64_t cap64 = (uint64_t)nr / parts;
  cap64 += cap64 / 8 + 1024;
  const uint32_t cap_r = (uint32_t)((cap64 + 31) / 32 * 32);
  if ((uint64_t)parts * cap_r + RS_TILE >= 0xfff00000ull || (uint64_t)nl >= 0xfff00000ull) return false;
  if (!map.dense) n_entries = cap_r * bucket_factor;
  const size_t table_bytes = (size_t)ring * n_entries * (map.dense ? 8 : sizeof(PJBucket));
  const uint32_t tiles_per_region = pj2_tiles_per_region(nl, parts);

  // control block (zeroed): per-region counters one per 128-byte line, region starts, then the single words
  constexpr size_t CTR_WORDS = (size_t)PJ_MAX_PARTS * PJ_CTR_STRIDE;
  DevBuf ctl(c, (6 * CTR_WORDS + (PJ_MAX_PARTS + 64) + 8 * 64) * 4, true);
  uint32_t* cnt_l = ctl.as<uint32_t>();          // probe rows per region (histogram)
  uint32_t* cur_l = cnt_l + CTR_WORDS;            // probe rows placed so far per region
  uint32_t* cnt_r = cnt_l + 2 * CTR_WORDS;        // build rows per region (the scatter's cursors)
  uint32_t* cleared = cnt_l + 3 * CTR_WORDS;
  uint32_t* inserted = cnt_l + 4 * CTR_WORDS;
  uint32_t* probed = cnt_l + 5 * CTR_WORDS;
  uint32_t* start_l = cnt_l + 6 * CTR_WORDS;      // [parts + 1]
  uint32_t* singles = start_l + PJ_MAX_PARTS + 64;
  uint32_t* flags = singles;
  unsigned long long* found = reinterpret_cast<unsigned long long*>(singles + 64);
  uint32_t* ticket = singles + 128;

  // ---- build side -> region-major copy (over-allocated regions; key first) ----
  std::vector<DevBuf> rp;
  std::vector<int> r_order{rki};
  for (int i = 0; i < ncr; ++i)
    if (i != rki) r_order.push_back(i);
  std::vector<void*> r_copy((size_t)ncr);
  {
    RSArgs a;
    memset(&a, 0, sizeof a);
    for (int s = 0; s < ncr; ++s) {
      rp.emplace_back(c, pad_bytes(((size_t)parts * cap_r + RS_TILE) * 8));
      a.in[s] = (const uint64_t*)right.cols[(size_t)r_order[(size_t)s]].data;
      a.out[s] = rp.back().as<uint64_t>();
      r_copy[(size_t)r_order[(size_t)s]] = rp.back().p;
    }
    a.n = nr;
    a.M = map;
    a.M.drop_outside = 0;
    a.cap = cap_r;
    a.ctr_stride = PJ_CTR_STRIDE;
    a.cursor = cnt_r;
    a.flags = flags;
    a.trash_row = parts * cap_r;
    launch_rscatter(c, a, ncr, nullptr, "k_rscatter_build");
  }

  // ---- probe side -> the output's left columns, exact placement ----
  PJMap lmap = map;
  lmap.drop_outside = left_join ? 0 : 1;  // INNER: probe keys outside the build range cannot match
  launch_region_hist(c, (const int64_t*)left.cols[(size_t)lki].data, nl, lmap, cnt_l, PJ_CTR_STRIDE);
  launch_region_offsets(c, cnt_l, PJ_CTR_STRIDE, parts, (tiles_per_region - 1) * PJ2_TILE, start_l, flags);
  out = new_table(c);
  {
    RSArgs a;
    memset(&a, 0, sizeof a);
    std::vector<int> l_order{lki};
    for (int i = 0; i < ncl; ++i)
      if (i != lki) l_order.push_back(i);
    for (int i = 0; i < ncl; ++i) {  // schema = left fields ++ right fields (executor.rs:532-537)
      out->fields.push_back(left.fields[(size_t)i]);
      out->cols.push_back(alloc_column(c, left.cols[(size_t)i].dtype, nl, false));
    }
    for (int s = 0; s < ncl; ++s) {
      a.in[s] = (const uint64_t*)left.cols[(size_t)l_order[(size_t)s]].data;
      a.out[s] = (uint64_t*)out->cols[(size_t)l_order[(size_t)s]].data;
    }
    a.n = nl;
    a.M = lmap;
    a.ctr_stride = PJ_CTR_STRIDE;
    a.start = start_l;
    a.cursor = cur_l;
    a.flags = flags;
    launch_rscatter(c, a, ncl, nullptr, "k_rscatter_probe");
  }
  hm.mark("pjoin2 scatter launches");

  // ---- region tables + probe: the right columns land at the probe rows' positions ----
  int inline_right = -1;
  if (ncr == 2) inline_right = 1 - rki;
  DevBuf tab(c, table_bytes + 256);
  DevBuf vbytes(c, pad_bytes((size_t)nl));
  PJ2Args A;
  memset(&A, 0, sizeof A);
  // The build-side key column of the result holds, row for row, the probe-side key wherever it is valid (an equi-join on
  // one Int64 key): it SHARES the probe-side output column's data buffer (Arrow buffers are immutable and may back
  // several arrays; the values under NULL slots of a LEFT join are unspecified) and owns only its validity bitmap -
  // 8 bytes per output row that are never written ("join.alias_key" = 0 materialises the column instead).
  const bool alias_key = c->opt("join.alias_key", 1) != 0 && right.cols[(size_t)rki].dtype == left.cols[(size_t)lki].dtype;
  for (int i = 0; i < ncr; ++i) {
    Column d;
    if (i == rki && alias_key) {
      d = out->cols[(size_t)lki];
      d.owned = false;
      d.valid = nullptr;
      d.has_range = false;
      if (left_join) {  // only the bitmap is this column's own
        d.valid = (uint32_t*)c->zalloc(bitmap_bytes(nl));
        d.owned_valid = true;
        d.null_count = -1;
      }
    } else {
      d = alloc_column(c, right.cols[(size_t)i].dtype, nl, left_join);
    }
    Field f = right.fields[(size_t)i];
    if (left_join) f.nullable = true;
    out->fields.push_back(f);
    out->cols.push_back(d);
    if (i == rki) A.out_right_key = alias_key ? nullptr : (uint64_t*)d.data;
    else if (i == inline_right) A.out_inline = (uint64_t*)d.data;
    else A.gather[A.n_gather++] = PJCol{r_copy[(size_t)i], d.data, 8, 0};
  }
  A.build_key = (const int64_t*)r_copy[(size_t)rki];
  A.build_val = inline_right >= 0 ? (const uint64_t*)r_copy[(size_t)inline_right] : nullptr;
  A.count_r = cnt_r;
  A.cap_r = cap_r;
  A.n_entries = n_entries;
  A.probe_key = (const uint64_t*)out->cols[(size_t)lki].data;
  A.start_l = start_l;
  A.probe_rows_alloc = (uint64_t)nl;
  A.tiles_per_region = tiles_per_region;
  A.ring = ring;
  A.table = tab.p;
  A.flags = flags;
  A.cleared = cleared;
  A.inserted = inserted;
  A.probed = probed;
  A.ticket = ticket;
  A.found = found;
  A.vbytes = vbytes.as<uint8_t>();
  launch_pj_join2(c, lmap, A);
  hm.mark("pjoin2 build/probe launch");

  uint32_t* h = (uint32_t*)c->scratch(32);
  QG_CUDA(cudaMemcpyAsync(h, flags, 4, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaMemcpyAsync(h + 2, found, 8, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaMemcpyAsync(h + 4, start_l + parts, 4, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));
  hm.mark("pjoin2 sync");
  const uint32_t fl = h[0];
  const int64_t n_found = (int64_t)(((uint64_t)h[3] << 32) | h[2]);
  const int64_t n_placed = (int64_t)h[4];  // probe rows that took part (INNER: keys outside the build range were dropped)
  if (fl != 0) {
    out.reset();
    return false;
  }
  const int64_t n_out = left_join ? n_placed : n_found;
  empty = n_out == 0;
  if (empty) {
    out.reset();
    return true;
  }
  if (!left_join && n_found < n_placed) {
    // INNER with unmatched probe rows: compact every output column by the match flags (order-preserving)
    const int64_t n_tiles = (n_placed + 1023) / 1024;
    DevBuf mask(c, (size_t)n_tiles * 128 + 256, true);
    DevBuf counts(c, (size_t)n_tiles * 4 + 256);
    DevBuf tile_off(c, (size_t)(n_tiles + 1) * 8 + 256);
    launch_bytes_to_bits(c, vbytes.as<uint8_t>(), n_placed, mask.as<uint32_t>());
    launch_mask_tile_counts(c, mask.as<uint32_t>(), n_tiles, counts.as<uint32_t>());
    scan_u32_to_i64(c, counts.as<uint32_t>(), n_tiles, tile_off.as<int64_t>());
    std::vector<CompactCol> cc;
    std::vector<Column> fresh(out->cols.size());
    const size_t alias_col = alias_key ? (size_t)(ncl + rki) : out->cols.size();
    for (size_t i = 0; i < out->cols.size(); ++i) {
      if (i == alias_col) continue;
      Column& col = out->cols[i];
      Column d = alloc_column(c, col.dtype, n_found, false);
      CompactCol x{};
      x.in = col.data;
      x.out = d.data;
      x.kind = CK_W8;
      cc.push_back(x);
      fresh[i] = d;
    }
    launch_compact(c, n_placed, 1024, mask.as<uint32_t>(), tile_off.as<int64_t>(), cc.data(), (int)cc.size(), n_found);
    for (size_t i = 0; i < out->cols.size(); ++i) {
      if (i == alias_col) continue;
      free_column(c, out->cols[i]);
      out->cols[i] = fresh[i];
    }
    if (alias_col < out->cols.size()) {  // the build-side key keeps sharing the (compacted) probe-side key
      Column d = out->cols[(size_t)lki];
      d.owned = false;
      d.has_range = false;
      out->cols[alias_col] = d;
    }
  }
  out->n_rows = n_out;
  out->batch_offsets = {0, n_out};
  for (Column& col : out->cols) col.n = n_out;
  if (left_join && ncr > 0) {
    Column& c0 = out->cols[(size_t)ncl];
    launch_bytes_to_bits(c, vbytes.as<uint8_t>(), n_out, c0.valid);
    for (int i = 1; i < ncr; ++i)
      QG_CUDA(cudaMemcpyAsync(out->cols[(size_t)ncl + i].valid, c0.valid, (size_t)((n_out + 31) / 32) * 4, cudaMemcpyDeviceToDevice,
                              c->stream));
  }
  return true;
}

bool equi_keys(const ExprTree& t, int i, int n_left, std::vector<std::pair<int, int>>& out) {
  const ExprNode& n = t.nodes[i];
  if (n.kind == QGPU_EXPR_BINARY && n.op == QGPU_AND) return equi_keys(t, n.left, n_left, out) && equi_keys(t, n.right, n_left, out);
  if (n.kind == QGPU_EXPR_BINARY && n.op == QGPU_EQUAL) {
    const ExprNode& a = t.nodes[n.left];
    const ExprNode& b = t.nodes[n.right];
    if (a.kind == QGPU_EXPR_COLUMN && b.kind == QGPU_EXPR_COLUMN) {
      if (a.op < n_left && b.op >= n_left) { out.push_back({a.op, b.op - n_left}); return true; }
      if (b.op < n_left && a.op >= n_left) { out.push_back({b.op, a.op - n_left}); return true; }
    }
  }
  return false;
}

// General equi-join on index pairs: any join type of INNER / LEFT / RIGHT / FULL, any number of key columns, Utf8 keys,
// duplicate build keys, validity bitmaps.  One flat table in HBM over ONE Int64 key per row: the key column itself for
// a single fixed-width key, else a 64-bit surrogate hashed from the key tuple (join_kernels.cu) - every emitted pair is
// then verified against the real keys and a collision makes the join start over with another seed.  RIGHT / FULL append
// the build rows that no probe row matched, with NULL left columns (Tier-2 semantics; the reference's
// execute_right_join / execute_full_join are Cartesian stubs, executor.rs:401-435).
std::shared_ptr<Table> join_general(Ctx* c, const Table& left, const Table& right, const std::vector<std::pair<int, int>>& keys,
                                    int join_type) {
  const int64_t nl = left.n_rows, nr = right.n_rows;
  const bool keep_left = join_type == QGPU_LEFT || join_type == QGPU_FULL;    // unmatched probe rows are emitted
  const bool keep_right = join_type == QGPU_RIGHT || join_type == QGPU_FULL;  // unmatched build rows are emitted
  const bool surrogate = !(keys.size() == 1 && left.cols[(size_t)keys[0].first].dtype != DT_UTF8);
  std::vector<ColRef> lkeys, rkeys;
  for (auto& k : keys) {
    lkeys.push_back(colref_of(left.cols[(size_t)k.first]));
    rkeys.push_back(colref_of(right.cols[(size_t)k.second]));
  }
  DevBuf out_left, out_right;
  int64_t n_out = 0;
  static const uint64_t kSeeds[4] = {0x0ull, 0x9e3779b97f4a7c15ull, 0xd1b54a32d192ed03ull, 0x8cb92ba72f3d8dd7ull};
  for (int attempt = 0;; ++attempt) {
    if (attempt == 4) throw_error(QGPU_ERR_UNSUPPORTED, "join keys collide under four 64-bit hashes");
    ColRef lkc = lkeys[0], rkc = rkeys[0];
    DevBuf lh, lv, rh, rv;
    if (surrogate) {
      lh = DevBuf(c, pad_bytes((size_t)nl * 8));
      lv = DevBuf(c, bitmap_bytes(nl));
      rh = DevBuf(c, pad_bytes((size_t)nr * 8));
      rv = DevBuf(c, bitmap_bytes(nr));
      launch_hash_rows(c, lkeys, nl, kSeeds[attempt], lh.as<int64_t>(), lv.as<uint32_t>());
      launch_hash_rows(c, rkeys, nr, kSeeds[attempt], rh.as<int64_t>(), rv.as<uint32_t>());
      lkc = ColRef{lh.p, lv.as<uint32_t>(), nullptr, DT_I64, 0};
      rkc = ColRef{rh.p, rv.as<uint32_t>(), nullptr, DT_I64, 0};
    }
    uint64_t table_cap = 1024;
    while (table_cap < (uint64_t)nr * 2) table_cap <<= 1;
    DevBuf mm(c, 256);
    launch_key_minmax(c, rkc, nr, mm.as<int64_t>());
    int64_t kmin, kmax;
    {
      int64_t* h = (int64_t*)c->scratch(16);
      QG_CUDA(cudaMemcpyAsync(h, mm.p, 16, cudaMemcpyDeviceToHost, c->stream));
      QG_CUDA(cudaStreamSynchronize(c->stream));
      kmin = h[0];
      kmax = h[1];
    }
    JoinTable jt;
    jt.cap = (uint32_t)table_cap;
    jt.compact = (c->opt("join.compact", 1) != 0 && kmax >= kmin && (uint64_t)kmax - (uint64_t)kmin < 0xffffffffull) ? 1 : 0;
    jt.key_min = kmin;
    jt.key_max = kmax;
    DevBuf tab(c, jt.bytes() + 256);
    jt.slots = tab.p;
    launch_join_build(c, rkc, nr, jt);
    const int64_t n_tiles = std::max<int64_t>(1, join_tiles(nl));
    DevBuf first(c, pad_bytes((size_t)std::max<int64_t>(nl, 1) * 4));
    DevBuf tile_counts(c, (size_t)n_tiles * 4 + 256, true);
    DevBuf tile_off(c, (size_t)(n_tiles + 1) * 8 + 256);
    DevBuf dup(c, 256, true);
    launch_join_count(c, lkc, nl, jt, keep_left, first.as<uint32_t>(), tile_counts.as<uint32_t>(), dup.as<uint32_t>());
    scan_u32_to_i64(c, tile_counts.as<uint32_t>(), n_tiles, tile_off.as<int64_t>());
    bool has_dups;
    {
      int64_t* h = (int64_t*)c->scratch(16);
      h[1] = 0;
      QG_CUDA(cudaMemcpyAsync(h, tile_off.as<int64_t>() + n_tiles, 8, cudaMemcpyDeviceToHost, c->stream));
      QG_CUDA(cudaMemcpyAsync(h + 1, dup.p, 4, cudaMemcpyDeviceToHost, c->stream));
      QG_CUDA(cudaStreamSynchronize(c->stream));
      n_out = nl > 0 ? h[0] : 0;
      has_dups = (uint32_t)h[1] != 0;
    }
    if (n_out + nr >= 0xfffffff0ll) throw_error(QGPU_ERR_UNSUPPORTED, "join result of 2^32 rows or more");
    // room for the unmatched build rows of RIGHT / FULL behind the pairs
    out_left = DevBuf(c, pad_bytes((size_t)(n_out + (keep_right ? nr : 0) + 1) * 4));
    out_right = DevBuf(c, pad_bytes((size_t)(n_out + (keep_right ? nr : 0) + 1) * 4));
    if (n_out > 0)
      launch_join_emit(c, lkc, nl, jt, keep_left, has_dups, first.as<uint32_t>(), tile_off.as<int64_t>(), out_left.as<uint32_t>(),
                       out_right.as<uint32_t>());
    if (!surrogate || n_out == 0) break;
    DevBuf bad(c, 256, true);
    launch_verify_pairs(c, lkeys, rkeys, out_left.as<uint32_t>(), out_right.as<uint32_t>(), n_out, bad.as<uint32_t>());
    if (read_u32(c, bad.as<uint32_t>()) == 0) break;
  }
  int64_t n_total = n_out;
  if (keep_right) {
    DevBuf bits(c, bitmap_bytes(nr)), cnt(c, 256);
    launch_unmatched_rows(c, out_right.as<uint32_t>(), n_out, nr, bits.as<uint32_t>(), out_right.as<uint32_t>() + n_out, cnt.as<uint32_t>());
    const int64_t m = (int64_t)read_u32(c, cnt.as<uint32_t>());
    launch_fill_u32(c, out_left.as<uint32_t>() + n_out, m, NULL_IDX);
    n_total += m;
  }
  if (n_total == 0) return nullptr;
  auto out = new_table(c);
  out->n_rows = n_total;
  out->batch_offsets = {0, n_total};
  std::vector<int> lcols(left.cols.size()), rcols(right.cols.size());
  for (size_t i = 0; i < lcols.size(); ++i) lcols[i] = (int)i;
  for (size_t i = 0; i < rcols.size(); ++i) rcols[i] = (int)i;
  auto lt = gather_rows(c, left, lcols, out_left.as<uint32_t>(), n_total, keep_right);
  auto rt = gather_rows(c, right, rcols, out_right.as<uint32_t>(), n_total, keep_left);
  for (size_t i = 0; i < lt->cols.size(); ++i) {  // schema = left fields ++ right fields (executor.rs:532-537)
    out->fields.push_back(lt->fields[i]);
    out->cols.push_back(lt->cols[i]);
    lt->cols[i].owned = false;
    lt->cols[i].owned_valid = false;
  }
  for (size_t i = 0; i < rt->cols.size(); ++i) {
    out->fields.push_back(rt->fields[i]);
    out->cols.push_back(rt->cols[i]);
    rt->cols[i].owned = false;
    rt->cols[i].owned_valid = false;
  }
  return out;
}
}  // namespace

std::shared_ptr<Table> op_join(Ctx* c, const Table& left, const Table& right, int join_type, const qgpu_expr* on) {
  const int64_t lb = (int64_t)left.batch_offsets.size() - 1, rb = (int64_t)right.batch_offsets.size() - 1;
  if (lb == 0 || rb == 0) return nullptr;  // executor.rs:350-352, even for LEFT
  if (join_type != QGPU_INNER && join_type != QGPU_LEFT && join_type != QGPU_RIGHT && join_type != QGPU_FULL)
    throw_error(QGPU_ERR_UNSUPPORTED, "CROSS joins keep the reference's Cartesian CPU path");
  if (!on) throw_error(QGPU_ERR_UNSUPPORTED, "join without ON keeps the reference's Cartesian CPU path");
  ExprTree t = parse_postfix(on);
  std::vector<std::pair<int, int>> keys;
  const int n_left = (int)left.cols.size();
  if (!equi_keys(t, t.root, n_left, keys))
    throw_error(QGPU_ERR_UNSUPPORTED, "ON must be Column = Column (AND ...) across the two sides to be accelerated");
  for (auto& k : keys)
    if (k.second >= (int)right.cols.size()) throw_error(QGPU_ERR_EXECUTION, "Column index %d out of bounds", k.second + n_left);
  if ((int)keys.size() > MAX_KEYS) throw_error(QGPU_ERR_UNSUPPORTED, "more than %d join key columns", MAX_KEYS);
  auto int_like = [](uint8_t d) { return d == DT_I64 || d == DT_I32; };
  for (auto& k : keys) {
    const uint8_t ld = left.cols[(size_t)k.first].dtype, rd = right.cols[(size_t)k.second].dtype;
    const bool ok = (int_like(ld) && int_like(rd)) || (ld == rd && (ld == DT_F64 || ld == DT_F32 || ld == DT_BOOL || ld == DT_UTF8));
    if (!ok) {
      if (ld == rd || (ld == DT_F64 && int_like(rd)) || (rd == DT_F64 && int_like(ld)))
        throw_error(QGPU_ERR_UNSUPPORTED, "join key type is not on the accelerated path");
      throw_error(QGPU_ERR_ARROW, "Invalid argument error: Invalid comparison operation: join keys of different types");
    }
  }
  const Column& lk = left.cols[keys[0].first];
  const Column& rk = right.cols[keys[0].second];
  // several key columns, a Utf8 key, RIGHT / FULL: the general index-pair join
  if (keys.size() != 1 || lk.dtype == DT_UTF8 || join_type == QGPU_RIGHT || join_type == QGPU_FULL)
    return join_general(c, left, right, keys, join_type);
  const bool left_join = join_type == QGPU_LEFT;
  const int64_t nl = left.n_rows, nr = right.n_rows;
  HostMarks hm(c);
  const int lki = keys[0].first, rki = keys[0].second;

  // "simple" tables (fixed-width columns, no validity bitmaps) take the radix-partitioned + fused-emit path
  auto is_simple = [](const Table& t) {
    for (const Column& col : t.cols)
      if (dtype_width(col.dtype) == 0 || col.valid != nullptr) return false;
    return (int)t.cols.size() <= MAX_JOIN_COLS;
  };
  const bool simple = is_simple(left) && is_simple(right);

  // ---- fast path: radix-partitioned join with region tables that stay L2-resident ----
  if (simple && lk.dtype == DT_I64 && rk.dtype == DT_I64 && c->opt("join.radix", 1) != 0 && nl > 0 &&
      nr >= c->opt("join.radix_min_build_rows", 1 << 19) && (int)left.cols.size() <= PJ_MAX_COLS &&
      (int)right.cols.size() <= PJ_MAX_COLS) {
    // dense (direct-addressed) layout first when the key range allows it, the hashed layout when that bails out
    // (a skewed region), the generic path below when that bails out too (duplicate build keys)
    std::shared_ptr<Table> out;
    bool empty = false, was_dense = false;
    // second generation when every column is 8 bytes wide and both sides are narrow enough to be staged whole
    auto all8 = [](const Table& t) {
      for (const Column& col : t.cols)
        if (dtype_width(col.dtype) != 8) return false;
      return (int)t.cols.size() <= RS_MAX_COLS;
    };
    auto pj = (c->opt("join.gen2", 1) != 0 && all8(left) && all8(right)) ? pjoin2 : pjoin;
    if (pj(c, left, right, lki, rki, left_join, c->opt("join.dense", 1) != 0, hm, out, empty, was_dense)) return empty ? nullptr : out;
    if (was_dense && pj(c, left, right, lki, rki, left_join, false, hm, out, empty, was_dense)) return empty ? nullptr : out;
  }

  // ---- generic path: one flat table in HBM, count + emit passes (any key type, duplicate build keys, validity) ----
  uint64_t table_cap = 1024;
  while (table_cap < (uint64_t)nr * 2) table_cap <<= 1;
  const Table& L = left;
  const Table& Rt = right;
  const Column& lkc = L.cols[lki];
  const Column& rkc = Rt.cols[rki];

  // key range of the build side: a span below 2^32 takes the compact 8-byte slots (4 per 32-byte bucket)
  int64_t kmin = 0, kmax = -1;
  column_range(c, rkc, kmin, kmax);
  hm.mark("key min/max (+sync)");
  const bool compact = c->opt("join.compact", 1) != 0 && kmax >= kmin && (uint64_t)kmax - (uint64_t)kmin < 0xffffffffull;
  JoinTable jt;
  jt.cap = (uint32_t)table_cap;
  jt.compact = compact ? 1 : 0;
  jt.key_min = kmin;
  jt.key_max = kmax;
  DevBuf tab(c, jt.bytes() + 256);
  jt.slots = tab.p;
  hm.mark("alloc table");
  launch_join_build(c, colref_of(rkc), nr, jt);
  hm.mark("build launch");

  const int64_t n_tiles = std::max<int64_t>(1, join_tiles(nl));
  DevBuf first(c, pad_bytes((size_t)nl * 4));
  DevBuf tile_counts(c, (size_t)n_tiles * 4 + 256, true);
  DevBuf tile_off(c, (size_t)(n_tiles + 1) * 8 + 256);
  DevBuf dup(c, 256, true);
  launch_join_count(c, colref_of(lkc), nl, jt, left_join, first.as<uint32_t>(), tile_counts.as<uint32_t>(), dup.as<uint32_t>());
  scan_u32_to_i64(c, tile_counts.as<uint32_t>(), n_tiles, tile_off.as<int64_t>());
  int64_t n_out;
  bool has_dups;
  {
    int64_t* h = (int64_t*)c->scratch(16);
    h[1] = 0;
    QG_CUDA(cudaMemcpyAsync(h, tile_off.as<int64_t>() + n_tiles, 8, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaMemcpyAsync(h + 1, dup.p, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    n_out = h[0];
    has_dups = (uint32_t)h[1] != 0;
  }
  hm.mark("count + scan (+sync)");
  if (n_out == 0) return nullptr;  // no joined batch has rows
  if (n_out >= 0xfffffff0ll) throw_error(QGPU_ERR_UNSUPPORTED, "join result of 2^32 rows or more");

  auto out = new_table(c);
  out->n_rows = n_out;
  out->batch_offsets = {0, n_out};

  if (simple && !has_dups) {
    // ---- fused emit: output columns written straight from the probe pass's matches ----
    JoinEmitCols E;
    memset(&E, 0, sizeof E);
    E.n_left = (int)L.cols.size();
    E.n_right = (int)Rt.cols.size();
    for (int i = 0; i < E.n_left; ++i) {  // schema = left fields ++ right fields (executor.rs:532-537)
      Column d = alloc_column(c, L.cols[i].dtype, n_out, false);
      E.left[i] = JoinOutCol{L.cols[i].data, d.data, dtype_width(d.dtype), 0};
      out->fields.push_back(left.fields[i]);
      out->cols.push_back(d);
    }
    DevBuf vbytes;
    if (left_join) {
      vbytes = DevBuf(c, pad_bytes((size_t)n_out));
      E.right_valid_bytes = vbytes.as<uint8_t>();
    }
    for (int i = 0; i < E.n_right; ++i) {
      Column d = alloc_column(c, Rt.cols[i].dtype, n_out, left_join);
      E.right[i] = JoinOutCol{Rt.cols[i].data, d.data, dtype_width(d.dtype), 0};
      Field f = right.fields[i];
      if (left_join) f.nullable = true;
      out->fields.push_back(f);
      out->cols.push_back(d);
    }
    hm.mark("alloc outputs");
    launch_join_emit_fused(c, nl, left_join, first.as<uint32_t>(), tile_off.as<int64_t>(), E);
    hm.mark("emit launch");
    if (left_join && E.n_right > 0) {
      Column& c0 = out->cols[(size_t)E.n_left];
      launch_bytes_to_bits(c, vbytes.as<uint8_t>(), n_out, c0.valid);
      for (int i = 1; i < E.n_right; ++i)
        QG_CUDA(cudaMemcpyAsync(out->cols[(size_t)E.n_left + i].valid, c0.valid, (size_t)((n_out + 31) / 32) * 4,
                                cudaMemcpyDeviceToDevice, c->stream));
    }
    return out;
  }

  // ---- generic: index pairs, then column gathers (any column layout, duplicate build keys) ----
  DevBuf out_left(c, pad_bytes((size_t)n_out * 4)), out_right(c, pad_bytes((size_t)n_out * 4));
  launch_join_emit(c, colref_of(lkc), nl, jt, left_join, has_dups, first.as<uint32_t>(), tile_off.as<int64_t>(),
                   out_left.as<uint32_t>(), out_right.as<uint32_t>());
  std::vector<int> lcols(L.cols.size()), rcols(Rt.cols.size());
  for (size_t i = 0; i < lcols.size(); ++i) lcols[i] = (int)i;
  for (size_t i = 0; i < rcols.size(); ++i) rcols[i] = (int)i;
  auto lt = gather_rows(c, L, lcols, out_left.as<uint32_t>(), n_out, false);
  auto rt = gather_rows(c, Rt, rcols, out_right.as<uint32_t>(), n_out, left_join);
  for (size_t i = 0; i < lt->cols.size(); ++i) {  // schema = left fields ++ right fields (executor.rs:532-537)
    out->fields.push_back(lt->fields[i]);
    out->cols.push_back(lt->cols[i]);
    lt->cols[i].owned = false;
    lt->cols[i].owned_valid = false;
  }
  for (size_t i = 0; i < rt->cols.size(); ++i) {
    out->fields.push_back(rt->fields[i]);
    out->cols.push_back(rt->cols[i]);
    rt->cols[i].owned = false;
    rt->cols[i].owned_valid = false;
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// Multi-GPU join: the radix scatter is the exchange (query-distributed/src/planner.rs:228-249: "shuffle both sides on
// the join key, join locally")
// ----------------------------------------------------------------------------------------------
// Every rank holds a row shard of both inputs.  The regions of the radix join are spread over the GPUs (rank r owns
// regions [r * F, (r + 1) * F)); k_pj_scatter<MULTI> writes every row - and the atomicAdd that reserves its place -
// straight into the owner's exchange arena over NVLink, two stream-ordered barriers fence the phase, and every rank
// runs k_pj_join on the regions it owns.  The output stays sharded (Merge::Concat semantics).  All decisions below
// are taken from all-gathered numbers, so every rank takes the same branch.
std::shared_ptr<Table> op_join_exchange(Ctx* c, const Table& left, const Table& right, int join_type, const qgpu_expr* on, bool* done) {
  *done = false;
  if (!c->nccl_comm || c->world < 2 || c->world > PJ_MAX_WORLD) return nullptr;
  if (join_type != QGPU_INNER && join_type != QGPU_LEFT) return nullptr;
  if (!on || c->opt("join.exchange", 1) == 0) return nullptr;
  ExprTree t = parse_postfix(on);
  std::vector<std::pair<int, int>> keys;
  const int ncl = (int)left.cols.size(), ncr = (int)right.cols.size();
  if (!equi_keys(t, t.root, ncl, keys) || keys.size() != 1) return nullptr;
  const int lki = keys[0].first, rki = keys[0].second;
  if (rki >= ncr || ncl > PJ_MAX_COLS || ncr > PJ_MAX_COLS) return nullptr;
  auto simple = [](const Table& tb) {
    for (const Column& col : tb.cols)
      if (dtype_width(col.dtype) == 0 || col.valid != nullptr) return false;
    return true;
  };
  // (schemas are the same on every rank; validity bitmaps may differ, hence the all-gathered `ok` below)
  const bool ok_local = simple(left) && simple(right) && left.cols[(size_t)lki].dtype == DT_I64 && right.cols[(size_t)rki].dtype == DT_I64;
  const bool left_join = join_type == QGPU_LEFT;
  const int W = c->world, rank = c->rank;
  HostMarks hm(c);

  int64_t kmin = 0, kmax = -1;
  if (ok_local && right.n_rows > 0) column_range(c, right.cols[(size_t)rki], kmin, kmax);
  const std::vector<int64_t> all = allgather_i64(c, {ok_local ? 1 : 0, left.n_rows, right.n_rows, kmin, kmax});
  int64_t nl = 0, nr = 0, gmin = INT64_MAX, gmax = INT64_MIN;
  for (int r = 0; r < W; ++r) {
    if (!all[(size_t)r * 5]) return nullptr;
    nl += all[(size_t)r * 5 + 1];
    nr += all[(size_t)r * 5 + 2];
    if (all[(size_t)r * 5 + 2] > 0) {
      gmin = std::min(gmin, all[(size_t)r * 5 + 3]);
      gmax = std::max(gmax, all[(size_t)r * 5 + 4]);
    }
  }
  if (nl == 0 || nr < c->opt("join.radix_min_build_rows", 1 << 19)) return nullptr;
  hm.mark("exchange join: totals");

  // ---- regions: F per rank, W * F <= PJ_MAX_PARTS in all ----
  const uint64_t region_bytes = (uint64_t)c->opt("join.region_bytes", 16 << 20);
  const uint32_t bucket_factor = (uint32_t)std::max<int64_t>(1, c->opt("join.bucket_factor", 2));
  PJMap map;
  memset(&map, 0, sizeof map);
  uint32_t F = 1, n_entries = 0;
  const uint64_t range = gmax >= gmin ? (uint64_t)gmax - (uint64_t)gmin + 1 : 0;
  if (c->opt("join.dense", 1) != 0 && range != 0 && range <= (uint64_t)nr * (uint64_t)c->opt("join.dense_max_sparsity", 8) &&
      range < (1ull << 36)) {
    uint32_t shift = 1;
    while ((8ull << shift) < region_bytes) ++shift;
    while (((range - 1) >> shift) + 1 > PJ_MAX_PARTS) ++shift;
    const uint32_t regions = (uint32_t)(((range - 1) >> shift) + 1);
    F = (regions + W - 1) / W;
    while ((uint64_t)F * W > PJ_MAX_PARTS) {  // rounding F up may overshoot: coarser regions
      ++shift;
      F = ((uint32_t)(((range - 1) >> shift) + 1) + W - 1) / W;
    }
    map.dense = 1;
    map.shift = shift;
    map.kmin = gmin;
    map.range = range;
    n_entries = 1u << shift;
  } else {
    if ((W & (W - 1)) != 0) return nullptr;  // hashed regions are a power of two: so must the world be
    const uint64_t region_bytes_h = (uint64_t)c->opt("join.region_bytes_hashed", 16 << 20);
    uint32_t parts = (uint32_t)W;
    while (parts < PJ_MAX_PARTS && (uint64_t)nr * 34 * bucket_factor / parts > region_bytes_h) parts <<= 1;
    F = parts / (uint32_t)W;
    while ((1u << map.log_parts) < parts) ++map.log_parts;
  }
  map.parts = F * (uint32_t)W;
  auto region_cap = [&](int64_t n, uint32_t mult) {
    uint64_t cap = (uint64_t)n / (map.dense ? std::max<uint64_t>(1, ((map.range - 1) >> map.shift) + 1) : map.parts);
    cap += cap / 8 + 1024;
    return (uint32_t)((cap + mult - 1) / mult * mult);
  };
  const uint32_t cap_l = region_cap(nl, PJ_TILE), cap_r = region_cap(nr, 32);
  if ((uint64_t)map.parts * cap_l >= 0xfff00000ull || (uint64_t)map.parts * cap_r >= 0xfff00000ull) return nullptr;
  if (!map.dense) n_entries = cap_r * bucket_factor;

  // ---- arena layout (identical on every rank): control block, then the region-major copies of both sides ----
  constexpr size_t CTR_WORDS = (size_t)PJ_MAX_PARTS * PJ_CTR_STRIDE;
  const size_t ctl_bytes = (4 * CTR_WORDS + 8 * 64) * 4;
  size_t off = (ctl_bytes + 255) / 256 * 256;
  std::vector<size_t> off_r((size_t)ncr), off_l((size_t)ncl);
  for (int i = 0; i < ncr; ++i) {
    off_r[(size_t)i] = off;
    off += pad_bytes((size_t)F * cap_r * dtype_width(right.cols[(size_t)i].dtype));
  }
  for (int i = 0; i < ncl; ++i) {
    off_l[(size_t)i] = off;
    off += pad_bytes((size_t)F * cap_l * dtype_width(left.cols[(size_t)i].dtype));
  }
  arena_ensure(c, off);
  hm.mark("exchange join: arena");
  char* A0 = (char*)c->arena;
  uint32_t* cnt_l = (uint32_t*)A0;
  uint32_t* cnt_r = cnt_l + CTR_WORDS;
  uint32_t* cleared = cnt_l + 2 * CTR_WORDS;
  uint32_t* inserted = cnt_l + 3 * CTR_WORDS;
  uint32_t* singles = cnt_l + 4 * CTR_WORDS;
  uint32_t* flags = singles;
  unsigned long long* out_cursor = reinterpret_cast<unsigned long long*>(singles + 64);
  uint32_t* ticket = singles + 128;

  // ---- phase 1: everybody clears its control block; barrier; everybody scatters both sides to the owners; barrier ----
  QG_CUDA(cudaMemsetAsync(A0, 0, ctl_bytes, c->stream));
  stream_barrier(c);
  auto scatter = [&](const Table& tb, int kc, uint32_t cap, size_t cnt_word_off, const std::vector<size_t>& offs, const PJMap& m,
                     const char* name) {
    PJCols pc;
    memset(&pc, 0, sizeof pc);
    pc.n = (int)tb.cols.size();
    pc.key_col = kc;
    PJPeers pe;
    memset(&pe, 0, sizeof pe);
    pe.world = W;
    pe.regions_per_rank = F;
    for (int r = 0; r < W; ++r) {
      pe.cursor[r] = (uint32_t*)c->arena_peer[r] + cnt_word_off;
      for (int i = 0; i < pc.n; ++i) pe.out[r][i] = (char*)c->arena_peer[r] + offs[(size_t)i];
    }
    for (int i = 0; i < pc.n; ++i) pc.c[i] = PJCol{tb.cols[(size_t)i].data, nullptr, dtype_width(tb.cols[(size_t)i].dtype), 0};
    // (flags: the writer's own word; they are OR-ed over the ranks at the end)
    launch_pj_scatter(c, (const int64_t*)tb.cols[(size_t)kc].data, tb.n_rows, m, cap, nullptr, flags, pc, &pe, name);
  };
  PJMap lmap = map;
  lmap.drop_outside = left_join ? 0 : 1;
  scatter(right, rki, cap_r, CTR_WORDS, off_r, map, "k_pj_scatter_build_nvlink");
  scatter(left, lki, cap_l, 0, off_l, lmap, "k_pj_scatter_probe_nvlink");
  stream_barrier(c);
  hm.mark("exchange join: scatter launches");

  // ---- phase 2: the local join over the F regions this rank owns ----
  PJMap mine = lmap;
  mine.parts = F;
  if (map.dense) {
    const uint64_t first = (uint64_t)rank * F << map.shift;  // first key (offset) of this rank's regions
    mine.kmin = (int64_t)((uint64_t)map.kmin + first);
    mine.range = first < map.range ? std::min<uint64_t>(map.range - first, (uint64_t)F << map.shift) : 0;
  }
  int inline_right = -1;
  if (ncr == 2 && dtype_width(right.cols[(size_t)(1 - rki)].dtype) == 8) inline_right = 1 - rki;
  const size_t table_bytes = (size_t)F * n_entries * (map.dense ? 8 : sizeof(PJBucket));
  DevBuf tab(c, table_bytes + 256);
  // upper bound of this rank's output rows: the probe rows that arrived (cursor sum) - read after the join; allocate
  // for the capacity of its regions
  const int64_t out_cap = (int64_t)F * cap_l;
  auto out = new_table(c);
  PJEmit E;
  memset(&E, 0, sizeof E);
  E.n_left = ncl;
  E.n_right = ncr;
  E.left_key = lki;
  E.right_key = rki;
  E.inline_right = inline_right;
  E.left_join = left_join ? 1 : 0;
  E.prefetch_left = -1;
  for (int i = 0; i < ncl && E.prefetch_left < 0; ++i)
    if (i != lki && dtype_width(left.cols[(size_t)i].dtype) == 8) E.prefetch_left = i;
  for (int i = 0; i < ncl; ++i) {
    Column d = alloc_column(c, left.cols[(size_t)i].dtype, out_cap, false);
    E.left[i] = PJCol{A0 + off_l[(size_t)i], d.data, dtype_width(d.dtype), 0};
    out->fields.push_back(left.fields[(size_t)i]);
    out->cols.push_back(d);
  }
  DevBuf vbytes;
  if (left_join) {
    vbytes = DevBuf(c, pad_bytes((size_t)out_cap));
    E.right_valid_bytes = vbytes.as<uint8_t>();
  }
  for (int i = 0; i < ncr; ++i) {
    Column d = alloc_column(c, right.cols[(size_t)i].dtype, out_cap, left_join);
    E.right[i] = PJCol{A0 + off_r[(size_t)i], d.data, dtype_width(d.dtype), 0};
    Field f = right.fields[(size_t)i];
    if (left_join) f.nullable = true;
    out->fields.push_back(f);
    out->cols.push_back(d);
  }
  PJJoinArgs JA;
  memset(&JA, 0, sizeof JA);
  JA.build_key = (const int64_t*)(A0 + off_r[(size_t)rki]);
  JA.build_val = inline_right >= 0 ? (const uint64_t*)(A0 + off_r[(size_t)inline_right]) : nullptr;
  JA.probe_key = (const int64_t*)(A0 + off_l[(size_t)lki]);
  JA.count_r = cnt_r;
  JA.count_l = cnt_l;
  JA.table = tab.p;
  JA.cap_r = cap_r;
  JA.cap_l = cap_l;
  JA.n_entries = n_entries;
  JA.flags = flags;
  JA.cleared = cleared;
  JA.inserted = inserted;
  JA.out_cursor = out_cursor;
  JA.ticket = ticket;
  JA.timing = nullptr;
  launch_pj_join(c, mine, JA, E);
  hm.mark("exchange join: join launch");

  uint32_t* h = (uint32_t*)c->scratch(16);
  QG_CUDA(cudaMemcpyAsync(h, flags, 4, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaMemcpyAsync(h + 2, out_cursor, 8, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));
  const uint32_t fl = h[0];
  const int64_t n_out = (int64_t)(((uint64_t)h[3] << 32) | h[2]);
  // a bail-out anywhere (overflowing region, duplicate build key, marker value) sends every rank to the other path
  const std::vector<int64_t> all_fl = allgather_i64(c, {(int64_t)fl});
  for (int r = 0; r < W; ++r)
    if (all_fl[(size_t)r]) return nullptr;
  *done = true;
  if (n_out == 0) return nullptr;
  out->n_rows = n_out;
  out->batch_offsets = {0, n_out};
  for (Column& col : out->cols) col.n = n_out;
  if (left_join && ncr > 0) {
    Column& c0 = out->cols[(size_t)ncl];
    launch_bytes_to_bits(c, vbytes.as<uint8_t>(), n_out, c0.valid);
    for (int i = 1; i < ncr; ++i)
      QG_CUDA(cudaMemcpyAsync(out->cols[(size_t)ncl + i].valid, c0.valid, (size_t)((n_out + 31) / 32) * 4, cudaMemcpyDeviceToDevice,
                              c->stream));
  }
  return out;
}

}  // namespace qgpu
