// ops_groupby.cu — high-cardinality GROUP BY on one Int64 key (BASELINE config 5: 10 M distinct keys, MIN / MAX / SUM)
// without a hash table in HBM.
//
// The HBM-table tier of op_aggregate (ops.cu, SK_GROUP_GLOBAL) pays ~8 L2 operations per row (probe, CAS, key compare,
// one atomic per accumulator) and sits at the L2 atomic rate: 8.4 ms for 100 M rows (profiles/r01h).  Here the rows are
// radix-partitioned twice on the bits of the key hash (rscatter_kernels.cu; 2^7..2^8 regions, then up to 2^8
// sub-regions each) until a partition holds a few thousand rows, i.e. a few hundred groups: the aggregation pass
// (scan_body.cuh SK_GROUP_PART) then gives every partition to one CTA, which aggregates it in a shared-memory table and
// appends the finished groups - all rows of a key are in one partition - to the output.  HBM sees three streaming
// passes; nothing random, no global atomics per row.
//
// Semantics are those of the grouped HashAggregate of SURVEY.md 8c Tier 2 (the reference's execute_aggregate returns
// Ok(vec![]) for GROUP BY, crates/query-executor/src/executor.rs:189): one output row per distinct key, aggregates per
// evaluate_aggregate (operators.rs:745-848), row order unspecified.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "qgpu_internal.h"
#include "rscatter_kernels.cuh"
#include "scan_kernels.cuh"

namespace qgpu {

namespace {
uint32_t pow2_ceil(uint64_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}
uint32_t log2_of(uint32_t p) {
  uint32_t l = 0;
  while ((1u << l) < p) ++l;
  return l;
}
}  // namespace

bool group_by_partitioned(Ctx* c, const Table& in, const AggPlan& ap, int n_aggr, std::shared_ptr<Table>* out_table, uint32_t* err_out) {
  if (c->opt("group.radix", 1) == 0 || ap.key_cols.size() != 1) return false;
  const int kc = ap.key_cols[0];
  const Program& prog = ap.prog;
  if (in.cols[(size_t)kc].dtype != DT_I64 || in.cols[(size_t)kc].valid != nullptr) return false;
  if (in.n_rows < c->opt("group.partition_min_rows", 1 << 22)) return false;
  // the columns that travel: the key first, then whatever else the program reads
  std::vector<int> carry{kc};
  for (int tc : prog.col_of_slot)
    if (std::find(carry.begin(), carry.end(), tc) == carry.end()) carry.push_back(tc);
  if ((int)carry.size() > RS_MAX_COLS) return false;
  for (int tc : carry)
    if (dtype_width(in.cols[(size_t)tc].dtype) != 8 || in.cols[(size_t)tc].valid != nullptr) return false;
  const int n_acc = (int)prog.acc_kinds.size();
  const int nc = (int)carry.size();
  const int R = rows_per_thread(c);
  const uint32_t scan_tile = (uint32_t)SCAN_THREADS * (uint32_t)R;

  // ---- geometry: P1 regions x P2 sub-regions, a partition of ~target rows for a table of S slots ----
  uint32_t S = 1024;
  while (S > 256 && (size_t)(S + 1) * (n_acc + 1) * 8 > 72 * 1024) S >>= 1;  // three CTAs per SM
  if (c->opt("group.part_slots", 0) > 0) S = pow2_ceil((uint64_t)c->opt("group.part_slots", 0));
  const uint64_t n = (uint64_t)in.n_rows;
  const uint64_t target = (uint64_t)c->opt("group.part_rows", 3 * (int64_t)S);
  uint32_t bits1 = GROUP_REGION_BITS;
  if (n / ((uint64_t)1 << bits1) / PJ_MAX_PARTS > target) bits1 = 8;
  if (c->opt("group.bits1", 0) > 0) bits1 = (uint32_t)c->opt("group.bits1", 0);
  const uint32_t P1 = 1u << bits1;
  uint32_t P2 = pow2_ceil(std::max<uint64_t>(1, n / P1 / target));
  if (P2 > PJ_MAX_PARTS) P2 = PJ_MAX_PARTS;
  if (c->opt("group.parts2", 0) > 0) P2 = pow2_ceil((uint64_t)c->opt("group.parts2", 0));
  auto cap_of = [](uint64_t mean, uint32_t mult) {
    uint64_t cap = mean + mean / 8 + 8 * (uint64_t)std::sqrt((double)mean) + 64;
    return (uint32_t)((cap + mult - 1) / mult * mult);
  };
  const uint32_t cap1 = cap_of(n / P1, RS_TILE);
  const uint32_t cap2 = cap_of(n / P1 / P2, std::max<uint32_t>(scan_tile, 1024));
  const uint64_t rows1 = (uint64_t)P1 * cap1 + RS_TILE, rows2 = (uint64_t)P1 * P2 * cap2 + RS_TILE;
  if (rows1 >= 0xfff00000ull || rows2 >= 0xfff00000ull) return false;

  constexpr uint32_t STRIDE1 = PJ_CTR_STRIDE, STRIDE2 = 8;
  DevBuf ctl(c, ((size_t)P1 * STRIDE1 + (size_t)P1 * P2 * STRIDE2 + 64) * 4, true);
  uint32_t* cur1 = ctl.as<uint32_t>();
  uint32_t* cur2 = cur1 + (size_t)P1 * STRIDE1;
  uint32_t* flags = cur2 + (size_t)P1 * P2 * STRIDE2;  // [0] PJFlags, [16] group cursor

  // ---- level 1 ----
  std::vector<DevBuf> l1, l2;
  RSArgs a1;
  memset(&a1, 0, sizeof a1);
  for (int i = 0; i < nc; ++i) {
    l1.emplace_back(c, pad_bytes((size_t)rows1 * 8));
    a1.in[i] = (const uint64_t*)in.cols[(size_t)carry[(size_t)i]].data;
    a1.out[i] = l1.back().as<uint64_t>();
  }
  a1.n = in.n_rows;
  a1.M.dense = 2;
  a1.M.parts = P1;
  a1.M.log_parts = bits1;
  a1.cap = cap1;
  a1.ctr_stride = STRIDE1;
  a1.cursor = cur1;
  a1.flags = flags;
  a1.trash_row = (uint32_t)((uint64_t)P1 * cap1);
  launch_rscatter(c, a1, nc, nullptr, "k_rscatter_groups_l1");

  // ---- level 2: the rows of region g go to sub-regions g * P2 + [0, P2) ----
  RSArgs a2;
  memset(&a2, 0, sizeof a2);
  for (int i = 0; i < nc; ++i) {
    l2.emplace_back(c, pad_bytes((size_t)rows2 * 8));
    a2.in[i] = l1[(size_t)i].as<uint64_t>();
    a2.out[i] = l2.back().as<uint64_t>();
  }
  a2.n = (int64_t)P1 * cap1;
  a2.M.dense = 3;
  a2.M.parts = P2;
  a2.M.log_parts = log2_of(P2);
  a2.M.skip_bits = bits1;
  a2.cap = cap2;
  a2.ctr_stride = STRIDE2;
  a2.cursor = cur2;
  a2.flags = flags;
  a2.in_cap = cap1;
  a2.in_ctr_stride = STRIDE1;
  a2.in_count = cur1;
  a2.trash_row = (uint32_t)((uint64_t)P1 * P2 * cap2);
  if (P2 > 1) {
    launch_rscatter(c, a2, nc, nullptr, "k_rscatter_groups_l2");
    l1.clear();  // (stream-ordered: the blocks are reused only by later work on this stream)
  }
  const std::vector<DevBuf>& parted = P2 > 1 ? l2 : l1;
  const uint32_t cap_p = P2 > 1 ? cap2 : cap1;
  const uint32_t* cnt_p = P2 > 1 ? cur2 : cur1;
  const uint32_t stride_p = P2 > 1 ? STRIDE2 : STRIDE1;
  const uint64_t n_parts = P2 > 1 ? (uint64_t)P1 * P2 : P1;
  if (P2 == 1) l2.clear();

  // ---- aggregation pass over the partition-major copy ----
  Table view;  // (no context: owns nothing)
  view.fields = in.fields;
  view.n_rows = (int64_t)(n_parts * cap_p);
  view.batch_offsets = {0, view.n_rows};
  view.cols.resize(in.cols.size());
  for (int i = 0; i < nc; ++i) {
    Column v;
    v.dtype = in.cols[(size_t)carry[(size_t)i]].dtype;
    v.n = view.n_rows;
    v.data = parted[(size_t)i].p;
    v.owned = false;
    view.cols[(size_t)carry[(size_t)i]] = v;
  }
  Program p3 = prog;
  const int kslot = p3.slot_for(kc);

  uint64_t out_cap = std::min<uint64_t>(n, std::max<uint64_t>(n / 4, (uint64_t)1 << 20));
  if (c->opt("group.out_cap", 0) > 0) out_cap = (uint64_t)c->opt("group.out_cap", 0);
  for (int attempt = 0; attempt < 2; ++attempt) {
    Column keys = alloc_column(c, DT_I64, (int64_t)out_cap, false);
    DevBuf acc(c, (size_t)out_cap * n_acc * 8 + 256);
    ScanParams P;
    DeviceProgram dp;
    bind_program(c, view, p3, P, dp);
    P.n_pred = ap.n_pred;
    P.n_keys = 1;
    P.key_slots[0] = kslot;
    P.region_cap = cap_p;
    P.region_count = cnt_p;
    P.region_count_stride = stride_p;
    P.part_regions = 1;
    P.part_slots = S;
    P.gp_out_cap = (uint32_t)std::min<uint64_t>(out_cap, 0xffffffffull);
    P.gp_out_keys = (uint64_t*)keys.data;
    P.gp_out_acc = acc.as<uint64_t>();
    P.gp_cursor = flags + 16;
    QG_CUDA(cudaMemsetAsync(flags + 16, 0, 4, c->stream));
    launch_scan(c, SK_GROUP_PART, R, P, "k_scan_filter_group_part", &p3);
    uint32_t* h = (uint32_t*)c->scratch(16);
    QG_CUDA(cudaMemcpyAsync(h, flags, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaMemcpyAsync(h + 1, P.err, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaMemcpyAsync(h + 2, flags + 16, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    const uint32_t fl = h[0], err = h[1];
    const int64_t ng = (int64_t)h[2];
    if (fl != 0 || (err & ERR_CAPACITY)) {  // a skewed partition, or more groups in a partition than its table holds
      free_column(c, keys);
      return false;
    }
    if (err & ERR_OUTCAP) {
      free_column(c, keys);
      if (out_cap >= n) return false;
      out_cap = n;
      continue;
    }
    *err_out = err;
    if (ng == 0) {
      free_column(c, keys);
      *out_table = nullptr;
      return true;
    }
    auto out = new_table(c);
    out->n_rows = ng;
    out->batch_offsets = {0, ng};
    Field f = in.fields[(size_t)kc];
    f.name = ap.key_names[0];
    f.nullable = true;
    keys.n = ng;
    out->fields.push_back(f);
    out->cols.push_back(keys);
    add_agg_outputs(c, prog, n_aggr, acc.as<uint64_t>(), nullptr, ng, *out);
    *out_table = out;
    return true;
  }
  return false;
}

}  // namespace qgpu
