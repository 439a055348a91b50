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

// the columns that travel: the key first, then whatever else the program reads; empty = not eligible
static std::vector<int> carry_columns(const Table& in, const AggPlan& ap) {
  if (ap.key_cols.size() != 1) return {};
  const int kc = ap.key_cols[0];
  if (in.cols[(size_t)kc].dtype != DT_I64 || in.cols[(size_t)kc].valid != nullptr) return {};
  std::vector<int> carry{kc};
  for (int tc : ap.prog.col_of_slot)
    if (std::find(carry.begin(), carry.end(), tc) == carry.end()) carry.push_back(tc);
  if ((int)carry.size() > RS_MAX_COLS) return {};
  for (int tc : carry)
    if (dtype_width(in.cols[(size_t)tc].dtype) != 8 || in.cols[(size_t)tc].valid != nullptr) return {};
  return carry;
}

// where the rows of the two local radix levels come from
struct PartInput {
  const uint64_t* cols[RS_MAX_COLS];  // the carry columns
  int64_t n_phys = 0;                 // physical rows (gaps included)
  uint32_t in_cap = 0;                // region-major input with gaps (the exchange arena), or 0
  uint32_t in_stride = 0;
  const uint32_t* in_count = nullptr;
  uint64_t n_expected = 0;            // rows expected: the geometry is derived from it
  uint32_t skip_bits = 0;             // hash bits already consumed (the exchange: log2 world)
  const uint32_t* exchange_flags = nullptr;  // PJFlags raised by the exchange scatter (device word), or nullptr
};

// status of the local pipeline: 0 = ok (out / err filled), 1 = bail out (skew, table full)
static int run_partitioned(Ctx* c, const Table& in, const std::vector<int>& carry, const PartInput& I, const AggPlan& ap, int n_aggr,
                           std::shared_ptr<Table>* out_table, uint32_t* err_out) {
  const int kc = ap.key_cols[0];
  const Program& prog = ap.prog;
  const int n_acc = (int)prog.acc_kinds.size();
  const int nc = (int)carry.size();
  const int R = rows_per_thread(c);
  const uint32_t scan_tile = (uint32_t)SCAN_THREADS * (uint32_t)R;

  // ---- geometry: P1 regions x P2 sub-regions, a partition of ~target rows for a table of S slots ----
  uint32_t S = 1024;
  while (S > 256 && (size_t)(S + 1) * (n_acc + 1) * 8 > 72 * 1024) S >>= 1;  // three CTAs per SM
  if (c->opt("group.part_slots", 0) > 0) S = pow2_ceil((uint64_t)c->opt("group.part_slots", 0));
  const uint64_t n = std::max<uint64_t>(I.n_expected, 1);
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
  if (rows1 >= 0xfff00000ull || rows2 >= 0xfff00000ull) return 1;

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
    a1.in[i] = I.cols[i];
    a1.out[i] = l1.back().as<uint64_t>();
  }
  a1.n = I.n_phys;
  a1.M.dense = 3;
  a1.M.parts = P1;
  a1.M.log_parts = bits1;
  a1.M.skip_bits = I.skip_bits;
  a1.cap = cap1;
  a1.ctr_stride = STRIDE1;
  a1.cursor = cur1;
  a1.flags = flags;
  a1.in_cap = I.in_cap;
  a1.in_ctr_stride = I.in_stride;
  a1.in_count = I.in_count;
  a1.merge_in = 1;
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
  a2.M.skip_bits = I.skip_bits + bits1;
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
    v.owned_valid = false;
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
    P.gp_preload = c->opt("group.preload", 1) != 0 ? 1u : 0u;
    P.gp_out_cap = (uint32_t)std::min<uint64_t>(out_cap, 0xffffffffull);
    P.gp_out_keys = (uint64_t*)keys.data;
    P.gp_out_acc = acc.as<uint64_t>();
    P.gp_cursor = flags + 16;
    QG_CUDA(cudaMemsetAsync(flags + 16, 0, 4, c->stream));
    launch_scan(c, SK_GROUP_PART, R, P, "k_scan_filter_group_part", &p3);
    uint32_t* h = (uint32_t*)c->scratch(32);
    h[3] = 0;
    QG_CUDA(cudaMemcpyAsync(h, flags, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaMemcpyAsync(h + 1, P.err, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaMemcpyAsync(h + 2, flags + 16, 4, cudaMemcpyDeviceToHost, c->stream));
    if (I.exchange_flags) QG_CUDA(cudaMemcpyAsync(h + 3, I.exchange_flags, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    const uint32_t fl = h[0] | h[3], err = h[1];
    const int64_t ng = (int64_t)h[2];
    if (fl != 0 || (err & ERR_CAPACITY)) {  // a skewed partition, or more groups in a partition than its table holds
      free_column(c, keys);
      return 1;
    }
    if (err & ERR_OUTCAP) {
      free_column(c, keys);
      if (out_cap >= n + (n >> 2)) return 1;
      out_cap = n + (n >> 2);  // (the exchanged form only knows the EXPECTED rows of this rank)
      continue;
    }
    *err_out = err;
    if (ng == 0) {
      free_column(c, keys);
      *out_table = nullptr;
      return 0;
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
    return 0;
  }
  return 1;
}

bool group_by_partitioned(Ctx* c, const Table& in, const AggPlan& ap, int n_aggr, std::shared_ptr<Table>* out_table, uint32_t* err_out) {
  if (c->opt("group.radix", 1) == 0) return false;
  if (in.n_rows < c->opt("group.partition_min_rows", 1 << 22)) return false;
  const std::vector<int> carry = carry_columns(in, ap);
  if (carry.empty()) return false;
  PartInput I;
  for (size_t i = 0; i < carry.size(); ++i) I.cols[i] = (const uint64_t*)in.cols[(size_t)carry[i]].data;
  I.n_phys = in.n_rows;
  I.n_expected = (uint64_t)in.n_rows;
  return run_partitioned(c, in, carry, I, ap, n_aggr, out_table, err_out) == 0;
}

// ----------------------------------------------------------------------------------------------
// multi-GPU: exchange the rows by owner over NVLink, then the local pipeline (query-distributed/src/planner.rs:200-226
// would merge partial groups on one node; with millions of groups every rank keeps the keys it owns instead and the
// result stays partitioned by key - Merge::Concat finishes it)
// ----------------------------------------------------------------------------------------------
bool group_by_exchanged(Ctx* c, const Table& in, const AggPlan& ap, int n_aggr, int64_t total_rows, int64_t max_rows,
                        std::shared_ptr<Table>* out_table, uint32_t* err_out) {
  const int W = c->world;
  if (c->opt("group.radix", 1) == 0 || W < 2 || W > PJ_MAX_WORLD || (W & (W - 1)) != 0) return false;
  // eligibility must be the same on every rank: schemas are, validity bitmaps may not be
  const std::vector<int> carry = carry_columns(in, ap);
  {
    bool all = true;
    for (int64_t e : allgather_i64(c, {carry.empty() ? 0 : 1})) all &= e != 0;
    if (!all) return false;
  }
  const int nc = (int)carry.size();
  const uint32_t logW = log2_of((uint32_t)W);
  // a sender's segment of an owner's arena: its rows / W with slack
  uint64_t seg = (uint64_t)max_rows / W;
  seg += seg / 8 + 8 * (uint64_t)std::sqrt((double)seg) + 64;
  const uint32_t seg_cap = (uint32_t)((seg + RS_TILE - 1) / RS_TILE * RS_TILE);
  const uint64_t arena_rows = (uint64_t)W * seg_cap + RS_TILE;
  if (arena_rows >= 0xfff00000ull) return false;
  const size_t col_bytes = pad_bytes((size_t)arena_rows * 8);
  arena_ensure(c, (size_t)nc * col_bytes + 4096);
  char* A0 = (char*)c->arena;

  // local cursors: rows this rank has placed in its segment on every owner
  DevBuf ctl(c, ((size_t)W * PJ_CTR_STRIDE + 64 + (size_t)W * W + 2 * W) * 4, true);
  uint32_t* cur0 = ctl.as<uint32_t>();
  uint32_t* xflags = cur0 + (size_t)W * PJ_CTR_STRIDE;
  uint32_t* sent = xflags + 64;            // [W] compact copy of the cursors
  uint32_t* all_sent = sent + W;           // [W senders][W owners]
  uint32_t* arrived = all_sent + (size_t)W * W;  // [W] rows that arrived here from every sender
  stream_barrier(c);  // every arena exists and nobody still reads its previous content
  RSArgs a0;
  memset(&a0, 0, sizeof a0);
  RSPeers pe;
  memset(&pe, 0, sizeof pe);
  pe.world = W;
  pe.regions_per_rank = 1;
  for (int i = 0; i < nc; ++i) a0.in[i] = (const uint64_t*)in.cols[(size_t)carry[(size_t)i]].data;
  for (int r = 0; r < W; ++r) {
    pe.cursor[r] = cur0 + (size_t)r * PJ_CTR_STRIDE;  // LOCAL: a sender reserves rows in its own segment, no remote atomics
    for (int i = 0; i < nc; ++i) pe.out[r][i] = (uint64_t*)((char*)c->arena_peer[r] + (size_t)i * col_bytes);
  }
  a0.n = in.n_rows;
  a0.M.dense = 3;
  a0.M.parts = (uint32_t)W;
  a0.M.log_parts = logW;
  a0.M.skip_bits = 0;
  a0.cap = seg_cap;
  a0.ctr_stride = 0;  // (the owners' cursor pointers above already address one counter each)
  a0.flags = xflags;
  a0.first_base = (uint32_t)c->rank * seg_cap;
  a0.trash_row = (uint32_t)((uint64_t)W * seg_cap);
  launch_rscatter(c, a0, nc, &pe, "k_rscatter_groups_nvlink");
  // who sent how much to whom: all-gather the cursors on the device (this collective also orders "all rows have landed")
  QG_CUDA(cudaMemcpy2DAsync(sent, 4, cur0, (size_t)PJ_CTR_STRIDE * 4, 4, (size_t)W, cudaMemcpyDeviceToDevice, c->stream));
  allgather_device(c, sent, all_sent, (size_t)W * 4);
  QG_CUDA(cudaMemcpy2DAsync(arrived, 4, all_sent + c->rank, (size_t)W * 4, 4, (size_t)W, cudaMemcpyDeviceToDevice, c->stream));

  PartInput I;
  for (int i = 0; i < nc; ++i) I.cols[i] = (const uint64_t*)(A0 + (size_t)i * col_bytes);
  I.n_phys = (int64_t)W * seg_cap;
  I.in_cap = seg_cap;
  I.in_stride = 1;
  I.in_count = arrived;
  I.n_expected = (uint64_t)total_rows / W + 1;
  I.skip_bits = logW;
  I.exchange_flags = xflags;
  std::shared_ptr<Table> mine;
  uint32_t err = 0;
  const int status = run_partitioned(c, in, carry, I, ap, n_aggr, &mine, &err);
  // everybody leaves alike: one rank's bail-out (or error) is everybody's
  bool any_bail = false;
  uint32_t any_err = 0;
  for (int64_t v : allgather_i64(c, {((int64_t)status << 32) | (int64_t)err})) {
    any_bail |= (v >> 32) != 0;
    any_err |= (uint32_t)v;
  }
  if (any_bail) return false;
  *err_out = any_err;
  *out_table = mine;
  return true;
}

}  // namespace qgpu
