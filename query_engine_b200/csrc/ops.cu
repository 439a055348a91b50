// ops.cu — operator orchestration: the bodies of QueryExecutor::execute_filter / execute_aggregate /
// execute_join / execute_projection (crates/query-executor/src/executor.rs:93-190, 343-399) and of
// Partitioner::partition_by_hash (crates/query-distributed/src/partition.rs:151-212) on device tables.
#include <algorithm>
#include <cstring>

#include "exchange_kernels.cuh"
#include "interp.cuh"
#include "join_kernels.cuh"
#include "pjoin_kernels.cuh"
#include "qgpu_internal.h"
#include "scan_kernels.cuh"
#include "util_kernels.cuh"

namespace qgpu {

namespace {

ColRef colref_of(const Column& col) {
  ColRef r;
  r.data = col.data;
  r.valid = col.valid;
  r.chars = col.chars;
  r.dtype = col.dtype;
  r.pad = 0;
  return r;
}

int rows_per_thread(Ctx* c) {
  const int64_t r = c->opt("scan.rows_per_thread", 4);
  return r == 8 ? 8 : 4;
}

// upload a program and fill the column table of the scan parameters
struct DeviceProgram {
  DevBuf code;
  DevBuf err;
};

void bind_program(Ctx* c, const Table& in, const Program& prog, ScanParams& P, DeviceProgram& dp) {
  memset(&P, 0, sizeof P);
  P.n_rows = in.n_rows;
  P.n_insn = (int)prog.code.size();
  P.n_cols = (int)prog.col_of_slot.size();
  P.stack_depth = prog.max_stack;
  P.l2_prefetch = (int)c->opt("scan.l2_prefetch", 1);
  for (int i = 0; i < P.n_cols; ++i) {
    const Column& col = in.cols[prog.col_of_slot[i]];
    P.cols[i] = colref_of(col);
    P.col_width[i] = dtype_width(col.dtype);
  }
  const size_t bytes = prog.code.size() * sizeof(Insn);
  dp.code = DevBuf(c, bytes ? bytes : 16);
  // pageable source: the runtime stages the (tiny) program before returning, so no synchronisation is needed
  if (bytes) QG_CUDA(cudaMemcpyAsync(dp.code.p, prog.code.data(), bytes, cudaMemcpyHostToDevice, c->stream));
  dp.err = DevBuf(c, 256, true);
  P.prog = dp.code.as<Insn>();
  P.err = dp.err.as<uint32_t>();
  P.n_acc = (int)prog.acc_kinds.size();
  for (size_t i = 0; i < prog.acc_kinds.size(); ++i) P.acc_kinds[i] = prog.acc_kinds[i];
}

uint32_t read_u32(Ctx* c, const uint32_t* d) {
  uint32_t* h = (uint32_t*)c->scratch(8);
  QG_CUDA(cudaMemcpyAsync(h, d, 4, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));
  return *h;
}

void raise_device_error(uint32_t err) {
  // arrow-rs texts: ArrowError::DivideByZero -> "Divide by zero error"; ComputeError("Overflow happened on: ..")
  if (err & ERR_DIVZERO) throw_error(QGPU_ERR_ARROW, "Divide by zero error");
  if (err & ERR_OVERFLOW) throw_error(QGPU_ERR_ARROW, "Compute error: Overflow happened on: checked integer arithmetic");
  if (err & ERR_SPIN) throw_error(QGPU_ERR_CUDA, "device spin bound exceeded");
}

std::shared_ptr<Table> new_table(Ctx* c) {
  auto t = std::make_shared<Table>();
  t->ctx = c;
  t->batch_offsets.push_back(0);
  return t;
}

Column borrow(const Column& col) {
  Column b = col;
  b.owned = false;
  return b;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// gather_rows: out[i] = in[idx[i]] for the listed columns (all Arrow layouts of the path)
// ----------------------------------------------------------------------------------------------
std::shared_ptr<Table> gather_rows(Ctx* c, const Table& in, const std::vector<int>& cols, const uint32_t* d_idx, int64_t n_out,
                                   bool idx_may_be_null) {
  auto out = new_table(c);
  out->n_rows = n_out;
  out->batch_offsets = {0, n_out};
  std::vector<GatherCol> gc;
  for (int ci : cols) {
    const Column& src = in.cols[ci];
    Field f = in.fields[ci];
    const bool need_valid = src.valid != nullptr || idx_may_be_null;
    if (idx_may_be_null) f.nullable = true;
    Column dst;
    if (src.dtype == DT_UTF8) {
      dst.dtype = DT_UTF8;
      dst.n = n_out;
      dst.owned = true;
      DevBuf lens(c, pad_bytes((size_t)(n_out + 1) * 8));
      DevBuf offs(c, pad_bytes((size_t)(n_out + 1) * 8));
      launch_gather_utf8_len(c, (const int32_t*)src.data, d_idx, n_out, lens.as<int64_t>());
      scan_i64_to_i64(c, lens.as<int64_t>(), n_out, offs.as<int64_t>());
      int64_t* h = (int64_t*)c->scratch(8);
      QG_CUDA(cudaMemcpyAsync(h, offs.as<int64_t>() + n_out, 8, cudaMemcpyDeviceToHost, c->stream));
      QG_CUDA(cudaStreamSynchronize(c->stream));
      const int64_t n_chars = *h;
      if (n_chars > 0x7ffffff0ll) throw_error(QGPU_ERR_UNSUPPORTED, "Utf8 result exceeds 2 GiB");
      dst.data = c->alloc(pad_bytes((size_t)(n_out + 1) * 4));
      dst.chars = (uint8_t*)c->alloc(pad_bytes((size_t)n_chars));
      dst.n_chars = n_chars;
      launch_offsets_i64_to_i32(c, offs.as<int64_t>(), n_out + 1, (int32_t*)dst.data);
      launch_gather_utf8_copy(c, (const int32_t*)src.data, src.chars, d_idx, n_out, (const int32_t*)dst.data, dst.chars);
      if (need_valid) {
        dst.valid = (uint32_t*)c->zalloc(bitmap_bytes(n_out));
        GatherCol g{};
        g.in = nullptr;
        g.out = nullptr;
        g.kind = -1;
        g.in_valid = src.valid;
        g.out_valid = dst.valid;
        gc.push_back(g);
      }
    } else if (src.dtype == DT_NULL) {
      dst.dtype = DT_NULL;
      dst.n = n_out;
    } else {
      dst = alloc_column(c, src.dtype, n_out, need_valid);
      GatherCol g{};
      g.in = src.data;
      g.out = dst.data;
      g.in_valid = src.valid;
      g.out_valid = dst.valid;
      g.kind = src.dtype == DT_BOOL ? CK_BITS : (dtype_width(src.dtype) == 8 ? CK_W8 : CK_W4);
      gc.push_back(g);
    }
    out->fields.push_back(f);
    out->cols.push_back(dst);
  }
  if (!gc.empty()) launch_gather(c, d_idx, n_out, gc.data(), (int)gc.size());
  return out;
}

// ----------------------------------------------------------------------------------------------
// Filter (executor.rs:131-155): per input batch, keep rows whose predicate is true and valid;
// row order kept; batches that end up empty are dropped.
// ----------------------------------------------------------------------------------------------
std::shared_ptr<Table> op_filter(Ctx* c, const Table& in, const qgpu_expr* pred) {
  auto out = new_table(c);
  out->fields = in.fields;
  const int64_t n_batches = (int64_t)in.batch_offsets.size() - 1;
  if (n_batches == 0) {
    out->cols.resize(in.cols.size());
    for (size_t i = 0; i < in.cols.size(); ++i) out->cols[i].dtype = in.cols[i].dtype;
    return out;  // nothing is evaluated for zero batches, not even the predicate's types
  }
  ExprTree tree = parse_postfix(pred);
  Program prog;
  compile_predicate(tree, schema_view(in), prog);

  const int R = rows_per_thread(c);
  const int tile_rows = SCAN_THREADS * R;
  const int64_t n = in.n_rows;
  const int64_t n_tiles = std::max<int64_t>(1, (n + tile_rows - 1) / tile_rows);

  ScanParams P;
  DeviceProgram dp;
  bind_program(c, in, prog, P, dp);
  DevBuf mask(c, (size_t)n_tiles * tile_rows / 8 + 256);
  DevBuf tile_counts(c, (size_t)n_tiles * 4 + 256);
  DevBuf tile_off(c, (size_t)(n_tiles + 1) * 8 + 256);
  P.mask = mask.as<uint32_t>();
  P.tile_counts = tile_counts.as<uint32_t>();
  launch_scan(c, SK_MASK, R, P, "k_scan_filter_mask", &prog);
  scan_u32_to_i64(c, tile_counts.as<uint32_t>(), n_tiles, tile_off.as<int64_t>());

  // survivors before each batch boundary -> output batch offsets
  const int nbnd = (int)in.batch_offsets.size();
  DevBuf d_bnd(c, (size_t)nbnd * 8);
  DevBuf d_pre(c, (size_t)nbnd * 8);
  {
    int64_t* h = (int64_t*)c->scratch((size_t)nbnd * 8);
    memcpy(h, in.batch_offsets.data(), (size_t)nbnd * 8);
    QG_CUDA(cudaMemcpyAsync(d_bnd.p, h, (size_t)nbnd * 8, cudaMemcpyHostToDevice, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
  }
  launch_prefix_at(c, d_bnd.as<int64_t>(), nbnd, tile_rows, mask.as<uint32_t>(), tile_off.as<int64_t>(), d_pre.as<int64_t>());
  std::vector<int64_t> pre(nbnd);
  {
    int64_t* h = (int64_t*)c->scratch((size_t)nbnd * 8);
    QG_CUDA(cudaMemcpyAsync(h, d_pre.p, (size_t)nbnd * 8, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(pre.data(), h, (size_t)nbnd * 8);
  }
  raise_device_error(read_u32(c, P.err));
  const int64_t k = pre.back();
  out->n_rows = k;
  for (int i = 1; i < nbnd; ++i)
    if (pre[i] > pre[i - 1]) out->batch_offsets.push_back(pre[i]);  // empty result batches are skipped (:149-151)

  // compaction of every column (the reference's filter_record_batch copies all of them, :147)
  std::vector<CompactCol> cc;
  bool any_utf8 = false;
  out->cols.resize(in.cols.size());
  for (size_t ci = 0; ci < in.cols.size(); ++ci) {
    const Column& src = in.cols[ci];
    if (src.dtype == DT_UTF8) {
      any_utf8 = true;
      continue;
    }
    if (src.dtype == DT_NULL) {
      out->cols[ci].dtype = DT_NULL;
      out->cols[ci].n = k;
      continue;
    }
    Column dst = alloc_column(c, src.dtype, k, src.valid != nullptr);
    CompactCol x{};
    x.in = src.data;
    x.out = dst.data;
    x.in_valid = src.valid;
    x.out_valid = dst.valid;
    x.kind = src.dtype == DT_BOOL ? CK_BITS : (dtype_width(src.dtype) == 8 ? CK_W8 : CK_W4);
    cc.push_back(x);
    out->cols[ci] = dst;
  }
  DevBuf rowidx;
  if (any_utf8) {
    rowidx = DevBuf(c, pad_bytes((size_t)k * 4));
    CompactCol x{};
    x.out = rowidx.p;
    x.kind = CK_ROWIDX;
    cc.push_back(x);
  }
  if (k > 0 && !cc.empty()) launch_compact(c, n, tile_rows, mask.as<uint32_t>(), tile_off.as<int64_t>(), cc.data(), (int)cc.size());
  if (any_utf8) {
    std::vector<int> ucols;
    for (size_t ci = 0; ci < in.cols.size(); ++ci)
      if (in.cols[ci].dtype == DT_UTF8) ucols.push_back((int)ci);
    auto g = gather_rows(c, in, ucols, rowidx.as<uint32_t>(), k, false);
    for (size_t j = 0; j < ucols.size(); ++j) {
      out->cols[ucols[j]] = g->cols[j];
      g->cols[j].owned = false;  // ownership moves to `out`
    }
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// compile (Filter +) HashAggregate into one device program: [predicate section | OP_GATE | aggregate section]
// ----------------------------------------------------------------------------------------------
AggPlan plan_aggregate(const Table& in, const qgpu_expr* pred, const qgpu_expr* groups, int n_group, const qgpu_agg_expr* aggs,
                       int n_aggr) {
  AggPlan ap;
  Program& prog = ap.prog;
  std::vector<int>& key_cols = ap.key_cols;
  std::vector<std::string>& key_names = ap.key_names;
  const SchemaView sv = schema_view(in);
  if (pred) {
    ExprTree pt = parse_postfix(pred);
    compile_predicate(pt, sv, prog);
  } else {
    Insn gate{};
    gate.op = OP_GATE;
    prog.code.push_back(gate);
  }
  ap.n_pred = (int)prog.code.size();

  // group keys must be plain columns on the fused path
  for (int g = 0; g < n_group; ++g) {
    ExprTree gt = parse_postfix(&groups[g]);
    infer_types(gt, sv);
    const ExprNode& root = gt.nodes[gt.root];
    if (root.kind != QGPU_EXPR_COLUMN)
      throw_error(QGPU_ERR_UNSUPPORTED, "GROUP BY on a computed expression is not on the accelerated path yet");
    key_cols.push_back(root.op);
    key_names.push_back(!root.str.empty() ? root.str : in.fields[root.op].name);
  }

  std::vector<ExprTree> trees;
  std::vector<int> funcs;
  for (int i = 0; i < n_aggr; ++i) {
    trees.push_back(parse_postfix(&aggs[i].expr));
    funcs.push_back(aggs[i].func);
  }
  compile_aggregates(trees, funcs, sv, prog);
  // hidden counter of rows that passed the predicate: with zero survivors the unfused plan feeds the
  // aggregate zero batches and the reference returns Ok(vec![])
  int& pass_slot = ap.pass_slot;
  if (pred) {
    if ((int)prog.acc_kinds.size() >= MAX_ACC) throw_error(QGPU_ERR_UNSUPPORTED, "too many accumulators");
    pass_slot = (int)prog.acc_kinds.size();
    prog.acc_kinds.push_back(ACC_COUNT);
    Insn ld{};
    ld.op = OP_LD;
    ld.ty = DT_BOOL;
    ld.mode = M_LIT;
    ld.imm = 1;
    prog.code.push_back(ld);
    Insn ag{};
    ag.op = OP_AGG;
    ag.ty = DT_BOOL;
    ag.mode = ACC_COUNT;
    ag.arg = pass_slot;
    prog.code.push_back(ag);
  }
  return ap;
}

// ----------------------------------------------------------------------------------------------
// HashAggregate (executor.rs:157-190; grouped path per SURVEY.md 8c Tier 2), optionally fused with
// the Filter directly below it.
// ----------------------------------------------------------------------------------------------
std::shared_ptr<Table> op_aggregate(Ctx* c, const Table& in, const qgpu_expr* pred, const qgpu_expr* groups, int n_group,
                                    const qgpu_agg_expr* aggs, int n_aggr) {
  if (n_aggr <= 0) return nullptr;  // :163-165
  const int64_t n_batches = (int64_t)in.batch_offsets.size() - 1;
  // "agg.keep_empty" (set by the multi-GPU partial-aggregate stage): a rank whose shard yields nothing must still
  // hand a typed, empty (grouped) or all-NULL one-row (scalar) table to the all-gather instead of Ok(vec![])
  const bool keep_empty = c->opt("agg.keep_empty", 0) != 0;
  // a rank of a multi-GPU merge must reach the collective even when its own shard is empty
  const bool must_participate = c->world > 1 && c->opt("agg.merge_ranks", 0) != 0;
  if (n_batches == 0 && !keep_empty && !must_participate) return nullptr;  // no batches -> nothing evaluated -> Ok(vec![])
  if (n_group > MAX_KEYS) throw_error(QGPU_ERR_UNSUPPORTED, "more than %d GROUP BY keys", MAX_KEYS);

  AggPlan ap = plan_aggregate(in, pred, groups, n_group, aggs, n_aggr);
  Program& prog = ap.prog;
  const int n_pred = ap.n_pred;
  const int pass_slot = ap.pass_slot;
  const std::vector<int>& key_cols = ap.key_cols;
  const std::vector<std::string>& key_names = ap.key_names;
  const int n_acc = (int)prog.acc_kinds.size();
  const int R = rows_per_thread(c);

  auto out = new_table(c);
  auto add_agg_outputs = [&](const uint64_t* acc, const uint32_t* slots, int64_t ng) {
    std::vector<AggFinalizeDesc> descs;
    for (int i = 0; i < n_aggr; ++i) {
      const AggOutput& o = prog.agg_out[i];
      Column col = alloc_column(c, o.out_type, ng, true);
      AggFinalizeDesc d;
      d.func = o.func;
      d.in_type = o.in_type;
      d.out_type = o.out_type;
      d.val_slot = o.val_slot < 0 ? 0 : o.val_slot;
      d.cnt_slot = o.cnt_slot;
      d.out_data = col.data;
      d.out_valid = col.valid;
      descs.push_back(d);
      Field f;
      f.name = "col_" + std::to_string(i);  // executor.rs:203-205
      f.dtype = o.out_type;
      f.nullable = true;
      out->fields.push_back(f);
      out->cols.push_back(col);
    }
    launch_agg_finalize(c, acc, slots, n_acc, ng, descs.data(), (int)descs.size());
  };

  ScanParams P;
  DeviceProgram dp;
  // "agg.merge_ranks" (multi-GPU): every rank aggregates its row shard and the raw accumulators are all-gathered over
  // NCCL and merged on the device before finalisation - PartialAggregate(n) -> FinalAggregate of
  // query-distributed/src/planner.rs:200-226 inside one call; every rank returns the same result.
  const bool merge_ranks = c->world > 1 && c->opt("agg.merge_ranks", 0) != 0;

  if (n_group == 0) {
    // ---- scalar aggregates: one fused pass, one host synchronisation ----
    bind_program(c, in, prog, P, dp);
    P.n_pred = n_pred;
    const int64_t max_grid = (int64_t)c->sm_count * 8;
    DevBuf partials(c, (size_t)max_grid * n_acc * 8 + 256);
    DevBuf finals(c, (size_t)(n_acc + 1) * 8 + 256);  // [n_acc] accumulators, then the device error word
    DevBuf ticket(c, 256, true);
    P.partials = partials.as<uint64_t>();
    P.finals = finals.as<uint64_t>();
    P.ticket = ticket.as<uint32_t>();
    launch_scan(c, SK_SCALAR, R, P, "k_scan_filter_agg", &prog);
    QG_CUDA(cudaMemsetAsync(finals.as<uint64_t>() + n_acc, 0, 8, c->stream));
    QG_CUDA(cudaMemcpyAsync(finals.as<uint64_t>() + n_acc, P.err, 4, cudaMemcpyDeviceToDevice, c->stream));
    if (merge_ranks) merge_scalar_across_ranks(c, finals.as<uint64_t>(), n_acc, prog.acc_kinds.data());
    out->n_rows = 1;
    out->batch_offsets = {0, 1};
    add_agg_outputs(finals.as<uint64_t>(), nullptr, 1);  // speculative: discarded if an error / empty input shows up
    uint64_t* h = (uint64_t*)c->scratch(16);
    QG_CUDA(cudaMemcpyAsync(h, finals.as<uint64_t>() + n_acc, 8, cudaMemcpyDeviceToHost, c->stream));
    if (pass_slot >= 0) QG_CUDA(cudaMemcpyAsync(h + 1, finals.as<uint64_t>() + pass_slot, 8, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    raise_device_error((uint32_t)h[0]);
    if (pass_slot >= 0 && h[1] == 0 && !keep_empty) return nullptr;
    return out;
  }

  // ---- GROUP BY ----
  auto empty_groups = [&]() -> std::shared_ptr<Table> {
    if (!keep_empty) return nullptr;
    for (int g = 0; g < n_group; ++g) {
      Field f = in.fields[key_cols[g]];
      f.name = key_names[g];
      f.nullable = true;
      out->fields.push_back(f);
      out->cols.push_back(alloc_column(c, f.dtype, 0, false));
    }
    for (int i = 0; i < n_aggr; ++i) {
      out->fields.push_back(Field{"col_" + std::to_string(i), prog.agg_out[i].out_type, true});
      out->cols.push_back(alloc_column(c, prog.agg_out[i].out_type, 0, false));
    }
    return out;
  };
  bool keys_fixed = true;
  for (int kc : key_cols)
    if (in.cols[kc].dtype == DT_UTF8) keys_fixed = false;
  if (merge_ranks && !keys_fixed) throw_error(QGPU_ERR_UNSUPPORTED, "multi-GPU GROUP BY on Utf8 keys (dictionary-encode them first)");

  GroupTable gt;
  memset(&gt, 0, sizeof gt);
  gt.n_acc = n_acc;
  gt.n_keys = n_group;
  gt.log_regions = GROUP_REGION_BITS;
  DevBuf rep, acc;
  bool table_made = false;
  auto make_table = [&](GroupTable& g, DevBuf& rep_b, DevBuf& acc_b, uint32_t cap) {
    rep_b = DevBuf(c, (size_t)cap * 4 + 256);
    acc_b = DevBuf(c, (size_t)cap * n_acc * 8 + 256);
    g.rep = rep_b.as<uint32_t>();
    g.acc = acc_b.as<uint64_t>();
    g.cap = cap;
    launch_group_init(c, g, prog.acc_kinds.data());
    table_made = true;
  };
  auto pow2_at_least = [](uint64_t x) {
    uint64_t p = 1024;
    while (p < x) p <<= 1;
    return p;
  };
  // err word + overflow flag + group count of one pass, read back together
  auto read_status = [&](const uint32_t* d_err, const uint32_t* d_ovf, uint32_t& err, uint32_t& ovf) {
    uint32_t* h = (uint32_t*)c->scratch(16);
    h[1] = 0;
    QG_CUDA(cudaMemcpyAsync(h, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
    if (d_ovf) QG_CUDA(cudaMemcpyAsync(h + 1, d_ovf, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    err = h[0];
    ovf = h[1];
  };

  uint32_t local_err = 0;
  const Table* src = &in;         // the table the group table's row ids index: `in`, or its region-major copy
  std::shared_ptr<Table> regio;
  DevBuf region_ctl;
  uint32_t region_cap = 0;
  const uint32_t* region_counts = nullptr;
  bool exchanged = false;         // multi-GPU: the rows were exchanged by key, this rank's groups are final
  int64_t exchanged_rows = 0;
  if (in.n_rows > 0 || merge_ranks) {  // (an empty shard of a multi-GPU aggregate still takes part in the collectives)
    for (int g = 0; g < n_group; ++g) gt.keys[g] = colref_of(in.cols[key_cols[g]]);
    bool done = false;
    // tier 1: few groups -> per-CTA directory + per-thread private accumulators in shared memory
    int gp = (int)c->opt("group.smem_groups", 8);
    if (in.n_rows > 0 && keys_fixed && gp > 0 && c->opt("group.force_global", 0) == 0) {
      Program p1 = prog;
      std::vector<int> kslots;
      for (int kc : key_cols) kslots.push_back(p1.slot_for(kc));
      bind_program(c, in, p1, P, dp);
      P.n_pred = n_pred;
      P.n_keys = n_group;
      for (int g = 0; g < n_group; ++g) P.key_slots[g] = kslots[g];
      while (gp > 1) {
        P.gp = gp;
        if (scan_smem_bytes(SK_GROUP_SMEM, R, P) <= 100 * 1024) break;
        gp /= 2;
      }
      P.gp = gp;
      if (scan_smem_bytes(SK_GROUP_SMEM, R, P) <= 200 * 1024) {
        make_table(gt, rep, acc, (uint32_t)pow2_at_least((uint64_t)c->sm_count * 8 * gp * 2));
        P.gt = gt;
        DevBuf overflow(c, 256, true);
        P.overflow = overflow.as<uint32_t>();
        launch_scan(c, SK_GROUP_SMEM, R, P, "k_scan_filter_group_smem", &p1);
        uint32_t err, ovf;
        read_status(P.err, P.overflow, err, ovf);
        if (!ovf && !(err & ERR_CAPACITY)) {
          done = true;
          local_err = err;
        } else if (!merge_ranks) {
          raise_device_error(err);
        }
      }
    }
    // ---- multi-GPU, many groups: exchange ROWS by key instead of all-gathering every rank's partial groups ----
    // (query-distributed/src/planner.rs:200-226 merges partials on one node; with ~10^7 groups per rank that is
    // gigabytes to every rank.)  The radix scatter of the rows (below) writes each row into the exchange arena of
    // the GPU that owns its hash region, over NVLink; every rank then aggregates the rows of its 128 / W regions and
    // returns ITS shard of the groups - the result is partitioned by key (Merge::Concat finishes it).  Whether to
    // take this path is decided from all-gathered facts, so every rank takes the same branch.
    if (merge_ranks && c->world >= 2 && c->world <= PJ_MAX_WORLD && (c->world & (c->world - 1)) == 0 &&
        c->opt("group.exchange", 1) != 0 && n_group == 1) {
      std::vector<int> carry(prog.col_of_slot.begin(), prog.col_of_slot.end());
      carry.push_back(key_cols[0]);
      std::sort(carry.begin(), carry.end());
      carry.erase(std::unique(carry.begin(), carry.end()), carry.end());
      bool elig = in.cols[key_cols[0]].dtype == DT_I64 && in.cols[key_cols[0]].valid == nullptr && (int)carry.size() <= PJ_MAX_COLS;
      for (int tc : carry)
        if (dtype_width(in.cols[tc].dtype) == 0 || in.cols[tc].valid != nullptr) elig = false;
      const std::vector<int64_t> all = allgather_i64(c, {(in.n_rows > 0 && !done) ? 1 : 0, elig ? 1 : 0, in.n_rows});
      bool any_need = false, all_elig = true;
      int64_t total_rows = 0;
      for (int r = 0; r < c->world; ++r) {
        any_need |= all[(size_t)r * 3] != 0;
        all_elig &= all[(size_t)r * 3 + 1] != 0;
        total_rows += all[(size_t)r * 3 + 2];
      }
      if (any_need && all_elig && total_rows >= c->opt("group.partition_min_rows", 1 << 22)) {
        const int W = c->world;
        constexpr uint32_t parts = 1u << GROUP_REGION_BITS;
        const uint32_t F = parts / (uint32_t)W;
        uint64_t cap64 = (uint64_t)total_rows / parts;
        cap64 += cap64 / 8 + 1024;
        cap64 = (cap64 + 1023) / 1024 * 1024;
        if (parts * cap64 < 0xfff00000ull) {
          const uint32_t cap = (uint32_t)cap64;
          const size_t ctl_bytes = ((size_t)PJ_MAX_PARTS * PJ_CTR_STRIDE + 64) * 4;
          size_t off = (ctl_bytes + 255) / 256 * 256;
          std::vector<size_t> offs(carry.size());
          for (size_t i = 0; i < carry.size(); ++i) {
            offs[i] = off;
            off += pad_bytes((size_t)F * cap * dtype_width(in.cols[carry[i]].dtype));
          }
          arena_ensure(c, off);
          char* A0 = (char*)c->arena;
          uint32_t* counts = (uint32_t*)A0;
          uint32_t* flags = counts + (size_t)PJ_MAX_PARTS * PJ_CTR_STRIDE;
          QG_CUDA(cudaMemsetAsync(A0, 0, ctl_bytes, c->stream));
          stream_barrier(c);  // every control block is zero before anybody reserves rows in it
          PJCols pc;
          memset(&pc, 0, sizeof pc);
          PJPeers pe;
          memset(&pe, 0, sizeof pe);
          pe.world = W;
          pe.regions_per_rank = F;
          for (size_t i = 0; i < carry.size(); ++i) {
            if (carry[i] == key_cols[0]) pc.key_col = pc.n;
            pc.c[pc.n++] = PJCol{in.cols[carry[i]].data, nullptr, dtype_width(in.cols[carry[i]].dtype), 0};
          }
          for (int r = 0; r < W; ++r) {
            pe.cursor[r] = (uint32_t*)c->arena_peer[r];
            for (size_t i = 0; i < carry.size(); ++i) pe.out[r][i] = (char*)c->arena_peer[r] + offs[i];
          }
          PJMap map;
          memset(&map, 0, sizeof map);
          map.dense = 2;
          map.parts = parts;
          map.log_parts = GROUP_REGION_BITS;
          if (in.n_rows > 0)
            launch_pj_scatter(c, (const int64_t*)in.cols[key_cols[0]].data, in.n_rows, map, cap, nullptr, flags, pc, &pe,
                              "k_pj_scatter_groups_nvlink");
          stream_barrier(c);  // every row has landed
          const std::vector<int64_t> fl = allgather_i64(c, {(int64_t)read_u32(c, flags)});
          bool overflow = false;
          for (int r = 0; r < W; ++r) overflow |= fl[(size_t)r] != 0;
          if (!overflow) {
            // this rank's regions as a region-major table view over its arena
            regio = new_table(c);
            regio->fields = in.fields;
            regio->n_rows = (int64_t)F * cap;
            regio->batch_offsets = {0, regio->n_rows};
            regio->cols.resize(in.cols.size());
            for (size_t i = 0; i < carry.size(); ++i) {
              Column v;
              v.dtype = in.cols[carry[i]].dtype;
              v.n = regio->n_rows;
              v.data = A0 + offs[i];
              v.owned = false;
              regio->cols[(size_t)carry[i]] = v;
            }
            src = regio.get();
            region_cap = cap;
            region_counts = counts;
            gt.region_base = (uint32_t)c->rank * F;
            gt.log_regions = 0;
            while ((1u << gt.log_regions) < F) ++gt.log_regions;
            exchanged = true;
            done = false;  // (a rank whose shared-memory pass had sufficed redoes its share too)
            exchanged_rows = 0;
            for (int r = 0; r < W; ++r) exchanged_rows += all[(size_t)r * 3 + 2];
          }
        }
      }
    }
    // tier 2: open-addressing table in HBM.  With many groups the table is far larger than the L2 and every row
    // would pay random HBM accesses (slot, representative key, one atomic per aggregate).  So the referenced columns
    // are first radix-scattered on the hash bits that pick the row's SUB-TABLE (pjoin_kernels.cu k_pj_scatter,
    // scan_params.cuh group_home_slot): the scan then walks the regions one after the other and each sub-table
    // (1/128 of the table) is L2-resident while its rows stream through.
    if (exchanged || (in.n_rows > 0 && !done)) {
      const int64_t rows_here = exchanged ? exchanged_rows / c->world + 1 : in.n_rows;  // (exchanged: the expected share)
      if (!exchanged && n_group == 1 && in.cols[key_cols[0]].dtype == DT_I64 && in.cols[key_cols[0]].valid == nullptr &&
          in.n_rows >= c->opt("group.partition_min_rows", 1 << 22) && c->opt("group.partition", 1) != 0 &&
          (int)prog.col_of_slot.size() < PJ_MAX_COLS) {
        // columns to carry: what the program reads, and the key (the table compares keys at representative rows)
        std::vector<int> carry(prog.col_of_slot.begin(), prog.col_of_slot.end());
        carry.push_back(key_cols[0]);
        bool ok = true;
        for (int tc : carry)
          if (dtype_width(in.cols[tc].dtype) == 0 || in.cols[tc].valid != nullptr) ok = false;
        constexpr uint32_t parts = 1u << GROUP_REGION_BITS;
        uint64_t cap64 = (uint64_t)in.n_rows / parts;
        cap64 += cap64 / 8 + 1024;
        cap64 = (cap64 + 1023) / 1024 * 1024;  // a multiple of the scan tile: a tile never straddles two regions
        if (ok && parts * cap64 < 0xfff00000ull) {
          const uint32_t cap = (uint32_t)cap64;
          region_ctl = DevBuf(c, ((size_t)PJ_MAX_PARTS * PJ_CTR_STRIDE + 64) * 4, true);
          uint32_t* counts = region_ctl.as<uint32_t>();
          uint32_t* flags = counts + (size_t)PJ_MAX_PARTS * PJ_CTR_STRIDE;
          regio = new_table(c);
          regio->fields = in.fields;
          regio->n_rows = (int64_t)parts * cap;
          regio->batch_offsets = {0, regio->n_rows};
          regio->cols.resize(in.cols.size());
          PJCols pc;
          memset(&pc, 0, sizeof pc);
          for (int tc : carry) {
            if (regio->cols[(size_t)tc].data) continue;
            Column d = alloc_column(c, in.cols[tc].dtype, regio->n_rows, false);
            regio->cols[(size_t)tc] = d;
            if (tc == key_cols[0]) pc.key_col = pc.n;
            pc.c[pc.n++] = PJCol{in.cols[tc].data, d.data, dtype_width(d.dtype), 0};
          }
          {
            PJMap map;
            memset(&map, 0, sizeof map);
            map.dense = 2;
            map.parts = parts;
            map.log_parts = GROUP_REGION_BITS;
            launch_pj_scatter(c, (const int64_t*)in.cols[key_cols[0]].data, in.n_rows, map, cap, counts, flags, pc, nullptr,
                              "k_pj_scatter_groups");
            if (read_u32(c, flags) != 0) {
              regio.reset();  // a region overflowed (one key holds a large share of the rows): plain pass instead
            } else {
              src = regio.get();
              region_cap = cap;
              region_counts = counts;
            }
          }
        }
      }
      bind_program(c, *src, prog, P, dp);
      for (int g = 0; g < n_group; ++g) gt.keys[g] = colref_of(src->cols[key_cols[g]]);
      if (region_cap) {
        P.region_cap = region_cap;
        P.region_count = region_counts;
        P.region_count_stride = PJ_CTR_STRIDE;
      }
      P.n_pred = n_pred;
      // first table: n / 4 slots (rounded up to a power of two) on large inputs (a failed pass costs more than the init of
      // a generous table), "group.capacity" slots on small ones; 4x larger on every retry
      const uint64_t first_cap = std::max<uint64_t>((uint64_t)c->opt("group.capacity", 1 << 22), (uint64_t)rows_here / 4);
      uint64_t cap = pow2_at_least(std::min<uint64_t>((uint64_t)rows_here * 2, first_cap));
      if (c->opt("group.capacity", 1 << 22) < (1 << 22)) cap = pow2_at_least((uint64_t)c->opt("group.capacity", 1 << 22));  // (tests)
      while (true) {
        make_table(gt, rep, acc, (uint32_t)cap);
        P.gt = gt;
        QG_CUDA(cudaMemsetAsync(P.err, 0, 4, c->stream));
        launch_scan(c, SK_GROUP_GLOBAL, R, P, "k_scan_filter_group_hbm", &prog);
        uint32_t err, ovf;
        read_status(P.err, nullptr, err, ovf);
        if (!(err & ERR_CAPACITY)) {
          local_err = err;
          break;
        }
        if (cap >= (1ull << 31)) throw_error(QGPU_ERR_UNSUPPORTED, "GROUP BY table would exceed 2^31 slots");
        cap <<= 2;  // retry with a 4x larger table
        if (cap > (1ull << 31)) cap = 1ull << 31;
      }
    }
  }
  if (!merge_ranks) raise_device_error(local_err);

  // collect the groups of a table: representative rows + slots
  auto collect = [&](const GroupTable& g, DevBuf& out_rep, DevBuf& out_slot) -> int64_t {
    out_rep = DevBuf(c, (size_t)g.cap * 4 + 256);
    out_slot = DevBuf(c, (size_t)g.cap * 4 + 256);
    DevBuf d_count(c, 256, true);
    launch_group_collect(c, g, out_rep.as<uint32_t>(), out_slot.as<uint32_t>(), d_count.as<uint32_t>());
    return (int64_t)read_u32(c, d_count.as<uint32_t>());
  };
  // keys (gathered from `src` at the representative rows) + finalised aggregates -> out
  auto emit = [&](const Table& src, const std::vector<int>& kcols, const GroupTable& g, DevBuf& out_rep, DevBuf& out_slot, int64_t ng) {
    auto keys_t = gather_rows(c, src, kcols, out_rep.as<uint32_t>(), ng, false);
    out->n_rows = ng;
    out->batch_offsets = {0, ng};
    for (int k = 0; k < n_group; ++k) {
      Field f = keys_t->fields[k];
      f.name = key_names[k];
      f.nullable = true;
      out->fields.push_back(f);
      out->cols.push_back(keys_t->cols[k]);
      keys_t->cols[k].owned = false;
    }
    add_agg_outputs(g.acc, out_slot.as<uint32_t>(), ng);
  };

  DevBuf out_rep, out_slot;
  int64_t ng = 0;
  if (exchanged) {  // an error anywhere is everybody's error (the ranks must leave the collective call alike)
    uint32_t any_err = 0;
    for (int64_t e : allgather_i64(c, {(int64_t)local_err})) any_err |= (uint32_t)e;
    raise_device_error(any_err);
  }
  if (table_made && gt.rep != nullptr) ng = collect(gt, out_rep, out_slot);
  if (!merge_ranks || exchanged) {
    if (ng == 0) return empty_groups();
    emit(*src, key_cols, gt, out_rep, out_slot, ng);
    if (regio) out->keepalive.push_back(regio);
    return out;
  }

  // ---- multi-GPU final merge: all-gather (group keys, raw accumulator rows) and re-group on the device ----
  std::vector<int> all_keys(n_group);
  for (int k = 0; k < n_group; ++k) all_keys[k] = k;
  std::shared_ptr<Table> local_keys;
  if (ng > 0) {
    local_keys = gather_rows(c, *src, key_cols, out_rep.as<uint32_t>(), ng, false);
  } else {  // typed, empty key table so the collective has a schema
    local_keys = new_table(c);
    for (int k = 0; k < n_group; ++k) {
      local_keys->fields.push_back(in.fields[key_cols[k]]);
      local_keys->cols.push_back(alloc_column(c, in.cols[key_cols[k]].dtype, 0, false));
    }
  }
  DevBuf local_acc(c, (size_t)std::max<int64_t>(ng, 1) * n_acc * 8 + 256);
  if (ng > 0) launch_gather_acc_rows(c, gt.acc, out_slot.as<uint32_t>(), n_acc, ng, local_acc.as<uint64_t>());
  DevBuf all_acc;
  uint32_t any_err = 0;
  std::shared_ptr<Table> gkeys = allgather_groups(c, *local_keys, local_acc.as<uint64_t>(), n_acc, local_err, all_acc, any_err);
  raise_device_error(any_err);
  const int64_t NG = gkeys ? gkeys->n_rows : 0;
  if (NG == 0) return empty_groups();
  GroupTable g2;
  memset(&g2, 0, sizeof g2);
  g2.n_acc = n_acc;
  g2.n_keys = n_group;
  for (int k = 0; k < n_group; ++k) g2.keys[k] = colref_of(gkeys->cols[k]);
  DevBuf rep2, acc2, d_err2(c, 256, true);
  make_table(g2, rep2, acc2, (uint32_t)pow2_at_least((uint64_t)NG * 2));
  launch_merge_groups(c, g2, all_acc.as<uint64_t>(), NG, prog.acc_kinds.data(), d_err2.as<uint32_t>());
  DevBuf rep_o, slot_o;
  const int64_t ng2 = collect(g2, rep_o, slot_o);
  emit(*gkeys, all_keys, g2, rep_o, slot_o, ng2);
  out->keepalive.push_back(gkeys);  // (key columns were gathered into `out`; keep the source until the stream drains)
  return out;
}

// ----------------------------------------------------------------------------------------------
// Projection (executor.rs:93-129)
// ----------------------------------------------------------------------------------------------
std::shared_ptr<Table> op_projection(Ctx* c, const std::shared_ptr<Table>& in_ptr, const qgpu_expr* exprs,
                                     const char* const* names, int n) {
  const Table& in = *in_ptr;
  auto out = new_table(c);
  out->batch_offsets = in.batch_offsets;
  out->n_rows = in.n_rows;
  if (n > MAX_OUT) throw_error(QGPU_ERR_UNSUPPORTED, "more than %d projected expressions", MAX_OUT);
  const int64_t n_batches = (int64_t)in.batch_offsets.size() - 1;
  const SchemaView sv = schema_view(in);
  Program prog;
  ScanParams P;
  memset(&P, 0, sizeof P);
  int n_store = 0;
  std::vector<void*> out_data;
  std::vector<uint32_t*> out_valid;
  std::vector<uint8_t> out_types;
  for (int i = 0; i < n; ++i) {
    ExprTree t = parse_postfix(&exprs[i]);
    if (n_batches == 0) {
      Field f{names && names[i] ? names[i] : "col_" + std::to_string(i), DT_NULL, true};
      out->fields.push_back(f);
      out->cols.push_back(Column{});
      continue;
    }
    const ExprNode& root = t.nodes[t.root];
    uint8_t ty = infer_types(t, sv);
    Field f;
    f.name = names && names[i] ? names[i] : "col_" + std::to_string(i);
    f.dtype = ty;
    f.nullable = true;
    if (root.kind == QGPU_EXPR_COLUMN) {
      out->fields.push_back(f);
      out->cols.push_back(borrow(in.cols[root.op]));  // evaluate_expr on Column is an Arc clone (operators.rs:22)
      if (out->keepalive.empty()) out->keepalive.push_back(in_ptr);
      continue;
    }
    if (ty == DT_UTF8) throw_error(QGPU_ERR_UNSUPPORTED, "Utf8 expressions are not on the accelerated path");
    Column col = ty == DT_NULL ? Column{} : alloc_column(c, ty, in.n_rows, t.nodes[t.root].maybe_null);
    col.dtype = ty;
    col.n = in.n_rows;
    if (ty != DT_NULL) {
      compile_value(t, sv, prog);
      Insn st{};
      st.op = OP_STORE;
      st.ty = ty;
      st.arg = n_store;
      prog.code.push_back(st);
      out_data.push_back(col.data);
      out_valid.push_back(col.valid);
      out_types.push_back(ty);
      ++n_store;
    }
    out->fields.push_back(f);
    out->cols.push_back(col);
  }
  if (n_store > 0 && in.n_rows > 0) {
    DeviceProgram dp;
    bind_program(c, in, prog, P, dp);
    P.n_out = n_store;
    for (int i = 0; i < n_store; ++i) {
      P.out_data[i] = out_data[i];
      P.out_valid[i] = out_valid[i];
      P.out_types[i] = out_types[i];
    }
    launch_scan(c, SK_STORE, rows_per_thread(c), P, "k_scan_project", &prog);
    raise_device_error(read_u32(c, P.err));
  }
  return out;
}

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
}  // namespace

std::shared_ptr<Table> op_join(Ctx* c, const Table& left, const Table& right, int join_type, const qgpu_expr* on) {
  const int64_t lb = (int64_t)left.batch_offsets.size() - 1, rb = (int64_t)right.batch_offsets.size() - 1;
  if (lb == 0 || rb == 0) return nullptr;  // executor.rs:350-352, even for LEFT
  if (join_type != QGPU_INNER && join_type != QGPU_LEFT)
    throw_error(QGPU_ERR_UNSUPPORTED, "only INNER and LEFT joins are on the accelerated path");
  if (!on) throw_error(QGPU_ERR_UNSUPPORTED, "join without ON keeps the reference's Cartesian CPU path");
  ExprTree t = parse_postfix(on);
  std::vector<std::pair<int, int>> keys;
  const int n_left = (int)left.cols.size();
  if (!equi_keys(t, t.root, n_left, keys))
    throw_error(QGPU_ERR_UNSUPPORTED, "ON must be Column = Column (AND ...) across the two sides to be accelerated");
  for (auto& k : keys)
    if (k.second >= (int)right.cols.size()) throw_error(QGPU_ERR_EXECUTION, "Column index %d out of bounds", k.second + n_left);
  if (keys.size() != 1) throw_error(QGPU_ERR_UNSUPPORTED, "multi-column join keys are not on the accelerated path yet");
  const Column& lk = left.cols[keys[0].first];
  const Column& rk = right.cols[keys[0].second];
  auto int_like = [](uint8_t d) { return d == DT_I64 || d == DT_I32; };
  const bool ok = (int_like(lk.dtype) && int_like(rk.dtype)) ||
                  (lk.dtype == rk.dtype && (lk.dtype == DT_F64 || lk.dtype == DT_F32 || lk.dtype == DT_BOOL));
  if (!ok) {
    if (lk.dtype == rk.dtype || (lk.dtype == DT_F64 && int_like(rk.dtype)) || (rk.dtype == DT_F64 && int_like(lk.dtype)))
      throw_error(QGPU_ERR_UNSUPPORTED, "join key type is not on the accelerated path");
    throw_error(QGPU_ERR_ARROW, "Invalid argument error: Invalid comparison operation: join keys of different types");
  }
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
    if (pjoin(c, left, right, lki, rki, left_join, c->opt("join.dense", 1) != 0, hm, out, empty, was_dense)) return empty ? nullptr : out;
    if (was_dense && pjoin(c, left, right, lki, rki, left_join, false, hm, out, empty, was_dense)) return empty ? nullptr : out;
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
  }
  for (size_t i = 0; i < rt->cols.size(); ++i) {
    out->fields.push_back(rt->fields[i]);
    out->cols.push_back(rt->cols[i]);
    rt->cols[i].owned = false;
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

// ----------------------------------------------------------------------------------------------
// Exchange: hash partitioning (partition.rs:151-212)
// ----------------------------------------------------------------------------------------------
std::vector<std::shared_ptr<Table>> op_exchange_hash(Ctx* c, const Table& in, const std::vector<int>& key_cols, int num_partitions,
                                                     bool siphash) {
  if (num_partitions < 1 || num_partitions > MAX_PARTS)
    throw_error(QGPU_ERR_UNSUPPORTED, "num_partitions must be in 1..%d", MAX_PARTS);
  if (key_cols.empty()) throw_error(QGPU_ERR_PARTITION, "No key columns found in batch");  // partition.rs:173-177
  if ((int)key_cols.size() > MAX_KEYS) throw_error(QGPU_ERR_UNSUPPORTED, "more than %d partition keys", MAX_KEYS);
  std::vector<std::shared_ptr<Table>> parts((size_t)num_partitions);
  const int64_t n = in.n_rows;
  if (n == 0) return parts;
  PartKeys K;
  memset(&K, 0, sizeof K);
  K.n_keys = (int)key_cols.size();
  for (int i = 0; i < K.n_keys; ++i) K.keys[i] = colref_of(in.cols[key_cols[i]]);
  const int64_t n_tiles = part_tiles(n);
  DevBuf part_id(c, pad_bytes((size_t)n));
  DevBuf hist(c, (size_t)num_partitions * n_tiles * 4 + 256);
  DevBuf hist_off(c, ((size_t)num_partitions * n_tiles + 1) * 8 + 256);
  DevBuf perm(c, pad_bytes((size_t)n * 4));
  launch_part_hist(c, K, n, num_partitions, siphash, part_id.as<uint8_t>(), hist.as<uint32_t>());
  scan_u32_to_i64(c, hist.as<uint32_t>(), (int64_t)num_partitions * n_tiles, hist_off.as<int64_t>());
  launch_part_perm(c, n, num_partitions, part_id.as<uint8_t>(), hist_off.as<int64_t>(), perm.as<uint32_t>());
  // partition starts
  std::vector<int64_t> starts((size_t)num_partitions + 1);
  {
    int64_t* h = (int64_t*)c->scratch((size_t)(num_partitions + 1) * 8);
    DevBuf d_starts(c, (size_t)(num_partitions + 1) * 8 + 256);
    launch_gather_strided_i64(c, hist_off.as<int64_t>(), n_tiles, num_partitions + 1, d_starts.as<int64_t>());
    QG_CUDA(cudaMemcpyAsync(h, d_starts.p, (size_t)(num_partitions + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(starts.data(), h, (size_t)(num_partitions + 1) * 8);
  }
  std::vector<int> all(in.cols.size());
  for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
  for (int p = 0; p < num_partitions; ++p) {
    const int64_t cnt = starts[p + 1] - starts[p];
    if (cnt == 0) continue;
    parts[p] = gather_rows(c, in, all, perm.as<uint32_t>() + starts[p], cnt, false);
  }
  c->sync();
  return parts;
}

}  // namespace qgpu
