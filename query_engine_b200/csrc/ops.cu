// ops.cu — operator orchestration: the bodies of QueryExecutor::execute_filter / execute_aggregate /
// execute_projection (crates/query-executor/src/executor.rs:93-190) and of Partitioner::partition_by_hash
// (crates/query-distributed/src/partition.rs:151-212) on device tables.  The join operators are in ops_join.cu.
#include <algorithm>
#include <cstring>

#include "exchange_kernels.cuh"
#include "interp.cuh"
#include "pjoin_kernels.cuh"
#include "qgpu_internal.h"
#include "scan_kernels.cuh"
#include "sort_kernels.cuh"
#include "util_kernels.cuh"

namespace qgpu {

int rows_per_thread(Ctx* c) {
  const int64_t r = c->opt("scan.rows_per_thread", 4);
  return r == 8 ? 8 : 4;
}

// upload a program and fill the column table of the scan parameters
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

// The same, with the reference's full message: arrow-rs names the operands of the failing operation ("Overflow happened
// on: 9223372036854775806 * 2"), and which operation fails FIRST is defined by its evaluation order - per expression,
// batch after batch, node after node, row after row.  So after an error the program is re-run once by the interpreter
// with the diagnostic sink (scan_body.cuh SinkDiag): pass 0 finds the minimal (section, batch, instruction, row) key
// among the failing operations, pass 1 records that operation's operands.  Error path only: cost is irrelevant.
void raise_device_error(Ctx* c, uint32_t err, const Table& in, const Program& prog, int n_pred) {
  if (!(err & (ERR_DIVZERO | ERR_OVERFLOW)) || in.n_rows <= 0 || prog.code.empty() || prog.code.size() > 255) {
    raise_device_error(err);
    return;
  }
  std::vector<uint8_t> section(prog.code.size());
  {
    int sec = 0;
    for (size_t pc = 0; pc < prog.code.size(); ++pc) {
      if ((int)pc == n_pred && n_pred > 0) sec = 1;
      section[pc] = (uint8_t)std::min(sec, 63);
      const bool ends = prog.code[pc].op == OP_AGG || prog.code[pc].op == OP_STORE;
      const bool next_ends = pc + 1 < prog.code.size() && (prog.code[pc + 1].op == OP_AGG || prog.code[pc + 1].op == OP_STORE);
      if (ends && !next_ends) ++sec;
    }
  }
  const int nb = (int)in.batch_offsets.size() - 1;
  DevBuf d_off(c, (size_t)(nb + 1) * 8 + 16), d_sec(c, section.size() + 16), d_key(c, 64), d_rec(c, sizeof(DiagRecord) + 16, true);
  QG_CUDA(cudaMemcpyAsync(d_off.p, in.batch_offsets.data(), (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  QG_CUDA(cudaMemcpyAsync(d_sec.p, section.data(), section.size(), cudaMemcpyHostToDevice, c->stream));
  QG_CUDA(cudaMemsetAsync(d_key.p, 0xff, 8, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));  // (the sources above are pageable host vectors)
  ScanParams P;
  DeviceProgram dp;
  bind_program(c, in, prog, P, dp);
  P.n_pred = n_pred;
  P.diag_n_batches = nb;
  P.diag_batch_offsets = d_off.as<int64_t>();
  P.diag_section_of_pc = d_sec.as<uint8_t>();
  P.diag_key = d_key.as<unsigned long long>();
  P.diag_rec = d_rec.as<DiagRecord>();
  for (int mode = 0; mode < 2; ++mode) {
    P.diag_mode = mode;
    launch_scan(c, SK_DIAG, 4, P, "k_scan_diag", nullptr);
  }
  DiagRecord rec;
  QG_CUDA(cudaMemcpyAsync(&rec, d_rec.p, sizeof rec, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));
  if (rec.err & ERR_DIVZERO) throw_error(QGPU_ERR_ARROW, "Divide by zero error");
  if (rec.err & ERR_OVERFLOW) {
    if (rec.op == OP_MOD) throw_error(QGPU_ERR_ARROW, "Compute error: Overflow happened on: remainder MIN %% -1");
    const char* sym = rec.op == OP_ADD ? "+" : rec.op == OP_SUB ? "-" : rec.op == OP_MUL ? "*" : "/";
    throw_error(QGPU_ERR_ARROW, "Compute error: Overflow happened on: %lld %s %lld", (long long)rec.lhs, sym, (long long)rec.rhs);
  }
  raise_device_error(err);  // (nothing found on this table: an error raised elsewhere, e.g. on another rank)
}

// output columns col_0.. of the aggregates (executor.rs:192-212) from raw accumulator rows: acc[slot(i)][*], slot(i) =
// slots ? slots[i] : i
void add_agg_outputs(Ctx* c, const Program& prog, int n_aggr, const uint64_t* acc, const uint32_t* slots, int64_t ng, Table& out,
                     const uint32_t* d_live) {
  const int n_acc = (int)prog.acc_kinds.size();
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
    out.fields.push_back(f);
    out.cols.push_back(col);
  }
  launch_agg_finalize(c, acc, slots, n_acc, ng, descs.data(), (int)descs.size(), d_live);
}

namespace {

Column borrow(const Column& col) {
  Column b = col;
  b.owned = false;
  b.owned_valid = false;
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
  std::vector<int64_t> pre(nbnd);
  {
    // one host round trip: the boundaries go up from one half of the pinned scratch, the survivor prefixes and the error
    // word come back into the other half
    int64_t* h = (int64_t*)c->scratch((size_t)(2 * nbnd + 1) * 8);
    int64_t* h_pre = h + nbnd;
    memcpy(h, in.batch_offsets.data(), (size_t)nbnd * 8);
    QG_CUDA(cudaMemcpyAsync(d_bnd.p, h, (size_t)nbnd * 8, cudaMemcpyHostToDevice, c->stream));
    launch_prefix_at(c, d_bnd.as<int64_t>(), nbnd, tile_rows, mask.as<uint32_t>(), tile_off.as<int64_t>(), d_pre.as<int64_t>());
    QG_CUDA(cudaMemcpyAsync(h_pre, d_pre.p, (size_t)nbnd * 8, cudaMemcpyDeviceToHost, c->stream));
    h_pre[nbnd] = 0;
    QG_CUDA(cudaMemcpyAsync(h_pre + nbnd, P.err, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(pre.data(), h_pre, (size_t)nbnd * 8);
    raise_device_error(c, (uint32_t)h_pre[nbnd], in, prog, 0);
  }
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
  if (k > 0 && !cc.empty()) launch_compact(c, n, tile_rows, mask.as<uint32_t>(), tile_off.as<int64_t>(), cc.data(), (int)cc.size(), k);
  if (any_utf8) {
    std::vector<int> ucols;
    for (size_t ci = 0; ci < in.cols.size(); ++ci)
      if (in.cols[ci].dtype == DT_UTF8) ucols.push_back((int)ci);
    auto g = gather_rows(c, in, ucols, rowidx.as<uint32_t>(), k, false);
    for (size_t j = 0; j < ucols.size(); ++j) {
      out->cols[ucols[j]] = g->cols[j];
      g->cols[j].owned = false;  // ownership moves to `out`
      g->cols[j].owned_valid = false;
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

  // GROUP BY on computed expressions (physical_plan.rs:35-39 allows any PhysicalExpr): the expressions are materialised
  // as extra columns (the projection scan), then the aggregation groups by those columns.  The reference evaluates group
  // expressions on the FILTERED batches, so a fused predicate is applied first (an error on a row the filter drops must
  // not surface).
  {
    std::vector<int> computed;
    for (int g = 0; g < n_group; ++g) {
      ExprTree gt = parse_postfix(&groups[g]);
      if (gt.nodes[gt.root].kind != QGPU_EXPR_COLUMN) computed.push_back(g);
    }
    if (!computed.empty()) {
      if (must_participate) throw_error(QGPU_ERR_UNSUPPORTED, "multi-GPU GROUP BY on a computed expression");
      std::shared_ptr<Table> filtered;
      const Table* src_t = &in;
      if (pred) {
        filtered = op_filter(c, in, pred);
        src_t = filtered.get();
        if ((int64_t)src_t->batch_offsets.size() - 1 == 0) return nullptr;  // no surviving batch -> Ok(vec![])
      }
      const Table& base = *src_t;
      std::vector<qgpu_expr> exprs;
      std::vector<std::string> names;
      for (int g : computed) {
        exprs.push_back(groups[g]);
        names.push_back("group_" + std::to_string(g));  // (the oracle's name of a computed key)
      }
      std::vector<const char*> name_ptrs;
      for (auto& nm : names) name_ptrs.push_back(nm.c_str());
      std::shared_ptr<Table> alias(std::shared_ptr<Table>(), const_cast<Table*>(&base));  // non-owning
      auto proj = op_projection(c, alias, exprs.data(), name_ptrs.data(), (int)exprs.size());
      Table ext;  // (no context: a view that owns nothing)
      ext.fields = base.fields;
      ext.n_rows = base.n_rows;
      ext.batch_offsets = base.batch_offsets;
      for (const Column& col : base.cols) {
        Column b = col;
        b.owned = false;
        b.owned_valid = false;
        ext.cols.push_back(b);
      }
      std::vector<qgpu_expr_node> nodes((size_t)n_group);
      std::vector<qgpu_expr> new_groups((size_t)n_group);
      std::vector<std::string> node_names((size_t)n_group);
      for (int g = 0; g < n_group; ++g) new_groups[(size_t)g] = groups[g];
      for (size_t j = 0; j < computed.size(); ++j) {
        const int g = computed[j];
        Column b = proj->cols[j];
        b.owned = false;
        b.owned_valid = false;
        ext.fields.push_back(proj->fields[j]);
        ext.cols.push_back(b);
        memset(&nodes[(size_t)g], 0, sizeof(qgpu_expr_node));
        nodes[(size_t)g].kind = QGPU_EXPR_COLUMN;
        nodes[(size_t)g].op = (int32_t)ext.cols.size() - 1;
        node_names[(size_t)g] = names[j];
        nodes[(size_t)g].str = node_names[(size_t)g].c_str();
        new_groups[(size_t)g].nodes = &nodes[(size_t)g];
        new_groups[(size_t)g].n_nodes = 1;
      }
      auto res = op_aggregate(c, ext, nullptr, new_groups.data(), n_group, aggs, n_aggr);
      if (res) {  // the result may borrow nothing from `ext` beyond this call: key columns are gathered copies
        res->keepalive.push_back(proj);
        if (filtered) res->keepalive.push_back(filtered);
      }
      return res;
    }
  }

  HostMarks hm(c);  // "debug.host_timing": where the host spends a step
  AggPlan ap = plan_aggregate(in, pred, groups, n_group, aggs, n_aggr);
  hm.mark("aggregate: plan + compile");
  Program& prog = ap.prog;
  const int n_pred = ap.n_pred;
  const int pass_slot = ap.pass_slot;
  const std::vector<int>& key_cols = ap.key_cols;
  const std::vector<std::string>& key_names = ap.key_names;
  const int n_acc = (int)prog.acc_kinds.size();
  const int R = rows_per_thread(c);

  auto out = new_table(c);
  auto add_agg_outputs = [&](const uint64_t* acc, const uint32_t* slots, int64_t ng, const uint32_t* d_live = nullptr) {
    qgpu::add_agg_outputs(c, prog, n_aggr, acc, slots, ng, *out, d_live);
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
    hm.mark("aggregate: bind + buffers");
    launch_scan(c, SK_SCALAR, R, P, "k_scan_filter_agg", &prog);
    hm.mark("aggregate: scan launch");
    QG_CUDA(cudaMemsetAsync(finals.as<uint64_t>() + n_acc, 0, 8, c->stream));
    QG_CUDA(cudaMemcpyAsync(finals.as<uint64_t>() + n_acc, P.err, 4, cudaMemcpyDeviceToDevice, c->stream));
    if (merge_ranks) merge_scalar_across_ranks(c, finals.as<uint64_t>(), n_acc, prog.acc_kinds.data());
    out->n_rows = 1;
    out->batch_offsets = {0, 1};
    add_agg_outputs(finals.as<uint64_t>(), nullptr, 1);  // speculative: discarded if an error / empty input shows up
    uint64_t* h = (uint64_t*)c->scratch(32);
    h[2] = 0;
    QG_CUDA(cudaMemcpyAsync(h, finals.as<uint64_t>() + n_acc, 8, cudaMemcpyDeviceToHost, c->stream));
    if (pass_slot >= 0) QG_CUDA(cudaMemcpyAsync(h + 1, finals.as<uint64_t>() + pass_slot, 8, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaMemcpyAsync(h + 2, P.err + 2, 4, cudaMemcpyDeviceToHost, c->stream));
    hm.mark("aggregate: outputs + copies queued");
    QG_CUDA(cudaStreamSynchronize(c->stream));
    hm.mark("aggregate: sync");
    if (pred) note_scan_feedback(c, h[2], (uint64_t)in.n_rows);
    if (merge_ranks) raise_device_error((uint32_t)h[0]);
    raise_device_error(c, (uint32_t)h[0], in, prog, pred ? n_pred : 0);
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
  // (d_count: an optional third word read in the same round trip - the group count of a speculative collect)
  auto read_status = [&](const uint32_t* d_err, const uint32_t* d_ovf, uint32_t& err, uint32_t& ovf, const uint32_t* d_count = nullptr,
                         uint32_t* count = nullptr) {
    uint32_t* h = (uint32_t*)c->scratch(32);
    h[1] = 0;
    h[2] = 0;
    h[3] = 0;
    QG_CUDA(cudaMemcpyAsync(h, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
    if (d_ovf) QG_CUDA(cudaMemcpyAsync(h + 1, d_ovf, 4, cudaMemcpyDeviceToHost, c->stream));
    if (d_ovf) QG_CUDA(cudaMemcpyAsync(h + 2, d_err + 2, 4, cudaMemcpyDeviceToHost, c->stream));  // (smem tier: predicate survivors)
    if (d_count) QG_CUDA(cudaMemcpyAsync(h + 3, d_count, 4, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    err = h[0];
    ovf = h[1];
    if (d_ovf && pred) note_scan_feedback(c, h[2], (uint64_t)in.n_rows);
    if (count) *count = h[3];
  };

  uint32_t local_err = 0;
  const Table* src = &in;         // the table the group table's row ids index: `in`, or its region-major copy
  std::shared_ptr<Table> regio;
  DevBuf region_ctl;
  uint32_t region_cap = 0;
  const uint32_t* region_counts = nullptr;
  bool exchanged = false;         // multi-GPU: the rows were exchanged by key, this rank's groups are final
  DevBuf pre_rep, pre_slot;       // smem tier: groups collected right behind the scan
  bool pre_collected = false;
  int64_t pre_ng = 0;
  int64_t exchanged_rows = 0;
  if (in.n_rows > 0 || merge_ranks) {  // (an empty shard of a multi-GPU aggregate still takes part in the collectives)
    for (int g = 0; g < n_group; ++g) gt.keys[g] = colref_of(in.cols[key_cols[g]]);
    bool done = false;
    // tier 1: few groups -> per-CTA directory + per-thread private accumulators in shared memory.  Whether the tier is
    // tried depends only on facts every rank of a multi-GPU aggregate shares (schema, program, options), so the ranks
    // agree on it without talking.
    int gp = (int)c->opt("group.smem_groups", 8);
    bool tier1 = keys_fixed && gp > 0 && c->opt("group.force_global", 0) == 0;
    Program p1 = prog;
    std::vector<int> kslots;
    for (int kc : key_cols) kslots.push_back(p1.slot_for(kc));
    if (tier1) {
      ScanParams F;
      memset(&F, 0, sizeof F);
      F.n_insn = (int32_t)p1.code.size();
      F.stack_depth = p1.max_stack;
      F.n_acc = n_acc;
      F.n_keys = n_group;
      while (gp > 1) {
        F.gp = gp;
        if (scan_smem_bytes(SK_GROUP_SMEM, R, F) <= 100 * 1024) break;
        gp /= 2;
      }
      F.gp = gp;
      tier1 = scan_smem_bytes(SK_GROUP_SMEM, R, F) <= 200 * 1024;
    }
    DevBuf overflow, d_count;
    bool launched1 = false;
    if (tier1 && in.n_rows > 0) {
      bind_program(c, in, p1, P, dp);
      P.n_pred = n_pred;
      P.n_keys = n_group;
      for (int g = 0; g < n_group; ++g) P.key_slots[g] = kslots[g];
      P.gp = gp;
      make_table(gt, rep, acc, (uint32_t)pow2_at_least((uint64_t)c->sm_count * 8 * gp * 2));
      P.gt = gt;
      overflow = DevBuf(c, 256, true);
      P.overflow = overflow.as<uint32_t>();
      launch_scan(c, SK_GROUP_SMEM, R, P, "k_scan_filter_group_smem", &p1);
      // the groups are collected speculatively, so that status and group count come back in ONE host round trip
      pre_rep = DevBuf(c, (size_t)gt.cap * 4 + 256, true);
      pre_slot = DevBuf(c, (size_t)gt.cap * 4 + 256, true);
      d_count = DevBuf(c, 256, true);
      launch_group_collect(c, gt, pre_rep.as<uint32_t>(), pre_slot.as<uint32_t>(), d_count.as<uint32_t>());
      launched1 = true;
    }
    uint32_t err1 = 0, ovf1 = 0, cnt1 = 0;
    bool have_status1 = false;
    bool keys8 = true;
    for (int kc : key_cols) keys8 &= dtype_width(in.cols[kc].dtype) == 8;
    if (merge_ranks && tier1 && keys8 && c->opt("agg.merge_small", 1) != 0) {
      // ---- multi-GPU, few groups (Q1): PartialAggregate(n) -> FinalAggregate (query-distributed/src/planner.rs:200-226)
      // with ONE host round trip.  Every rank packs its status words, up to G_MAX group keys and their raw accumulator
      // rows into a fixed-size block; one all-gather; every rank re-groups the blocks, collects and finalises the
      // result speculatively (outputs sized for world x G_MAX groups) and only then reads the status of all ranks and
      // the final group count.  Any rank that overflowed the tier, holds more groups, or has NULL-able keys sends
      // everybody - the headers are the same on every rank - down the general path below.
      constexpr int G_MAX = 64;
      const int W = c->world;
      const int bw = small_group_block_words(n_group, n_acc, G_MAX);
      DevBuf mine(c, (size_t)bw * 8 + 256, true), all(c, (size_t)W * bw * 8 + 256);
      SmallGroupPack a;
      a.n_keys = n_group;
      a.n_acc = n_acc;
      a.g_max = G_MAX;
      a.block = mine.as<uint64_t>();
      for (int kc : key_cols) a.key_valid |= in.cols[kc].valid != nullptr;
      if (launched1) {
        a.d_count = d_count.as<uint32_t>();
        a.d_err = P.err;
        a.d_ovf = P.overflow;
        a.rep = pre_rep.as<uint32_t>();
        a.slot = pre_slot.as<uint32_t>();
        a.acc = gt.acc;
        for (int g = 0; g < n_group; ++g) a.key[g] = (const uint64_t*)in.cols[key_cols[g]].data;
      }
      launch_pack_small_groups(c, a);
      allgather_device(c, mine.p, all.p, (size_t)bw * 8);
      const int64_t NG = (int64_t)W * G_MAX;
      DevBuf keys_all(c, (size_t)n_group * NG * 8 + 256), rows_all(c, (size_t)NG * n_acc * 8 + 256);
      launch_unpack_small_groups(c, all.as<uint64_t>(), W, n_group, n_acc, G_MAX, keys_all.as<uint64_t>(), rows_all.as<uint64_t>());
      auto gkeys = new_table(c);
      gkeys->n_rows = NG;
      gkeys->batch_offsets = {0, NG};
      std::vector<int> all_keys(n_group);
      for (int k = 0; k < n_group; ++k) {
        all_keys[k] = k;
        Column v;
        v.dtype = in.cols[key_cols[k]].dtype;
        v.n = NG;
        v.data = keys_all.as<uint64_t>() + (size_t)k * NG;
        v.owned = false;
        gkeys->fields.push_back(in.fields[key_cols[k]]);
        gkeys->cols.push_back(v);
      }
      GroupTable g2;
      memset(&g2, 0, sizeof g2);
      g2.n_acc = n_acc;
      g2.n_keys = n_group;
      for (int k = 0; k < n_group; ++k) g2.keys[k] = colref_of(gkeys->cols[k]);
      DevBuf rep2, acc2, d_err2(c, 256, true), d_count2(c, 256, true);
      const bool table_made_before = table_made;
      make_table(g2, rep2, acc2, (uint32_t)pow2_at_least((uint64_t)NG * 2));
      table_made = table_made_before;
      launch_merge_groups(c, g2, rows_all.as<uint64_t>(), NG, prog.acc_kinds.data(), d_err2.as<uint32_t>(), all.as<uint64_t>(), G_MAX, bw);
      DevBuf rep_o(c, (size_t)g2.cap * 4 + 256, true), slot_o(c, (size_t)g2.cap * 4 + 256, true);
      launch_group_collect(c, g2, rep_o.as<uint32_t>(), slot_o.as<uint32_t>(), d_count2.as<uint32_t>());
      {  // speculative result: keys gathered at the representative rows + finalised aggregates, NG rows of room
        auto keys_t = gather_rows(c, *gkeys, all_keys, rep_o.as<uint32_t>(), NG, false);
        for (int k = 0; k < n_group; ++k) {
          Field f = keys_t->fields[k];
          f.name = key_names[k];
          f.nullable = true;
          out->fields.push_back(f);
          out->cols.push_back(keys_t->cols[k]);
          keys_t->cols[k].owned = false;
          keys_t->cols[k].owned_valid = false;
        }
        add_agg_outputs(g2.acc, slot_o.as<uint32_t>(), NG, d_count2.as<uint32_t>());
      }
      uint64_t* h = (uint64_t*)c->scratch((size_t)W * 64 + 64);
      QG_CUDA(cudaMemcpy2DAsync(h, 64, all.p, (size_t)bw * 8, 64, (size_t)W, cudaMemcpyDeviceToHost, c->stream));
      uint32_t* h2 = (uint32_t*)(h + (size_t)W * 8);
      QG_CUDA(cudaMemcpyAsync(h2, d_count2.p, 4, cudaMemcpyDeviceToHost, c->stream));
      QG_CUDA(cudaMemcpyAsync(h2 + 1, d_err2.p, 4, cudaMemcpyDeviceToHost, c->stream));
      QG_CUDA(cudaStreamSynchronize(c->stream));
      bool fallback = (h2[1] & ERR_CAPACITY) != 0;
      uint32_t any_err = 0;
      for (int r = 0; r < W; ++r) {
        const uint64_t* hr = h + (size_t)r * 8;
        if (hr[0] > (uint64_t)G_MAX || hr[2] != 0 || hr[4] != 0 || (hr[1] & ERR_CAPACITY)) fallback = true;
        any_err |= (uint32_t)hr[1];
      }
      const uint64_t* hm = h + (size_t)c->rank * 8;
      err1 = (uint32_t)hm[1];
      ovf1 = (uint32_t)hm[2];
      cnt1 = (uint32_t)hm[0];
      have_status1 = true;
      if (launched1 && pred) note_scan_feedback(c, hm[3], (uint64_t)in.n_rows);
      if (!fallback) {
        raise_device_error(any_err & ~(uint32_t)ERR_CAPACITY);
        const int64_t ng2 = (int64_t)h2[0];
        if (ng2 == 0) {
          out = new_table(c);
          return empty_groups();
        }
        out->n_rows = ng2;
        out->batch_offsets = {0, ng2};
        for (Column& col : out->cols) col.n = ng2;
        return out;
      }
      out = new_table(c);  // drop the speculative result; everybody takes the general path
    }
    if (launched1) {
      if (!have_status1) read_status(P.err, P.overflow, err1, ovf1, d_count.as<uint32_t>(), &cnt1);
      if (!ovf1 && !(err1 & ERR_CAPACITY)) {
        done = true;
        pre_collected = true;
        pre_ng = (int64_t)cnt1;
        local_err = err1;
      } else if (!merge_ranks) {
        raise_device_error(c, err1, in, prog, pred ? n_pred : 0);
      }
    }
    // tier 2 (one GPU): radix-partition the rows twice on the key hash and aggregate every partition in a shared-memory
    // table (ops_groupby.cu) - three streaming passes, no hash table in HBM
    if (!done && !merge_ranks && in.n_rows > 0 && n_group == 1) {
      std::shared_ptr<Table> parted;
      uint32_t perr = 0;
      if (group_by_partitioned(c, in, ap, n_aggr, &parted, &perr)) {
        raise_device_error(c, perr, in, prog, pred ? n_pred : 0);
        return parted ? parted : empty_groups();
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
      if (any_need && total_rows >= c->opt("group.partition_min_rows", 1 << 22)) {
        // first choice: exchange by owner (8-way, long NVLink runs, no remote atomics), then the local two-level radix
        // pipeline with shared-memory tables (ops_groupby.cu)
        int64_t max_rows = 0;
        for (int r = 0; r < c->world; ++r) max_rows = std::max(max_rows, all[(size_t)r * 3 + 2]);
        std::shared_ptr<Table> shard;
        uint32_t xerr = 0;
        if (group_by_exchanged(c, in, ap, n_aggr, total_rows, max_rows, &shard, &xerr)) {
          raise_device_error(xerr);
          return shard ? shard : empty_groups();
        }
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
              v.owned_valid = false;
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
    // tier 3: open-addressing table in HBM.  With many groups the table is far larger than the L2 and every row
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
  if (!merge_ranks) raise_device_error(c, local_err, in, prog, pred ? n_pred : 0);

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
      keys_t->cols[k].owned_valid = false;
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
  if (pre_collected && !exchanged) {
    out_rep = std::move(pre_rep);
    out_slot = std::move(pre_slot);
    ng = pre_ng;
  } else if (table_made && gt.rep != nullptr) {
    ng = collect(gt, out_rep, out_slot);
  }
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
    raise_device_error(c, read_u32(c, P.err), in, prog, 0);
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// Limit (executor.rs:299-341): skip `skip` rows, keep at most `fetch` (negative: no bound) - batch by batch, so the output
// batches are slices of the input batches (empty slices are not emitted); nothing kept => Ok(vec![]).
// ----------------------------------------------------------------------------------------------
std::shared_ptr<Table> op_limit(Ctx* c, const Table& in, int64_t skip, int64_t fetch) {
  const int64_t n = in.n_rows;
  const int64_t start = std::min<int64_t>(std::max<int64_t>(skip, 0), n);
  const int64_t end = fetch < 0 ? n : std::min<int64_t>(n, start + std::min<int64_t>(fetch, n));
  if (start >= end) return nullptr;
  const int64_t k = end - start;
  DevBuf idx(c, pad_bytes((size_t)k * 4));
  launch_iota_u32(c, idx.as<uint32_t>(), k, (uint32_t)start);
  std::vector<int> all(in.cols.size());
  for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
  auto out = gather_rows(c, in, all, idx.as<uint32_t>(), k, false);
  // the kept range cut at the input's batch boundaries
  out->batch_offsets = {0};
  for (size_t b = 1; b < in.batch_offsets.size(); ++b) {
    const int64_t hi = std::min(in.batch_offsets[b], end) - start;
    if (hi > out->batch_offsets.back()) out->batch_offsets.push_back(hi);
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// Sort: what Merge::sorted_merge does after concatenating the partitions (query-distributed/src/operators.rs:141-194:
// lexsort_to_indices over the named columns + take of every column), one output batch.  Also the Tier-2 body of the
// executor's Sort node, which the reference leaves a pass-through (executor.rs:290-297).  `fetch` >= 0 keeps the first
// `fetch` rows only (ORDER BY .. LIMIT n: the gather touches n rows).
// ----------------------------------------------------------------------------------------------
std::shared_ptr<Table> op_sort(Ctx* c, const Table& in, const std::vector<int>& key_cols, const std::vector<int>& descending,
                               const std::vector<int>& nulls_first, int64_t fetch) {
  if ((int64_t)in.batch_offsets.size() - 1 == 0 || in.n_rows == 0) return nullptr;  // operators.rs:150-152
  std::vector<ColRef> keys;
  for (int kc : key_cols) {
    if (kc < 0 || kc >= (int)in.cols.size()) throw_error(QGPU_ERR_EXECUTION, "Column index %d out of bounds", kc);
    if (in.cols[(size_t)kc].dtype == DT_UTF8) throw_error(QGPU_ERR_UNSUPPORTED, "sorting by a Utf8 column is not on the accelerated path");
    keys.push_back(colref_of(in.cols[(size_t)kc]));
  }
  const int64_t n = in.n_rows;
  DevBuf perm(c, pad_bytes((size_t)n * 4));
  sort_permutation(c, keys, descending, nulls_first, n, perm.as<uint32_t>());
  const int64_t k = fetch >= 0 ? std::min(fetch, n) : n;
  if (k == 0) return nullptr;
  std::vector<int> all(in.cols.size());
  for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
  return gather_rows(c, in, all, perm.as<uint32_t>(), k, false);
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
