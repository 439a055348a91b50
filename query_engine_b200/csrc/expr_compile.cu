// expr_compile.cu — host side: PhysicalExpr (postfix) -> typed device program.
//
// Type rules and error texts mirror the reference's evaluator so that the Rust shim can surface the
// same QueryError the CPU path would (crates/query-executor/src/operators.rs:13-62, 349-380, 382-612,
// 616-709, 711-743, 745-848).  Everything here is plain host C++; no GPU is touched.
#include <cstring>
#include <sstream>

#include "interp.cuh"
#include "qgpu_internal.h"

namespace qgpu {

void throw_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error{code, std::string(buf)};
}

static const char* kUnsupportedMsg[] = {
    "Scalar subquery execution not yet implemented in synchronous context",   // ScalarSubquery  operators.rs:34-40
    "IN subquery execution not yet implemented in synchronous context",        // InSubquery      :41-46
    "EXISTS subquery execution not yet implemented in synchronous context",    // Exists          :47-52
    "Window function execution not yet implemented",                           // WindowFunction  :53-59
    "Scalar functions are evaluated by the reference's CPU path (string/math functions are not on the accelerated path)",
};

ExprTree parse_postfix(const qgpu_expr* e) {
  ExprTree t;
  if (!e || !e->nodes || e->n_nodes <= 0) throw_error(QGPU_ERR_INVALID, "empty expression");
  std::vector<int> stack;
  for (int i = 0; i < e->n_nodes; ++i) {
    const qgpu_expr_node& s = e->nodes[i];
    ExprNode n;
    n.kind = s.kind;
    n.op = s.op;
    if (s.str) n.str = s.str;
    switch (s.kind) {
      case QGPU_EXPR_COLUMN: break;
      case QGPU_EXPR_LITERAL:
        n.lit_type = s.op;
        n.is_null = s.is_null != 0 || s.op == QGPU_NULL;
        n.imm = s.value.u64;
        if (s.op == QGPU_FLOAT32) n.imm = f32_bits(s.value.f32);
        if (s.op == QGPU_BOOL) n.imm = s.value.b ? 1 : 0;
        if (s.op == QGPU_INT32) n.imm = (uint64_t)(int64_t)(int32_t)s.value.i64;
        break;
      case QGPU_EXPR_BINARY:
        if (stack.size() < 2) throw_error(QGPU_ERR_INVALID, "malformed postfix expression (binary)");
        n.right = stack.back(); stack.pop_back();
        n.left = stack.back(); stack.pop_back();
        break;
      case QGPU_EXPR_UNARY:
        if (stack.empty()) throw_error(QGPU_ERR_INVALID, "malformed postfix expression (unary)");
        n.left = stack.back(); stack.pop_back();
        break;
      case QGPU_EXPR_UNSUPPORTED: break;
      default: throw_error(QGPU_ERR_INVALID, "unknown expression node kind %d", s.kind);
    }
    t.nodes.push_back(n);
    stack.push_back((int)t.nodes.size() - 1);
  }
  if (stack.size() != 1) throw_error(QGPU_ERR_INVALID, "malformed postfix expression (%zu roots)", stack.size());
  t.root = stack.back();
  return t;
}

std::string ExprTree::canon(int i) const {
  const ExprNode& n = nodes[i];
  std::ostringstream o;
  switch (n.kind) {
    case QGPU_EXPR_COLUMN: o << "c" << n.op; break;
    case QGPU_EXPR_LITERAL: o << "l" << n.lit_type << ":" << (n.is_null ? "null" : "") << n.imm << ":" << n.str; break;
    case QGPU_EXPR_BINARY: o << "(" << canon(n.left) << " b" << n.op << " " << canon(n.right) << ")"; break;
    case QGPU_EXPR_UNARY: o << "(u" << n.op << " " << canon(n.left) << ")"; break;
    default: o << "x" << n.op; break;
  }
  return o.str();
}

static bool is_num(uint8_t t) { return t == DT_I64 || t == DT_I32 || t == DT_F64 || t == DT_F32; }

static const char* dt_name(uint8_t t) {
  switch (t) {
    case DT_NULL: return "Null";
    case DT_BOOL: return "Boolean";
    case DT_I32: return "Int32";
    case DT_I64: return "Int64";
    case DT_F32: return "Float32";
    case DT_F64: return "Float64";
    case DT_UTF8: return "Utf8";
    default: return "Other";
  }
}

static uint8_t lit_dt(const ExprNode& n) {
  if (n.is_null) return DT_NULL;  // every X(None) and ScalarValue::Null become a NullArray (operators.rs:344-345)
  switch (n.lit_type) {
    case QGPU_BOOL: return DT_BOOL;
    case QGPU_INT32: return DT_I32;
    case QGPU_INT64: return DT_I64;
    case QGPU_FLOAT32: return DT_F32;
    case QGPU_FLOAT64: return DT_F64;
    case QGPU_UTF8: return DT_UTF8;
    default: return DT_OTHER;  // Int8/16, UInt*: arrays none of the ladders accept
  }
}

// result type of coerce_numeric_types (operators.rs:616-675); DT_OTHER = left unchanged (mismatch)
static uint8_t coerce(uint8_t l, uint8_t r) {
  if (l == r) return l;
  if (l == DT_F64 && (r == DT_I64 || r == DT_I32)) return DT_F64;
  if (r == DT_F64 && (l == DT_I64 || l == DT_I32)) return DT_F64;
  if (l == DT_F32 || r == DT_F32) {
    // both sides through cast_to_float64, which only converts F64/F32/I64/I32
    const bool lc = is_num(l), rc = is_num(r);
    return (lc && rc) ? DT_F64 : DT_OTHER;
  }
  if ((l == DT_I64 && r == DT_I32) || (l == DT_I32 && r == DT_I64)) return DT_I64;
  return DT_OTHER;
}

static uint8_t infer_node(ExprTree& t, int i, const SchemaView& s) {
  ExprNode& n = t.nodes[i];
  switch (n.kind) {
    case QGPU_EXPR_COLUMN:
      if (n.op < 0 || n.op >= (int)s.dtypes.size())
        throw_error(QGPU_ERR_EXECUTION, "Column index %d out of bounds", n.op);  // operators.rs:16-21
      n.type = s.dtypes[n.op];
      n.maybe_null = s.has_nulls[n.op];
      return n.type;
    case QGPU_EXPR_LITERAL:
      n.type = lit_dt(n);
      n.maybe_null = n.type == DT_NULL;
      return n.type;
    case QGPU_EXPR_UNSUPPORTED: {
      int v = n.op < 0 || n.op > 4 ? 4 : n.op;
      throw_error(v == 4 ? QGPU_ERR_UNSUPPORTED : QGPU_ERR_EXECUTION, "%s", kUnsupportedMsg[v]);
    }
    case QGPU_EXPR_UNARY: {
      uint8_t c = infer_node(t, n.left, s);
      ExprNode& nn = t.nodes[i];
      nn.maybe_null = t.nodes[nn.left].maybe_null;
      if (nn.op == QGPU_NOT) {
        if (c != DT_BOOL) throw_error(QGPU_ERR_EXECUTION, "NOT operator requires boolean array");
        return nn.type = DT_BOOL;
      }
      if (nn.op == QGPU_CAST_FLOAT64) {
        if (!is_num(c)) throw_error(QGPU_ERR_EXECUTION, "Unsupported type for cast to Float64");
        return nn.type = DT_F64;
      }
      if (!is_num(c)) throw_error(QGPU_ERR_EXECUTION, "Unsupported type for negation");
      return nn.type = c;
    }
    case QGPU_EXPR_BINARY: {
      uint8_t l = infer_node(t, n.left, s);
      uint8_t r = infer_node(t, t.nodes[i].right, s);
      ExprNode& nn = t.nodes[i];
      nn.maybe_null = t.nodes[nn.left].maybe_null || t.nodes[nn.right].maybe_null;
      switch (nn.op) {
        case QGPU_ADD: case QGPU_SUBTRACT: case QGPU_MULTIPLY: case QGPU_DIVIDE: {
          static const char* nm[] = {"addition", "subtraction", "multiplication", "division"};
          if (l != r || !is_num(l)) throw_error(QGPU_ERR_EXECUTION, "Unsupported types for %s", nm[nn.op]);
          return nn.type = l;
        }
        case QGPU_MODULO:
          if (l != r || (l != DT_I64 && l != DT_I32))
            throw_error(QGPU_ERR_EXECUTION, "Modulo operation requires integer arrays");
          nn.maybe_null = true;  // r == 0 -> NULL
          return nn.type = l;
        case QGPU_EQUAL: case QGPU_NOT_EQUAL: case QGPU_LESS: case QGPU_LESS_EQUAL: case QGPU_GREATER:
        case QGPU_GREATER_EQUAL: {
          static const char* sym[] = {"==", "!=", "<", "<=", ">", ">="};
          uint8_t c = coerce(l, r);
          if (c == DT_OTHER)
            throw_error(QGPU_ERR_ARROW, "Invalid argument error: Invalid comparison operation: %s %s %s", dt_name(l),
                        sym[nn.op - QGPU_EQUAL], dt_name(r));
          if (c == DT_UTF8)
            throw_error(QGPU_ERR_UNSUPPORTED, "Utf8 comparison is not on the accelerated path");
          return nn.type = DT_BOOL;
        }
        case QGPU_AND: case QGPU_OR:
          if (l != DT_BOOL || r != DT_BOOL)
            throw_error(QGPU_ERR_EXECUTION, "%s requires boolean arrays", nn.op == QGPU_AND ? "AND" : "OR");
          return nn.type = DT_BOOL;
        case QGPU_TS_MATCH:
          throw_error(QGPU_ERR_UNSUPPORTED, "@@ (full-text match) is not on the accelerated path");
        default: throw_error(QGPU_ERR_INVALID, "unknown binary operator %d", nn.op);
      }
    }
    default: throw_error(QGPU_ERR_INVALID, "bad node kind");
  }
}

uint8_t infer_types(ExprTree& t, const SchemaView& s) { return infer_node(t, t.root, s); }

int Program::slot_for(int table_col) {
  for (size_t i = 0; i < col_of_slot.size(); ++i)
    if (col_of_slot[i] == table_col) return (int)i;
  if ((int)col_of_slot.size() >= MAX_COLS)
    throw_error(QGPU_ERR_UNSUPPORTED, "more than %d distinct columns in one fused program", MAX_COLS);
  col_of_slot.push_back(table_col);
  return (int)col_of_slot.size() - 1;
}

// ---- code generation ------------------------------------------------------------------------
namespace {

struct Gen {
  ExprTree& t;
  Program& p;
  int depth = 0;

  static uint8_t op_of_binary(int b) {
    switch (b) {
      case QGPU_ADD: return OP_ADD;
      case QGPU_SUBTRACT: return OP_SUB;
      case QGPU_MULTIPLY: return OP_MUL;
      case QGPU_DIVIDE: return OP_DIV;
      case QGPU_MODULO: return OP_MOD;
      case QGPU_EQUAL: return OP_EQ;
      case QGPU_NOT_EQUAL: return OP_NE;
      case QGPU_LESS: return OP_LT;
      case QGPU_LESS_EQUAL: return OP_LE;
      case QGPU_GREATER: return OP_GT;
      case QGPU_GREATER_EQUAL: return OP_GE;
      case QGPU_AND: return OP_AND;
      default: return OP_OR;
    }
  }

  bool is_leaf(int i) const {
    int k = t.nodes[i].kind;
    return k == QGPU_EXPR_COLUMN || k == QGPU_EXPR_LITERAL;
  }

  static uint64_t lit_convert(const ExprNode& n, uint8_t to) {
    if (n.type == DT_NULL) return 0;
    return convert(n.imm, n.type, to);
  }

  // fills the operand fields of `in` from leaf node i, converted to compute type ct
  void set_operand(Insn& in, int i, uint8_t ct) {
    const ExprNode& n = t.nodes[i];
    if (n.kind == QGPU_EXPR_COLUMN) {
      in.mode = M_COL;
      in.arg = p.slot_for(n.op);
      in.flags |= (uint8_t)(n.type << 4);
    } else {
      in.mode = M_LIT;
      in.imm = lit_convert(n, ct);
      if (n.type == DT_NULL) in.flags |= F_LITNULL;
    }
  }

  void emit(Insn in) { p.code.push_back(in); }

  void cast_acc(uint8_t from, uint8_t to) {
    if (from == to || from == DT_NULL) return;
    if (from == DT_I32 && to == DT_I64) return;  // representation already sign-extended
    Insn c{};
    c.op = OP_CAST;
    c.ty = from;
    c.arg = to;
    emit(c);
  }

  // leaves node i's value in acc, in the node's own type
  void gen(int i) {
    const ExprNode n = t.nodes[i];
    switch (n.kind) {
      case QGPU_EXPR_COLUMN:
      case QGPU_EXPR_LITERAL: {
        Insn in{};
        in.op = OP_LD;
        in.ty = n.type == DT_NULL ? (uint8_t)DT_BOOL : n.type;
        set_operand(in, i, in.ty);
        emit(in);
        return;
      }
      case QGPU_EXPR_UNARY: {
        gen(n.left);
        if (n.op == QGPU_CAST_FLOAT64) {
          cast_acc(t.nodes[n.left].type, DT_F64);
          return;
        }
        Insn in{};
        in.op = n.op == QGPU_NOT ? OP_NOT : OP_NEG;
        in.ty = n.type;
        emit(in);
        return;
      }
      case QGPU_EXPR_BINARY: {
        const uint8_t lt = t.nodes[n.left].type, rt = t.nodes[n.right].type;
        uint8_t ct = lt;  // compute type
        const bool is_cmp = n.op >= QGPU_EQUAL && n.op <= QGPU_GREATER_EQUAL;
        if (is_cmp) ct = coerce(lt, rt);
        if (ct == DT_NULL) {
          // NullArray cmp NullArray: all-null boolean array
          Insn in{};
          in.op = OP_LD;
          in.ty = DT_BOOL;
          in.mode = M_LIT;
          in.flags = F_LITNULL;
          emit(in);
          return;
        }
        Insn in{};
        in.op = op_of_binary(n.op);
        in.ty = ct;
        if (is_leaf(n.right)) {
          gen(n.left);
          cast_acc(lt, ct);
          set_operand(in, n.right, ct);
        } else if (is_leaf(n.left)) {
          gen(n.right);
          cast_acc(rt, ct);
          in.flags |= F_REV;
          set_operand(in, n.left, ct);
        } else {
          gen(n.left);
          cast_acc(lt, ct);
          Insn push{};
          push.op = OP_PUSH;
          emit(push);
          if (++depth > p.max_stack) p.max_stack = depth;
          if (depth > MAX_STACK) throw_error(QGPU_ERR_UNSUPPORTED, "expression too deep for the fused evaluator");
          gen(n.right);
          cast_acc(rt, ct);
          in.flags |= F_REV;
          in.mode = M_STK;
          --depth;
        }
        emit(in);
        return;
      }
      default: throw_error(QGPU_ERR_INVALID, "bad node");
    }
  }

  void conjuncts(int i, std::vector<int>& out) {
    const ExprNode& n = t.nodes[i];
    if (n.kind == QGPU_EXPR_BINARY && n.op == QGPU_AND) {
      conjuncts(n.left, out);
      conjuncts(n.right, out);
    } else {
      out.push_back(i);
    }
  }
};

int flip_cmp(int op) {
  switch (op) {
    case OP_LT: return OP_GT;
    case OP_LE: return OP_GE;
    case OP_GT: return OP_LT;
    case OP_GE: return OP_LE;
    default: return op;
  }
}

}  // namespace

// Predicate section.  A row passes a filter iff the predicate is TRUE and VALID (arrow `filter` treats a
// null mask slot as false; executor.rs:147).  With non-Kleene AND (null if either side null) that is
// exactly "every top-level conjunct is true and valid", so the AND chain is evaluated conjunct by
// conjunct into the `pass` mask; every conjunct is still evaluated on every row (no short circuit), so
// checked-arithmetic errors surface exactly as in the reference.
void compile_predicate(ExprTree& t, const SchemaView& s, Program& p) {
  uint8_t ty = infer_types(t, s);
  if (ty != DT_BOOL) throw_error(QGPU_ERR_EXECUTION, "Filter predicate must return boolean");  // executor.rs:140-145
  Gen g{t, p};
  std::vector<int> cj;
  g.conjuncts(t.root, cj);
  for (int i : cj) {
    const ExprNode& n = t.nodes[i];
    bool fused = false;
    if (n.kind == QGPU_EXPR_BINARY && n.op >= QGPU_EQUAL && n.op <= QGPU_GREATER_EQUAL) {
      const ExprNode& a = t.nodes[n.left];
      const ExprNode& b = t.nodes[n.right];
      const ExprNode* colp = nullptr;
      const ExprNode* litp = nullptr;
      bool flipped = false;
      if (a.kind == QGPU_EXPR_COLUMN && b.kind == QGPU_EXPR_LITERAL) { colp = &a; litp = &b; }
      if (b.kind == QGPU_EXPR_COLUMN && a.kind == QGPU_EXPR_LITERAL) { colp = &b; litp = &a; flipped = true; }
      if (colp && litp && litp->type != DT_NULL) {
        uint8_t ct = coerce(a.type, b.type);
        Insn in{};
        in.op = OP_CMPCL;
        in.ty = ct;
        int cmp = Gen::op_of_binary(n.op);
        in.mode = (uint8_t)(flipped ? flip_cmp(cmp) : cmp);
        in.arg = p.slot_for(colp->op);
        in.flags = (uint8_t)(colp->type << 4);
        in.imm = (uint64_t)cmp_key(convert(litp->imm, litp->type, ct), ct);
        p.code.push_back(in);
        fused = true;
      }
    }
    if (!fused) {
      g.gen(i);
      Insn pr{};
      pr.op = OP_PRED;
      p.code.push_back(pr);
    }
  }
  Insn gate{};
  gate.op = OP_GATE;
  p.code.push_back(gate);
}

uint8_t compile_value(ExprTree& t, const SchemaView& s, Program& p) {
  uint8_t ty = infer_types(t, s);
  Gen g{t, p};
  g.gen(t.root);
  return ty;
}

// Aggregate section (evaluate_aggregate, operators.rs:745-848).  Identical input expressions are
// evaluated once; SUM and AVG of the same input share the sum slot; inputs that cannot be NULL share
// one row counter.
void compile_aggregates(std::vector<ExprTree>& trees, const std::vector<int>& funcs, const SchemaView& s, Program& p) {
  struct Group {
    std::string key;
    int tree;
    std::vector<int> members;
  };
  std::vector<Group> groups;
  std::vector<uint8_t> types(trees.size());
  for (size_t i = 0; i < trees.size(); ++i) {
    types[i] = infer_types(trees[i], s);
    const int f = funcs[i];
    const uint8_t t = types[i];
    if (f != QGPU_COUNT && !is_num(t)) {
      static const char* nm[] = {"COUNT", "SUM", "AVG", "MIN", "MAX"};
      throw_error(QGPU_ERR_EXECUTION, "Unsupported type for %s", nm[f]);
    }
    std::string k = trees[i].canon(trees[i].root);
    bool found = false;
    for (auto& g : groups)
      if (g.key == k) { g.members.push_back((int)i); found = true; break; }
    if (!found) groups.push_back(Group{k, (int)i, {(int)i}});
  }
  p.agg_out.assign(trees.size(), AggOutput{});
  int rows_slot = -1;  // shared counter for inputs that are never NULL
  auto new_slot = [&](uint8_t kind) {
    if ((int)p.acc_kinds.size() >= MAX_ACC) throw_error(QGPU_ERR_UNSUPPORTED, "more than %d accumulators", MAX_ACC);
    p.acc_kinds.push_back(kind);
    return (int)p.acc_kinds.size() - 1;
  };
  for (auto& g : groups) {
    ExprTree& tr = trees[g.tree];
    const uint8_t ty = types[g.tree];
    const bool maybe_null = tr.nodes[tr.root].maybe_null || ty == DT_NULL;
    Gen gen{tr, p};
    bool only_count = true;
    for (int m : g.members) only_count &= funcs[m] == QGPU_COUNT;
    const int rk = tr.nodes[tr.root].kind;
    if (only_count && !maybe_null && (rk == QGPU_EXPR_COLUMN || rk == QGPU_EXPR_LITERAL)) {
      // COUNT of a column / literal that is never NULL = number of rows: do not read the column at all
      Insn one{};
      one.op = OP_LD;
      one.ty = DT_BOOL;
      one.mode = M_LIT;
      one.imm = 1;
      p.code.push_back(one);
    } else {
      gen.gen(tr.root);
    }
    int cnt_slot = -1, sum_slot = -1, min_slot = -1, max_slot = -1;
    auto emit_agg = [&](uint8_t kind, int slot) {
      Insn in{};
      in.op = OP_AGG;
      in.ty = ty == DT_NULL ? (uint8_t)DT_BOOL : ty;
      in.mode = kind;
      in.arg = slot;
      p.code.push_back(in);
    };
    // every aggregate needs the non-null count (COUNT itself; SUM/MIN/MAX -> NULL when 0; AVG divisor)
    if (!maybe_null) {
      if (rows_slot < 0) {
        rows_slot = new_slot(ACC_COUNT);
        emit_agg(ACC_COUNT, rows_slot);
      }
      cnt_slot = rows_slot;
    } else {
      cnt_slot = new_slot(ACC_COUNT);
      emit_agg(ACC_COUNT, cnt_slot);
    }
    for (int m : g.members) {
      AggOutput& o = p.agg_out[m];
      o.func = funcs[m];
      o.in_type = ty;
      o.cnt_slot = cnt_slot;
      o.val_slot = -1;
      switch (funcs[m]) {
        case QGPU_COUNT: o.out_type = DT_I64; break;
        case QGPU_SUM:
        case QGPU_AVG: {
          const bool isint = ty == DT_I64 || ty == DT_I32;
          if (sum_slot < 0) {
            sum_slot = new_slot(isint ? ACC_ADD_I64 : ACC_ADD_F64);
            emit_agg(isint ? ACC_ADD_I64 : ACC_ADD_F64, sum_slot);
          }
          o.val_slot = sum_slot;
          o.out_type = funcs[m] == QGPU_AVG ? DT_F64 : (isint ? DT_I64 : DT_F64);
          break;
        }
        case QGPU_MIN:
          if (min_slot < 0) { min_slot = new_slot(ACC_MIN_I64); emit_agg(ACC_MIN_I64, min_slot); }
          o.val_slot = min_slot;
          o.out_type = ty;
          break;
        case QGPU_MAX:
          if (max_slot < 0) { max_slot = new_slot(ACC_MAX_I64); emit_agg(ACC_MAX_I64, max_slot); }
          o.val_slot = max_slot;
          o.out_type = ty;
          break;
        default: throw_error(QGPU_ERR_INVALID, "unknown aggregate function %d", funcs[m]);
      }
    }
  }
}

static const char* op_name(uint8_t op) {
  static const char* n[] = {"LD", "ADD", "SUB", "MUL", "DIV", "MOD", "EQ", "NE", "LT", "LE", "GT", "GE", "AND",
                            "OR", "NOT", "NEG", "CAST", "PUSH", "PRED", "CMPCL", "GATE", "AGG", "STORE", "KEY", "NOP"};
  return op <= OP_NOP ? n[op] : "?";
}

std::string Program::listing() const {
  std::ostringstream o;
  for (size_t i = 0; i < code.size(); ++i) {
    const Insn& in = code[i];
    o << i << ": " << op_name(in.op) << " ty=" << dt_name(in.ty);
    if (in.op == OP_CMPCL) o << " cmp=" << op_name(in.mode) << " col=" << in.arg << " sty=" << dt_name(insn_sty(in)) << " key=" << (int64_t)in.imm;
    else if (in.op == OP_AGG) o << " kind=" << (int)in.mode << " slot=" << in.arg;
    else if (in.op == OP_CAST) o << " to=" << dt_name((uint8_t)in.arg);
    else if (in.op == OP_STORE) o << " out=" << in.arg;
    else if (in.mode == M_COL) o << " col=" << in.arg << " sty=" << dt_name(insn_sty(in));
    else if (in.mode == M_LIT) o << " lit=" << (int64_t)in.imm << ((in.flags & F_LITNULL) ? " (null)" : "");
    else if (in.mode == M_STK) o << " stk";
    if (in.flags & F_REV) o << " rev";
    o << "\n";
  }
  o << "cols:";
  for (int c : col_of_slot) o << " " << c;
  o << "\nstack=" << max_stack << " acc=" << acc_kinds.size() << "\n";
  return o.str();
}

}  // namespace qgpu
