// cpu_ref.cpp — CPU ORACLE / BASELINE, TEST INFRASTRUCTURE ONLY (never linked into or called by libqgpu).
//
// A C++ restatement ("port") of the reference's executor for the bench workloads, keeping the reference's
// ALGORITHMIC STRUCTURE so that its timing stands in for the Rust binary that cannot be built in this image
// (no cargo/rustc; arrow 53.4.1 is un-vendored):
//
//   * per RecordBatch, recursive expression evaluation that allocates one array per node, a full n-row array per
//     literal and one boolean array per comparison / AND        (crates/query-executor/src/operators.rs:13-62, 322-347, 382-612)
//   * comparisons after int -> f64 coercion through a freshly cast array                         (operators.rs:616-709)
//   * Filter = predicate array, then a gather of EVERY column of the batch (filter_record_batch)  (executor.rs:131-155)
//   * scalar aggregates = per aggregate: evaluate on every batch, concat, reduce                  (executor.rs:167-187, operators.rs:745-848)
//   * single-threaded, like the reference (no rayon use anywhere in query-executor)
//   * GROUP BY and the keyed join do not exist in the reference (executor.rs:189, 363-399): they are restated as the
//     Tier-2 completion with a plain chained hash table — "port" of the semantics, not of code.
//
// Only non-null Int64 / Float64 columns (what the bench configs use); the numpy oracle (oracle/oracle.py) is the
// semantic checker with nulls, other types and error cases.  This file is cross-checked against it in tests/.
//
// `threads` > 1 splits the per-batch work over std::threads (what a rayon par_iter over batches would give): it is
// reported separately as the "all host cores" figure and is NOT what the reference does.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/qgpu.h"

namespace {

enum Ty { T_BOOL = 1, T_I64 = 3, T_F64 = 5 };

struct Arr {
  int ty = 0;
  int64_t n = 0;
  const void* p = nullptr;               // view
  std::shared_ptr<std::vector<uint64_t>> own;  // owner when freshly allocated (8 bytes per element, bools too are 1 byte below)
  std::shared_ptr<std::vector<uint8_t>> ownb;
  const int64_t* i64() const { return (const int64_t*)p; }
  const double* f64() const { return (const double*)p; }
  const uint8_t* b() const { return (const uint8_t*)p; }
};

Arr new_arr(int ty, int64_t n) {
  Arr a;
  a.ty = ty;
  a.n = n;
  if (ty == T_BOOL) {
    a.ownb = std::make_shared<std::vector<uint8_t>>((size_t)n);
    a.p = a.ownb->data();
  } else {
    a.own = std::make_shared<std::vector<uint64_t>>((size_t)n);
    a.p = a.own->data();
  }
  return a;
}

struct Batch {
  std::vector<Arr> cols;
  int64_t n = 0;
};

struct Fail {
  std::string msg;
};

struct Node {
  int kind, op, lit_ty;
  int64_t li;
  double lf;
  int left = -1, right = -1;
};

std::vector<Node> parse(const qgpu_expr* e, int* root) {
  std::vector<Node> nodes;
  std::vector<int> st;
  for (int i = 0; i < e->n_nodes; ++i) {
    const qgpu_expr_node& s = e->nodes[i];
    Node n{};
    n.kind = s.kind;
    n.op = s.op;
    if (s.kind == QGPU_EXPR_LITERAL) {
      n.lit_ty = s.op;
      n.li = s.value.i64;
      n.lf = s.value.f64;
    } else if (s.kind == QGPU_EXPR_BINARY) {
      n.right = st.back(); st.pop_back();
      n.left = st.back(); st.pop_back();
    } else if (s.kind == QGPU_EXPR_UNARY) {
      n.left = st.back(); st.pop_back();
    }
    nodes.push_back(n);
    st.push_back((int)nodes.size() - 1);
  }
  *root = st.back();
  return nodes;
}

Arr cast_f64(const Arr& a) {  // coerce_numeric_types: `x as f64` into a new array
  if (a.ty == T_F64) return a;
  Arr r = new_arr(T_F64, a.n);
  double* o = (double*)r.p;
  const int64_t* v = a.i64();
  for (int64_t i = 0; i < a.n; ++i) o[i] = (double)v[i];
  return r;
}

template <class T, class F>
Arr cmp_arr(const Arr& l, const Arr& r, F f) {
  Arr o = new_arr(T_BOOL, l.n);
  uint8_t* ob = (uint8_t*)o.p;
  const T* a = (const T*)l.p;
  const T* b = (const T*)r.p;
  for (int64_t i = 0; i < l.n; ++i) ob[i] = f(a[i], b[i]) ? 1 : 0;
  return o;
}

int64_t f64_key(double d) {  // IEEE totalOrder key (arrow-rs cmp on floats)
  int64_t b;
  memcpy(&b, &d, 8);
  return b ^ (int64_t)((uint64_t)(b >> 63) >> 1);
}

Arr eval(const std::vector<Node>& nodes, int i, const Batch& b) {
  const Node& n = nodes[i];
  switch (n.kind) {
    case QGPU_EXPR_COLUMN:
      if (n.op < 0 || n.op >= (int)b.cols.size()) throw Fail{"Column index " + std::to_string(n.op) + " out of bounds"};
      return b.cols[n.op];  // Arc clone
    case QGPU_EXPR_LITERAL: {  // create_literal_array: a full n-row array
      if (n.lit_ty == QGPU_INT64) {
        Arr a = new_arr(T_I64, b.n);
        std::fill((int64_t*)a.p, (int64_t*)a.p + b.n, n.li);
        return a;
      }
      if (n.lit_ty == QGPU_FLOAT64) {
        Arr a = new_arr(T_F64, b.n);
        std::fill((double*)a.p, (double*)a.p + b.n, n.lf);
        return a;
      }
      throw Fail{"cpu_ref: literal type not supported by the port"};
    }
    case QGPU_EXPR_BINARY: {
      Arr l = eval(nodes, n.left, b), r = eval(nodes, n.right, b);
      const int64_t N = b.n;
      if (n.op <= QGPU_DIVIDE) {
        static const char* nm[] = {"addition", "subtraction", "multiplication", "division"};
        if (l.ty != r.ty || l.ty == T_BOOL) throw Fail{std::string("Unsupported types for ") + nm[n.op]};
        Arr o = new_arr(l.ty, N);
        if (l.ty == T_F64) {
          const double* a = l.f64();
          const double* c = r.f64();
          double* d = (double*)o.p;
          switch (n.op) {
            case QGPU_ADD: for (int64_t k = 0; k < N; ++k) d[k] = a[k] + c[k]; break;
            case QGPU_SUBTRACT: for (int64_t k = 0; k < N; ++k) d[k] = a[k] - c[k]; break;
            case QGPU_MULTIPLY: for (int64_t k = 0; k < N; ++k) d[k] = a[k] * c[k]; break;
            default: for (int64_t k = 0; k < N; ++k) d[k] = a[k] / c[k]; break;
          }
        } else {  // checked integer kernels
          const int64_t* a = l.i64();
          const int64_t* c = r.i64();
          int64_t* d = (int64_t*)o.p;
          bool ovf = false, dz = false;
          switch (n.op) {
            case QGPU_ADD: for (int64_t k = 0; k < N; ++k) ovf |= __builtin_add_overflow(a[k], c[k], &d[k]); break;
            case QGPU_SUBTRACT: for (int64_t k = 0; k < N; ++k) ovf |= __builtin_sub_overflow(a[k], c[k], &d[k]); break;
            case QGPU_MULTIPLY: for (int64_t k = 0; k < N; ++k) ovf |= __builtin_mul_overflow(a[k], c[k], &d[k]); break;
            default:
              for (int64_t k = 0; k < N; ++k) {
                if (c[k] == 0) { dz = true; d[k] = 0; }
                else if (a[k] == INT64_MIN && c[k] == -1) { ovf = true; d[k] = 0; }
                else d[k] = a[k] / c[k];
              }
          }
          if (dz) throw Fail{"Divide by zero error"};
          if (ovf) throw Fail{"Compute error: Overflow happened on: checked integer arithmetic"};
        }
        return o;
      }
      if (n.op >= QGPU_EQUAL && n.op <= QGPU_GREATER_EQUAL) {
        if (l.ty != r.ty) {  // coerce_numeric_types
          if (l.ty == T_BOOL || r.ty == T_BOOL) throw Fail{"Invalid argument error: Invalid comparison operation"};
          l = cast_f64(l);
          r = cast_f64(r);
        }
        if (l.ty == T_F64) {
          switch (n.op) {
            case QGPU_EQUAL: return cmp_arr<double>(l, r, [](double x, double y) { return f64_key(x) == f64_key(y); });
            case QGPU_NOT_EQUAL: return cmp_arr<double>(l, r, [](double x, double y) { return f64_key(x) != f64_key(y); });
            case QGPU_LESS: return cmp_arr<double>(l, r, [](double x, double y) { return f64_key(x) < f64_key(y); });
            case QGPU_LESS_EQUAL: return cmp_arr<double>(l, r, [](double x, double y) { return f64_key(x) <= f64_key(y); });
            case QGPU_GREATER: return cmp_arr<double>(l, r, [](double x, double y) { return f64_key(x) > f64_key(y); });
            default: return cmp_arr<double>(l, r, [](double x, double y) { return f64_key(x) >= f64_key(y); });
          }
        }
        if (l.ty == T_I64) {
          switch (n.op) {
            case QGPU_EQUAL: return cmp_arr<int64_t>(l, r, [](int64_t x, int64_t y) { return x == y; });
            case QGPU_NOT_EQUAL: return cmp_arr<int64_t>(l, r, [](int64_t x, int64_t y) { return x != y; });
            case QGPU_LESS: return cmp_arr<int64_t>(l, r, [](int64_t x, int64_t y) { return x < y; });
            case QGPU_LESS_EQUAL: return cmp_arr<int64_t>(l, r, [](int64_t x, int64_t y) { return x <= y; });
            case QGPU_GREATER: return cmp_arr<int64_t>(l, r, [](int64_t x, int64_t y) { return x > y; });
            default: return cmp_arr<int64_t>(l, r, [](int64_t x, int64_t y) { return x >= y; });
          }
        }
        throw Fail{"cpu_ref: boolean comparison not supported by the port"};
      }
      if (n.op == QGPU_AND || n.op == QGPU_OR) {
        if (l.ty != T_BOOL || r.ty != T_BOOL) throw Fail{n.op == QGPU_AND ? "AND requires boolean arrays" : "OR requires boolean arrays"};
        Arr o = new_arr(T_BOOL, N);
        uint8_t* d = (uint8_t*)o.p;
        const uint8_t* a = l.b();
        const uint8_t* c = r.b();
        if (n.op == QGPU_AND) for (int64_t k = 0; k < N; ++k) d[k] = a[k] & c[k];
        else for (int64_t k = 0; k < N; ++k) d[k] = a[k] | c[k];
        return o;
      }
      throw Fail{"cpu_ref: binary operator not supported by the port"};
    }
    default: throw Fail{"cpu_ref: expression kind not supported by the port"};
  }
}

// filter_record_batch: gather of every column
Batch filter_batch(const Batch& b, const Arr& mask) {
  if (mask.ty != T_BOOL) throw Fail{"Filter predicate must return boolean"};
  const uint8_t* m = mask.b();
  int64_t k = 0;
  for (int64_t i = 0; i < b.n; ++i) k += m[i];
  Batch o;
  o.n = k;
  for (const Arr& c : b.cols) {
    Arr d = new_arr(c.ty, k);
    uint64_t* dp = (uint64_t*)d.p;
    const uint64_t* sp = (const uint64_t*)c.p;
    int64_t j = 0;
    for (int64_t i = 0; i < b.n; ++i)
      if (m[i]) dp[j++] = sp[i];
    o.cols.push_back(d);
  }
  return o;
}

Arr concat(const std::vector<Arr>& parts) {
  int64_t n = 0;
  for (auto& a : parts) n += a.n;
  Arr o = new_arr(parts[0].ty, n);
  char* d = (char*)o.p;
  const size_t w = parts[0].ty == T_BOOL ? 1 : 8;
  for (auto& a : parts) {
    memcpy(d, a.p, (size_t)a.n * w);
    d += (size_t)a.n * w;
  }
  return o;
}

struct AggOut {
  uint64_t bits;
  uint8_t valid;
};

uint64_t bits_of(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

// evaluate_aggregate over rows idx[0..m) of a (idx == nullptr: all rows)
AggOut reduce(int func, const Arr& a, const int64_t* idx, int64_t m) {
  AggOut o{0, 0};
  if (func == QGPU_COUNT) { o.bits = (uint64_t)m; o.valid = 1; return o; }
  if (m == 0) return o;
  o.valid = 1;
  if (a.ty == T_I64) {
    const int64_t* v = a.i64();
    if (func == QGPU_SUM || func == QGPU_AVG) {
      uint64_t s = 0;
      for (int64_t i = 0; i < m; ++i) s += (uint64_t)v[idx ? idx[i] : i];
      o.bits = func == QGPU_SUM ? s : bits_of((double)(int64_t)s / (double)m);
    } else {
      int64_t r = v[idx ? idx[0] : 0];
      for (int64_t i = 1; i < m; ++i) {
        const int64_t x = v[idx ? idx[i] : i];
        r = func == QGPU_MIN ? std::min(r, x) : std::max(r, x);
      }
      o.bits = (uint64_t)r;
    }
  } else if (a.ty == T_F64) {
    const double* v = a.f64();
    if (func == QGPU_SUM || func == QGPU_AVG) {
      double s = 0;
      for (int64_t i = 0; i < m; ++i) s += v[idx ? idx[i] : i];
      o.bits = bits_of(func == QGPU_SUM ? s : s / (double)m);
    } else {
      double r = v[idx ? idx[0] : 0];
      for (int64_t i = 1; i < m; ++i) {
        const double x = v[idx ? idx[i] : i];
        const bool less = f64_key(x) < f64_key(r);
        if (func == QGPU_MIN ? less : !less) r = x;
      }
      o.bits = bits_of(r);
    }
  } else {
    throw Fail{"Unsupported type for aggregate"};
  }
  return o;
}

std::vector<Batch> make_batches(const void* const* cols, const int32_t* types, int n_cols, int64_t n_rows, int64_t batch_rows) {
  std::vector<Batch> out;
  if (batch_rows <= 0) batch_rows = n_rows > 0 ? n_rows : 1;
  for (int64_t o = 0; o < n_rows; o += batch_rows) {
    Batch b;
    b.n = std::min(batch_rows, n_rows - o);
    for (int c = 0; c < n_cols; ++c) {
      Arr a;
      a.ty = types[c];
      a.n = b.n;
      a.p = (const char*)cols[c] + o * 8;
      b.cols.push_back(a);
    }
    out.push_back(b);
  }
  return out;
}

template <class F>
void parallel_over(size_t n, int threads, F f) {
  if (threads <= 1 || n < 2) {
    for (size_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> ts;
  std::vector<std::string> errs((size_t)threads);
  for (int t = 0; t < threads; ++t)
    ts.emplace_back([&, t] {
      try {
        for (size_t i = (size_t)t; i < n; i += (size_t)threads) f(i);
      } catch (const Fail& e) {
        errs[(size_t)t] = e.msg;
      }
    });
  for (auto& t : ts) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw Fail{e};
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* cpuref_last_error() { return g_err.c_str(); }

// Filter (optional) + HashAggregate over non-null Int64 / Float64 columns.
//   types[c]: qgpu_dtype of column c (QGPU_INT64 / QGPU_FLOAT64)
//   out_keys:  [cap][n_group] raw 8-byte key values       out_vals: [cap][n_aggr] raw 8-byte results
//   out_valid: [cap][n_aggr]                               *out_groups: rows produced (-1: the reference's Ok(vec![]))
// Returns 0, or 1 with cpuref_last_error() set (the QueryError text).
int cpuref_filter_aggregate(const void* const* cols, const int32_t* types, int32_t n_cols, int64_t n_rows, int64_t batch_rows,
                            const qgpu_expr* pred, const qgpu_expr* groups, int32_t n_group, const qgpu_agg_expr* aggs,
                            int32_t n_aggr, int32_t threads, int64_t cap, uint64_t* out_keys, uint64_t* out_vals,
                            uint8_t* out_valid, int64_t* out_groups) {
  try {
    *out_groups = -1;
    std::vector<Batch> in = make_batches(cols, types, n_cols, n_rows, batch_rows);
    // ---- Filter (executor.rs:131-155)
    std::vector<Batch> batches;
    if (pred) {
      int root;
      std::vector<Node> pn = parse(pred, &root);
      std::vector<Batch> tmp(in.size());
      parallel_over(in.size(), threads, [&](size_t i) { tmp[i] = filter_batch(in[i], eval(pn, root, in[i])); });
      for (auto& b : tmp)
        if (b.n > 0) batches.push_back(std::move(b));
    } else {
      batches = in;
    }
    if (n_aggr == 0 || batches.empty()) return 0;
    // ---- aggregate inputs: per aggregate, evaluate on every batch, concat (executor.rs:170-180)
    std::vector<Arr> agg_in((size_t)n_aggr);
    for (int a = 0; a < n_aggr; ++a) {
      int root;
      std::vector<Node> an = parse(&aggs[a].expr, &root);
      std::vector<Arr> parts(batches.size());
      parallel_over(batches.size(), threads, [&](size_t i) { parts[i] = eval(an, root, batches[i]); });
      agg_in[(size_t)a] = concat(parts);
    }
    if (n_group == 0) {
      if (cap < 1) throw Fail{"cpu_ref: output capacity too small"};
      for (int a = 0; a < n_aggr; ++a) {
        const AggOut o = reduce(aggs[a].func, agg_in[(size_t)a], nullptr, agg_in[(size_t)a].n);
        out_vals[a] = o.bits;
        out_valid[a] = o.valid;
      }
      *out_groups = 1;
      return 0;
    }
    // ---- Tier-2 GROUP BY: chained hash table over the concatenated key columns
    std::vector<Arr> keys((size_t)n_group);
    for (int g = 0; g < n_group; ++g) {
      int root;
      std::vector<Node> gn = parse(&groups[g], &root);
      std::vector<Arr> parts(batches.size());
      parallel_over(batches.size(), threads, [&](size_t i) { parts[i] = eval(gn, root, batches[i]); });
      keys[(size_t)g] = concat(parts);
    }
    const int64_t n = keys[0].n;
    struct KeyHash {
      size_t operator()(const std::vector<uint64_t>& k) const {
        uint64_t h = 0x9e3779b97f4a7c15ull;
        for (uint64_t v : k) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
      }
    };
    std::unordered_map<std::vector<uint64_t>, int64_t, KeyHash> index;
    std::vector<std::vector<int64_t>> rows_of;
    std::vector<uint64_t> k((size_t)n_group);
    for (int64_t i = 0; i < n; ++i) {
      for (int g = 0; g < n_group; ++g) k[(size_t)g] = ((const uint64_t*)keys[(size_t)g].p)[i];
      auto it = index.find(k);
      if (it == index.end()) {
        it = index.emplace(k, (int64_t)rows_of.size()).first;
        rows_of.emplace_back();
      }
      rows_of[(size_t)it->second].push_back(i);
    }
    const int64_t ng = (int64_t)rows_of.size();
    if (ng > cap) throw Fail{"cpu_ref: output capacity too small"};
    for (auto& kv : index)
      for (int g = 0; g < n_group; ++g) out_keys[kv.second * n_group + g] = kv.first[(size_t)g];
    parallel_over((size_t)ng, threads, [&](size_t gi) {
      for (int a = 0; a < n_aggr; ++a) {
        const AggOut o = reduce(aggs[a].func, agg_in[(size_t)a], rows_of[gi].data(), (int64_t)rows_of[gi].size());
        out_vals[(int64_t)gi * n_aggr + a] = o.bits;
        out_valid[(int64_t)gi * n_aggr + a] = o.valid;
      }
    });
    *out_groups = ng;
    return 0;
  } catch (const Fail& e) {
    g_err = e.msg;
    return 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Tier-2 equi-join on one Int64 key (INNER = 0, LEFT = 1), probe = left, build = right; output in probe order.
// out_left / out_right: row index pairs (out_right = -1 for an unmatched LEFT row); returns the pair count, or -1 if it
// exceeds cap.
int64_t cpuref_hash_join(const int64_t* lkey, int64_t nl, const int64_t* rkey, int64_t nr, int32_t join_type, int64_t cap,
                         int64_t* out_left, int64_t* out_right) {
  std::unordered_map<int64_t, std::vector<int64_t>> table;
  table.reserve((size_t)nr * 2);
  for (int64_t j = 0; j < nr; ++j) table[rkey[j]].push_back(j);
  int64_t n = 0;
  for (int64_t i = 0; i < nl; ++i) {
    auto it = table.find(lkey[i]);
    if (it != table.end()) {
      for (int64_t j : it->second) {
        if (n >= cap) return -1;
        out_left[n] = i;
        out_right[n] = j;
        ++n;
      }
    } else if (join_type == 1) {
      if (n >= cap) return -1;
      out_left[n] = i;
      out_right[n] = -1;
      ++n;
    }
  }
  return n;
}

// gather of one 8-byte column through an index vector (arrow `take`; -1 -> 0 / null)
void cpuref_take8(const uint64_t* src, const int64_t* idx, int64_t n, uint64_t* dst) {
  for (int64_t i = 0; i < n; ++i) dst[i] = idx[i] < 0 ? 0 : src[idx[i]];
}

}  // extern "C"
