// interp.cuh — the per-thread evaluator of compiled PhysicalExpr programs.
//
// Semantics restated from the reference (crates/query-executor/src/operators.rs):
//   * arithmetic only on identical types, integer ops CHECKED (overflow / div-by-zero raise)      :384-507
//   * modulo: r == 0 -> NULL, sign of the dividend                                               :711-743
//   * comparisons after coerce_numeric_types, floats ordered by IEEE totalOrder (arrow-rs cmp)   :509-538, 616-675
//   * AND / OR are not Kleene: NULL if either side is NULL                                       :539-570
//   * unary minus wraps (release build)                                                          :360-378
//
// One thread evaluates R rows at a time ("lanes"); the program is uniform across the grid so the
// switch dispatch never diverges.  The same code compiles for the host (tests/native checks the
// evaluator against the oracle without a GPU) — the host build is test scaffolding, not a fallback.
#pragma once
#include "common.cuh"

namespace qgpu {

template <int R>
struct Lanes {
  uint64_t v[R];
  uint32_t ok;  // bit r = row r valid (non-null)
};

QG_HD int64_t as_i64(uint64_t x) { return (int64_t)x; }

QG_HD double bits_f64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)x);
#else
  double d;
  __builtin_memcpy(&d, &x, 8);
  return d;
#endif
}
QG_HD uint64_t f64_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t x;
  __builtin_memcpy(&x, &d, 8);
  return x;
#endif
}
QG_HD float bits_f32(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float((uint32_t)x);
#else
  uint32_t u = (uint32_t)x;
  float f;
  __builtin_memcpy(&f, &u, 4);
  return f;
#endif
}
QG_HD uint64_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__float_as_uint(f);
#else
  uint32_t u;
  __builtin_memcpy(&u, &f, 4);
  return (uint64_t)u;
#endif
}

// IEEE-754 totalOrder as a signed integer key: -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN
QG_HD int64_t f64_order_key(uint64_t bits) {
  int64_t b = (int64_t)bits;
  return b ^ (int64_t)((uint64_t)(b >> 63) >> 1);
}
QG_HD uint64_t f64_from_order_key(int64_t k) { return (uint64_t)(k ^ (int64_t)((uint64_t)(k >> 63) >> 1)); }
QG_HD int64_t f32_order_key(uint64_t bits) {
  int32_t b = (int32_t)(uint32_t)bits;
  return (int64_t)(b ^ (int32_t)((uint32_t)(b >> 31) >> 1));
}
QG_HD uint64_t f32_from_order_key(int64_t k) {
  int32_t b = (int32_t)k;
  return (uint64_t)(uint32_t)(b ^ (int32_t)((uint32_t)(b >> 31) >> 1));
}

// value in compute type `ty` -> orderable / comparable signed key
QG_HD int64_t cmp_key(uint64_t v, uint8_t ty) {
  if (ty == DT_F64) return f64_order_key(v);
  if (ty == DT_F32) return f32_order_key(v);
  return (int64_t)v;  // BOOL, I32 (sign-extended), I64
}

// storage -> compute type conversion done by coerce_numeric_types / cast_to_float64 (operators.rs:616-709)
QG_HD uint64_t convert(uint64_t v, uint8_t from, uint8_t to) {
  if (from == to) return v;
  if (to == DT_F64) {
    if (from == DT_I64 || from == DT_I32) return f64_bits((double)(int64_t)v);  // `x as f64`
    if (from == DT_F32) return f64_bits((double)bits_f32(v));
  }
  if (to == DT_I64 && from == DT_I32) return v;  // already sign-extended
  return v;
}

QG_HD bool mul_overflow_i64(int64_t a, int64_t b, int64_t* r) {
#if defined(__CUDA_ARCH__)
  int64_t lo = (int64_t)((uint64_t)a * (uint64_t)b);
  int64_t hi = __mul64hi(a, b);
  *r = lo;
  return hi != (lo >> 63);
#else
  return __builtin_mul_overflow(a, b, r);
#endif
}

// one checked / IEEE arithmetic op on one row; returns result, raises err bits, may clear *ok (modulo by 0)
QG_HD uint64_t arith1(uint8_t op, uint8_t ty, uint64_t ua, uint64_t ub, bool live, uint32_t& err, bool& ok) {
  switch (ty) {
    case DT_I64: {
      int64_t a = (int64_t)ua, b = (int64_t)ub, r = 0;
      switch (op) {
        case OP_ADD:
          r = (int64_t)(ua + ub);
          if (live && (((a ^ r) & (b ^ r)) < 0)) err |= ERR_OVERFLOW;
          break;
        case OP_SUB:
          r = (int64_t)(ua - ub);
          if (live && (((a ^ b) & (a ^ r)) < 0)) err |= ERR_OVERFLOW;
          break;
        case OP_MUL:
          if (mul_overflow_i64(a, b, &r) && live) err |= ERR_OVERFLOW;
          break;
        case OP_DIV:
          if (b == 0) {
            if (live) err |= ERR_DIVZERO;
          } else if (a == INT64_MIN && b == -1) {
            if (live) err |= ERR_OVERFLOW;
          } else {
            r = a / b;
          }
          break;
        case OP_MOD:
          if (b == 0) {
            ok = false;  // NULL, not an error (operators.rs:720)
          } else if (a == INT64_MIN && b == -1) {
            if (live) err |= ERR_OVERFLOW;  // Rust aborts here; surfaced as an error
          } else {
            r = a % b;
          }
          break;
        default: break;
      }
      return (uint64_t)r;
    }
    case DT_I32: {
      int64_t a = (int64_t)ua, b = (int64_t)ub, r = 0;
      switch (op) {
        case OP_ADD: r = a + b; break;
        case OP_SUB: r = a - b; break;
        case OP_MUL: r = a * b; break;
        case OP_DIV:
          if (b == 0) {
            if (live) err |= ERR_DIVZERO;
          } else {
            r = a / b;
          }
          break;
        case OP_MOD:
          if (b == 0) {
            ok = false;
          } else if (a == INT32_MIN && b == -1) {
            if (live) err |= ERR_OVERFLOW;
          } else {
            r = a % b;
          }
          break;
        default: break;
      }
      if (live && (r < INT32_MIN || r > INT32_MAX)) {
        err |= ERR_OVERFLOW;
        r = 0;
      }
      return (uint64_t)r;
    }
    case DT_F64: {
      double a = bits_f64(ua), b = bits_f64(ub), r = 0.0;
      switch (op) {
        case OP_ADD: r = a + b; break;
        case OP_SUB: r = a - b; break;
        case OP_MUL: r = a * b; break;
        case OP_DIV: r = a / b; break;  // IEEE: no error
        default: break;
      }
      return f64_bits(r);
    }
    case DT_F32: {
      float a = bits_f32(ua), b = bits_f32(ub), r = 0.f;
      switch (op) {
        case OP_ADD: r = a + b; break;
        case OP_SUB: r = a - b; break;
        case OP_MUL: r = a * b; break;
        case OP_DIV: r = a / b; break;
        default: break;
      }
      return f32_bits(r);
    }
    default: return 0;
  }
}

QG_HD bool cmp1(uint8_t op, int64_t a, int64_t b) {
  switch (op) {
    case OP_EQ: return a == b;
    case OP_NE: return a != b;
    case OP_LT: return a < b;
    case OP_LE: return a <= b;
    case OP_GT: return a > b;
    default: return a >= b;  // OP_GE
  }
}

// ---------------------------------------------------------------------------------------------
// The accumulator machine.  `exec_insn` executes ONE instruction on the thread's R rows; it is the single
// definition of the evaluator's semantics and is used two ways:
//   * interpreted: run_program loops over a program held in shared memory (instruction fields are run-time values,
//     the switches are uniform branches);
//   * specialised: jit.cu emits `exec_insn<R>(Insn{...literal fields...}, ...)` once per instruction and NVRTC
//     inlines it with constant fields, which folds every switch below away (first Futamura projection).
// Row loops sit INSIDE the opcode / type switches so the interpreted form pays one dispatch per R rows.
// Loader supplies column values, Stack the spill space, Sink consumes OP_AGG / OP_STORE / OP_KEY.
// ---------------------------------------------------------------------------------------------
template <int R>
struct Machine {
  Lanes<R> acc;
  uint32_t pass;  // rows still selected (enters as the in-bounds mask)
  uint32_t gate;  // rows on which checked-arithmetic errors count
  uint32_t err;   // DevErr bits
};

template <int R>
QG_HD void machine_reset(Machine<R>& m, uint32_t inb) {
#pragma unroll
  for (int r = 0; r < R; ++r) m.acc.v[r] = 0;
  m.acc.ok = 0;
  m.pass = inb;
  m.gate = inb;
}

template <int R, class Loader>
QG_HD void load_operand(const Insn& in, Loader& ld, Lanes<R>& x) {
  constexpr uint32_t ALL = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
  if (in.mode == M_COL) {
    ld.template load<R>(in.arg, x);
    const uint8_t sty = insn_sty(in);
    if (sty != in.ty) {
#pragma unroll
      for (int r = 0; r < R; ++r) x.v[r] = convert(x.v[r], sty, in.ty);
    }
  } else {  // M_LIT
#pragma unroll
    for (int r = 0; r < R; ++r) x.v[r] = in.imm;
    x.ok = (in.flags & F_LITNULL) ? 0u : ALL;
  }
}

// comparison of R rows against R rows -> bit mask; the (type, op) dispatch is outside the row loop
template <int R>
QG_HD uint32_t cmp_rows(uint8_t op, uint8_t ty, const uint64_t* a, const uint64_t* b) {
  uint32_t t = 0;
#define QG_CMP_LOOP(EXPR)                                   \
  _Pragma("unroll") for (int r = 0; r < R; ++r) {           \
    const int64_t x = cmp_key(a[r], ty), y = cmp_key(b[r], ty); \
    t |= ((EXPR) ? 1u : 0u) << r;                           \
  }
  switch (op) {
    case OP_EQ: QG_CMP_LOOP(x == y) break;
    case OP_NE: QG_CMP_LOOP(x != y) break;
    case OP_LT: QG_CMP_LOOP(x < y) break;
    case OP_LE: QG_CMP_LOOP(x <= y) break;
    case OP_GT: QG_CMP_LOOP(x > y) break;
    default: QG_CMP_LOOP(x >= y) break;
  }
#undef QG_CMP_LOOP
  return t;
}

// comparison of R rows against one pre-keyed literal
template <int R>
QG_HD uint32_t cmp_rows_lit(uint8_t op, uint8_t sty, uint8_t ty, const uint64_t* a, int64_t kl) {
  uint32_t t = 0;
#define QG_CMP_LOOP(EXPR)                                                   \
  _Pragma("unroll") for (int r = 0; r < R; ++r) {                           \
    const int64_t x = cmp_key(convert(a[r], sty, ty), ty);                  \
    t |= ((EXPR) ? 1u : 0u) << r;                                           \
  }
  switch (op) {
    case OP_EQ: QG_CMP_LOOP(x == kl) break;
    case OP_NE: QG_CMP_LOOP(x != kl) break;
    case OP_LT: QG_CMP_LOOP(x < kl) break;
    case OP_LE: QG_CMP_LOOP(x <= kl) break;
    case OP_GT: QG_CMP_LOOP(x > kl) break;
    default: QG_CMP_LOOP(x >= kl) break;
  }
#undef QG_CMP_LOOP
  return t;
}

template <int R, class Loader, class Stack, class Sink>
QG_HD void exec_insn(const Insn in, Machine<R>& m, Loader& ld, Stack& st, Sink& sk) {
  Lanes<R>& acc = m.acc;
  switch (in.op) {
    case OP_LD: load_operand<R>(in, ld, acc); break;
    case OP_ADD: case OP_SUB: case OP_MUL: case OP_DIV: case OP_MOD:
    case OP_EQ: case OP_NE: case OP_LT: case OP_LE: case OP_GT: case OP_GE:
    case OP_AND: case OP_OR: {
      Lanes<R> x;
      if (in.mode == M_STK) st.template pop<R>(x);
      else load_operand<R>(in, ld, x);
      const bool rev = in.flags & F_REV;
      const uint32_t both = acc.ok & x.ok;
      uint32_t okout = both;
      if (in.op <= OP_MOD) {
        const bool fp = in.ty == DT_F64 || in.ty == DT_F32;
        if (fp) {
          // IEEE arithmetic never raises and never nulls: no per-row bookkeeping
          uint32_t dummy_err = 0;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            bool ok = true;
            acc.v[r] = arith1(in.op, in.ty, rev ? x.v[r] : acc.v[r], rev ? acc.v[r] : x.v[r], false, dummy_err, ok);
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            bool ok = true;
            const bool live = (both >> r) & (m.gate >> r) & 1u;
            const uint64_t lhs = rev ? x.v[r] : acc.v[r], rhs = rev ? acc.v[r] : x.v[r];
            uint32_t e = 0;
            acc.v[r] = arith1(in.op, in.ty, lhs, rhs, live, e, ok);
            m.err |= e;
            if (e) sk.fault(in, r, lhs, rhs, e);  // (a no-op in every sink but the diagnostic one)
            if (!ok) okout &= ~(1u << r);
          }
        }
      } else if (in.op <= OP_GE) {
        const uint32_t t = rev ? cmp_rows<R>(in.op, in.ty, x.v, acc.v) : cmp_rows<R>(in.op, in.ty, acc.v, x.v);
#pragma unroll
        for (int r = 0; r < R; ++r) acc.v[r] = (t >> r) & 1u;
      } else if (in.op == OP_AND) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc.v[r] = acc.v[r] & x.v[r] & 1u;
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) acc.v[r] = (acc.v[r] | x.v[r]) & 1u;
      }
      acc.ok = okout;
      break;
    }
    case OP_NOT:
#pragma unroll
      for (int r = 0; r < R; ++r) acc.v[r] = (~acc.v[r]) & 1u;
      break;
    case OP_NEG:
      if (in.ty == DT_F64) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc.v[r] ^= 0x8000000000000000ull;
      } else if (in.ty == DT_F32) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc.v[r] = (acc.v[r] ^ 0x80000000ull) & 0xffffffffull;
      } else if (in.ty == DT_I32) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc.v[r] = (uint64_t)(int64_t)(int32_t)(0u - (uint32_t)acc.v[r]);  // wraps in i32
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) acc.v[r] = 0ull - acc.v[r];  // wraps in i64 (release-mode Rust `-x`)
      }
      break;
    case OP_CAST:
#pragma unroll
      for (int r = 0; r < R; ++r) acc.v[r] = convert(acc.v[r], in.ty, (uint8_t)in.arg);
      break;
    case OP_PUSH: st.template push<R>(acc); break;
    case OP_PRED: {
      uint32_t t = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) t |= (uint32_t)(acc.v[r] & 1u) << r;
      m.pass &= t & acc.ok;  // null predicate slot == drop (arrow filter)
      break;
    }
    case OP_CMPCL: {
      Lanes<R> x;
      ld.template load<R>(in.arg, x);
      // literal already an order key in the compute type
      m.pass &= cmp_rows_lit<R>(in.mode, insn_sty(in), in.ty, x.v, (int64_t)in.imm) & x.ok;
      break;
    }
    case OP_GATE:
      m.gate = m.pass;
      ld.gate(m.pass);
      break;
    case OP_AGG: sk.template agg<R>(in, acc, m.pass); break;
    case OP_STORE: sk.template store<R>(in, acc, m.pass); break;
    case OP_KEY: sk.template key<R>(in, acc, m.pass); break;
    default: break;
  }
}

// interpreted form: program in (shared) memory
template <int R, class Loader, class Stack, class Sink>
QG_HD void run_program(const Insn* prog, int n_insn, Machine<R>& m, Loader& ld, Stack& st, Sink& sk) {
  for (int pc = 0; pc < n_insn; ++pc) exec_insn<R>(prog[pc], m, ld, st, sk);
}

}  // namespace qgpu
