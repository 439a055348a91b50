// common.cuh — shared types of libqgpu (device program encoding, column refs, error plumbing).
//
// The device "program" is the flattened form of the reference's PhysicalExpr trees
// (crates/query-executor/src/physical_plan.rs:74-117): a one-address accumulator
// machine whose instructions are typed at compile time by the host (expr_compile.cu),
// so the kernels never look at Arrow types or do dynamic dispatch per row.
//
// This header (with interp.cuh, device_utils.cuh, scan_params.cuh, scan_body.cuh) is also compiled at run time by
// NVRTC when a scan kernel is specialised on a query (jit.cu), so everything outside the __CUDACC_RTC__ guards must
// be plain device-compatible C++ without standard-library includes.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#define INT64_MAX 0x7fffffffffffffffll
#define INT64_MIN (-INT64_MAX - 1)
#define INT32_MAX 0x7fffffff
#define INT32_MIN (-INT32_MAX - 1)
#else
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#endif

#if defined(__CUDACC__)
#define QG_HD __host__ __device__ __forceinline__
#define QG_D __device__ __forceinline__
#else
#define QG_HD inline
#define QG_D inline
#endif

namespace qgpu {

// ---- value types (numbering = qgpu_dtype in include/qgpu.h) -------------------------------
enum DT : uint8_t { DT_NULL = 0, DT_BOOL = 1, DT_I32 = 2, DT_I64 = 3, DT_F32 = 4, DT_F64 = 5, DT_UTF8 = 6, DT_OTHER = 7 };

// In registers every value is one 64-bit word:
//   BOOL 0/1 · I32 sign-extended to 64 bit · I64 · F32 bits in the low word · F64 bits.

// ---- opcodes -------------------------------------------------------------------------------
enum Op : uint8_t {
  OP_LD = 0,   // acc = operand
  OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MOD,            // acc = acc (op) operand     [REV: operand (op) acc]
  OP_EQ, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE,          // acc(bool) = acc cmp operand  (floats: IEEE totalOrder, as arrow-rs)
  OP_AND, OP_OR,                                     // non-Kleene: null if either side null (operators.rs:539-570)
  OP_NOT, OP_NEG,                                    // unary on acc
  OP_CAST,                                           // acc: ty -> arg
  OP_PUSH,                                           // spill acc to the per-thread stack
  OP_PRED,                                           // pass &= acc && valid(acc)      (end of one conjunct)
  OP_CMPCL,                                          // fused: pass &= (column[arg] <mode> literal imm), no stack traffic
  OP_GATE,                                           // end of the predicate section: error gating switches to `pass`
  OP_AGG,                                            // accumulate acc into accumulator slot arg (kind in mode)
  OP_STORE,                                          // write acc to output column arg (values + validity)
  OP_KEY,                                            // acc -> group-key word arg
  OP_NOP
};
enum Mode : uint8_t { M_NONE = 0, M_COL = 1, M_LIT = 2, M_STK = 3 };
enum AccKind : uint8_t { ACC_ADD_I64 = 0, ACC_ADD_F64 = 1, ACC_MIN_I64 = 2, ACC_MAX_I64 = 3, ACC_COUNT = 4 };

enum InsnFlags : uint8_t {
  F_REV = 1,       // operand is the LEFT side of the binary op
  F_LITNULL = 2,   // literal operand is NULL
  // bits 4..6: storage dtype of a column operand (converted to `ty` on load)
};

struct __align__(16) Insn {
  uint8_t op;
  uint8_t ty;     // compute type (DT)
  uint8_t mode;   // Mode; OP_CMPCL: the comparison Op; OP_AGG: AccKind
  uint8_t flags;
  int32_t arg;
  uint64_t imm;
};
static_assert(sizeof(Insn) == 16, "Insn must be 16 bytes");

QG_HD uint8_t insn_sty(const Insn& in) { return (in.flags >> 4) & 7; }

// device error bits raised by checked arithmetic (arrow-rs numeric::{add,sub,mul,div} are checked)
enum DevErr : uint32_t { ERR_OVERFLOW = 1, ERR_DIVZERO = 2, ERR_SPIN = 4, ERR_CAPACITY = 8, ERR_OUTCAP = 16 };

// ---- a column as the kernels see it ----------------------------------------------------------
struct ColRef {
  const void* data;        // values (fixed width) or bit-packed booleans; Utf8: int32 offsets
  const uint32_t* valid;   // Arrow validity bitmap (LSB first) or nullptr = no nulls
  const uint8_t* chars;    // Utf8 only
  int32_t dtype;           // DT
  int32_t pad;
};

constexpr int MAX_COLS = 16;     // distinct columns one fused program may reference
constexpr int MAX_ACC = 32;      // accumulator slots per aggregate program
constexpr int MAX_KEYS = 4;      // group-by / join key columns
constexpr int MAX_STACK = 4;     // spill depth of the accumulator machine
constexpr int MAX_OUT = 16;      // OP_STORE outputs of one program

#ifndef __CUDACC_RTC__
// ---- host-side error carrying the reference's QueryError variant ---------------------------
struct Error {
  int code;  // qgpu_status
  std::string msg;
};
#endif

}  // namespace qgpu
