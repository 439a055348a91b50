// device_utils.cuh — small sm_100a device helpers: streaming loads, warp/block reductions, hashing.
#pragma once
#include "common.cuh"
#include "interp.cuh"

namespace qgpu {

constexpr int WARP = 32;

// streaming (read-once) loads: read-only path, do not allocate in L1.  Not `volatile`: two loads of the same
// address in one specialised program are one load after CSE.
QG_D uint64_t ld_stream_u64(const uint64_t* p) {
  uint64_t v;
  asm("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
QG_D uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
QG_D void ld_stream_v2u64(const uint64_t* p, uint64_t& a, uint64_t& b) {
  asm("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
// 16-byte read-only load that MAY allocate in L1: for threads that read 32 contiguous bytes with two instructions
// (the second one then hits the sector the first brought in instead of going back to L2)
QG_D void ld_nc_v2u64(const uint64_t* p, uint64_t& a, uint64_t& b) {
  asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
QG_D void st_stream_u64(uint64_t* p, uint64_t v) { asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(v)); }
QG_D void st_stream_u32(uint32_t* p, uint32_t v) { asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(v)); }

// bulk L2 prefetch of the next tile of a column (sm_90+; one thread issues it for the CTA)
QG_D void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

QG_D uint32_t lane_id() { return threadIdx.x & 31; }
QG_D uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

QG_D uint64_t shfl_xor_u64(uint64_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
QG_D uint64_t shfl_u64(uint64_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// combine two accumulator words of the given AccKind
QG_HD uint64_t acc_combine(uint8_t kind, uint64_t a, uint64_t b) {
  switch (kind) {
    case ACC_ADD_F64: return f64_bits(bits_f64(a) + bits_f64(b));
    case ACC_MIN_I64: return (uint64_t)((int64_t)a < (int64_t)b ? (int64_t)a : (int64_t)b);
    case ACC_MAX_I64: return (uint64_t)((int64_t)a > (int64_t)b ? (int64_t)a : (int64_t)b);
    default: return a + b;  // ACC_ADD_I64 (wrapping, like arrow sum), ACC_COUNT
  }
}
QG_HD uint64_t acc_identity(uint8_t kind) {
  switch (kind) {
    case ACC_MIN_I64: return (uint64_t)INT64_MAX;
    case ACC_MAX_I64: return (uint64_t)INT64_MIN;
    default: return 0;  // +0.0 / 0
  }
}

QG_D uint64_t warp_reduce_acc(uint8_t kind, uint64_t v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = acc_combine(kind, v, shfl_xor_u64(v, m));
  return v;
}

QG_D uint32_t warp_excl_scan_u32(uint32_t v, uint32_t& total) {
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane_id() >= (uint32_t)d) x += y;
  }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

// 64-bit finalizer (splitmix64 / murmur3 fmix64 family): used for group-by / join tables and the
// NCCL shuffle.  (The reference-visible Exchange uses SipHash-1-3 instead, see exchange_kernels.cu.)
QG_HD uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// partition hash of ONE key value (register representation; NULL -> the marker): what the shuffle / radix-join
// partitioning passes and the partitioned join kernels must agree on
constexpr uint64_t PART_NULL_MARK = 0x6e756c6cull;
QG_HD uint64_t part_hash1(uint64_t v) { return mix64(0x9e3779b97f4a7c15ull ^ v); }

// value of row `row` of a fixed-width / bool column in register representation
QG_D uint64_t load_value(const ColRef& c, int64_t row) {
  switch (c.dtype) {
    case DT_I64: case DT_F64: return ((const uint64_t*)c.data)[row];
    case DT_I32: return (uint64_t)(int64_t)((const int32_t*)c.data)[row];
    case DT_F32: return (uint64_t)((const uint32_t*)c.data)[row];
    case DT_BOOL: return (((const uint32_t*)c.data)[row >> 5] >> (row & 31)) & 1u;
    default: return 0;
  }
}
QG_D bool row_valid(const ColRef& c, int64_t row) {
  if (c.dtype == DT_NULL) return false;
  return c.valid == nullptr || ((c.valid[row >> 5] >> (row & 31)) & 1u);
}

}  // namespace qgpu
