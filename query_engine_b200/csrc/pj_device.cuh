// pj_device.cuh — device helpers shared by the radix-partitioning / join kernels (pjoin_kernels.cu, rscatter_kernels.cu,
// pjoin2_kernels.cu): key -> region maps, L2 eviction-policy loads / stores, bulk-copy (TMA, 1-D) staging with mbarriers.
#pragma once
#include "device_utils.cuh"
#include "pjoin_kernels.cuh"

namespace qgpu {

constexpr unsigned long long PJ_EMPTY = ~0ull;

// one 64-bit mix per key: the TOP bits choose the region, the LOW 32 bits the bucket inside the region table
QG_D uint32_t pj_region(uint64_t h, uint32_t log_parts) { return log_parts ? (uint32_t)(h >> (64 - log_parts)) : 0u; }
// region of a key; M.parts = "no region" (the row is dropped)
QG_D uint32_t pj_region_of(const PJMap& M, uint64_t key) {
  if (!M.dense) return pj_region(mix64(key), M.log_parts);
  // 2: the sub-table of the GROUP BY table a single 8-byte key lives in (scan_body.cuh hash_row + group_home_slot)
  if (M.dense == 2) return (uint32_t)(mix64(0x9e3779b97f4a7c15ull ^ key) >> (64 - M.log_parts));
  // 3: the next log_parts bits of the same hash below its top skip_bits (second radix level of the GROUP BY partitioning)
  if (M.dense == 3) return (uint32_t)(mix64(0x9e3779b97f4a7c15ull ^ key) >> (64 - M.skip_bits - M.log_parts)) & (M.parts - 1u);
  const uint64_t idx = key - (uint64_t)M.kmin;
  if (idx < M.range) return (uint32_t)(idx >> M.shift);
  return M.drop_outside ? M.parts : M.parts - 1;
}

// the same with the map kind (M.dense) as a compile-time constant: no per-row branch chain in unrolled loops
template <int KIND>
QG_D uint32_t pj_region_of_k(const PJMap& M, uint64_t key) {
  if (KIND == 0) return pj_region(mix64(key), M.log_parts);
  if (KIND == 2) return (uint32_t)(mix64(0x9e3779b97f4a7c15ull ^ key) >> (64 - M.log_parts));
  if (KIND == 3) return (uint32_t)(mix64(0x9e3779b97f4a7c15ull ^ key) >> (64 - M.skip_bits - M.log_parts)) & (M.parts - 1u);
  const uint64_t idx = key - (uint64_t)M.kmin;
  if (idx < M.range) return (uint32_t)(idx >> M.shift);
  return M.drop_outside ? M.parts : M.parts - 1;
}

// ---- L2 eviction hints: what streams through once (staged probe tiles, build rows, output rows) is marked
// evict-first so that it does not push the regions' tables out of L2 between their clear, fill and probe ----------
QG_D uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
QG_D uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
QG_D void st_hint_v2u64(void* p, unsigned long long a, unsigned long long b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.b64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(a), "l"(b), "l"(pol) : "memory");
}
QG_D void st_evict_first_u64(uint64_t* p, uint64_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
QG_D uint64_t ld_hint_u64(const uint64_t* p, uint64_t pol) {
  uint64_t v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}

// ---- bulk-async staging (TMA engine, 1-D): global -> shared, completion on an mbarrier -------------------------
QG_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
QG_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
QG_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
QG_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
QG_D void bulk_g2s(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

QG_D uint32_t pj_bucket(uint64_t h, uint32_t n_buckets) { return __umulhi((uint32_t)h, n_buckets); }

QG_D uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// table reads: the table is written earlier in the SAME kernel by other CTAs, so not the read-only (.nc) path
QG_D void ld_bucket(const PJBucket* p, unsigned long long& k0, unsigned long long& k1, unsigned long long& v0, unsigned long long& v1,
                    uint64_t pol) {
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(k0), "=l"(k1) : "l"(p), "l"(pol) : "memory");
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;"
               : "=l"(v0), "=l"(v1)
               : "l"((const char*)p + 16), "l"(pol)
               : "memory");
}
QG_D unsigned long long ld_entry(const unsigned long long* p, uint64_t pol) {
  unsigned long long v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}

QG_D unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

}  // namespace qgpu
