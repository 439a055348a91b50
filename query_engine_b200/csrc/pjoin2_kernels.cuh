// pjoin2_kernels.cuh — build + probe of the radix-partitioned hash join, second generation (see pjoin2_kernels.cu).
#pragma once
#include "pjoin_kernels.cuh"

namespace qgpu {

constexpr uint32_t PJ2_TILE = 2048;  // probe window (rows)

struct PJ2Args {
  // build side: region-major copy, over-allocated (region r at rows [r * cap_r, r * cap_r + count_r[r]))
  const int64_t* build_key;
  const uint64_t* build_val;   // 8-byte payload riding in the table, or nullptr (value = build row position)
  const uint32_t* count_r;     // counters PJ_CTR_STRIDE words apart
  uint32_t cap_r;
  uint32_t n_entries;          // table entries per region: buckets (hashed) or values (dense, even)
  // probe side: the OUTPUT key column, region p at rows [start_l[p], start_l[p + 1])
  const uint64_t* probe_key;
  const uint32_t* start_l;     // [parts + 1]
  uint64_t probe_rows_alloc;   // rows the probe key column holds (bulk copies stop there)
  uint32_t tiles_per_region;   // probe item slots per region (pj2_tiles_per_region)
  uint32_t ring;               // region tables in the ring
  void* table;                 // ring x n_entries x (8 dense | sizeof(PJBucket) hashed)
  // control words (zeroed by the caller)
  uint32_t* flags;
  uint32_t* cleared;           // per region: finished clear items
  uint32_t* inserted;          // per region: finished insert items
  uint32_t* probed;            // per region: finished probe windows
  uint32_t* ticket;
  unsigned long long* found;   // matched probe rows
  // build-side output columns, written at the probe row's position
  uint64_t* out_right_key;     // or nullptr
  uint64_t* out_inline;        // the payload that rides in the table, or nullptr
  int32_t n_gather;
  int32_t pad;
  PJCol gather[PJ_MAX_COLS];   // other build columns: in = region-major copy, out = output column (8-byte)
  uint8_t* vbytes;             // 1 byte per output row: matched
};

uint32_t pj2_tiles_per_region(int64_t probe_rows, uint32_t parts);
void launch_pj_join2(Ctx* c, const PJMap& map, const PJ2Args& args);

}  // namespace qgpu
