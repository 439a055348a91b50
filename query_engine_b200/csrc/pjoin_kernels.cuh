// pjoin_kernels.cuh — launchers of the radix-partitioned hash join (the fast path of op_join, ops.cu).
//
// Why partition: a 15 M-row build side needs a >= 240 MB table, the 126 MB L2 cannot hold it, and a probe that misses
// L2 pays one random DRAM sector per row (measured: 60 M probes ~ 1.2 ms, 60 M payload gathers another ~ 1 ms).  One
// coarse radix pass (F <= 256 regions, warp-match ranking, shared-memory staging so every global write is a coalesced
// run) makes each region's table <= ~12 MB: the rows of a region stream through while its table sits in L2.
//
//   k_pj_scatter   both sides -> region-major copies (SoA columns, region r at rows [r*cap, r*cap + count[r]))
//   k_pj_build     region tables of 16-byte slots {key, value}; value = the single 8-byte payload word of the build
//                  side when it has one, else the build row's position (gathered from the L2-resident region copy)
//   k_pj_probe     fused probe + emit: one table sector per probe row, ballot/popc compaction, one output-cursor
//                  reservation per 1024-row tile, output columns written directly (no index arrays, no gather pass)
//
// Regions are over-allocated by 1/8 instead of counted first (no histogram pass, no host round trip); an overflow (a
// skewed key), a duplicate build key (an output row count that is not bounded by the probe side) or a key equal to
// the empty marker raise a flag and op_join falls back to the generic count/emit path.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

constexpr int PJ_MAX_COLS = 16;
constexpr uint32_t PJ_MAX_PARTS = 256;
constexpr uint32_t PJ_TILE = 1024;  // probe tile (rows); the probe side's region capacity is a multiple of it

enum PJFlags : uint32_t { PJ_OVERFLOW = 1, PJ_DUPLICATE = 2, PJ_SENTINEL = 4 };

struct PJCol {
  const void* in;
  void* out;
  int32_t width;  // 4 or 8
  int32_t pad;
};
struct PJCols {
  int32_t n;
  int32_t key_col;  // index of the key column in c[] (its values are already in registers)
  PJCol c[PJ_MAX_COLS];
};

struct PJSlot {
  unsigned long long key;  // ~0 = empty
  unsigned long long val;
};

struct PJEmit {
  int32_t n_left, n_right;
  int32_t left_key, right_key;  // key column indices on either side
  int32_t inline_right;         // right column whose 8-byte value rides in the slot, or -1
  int32_t left_join;
  PJCol left[PJ_MAX_COLS];      // in = region-major probe column, out = output column
  PJCol right[PJ_MAX_COLS];     // in = region-major build column, out = output column
  uint8_t* right_valid_bytes;   // LEFT: 1 byte per output row (1 = matched)
};

// rows -> regions.  `cursor` (parts x uint32, zeroed by the caller) ends up holding every region's row count.
void launch_pj_scatter(Ctx* c, const int64_t* key, int64_t n, uint32_t parts, uint32_t cap, uint32_t* cursor, uint32_t* flags,
                       const PJCols& cols, const char* name);
// table must be filled with 0xff bytes (empty slots) by the caller
void launch_pj_build(Ctx* c, const int64_t* key_part, const uint64_t* val_part, uint32_t parts, uint32_t cap, const uint32_t* count,
                     PJSlot* table, uint32_t tcap, uint32_t* flags);
void launch_pj_probe(Ctx* c, const int64_t* key_part, uint32_t parts, uint32_t cap, const uint32_t* count, const PJSlot* table,
                     uint32_t tcap, unsigned long long* out_cursor, const PJEmit& e);

}  // namespace qgpu
