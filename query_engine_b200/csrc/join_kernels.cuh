// join_kernels.cuh — launchers of the hash-join kernels.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

// The generic path's table: one flat array of 32-byte BUCKETS in HBM probed linearly (one 32 B sector = one DRAM / L2
// access granule):
//   compact: 4 slots of 8 B  = (uint32 (key - key_min)) << 32 | (build row + 1)   when key_max - key_min < 2^32
//   wide:    2 slots of 16 B = {int64 key, uint32 build row + 1, pad}
// 0 = empty.  A chain ends at the first bucket that still has an empty slot.
struct JoinTable {
  void* slots;
  uint32_t cap;       // slots (power of two, >= 4)
  int32_t compact;
  int64_t key_min;    // compact only
  int64_t key_max;
  size_t bytes() const { return (size_t)cap * (compact ? 8 : 16); }
};

// min / max of the valid keys of a column -> d_minmax[0..1] (int64 register representation)
void launch_key_minmax(Ctx* c, const ColRef& key, int64_t n, int64_t* d_minmax);

void launch_join_build(Ctx* c, const ColRef& key, int64_t n, const JoinTable& t);
int64_t join_tiles(int64_t n);
void launch_join_count(Ctx* c, const ColRef& key, int64_t n, const JoinTable& t, bool left_join, uint32_t* first_match,
                       uint32_t* tile_counts, uint32_t* dup_flag);
// generic emit: (left row, right row) index pairs in probe order; re-probes only when the build side has duplicates
void launch_join_emit(Ctx* c, const ColRef& key, int64_t n, const JoinTable& t, bool left_join, bool has_dups,
                      const uint32_t* first_match, const int64_t* tile_off, uint32_t* out_left, uint32_t* out_right);

// fused emit for a duplicate-free build side: writes the output columns directly (left values from the probe row,
// right values gathered at the matched build row) - no index arrays, no separate gather passes
constexpr int MAX_JOIN_COLS = 16;
struct JoinOutCol {
  const void* in;
  void* out;
  int32_t width;  // 4 or 8
  int32_t pad;
};
struct JoinEmitCols {
  int32_t n_left, n_right;
  JoinOutCol left[MAX_JOIN_COLS];
  JoinOutCol right[MAX_JOIN_COLS];
  uint8_t* right_valid_bytes;  // LEFT join: 1 byte per output row (1 = matched), else nullptr
};
void launch_join_emit_fused(Ctx* c, int64_t n, bool left_join, const uint32_t* first_match, const int64_t* tile_off,
                            const JoinEmitCols& cols);

// composite / Utf8 join keys: a 64-bit surrogate per row (NULL component -> invalid), and the check that every emitted
// pair really has equal keys (*d_mismatches += pairs that do not: a 64-bit collision, the host retries with a new seed)
void launch_hash_rows(Ctx* c, const std::vector<ColRef>& keys, int64_t n, uint64_t seed, int64_t* out, uint32_t* valid);
void launch_verify_pairs(Ctx* c, const std::vector<ColRef>& lkeys, const std::vector<ColRef>& rkeys, const uint32_t* out_left,
                         const uint32_t* out_right, int64_t n, uint32_t* d_mismatches);
// RIGHT / FULL joins: list[0 .. *d_count) = the rows of [0, n_rows) that no entry of matched_idx names (NULL_IDX entries
// are skipped); bits: scratch bitmap of n_rows bits
void launch_unmatched_rows(Ctx* c, const uint32_t* matched_idx, int64_t n_idx, int64_t n_rows, uint32_t* bits, uint32_t* list,
                           uint32_t* d_count);
void launch_fill_u32(Ctx* c, uint32_t* p, int64_t n, uint32_t v);

}  // namespace qgpu
