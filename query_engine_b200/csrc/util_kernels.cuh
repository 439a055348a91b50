// util_kernels.cuh — prefix scans, order-preserving stream compaction, gathers.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

constexpr uint32_t NULL_IDX = 0xffffffffu;  // "no row": LEFT-join misses gather NULLs

// out[i] = sum(in[0..i)), out[n] = total.  in: uint32 counts (n entries), out: int64 (n+1 entries)
void scan_u32_to_i64(Ctx* c, const uint32_t* in, int64_t n, int64_t* out);
// same for int64 input
void scan_i64_to_i64(Ctx* c, const int64_t* in, int64_t n, int64_t* out);

enum CompactKind { CK_W8 = 0, CK_W4 = 1, CK_BITS = 2, CK_ROWIDX = 3 };
struct CompactCol {
  const void* in;
  void* out;
  const uint32_t* in_valid;  // nullptr: column has no nulls
  uint32_t* out_valid;       // zero-initialised bitmap, or nullptr
  int32_t kind;
  int32_t pad;
};
constexpr int MAX_COMPACT_COLS = 24;
// order-preserving compaction of up to MAX_COMPACT_COLS columns in one pass over the selection bitmap
// n_selected (when the caller knows it): sparse selections take a warp-per-tile kernel whose work follows the survivors
void launch_compact(Ctx* c, int64_t n_rows, int tile_rows, const uint32_t* mask, const int64_t* tile_off,
                    const CompactCol* cols, int n_cols, int64_t n_selected = -1);

// counts[t] = selected rows of the 1024-row tile t of a selection bitmap (whole tiles: the bitmap is zero-padded)
void launch_mask_tile_counts(Ctx* c, const uint32_t* mask, int64_t n_tiles, uint32_t* counts);

// out[i] = number of selected rows with index < boundaries[i]
void launch_prefix_at(Ctx* c, const int64_t* d_boundaries, int nb, int tile_rows, const uint32_t* mask,
                      const int64_t* tile_off, int64_t* d_out);

struct GatherCol {
  const void* in;
  void* out;
  const uint32_t* in_valid;
  uint32_t* out_valid;  // written with whole words (ballot); nullptr = not needed
  int32_t kind;         // CK_W8 / CK_W4 / CK_BITS
  int32_t pad;
};
constexpr int MAX_GATHER_COLS = 16;
void launch_gather(Ctx* c, const uint32_t* idx, int64_t n_out, const GatherCol* cols, int n_cols);

// Utf8 gather: lengths -> (scan) -> offsets -> bytes
void launch_gather_utf8_len(Ctx* c, const int32_t* in_off, const uint32_t* idx, int64_t n_out, int64_t* out_len);
void launch_offsets_i64_to_i32(Ctx* c, const int64_t* in, int64_t n, int32_t* out);
void launch_gather_utf8_copy(Ctx* c, const int32_t* in_off, const uint8_t* in_chars, const uint32_t* idx, int64_t n_out,
                             const int32_t* out_off, uint8_t* out_chars);

// multi-GPU scalar aggregate: all [world][n_acc + 1] -> finals [n_acc + 1] (accumulators folded in rank order by
// kind; last word = OR of the ranks' device error words)
void launch_fold_finals(Ctx* c, const uint64_t* all, int world, int n_acc, const uint8_t* kinds_host, uint64_t* finals);
// out[i] = in[i * stride], i < count (partition starts out of a partition-major histogram scan)
void launch_gather_strided_i64(Ctx* c, const int64_t* in, int64_t stride, int count, int64_t* out);
void launch_iota_u32(Ctx* c, uint32_t* out, int64_t n, uint32_t first = 0);  // out[i] = first + i
// validity bitmaps <-> one byte per row (NCCL moves whole bytes); bits == nullptr means all valid
void launch_bits_to_bytes(Ctx* c, const uint32_t* bits, int64_t n, uint8_t* bytes);
void launch_bytes_to_bits(Ctx* c, const uint8_t* bytes, int64_t n, uint32_t* bits);

}  // namespace qgpu
