// rscatter_kernels.cuh — the radix scatter of 8-byte columns (second generation of k_pj_scatter, pjoin_kernels.cu).
//
// One pass moves every row of up to RS_MAX_COLS 8-byte columns into the region its key maps to (PJMap).  All columns of
// a 2048-row tile arrive in shared memory through the bulk-copy engine (two stages, mbarrier completion) while the
// previous tile is processed; rows are ranked inside the tile with MATCH.ANY + warp-private histograms, the tile's
// region-major order is recorded as one packed (region, source row) word per position, one atomicAdd per (tile, region)
// reserves the run, and every column leaves shared memory as coalesced runs.  Everything per-row is compile-time
// unrolled over the column count (no per-column dispatch in the inner loops): about half the instructions per row of
// the first version, which was SM-issue-bound (profiles/r01h_ncu_full.csv: sm 61 %, DRAM 2.9 TB/s).
//
// Placement of a region's rows:
//   over-allocated (start == nullptr): region q owns rows [q * cap, (q + 1) * cap); a run that does not fit raises
//       PJ_OVERFLOW and lands in the trash rows at the end of every output column (the caller retries another way);
//   exact (start != nullptr): region q owns rows [start[q], start[q + 1]) - counted first by launch_region_hist - so the
//       scattered columns have no gaps: the hash join writes its probe-side OUTPUT columns this way.
// Input may itself be region-major with gaps (in_cap != 0): the second radix level of the high-cardinality GROUP BY
// then maps the rows of input region g to output regions g * M.parts + sub.
#pragma once
#include "pjoin_kernels.cuh"

namespace qgpu {

constexpr int RS_MAX_COLS = 4;
constexpr int RS_TILE = 2048;

struct RSArgs {
  const uint64_t* in[RS_MAX_COLS];   // [0] = the key column
  uint64_t* out[RS_MAX_COLS];
  int64_t n;                         // physical input rows (gaps included for region-major input)
  PJMap M;
  uint32_t cap;                      // rows per output region (over-allocated placement)
  uint32_t ctr_stride;               // words between the per-region cursors
  const uint32_t* start;             // exact placement: first output row of every region, or nullptr
  uint32_t* cursor;                  // rows placed so far per output region (zeroed by the caller)
  uint32_t* flags;
  uint32_t in_cap;                   // region-major input: rows per input region (a multiple of RS_TILE); 0 = plain
  uint32_t in_ctr_stride;
  const uint32_t* in_count;          // region-major input: rows per input region
  uint32_t trash_row;                // first of RS_TILE rows at the end of every output column that absorb overflowing runs
  uint32_t plain_stores;             // 1: ordinary stores (L2 may combine the tails of consecutive runs), 0: evict-first
  uint32_t first_base;               // added to the first row of every output region (a sender's segment of a shared region)
  uint32_t merge_in;                 // region-major input: 1 = all input regions feed the SAME output regions [0, parts)
};

struct RSPeers {
  int32_t world;
  uint32_t regions_per_rank;               // output region q lives on rank q / regions_per_rank, as its region q % regions_per_rank
  uint32_t* cursor[PJ_MAX_WORLD];          // the owner's cursors
  uint64_t* out[PJ_MAX_WORLD][RS_MAX_COLS];  // the owner's copy of column ci
};

// n_cols in 1..RS_MAX_COLS; peers == nullptr: single GPU
void launch_rscatter(Ctx* c, const RSArgs& args, int n_cols, const RSPeers* peers, const char* name);

// counts[q * ctr_stride] += rows of `key` that map to region q (rows that map to no region are not counted)
void launch_region_hist(Ctx* c, const int64_t* key, int64_t n, const PJMap& map, uint32_t* counts, uint32_t ctr_stride);
// start[q] = sum of counts[< q] for q in 0..parts (start[parts] = total); raises PJ_OVERFLOW in *flags when a count
// exceeds max_count
void launch_region_offsets(Ctx* c, const uint32_t* counts, uint32_t ctr_stride, uint32_t parts, uint32_t max_count, uint32_t* start,
                           uint32_t* flags);

}  // namespace qgpu
