// pjoin_kernels.cuh — launchers of the radix-partitioned hash join (the fast path of op_join, ops.cu).
//
// Why partition: a 15 M-row build side needs a >= 240 MB table, the 126 MB L2 cannot hold it, and a probe that misses
// L2 pays one random DRAM sector per row (measured on B200: 50 G random 16-byte loads/s from a 512 MB table against
// 283 G/s = one per SM cycle from a 32 MB one; CAS 25 G/s against 122 G/s).  One coarse radix pass (F <= 256
// regions, warp-match ranking, shared-memory staging so every global write is a coalesced run) makes each region's
// table <= ~12 MB: the rows of a region stream through while its table sits in L2.
//
//   k_pj_scatter   both sides -> region-major copies (SoA columns, region r at rows [r*cap, r*cap + count[r]))
//   k_pj_join      ONE cooperative kernel pipelines, region after region, [clear table] [insert build rows] [probe +
//                  emit], so that a region's table is created, filled and probed while it sits in L2.  The value of
//                  a key is the single 8-byte payload word of the build side when it has one, else the build row's
//                  position (gathered from the L2-resident region copy).  Probe + emit: key / payload tiles staged
//                  by the bulk-copy engine, table lookup, ballot/popc compaction, one output-cursor reservation per
//                  1024-row tile, output columns written directly (no index arrays, no gather pass)
//
// Two table layouts, chosen by the host from the build key's range:
//   dense  (max - min + 1 <= 8 x build rows, e.g. any surrogate / order key): regions are key RANGES, the table is
//          direct-addressed, one 8-byte value per key of the range (~0 = absent).  One load per probe row, no key
//          compare, no collision chain, no CAS chain in the build (one atomicExch).
//   hashed (anything else): regions by the top bits of a 64-bit mix, 32-byte two-slot buckets, linear over buckets.
//
// Regions are over-allocated by 1/8 instead of counted first (no histogram pass, no host round trip); an overflow (a
// skewed key), a duplicate build key (an output row count that is not bounded by the probe side) or a key / value
// equal to the empty marker raise a flag and op_join falls back to the generic count/emit path.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

constexpr int PJ_MAX_COLS = 16;
constexpr uint32_t PJ_MAX_PARTS = 256;
// per-region counters (scatter cursors, finished-item counts) live one per 128-byte line: neighbours in one 32-byte
// sector would serialise their atomics on one L2 slice
constexpr uint32_t PJ_CTR_STRIDE = 32;
constexpr uint32_t PJ_TILE = 1024;  // probe tile (rows); the probe side's region capacity is a multiple of it

enum PJFlags : uint32_t { PJ_OVERFLOW = 1, PJ_DUPLICATE = 2, PJ_SENTINEL = 4, PJ_TIMEOUT = 8 };

struct PJCol {
  const void* in;
  void* out;
  int32_t width;  // 4 or 8
  int32_t pad;
};
// how a key is mapped to a region (and, for the dense layout, to its table entry)
struct PJMap {
  int32_t dense;        // 1: region = (key - kmin) >> shift, 0: region = top log_parts bits of mix64(key),
                        // 2: region = top log_parts bits of the GROUP BY table hash of a single 8-byte key
                        // 3: region = the log_parts bits below the top skip_bits of that hash (second radix level)
  uint32_t parts;       // number of regions (hashed: a power of two)
  uint32_t log_parts;   // hashed
  uint32_t shift;       // dense: log2(keys per region)
  int64_t kmin;         // dense
  uint64_t range;       // dense: kmax - kmin + 1
  int32_t drop_outside; // dense: rows whose key is outside [kmin, kmax] cannot match: 1 = drop them (INNER probe side)
  uint32_t skip_bits;   // dense == 3
};

struct PJCols {
  int32_t n;
  int32_t key_col;  // index of the key column in c[] (its values are already in registers)
  PJCol c[PJ_MAX_COLS];
};

// multi-GPU scatter: where every GPU's copy of the regions lives (peer pointers into the exchange arenas)
constexpr int PJ_MAX_WORLD = 8;
struct PJPeers {
  int32_t world;
  uint32_t regions_per_rank;              // region q belongs to rank q / regions_per_rank
  uint32_t* cursor[PJ_MAX_WORLD];         // the owner's per-region row cursors (PJ_CTR_STRIDE words apart)
  void* out[PJ_MAX_WORLD][PJ_MAX_COLS];   // the owner's region-major copy of column ci
};

struct __align__(32) PJBucket {  // one 32-byte sector: two slots
  unsigned long long k[2];  // ~0 = empty
  unsigned long long v[2];
};

struct PJEmit {
  int32_t n_left, n_right;
  int32_t left_key, right_key;  // key column indices on either side
  int32_t inline_right;         // right column whose 8-byte value rides in the slot, or -1
  int32_t left_join;
  int32_t prefetch_left;        // 8-byte probe column loaded together with the keys, or -1
  int32_t pad;
  PJCol left[PJ_MAX_COLS];      // in = region-major probe column, out = output column
  PJCol right[PJ_MAX_COLS];     // in = region-major build column, out = output column
  uint8_t* right_valid_bytes;   // LEFT: 1 byte per output row (1 = matched)
};

// rows -> regions.  `cursor` (parts counters at PJ_CTR_STRIDE words, zeroed by the caller) ends up holding every
// region's row count.
// `peers` != nullptr: the regions are spread over the GPUs (see PJPeers); `cursor` and the `out` pointers of `cols`
// are then unused.
void launch_pj_scatter(Ctx* c, const int64_t* key, int64_t n, const PJMap& map, uint32_t cap, uint32_t* cursor, uint32_t* flags,
                       const PJCols& cols, const PJPeers* peers, const char* name);
// build + probe in one cooperative kernel (see pjoin_kernels.cu).  `cleared` / `inserted` (parts words each),
// *ticket, *out_cursor and *flags must be zero; *out_cursor ends up holding the number of output rows.
// hashed: table = parts x n_entries PJBucket; dense: table = parts x n_entries uint64 (n_entries = keys per region)
struct PJJoinArgs {
  const int64_t* build_key;   // region-major build side
  const uint64_t* build_val;  // 8-byte payload riding in the table, or nullptr (value = build row position)
  const int64_t* probe_key;   // region-major probe side
  const uint32_t* count_r;    // rows per region, build / probe (counters at PJ_CTR_STRIDE words, like cleared / inserted)
  const uint32_t* count_l;
  void* table;
  uint32_t cap_r, cap_l;      // region capacities (rows); cap_l is a multiple of PJ_TILE
  uint32_t n_entries;         // table entries per region: buckets (hashed) or values (dense, even)
  uint32_t pad;
  uint32_t* flags;
  uint32_t* cleared;          // per region: finished clear items
  uint32_t* inserted;         // per region: finished insert items
  uint32_t* ticket;           // next work item
  unsigned long long* out_cursor;
  unsigned long long* timing;  // debug: 24 zeroed words of per-item-kind counters, or nullptr
};
void launch_pj_join(Ctx* c, const PJMap& map, const PJJoinArgs& args, const PJEmit& emit);

}  // namespace qgpu
