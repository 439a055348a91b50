// scan_params.cuh — parameter blocks of the fused column-scan kernels (device-safe: also compiled by NVRTC).
#pragma once
#include "common.cuh"

namespace qgpu {

constexpr int SCAN_THREADS = 256;

// open-addressing table of group representatives in HBM (lock-free, no sentinels, no spinning):
// a slot holds (row index + 1) of the first row that claimed it; key equality is checked against the
// key columns at that row, so any number / type of key columns (Utf8 included) works.
//
// Slot of a key: the TOP GROUP_REGION_BITS bits of its hash pick one of 128 sub-tables, the middle bits the slot inside
// it (tables below 2^14 slots are not divided).  Which sub-table a key lives in therefore does not depend on the
// table's size, and rows that were radix-scattered on those same hash bits (op_aggregate's high-cardinality path)
// touch one sub-table after the other: each stays L2-resident while its rows stream through.
constexpr int GROUP_REGION_BITS = 7;
struct GroupTable {
  uint32_t* rep;     // [cap]  0 = empty
  uint64_t* acc;     // [cap][n_acc]
  uint32_t cap;      // power of two
  int32_t n_acc;
  int32_t n_keys;
  // the table is 2^log_regions sub-tables for the hash regions [region_base, region_base + 2^log_regions) of the 128:
  // 7 / 0 on one GPU; a rank of the key-partitioned multi-GPU aggregate only ever sees its own 128 / W regions
  int32_t log_regions;
  uint32_t region_base;
  int32_t pad;
  ColRef keys[MAX_KEYS];
};
QG_HD uint32_t group_home_slot(uint64_t h, const GroupTable& g) {
  const uint32_t mask = g.cap - 1;
  if (g.cap < (1u << 14) || g.log_regions == 0) return (uint32_t)(h >> 20) & mask;
  const uint32_t sub = g.cap >> g.log_regions;  // slots per sub-table
  const uint32_t r = ((uint32_t)(h >> (64 - GROUP_REGION_BITS)) - g.region_base) & ((1u << g.log_regions) - 1u);
  return r * sub + ((uint32_t)(h >> 20) & (sub - 1));
}

enum SinkKind { SK_MASK = 0, SK_SCALAR = 1, SK_GROUP_SMEM = 2, SK_GROUP_GLOBAL = 3, SK_STORE = 4, SK_GROUP_PART = 5, SK_DIAG = 6 };

// what the diagnostic pass records about the FIRST failing checked operation (evaluation order of the reference)
struct DiagRecord {
  uint64_t lhs, rhs;
  uint32_t op, ty, err, pad;
};

struct ScanParams {
  int64_t n_rows;
  int64_t n_tiles;
  // region-major input with gaps (the radix scatter's output): region r holds rows [r * region_cap,
  // r * region_cap + region_count[r * region_count_stride]); region_cap (a multiple of the tile) = 0: plain table
  uint32_t region_cap;
  uint32_t region_count_stride;
  const uint32_t* region_count;
  const Insn* prog;
  int32_t n_insn;
  int32_t n_pred;        // instructions [0, n_pred) are the predicate section (ends with OP_GATE)
  int32_t n_cols;
  int32_t stack_depth;
  int32_t l2_prefetch;
  int32_t col_width[MAX_COLS];
  ColRef cols[MAX_COLS];
  uint32_t* err;         // [0] DevErr bits, [2] rows that passed the predicate (specialised aggregate scans)
  // SK_MASK
  uint32_t* mask;
  uint32_t* tile_counts;
  // SK_SCALAR / grouped
  int32_t n_acc;
  uint8_t acc_kinds[MAX_ACC];
  uint64_t* partials;
  uint64_t* finals;
  uint32_t* ticket;
  // grouped
  int32_t n_keys;
  int32_t key_slots[MAX_KEYS];
  int32_t gp;            // SK_GROUP_SMEM: per-CTA directory capacity
  uint32_t* overflow;    // SK_GROUP_SMEM: set when a CTA saw more than gp groups
  GroupTable gt;
  // SK_GROUP_PART: the input is partition-major (radix-partitioned on the key hash, rscatter_kernels.cu): `part_regions`
  // consecutive input regions hold all rows of one set of keys.  A CTA aggregates a partition in a shared-memory table
  // of `part_slots` slots and appends its groups - key, raw accumulator row - to the output arrays.
  int32_t part_regions;
  uint32_t part_slots;       // power of two
  uint32_t gp_out_cap;       // capacity of the output arrays (groups)
  uint32_t gp_preload;       // request every referenced column of a tile up front (specialised kernels: QGPU_GP_PRELOAD)
  uint64_t* gp_out_keys;     // [gp_out_cap]
  uint64_t* gp_out_acc;      // [gp_out_cap][n_acc]
  uint32_t* gp_cursor;       // groups written so far
  // SK_DIAG: after a checked-arithmetic error, find the failing operation that the reference would have reported - the
  // first in (section, batch, instruction, row) order, sections being the predicate and then every aggregate /
  // projected expression - and record its operands (diag_mode 0: atomicMin of the order key; 1: the holder of the
  // minimal key writes the record)
  int32_t diag_mode;
  int32_t diag_n_batches;
  const int64_t* diag_batch_offsets;   // [diag_n_batches + 1]
  const uint8_t* diag_section_of_pc;   // [n_insn]
  unsigned long long* diag_key;
  DiagRecord* diag_rec;
  // SK_STORE
  int32_t n_out;
  uint8_t out_types[MAX_OUT];
  void* out_data[MAX_OUT];
  uint32_t* out_valid[MAX_OUT];
};

}  // namespace qgpu
