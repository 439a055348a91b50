// scan_kernels.cuh — parameter blocks and launchers of the fused column-scan kernels.
#pragma once
#include "common.cuh"
#include "qgpu_internal.h"

namespace qgpu {

constexpr int SCAN_THREADS = 256;

// open-addressing table of group representatives in HBM (lock-free, no sentinels, no spinning):
// a slot holds (row index + 1) of the first row that claimed it; key equality is checked against the
// key columns at that row, so any number / type of key columns (Utf8 included) works.
struct GroupTable {
  uint32_t* rep;     // [cap]  0 = empty
  uint64_t* acc;     // [cap][n_acc]
  uint32_t cap;      // power of two
  int32_t n_acc;
  int32_t n_keys;
  int32_t pad;
  ColRef keys[MAX_KEYS];
};

enum SinkKind { SK_MASK = 0, SK_SCALAR = 1, SK_GROUP_SMEM = 2, SK_GROUP_GLOBAL = 3, SK_STORE = 4 };

struct ScanParams {
  int64_t n_rows;
  int64_t n_tiles;
  const Insn* prog;
  int32_t n_insn;
  int32_t n_pred;        // instructions [0, n_pred) are the predicate section (ends with OP_GATE)
  int32_t n_cols;
  int32_t stack_depth;
  int32_t l2_prefetch;
  int32_t col_width[MAX_COLS];
  ColRef cols[MAX_COLS];
  uint32_t* err;
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
  // SK_STORE
  int32_t n_out;
  uint8_t out_types[MAX_OUT];
  void* out_data[MAX_OUT];
  uint32_t* out_valid[MAX_OUT];
};

// rows_per_thread in {4, 8}; grid is chosen by the launcher (persistent: sm_count x ctas_per_sm)
void launch_scan(Ctx* c, int sink, int rows_per_thread, ScanParams& p, const char* name);
size_t scan_smem_bytes(int sink, int rows_per_thread, const ScanParams& p);

// grouped aggregation helpers
void launch_group_init(Ctx* c, const GroupTable& gt, const uint8_t* acc_kinds_host);
// collects used slots: out_rep[i] = representative row, out_slot[i] = table slot; returns count via d_count
void launch_group_collect(Ctx* c, const GroupTable& gt, uint32_t* out_rep, uint32_t* out_slot, uint32_t* d_count);

struct AggFinalizeDesc {
  int32_t func, in_type, out_type, val_slot, cnt_slot;
  void* out_data;
  uint32_t* out_valid;
};
// out columns from accumulator rows: acc_base[slot_of_group(i)*n_acc + s]; slots == nullptr -> group i uses row i
void launch_agg_finalize(Ctx* c, const uint64_t* acc, const uint32_t* slots, int n_acc, int64_t n_groups,
                         const AggFinalizeDesc* descs_host, int n_desc);

}  // namespace qgpu
