// scan_kernels.cuh — parameter blocks and launchers of the fused column-scan kernels.
#pragma once
#include "common.cuh"
#include "qgpu_internal.h"
#include "scan_params.cuh"

namespace qgpu {

// rows_per_thread in {4, 8}; grid is chosen by the launcher (persistent: sm_count x ctas_per_sm)
// `prog` != nullptr lets the launcher specialise the kernel on the program with NVRTC (jit.cu) when the context's
// "scan.jit" policy says so; otherwise the ahead-of-time interpreter kernel runs the program from shared memory.
void launch_scan(Ctx* c, int sink, int rows_per_thread, ScanParams& p, const char* name, const Program* prog = nullptr);
size_t scan_smem_bytes(int sink, int rows_per_thread, const ScanParams& p, bool specialised = false);
// after a specialised aggregate scan: how many of the rows passed its predicate (ScanParams.err[2]) - remembered per
// program and used to pick the kernel's form the next time it runs (launch_scan)
void note_scan_feedback(Ctx* c, uint64_t kept, uint64_t rows);

// grouped aggregation helpers
void launch_group_init(Ctx* c, const GroupTable& gt, const uint8_t* acc_kinds_host);
// collects used slots: out_rep[i] = representative row, out_slot[i] = table slot; returns count via d_count
void launch_group_collect(Ctx* c, const GroupTable& gt, uint32_t* out_rep, uint32_t* out_slot, uint32_t* d_count);

// compact the accumulator rows of the collected groups: out[i][*] = acc[slots[i]][*]
void launch_gather_acc_rows(Ctx* c, const uint64_t* acc, const uint32_t* slots, int n_acc, int64_t ng, uint64_t* out);
// multi-GPU final merge: row i of the gathered key table joins group find_or_insert(g, i); its raw accumulators
// rows[i][*] are folded into the group's with the accumulator kind's atomic
// blk_hdr != nullptr: the rows come from fixed-size per-rank blocks and row i is live when i % g_max is below the group
// count in its rank's block header (blk_hdr[(i / g_max) * blk_words])
void launch_merge_groups(Ctx* c, const GroupTable& g, const uint64_t* rows, int64_t n, const uint8_t* kinds_host, uint32_t* d_err,
                         const uint64_t* blk_hdr = nullptr, int g_max = 0, int blk_words = 0);

// Multi-GPU merge of SMALL group sets (Q1: 4 groups per rank) without a host round trip in the middle: every rank packs
// (status words, up to g_max group keys, their raw accumulator rows) into one fixed-size block, ONE all-gather moves the
// blocks, every rank re-groups them.  Sizes are fixed, so nothing has to be known on the host before the collective.
inline int small_group_block_words(int n_keys, int n_acc, int g_max) { return 8 + g_max * (n_keys + n_acc); }
struct SmallGroupPack {
  const uint32_t* d_count = nullptr;  // nullptr: this rank scanned nothing (empty shard)
  const uint32_t* d_err = nullptr;
  const uint32_t* d_ovf = nullptr;
  const uint32_t* rep = nullptr;
  const uint32_t* slot = nullptr;
  const uint64_t* acc = nullptr;
  const uint64_t* key[MAX_KEYS] = {nullptr, nullptr, nullptr, nullptr};
  int n_keys = 0, n_acc = 0, g_max = 0;
  bool key_valid = false;
  uint64_t* block = nullptr;
};
void launch_pack_small_groups(Ctx* c, const SmallGroupPack& a);
void launch_unpack_small_groups(Ctx* c, const uint64_t* blocks, int world, int n_keys, int n_acc, int g_max, uint64_t* keys_out, uint64_t* rows_out);

struct AggFinalizeDesc {
  int32_t func, in_type, out_type, val_slot, cnt_slot;
  void* out_data;
  uint32_t* out_valid;
};
// out columns from accumulator rows: acc_base[slot_of_group(i)*n_acc + s]; slots == nullptr -> group i uses row i
// d_live (optional): a device word holding how many of the n_groups rows are real (the rest stay NULL)
void launch_agg_finalize(Ctx* c, const uint64_t* acc, const uint32_t* slots, int n_acc, int64_t n_groups,
                         const AggFinalizeDesc* descs_host, int n_desc, const uint32_t* d_live = nullptr);

}  // namespace qgpu
