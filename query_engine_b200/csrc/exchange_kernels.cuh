// exchange_kernels.cuh — launchers of the hash-partition kernels.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

constexpr int MAX_PARTS = 128;

struct PartKeys {
  int32_t n_keys;
  int32_t pad;
  ColRef keys[MAX_KEYS];
};

int64_t part_tiles(int64_t n);
// hist is partition-major: hist[p * n_tiles + tile]
void launch_part_hist(Ctx* c, const PartKeys& K, int64_t n, int num_parts, bool use_sip, uint8_t* part_id, uint32_t* hist);
// pass 2, materialising form: scatter up to MAX_SCATTER_COLS fixed-width columns straight to their partition-major
// positions (coalesced reads; a tile writes one contiguous run per partition).  Rows keep their order in a partition.
constexpr int MAX_SCATTER_COLS = 16;
struct ScatterCol {
  const void* in;
  void* out;
  int32_t width;  // 4 or 8
  int32_t pad;
};
void launch_part_scatter(Ctx* c, int64_t n, int num_parts, const uint8_t* part_id, const int64_t* hist_off, const ScatterCol* cols,
                         int n_cols);
void launch_part_perm(Ctx* c, int64_t n, int num_parts, const uint8_t* part_id, const int64_t* hist_off, uint32_t* perm);

}  // namespace qgpu
