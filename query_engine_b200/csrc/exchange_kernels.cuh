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
// pass 2: stable permutation (rows keep their input order inside a partition), consumed by the column gathers
// `carry` != nullptr: perm[pos] = carry[row] instead of row (a sort pass composing with the order so far)
void launch_part_perm(Ctx* c, int64_t n, int num_parts, const uint8_t* part_id, const int64_t* hist_off, uint32_t* perm,
                      const uint32_t* carry = nullptr);

}  // namespace qgpu
