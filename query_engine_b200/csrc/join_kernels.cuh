// join_kernels.cuh — launchers of the hash-join kernels.
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

struct __align__(16) JoinSlot {
  int64_t key;
  uint32_t row1;  // build row + 1, 0 = empty
  uint32_t pad;
};

void launch_join_build(Ctx* c, const ColRef& key, int64_t n, JoinSlot* tab, uint32_t cap);
int64_t join_tiles(int64_t n);
void launch_join_count(Ctx* c, const ColRef& key, int64_t n, const JoinSlot* tab, uint32_t cap, bool left_join,
                       uint32_t* first_match, uint32_t* tile_counts, uint32_t* dup_flag);
void launch_join_emit(Ctx* c, const ColRef& key, int64_t n, const JoinSlot* tab, uint32_t cap, bool left_join, bool has_dups,
                      const uint32_t* first_match, const int64_t* tile_off, uint32_t* out_left, uint32_t* out_right);

}  // namespace qgpu
