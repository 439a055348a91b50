// sort_kernels.cuh — device lexsort (see sort_kernels.cu)
#pragma once
#include "qgpu_internal.h"

namespace qgpu {

// perm[0..n) = the row order of a lexicographic sort by `cols` (first column most significant); per column: descending,
// nulls_first (arrow SortOptions).  Fixed-width and boolean columns; ties keep their input order (stable).
void sort_permutation(Ctx* c, const std::vector<ColRef>& cols, const std::vector<int>& descending, const std::vector<int>& nulls_first,
                      int64_t n, uint32_t* perm);

}  // namespace qgpu
