// jit.h — run-time specialisation of the scan kernels with NVRTC (see jit.cu).
#pragma once
#include <functional>
#include <string>

#include "qgpu_internal.h"
#include "scan_params.cuh"

namespace qgpu {

struct JitKernel;

// CUDA source of scan_body<R, sink, JitProg> specialised on `prog` and on the column types / validity of `p`
// `eager`: the aggregate scans also carry the fetch-everything-up-front copy of the tile code (scan_body.cuh)
// `min_blocks`: the kernel's __launch_bounds__ minimum CTAs per SM (register cap)
std::string jit_scan_source(int sink, int R, const ScanParams& p, const Program& prog, bool eager = false, int min_blocks = 2);
// compact binary key of everything jit_scan_source depends on (cache lookups do not build the source text)
std::string jit_scan_key(int sink, int R, const ScanParams& p, const Program& prog);
// compile only (works without a GPU): returns the cubin, appends the NVRTC log to *log when given
std::string jit_compile(const std::string& src, std::string* log);
bool jit_cached(const std::string& key);
// `make_source` is only called on a cache miss
JitKernel* jit_get(Ctx* c, const std::string& key, const std::function<std::string()>& make_source, const char* entry);
// raises the dynamic shared memory limit when needed; returns resident CTAs per SM at (threads, smem)
int jit_prepare(Ctx* c, JitKernel* k, int threads, size_t smem);
void jit_launch(Ctx* c, JitKernel* k, unsigned grid, unsigned block, size_t smem, void** args);
// stable storage for profile names built at run time
const char* intern_name(const std::string& s);

}  // namespace qgpu
