// scan_kernels.cu — ahead-of-time instantiations and launchers of the fused column-scan kernels (sm_100a).
//
// The kernel body lives in scan_body.cuh.  Here it is instantiated with the interpreter program policy (program read
// from shared memory) for R in {4, 8} x five sinks; launch_scan() picks between these and the NVRTC-specialised
// kernel of the same body (jit.cu) according to the context's "scan.jit" policy:
//     0 = always interpret, 1 = always specialise, 2 (default) = specialise when the table has at least
//     "scan.jit_min_rows" rows (default 1 Mi) or the specialised kernel is already cached.
#include <cstring>

#include "jit.h"
#include "scan_body.cuh"
#include "scan_kernels.cuh"

namespace qgpu {

namespace {

constexpr int T = SCAN_THREADS;
constexpr int NW = SCAN_NW;

template <int R, int SINK>
__global__ void __launch_bounds__(SCAN_THREADS, 2) k_scan(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  scan_body<R, SINK, InterpProg>(P, InterpProg{}, smem_raw);
}

// ---------------- group table helpers ------------------------------------------------------------
__global__ void k_group_init(uint64_t* acc, size_t n, int n_acc, const uint64_t* ident) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[i] = ident[i % n_acc];
}

__global__ void k_group_collect(const uint32_t* rep, uint32_t cap, uint32_t* out_rep, uint32_t* out_slot, uint32_t* count) {
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x) {
    const uint32_t r = rep[s];
    if (r != 0) {
      const uint32_t i = atomicAdd(count, 1u);
      out_rep[i] = r - 1;
      out_slot[i] = s;
    }
  }
}

__global__ void k_gather_acc_rows(const uint64_t* acc, const uint32_t* slots, int n_acc, int64_t ng, uint64_t* out) {
  const int64_t total = ng * n_acc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = acc[(size_t)slots[i / n_acc] * n_acc + (i % n_acc)];
}

struct MergeParams {
  GroupTable g;
  const uint64_t* rows;
  int64_t n;
  uint32_t* err;
  // fixed-size blocks (merge of small group sets): row i belongs to rank i / g_max and is live when i % g_max is below
  // that rank's group count, the first word of its block header; nullptr: every row is live
  const uint64_t* blk_hdr;
  int32_t g_max, blk_words;
  uint8_t kinds[MAX_ACC];
};
__global__ void k_merge_groups(const __grid_constant__ MergeParams M) {
  uint32_t err = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M.n; i += (int64_t)gridDim.x * blockDim.x) {
    if (M.blk_hdr != nullptr && (uint64_t)(i % M.g_max) >= M.blk_hdr[(size_t)(i / M.g_max) * M.blk_words]) continue;
    const uint32_t slot = find_or_insert(M.g, i, err);
    for (int s = 0; s < M.g.n_acc; ++s) atomic_acc(M.kinds[s], M.g.acc + (size_t)slot * M.g.n_acc + s, M.rows[(size_t)i * M.g.n_acc + s]);
  }
  if (err) atomicOr(M.err, err);
}

// ---- small group sets of a multi-GPU aggregate: one fixed-size block per rank (SmallGroupBlock, scan_kernels.cuh) ----
// block = [8 header words: group count, device error word, shared-memory-tier overflow, predicate survivors, key
// validity present, 3 spare] [n_keys x g_max key words, key-major] [g_max x n_acc accumulator rows]
struct PackSmallParams {
  const uint32_t* d_count;   // groups the speculative collect found (nullptr: this rank has no rows)
  const uint32_t* d_err;     // ScanParams.err ([0] error bits, [2] predicate survivors)
  const uint32_t* d_ovf;
  const uint32_t* rep;       // representative rows / table slots of the collected groups
  const uint32_t* slot;
  const uint64_t* acc;       // the group table's accumulator rows
  const uint64_t* key[MAX_KEYS];  // 8-byte key columns
  int32_t n_keys, n_acc, g_max;
  uint32_t key_valid;        // a key column of this rank carries a validity bitmap (the merge is then not taken)
  uint64_t* block;
};
__global__ void k_pack_small_groups(const __grid_constant__ PackSmallParams P) {
  const uint32_t total = P.d_count ? *P.d_count : 0u;
  const uint32_t n = total < (uint32_t)P.g_max ? total : (uint32_t)P.g_max;
  if (threadIdx.x == 0) {
    P.block[0] = total;
    P.block[1] = P.d_err ? P.d_err[0] : 0u;
    P.block[2] = P.d_ovf ? *P.d_ovf : 0u;
    P.block[3] = P.d_err ? P.d_err[2] : 0u;
    P.block[4] = P.key_valid;
  }
  uint64_t* keys = P.block + 8;
  uint64_t* rows = keys + (size_t)P.n_keys * P.g_max;
  for (uint32_t i = threadIdx.x; i < n * (uint32_t)P.n_keys; i += blockDim.x) {
    const uint32_t k = i / n, g = i - k * n;
    keys[(size_t)k * P.g_max + g] = P.key[k][P.rep[g]];
  }
  for (uint32_t i = threadIdx.x; i < n * (uint32_t)P.n_acc; i += blockDim.x) {
    const uint32_t g = i / (uint32_t)P.n_acc, a = i - g * (uint32_t)P.n_acc;
    rows[i] = P.acc[(size_t)P.slot[g] * P.n_acc + a];
  }
}

// gathered blocks -> contiguous key columns [world * g_max] and accumulator rows [world * g_max][n_acc] (dead rows zero)
__global__ void k_unpack_small_groups(const uint64_t* blocks, int world, int n_keys, int n_acc, int g_max, int blk_words, uint64_t* keys_out,
                                      uint64_t* rows_out) {
  const int per = n_keys + n_acc;
  const int total = world * g_max * per;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i / per, f = i - row * per;
    const int r = row / g_max, g = row - r * g_max;
    const uint64_t* b = blocks + (size_t)r * blk_words;
    const bool live = (uint64_t)g < b[0];
    if (f < n_keys) keys_out[(size_t)f * world * g_max + row] = live ? b[8 + (size_t)f * g_max + g] : 0ull;
    else rows_out[(size_t)row * n_acc + (f - n_keys)] = live ? b[8 + (size_t)n_keys * g_max + (size_t)g * n_acc + (f - n_keys)] : 0ull;
  }
}

struct FinalizeParams {
  const uint64_t* acc;
  const uint32_t* slots;
  int n_acc;
  int n_desc;
  int64_t n_groups;
  const uint32_t* d_live;  // optional device word: only the first *d_live groups are real (speculative, over-sized outputs)
  AggFinalizeDesc d[MAX_ACC];
};

// evaluate_aggregate (operators.rs:745-848) on the reduced accumulators of each group
__global__ void k_agg_finalize(const __grid_constant__ FinalizeParams F) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool inb = i < F.n_groups && (F.d_live == nullptr || i < (int64_t)*F.d_live);
  const uint64_t* a = F.acc + (size_t)(inb ? (F.slots ? F.slots[i] : (uint32_t)i) : 0) * F.n_acc;
  for (int j = 0; j < F.n_desc; ++j) {
    const AggFinalizeDesc& d = F.d[j];
    bool valid = false;
    uint64_t out = 0;
    if (inb) {
      const uint64_t cnt = a[d.cnt_slot];
      switch (d.func) {
        case QGPU_COUNT: out = cnt; valid = true; break;
        case QGPU_SUM:
          valid = cnt > 0;
          out = a[d.val_slot];
          if (d.in_type == DT_I32) out = (uint64_t)(int64_t)(int32_t)(uint32_t)out;  // i32 sum wraps, then widens (:755-757)
          // Float32: arrow sums in f32 and the reference widens the f32 result (:761-763); we accumulate in f64 and
          // round once to f32, so the value is an f32 like the reference's (within f32 accumulation error of it)
          if (d.in_type == DT_F32) out = f64_bits((double)(float)bits_f64(out));
          break;
        case QGPU_AVG: {
          valid = cnt > 0;
          double s;
          if (d.in_type == DT_I64) s = (double)(int64_t)a[d.val_slot];
          else if (d.in_type == DT_I32) s = (double)(int32_t)(uint32_t)a[d.val_slot];
          else if (d.in_type == DT_F32) s = (double)(float)bits_f64(a[d.val_slot]);  // f32 sum `as f64` (:797-798)
          else s = bits_f64(a[d.val_slot]);
          out = f64_bits(valid ? s / (double)cnt : 0.0);
          break;
        }
        default: {  // MIN / MAX: decode the order key back to the input type
          valid = cnt > 0;
          const int64_t k = (int64_t)a[d.val_slot];
          if (d.in_type == DT_F64) out = f64_from_order_key(k);
          else if (d.in_type == DT_F32) out = f32_from_order_key(k);
          else out = (uint64_t)k;
          break;
        }
      }
      if (!valid) out = 0;
      if (d.out_type == DT_I64 || d.out_type == DT_F64) ((uint64_t*)d.out_data)[i] = out;
      else ((uint32_t*)d.out_data)[i] = (uint32_t)out;
    }
    const uint32_t w = __ballot_sync(0xffffffffu, inb && valid);
    if ((threadIdx.x & 31) == 0 && inb) d.out_valid[i >> 5] = w;
  }
}

int64_t pick_grid(Ctx* c, int per_sm, const ScanParams& p) {
  const int64_t want = c->opt("scan.ctas_per_sm", 0);
  if (want > 0 && want < per_sm) per_sm = (int)want;
  int64_t grid = (int64_t)c->sm_count * per_sm;
  if (grid > p.n_tiles) grid = p.n_tiles;
  return grid < 1 ? 1 : grid;
}

template <int R, int SINK>
void launch_one(Ctx* c, ScanParams& p, const char* name) {
  const size_t smem = scan_smem_bytes(SINK, R, p, false);
  auto kern = k_scan<R, SINK>;
  if (smem > 48 * 1024) QG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SCAN_THREADS, smem));
  if (per_sm < 1) throw_error(QGPU_ERR_CUDA, "scan kernel %s does not fit on an SM (smem %zu)", name, smem);
  const int64_t grid = pick_grid(c, per_sm, p);
  ProfScope ps(c, name);
  kern<<<(unsigned)grid, SCAN_THREADS, smem, c->stream>>>(p);
  QG_CUDA(cudaGetLastError());
}

// "<name>_jit", interned once per distinct operator name
const char* jit_profile_name(const char* name) {
  static thread_local const char* last_in = nullptr;
  static thread_local const char* last_out = nullptr;
  if (name != last_in) {
    last_out = intern_name(std::string(name) + "_jit");
    last_in = name;
  }
  return last_out;
}

bool launch_specialised(Ctx* c, int sink, int R, ScanParams& p, const char* name, const Program& prog) {
  const int64_t policy = c->opt("scan.jit", 2);
  if (policy == 0) return false;
  std::string key = jit_scan_key(sink, R, p, prog);
  // Feedback-directed choice of the kernel's form: the last run of this very program reported how many rows its
  // predicate kept (ScanParams.err[2], note_scan_feedback); when most rows survive, the form that fetches every column
  // up front is used (Q1), else the form whose loads are gated on the survivors (Q6).  First runs use the gated form.
  c->last_scan_key = key;
  bool eager = false;
  if ((sink == SK_SCALAR || sink == SK_GROUP_SMEM) && c->opt("scan.eager_loads", 0) != 0) {
    auto it = c->scan_feedback.find(key);
    eager = it != c->scan_feedback.end() && it->second >= 0.5f;
  }
  key.push_back(eager ? 'E' : 'G');
  // __launch_bounds__ minimum blocks: 2 (128 registers) for the streaming scans; the partition pass of the two-level
  // GROUP BY is issue-bound and wants four CTAs per SM (64 registers: 1.81 ms against 2.11 ms at three)
  const int min_blocks = (int)std::min<int64_t>(4, std::max<int64_t>(1, sink == SK_GROUP_PART ? c->opt("group.jit_min_blocks", 4)
                                                                                               : c->opt("scan.jit_min_blocks", 2)));
  key.push_back((char)('0' + min_blocks));
  if (policy == 2 && p.n_rows < c->opt("scan.jit_min_rows", 1 << 20) && !jit_cached(key)) return false;
  JitKernel* k = jit_get(c, key, [&] { return jit_scan_source(sink, R, p, prog, eager, min_blocks); }, "qgpu_jit_scan");
  const size_t smem = scan_smem_bytes(sink, R, p, true);
  const int per_sm = jit_prepare(c, k, SCAN_THREADS, smem);
  if (per_sm < 1) throw_error(QGPU_ERR_CUDA, "specialised scan kernel %s does not fit on an SM (smem %zu)", name, smem);
  const int64_t grid = pick_grid(c, per_sm, p);
  void* args[] = {(void*)&p};
  // distinct profile name: the same operator, specialised
  ProfScope ps(c, jit_profile_name(name));
  jit_launch(c, k, (unsigned)grid, SCAN_THREADS, smem, args);
  return true;
}

}  // namespace

void note_scan_feedback(Ctx* c, uint64_t kept, uint64_t rows) {
  if (c->last_scan_key.empty() || rows == 0) return;
  c->scan_feedback[c->last_scan_key] = (float)((double)kept / (double)rows);
}

size_t scan_smem_bytes(int sink, int R, const ScanParams& p, bool specialised) {
  size_t b = 0;
  if (!specialised) {
    b += (size_t)p.n_insn * sizeof(Insn);
    b += (size_t)p.stack_depth * R * T * 8 + (size_t)p.stack_depth * T * 4;
  }
  switch (sink) {
    case SK_MASK: b += 16; break;
    case SK_SCALAR: b += (specialised ? 0 : (size_t)p.n_acc * T * 8) + (size_t)p.n_acc * NW * 8 + 16; break;
    case SK_GROUP_SMEM:
      b += (size_t)p.gp * p.n_acc * T * 8 + (size_t)p.gp * p.n_keys * 8 + (size_t)p.gp * 8 + 16 + (size_t)p.gp * 4;
      break;
    case SK_GROUP_PART:
      b += (size_t)(p.part_slots + 1) * 8 + (size_t)(p.part_slots + 1) * p.n_acc * 8 + 16 + (size_t)(p.part_slots + 2) * 2;
      break;
    default: break;
  }
  return (b + 15) & ~(size_t)15;
}

void launch_scan(Ctx* c, int sink, int R, ScanParams& p, const char* name, const Program* prog) {
  const int64_t tile = (int64_t)SCAN_THREADS * R;
  p.n_tiles = (p.n_rows + tile - 1) / tile;
  if (p.n_tiles == 0) p.n_tiles = 1;
  if (prog != nullptr && launch_specialised(c, sink, R, p, name, *prog)) return;
#define QG_CASE(RR, SK) \
  if (R == RR && sink == SK) return launch_one<RR, SK>(c, p, name);
  QG_CASE(4, SK_MASK) QG_CASE(8, SK_MASK)
  QG_CASE(4, SK_SCALAR) QG_CASE(8, SK_SCALAR)
  QG_CASE(4, SK_GROUP_SMEM) QG_CASE(8, SK_GROUP_SMEM)
  QG_CASE(4, SK_GROUP_GLOBAL) QG_CASE(8, SK_GROUP_GLOBAL)
  QG_CASE(4, SK_STORE) QG_CASE(8, SK_STORE)
  QG_CASE(4, SK_GROUP_PART) QG_CASE(8, SK_GROUP_PART)
  QG_CASE(4, SK_DIAG) QG_CASE(8, SK_DIAG)
#undef QG_CASE
  throw_error(QGPU_ERR_INVALID, "no scan kernel for rows_per_thread=%d sink=%d", R, sink);
}

void launch_group_init(Ctx* c, const GroupTable& gt, const uint8_t* kinds) {
  uint64_t ident[MAX_ACC];
  for (int i = 0; i < gt.n_acc; ++i) ident[i] = acc_identity(kinds[i]);
  DevBuf d_ident(c, sizeof ident);
  QG_CUDA(cudaMemcpyAsync(d_ident.p, ident, sizeof ident, cudaMemcpyHostToDevice, c->stream));  // pageable: staged on return
  QG_CUDA(cudaMemsetAsync(gt.rep, 0, (size_t)gt.cap * 4, c->stream));
  const size_t n = (size_t)gt.cap * gt.n_acc;
  int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 8);
  if (grid < 1) grid = 1;
  ProfScope ps(c, "k_group_init");
  k_group_init<<<grid, 256, 0, c->stream>>>(gt.acc, n, gt.n_acc, d_ident.as<uint64_t>());
  QG_CUDA(cudaGetLastError());
}

void launch_group_collect(Ctx* c, const GroupTable& gt, uint32_t* out_rep, uint32_t* out_slot, uint32_t* d_count) {
  QG_CUDA(cudaMemsetAsync(d_count, 0, 4, c->stream));
  int grid = (int)std::min<size_t>(((size_t)gt.cap + 255) / 256, (size_t)c->sm_count * 8);
  ProfScope ps(c, "k_group_collect");
  k_group_collect<<<grid, 256, 0, c->stream>>>(gt.rep, gt.cap, out_rep, out_slot, d_count);
  QG_CUDA(cudaGetLastError());
}

void launch_gather_acc_rows(Ctx* c, const uint64_t* acc, const uint32_t* slots, int n_acc, int64_t ng, uint64_t* out) {
  if (ng <= 0) return;
  int grid = (int)std::min<int64_t>((ng * n_acc + 255) / 256, (int64_t)c->sm_count * 8);
  ProfScope ps(c, "k_gather_acc_rows");
  k_gather_acc_rows<<<grid, 256, 0, c->stream>>>(acc, slots, n_acc, ng, out);
  QG_CUDA(cudaGetLastError());
}

void launch_pack_small_groups(Ctx* c, const SmallGroupPack& a) {
  PackSmallParams P;
  memset(&P, 0, sizeof P);
  P.d_count = a.d_count;
  P.d_err = a.d_err;
  P.d_ovf = a.d_ovf;
  P.rep = a.rep;
  P.slot = a.slot;
  P.acc = a.acc;
  for (int k = 0; k < a.n_keys; ++k) P.key[k] = a.key[k];
  P.n_keys = a.n_keys;
  P.n_acc = a.n_acc;
  P.g_max = a.g_max;
  P.key_valid = a.key_valid ? 1u : 0u;
  P.block = a.block;
  ProfScope ps(c, "k_pack_small_groups");
  k_pack_small_groups<<<1, 256, 0, c->stream>>>(P);
  QG_CUDA(cudaGetLastError());
}

void launch_unpack_small_groups(Ctx* c, const uint64_t* blocks, int world, int n_keys, int n_acc, int g_max, uint64_t* keys_out, uint64_t* rows_out) {
  const int total = world * g_max * (n_keys + n_acc);
  ProfScope ps(c, "k_unpack_small_groups");
  k_unpack_small_groups<<<(total + 255) / 256, 256, 0, c->stream>>>(blocks, world, n_keys, n_acc, g_max, small_group_block_words(n_keys, n_acc, g_max),
                                                                    keys_out, rows_out);
  QG_CUDA(cudaGetLastError());
}

void launch_merge_groups(Ctx* c, const GroupTable& g, const uint64_t* rows, int64_t n, const uint8_t* kinds_host, uint32_t* d_err,
                         const uint64_t* blk_hdr, int g_max, int blk_words) {
  if (n <= 0) return;
  MergeParams M;
  M.g = g;
  M.rows = rows;
  M.n = n;
  M.err = d_err;
  M.blk_hdr = blk_hdr;
  M.g_max = g_max;
  M.blk_words = blk_words;
  for (int i = 0; i < g.n_acc && i < MAX_ACC; ++i) M.kinds[i] = kinds_host[i];
  int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
  ProfScope ps(c, "k_merge_groups");
  k_merge_groups<<<grid, 256, 0, c->stream>>>(M);
  QG_CUDA(cudaGetLastError());
}

void launch_agg_finalize(Ctx* c, const uint64_t* acc, const uint32_t* slots, int n_acc, int64_t n_groups,
                         const AggFinalizeDesc* descs, int n_desc, const uint32_t* d_live) {
  if (n_desc > MAX_ACC) throw_error(QGPU_ERR_UNSUPPORTED, "too many aggregates (%d)", n_desc);
  FinalizeParams F;
  F.acc = acc;
  F.slots = slots;
  F.n_acc = n_acc;
  F.n_desc = n_desc;
  F.n_groups = n_groups;
  F.d_live = d_live;
  for (int i = 0; i < n_desc; ++i) F.d[i] = descs[i];
  const int64_t blocks = (n_groups + 255) / 256;
  ProfScope ps(c, "k_agg_finalize");
  k_agg_finalize<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, c->stream>>>(F);
  QG_CUDA(cudaGetLastError());
}

}  // namespace qgpu
