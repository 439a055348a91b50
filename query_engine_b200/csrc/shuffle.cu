// shuffle.cu — the in-box multi-GPU exchange: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Replaces what the reference's DistributedExecutor does between stages — Exchange::execute on the stage input
// followed by Merge::concat of the partition outputs (crates/query-distributed/src/executor.rs:230-263,
// operators.rs:58-61,136-138) — for the two shuffle shapes its planner emits: shuffle-by-group-key before the final
// aggregate and shuffle-both-inputs before a join (planner.rs:200-249, 281-302).
//
//   1. rows are ranked into `world` partitions by key hash (exchange_kernels.cu) and gathered partition-major, so
//      the rows bound for peer p are one contiguous slice of every column;
//   2. the world x world row-count matrix is all-gathered (8 int64 per rank);
//   3. per column buffer, one grouped ncclSend/ncclRecv all-to-all moves the slices; a rank's own slice goes
//      through the same group (NCCL turns it into a local copy);
//   4. arrivals are laid out in source-rank order == Merge::concat of the partition outputs.
//
// libnccl is dlopen'ed on first use so that single-GPU users (and the CPU-only symbol tests) do not need it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "exchange_kernels.cuh"
#include "qgpu_internal.h"
#include "util_kernels.cuh"

namespace qgpu {
namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
#define QG_SYM(field, name)                                   \
  api.field = (decltype(api.field))dlsym(h, name);            \
  if (!api.field) err = std::string("libnccl lacks ") + name;
    QG_SYM(GetUniqueId, "ncclGetUniqueId")
    QG_SYM(CommInitRank, "ncclCommInitRank")
    QG_SYM(CommDestroy, "ncclCommDestroy")
    QG_SYM(GetErrorString, "ncclGetErrorString")
    QG_SYM(GroupStart, "ncclGroupStart")
    QG_SYM(GroupEnd, "ncclGroupEnd")
    QG_SYM(Send, "ncclSend")
    QG_SYM(Recv, "ncclRecv")
    QG_SYM(AllGather, "ncclAllGather")
#undef QG_SYM
  });
  if (!err.empty()) throw_error(QGPU_ERR_CUDA, "%s", err.c_str());
  return &api;
}

#define QG_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t _r = (call);                                                                      \
    if (_r != ncclSuccess)                                                                         \
      ::qgpu::throw_error(QGPU_ERR_CUDA, "NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl()->GetErrorString(_r)); \
  } while (0)

}  // namespace

// all-gather `n` int64 per rank through the pinned scratch
std::vector<int64_t> allgather_i64(Ctx* c, const std::vector<int64_t>& mine) {
  const size_t n = mine.size();
  const int W = c->world;
  DevBuf d_in(c, n * 8 + 16), d_out(c, n * 8 * W + 16);
  int64_t* h = (int64_t*)c->scratch(n * 8 * W + 16);
  memcpy(h, mine.data(), n * 8);
  QG_CUDA(cudaMemcpyAsync(d_in.p, h, n * 8, cudaMemcpyHostToDevice, c->stream));
  QG_NCCL(nccl()->AllGather(d_in.p, d_out.p, n, ncclInt64, (ncclComm_t)c->nccl_comm, c->stream));
  QG_CUDA(cudaMemcpyAsync(h, d_out.p, n * 8 * W, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));
  return std::vector<int64_t>(h, h + n * W);
}

namespace {

// one grouped all-to-all of byte slices: send [soff[p], soff[p+1]) of `src` to p, receive into [roff[p], roff[p+1])
void all_to_all_bytes(Ctx* c, const char* src, const std::vector<int64_t>& soff, char* dst, const std::vector<int64_t>& roff,
                      int64_t width) {
  NcclApi* N = nccl();
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  ProfScope ps(c, "nccl_all_to_all");
  QG_NCCL(N->GroupStart());
  for (int p = 0; p < c->world; ++p) {
    const int64_t sb = (soff[p + 1] - soff[p]) * width, rb = (roff[p + 1] - roff[p]) * width;
    if (sb > 0) QG_NCCL(N->Send(src + soff[p] * width, (size_t)sb, ncclUint8, p, comm, c->stream));
    if (rb > 0) QG_NCCL(N->Recv(dst + roff[p] * width, (size_t)rb, ncclUint8, p, comm, c->stream));
  }
  QG_NCCL(N->GroupEnd());
}

void need_comm(Ctx* c) {
  if (!c->nccl_comm) throw_error(QGPU_ERR_INVALID, "no communicator: call qgpu_comm_init first");
}

void reject_utf8(const Table& in, const char* what) {
  for (const Column& col : in.cols)
    if (col.dtype == DT_UTF8)
      throw_error(QGPU_ERR_UNSUPPORTED, "%s of Utf8 columns is not on the accelerated path (dictionary-encode them first)", what);
}

// the receive side of steps 3-4 for a table whose rows are already partition-major
std::shared_ptr<Table> exchange_slices(Ctx* c, const Table& src, const std::vector<int64_t>& soff,
                                       const std::vector<int64_t>& roff) {
  const int W = c->world;
  const int64_t n_in = soff[W], n_out = roff[W];
  auto out = std::make_shared<Table>();
  out->adopt(c);
  out->fields = src.fields;
  out->n_rows = n_out;
  out->batch_offsets = {0};
  if (n_out > 0) out->batch_offsets.push_back(n_out);
  // does any rank carry a validity bitmap for column ci?  (must agree across ranks: it adds a transfer)
  std::vector<int64_t> has_valid(src.cols.size());
  for (size_t ci = 0; ci < src.cols.size(); ++ci) has_valid[ci] = src.cols[ci].valid != nullptr ? 1 : 0;
  const std::vector<int64_t> all_valid = allgather_i64(c, has_valid);
  for (size_t ci = 0; ci < src.cols.size(); ++ci) {
    const Column& s = src.cols[ci];
    bool any_valid = false;
    for (int r = 0; r < W; ++r) any_valid |= all_valid[(size_t)r * src.cols.size() + ci] != 0;
    Column d;
    if (s.dtype == DT_NULL) {
      d.dtype = DT_NULL;
      d.n = n_out;
    } else if (s.dtype == DT_BOOL) {
      // bit-packed values travel as one byte per row
      d = alloc_column(c, DT_BOOL, n_out, false);
      DevBuf sb(c, pad_bytes((size_t)n_in)), rb(c, pad_bytes((size_t)n_out));
      launch_bits_to_bytes(c, (const uint32_t*)s.data, n_in, sb.as<uint8_t>());
      all_to_all_bytes(c, sb.as<char>(), soff, rb.as<char>(), roff, 1);
      launch_bytes_to_bits(c, rb.as<uint8_t>(), n_out, (uint32_t*)d.data);
    } else {
      const int w = dtype_width(s.dtype);
      d = alloc_column(c, s.dtype, n_out, false);
      all_to_all_bytes(c, (const char*)s.data, soff, (char*)d.data, roff, w);
    }
    if (any_valid && s.dtype != DT_NULL) {
      DevBuf sb(c, pad_bytes((size_t)n_in)), rb(c, pad_bytes((size_t)n_out));
      launch_bits_to_bytes(c, s.valid, n_in, sb.as<uint8_t>());
      all_to_all_bytes(c, sb.as<char>(), soff, rb.as<char>(), roff, 1);
      d.valid = (uint32_t*)c->zalloc(bitmap_bytes(n_out));
      launch_bytes_to_bits(c, rb.as<uint8_t>(), n_out, d.valid);
      d.null_count = -1;
      out->fields[ci].nullable = true;
    }
    out->cols.push_back(d);
  }
  c->sync();
  return out;
}

}  // namespace

void comm_unique_id(void* out_id128) {
  static_assert(sizeof(ncclUniqueId) == QGPU_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  QG_NCCL(nccl()->GetUniqueId(&id));
  memcpy(out_id128, &id, sizeof id);
}

void comm_init(Ctx* c, const void* id128, int rank, int world) {
  if (world < 1 || rank < 0 || rank >= world) throw_error(QGPU_ERR_INVALID, "bad rank %d / world size %d", rank, world);
  if (world > MAX_PARTS) throw_error(QGPU_ERR_UNSUPPORTED, "world size above %d", MAX_PARTS);
  comm_destroy(c);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t comm = nullptr;
  QG_NCCL(nccl()->CommInitRank(&comm, world, id, rank));
  c->nccl_comm = comm;
  c->rank = rank;
  c->world = world;
}

// ---- exchange arena -------------------------------------------------------------------------------------------------
// One cudaMalloc'ed buffer per GPU, exported with CUDA IPC and mapped by every other process of the box, so that a
// kernel on any GPU can store (and atomicAdd) straight into it over NVLink.  (Re)allocation is a collective: all
// ranks call arena_ensure with the same size.
// (CUDA leaves freeing an exported allocation while an importer still maps it undefined, hence the order: every rank
// unmaps its peers, a host-synchronised barrier, then every rank frees its own buffer.)
void arena_release(Ctx* c, bool collective) {
  bool had_peers = false;
  for (int r = 0; r < 8; ++r) {
    if (c->arena_peer[r] && r != c->rank) {
      cudaIpcCloseMemHandle(c->arena_peer[r]);
      had_peers = true;
    }
    c->arena_peer[r] = nullptr;
  }
  // growing the arena is a collective, so the barrier is safe there; at shutdown the ranks may leave at different
  // times (or after a failure) and the process exit reclaims everything: no barrier that could hang
  if (collective && had_peers && c->nccl_comm) {
    try {
      stream_barrier(c);
      cudaStreamSynchronize(c->stream);
    } catch (const Error&) {
      cudaGetLastError();  // shutting down after a failure: free anyway
    }
  }
  if (c->arena) cudaFree(c->arena);
  c->arena = nullptr;
  c->arena_bytes = 0;
}

void arena_ensure(Ctx* c, size_t bytes) {
  need_comm(c);
  if (c->world > 8) throw_error(QGPU_ERR_UNSUPPORTED, "exchange arena for more than 8 GPUs");
  if (c->arena_bytes >= bytes) return;
  QG_CUDA(cudaStreamSynchronize(c->stream));
  stream_barrier(c);  // nobody still writes into an arena that is about to go
  QG_CUDA(cudaStreamSynchronize(c->stream));
  arena_release(c, true);
  const size_t round = (size_t)std::max<int64_t>(1 << 16, c->opt("exchange.arena_round", 64 << 20));
  bytes = (bytes + round - 1) / round * round;
  QG_CUDA(cudaMalloc(&c->arena, bytes));
  c->arena_bytes = bytes;
  cudaIpcMemHandle_t mine;
  QG_CUDA(cudaIpcGetMemHandle(&mine, c->arena));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  const int W = c->world;
  DevBuf d_in(c, 64 + 16), d_out(c, (size_t)64 * W + 16);
  char* h = (char*)c->scratch((size_t)64 * W + 16);
  memcpy(h, &mine, 64);
  QG_CUDA(cudaMemcpyAsync(d_in.p, h, 64, cudaMemcpyHostToDevice, c->stream));
  QG_NCCL(nccl()->AllGather(d_in.p, d_out.p, 64, ncclUint8, (ncclComm_t)c->nccl_comm, c->stream));
  QG_CUDA(cudaMemcpyAsync(h, d_out.p, (size_t)64 * W, cudaMemcpyDeviceToHost, c->stream));
  QG_CUDA(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < W; ++r) {
    if (r == c->rank) {
      c->arena_peer[r] = c->arena;
      continue;
    }
    cudaIpcMemHandle_t hr;
    memcpy(&hr, h + (size_t)64 * r, 64);
    QG_CUDA(cudaIpcOpenMemHandle(&c->arena_peer[r], hr, cudaIpcMemLazyEnablePeerAccess));
  }
}

void stream_barrier(Ctx* c) {
  need_comm(c);
  DevBuf d_in(c, 256, true), d_out(c, (size_t)8 * c->world + 256);
  ProfScope ps(c, "nccl_barrier");
  QG_NCCL(nccl()->AllGather(d_in.p, d_out.p, 8, ncclUint8, (ncclComm_t)c->nccl_comm, c->stream));
}

// all-gather of `bytes` per rank between device buffers, stream-ordered (no host synchronisation)
void allgather_device(Ctx* c, const void* d_in, void* d_out, size_t bytes) {
  need_comm(c);
  ProfScope ps(c, "nccl_all_gather_small");
  QG_NCCL(nccl()->AllGather(d_in, d_out, bytes, ncclUint8, (ncclComm_t)c->nccl_comm, c->stream));
}

void comm_destroy(Ctx* c) {
  arena_release(c, false);
  if (c->nccl_comm) {
    nccl()->CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  c->rank = 0;
  c->world = 1;
}

std::shared_ptr<Table> op_shuffle_hash(Ctx* c, const Table& in, const std::vector<int>& key_cols) {
  need_comm(c);
  reject_utf8(in, "the NCCL shuffle");
  if (key_cols.empty()) throw_error(QGPU_ERR_PARTITION, "No key columns found in batch");  // partition.rs:173-177
  if ((int)key_cols.size() > MAX_KEYS) throw_error(QGPU_ERR_UNSUPPORTED, "more than %d partition keys", MAX_KEYS);
  const int W = c->world;
  const int64_t n = in.n_rows;

  // 1. rank rows into W partitions (mix64 hash: partition placement is not observable in query results,
  //    SURVEY.md 8a13) and gather them partition-major
  std::vector<int64_t> soff((size_t)W + 1, 0);
  std::shared_ptr<Table> sorted;
  if (n > 0) {
    PartKeys K;
    memset(&K, 0, sizeof K);
    K.n_keys = (int)key_cols.size();
    for (int i = 0; i < K.n_keys; ++i) K.keys[i] = colref_of(in.cols[key_cols[i]]);
    const int64_t n_tiles = part_tiles(n);
    DevBuf part_id(c, pad_bytes((size_t)n));
    DevBuf hist(c, (size_t)W * n_tiles * 4 + 256);
    DevBuf hist_off(c, ((size_t)W * n_tiles + 1) * 8 + 256);
    DevBuf perm(c, pad_bytes((size_t)n * 4));
    launch_part_hist(c, K, n, W, false, part_id.as<uint8_t>(), hist.as<uint32_t>());
    scan_u32_to_i64(c, hist.as<uint32_t>(), (int64_t)W * n_tiles, hist_off.as<int64_t>());
    launch_part_perm(c, n, W, part_id.as<uint8_t>(), hist_off.as<int64_t>(), perm.as<uint32_t>());
    int64_t* h = (int64_t*)c->scratch((size_t)(W + 1) * 8);
    DevBuf d_starts(c, (size_t)(W + 1) * 8 + 256);
    launch_gather_strided_i64(c, hist_off.as<int64_t>(), n_tiles, W + 1, d_starts.as<int64_t>());
    QG_CUDA(cudaMemcpyAsync(h, d_starts.p, (size_t)(W + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    QG_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(soff.data(), h, (size_t)(W + 1) * 8);
    std::vector<int> all(in.cols.size());
    for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
    sorted = gather_rows(c, in, all, perm.as<uint32_t>(), n, false);
  }
  const Table& src = sorted ? *sorted : in;

  // 2. counts[r][p] = rows rank r sends to p
  std::vector<int64_t> mine((size_t)W);
  for (int p = 0; p < W; ++p) mine[p] = soff[p + 1] - soff[p];
  const std::vector<int64_t> counts = allgather_i64(c, mine);
  std::vector<int64_t> roff((size_t)W + 1, 0);
  for (int r = 0; r < W; ++r) roff[r + 1] = roff[r] + counts[(size_t)r * W + c->rank];
  if (roff[W] >= 0xfffffff0ll) throw_error(QGPU_ERR_UNSUPPORTED, "shuffle would deliver 2^32 rows or more to one GPU");

  // 3.-4.
  auto out = exchange_slices(c, src, soff, roff);
  if (out->n_rows > 0) return out;
  // "shuffle.keep_empty": a rank that received nothing keeps a typed table of one zero-row batch - the build side of a
  // multi-GPU LEFT join must reach the local join even when empty (its probe rows come out with NULL right columns)
  if (c->opt("shuffle.keep_empty", 0) == 0) return nullptr;
  out->batch_offsets = {0, 0};
  return out;
}

namespace {
std::shared_ptr<Table> allgather_table(Ctx* c, const Table& in, const std::vector<int64_t>& ns, const std::vector<int64_t>& all_valid);
}

// finals: [n_acc] accumulators + [1] device error word per rank -> folded over ranks in rank order, in place
void merge_scalar_across_ranks(Ctx* c, uint64_t* finals, int n_acc, const uint8_t* kinds) {
  need_comm(c);
  DevBuf all(c, (size_t)c->world * (n_acc + 1) * 8 + 256);
  {
    ProfScope ps(c, "nccl_all_gather_acc");
    QG_NCCL(nccl()->AllGather(finals, all.p, (size_t)(n_acc + 1), ncclUint64, (ncclComm_t)c->nccl_comm, c->stream));
  }
  launch_fold_finals(c, all.as<uint64_t>(), c->world, n_acc, kinds, finals);
}

// all-gather of every rank's (group key rows, raw accumulator rows, error word); returns the key table of all ranks
// (rank order) and fills all_acc [NG][n_acc]
std::shared_ptr<Table> allgather_groups(Ctx* c, const Table& local_keys, const uint64_t* local_acc, int n_acc, uint32_t local_err,
                                        DevBuf& all_acc, uint32_t& any_err) {
  need_comm(c);
  const int W = c->world;
  const int64_t ng = local_keys.n_rows;
  // one host round trip: group count, error word and the key columns' validity flags of every rank
  std::vector<int64_t> mine{ng, (int64_t)local_err};
  for (const Column& col : local_keys.cols) mine.push_back(col.valid != nullptr ? 1 : 0);
  const size_t ms = mine.size(), nk = local_keys.cols.size();
  const std::vector<int64_t> meta_all = allgather_i64(c, mine);
  std::vector<int64_t> meta((size_t)W * 2), key_ns((size_t)W), key_valid((size_t)W * nk);
  for (int r = 0; r < W; ++r) {
    meta[(size_t)r * 2] = meta_all[(size_t)r * ms];
    meta[(size_t)r * 2 + 1] = meta_all[(size_t)r * ms + 1];
    key_ns[(size_t)r] = meta_all[(size_t)r * ms];
    for (size_t k = 0; k < nk; ++k) key_valid[(size_t)r * nk + k] = meta_all[(size_t)r * ms + 2 + k];
  }
  std::vector<int64_t> roff((size_t)W + 1, 0);
  any_err = 0;
  for (int r = 0; r < W; ++r) {
    roff[r + 1] = roff[r] + meta[(size_t)r * 2];
    any_err |= (uint32_t)meta[(size_t)r * 2 + 1];
  }
  const int64_t NG = roff[W];
  if (NG == 0) return nullptr;
  all_acc = DevBuf(c, (size_t)NG * n_acc * 8 + 256);
  {
    NcclApi* N = nccl();
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const size_t w = (size_t)n_acc * 8;
    ProfScope ps(c, "nccl_all_gather_acc");
    QG_NCCL(N->GroupStart());
    for (int p = 0; p < W; ++p) {
      if (ng > 0) QG_NCCL(N->Send(local_acc, (size_t)ng * w, ncclUint8, p, comm, c->stream));
      const int64_t cnt = meta[(size_t)p * 2];
      if (cnt > 0) QG_NCCL(N->Recv(all_acc.as<char>() + (size_t)roff[p] * w, (size_t)cnt * w, ncclUint8, p, comm, c->stream));
    }
    QG_NCCL(N->GroupEnd());
  }
  reject_utf8(local_keys, "all-gather");
  return allgather_table(c, local_keys, key_ns, key_valid);
}

std::shared_ptr<Table> op_allgather(Ctx* c, const Table& in) {
  need_comm(c);
  reject_utf8(in, "all-gather");
  // one host round trip for everything the ranks must know of each other: row count, which columns carry validity
  std::vector<int64_t> mine{in.n_rows};
  for (const Column& col : in.cols) mine.push_back(col.valid != nullptr ? 1 : 0);
  const std::vector<int64_t> all = allgather_i64(c, mine);
  const size_t stride = mine.size();
  std::vector<int64_t> ns((size_t)c->world), all_valid((size_t)c->world * in.cols.size());
  for (int r = 0; r < c->world; ++r) {
    ns[(size_t)r] = all[(size_t)r * stride];
    for (size_t ci = 0; ci < in.cols.size(); ++ci) all_valid[(size_t)r * in.cols.size() + ci] = all[(size_t)r * stride + 1 + ci];
  }
  return allgather_table(c, in, ns, all_valid);
}

namespace {
// every rank sends its whole table to every rank; `ns` = rows per rank, `all_valid[r * n_cols + ci]` = rank r's column
// ci carries a validity bitmap (both already known to every rank)
std::shared_ptr<Table> allgather_table(Ctx* c, const Table& in, const std::vector<int64_t>& ns, const std::vector<int64_t>& all_valid) {
  const int W = c->world;
  std::vector<int64_t> roff((size_t)W + 1, 0);
  for (int r = 0; r < W; ++r) roff[r + 1] = roff[r] + ns[r];
  const int64_t n_out = roff[W];
  if (n_out >= 0xfffffff0ll) throw_error(QGPU_ERR_UNSUPPORTED, "all-gather of 2^32 rows or more");
  auto out = std::make_shared<Table>();
  out->adopt(c);
  out->fields = in.fields;
  out->n_rows = n_out;
  out->batch_offsets = {0};
  if (n_out > 0) out->batch_offsets.push_back(n_out);
  NcclApi* N = nccl();
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  auto bcast_all = [&](const char* src, char* dst, int64_t width) {
    ProfScope ps(c, "nccl_all_gather_v");
    QG_NCCL(N->GroupStart());
    for (int p = 0; p < W; ++p) {
      if (in.n_rows > 0) QG_NCCL(N->Send(src, (size_t)(in.n_rows * width), ncclUint8, p, comm, c->stream));
      if (ns[p] > 0) QG_NCCL(N->Recv(dst + roff[p] * width, (size_t)(ns[p] * width), ncclUint8, p, comm, c->stream));
    }
    QG_NCCL(N->GroupEnd());
  };
  for (size_t ci = 0; ci < in.cols.size(); ++ci) {
    const Column& s = in.cols[ci];
    bool any_valid = false;
    for (int r = 0; r < W; ++r) any_valid |= all_valid[(size_t)r * in.cols.size() + ci] != 0;
    Column d;
    if (s.dtype == DT_NULL) {
      d.dtype = DT_NULL;
      d.n = n_out;
    } else if (s.dtype == DT_BOOL) {
      d = alloc_column(c, DT_BOOL, n_out, false);
      DevBuf sb(c, pad_bytes((size_t)in.n_rows)), rb(c, pad_bytes((size_t)n_out));
      launch_bits_to_bytes(c, (const uint32_t*)s.data, in.n_rows, sb.as<uint8_t>());
      bcast_all(sb.as<char>(), rb.as<char>(), 1);
      launch_bytes_to_bits(c, rb.as<uint8_t>(), n_out, (uint32_t*)d.data);
    } else {
      d = alloc_column(c, s.dtype, n_out, false);
      bcast_all((const char*)s.data, (char*)d.data, dtype_width(s.dtype));
    }
    if (any_valid && s.dtype != DT_NULL) {
      DevBuf sb(c, pad_bytes((size_t)in.n_rows)), rb(c, pad_bytes((size_t)n_out));
      launch_bits_to_bytes(c, s.valid, in.n_rows, sb.as<uint8_t>());
      bcast_all(sb.as<char>(), rb.as<char>(), 1);
      d.valid = (uint32_t*)c->zalloc(bitmap_bytes(n_out));
      launch_bytes_to_bits(c, rb.as<uint8_t>(), n_out, d.valid);
      d.null_count = -1;
      out->fields[ci].nullable = true;
    }
    out->cols.push_back(d);
  }
  return out->n_rows > 0 ? out : nullptr;
}
}  // namespace

}  // namespace qgpu
