// qgpu_internal.h — host-side structures of libqgpu shared by the translation units.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <ctime>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/qgpu.h"
#include "common.cuh"

namespace qgpu {

[[noreturn]] void throw_error(int code, const char* fmt, ...);

#define QG_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      ::qgpu::throw_error(QGPU_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), \
                          __FILE__, __LINE__, cudaGetErrorString(_e));                        \
  } while (0)

// ---- device-resident table: a Vec<RecordBatch> concatenated column-wise in HBM --------------
struct Field {
  std::string name;
  uint8_t dtype;  // DT
  bool nullable;
};

struct Column {
  uint8_t dtype = DT_NULL;
  int64_t n = 0;
  void* data = nullptr;        // fixed width values / bit-packed bools / Utf8 int32 offsets (n+1)
  uint32_t* valid = nullptr;   // validity bitmap or nullptr
  uint8_t* chars = nullptr;    // Utf8 bytes
  int64_t n_chars = 0;
  int64_t null_count = 0;      // -1 = unknown
  bool owned = true;           // false: borrowed from another table (zero-copy pass-through)
  bool owned_valid = false;    // a borrowed column whose validity bitmap is nevertheless its own (the radix join's aliased key)
  // min / max of the valid values of an integer column, computed on first use (tables are immutable) - the join
  // picks its table layout from the build key's range
  mutable bool has_range = false;
  mutable int64_t vmin = 0, vmax = -1;
};

struct Ctx;

struct Table {
  Ctx* ctx = nullptr;  // owner of the device memory; nullptr: schema-only, or orphaned by qgpu_shutdown
  std::vector<Field> fields;
  std::vector<Column> cols;
  int64_t n_rows = 0;
  std::vector<int64_t> batch_offsets;  // size n_batches + 1
  std::vector<std::shared_ptr<Table>> keepalive;  // tables whose buffers we borrow
  // concurrency: a context runs callers on several lanes (streams).  `ready` is recorded on the producing lane's stream
  // when an operator hands the table out; an operator on ANOTHER lane makes its stream wait for it and marks the table
  // `shared`, which makes the destructor drain every lane before the memory goes back to the owner's block cache.
  cudaEvent_t ready = nullptr;
  bool shared = false;
  void adopt(Ctx* c);  // ties the table to its context (registered there so that qgpu_shutdown can orphan it)
  ~Table();
};

inline int dtype_width(uint8_t dt) {
  switch (dt) {
    case DT_I32: case DT_F32: return 4;
    case DT_I64: case DT_F64: return 8;
    default: return 0;
  }
}

// ---- compiled program -----------------------------------------------------------------------
struct ExprNode {
  int kind = 0, op = 0, lit_type = 0;
  bool is_null = false;
  uint64_t imm = 0;
  std::string str;
  int left = -1, right = -1;
  // inferred
  uint8_t type = DT_NULL;
  bool maybe_null = false;
};

struct ExprTree {
  std::vector<ExprNode> nodes;
  int root = -1;
  std::string canon(int n) const;  // canonical text, used to dedupe identical aggregate inputs
};

struct SchemaView {
  std::vector<uint8_t> dtypes;
  std::vector<bool> has_nulls;
};

struct AggOutput {
  int func;         // qgpu_agg_func
  uint8_t in_type;  // type of the aggregate's input expression
  uint8_t out_type; // per operators.rs:745-848
  int val_slot;     // -1 for COUNT
  int cnt_slot;     // non-null count slot
};

struct Program {
  std::vector<Insn> code;
  std::vector<int> col_of_slot;   // table column index for each program column slot
  int max_stack = 0;
  // aggregate programs
  std::vector<uint8_t> acc_kinds;  // AccKind per accumulator slot
  std::vector<AggOutput> agg_out;
  // projection programs
  std::vector<uint8_t> out_types;
  int slot_for(int table_col);
  std::string listing() const;
};

ExprTree parse_postfix(const qgpu_expr* e);
uint8_t infer_types(ExprTree& t, const SchemaView& s);  // throws Error with the reference's messages
void compile_predicate(ExprTree& t, const SchemaView& s, Program& p);
uint8_t compile_value(ExprTree& t, const SchemaView& s, Program& p);  // leaves the value in acc
void compile_aggregates(std::vector<ExprTree>& trees, const std::vector<int>& funcs, const SchemaView& s, Program& p);

// ---- context ----------------------------------------------------------------------------------
struct ProfileEntry {
  const char* name;
  cudaEvent_t start, stop;
  bool timed;
};

struct Ctx {
  qgpu_ctx* parent = nullptr;  // the context this lane belongs to
  int lane = 0;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;
  std::recursive_mutex mu;   // held by the one caller that runs on this lane
  std::mutex cache_mu;       // guards the block cache and the table registry: tables may be freed from any lane
  std::map<std::string, int64_t> options;
  // profiling
  bool profile_events = false;
  std::vector<ProfileEntry> prof;
  int64_t launch_count = 0;
  cudaEvent_t timer_start = nullptr, timer_stop = nullptr;
  // NCCL
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  // exchange arena: a cudaMalloc'ed buffer every GPU of the box can write into (CUDA IPC over NVLink)
  void* arena = nullptr;
  size_t arena_bytes = 0;
  void* arena_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // pinned staging for small D2H results
  void* pinned_scratch = nullptr;
  size_t pinned_scratch_bytes = 0;
  // exact-size block cache over the stream-ordered pool (see Ctx::alloc)
  std::map<size_t, std::vector<void*>> free_blocks;
  std::map<void*, size_t> live_blocks;
  size_t cached_bytes = 0;
  // tables whose device memory belongs to this context and are still alive (handles the caller has not released):
  // qgpu_shutdown frees their memory and orphans them, so a late qgpu_table_release never touches a dead context
  std::set<Table*> live_tables;
  // specialised scans: key of the last launched program, and per program the fraction of rows its predicate kept
  // resident tables by name: what MemoryDataSource is to the reference (query-storage/src/memory.rs:16-43)
  std::map<std::string, std::shared_ptr<Table>> catalog;
  std::string last_scan_key;
  std::map<std::string, float> scan_feedback;

  int64_t opt(const char* key, int64_t dflt) const {
    auto it = options.find(key);
    return it == options.end() ? dflt : it->second;
  }
  void* alloc(size_t bytes);            // cached block of the rounded size, else cudaMallocAsync
  void free(void* p);                   // back to the cache (all work is ordered on the one stream)
  void trim(size_t keep_bytes);         // return cached blocks to the driver until <= keep_bytes stay cached
  void trim_locked(size_t keep_bytes);  // (cache_mu held)
  void* zalloc(size_t bytes);           // alloc + memset 0
  void prof_begin(const char* name);
  void prof_end();
  void sync();
  void* scratch(size_t bytes);          // pinned host scratch
};

// grid of a persistent (grid-stride over tiles) kernel: SMs x the CTAs that are actually resident per SM for this
// kernel's register / shared-memory footprint, so the whole grid is one wave (a hard-coded "8 per SM" leaves a
// partial second wave that runs at a fraction of the occupancy)
int resident_ctas_per_sm(const void* kernel, int threads, size_t smem);
template <class K>
unsigned persistent_grid(Ctx* c, K kernel, int threads, int64_t tiles, size_t smem = 0) {
  int64_t g = (int64_t)c->sm_count * resident_ctas_per_sm((const void*)kernel, threads, smem);
  if (g > tiles) g = tiles;
  return (unsigned)(g < 1 ? 1 : g);
}

// host-side phase timer for "debug.host_timing": prints the host wall time spent since the previous mark (no
// device synchronisation: it shows where the HOST is busy or blocked while queueing work)
struct HostMarks {
  Ctx* c;
  bool on;
  double t0;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  explicit HostMarks(Ctx* c_) : c(c_), on(c_->opt("debug.host_timing", 0) != 0), t0(now()) {}
  void mark(const char* what) {
    if (!on) return;
    const double t = now();
    fprintf(stderr, "[qgpu host] %-28s %8.3f ms\n", what, t - t0);
    t0 = t;
  }
};

struct ProfScope {
  Ctx* c;
  ProfScope(Ctx* c_, const char* name) : c(c_) { c->prof_begin(name); }
  ~ProfScope() { c->prof_end(); }
};

// RAII device buffer from the context pool
struct DevBuf {
  Ctx* c = nullptr;
  void* p = nullptr;
  DevBuf() {}
  DevBuf(Ctx* c_, size_t bytes, bool zero = false) : c(c_) { p = zero ? c->zalloc(bytes) : c->alloc(bytes); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : c(o.c), p(o.p) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { reset(); c = o.c; p = o.p; o.p = nullptr; }
    return *this;
  }
  ~DevBuf() { reset(); }
  void reset() { if (p && c) c->free(p); p = nullptr; }
  void* release() { void* r = p; p = nullptr; return r; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline ColRef colref_of(const Column& col) {
  ColRef r;
  r.data = col.data;
  r.valid = col.valid;
  r.chars = col.chars;
  r.dtype = col.dtype;
  r.pad = 0;
  return r;
}

inline std::shared_ptr<Table> new_table(Ctx* c) {
  auto t = std::make_shared<Table>();
  t->adopt(c);
  t->batch_offsets.push_back(0);
  return t;
}

inline size_t pad_bytes(size_t b) { return ((b + 255) / 256) * 256 + 256; }  // slack lets vector loads over-read safely
inline size_t bitmap_bytes(int64_t n) { return pad_bytes((size_t)((n + 31) / 32) * 4); }

// ---- table.cu -----------------------------------------------------------------------------------
std::shared_ptr<Table> import_batches(Ctx* c, ArrowArray* batches, int64_t n, ArrowSchema* schema);
std::shared_ptr<Table> table_from_schema(ArrowSchema* schema);
void export_batch(Ctx* c, const Table& t, int64_t index, ArrowArray* out, ArrowSchema* out_schema);
Column alloc_column(Ctx* c, uint8_t dtype, int64_t n, bool with_valid);
void free_column(Ctx* c, Column& col);
SchemaView schema_view(const Table& t);


struct AggPlan {
  Program prog;
  int n_pred = 0;
  int pass_slot = -1;  // hidden COUNT of rows that passed the fused predicate
  std::vector<int> key_cols;
  std::vector<std::string> key_names;
};
AggPlan plan_aggregate(const Table& in, const qgpu_expr* pred, const qgpu_expr* groups, int n_group, const qgpu_agg_expr* aggs,
                       int n_aggr);

// ---- shared by the operator files (ops.cu, ops_groupby.cu) ------------------------------------------
struct ScanParams;
struct DeviceProgram {
  DevBuf code;
  DevBuf err;
};
int rows_per_thread(Ctx* c);
void bind_program(Ctx* c, const Table& in, const Program& prog, ScanParams& P, DeviceProgram& dp);
uint32_t read_u32(Ctx* c, const uint32_t* d);
void raise_device_error(uint32_t err);
void raise_device_error(Ctx* c, uint32_t err, const Table& in, const Program& prog, int n_pred);
void add_agg_outputs(Ctx* c, const Program& prog, int n_aggr, const uint64_t* acc, const uint32_t* slots, int64_t ng, Table& out,
                     const uint32_t* d_live = nullptr);
// High-cardinality GROUP BY on one Int64 key through two radix levels + shared-memory tables (ops_groupby.cu).
// false: not eligible, or a device-side bail-out (skewed partition, table full) - the caller takes the HBM-table path.
// true: *out holds the groups (nullptr when there are none); *err_out the device error word of the pass.
bool group_by_partitioned(Ctx* c, const Table& in, const AggPlan& ap, int n_aggr, std::shared_ptr<Table>* out, uint32_t* err_out);
// The multi-GPU form (collective): the rows are first exchanged by the top bits of the key hash - every rank scatters its
// rows straight into the owners' exchange arenas over NVLink - then every rank runs the same two radix levels +
// shared-memory aggregation on the rows it owns and returns ITS shard of the groups.  `total_rows` / `max_rows`: sum and
// maximum of the ranks' row counts.  Returns false on EVERY rank alike when the shape is not eligible or any rank bailed
// out (the caller then takes the older exchange path).
bool group_by_exchanged(Ctx* c, const Table& in, const AggPlan& ap, int n_aggr, int64_t total_rows, int64_t max_rows,
                        std::shared_ptr<Table>* out, uint32_t* err_out);

// ---- operators (ops.cu) --------------------------------------------------------------------------
std::shared_ptr<Table> op_filter(Ctx* c, const Table& in, const qgpu_expr* pred);
std::shared_ptr<Table> op_aggregate(Ctx* c, const Table& in, const qgpu_expr* pred, const qgpu_expr* groups, int n_group,
                                    const qgpu_agg_expr* aggs, int n_aggr);
std::shared_ptr<Table> op_join(Ctx* c, const Table& left, const Table& right, int join_type, const qgpu_expr* on);
// multi-GPU join whose radix scatter is the exchange; *done = false: not eligible here (on every rank alike) - the
// caller shuffles and joins locally instead
std::shared_ptr<Table> op_join_exchange(Ctx* c, const Table& left, const Table& right, int join_type, const qgpu_expr* on, bool* done);
std::shared_ptr<Table> op_sort(Ctx* c, const Table& in, const std::vector<int>& key_cols, const std::vector<int>& descending,
                               const std::vector<int>& nulls_first, int64_t fetch);
std::shared_ptr<Table> op_limit(Ctx* c, const Table& in, int64_t skip, int64_t fetch);
std::shared_ptr<Table> op_projection(Ctx* c, const std::shared_ptr<Table>& in, const qgpu_expr* exprs, const char* const* names,
                                     int n);
std::vector<std::shared_ptr<Table>> op_exchange_hash(Ctx* c, const Table& in, const std::vector<int>& key_cols,
                                                     int num_partitions, bool siphash);
std::shared_ptr<Table> op_shuffle_hash(Ctx* c, const Table& in, const std::vector<int>& key_cols);
std::shared_ptr<Table> op_allgather(Ctx* c, const Table& in);
std::shared_ptr<Table> gather_rows(Ctx* c, const Table& in, const std::vector<int>& cols, const uint32_t* d_idx, int64_t n_out,
                                   bool idx_may_be_null);

// ---- shuffle.cu: NCCL communicator (libnccl is dlopen'ed on first use) ---------------------------
void comm_unique_id(void* out_id128);
void comm_init(Ctx* c, const void* id128, int rank, int world);
void comm_destroy(Ctx* c);
void merge_scalar_across_ranks(Ctx* c, uint64_t* finals, int n_acc, const uint8_t* kinds);
// exchange arena + small collectives for the NVLink-scatter join (ops.cu)
void arena_ensure(Ctx* c, size_t bytes);   // collective: every rank ends up with >= bytes, mapped on every peer
void arena_release(Ctx* c, bool collective);
void stream_barrier(Ctx* c);               // stream-ordered barrier over the communicator (no host sync)
void allgather_device(Ctx* c, const void* d_in, void* d_out, size_t bytes);  // device buffers, stream-ordered
std::vector<int64_t> allgather_i64(Ctx* c, const std::vector<int64_t>& mine);
std::shared_ptr<Table> allgather_groups(Ctx* c, const Table& local_keys, const uint64_t* local_acc, int n_acc, uint32_t local_err,
                                        DevBuf& all_acc, uint32_t& any_err);

}  // namespace qgpu

// One context = one device, several LANES.  A lane is a complete execution resource - stream, block cache, pinned
// scratch, profile list - used by one caller at a time; concurrent callers (the reference's executor is called from
// pgwire's blocking pool, crates/query-pgwire/src/backend.rs:214-217, 569) run on different lanes side by side.  Lane 0
// also owns the NCCL communicator, the exchange arena and the table catalog: collectives always run there.
struct qgpu_ctx {
  qgpu::Ctx c;  // lane 0
  std::vector<std::unique_ptr<qgpu::Ctx>> more;  // lanes 1..
  std::atomic<int> next_lane{0};
  int n_lanes() const { return 1 + (int)more.size(); }
  qgpu::Ctx* lane(int i) { return i == 0 ? &c : more[(size_t)i - 1].get(); }
};
struct qgpu_table {
  std::shared_ptr<qgpu::Table> t;
};
