// capi.cu — the extern "C" surface declared in include/qgpu.h.
//
// Every entry point converts C++ exceptions into a qgpu_status + thread-local message; the message is the text
// the reference would have carried in QueryError::ExecutionError / ArrowError (crates/query-core/src/error.rs:4-50).
// A context runs concurrent callers side by side (the reference's QueryExecutor is called concurrently from pgwire's
// blocking pool, crates/query-pgwire/src/backend.rs:214-217,569): every caller thread gets a LANE - stream, block cache,
// pinned scratch - and only callers that share a lane serialise (qgpu_internal.h: qgpu_ctx).
#include <cstdlib>
#include <cstring>

#include "jit.h"
#include "qgpu_internal.h"
#include "scan_kernels.cuh"

namespace qgpu {
thread_local std::string g_last_error;

static int32_t fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
}  // namespace qgpu

using namespace qgpu;

#define QG_TRY try {
#define QG_CATCH                                                             \
  }                                                                          \
  catch (const qgpu::Error& e) { return qgpu::fail(e.code, e.msg); }         \
  catch (const std::bad_alloc&) { return qgpu::fail(QGPU_ERR_CUDA, "host allocation failed"); } \
  catch (const std::exception& e) { return qgpu::fail(QGPU_ERR_EXECUTION, e.what()); }

#define QG_NEED(ptr, what)                                                   \
  if (!(ptr)) return qgpu::fail(QGPU_ERR_INVALID, std::string("null argument: ") + what)

namespace qgpu {
// every lane's stream idle (a table shared between lanes is about to be freed)
void drain_all_lanes(Ctx* c) {
  if (!c->parent) return;
  for (int i = 0; i < c->parent->n_lanes(); ++i) cudaStreamSynchronize(c->parent->lane(i)->stream);
}
}  // namespace qgpu

extern "C" {

int32_t qgpu_version(void) { return 100; }

const char* qgpu_last_error(void) { return g_last_error.c_str(); }

int32_t qgpu_init(int32_t device, qgpu_ctx** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_TRY
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    throw_error(QGPU_ERR_CUDA, "no usable CUDA device (%s); libqgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= n_dev) throw_error(QGPU_ERR_INVALID, "device %d out of range (have %d)", device, n_dev);
  QG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  QG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    throw_error(QGPU_ERR_CUDA, "device %d is sm_%d%d; libqgpu is built for sm_100a only", device, prop.major, prop.minor);
  auto* h = new qgpu_ctx();
  // lanes: concurrent callers run side by side, each on a stream / block cache of its own (QGPU_LANES, default 4)
  int n_lanes = 4;
  if (const char* e = getenv("QGPU_LANES")) n_lanes = std::max(1, std::min(16, atoi(e)));
  for (int i = 1; i < n_lanes; ++i) h->more.emplace_back(new Ctx());
  try {
    for (int i = 0; i < n_lanes; ++i) {
      Ctx& c = *h->lane(i);
      c.parent = h;
      c.lane = i;
      c.device = device;
      c.sm_count = prop.multiProcessorCount;
      QG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
      QG_CUDA(cudaDeviceGetDefaultMemPool(&c.pool, device));
      // keep freed blocks cached in the pool: operators allocate and free multi-GB temporaries per call
      uint64_t thresh = UINT64_MAX;
      QG_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &thresh));
      QG_CUDA(cudaEventCreate(&c.timer_start));
      QG_CUDA(cudaEventCreate(&c.timer_stop));
    }
  } catch (...) {
    qgpu_shutdown(h);
    throw;
  }
  *out = h;
  return QGPU_OK;
  QG_CATCH
}

void qgpu_shutdown(qgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->c.device);
  for (int i = 0; i < ctx->n_lanes(); ++i)
    if (ctx->lane(i)->stream) cudaStreamSynchronize(ctx->lane(i)->stream);
  ctx->c.catalog.clear();
  for (int i = ctx->n_lanes() - 1; i >= 0; --i) {
    Ctx& c = *ctx->lane(i);
    {
      // handles the caller still holds (or a garbage collector has not reached yet): free their device memory now and
      // orphan them, so that a later qgpu_table_release only deletes host structures
      std::vector<Table*> live;
      {
        std::lock_guard<std::mutex> lk(c.cache_mu);
        live.assign(c.live_tables.begin(), c.live_tables.end());
        c.live_tables.clear();
      }
      for (Table* t : live) {
        for (auto& col : t->cols) free_column(&c, col);
        t->ctx = nullptr;
      }
    }
    c.trim(0);
    if (i == 0) comm_destroy(&c);
    for (auto& e : c.prof) {
      cudaEventDestroy(e.start);
      cudaEventDestroy(e.stop);
    }
    if (c.timer_start) cudaEventDestroy(c.timer_start);
    if (c.timer_stop) cudaEventDestroy(c.timer_stop);
    if (c.pinned_scratch) cudaFreeHost(c.pinned_scratch);
    if (c.stream) cudaStreamDestroy(c.stream);
  }
  delete ctx;
}

// The lane of the calling thread: sticky (a thread's import -> operators -> export run in one stream, in order), handed
// out round-robin on a thread's first call into a context.  With a communicator (one process per GPU, multi-GPU runs)
// everything runs on lane 0, where the collectives live.
static Ctx* pick_lane(qgpu_ctx* ctx) {
  if (ctx->n_lanes() == 1 || ctx->c.world > 1) return &ctx->c;
  thread_local std::map<qgpu_ctx*, int> mine;
  auto it = mine.find(ctx);
  if (it == mine.end()) it = mine.emplace(ctx, ctx->next_lane.fetch_add(1) % ctx->n_lanes()).first;
  return ctx->lane(it->second % ctx->n_lanes());
}

// an input table produced on another lane: this lane's stream waits for its producer, and the table remembers that it
// is shared (its destructor then drains every lane before the memory is reused)
static void use_table(Ctx* c, const qgpu_table* t) {
  if (!t || !t->t) return;
  Table* tab = t->t.get();
  if (tab->ctx == nullptr || tab->ctx == c) return;
  tab->shared = true;
  if (tab->ready) cudaStreamWaitEvent(c->stream, tab->ready, 0);
  for (auto& k : tab->keepalive)
    if (k) k->shared = true;
}

// the caller's lane, locked for the duration of the call
#define QG_ENTER(ctx)                                         \
  QG_NEED(ctx, "ctx");                                        \
  Ctx* c = pick_lane(ctx);                                    \
  std::lock_guard<std::recursive_mutex> _lk(c->mu);           \
  QG_CUDA(cudaSetDevice(c->device))
// lane 0 (communicator, exchange arena, catalog, the bench's profile / timer calls)
#define QG_ENTER0(ctx)                                        \
  QG_NEED(ctx, "ctx");                                        \
  Ctx* c = &(ctx)->c;                                         \
  std::lock_guard<std::recursive_mutex> _lk(c->mu);           \
  QG_CUDA(cudaSetDevice(c->device))

int32_t qgpu_set_option(qgpu_ctx* ctx, const char* key, int64_t value) {
  QG_NEED(key, "key");
  QG_TRY
  QG_NEED(ctx, "ctx");
  for (int i = 0; i < ctx->n_lanes(); ++i) {
    Ctx* c = ctx->lane(i);
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    c->options[key] = value;
  }
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_synchronize(qgpu_ctx* ctx) {
  QG_TRY
  QG_NEED(ctx, "ctx");
  QG_CUDA(cudaSetDevice(ctx->c.device));
  for (int i = 0; i < ctx->n_lanes(); ++i) QG_CUDA(cudaStreamSynchronize(ctx->lane(i)->stream));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_host_register(qgpu_ctx* ctx, void* ptr, size_t bytes) {
  QG_NEED(ptr, "ptr");
  QG_TRY
  QG_ENTER(ctx);
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return QGPU_OK;
  }
  QG_CUDA(e);
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_host_unregister(qgpu_ctx* ctx, void* ptr) {
  QG_NEED(ptr, "ptr");
  QG_TRY
  QG_ENTER(ctx);
  cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    cudaGetLastError();
    return QGPU_OK;
  }
  QG_CUDA(e);
  return QGPU_OK;
  QG_CATCH
}

// ---- tables -------------------------------------------------------------------------------------
static qgpu_table* wrap(std::shared_ptr<Table> t) {
  if (!t) return nullptr;
  // "ready": everything queued on the producing lane so far (an operator may queue kernels after its last host sync)
  if (t->ctx && t->ctx->parent && t->ctx->parent->n_lanes() > 1 && t->ctx->stream) {
    if (!t->ready) cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming);
    if (t->ready) cudaEventRecord(t->ready, t->ctx->stream);
  }
  auto* h = new qgpu_table();
  h->t = std::move(t);
  return h;
}

int32_t qgpu_table_import(qgpu_ctx* ctx, struct ArrowArray* batches, int64_t n_batches, struct ArrowSchema* schema,
                          qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(schema, "schema");
  if (n_batches < 0 || (n_batches > 0 && !batches)) return fail(QGPU_ERR_INVALID, "bad batch array");
  // the inputs are consumed whatever happens: released on success and on every error path
  struct Releaser {
    ArrowArray* b;
    int64_t n;
    ArrowSchema* s;
    ~Releaser() {
      for (int64_t i = 0; i < n; ++i)
        if (b[i].release) b[i].release(&b[i]);
      if (s->release) s->release(s);
    }
  } releaser{batches, n_batches, schema};
  QG_TRY
  QG_ENTER(ctx);
  *out = wrap(import_batches(c, batches, n_batches, schema));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_table_from_schema(struct ArrowSchema* schema, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(schema, "schema");
  QG_TRY
  *out = wrap(table_from_schema(schema));
  return QGPU_OK;
  QG_CATCH
}

int64_t qgpu_table_num_batches(const qgpu_table* t) { return t ? (int64_t)t->t->batch_offsets.size() - 1 : 0; }
int64_t qgpu_table_num_rows(const qgpu_table* t) { return t ? t->t->n_rows : 0; }
int32_t qgpu_table_num_columns(const qgpu_table* t) { return t ? (int32_t)t->t->cols.size() : 0; }

int32_t qgpu_table_export(qgpu_ctx* ctx, const qgpu_table* t, int64_t index, struct ArrowArray* out_array,
                          struct ArrowSchema* out_schema) {
  QG_NEED(t, "table");
  QG_NEED(out_array, "out_array");
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, t);
  export_batch(c, *t->t, index, out_array, out_schema);
  return QGPU_OK;
  QG_CATCH
}

void qgpu_table_release(qgpu_ctx* ctx, qgpu_table* t) {
  if (!t) return;
  // The table knows its owning lane (nullptr once that context was shut down: only host memory is left then); `ctx` may
  // be NULL.  The owner's block cache and table registry are thread-safe, so no lane lock is taken: memory freed here is
  // reused in the owner's stream order, and a table that another lane has read drains every lane first (~Table).
  (void)ctx;
  Ctx* owner = t->t ? t->t->ctx : nullptr;
  if (!owner && t->t)
    for (auto& k : t->t->keepalive)
      if (k && k->ctx) owner = k->ctx;
  if (owner) cudaSetDevice(owner->device);
  delete t;
}

// ---- resident tables ------------------------------------------------------------------------------
int32_t qgpu_catalog_put(qgpu_ctx* ctx, const char* name, qgpu_table* t) {
  QG_NEED(name, "name");
  QG_NEED(t, "table");
  QG_TRY
  QG_ENTER0(ctx);
  t->t->shared = true;  // (scanned from every lane: drained before its memory is reused)
  c->catalog[name] = t->t;  // shared: the caller's handle stays valid and may be released independently
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_catalog_get(qgpu_ctx* ctx, const char* name, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(name, "name");
  QG_TRY
  QG_ENTER0(ctx);
  auto it = c->catalog.find(name);
  if (it != c->catalog.end()) *out = wrap(it->second);
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_catalog_drop(qgpu_ctx* ctx, const char* name) {
  QG_NEED(name, "name");
  QG_TRY
  QG_ENTER0(ctx);
  c->catalog.erase(name);
  return QGPU_OK;
  QG_CATCH
}

// ---- operators ----------------------------------------------------------------------------------
int32_t qgpu_execute_filter(qgpu_ctx* ctx, const qgpu_table* input, const qgpu_expr* predicate, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(input, "input");
  QG_NEED(predicate, "predicate");
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, input);
  *out = wrap(op_filter(c, *input->t, predicate));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_execute_aggregate(qgpu_ctx* ctx, const qgpu_table* input, const qgpu_expr* predicate,
                               const qgpu_expr* group_exprs, int32_t n_group, const qgpu_agg_expr* aggr_exprs,
                               int32_t n_aggr, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(input, "input");
  if (n_group < 0 || n_aggr < 0 || (n_group > 0 && !group_exprs) || (n_aggr > 0 && !aggr_exprs))
    return fail(QGPU_ERR_INVALID, "bad group / aggregate expression arrays");
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, input);
  *out = wrap(op_aggregate(c, *input->t, predicate, group_exprs, n_group, aggr_exprs, n_aggr));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_execute_join(qgpu_ctx* ctx, const qgpu_table* left, const qgpu_table* right, int32_t join_type,
                          const qgpu_expr* on, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(left, "left");
  QG_NEED(right, "right");
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, left);
  use_table(c, right);
  *out = wrap(op_join(c, *left->t, *right->t, join_type, on));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_execute_projection(qgpu_ctx* ctx, const qgpu_table* input, const qgpu_expr* exprs, const char* const* names,
                                int32_t n_exprs, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(input, "input");
  if (n_exprs < 0 || (n_exprs > 0 && !exprs)) return fail(QGPU_ERR_INVALID, "bad projection expression array");
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, input);
  *out = wrap(op_projection(c, input->t, exprs, names, n_exprs));
  return QGPU_OK;
  QG_CATCH
}

static std::vector<int> key_indices(const Table& t, const char* const* names, int n);

int32_t qgpu_execute_limit(qgpu_ctx* ctx, const qgpu_table* input, int64_t skip, int64_t fetch, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(input, "input");
  if (skip < 0) return fail(QGPU_ERR_INVALID, "negative skip");
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, input);
  *out = wrap(op_limit(c, *input->t, skip, fetch));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_sort(qgpu_ctx* ctx, const qgpu_table* input, const char* const* sort_columns, const int32_t* ascending,
                  const int32_t* nulls_first, int32_t n_columns, int64_t fetch, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(input, "input");
  if (n_columns < 0 || (n_columns > 0 && (!sort_columns || !ascending || !nulls_first))) return fail(QGPU_ERR_INVALID, "bad sort column arrays");
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, input);
  // columns are looked up by name; names the schema does not have are skipped (operators.rs:163-170 filter_map)
  std::vector<int> cols, desc, nf;
  for (int i = 0; i < n_columns; ++i) {
    const std::vector<int> idx = key_indices(*input->t, sort_columns + i, 1);
    if (idx.empty()) continue;
    cols.push_back(idx[0]);
    desc.push_back(ascending[i] ? 0 : 1);
    nf.push_back(nulls_first[i] ? 1 : 0);
  }
  *out = wrap(op_sort(c, *input->t, cols, desc, nf, fetch));
  return QGPU_OK;
  QG_CATCH
}

// ---- exchange -----------------------------------------------------------------------------------
// key columns are looked up by name, and names that are absent are skipped, exactly like
// `batch.schema().index_of(col_name)` + filter_map in partition.rs:165-171
static std::vector<int> key_indices(const Table& t, const char* const* names, int n) {
  std::vector<int> idx;
  for (int i = 0; i < n; ++i) {
    if (!names[i]) continue;
    for (size_t ci = 0; ci < t.fields.size(); ++ci)
      if (t.fields[ci].name == names[i]) {
        idx.push_back((int)ci);
        break;
      }
  }
  return idx;
}

int32_t qgpu_exchange_hash(qgpu_ctx* ctx, const qgpu_table* input, const char* const* key_columns, int32_t n_keys,
                           int32_t num_partitions, qgpu_table** out_partitions) {
  QG_NEED(input, "input");
  QG_NEED(out_partitions, "out_partitions");
  if (n_keys < 0 || (n_keys > 0 && !key_columns)) return fail(QGPU_ERR_INVALID, "bad key column array");
  for (int p = 0; p < num_partitions; ++p) out_partitions[p] = nullptr;
  QG_TRY
  QG_ENTER(ctx);
  use_table(c, input);
  auto parts = op_exchange_hash(c, *input->t, key_indices(*input->t, key_columns, n_keys), num_partitions, true);
  for (int p = 0; p < num_partitions; ++p) out_partitions[p] = wrap(parts[(size_t)p]);
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_comm_unique_id(void* out_id128) {
  QG_NEED(out_id128, "out_id128");
  QG_TRY
  comm_unique_id(out_id128);
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_comm_init(qgpu_ctx* ctx, const void* id128, int32_t rank, int32_t world_size) {
  QG_NEED(id128, "id128");
  QG_TRY
  QG_ENTER0(ctx);
  comm_init(c, id128, rank, world_size);
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_comm_rank(const qgpu_ctx* ctx) { return ctx ? ctx->c.rank : 0; }
int32_t qgpu_comm_size(const qgpu_ctx* ctx) { return ctx ? ctx->c.world : 1; }

int32_t qgpu_shuffle_hash(qgpu_ctx* ctx, const qgpu_table* input, const char* const* key_columns, int32_t n_keys,
                          qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(input, "input");
  if (n_keys < 0 || (n_keys > 0 && !key_columns)) return fail(QGPU_ERR_INVALID, "bad key column array");
  QG_TRY
  QG_ENTER0(ctx);
  use_table(c, input);
  *out = wrap(op_shuffle_hash(c, *input->t, key_indices(*input->t, key_columns, n_keys)));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_execute_join_exchange(qgpu_ctx* ctx, const qgpu_table* left, const qgpu_table* right, int32_t join_type,
                                   const qgpu_expr* on, qgpu_table** out, int32_t* done) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(done, "done");
  *done = 0;
  QG_NEED(left, "left");
  QG_NEED(right, "right");
  QG_TRY
  QG_ENTER0(ctx);
  use_table(c, left);
  use_table(c, right);
  bool d = false;
  *out = wrap(op_join_exchange(c, *left->t, *right->t, join_type, on, &d));
  *done = d ? 1 : 0;
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_allgather_table(qgpu_ctx* ctx, const qgpu_table* input, qgpu_table** out) {
  QG_NEED(out, "out");
  *out = nullptr;
  QG_NEED(input, "input");
  QG_TRY
  QG_ENTER0(ctx);
  use_table(c, input);
  *out = wrap(op_allgather(c, *input->t));
  return QGPU_OK;
  QG_CATCH
}

// ---- measurement --------------------------------------------------------------------------------
int32_t qgpu_profile_enable(qgpu_ctx* ctx, int32_t on) {
  QG_TRY
  QG_ENTER(ctx);
  c->profile_events = on != 0;
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_profile_reset(qgpu_ctx* ctx) {
  QG_TRY
  QG_ENTER(ctx);
  c->sync();
  for (auto& e : c->prof) {
    cudaEventDestroy(e.start);
    cudaEventDestroy(e.stop);
  }
  c->prof.clear();
  c->launch_count = 0;
  return QGPU_OK;
  QG_CATCH
}

int64_t qgpu_profile_count(qgpu_ctx* ctx) {
  if (!ctx) return 0;
  Ctx* c = pick_lane(ctx);
  std::lock_guard<std::recursive_mutex> lk(c->mu);
  return c->launch_count;
}

int32_t qgpu_profile_get(qgpu_ctx* ctx, int64_t i, const char** name, float* ms) {
  QG_TRY
  QG_ENTER(ctx);
  if (i < 0 || i >= (int64_t)c->prof.size()) throw_error(QGPU_ERR_INVALID, "profile entry %lld out of range", (long long)i);
  ProfileEntry& e = c->prof[(size_t)i];
  if (name) *name = e.name;
  if (ms) {
    QG_CUDA(cudaEventSynchronize(e.stop));
    QG_CUDA(cudaEventElapsedTime(ms, e.start, e.stop));
  }
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_timer_start(qgpu_ctx* ctx) {
  QG_TRY
  QG_ENTER(ctx);
  QG_CUDA(cudaEventRecord(c->timer_start, c->stream));
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_timer_stop(qgpu_ctx* ctx, float* ms) {
  QG_NEED(ms, "ms");
  QG_TRY
  QG_ENTER(ctx);
  QG_CUDA(cudaEventRecord(c->timer_stop, c->stream));
  QG_CUDA(cudaEventSynchronize(c->timer_stop));
  QG_CUDA(cudaEventElapsedTime(ms, c->timer_start, c->timer_stop));
  return QGPU_OK;
  QG_CATCH
}

// ---- debugging ----------------------------------------------------------------------------------
int32_t qgpu_debug_compile(qgpu_ctx* ctx, const qgpu_table* schema_of, const qgpu_expr* expr, char* buf, size_t buf_len) {
  (void)ctx;
  QG_NEED(schema_of, "schema_of");
  QG_NEED(expr, "expr");
  QG_TRY
  ExprTree t = parse_postfix(expr);
  const SchemaView sv = schema_view(*schema_of->t);
  Program p;
  ExprTree probe = t;
  const uint8_t ty = infer_types(probe, sv);
  if (ty == DT_BOOL) compile_predicate(t, sv, p);
  else compile_value(t, sv, p);
  const std::string s = p.listing();
  if (buf && buf_len > 0) {
    const size_t n = s.size() < buf_len - 1 ? s.size() : buf_len - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return QGPU_OK;
  QG_CATCH
}

int32_t qgpu_debug_jit_aggregate(const qgpu_table* schema_of, const qgpu_expr* predicate, const qgpu_expr* group_exprs,
                                 int32_t n_group, const qgpu_agg_expr* aggr_exprs, int32_t n_aggr, int32_t rows_per_thread,
                                 char* buf, size_t buf_len) {
  QG_NEED(schema_of, "schema_of");
  QG_TRY
  const Table& t = *schema_of->t;
  AggPlan ap = plan_aggregate(t, predicate, group_exprs, n_group, aggr_exprs, n_aggr);
  Program prog = ap.prog;
  ScanParams P;
  memset(&P, 0, sizeof P);
  P.n_pred = ap.n_pred;
  P.n_keys = n_group;
  for (int g = 0; g < n_group; ++g) P.key_slots[g] = prog.slot_for(ap.key_cols[(size_t)g]);
  const SchemaView sv = schema_view(t);
  for (size_t i = 0; i < prog.col_of_slot.size(); ++i) {
    P.cols[i].dtype = sv.dtypes[(size_t)prog.col_of_slot[i]];
    P.cols[i].valid = sv.has_nulls[(size_t)prog.col_of_slot[i]] ? (const uint32_t*)8 : nullptr;  // presence flag only
  }
  int sink = n_group == 0 ? SK_SCALAR : SK_GROUP_SMEM;
  if (const char* e = getenv("QGPU_DEBUG_SINK")) {  // developer aid: look at another sink's code without a GPU (QGPU_JIT_DUMP keeps the cubin)
    sink = atoi(e);
    P.part_slots = 1024;
    P.gp_preload = getenv("QGPU_DEBUG_GP_NOPRELOAD") ? 0u : 1u;
  }
  // (the partition pass is launched with four CTAs per SM, scan_kernels.cu::launch_specialised)
  const std::string src = jit_scan_source(sink, rows_per_thread == 8 ? 8 : 4, P, prog, false, sink == SK_GROUP_PART ? 4 : 2);
  std::string log;
  jit_compile(src, &log);
  if (buf && buf_len > 0) {
    const size_t n = log.size() < buf_len - 1 ? log.size() : buf_len - 1;
    memcpy(buf, log.data(), n);
    buf[n] = 0;
  }
  return QGPU_OK;
  QG_CATCH
}

}  // extern "C"
