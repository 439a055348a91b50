/*
 * qgpu.h — C ABI of libqgpu, the B200-native (sm_100a) implementation of
 * query-engine's vectorised relational hot path.
 *
 * This header is the drop-in boundary.  The reference (AarambhDevHub/query-engine)
 * is pure Rust and has no FFI of its own; each entry point below names the
 * reference function whose BODY it replaces.  A Rust `query-gpu` crate binds
 * these symbols with a plain `extern "C"` block and the
 * three private arms of `QueryExecutor` call it:
 *
 *   qgpu_execute_filter      <- QueryExecutor::execute_filter     crates/query-executor/src/executor.rs:131-155
 *   qgpu_execute_aggregate   <- QueryExecutor::execute_aggregate  crates/query-executor/src/executor.rs:157-190
 *   qgpu_execute_join        <- QueryExecutor::execute_join       crates/query-executor/src/executor.rs:343-361
 *   qgpu_execute_projection  <- QueryExecutor::execute_projection crates/query-executor/src/executor.rs:93-129
 *   qgpu_execute_limit       <- QueryExecutor::execute_limit      crates/query-executor/src/executor.rs:299-341
 *   qgpu_sort                <- Merge::sorted_merge (after its concat) crates/query-distributed/src/operators.rs:141-194;
 *                               also the Tier-2 body of QueryExecutor::execute_sort, a pass-through in the reference
 *                               (executor.rs:290-297)
 *   qgpu_exchange_hash       <- Partitioner::partition_by_hash    crates/query-distributed/src/partition.rs:151-212
 *   qgpu_shuffle_hash        <- Exchange::execute between stages  crates/query-distributed/src/operators.rs:58-61
 *                               (+ executor.rs:230-263), as an in-box NCCL all-to-all
 *   qgpu_execute_join_exchange <- the "shuffle both sides, join locally" stage pair of the distributed planner
 *                               crates/query-distributed/src/planner.rs:228-249, fused: NVLink scatter + local join
 *   qgpu_expr / qgpu_agg_expr<- PhysicalExpr / AggregateExpr      crates/query-executor/src/physical_plan.rs:74-157
 *   qgpu_table_import/export <- Vec<RecordBatch> in / out of the operators (Arrow C Data Interface;
 *                               arrow-rs side: arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema})
 *
 * Conventions: plain pointers and sizes only; every function returns a
 * qgpu_status (0 = ok); on failure `qgpu_last_error()` returns the message the
 * reference would have put in QueryError::ExecutionError / ArrowError.  There
 * is no CPU fallback: without a usable CUDA device `qgpu_init` fails.
 */
#ifndef QGPU_H
#define QGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Arrow C Data Interface (standard, ABI-stable) ---------------------- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif

/* ---- status codes: which QueryError variant the Rust shim raises -------- */
typedef enum qgpu_status {
  QGPU_OK = 0,
  QGPU_ERR_EXECUTION = 1,   /* QueryError::ExecutionError(msg)   query-core/src/error.rs */
  QGPU_ERR_ARROW = 2,       /* QueryError::ArrowError(msg)  (checked arithmetic, type mismatch in cmp) */
  QGPU_ERR_CUDA = 3,        /* CUDA / NCCL runtime failure -> ExecutionError */
  QGPU_ERR_INVALID = 4,     /* bad argument to the C ABI itself */
  QGPU_ERR_PARTITION = 5,   /* DistributedError::PartitionError(msg) query-distributed/src/error.rs */
  QGPU_ERR_UNSUPPORTED = 6  /* shape not on the accelerated path: caller keeps the reference's CPU arm */
} qgpu_status;

/* ---- data types: ScalarValue / Arrow column types the path handles ------ */
typedef enum qgpu_dtype {
  QGPU_NULL = 0,
  QGPU_BOOL = 1,
  QGPU_INT32 = 2,
  QGPU_INT64 = 3,
  QGPU_FLOAT32 = 4,
  QGPU_FLOAT64 = 5,
  QGPU_UTF8 = 6,
  /* literal-only types (ScalarValue variants the operator ladders reject) */
  QGPU_INT8 = 7,
  QGPU_INT16 = 8,
  QGPU_UINT8 = 9,
  QGPU_UINT16 = 10,
  QGPU_UINT32 = 11,
  QGPU_UINT64 = 12
} qgpu_dtype;

/* BinaryOp, physical_plan.rs:119-136 (same order) */
typedef enum qgpu_binary_op {
  QGPU_ADD = 0, QGPU_SUBTRACT = 1, QGPU_MULTIPLY = 2, QGPU_DIVIDE = 3, QGPU_MODULO = 4,
  QGPU_EQUAL = 5, QGPU_NOT_EQUAL = 6, QGPU_LESS = 7, QGPU_LESS_EQUAL = 8,
  QGPU_GREATER = 9, QGPU_GREATER_EQUAL = 10, QGPU_AND = 11, QGPU_OR = 12, QGPU_TS_MATCH = 13
} qgpu_binary_op;

/* UnaryOp, physical_plan.rs:138-142 */
typedef enum qgpu_unary_op {
  QGPU_NOT = 0,
  QGPU_MINUS = 1,
  /* extension (not a reference UnaryOp): `x as f64` of a numeric value, NULL stays NULL.  Used by the multi-GPU final
   * aggregate to finish AVG = SUM(partial sums) as f64 / SUM(partial counts) as f64 (operators.rs:770-808) */
  QGPU_CAST_FLOAT64 = 2
} qgpu_unary_op;

/* AggregateFunction, physical_plan.rs:150-157 */
typedef enum qgpu_agg_func { QGPU_COUNT = 0, QGPU_SUM = 1, QGPU_AVG = 2, QGPU_MIN = 3, QGPU_MAX = 4 } qgpu_agg_func;

/* JoinType, query-parser/src/ast.rs:101-107 */
typedef enum qgpu_join_type { QGPU_INNER = 0, QGPU_LEFT = 1, QGPU_RIGHT = 2, QGPU_FULL = 3, QGPU_CROSS = 4 } qgpu_join_type;

/* ---- PhysicalExpr, flattened to postfix --------------------------------- */
typedef enum qgpu_expr_kind {
  QGPU_EXPR_COLUMN = 0,      /* PhysicalExpr::Column{name,index}:   op = index, str = name (may be NULL) */
  QGPU_EXPR_LITERAL = 1,     /* PhysicalExpr::Literal(ScalarValue): op = qgpu_dtype, is_null, value / str */
  QGPU_EXPR_BINARY = 2,      /* PhysicalExpr::BinaryExpr: op = qgpu_binary_op; pops right, then left */
  QGPU_EXPR_UNARY = 3,       /* PhysicalExpr::UnaryExpr:  op = qgpu_unary_op;  pops one */
  QGPU_EXPR_UNSUPPORTED = 4  /* ScalarSubquery/InSubquery/Exists/WindowFunction/ScalarFunction: op = variant 0..4;
                                evaluating it fails with the reference's message (operators.rs:34-60) */
} qgpu_expr_kind;

typedef struct qgpu_expr_node {
  int32_t kind;     /* qgpu_expr_kind */
  int32_t op;       /* see qgpu_expr_kind */
  int32_t is_null;  /* LITERAL: ScalarValue::X(None) or ScalarValue::Null */
  int32_t reserved;
  union {
    int64_t i64;
    uint64_t u64;
    double f64;
    float f32;
    int32_t b;
  } value;          /* LITERAL payload for fixed-width types */
  const char* str;  /* LITERAL Utf8 payload / COLUMN name, NUL-terminated, may be NULL */
} qgpu_expr_node;

typedef struct qgpu_expr {
  const qgpu_expr_node* nodes; /* postfix order */
  int32_t n_nodes;
} qgpu_expr;

/* AggregateExpr{func, expr}, physical_plan.rs:144-148 */
typedef struct qgpu_agg_expr {
  int32_t func; /* qgpu_agg_func */
  qgpu_expr expr;
} qgpu_agg_expr;

/* ---- handles ------------------------------------------------------------ */
typedef struct qgpu_ctx qgpu_ctx;     /* one CUDA device + stream + memory pool (+ NCCL communicator) */
typedef struct qgpu_table qgpu_table; /* a device-resident Vec<RecordBatch>: columns in HBM + batch row offsets */

/* ---- context ------------------------------------------------------------ */
int32_t qgpu_version(void);
const char* qgpu_last_error(void); /* thread-local; valid until the next call on this thread */
int32_t qgpu_init(int32_t device, qgpu_ctx** out);
void qgpu_shutdown(qgpu_ctx* ctx);
/* tunables; all optional, the full list with defaults is the "Runtime options" table of DESIGN.md ("scan.jit",
 * "group.capacity", "join.dense", ...).  Two are set by a multi-GPU host around its calls: "agg.merge_ranks" = 1 makes
 * qgpu_execute_aggregate a collective over the communicator's ranks - every rank returns the merged result, or, for a
 * GROUP BY with many groups, its key-partitioned shard of it (rows are exchanged by key over NVLink); "agg.keep_empty". */
int32_t qgpu_set_option(qgpu_ctx* ctx, const char* key, int64_t value);
int32_t qgpu_synchronize(qgpu_ctx* ctx);

/* pinned host memory for zero-copy ingest of Arrow buffers (north_star) */
int32_t qgpu_host_register(qgpu_ctx* ctx, void* ptr, size_t bytes);
int32_t qgpu_host_unregister(qgpu_ctx* ctx, void* ptr);

/* ---- Vec<RecordBatch> <-> device --------------------------------------- */
/* `batches` points at n_batches ArrowArray structs (each a struct array = one RecordBatch) sharing `schema`.
 * The buffers are read for the duration of the call (H2D copies); the structs are released before return. */
int32_t qgpu_table_import(qgpu_ctx* ctx, struct ArrowArray* batches, int64_t n_batches, struct ArrowSchema* schema,
                          qgpu_table** out);
/* A zero-batch table that carries only a schema (no device memory, no context needed): what an operator sees
 * when its child produced no batches; also lets the CPU-only test-suite drive qgpu_debug_compile.
 * The schema is released before return. */
int32_t qgpu_table_from_schema(struct ArrowSchema* schema, qgpu_table** out);
int64_t qgpu_table_num_batches(const qgpu_table* t);
int64_t qgpu_table_num_rows(const qgpu_table* t);
int32_t qgpu_table_num_columns(const qgpu_table* t);
/* export batch `index` into freshly allocated host buffers owned by the returned ArrowArray (release callback) */
int32_t qgpu_table_export(qgpu_ctx* ctx, const qgpu_table* t, int64_t index, struct ArrowArray* out_array,
                          struct ArrowSchema* out_schema);
void qgpu_table_release(qgpu_ctx* ctx, qgpu_table* t);

/* ---- resident tables ------------------------------- */
/* A catalog of device-resident tables keyed by name: the role MemoryDataSource plays in the reference
 * (crates/query-storage/src/memory.rs:16-43, 305-313) - a table is uploaded once and scanned by many queries without
 * crossing PCIe again.  The catalog shares ownership: the caller may release its own handle after `put`; `get`
 * returns a new handle (NULL when the name is unknown). */
int32_t qgpu_catalog_put(qgpu_ctx* ctx, const char* name, qgpu_table* t);
int32_t qgpu_catalog_get(qgpu_ctx* ctx, const char* name, qgpu_table** out);
int32_t qgpu_catalog_drop(qgpu_ctx* ctx, const char* name);

/* ---- operators (bodies of the reference's execute_* arms) --------------- */
int32_t qgpu_execute_filter(qgpu_ctx* ctx, const qgpu_table* input, const qgpu_expr* predicate, qgpu_table** out);

/* `predicate` may be NULL; when given, it is the Filter node directly below the HashAggregate, fused into the
 * same pass (rows failing it are not evaluated by the aggregate expressions, exactly as in the unfused plan).
 * `*out` is NULL when the reference returns `Ok(vec![])`. */
int32_t qgpu_execute_aggregate(qgpu_ctx* ctx, const qgpu_table* input, const qgpu_expr* predicate,
                               const qgpu_expr* group_exprs, int32_t n_group, const qgpu_agg_expr* aggr_exprs,
                               int32_t n_aggr, qgpu_table** out);

/* `on`: Column = Column (or an AND chain of those) over the merged left++right schema (planner.rs:148-157). */
int32_t qgpu_execute_join(qgpu_ctx* ctx, const qgpu_table* left, const qgpu_table* right, int32_t join_type,
                          const qgpu_expr* on, qgpu_table** out);

int32_t qgpu_execute_projection(qgpu_ctx* ctx, const qgpu_table* input, const qgpu_expr* exprs,
                                const char* const* names, int32_t n_exprs, qgpu_table** out);

/* LIMIT / OFFSET: skip `skip` rows, then keep at most `fetch` rows (fetch < 0: Option::None, no bound).  The output
 * batches are the slices of the input batches the kept range covers; `*out` is NULL when nothing is kept. */
int32_t qgpu_execute_limit(qgpu_ctx* ctx, const qgpu_table* input, int64_t skip, int64_t fetch, qgpu_table** out);

/* Lexicographic sort by the named columns (SortColumn{name, ascending, nulls_first}, query-distributed/src/operators.rs:
 * 97-104; names the schema lacks are skipped like the reference's filter_map), every column taken in that order, ONE
 * output batch; floats order by IEEE totalOrder like arrow.  fetch >= 0 keeps the first `fetch` rows (Top-K). */
int32_t qgpu_sort(qgpu_ctx* ctx, const qgpu_table* input, const char* const* sort_columns, const int32_t* ascending,
                  const int32_t* nulls_first, int32_t n_columns, int64_t fetch, qgpu_table** out);

/* ---- exchange ----------------------------------------------------------- */
/* Local hash partitioning: out_partitions[p] receives the rows whose SipHash-1-3(key columns) % num_partitions == p,
 * in input order (NULL for an empty partition). */
int32_t qgpu_exchange_hash(qgpu_ctx* ctx, const qgpu_table* input, const char* const* key_columns, int32_t n_keys,
                           int32_t num_partitions, qgpu_table** out_partitions);

/* In-box multi-GPU exchange: one process per GPU; rank 0 creates the id, the host distributes the 128 bytes. */
#define QGPU_UNIQUE_ID_BYTES 128
int32_t qgpu_comm_unique_id(void* out_id128);
int32_t qgpu_comm_init(qgpu_ctx* ctx, const void* id128, int32_t rank, int32_t world_size);
int32_t qgpu_comm_rank(const qgpu_ctx* ctx);
int32_t qgpu_comm_size(const qgpu_ctx* ctx);
/* hash-partition `input` by its key columns into world_size parts, all-to-all them over NCCL, return the rows this
 * rank owns (concatenated in source-rank order).  *out may be NULL if nothing arrived. */
int32_t qgpu_shuffle_hash(qgpu_ctx* ctx, const qgpu_table* input, const char* const* key_columns, int32_t n_keys,
                          qgpu_table** out);
/* Multi-GPU join of row-sharded inputs in ONE call: the radix scatter of the join writes every row straight into the
 * owning GPU's memory over NVLink (CUDA-IPC exchange arena, created on first use) instead of a separate shuffle, then
 * every rank joins the regions it owns; the output stays sharded, like shuffle + local join
 * (query-distributed/src/planner.rs:228-249).  Collective: every rank calls it with its shard.  *done = 0 means the
 * shape is not eligible (decided alike on every rank) and nothing was produced: the caller then uses
 * qgpu_shuffle_hash + qgpu_execute_join. */
int32_t qgpu_execute_join_exchange(qgpu_ctx* ctx, const qgpu_table* left, const qgpu_table* right, int32_t join_type,
                                   const qgpu_expr* on, qgpu_table** out, int32_t* done);
/* gather every rank's table on every rank (final merge of partial aggregates); concatenated in rank order */
int32_t qgpu_allgather_table(qgpu_ctx* ctx, const qgpu_table* input, qgpu_table** out);

/* ---- measurement -------------------------------------------------------- */
/* When enabled, every kernel launch is bracketed by CUDA events on the context's stream. */
int32_t qgpu_profile_enable(qgpu_ctx* ctx, int32_t on);
int32_t qgpu_profile_reset(qgpu_ctx* ctx);
int64_t qgpu_profile_count(qgpu_ctx* ctx); /* kernels launched since reset (counted even when events are off) */
/* i-th launch since reset: name and duration (ms; 0 when events are off) */
int32_t qgpu_profile_get(qgpu_ctx* ctx, int64_t i, const char** name, float* ms);
/* events on the context's stream for timing a whole operator call */
int32_t qgpu_timer_start(qgpu_ctx* ctx);
int32_t qgpu_timer_stop(qgpu_ctx* ctx, float* ms);

/* ---- debugging ---------------------------------------------------------- */
/* Type-check + compile an expression against a table's schema without running it; writes a listing of the
 * device program into buf.  Needs no GPU work (used by the CPU test suite). */
int32_t qgpu_debug_compile(qgpu_ctx* ctx, const qgpu_table* schema_of, const qgpu_expr* expr, char* buf, size_t buf_len);

/* Compile (Filter +) HashAggregate over `schema_of`'s schema into a device program, specialise the scan kernel on it
 * with NVRTC for sm_100a and write the NVRTC / ptxas log into buf.  Nothing is loaded or launched, so this works
 * without a GPU; with QGPU_JIT_DUMP=<dir> set the generated source and cubin are written there. */
int32_t qgpu_debug_jit_aggregate(const qgpu_table* schema_of, const qgpu_expr* predicate, const qgpu_expr* group_exprs,
                                 int32_t n_group, const qgpu_agg_expr* aggr_exprs, int32_t n_aggr, int32_t rows_per_thread,
                                 char* buf, size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif /* QGPU_H */
