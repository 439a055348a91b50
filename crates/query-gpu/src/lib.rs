//! query-gpu — safe wrappers over libqgpu, the B200-native implementation of query-engine's vectorised relational hot
//! path (filter + compaction, hash GROUP BY, hash join, partition exchange).
//!
//! The crate knows nothing of `query_executor::PhysicalExpr`: `query-executor` depends on it (not the other way round)
//! and flattens its expression trees into [`ExprBuf`] postfix programs (see `crates/patches/query-executor.patch`,
//! new file `src/gpu.rs`).  RecordBatches cross the boundary through the Arrow C Data Interface
//! (`arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema}`): buffers are read for the duration of the call - register them
//! with [`GpuContext::host_register`] for zero-copy DMA from pinned memory - and results come back in pinned host
//! buffers owned by the returned arrays.
//!
//! Error mapping (`query_core::QueryError`, crates/query-core/src/error.rs:4-50): status 1 -> `ExecutionError(msg)`;
//! 2 -> `ExecutionError(msg)` carrying arrow-rs' own text (the reference gets there through `#[from] ArrowError`);
//! 3 -> `ExecutionError` (CUDA / NCCL failure); 5 -> the distributed crate maps it to
//! `DistributedError::PartitionError`; 6 (`Unsupported`) is NOT an error for the caller: it keeps the reference's CPU
//! arm for that shape (string functions, CROSS joins, joins without an equi-`on`).

pub mod ffi;

use arrow::array::{Array, StructArray};
use arrow::datatypes::{Schema, SchemaRef};
use arrow::ffi::{from_ffi, to_ffi, FFI_ArrowArray, FFI_ArrowSchema};
use arrow::record_batch::RecordBatch;
use query_core::{QueryError, Result};
use std::ffi::{CStr, CString};
use std::os::raw::c_char;
use std::ptr;
use std::sync::Arc;

/// What a non-zero status means to the caller.
#[derive(Debug)]
pub enum GpuError {
    /// The shape is not on the accelerated path: run the reference's CPU arm instead.
    Unsupported(String),
    /// A real failure, already phrased like the reference would have.
    Query(QueryError),
    /// `DistributedError::PartitionError` text (status 5).
    Partition(String),
}

impl From<GpuError> for QueryError {
    fn from(e: GpuError) -> Self {
        match e {
            GpuError::Unsupported(m) | GpuError::Partition(m) => QueryError::ExecutionError(m),
            GpuError::Query(q) => q,
        }
    }
}

pub type GpuResult<T> = std::result::Result<T, GpuError>;

fn last_error() -> String {
    unsafe {
        let p = ffi::qgpu_last_error();
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

fn check(status: i32) -> GpuResult<()> {
    match status {
        ffi::QGPU_OK => Ok(()),
        ffi::QGPU_ERR_UNSUPPORTED => Err(GpuError::Unsupported(last_error())),
        ffi::QGPU_ERR_PARTITION => Err(GpuError::Partition(last_error())),
        _ => Err(GpuError::Query(QueryError::ExecutionError(last_error()))),
    }
}

/// One CUDA device + stream + memory pool (+ NCCL communicator).  Thread-safe: libqgpu serialises the callers of a
/// context internally, which is how pgwire uses the executor (`spawn_blocking`, backend.rs:214-217).
pub struct GpuContext {
    raw: *mut ffi::qgpu_ctx,
}

unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}

impl GpuContext {
    /// Fails when there is no usable sm_100 device: libqgpu has no CPU fallback.
    pub fn new(device: i32) -> Result<Arc<Self>> {
        let mut raw = ptr::null_mut();
        check(unsafe { ffi::qgpu_init(device, &mut raw) }).map_err(QueryError::from)?;
        Ok(Arc::new(Self { raw }))
    }

    pub fn raw(&self) -> *mut ffi::qgpu_ctx {
        self.raw
    }

    pub fn set_option(&self, key: &str, value: i64) -> GpuResult<()> {
        let k = CString::new(key).unwrap();
        check(unsafe { ffi::qgpu_set_option(self.raw, k.as_ptr(), value) })
    }

    /// Pin a host buffer (e.g. an Arrow `Buffer`'s bytes) so that imports DMA straight out of it.
    pub fn host_register(&self, ptr: *mut u8, bytes: usize) -> GpuResult<()> {
        check(unsafe { ffi::qgpu_host_register(self.raw, ptr as *mut _, bytes) })
    }

    /// `Vec<RecordBatch>` -> device table.  Zero batches need the schema (an operator below produced nothing).
    pub fn import(self: &Arc<Self>, batches: &[RecordBatch], schema: Option<SchemaRef>) -> GpuResult<GpuTable> {
        let schema: SchemaRef = match (batches.first(), schema) {
            (Some(b), _) => b.schema(),
            (None, Some(s)) => s,
            (None, None) => return Err(GpuError::Query(QueryError::ExecutionError("import of zero batches needs a schema".into()))),
        };
        let mut arrays: Vec<FFI_ArrowArray> = Vec::with_capacity(batches.len());
        for b in batches {
            // a RecordBatch crosses the C Data Interface as a struct array
            let sa: StructArray = b.clone().into();
            let (a, _s) = to_ffi(&sa.to_data()).map_err(|e| GpuError::Query(QueryError::ExecutionError(e.to_string())))?;
            arrays.push(a);
        }
        let mut ffi_schema = FFI_ArrowSchema::try_from(arrow::datatypes::DataType::Struct(schema.fields().clone()))
            .map_err(|e| GpuError::Query(QueryError::ExecutionError(e.to_string())))?;
        let mut out = ptr::null_mut();
        // libqgpu releases the exported structs before it returns, on success and on failure
        let status = unsafe { ffi::qgpu_table_import(self.raw, arrays.as_mut_ptr(), arrays.len() as i64, &mut ffi_schema, &mut out) };
        check(status)?;
        Ok(GpuTable { ctx: Arc::clone(self), raw: out })
    }

    /// Resident tables keyed by name - the role `MemoryDataSource` plays in the reference (query-storage/src/memory.rs).
    pub fn catalog_put(&self, name: &str, table: &GpuTable) -> GpuResult<()> {
        let n = CString::new(name).unwrap();
        check(unsafe { ffi::qgpu_catalog_put(self.raw, n.as_ptr(), table.raw) })
    }

    pub fn catalog_get(self: &Arc<Self>, name: &str) -> GpuResult<Option<GpuTable>> {
        let n = CString::new(name).unwrap();
        let mut out = ptr::null_mut();
        check(unsafe { ffi::qgpu_catalog_get(self.raw, n.as_ptr(), &mut out) })?;
        Ok(if out.is_null() { None } else { Some(GpuTable { ctx: Arc::clone(self), raw: out }) })
    }

    pub fn catalog_drop(&self, name: &str) -> GpuResult<()> {
        let n = CString::new(name).unwrap();
        check(unsafe { ffi::qgpu_catalog_drop(self.raw, n.as_ptr()) })
    }
}

impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { ffi::qgpu_shutdown(self.raw) }
    }
}

/// A device-resident `Vec<RecordBatch>`.
pub struct GpuTable {
    ctx: Arc<GpuContext>,
    raw: *mut ffi::qgpu_table,
}

unsafe impl Send for GpuTable {}
unsafe impl Sync for GpuTable {}

impl GpuTable {
    fn wrap(ctx: &Arc<GpuContext>, raw: *mut ffi::qgpu_table) -> Option<GpuTable> {
        if raw.is_null() {
            None // the reference's `Ok(vec![])`
        } else {
            Some(GpuTable { ctx: Arc::clone(ctx), raw })
        }
    }

    pub fn raw(&self) -> *const ffi::qgpu_table {
        self.raw
    }
    pub fn num_rows(&self) -> i64 {
        unsafe { ffi::qgpu_table_num_rows(self.raw) }
    }
    pub fn num_batches(&self) -> i64 {
        unsafe { ffi::qgpu_table_num_batches(self.raw) }
    }

    /// Device -> host: one RecordBatch per device batch, in pinned buffers owned by the arrays.
    pub fn to_batches(&self) -> GpuResult<Vec<RecordBatch>> {
        let mut out = Vec::with_capacity(self.num_batches() as usize);
        for i in 0..self.num_batches() {
            let mut a = FFI_ArrowArray::empty();
            let mut s = FFI_ArrowSchema::empty();
            check(unsafe { ffi::qgpu_table_export(self.ctx.raw, self.raw, i, &mut a, &mut s) })?;
            let data = unsafe { from_ffi(a, &s) }.map_err(|e| GpuError::Query(QueryError::ExecutionError(e.to_string())))?;
            out.push(RecordBatch::from(StructArray::from(data)));
        }
        Ok(out)
    }

    // ---- the operators: bodies of QueryExecutor::execute_* -------------------------------------------------------
    /// `execute_filter` (executor.rs:131-155)
    pub fn filter(&self, predicate: &ExprBuf) -> GpuResult<Option<GpuTable>> {
        let mut out = ptr::null_mut();
        let e = predicate.as_ffi();
        check(unsafe { ffi::qgpu_execute_filter(self.ctx.raw, self.raw, &e, &mut out) })?;
        Ok(GpuTable::wrap(&self.ctx, out))
    }

    /// `execute_aggregate` (executor.rs:157-190); `predicate`: the Filter directly below the HashAggregate, fused.
    pub fn aggregate(&self, predicate: Option<&ExprBuf>, groups: &[ExprBuf], aggs: &[(i32, ExprBuf)]) -> GpuResult<Option<GpuTable>> {
        let pred = predicate.map(|p| p.as_ffi());
        let g: Vec<ffi::qgpu_expr> = groups.iter().map(|e| e.as_ffi()).collect();
        let a: Vec<ffi::qgpu_agg_expr> = aggs.iter().map(|(f, e)| ffi::qgpu_agg_expr { func: *f, expr: e.as_ffi() }).collect();
        let mut out = ptr::null_mut();
        check(unsafe {
            ffi::qgpu_execute_aggregate(
                self.ctx.raw,
                self.raw,
                pred.as_ref().map_or(ptr::null(), |p| p as *const _),
                g.as_ptr(),
                g.len() as i32,
                a.as_ptr(),
                a.len() as i32,
                &mut out,
            )
        })?;
        Ok(GpuTable::wrap(&self.ctx, out))
    }

    /// `execute_join` (executor.rs:343-361): `join_type` = `query_parser::JoinType as i32`, `on` over left ++ right.
    pub fn join(&self, right: &GpuTable, join_type: i32, on: Option<&ExprBuf>) -> GpuResult<Option<GpuTable>> {
        let on_ffi = on.map(|e| e.as_ffi());
        let mut out = ptr::null_mut();
        check(unsafe {
            ffi::qgpu_execute_join(self.ctx.raw, self.raw, right.raw, join_type, on_ffi.as_ref().map_or(ptr::null(), |p| p as *const _), &mut out)
        })?;
        Ok(GpuTable::wrap(&self.ctx, out))
    }

    /// `execute_projection` (executor.rs:93-129)
    pub fn project(&self, exprs: &[ExprBuf], names: &[String]) -> GpuResult<Option<GpuTable>> {
        let e: Vec<ffi::qgpu_expr> = exprs.iter().map(|x| x.as_ffi()).collect();
        let keep: Vec<CString> = names.iter().map(|n| CString::new(n.as_str()).unwrap()).collect();
        let ptrs: Vec<*const c_char> = keep.iter().map(|c| c.as_ptr()).collect();
        let mut out = ptr::null_mut();
        check(unsafe { ffi::qgpu_execute_projection(self.ctx.raw, self.raw, e.as_ptr(), ptrs.as_ptr(), e.len() as i32, &mut out) })?;
        Ok(GpuTable::wrap(&self.ctx, out))
    }

    /// `execute_limit` (executor.rs:299-341)
    pub fn limit(&self, skip: usize, fetch: Option<usize>) -> GpuResult<Option<GpuTable>> {
        let mut out = ptr::null_mut();
        check(unsafe { ffi::qgpu_execute_limit(self.ctx.raw, self.raw, skip as i64, fetch.map_or(-1, |f| f as i64), &mut out) })?;
        Ok(GpuTable::wrap(&self.ctx, out))
    }

    /// `Merge::sorted_merge` after its concat (query-distributed/src/operators.rs:141-194); `fetch`: Top-K.
    pub fn sort(&self, columns: &[(String, bool, bool)], fetch: Option<usize>) -> GpuResult<Option<GpuTable>> {
        let keep: Vec<CString> = columns.iter().map(|c| CString::new(c.0.as_str()).unwrap()).collect();
        let ptrs: Vec<*const c_char> = keep.iter().map(|c| c.as_ptr()).collect();
        let asc: Vec<i32> = columns.iter().map(|c| c.1 as i32).collect();
        let nf: Vec<i32> = columns.iter().map(|c| c.2 as i32).collect();
        let mut out = ptr::null_mut();
        check(unsafe {
            ffi::qgpu_sort(self.ctx.raw, self.raw, ptrs.as_ptr(), asc.as_ptr(), nf.as_ptr(), ptrs.len() as i32, fetch.map_or(-1, |f| f as i64), &mut out)
        })?;
        Ok(GpuTable::wrap(&self.ctx, out))
    }

    /// `Partitioner::partition_by_hash` (query-distributed/src/partition.rs:151-212): SipHash-1-3 placement, bit-compatible.
    pub fn exchange_hash(&self, key_columns: &[String], num_partitions: usize) -> GpuResult<Vec<Option<GpuTable>>> {
        let keep: Vec<CString> = key_columns.iter().map(|n| CString::new(n.as_str()).unwrap()).collect();
        let ptrs: Vec<*const c_char> = keep.iter().map(|c| c.as_ptr()).collect();
        let mut parts = vec![ptr::null_mut(); num_partitions];
        check(unsafe {
            ffi::qgpu_exchange_hash(self.ctx.raw, self.raw, ptrs.as_ptr(), ptrs.len() as i32, num_partitions as i32, parts.as_mut_ptr())
        })?;
        Ok(parts.into_iter().map(|p| GpuTable::wrap(&self.ctx, p)).collect())
    }
}

impl Drop for GpuTable {
    fn drop(&mut self) {
        unsafe { ffi::qgpu_table_release(self.ctx.raw, self.raw) }
    }
}

/// A `PhysicalExpr` flattened to postfix (`qgpu_expr`): the nodes plus the C strings they point into.
#[derive(Default)]
pub struct ExprBuf {
    nodes: Vec<ffi::qgpu_expr_node>,
    strings: Vec<CString>,
}

impl ExprBuf {
    pub fn new() -> Self {
        Self::default()
    }
    fn node(kind: i32, op: i32) -> ffi::qgpu_expr_node {
        ffi::qgpu_expr_node { kind, op, is_null: 0, reserved: 0, value: ffi::qgpu_value { u64_: 0 }, str_: ptr::null() }
    }
    fn keep(&mut self, s: &str) -> *const c_char {
        self.strings.push(CString::new(s).unwrap_or_default());
        self.strings.last().unwrap().as_ptr() // (a CString's heap buffer does not move when the Vec grows)
    }
    /// `PhysicalExpr::Column { name, index }`
    pub fn column(&mut self, name: &str, index: usize) {
        let mut n = Self::node(ffi::QGPU_EXPR_COLUMN, index as i32);
        n.str_ = self.keep(name);
        self.nodes.push(n);
    }
    /// `PhysicalExpr::Literal(ScalarValue)`: `dtype` = qgpu_dtype, `bits` = the value's 8-byte payload, `null` for X(None) / Null
    pub fn literal(&mut self, dtype: i32, bits: u64, null: bool) {
        let mut n = Self::node(ffi::QGPU_EXPR_LITERAL, dtype);
        n.is_null = null as i32;
        n.value = ffi::qgpu_value { u64_: bits };
        self.nodes.push(n);
    }
    pub fn literal_utf8(&mut self, s: Option<&str>) {
        let mut n = Self::node(ffi::QGPU_EXPR_LITERAL, ffi::QGPU_UTF8);
        match s {
            Some(s) => n.str_ = self.keep(s),
            None => n.is_null = 1,
        }
        self.nodes.push(n);
    }
    /// `PhysicalExpr::BinaryExpr`: push after both operands; `op` = `BinaryOp as i32` (same order as qgpu_binary_op)
    pub fn binary(&mut self, op: i32) {
        self.nodes.push(Self::node(ffi::QGPU_EXPR_BINARY, op));
    }
    /// `PhysicalExpr::UnaryExpr`: push after the operand; `op` = `UnaryOp as i32`
    pub fn unary(&mut self, op: i32) {
        self.nodes.push(Self::node(ffi::QGPU_EXPR_UNARY, op));
    }
    /// ScalarSubquery / InSubquery / Exists / WindowFunction / ScalarFunction (`variant` 0..4): evaluating it fails with
    /// the reference's own message (operators.rs:34-60)
    pub fn unsupported(&mut self, variant: i32) {
        self.nodes.push(Self::node(ffi::QGPU_EXPR_UNSUPPORTED, variant));
    }
    pub fn as_ffi(&self) -> ffi::qgpu_expr {
        ffi::qgpu_expr { nodes: self.nodes.as_ptr(), n_nodes: self.nodes.len() as i32 }
    }
}

/// Schema helper for callers that hold a `query_core::Schema`.
pub fn arrow_schema_of(batches: &[RecordBatch]) -> Option<Arc<Schema>> {
    batches.first().map(|b| b.schema())
}
