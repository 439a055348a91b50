//! Raw bindings of `include/qgpu.h` (libqgpu's C ABI).  One declaration per exported symbol, same order as the header.
#![allow(non_camel_case_types)]

use arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema};
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct qgpu_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct qgpu_table {
    _p: [u8; 0],
}

/// qgpu_status
pub const QGPU_OK: i32 = 0;
pub const QGPU_ERR_EXECUTION: i32 = 1;
pub const QGPU_ERR_ARROW: i32 = 2;
pub const QGPU_ERR_CUDA: i32 = 3;
pub const QGPU_ERR_INVALID: i32 = 4;
pub const QGPU_ERR_PARTITION: i32 = 5;
pub const QGPU_ERR_UNSUPPORTED: i32 = 6;

/// qgpu_dtype
pub const QGPU_NULL: i32 = 0;
pub const QGPU_BOOL: i32 = 1;
pub const QGPU_INT32: i32 = 2;
pub const QGPU_INT64: i32 = 3;
pub const QGPU_FLOAT32: i32 = 4;
pub const QGPU_FLOAT64: i32 = 5;
pub const QGPU_UTF8: i32 = 6;
pub const QGPU_INT8: i32 = 7;
pub const QGPU_INT16: i32 = 8;
pub const QGPU_UINT8: i32 = 9;
pub const QGPU_UINT16: i32 = 10;
pub const QGPU_UINT32: i32 = 11;
pub const QGPU_UINT64: i32 = 12;

/// qgpu_expr_kind
pub const QGPU_EXPR_COLUMN: i32 = 0;
pub const QGPU_EXPR_LITERAL: i32 = 1;
pub const QGPU_EXPR_BINARY: i32 = 2;
pub const QGPU_EXPR_UNARY: i32 = 3;
pub const QGPU_EXPR_UNSUPPORTED: i32 = 4;

#[repr(C)]
#[derive(Clone, Copy)]
pub union qgpu_value {
    pub i64_: i64,
    pub u64_: u64,
    pub f64_: f64,
    pub f32_: f32,
    pub b: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qgpu_expr_node {
    pub kind: i32,
    pub op: i32,
    pub is_null: i32,
    pub reserved: i32,
    pub value: qgpu_value,
    pub str_: *const c_char,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qgpu_expr {
    pub nodes: *const qgpu_expr_node,
    pub n_nodes: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qgpu_agg_expr {
    pub func: i32,
    pub expr: qgpu_expr,
}

pub const QGPU_UNIQUE_ID_BYTES: usize = 128;

extern "C" {
    // ---- context
    pub fn qgpu_version() -> i32;
    pub fn qgpu_last_error() -> *const c_char;
    pub fn qgpu_init(device: i32, out: *mut *mut qgpu_ctx) -> i32;
    pub fn qgpu_shutdown(ctx: *mut qgpu_ctx);
    pub fn qgpu_set_option(ctx: *mut qgpu_ctx, key: *const c_char, value: i64) -> i32;
    pub fn qgpu_synchronize(ctx: *mut qgpu_ctx) -> i32;
    pub fn qgpu_host_register(ctx: *mut qgpu_ctx, ptr: *mut c_void, bytes: usize) -> i32;
    pub fn qgpu_host_unregister(ctx: *mut qgpu_ctx, ptr: *mut c_void) -> i32;
    // ---- Vec<RecordBatch> <-> device
    pub fn qgpu_table_import(
        ctx: *mut qgpu_ctx,
        batches: *mut FFI_ArrowArray,
        n_batches: i64,
        schema: *mut FFI_ArrowSchema,
        out: *mut *mut qgpu_table,
    ) -> i32;
    pub fn qgpu_table_from_schema(schema: *mut FFI_ArrowSchema, out: *mut *mut qgpu_table) -> i32;
    pub fn qgpu_table_num_batches(t: *const qgpu_table) -> i64;
    pub fn qgpu_table_num_rows(t: *const qgpu_table) -> i64;
    pub fn qgpu_table_num_columns(t: *const qgpu_table) -> i32;
    pub fn qgpu_table_export(
        ctx: *mut qgpu_ctx,
        t: *const qgpu_table,
        index: i64,
        out_array: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
    ) -> i32;
    pub fn qgpu_table_release(ctx: *mut qgpu_ctx, t: *mut qgpu_table);
    // ---- resident tables (MemoryDataSource-like catalog, query-storage/src/memory.rs:16-43)
    pub fn qgpu_catalog_put(ctx: *mut qgpu_ctx, name: *const c_char, t: *mut qgpu_table) -> i32;
    pub fn qgpu_catalog_get(ctx: *mut qgpu_ctx, name: *const c_char, out: *mut *mut qgpu_table) -> i32;
    pub fn qgpu_catalog_drop(ctx: *mut qgpu_ctx, name: *const c_char) -> i32;
    // ---- operators
    pub fn qgpu_execute_filter(
        ctx: *mut qgpu_ctx,
        input: *const qgpu_table,
        predicate: *const qgpu_expr,
        out: *mut *mut qgpu_table,
    ) -> i32;
    pub fn qgpu_execute_aggregate(
        ctx: *mut qgpu_ctx,
        input: *const qgpu_table,
        predicate: *const qgpu_expr,
        group_exprs: *const qgpu_expr,
        n_group: i32,
        aggr_exprs: *const qgpu_agg_expr,
        n_aggr: i32,
        out: *mut *mut qgpu_table,
    ) -> i32;
    pub fn qgpu_execute_join(
        ctx: *mut qgpu_ctx,
        left: *const qgpu_table,
        right: *const qgpu_table,
        join_type: i32,
        on: *const qgpu_expr,
        out: *mut *mut qgpu_table,
    ) -> i32;
    pub fn qgpu_execute_projection(
        ctx: *mut qgpu_ctx,
        input: *const qgpu_table,
        exprs: *const qgpu_expr,
        names: *const *const c_char,
        n_exprs: i32,
        out: *mut *mut qgpu_table,
    ) -> i32;
    pub fn qgpu_execute_limit(ctx: *mut qgpu_ctx, input: *const qgpu_table, skip: i64, fetch: i64, out: *mut *mut qgpu_table) -> i32;
    pub fn qgpu_sort(
        ctx: *mut qgpu_ctx,
        input: *const qgpu_table,
        sort_columns: *const *const c_char,
        ascending: *const i32,
        nulls_first: *const i32,
        n_columns: i32,
        fetch: i64,
        out: *mut *mut qgpu_table,
    ) -> i32;
    // ---- exchange
    pub fn qgpu_exchange_hash(
        ctx: *mut qgpu_ctx,
        input: *const qgpu_table,
        key_columns: *const *const c_char,
        n_keys: i32,
        num_partitions: i32,
        out_partitions: *mut *mut qgpu_table,
    ) -> i32;
    pub fn qgpu_comm_unique_id(out_id128: *mut c_void) -> i32;
    pub fn qgpu_comm_init(ctx: *mut qgpu_ctx, id128: *const c_void, rank: i32, world_size: i32) -> i32;
    pub fn qgpu_comm_rank(ctx: *const qgpu_ctx) -> i32;
    pub fn qgpu_comm_size(ctx: *const qgpu_ctx) -> i32;
    pub fn qgpu_shuffle_hash(
        ctx: *mut qgpu_ctx,
        input: *const qgpu_table,
        key_columns: *const *const c_char,
        n_keys: i32,
        out: *mut *mut qgpu_table,
    ) -> i32;
    pub fn qgpu_execute_join_exchange(
        ctx: *mut qgpu_ctx,
        left: *const qgpu_table,
        right: *const qgpu_table,
        join_type: i32,
        on: *const qgpu_expr,
        out: *mut *mut qgpu_table,
        done: *mut i32,
    ) -> i32;
    pub fn qgpu_allgather_table(ctx: *mut qgpu_ctx, input: *const qgpu_table, out: *mut *mut qgpu_table) -> i32;
    // ---- measurement
    pub fn qgpu_profile_enable(ctx: *mut qgpu_ctx, on: i32) -> i32;
    pub fn qgpu_profile_reset(ctx: *mut qgpu_ctx) -> i32;
    pub fn qgpu_profile_count(ctx: *mut qgpu_ctx) -> i64;
    pub fn qgpu_profile_get(ctx: *mut qgpu_ctx, i: i64, name: *mut *const c_char, ms: *mut f32) -> i32;
    pub fn qgpu_timer_start(ctx: *mut qgpu_ctx) -> i32;
    pub fn qgpu_timer_stop(ctx: *mut qgpu_ctx, ms: *mut f32) -> i32;
}
