"""ctypes binding of libqgpu (include/qgpu.h) — the same symbols the Rust `query-gpu` crate binds with
`extern "C"`.  No torch types, no numpy: plain pointers, sizes and Arrow C Data Interface structs.

There is no CPU fallback: if the library is missing or no CUDA device is usable, loading / `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import struct
from pathlib import Path
from typing import List, Optional, Sequence

import pyarrow as pa

from .physical_plan import (
    AggregateExpr,
    ArrowError,
    ExecutionError,
    PhysicalExpr,
    QueryError,
    ScalarType,
    ScalarValue,
)

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libqgpu.so"

QGPU_OK = 0
QGPU_ERR_EXECUTION = 1
QGPU_ERR_ARROW = 2
QGPU_ERR_CUDA = 3
QGPU_ERR_INVALID = 4
QGPU_ERR_PARTITION = 5
QGPU_ERR_UNSUPPORTED = 6

EXPR_COLUMN, EXPR_LITERAL, EXPR_BINARY, EXPR_UNARY, EXPR_UNSUPPORTED = range(5)
_UNSUPPORTED_VARIANTS = ["ScalarSubquery", "InSubquery", "Exists", "WindowFunction", "ScalarFunction"]


class UnsupportedOnGpu(QueryError):
    """QGPU_ERR_UNSUPPORTED: the shape is not on the accelerated path; a caller embedding libqgpu keeps the
    reference's CPU arm for it.  (This package has no such arm: the error propagates.)"""

    def __init__(self, msg: str):
        super().__init__("Unsupported", msg)


class ArrowSchema(C.Structure):
    pass


class ArrowArray(C.Structure):
    pass


ArrowSchema._fields_ = [
    ("format", C.c_char_p),
    ("name", C.c_char_p),
    ("metadata", C.c_char_p),
    ("flags", C.c_int64),
    ("n_children", C.c_int64),
    ("children", C.POINTER(C.POINTER(ArrowSchema))),
    ("dictionary", C.POINTER(ArrowSchema)),
    ("release", C.c_void_p),
    ("private_data", C.c_void_p),
]
ArrowArray._fields_ = [
    ("length", C.c_int64),
    ("null_count", C.c_int64),
    ("offset", C.c_int64),
    ("n_buffers", C.c_int64),
    ("n_children", C.c_int64),
    ("buffers", C.POINTER(C.c_void_p)),
    ("children", C.POINTER(C.POINTER(ArrowArray))),
    ("dictionary", C.POINTER(ArrowArray)),
    ("release", C.c_void_p),
    ("private_data", C.c_void_p),
]


class _Value(C.Union):
    _fields_ = [("i64", C.c_int64), ("u64", C.c_uint64), ("f64", C.c_double), ("f32", C.c_float), ("b", C.c_int32)]


class ExprNode(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("op", C.c_int32),
        ("is_null", C.c_int32),
        ("reserved", C.c_int32),
        ("value", _Value),
        ("str", C.c_char_p),
    ]


class Expr(C.Structure):
    _fields_ = [("nodes", C.POINTER(ExprNode)), ("n_nodes", C.c_int32)]


class AggExpr(C.Structure):
    _fields_ = [("func", C.c_int32), ("expr", Expr)]


_lib = None

# every symbol include/qgpu.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIGNATURES = [
    ("qgpu_version", C.c_int32, []),
    ("qgpu_last_error", C.c_char_p, []),
    ("qgpu_init", C.c_int32, [C.c_int32, C.POINTER(_P)]),
    ("qgpu_shutdown", None, [_P]),
    ("qgpu_set_option", C.c_int32, [_P, C.c_char_p, C.c_int64]),
    ("qgpu_synchronize", C.c_int32, [_P]),
    ("qgpu_host_register", C.c_int32, [_P, _P, C.c_size_t]),
    ("qgpu_host_unregister", C.c_int32, [_P, _P]),
    ("qgpu_table_import", C.c_int32, [_P, C.POINTER(ArrowArray), C.c_int64, C.POINTER(ArrowSchema), C.POINTER(_P)]),
    ("qgpu_table_from_schema", C.c_int32, [C.POINTER(ArrowSchema), C.POINTER(_P)]),
    ("qgpu_table_num_batches", C.c_int64, [_P]),
    ("qgpu_table_num_rows", C.c_int64, [_P]),
    ("qgpu_table_num_columns", C.c_int32, [_P]),
    ("qgpu_table_export", C.c_int32, [_P, _P, C.c_int64, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)]),
    ("qgpu_table_release", None, [_P, _P]),
    ("qgpu_catalog_put", C.c_int32, [_P, C.c_char_p, _P]),
    ("qgpu_catalog_get", C.c_int32, [_P, C.c_char_p, C.POINTER(_P)]),
    ("qgpu_catalog_drop", C.c_int32, [_P, C.c_char_p]),
    ("qgpu_execute_filter", C.c_int32, [_P, _P, C.POINTER(Expr), C.POINTER(_P)]),
    (
        "qgpu_execute_aggregate",
        C.c_int32,
        [_P, _P, C.POINTER(Expr), C.POINTER(Expr), C.c_int32, C.POINTER(AggExpr), C.c_int32, C.POINTER(_P)],
    ),
    ("qgpu_execute_join", C.c_int32, [_P, _P, _P, C.c_int32, C.POINTER(Expr), C.POINTER(_P)]),
    ("qgpu_execute_projection", C.c_int32, [_P, _P, C.POINTER(Expr), C.POINTER(C.c_char_p), C.c_int32, C.POINTER(_P)]),
    ("qgpu_execute_limit", C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.POINTER(_P)]),
    ("qgpu_sort", C.c_int32, [_P, _P, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.POINTER(_P)]),
    ("qgpu_exchange_hash", C.c_int32, [_P, _P, C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.POINTER(_P)]),
    ("qgpu_comm_unique_id", C.c_int32, [_P]),
    ("qgpu_comm_init", C.c_int32, [_P, _P, C.c_int32, C.c_int32]),
    ("qgpu_comm_rank", C.c_int32, [_P]),
    ("qgpu_comm_size", C.c_int32, [_P]),
    ("qgpu_shuffle_hash", C.c_int32, [_P, _P, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(_P)]),
    ("qgpu_execute_join_exchange", C.c_int32, [_P, _P, _P, C.c_int32, C.POINTER(Expr), C.POINTER(_P), C.POINTER(C.c_int32)]),
    ("qgpu_allgather_table", C.c_int32, [_P, _P, C.POINTER(_P)]),
    ("qgpu_profile_enable", C.c_int32, [_P, C.c_int32]),
    ("qgpu_profile_reset", C.c_int32, [_P]),
    ("qgpu_profile_count", C.c_int64, [_P]),
    ("qgpu_profile_get", C.c_int32, [_P, C.c_int64, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]),
    ("qgpu_timer_start", C.c_int32, [_P]),
    ("qgpu_timer_stop", C.c_int32, [_P, C.POINTER(C.c_float)]),
    ("qgpu_debug_compile", C.c_int32, [_P, _P, C.POINTER(Expr), C.c_char_p, C.c_size_t]),
    (
        "qgpu_debug_jit_aggregate",
        C.c_int32,
        [_P, C.POINTER(Expr), C.POINTER(Expr), C.c_int32, C.POINTER(AggExpr), C.c_int32, C.c_int32, C.c_char_p, C.c_size_t],
    ),
]
DECLARED_SYMBOLS = [s[0] for s in _SIGNATURES]


def lib_path() -> Path:
    return Path(os.environ.get("QGPU_LIBRARY", _LIB_PATH))


def load():
    """dlopen libqgpu.so and type every entry point.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m query_engine_b200.build` (there is no CPU fallback)"
        )
    lib = C.CDLL(str(path))
    for name, restype, argtypes in _SIGNATURES:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int):
    if status == QGPU_OK:
        return
    msg = (load().qgpu_last_error() or b"").decode("utf-8", "replace")
    if status == QGPU_ERR_EXECUTION:
        raise ExecutionError(msg)
    if status == QGPU_ERR_ARROW:
        raise ArrowError(msg)
    if status == QGPU_ERR_PARTITION:
        raise QueryError("PartitionError", msg)
    if status == QGPU_ERR_UNSUPPORTED:
        raise UnsupportedOnGpu(msg)
    if status == QGPU_ERR_CUDA:
        raise RuntimeError(f"libqgpu CUDA failure: {msg}")
    raise ValueError(f"libqgpu invalid argument: {msg}")


# ----------------------------------------------------------------------------------------------
# PhysicalExpr -> postfix qgpu_expr
# ----------------------------------------------------------------------------------------------
_INT_RANGES = {
    ScalarType.Int8: (-(2**7), 2**7 - 1),
    ScalarType.Int16: (-(2**15), 2**15 - 1),
    ScalarType.Int32: (-(2**31), 2**31 - 1),
    ScalarType.Int64: (-(2**63), 2**63 - 1),
    ScalarType.UInt8: (0, 2**8 - 1),
    ScalarType.UInt16: (0, 2**16 - 1),
    ScalarType.UInt32: (0, 2**32 - 1),
    ScalarType.UInt64: (0, 2**64 - 1),
}


class FlatExpr:
    """Owns the node array (and the byte strings it points to) of one qgpu_expr."""

    def __init__(self, expr: PhysicalExpr):
        nodes: List[ExprNode] = []
        self._keep: List[bytes] = []
        self._emit(expr, nodes)
        self.array = (ExprNode * len(nodes))(*nodes)
        self.c = Expr(C.cast(self.array, C.POINTER(ExprNode)), len(nodes))

    def _str(self, s: Optional[str]):
        if s is None:
            return None
        b = s.encode("utf-8")
        self._keep.append(b)
        return b

    def _emit(self, e: PhysicalExpr, out: List[ExprNode]):
        n = ExprNode()
        if isinstance(e, PhysicalExpr.Column):
            n.kind, n.op, n.str = EXPR_COLUMN, int(e.index), self._str(e.name)
        elif isinstance(e, PhysicalExpr.Literal):
            v: ScalarValue = e.value
            n.kind, n.op = EXPR_LITERAL, int(v.type)
            n.is_null = 1 if (v.value is None or v.type == ScalarType.Null) else 0
            if not n.is_null:
                if v.type == ScalarType.Boolean:
                    n.value.b = 1 if v.value else 0
                elif v.type == ScalarType.Float64:
                    n.value.f64 = float(v.value)
                elif v.type == ScalarType.Float32:
                    n.value.f32 = float(v.value)
                elif v.type == ScalarType.Utf8:
                    n.str = self._str(str(v.value))
                elif v.type in _INT_RANGES:
                    lo, hi = _INT_RANGES[v.type]
                    iv = int(v.value)
                    if not lo <= iv <= hi:
                        raise ValueError(f"literal {iv} out of range for {v.type.name}")
                    if v.type == ScalarType.UInt64:
                        n.value.u64 = iv
                    else:
                        n.value.i64 = iv
        elif isinstance(e, PhysicalExpr.BinaryExpr):
            self._emit(e.left, out)
            self._emit(e.right, out)
            n.kind, n.op = EXPR_BINARY, int(e.op)
        elif isinstance(e, PhysicalExpr.UnaryExpr):
            self._emit(e.expr, out)
            n.kind, n.op = EXPR_UNARY, int(e.op)
        elif isinstance(e, PhysicalExpr.Unsupported):
            n.kind = EXPR_UNSUPPORTED
            n.op = _UNSUPPORTED_VARIANTS.index(e.variant) if e.variant in _UNSUPPORTED_VARIANTS else 4
        else:
            raise TypeError(f"not a PhysicalExpr: {e!r}")
        out.append(n)


@functools.lru_cache(maxsize=512)
def flat_expr(expr: PhysicalExpr) -> FlatExpr:
    """Flattened form of a (hashable, frozen) expression tree; cached because a served plan is executed many times."""
    return FlatExpr(expr)


@functools.lru_cache(maxsize=256)
def _flat_expr_array(exprs: tuple) -> "FlatExprArray":
    return FlatExprArray(exprs)


@functools.lru_cache(maxsize=256)
def _flat_agg_array(aggs: tuple) -> "FlatAggArray":
    return FlatAggArray(aggs)


class FlatExprArray:
    def __init__(self, exprs: Sequence[PhysicalExpr]):
        self.flat = [FlatExpr(e) for e in exprs]
        self.array = (Expr * max(1, len(self.flat)))(*[f.c for f in self.flat])
        self.n = len(self.flat)


class FlatAggArray:
    def __init__(self, aggs: Sequence[AggregateExpr]):
        self.flat = [FlatExpr(a.expr) for a in aggs]
        self.array = (AggExpr * max(1, len(aggs)))(*[AggExpr(int(a.func), f.c) for a, f in zip(aggs, self.flat)])
        self.n = len(aggs)


def _names(names: Sequence[str]):
    enc = [s.encode("utf-8") for s in names]
    arr = (C.c_char_p * max(1, len(enc)))(*enc)
    return arr, enc


# ----------------------------------------------------------------------------------------------
# Context / Table
# ----------------------------------------------------------------------------------------------
class Table:
    """A device-resident `Vec<RecordBatch>` (qgpu_table*)."""

    def __init__(self, ctx: Optional["Context"], handle: int):
        self.ctx = ctx
        self.handle = handle

    @property
    def num_rows(self) -> int:
        return load().qgpu_table_num_rows(self.handle)

    @property
    def num_batches(self) -> int:
        return load().qgpu_table_num_batches(self.handle)

    @property
    def num_columns(self) -> int:
        return load().qgpu_table_num_columns(self.handle)

    def to_batches(self) -> List[pa.RecordBatch]:
        """Export every batch to host Arrow memory (D2H copies)."""
        lib = load()
        out = []
        for i in range(self.num_batches):
            a, s = ArrowArray(), ArrowSchema()
            check(lib.qgpu_table_export(self.ctx.handle, self.handle, i, C.byref(a), C.byref(s)))
            out.append(pa.RecordBatch._import_from_c(C.addressof(a), C.addressof(s)))
        return out

    def release(self):
        if self.handle:
            # the library finds the table's owning context itself (a table orphaned by qgpu_shutdown only frees
            # host structures), so a closed Context is fine here
            load().qgpu_table_release(self.ctx.handle if self.ctx and self.ctx.handle else None, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def schema_table(schema: pa.Schema) -> Table:
    """Zero-batch, schema-only table; needs no GPU."""
    lib = load()
    s = ArrowSchema()
    schema._export_to_c(C.addressof(s))
    h = _P()
    check(lib.qgpu_table_from_schema(C.byref(s), C.byref(h)))
    return Table(None, h.value)


def debug_compile(schema: pa.Schema, expr: PhysicalExpr) -> str:
    """Type-check + compile `expr` against `schema`; returns the device program listing.  Needs no GPU."""
    lib = load()
    t = schema_table(schema)
    f = FlatExpr(expr)
    buf = C.create_string_buffer(1 << 16)
    check(lib.qgpu_debug_compile(None, t.handle, C.byref(f.c), buf, len(buf)))
    return buf.value.decode()


def debug_jit_aggregate(schema: pa.Schema, predicate, group_exprs, aggr_exprs, rows_per_thread: int = 4) -> str:
    """Specialise the (Filter +) HashAggregate scan kernel for `schema` with NVRTC and return the compile log
    (ptxas -v resource usage).  Compiles for sm_100a without needing a GPU."""
    lib = load()
    t = schema_table(schema)
    fp = FlatExpr(predicate) if predicate is not None else None
    g = FlatExprArray(group_exprs)
    a = FlatAggArray(aggr_exprs)
    buf = C.create_string_buffer(1 << 16)
    check(
        lib.qgpu_debug_jit_aggregate(
            t.handle, C.byref(fp.c) if fp else None, g.array, g.n, a.array, a.n, rows_per_thread, buf, len(buf)
        )
    )
    return buf.value.decode()


class Context:
    """qgpu_ctx*: one CUDA device, one stream, one memory pool (+ an NCCL communicator after comm_init)."""

    def __init__(self, device: int = 0):
        lib = load()
        h = _P()
        check(lib.qgpu_init(device, C.byref(h)))
        self.handle = h.value
        self.device = device

    def close(self):
        if self.handle:
            load().qgpu_shutdown(self.handle)
            self.handle = None

    def __del__(self):
        # tables may outlive us in arbitrary GC order; leave the context to process teardown
        pass

    def set_option(self, key: str, value: int):
        check(load().qgpu_set_option(self.handle, key.encode(), int(value)))

    def synchronize(self):
        check(load().qgpu_synchronize(self.handle))

    def host_register(self, address: int, nbytes: int):
        check(load().qgpu_host_register(self.handle, address, nbytes))

    def host_unregister(self, address: int):
        check(load().qgpu_host_unregister(self.handle, address))

    # ---- tables -----------------------------------------------------------------------------
    def import_batches(self, batches: Sequence[pa.RecordBatch], schema: Optional[pa.Schema] = None) -> Table:
        lib = load()
        if schema is None:
            if not batches:
                raise ValueError("need a schema to import zero batches")
            schema = batches[0].schema
        n = len(batches)
        arr = (ArrowArray * max(1, n))()
        s = ArrowSchema()
        try:
            for i, b in enumerate(batches):
                if b.schema != schema:
                    b = b.cast(schema)
                b._export_to_c(C.addressof(arr[i]))
            schema._export_to_c(C.addressof(s))
        except Exception:
            # nothing was handed to the library yet: give back what was exported so far
            for a in list(arr) + [s]:
                if a.release:
                    C.CFUNCTYPE(None, C.c_void_p)(a.release)(C.addressof(a))
            raise
        h = _P()
        # the library releases the exported structs before returning, on success and on failure
        check(lib.qgpu_table_import(self.handle, arr, n, C.byref(s), C.byref(h)))
        return Table(self, h.value)

    # ---- resident tables by name ---------------------------------------------------------------
    def catalog_put(self, name: str, table: Table):
        check(load().qgpu_catalog_put(self.handle, name.encode(), table.handle))

    def catalog_get(self, name: str) -> Optional[Table]:
        h = _P()
        check(load().qgpu_catalog_get(self.handle, name.encode(), C.byref(h)))
        return self._wrap(h)

    def catalog_drop(self, name: str):
        check(load().qgpu_catalog_drop(self.handle, name.encode()))

    def _wrap(self, h) -> Optional[Table]:
        return Table(self, h.value) if h.value else None

    # ---- operators --------------------------------------------------------------------------
    def execute_filter(self, table: Table, predicate: PhysicalExpr) -> Optional[Table]:
        f = flat_expr(predicate)
        h = _P()
        check(load().qgpu_execute_filter(self.handle, table.handle, C.byref(f.c), C.byref(h)))
        return self._wrap(h)

    def execute_aggregate(
        self,
        table: Table,
        group_exprs: Sequence[PhysicalExpr],
        aggr_exprs: Sequence[AggregateExpr],
        predicate: Optional[PhysicalExpr] = None,
    ) -> Optional[Table]:
        fp = flat_expr(predicate) if predicate is not None else None
        g = _flat_expr_array(tuple(group_exprs))
        a = _flat_agg_array(tuple(aggr_exprs))
        h = _P()
        check(
            load().qgpu_execute_aggregate(
                self.handle, table.handle, C.byref(fp.c) if fp else None, g.array, g.n, a.array, a.n, C.byref(h)
            )
        )
        return self._wrap(h)

    def execute_join(self, left: Table, right: Table, join_type: int, on: Optional[PhysicalExpr]) -> Optional[Table]:
        f = flat_expr(on) if on is not None else None
        h = _P()
        check(
            load().qgpu_execute_join(
                self.handle, left.handle, right.handle, int(join_type), C.byref(f.c) if f else None, C.byref(h)
            )
        )
        return self._wrap(h)

    def execute_projection(self, table: Table, exprs: Sequence[PhysicalExpr], names: Sequence[str]) -> Optional[Table]:
        e = FlatExprArray(exprs)
        arr, _keep = _names(names)
        h = _P()
        check(load().qgpu_execute_projection(self.handle, table.handle, e.array, arr, e.n, C.byref(h)))
        return self._wrap(h)

    def execute_limit(self, table: Table, skip: int, fetch: Optional[int]) -> Optional[Table]:
        h = _P()
        check(load().qgpu_execute_limit(self.handle, table.handle, int(skip), -1 if fetch is None else int(fetch), C.byref(h)))
        return self._wrap(h)

    def sort(self, table: Table, columns: Sequence[str], ascending: Sequence[bool], nulls_first: Sequence[bool],
             fetch: Optional[int] = None) -> Optional[Table]:
        """lexsort by the named columns + take (Merge::sorted_merge after its concat); `fetch`: keep the first rows only"""
        arr, _keep = _names(columns)
        n = len(columns)
        asc = (C.c_int32 * max(1, n))(*[1 if a else 0 for a in ascending])
        nf = (C.c_int32 * max(1, n))(*[1 if a else 0 for a in nulls_first])
        h = _P()
        check(load().qgpu_sort(self.handle, table.handle, arr, asc, nf, n, -1 if fetch is None else int(fetch), C.byref(h)))
        return self._wrap(h)

    # ---- exchange ---------------------------------------------------------------------------
    def exchange_hash(self, table: Table, key_columns: Sequence[str], num_partitions: int) -> List[Optional[Table]]:
        arr, _keep = _names(key_columns)
        outs = (_P * max(1, num_partitions))()
        check(load().qgpu_exchange_hash(self.handle, table.handle, arr, len(key_columns), num_partitions, outs))
        return [Table(self, outs[i]) if outs[i] else None for i in range(num_partitions)]

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(load().qgpu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world_size: int):
        if len(unique_id) != 128:
            raise ValueError("unique id must be 128 bytes")
        check(load().qgpu_comm_init(self.handle, unique_id, rank, world_size))

    @property
    def rank(self) -> int:
        return load().qgpu_comm_rank(self.handle)

    @property
    def world_size(self) -> int:
        return load().qgpu_comm_size(self.handle)

    def shuffle_hash(self, table: Table, key_columns: Sequence[str]) -> Optional[Table]:
        arr, _keep = _names(key_columns)
        h = _P()
        check(load().qgpu_shuffle_hash(self.handle, table.handle, arr, len(key_columns), C.byref(h)))
        return self._wrap(h)

    def execute_join_exchange(self, left: Table, right: Table, join_type: int, on: Optional[PhysicalExpr]):
        """(done, table): the multi-GPU join whose radix scatter is the exchange (collective).  done = False: not
        eligible on this input (every rank sees the same answer) - shuffle and join locally instead."""
        f = flat_expr(on) if on is not None else None
        h = _P()
        done = C.c_int32(0)
        check(
            load().qgpu_execute_join_exchange(
                self.handle, left.handle, right.handle, int(join_type), C.byref(f.c) if f else None, C.byref(h), C.byref(done)
            )
        )
        return bool(done.value), self._wrap(h)

    def allgather(self, table: Table) -> Optional[Table]:
        h = _P()
        check(load().qgpu_allgather_table(self.handle, table.handle, C.byref(h)))
        return self._wrap(h)

    # ---- measurement ------------------------------------------------------------------------
    def profile_enable(self, on: bool):
        check(load().qgpu_profile_enable(self.handle, 1 if on else 0))

    def profile_reset(self):
        check(load().qgpu_profile_reset(self.handle))

    def profile_count(self) -> int:
        return load().qgpu_profile_count(self.handle)

    def profile_entries(self):
        """[(kernel name, ms)] of the timed launches since the last reset (events must have been enabled)."""
        lib = load()
        out = []
        i = 0
        name, ms = C.c_char_p(), C.c_float()
        while lib.qgpu_profile_get(self.handle, i, C.byref(name), C.byref(ms)) == QGPU_OK:
            out.append((name.value.decode(), ms.value))
            i += 1
        return out

    def timer_start(self):
        check(load().qgpu_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(load().qgpu_timer_stop(self.handle, C.byref(ms)))
        return ms.value


_ = struct  # (kept for callers that pack unique ids)
