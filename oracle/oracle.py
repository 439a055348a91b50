"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

A numpy/pyarrow restatement of the reference's algorithm for the hot path
(filter / hash aggregate / hash join of `query-executor`, hash exchange of
`query-distributed`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module,
and only as the checker.

Pinning status
--------------
The reference is Rust + the un-vendored `arrow` crate 53.4.1 (Cargo.lock:156-157)
and cannot be compiled in this image (no cargo/rustc).  The reference holds NO
test, golden vector or fixture for this path.  Therefore:

* Tier 1 (filter, expression evaluation, scalar aggregates): restated line by
  line from the Rust cited on each function, and pinned by (a) the hand-derived
  golden on the reference's own `data/employees.csv` (tests/golden/), and (b)
  cross-checks against pyarrow.compute wherever Arrow C++ and arrow-rs agree.
* Tier 2 (GROUP BY, equi-join): the reference returns `Ok(vec![])` for grouped
  aggregation (executor.rs:189) and a Cartesian product that ignores `on`
  (executor.rs:363-399,500-540).  The oracle implements the documented
  completion and ALSO the literal behaviour (`reference_asis=True`).
  **Parity for Tier 2 is unpinned by the reference.**
* Exchange: SipHash-1-3 (Rust `DefaultHasher`, keys 0,0) restated from the
  published algorithm; the skeleton is checked against the official SipHash-2-4
  vectors; the reference only pins row conservation (partition.rs:401-415).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa

from query_engine_b200.physical_plan import (
    AggregateExpr,
    AggregateFunction,
    ArrowError,
    BinaryOp,
    ExecutionError,
    JoinType,
    PhysicalExpr,
    PhysicalPlan,
    QueryError,
    ScalarType,
    ScalarValue,
    UnaryOp,
    UNSUPPORTED_EXPR_MESSAGES,
)

I64_MIN = -(2**63)
I64_MAX = 2**63 - 1
I32_MIN = -(2**31)
I32_MAX = 2**31 - 1


# ---------------------------------------------------------------------------
# internal typed array
# ---------------------------------------------------------------------------
@dataclass
class Arr:
    """dtype in {'null','bool','i32','i64','f32','f64','utf8', 'other:<arrow type>'}"""

    dtype: str
    values: np.ndarray
    valid: Optional[np.ndarray] = None  # bool array, None = all valid

    def __len__(self):
        return len(self.values)

    def validity(self) -> np.ndarray:
        if self.valid is None:
            return np.ones(len(self.values), dtype=bool)
        return self.valid

    def null_count(self) -> int:
        if self.dtype == "null":
            return len(self.values)
        return 0 if self.valid is None else int((~self.valid).sum())


_NP = {"bool": np.bool_, "i32": np.int32, "i64": np.int64, "f32": np.float32, "f64": np.float64}
_PA = {
    "bool": pa.bool_(),
    "i32": pa.int32(),
    "i64": pa.int64(),
    "f32": pa.float32(),
    "f64": pa.float64(),
    "utf8": pa.string(),
    "null": pa.null(),
}


def _dtype_of_arrow(t: pa.DataType) -> str:
    if pa.types.is_boolean(t):
        return "bool"
    if pa.types.is_int32(t):
        return "i32"
    if pa.types.is_int64(t):
        return "i64"
    if pa.types.is_float32(t):
        return "f32"
    if pa.types.is_float64(t):
        return "f64"
    if pa.types.is_string(t) or pa.types.is_large_string(t):
        return "utf8"
    if pa.types.is_null(t):
        return "null"
    return f"other:{t}"


def from_arrow(a: pa.Array) -> Arr:
    if isinstance(a, pa.ChunkedArray):
        a = a.combine_chunks()
    dt = _dtype_of_arrow(a.type)
    n = len(a)
    if dt == "null":
        return Arr("null", np.zeros(n, dtype=np.int8), np.zeros(n, dtype=bool))
    valid = None
    if a.null_count:
        valid = ~a.is_null().to_numpy(zero_copy_only=False)
    if dt == "utf8":
        return Arr(dt, np.array(a.to_pylist(), dtype=object), valid)
    if dt.startswith("other"):
        return Arr(dt, np.array(a.to_pylist(), dtype=object), valid)
    if dt == "bool":
        vals = a.fill_null(False).to_numpy(zero_copy_only=False).astype(bool)
    else:
        vals = a.fill_null(0).to_numpy(zero_copy_only=False).astype(_NP[dt], copy=False)
    return Arr(dt, np.ascontiguousarray(vals), valid)


def to_arrow(x: Arr) -> pa.Array:
    n = len(x)
    if x.dtype == "null":
        return pa.nulls(n)
    mask = None if x.valid is None else ~x.valid
    if x.dtype == "utf8" or x.dtype.startswith("other"):
        vals = [None if (mask is not None and mask[i]) else x.values[i] for i in range(n)]
        return pa.array(vals, type=_PA.get(x.dtype, None))
    return pa.array(x.values, type=_PA[x.dtype], mask=mask)


# ---------------------------------------------------------------------------
# float totalOrder keys (arrow-rs cmp / min / max order floats by IEEE-754
# totalOrder: -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN)
# ---------------------------------------------------------------------------
def total_order_key(vals: np.ndarray) -> np.ndarray:
    if vals.dtype == np.float64:
        b = vals.view(np.int64)
        return b ^ ((b >> 63) & np.int64(0x7FFFFFFFFFFFFFFF))
    if vals.dtype == np.float32:
        b = vals.view(np.int32)
        return b ^ ((b >> 31) & np.int32(0x7FFFFFFF))
    raise TypeError(vals.dtype)


# ---------------------------------------------------------------------------
# literals  (operators.rs:322-347 create_literal_array)
# ---------------------------------------------------------------------------
def create_literal_array(val: ScalarValue, size: int) -> Arr:
    t = val.type
    v = val.value
    if t == ScalarType.Null or v is None:
        # `ScalarValue::Null` and every `X(None)` fall to NullArray (:344-345)
        return Arr("null", np.zeros(size, dtype=np.int8), np.zeros(size, dtype=bool))
    if t == ScalarType.Boolean:
        return Arr("bool", np.full(size, bool(v), dtype=bool))
    if t == ScalarType.Int32:
        return Arr("i32", np.full(size, v, dtype=np.int32))
    if t == ScalarType.Int64:
        return Arr("i64", np.full(size, v, dtype=np.int64))
    if t == ScalarType.Float32:
        return Arr("f32", np.full(size, v, dtype=np.float32))
    if t == ScalarType.Float64:
        return Arr("f64", np.full(size, v, dtype=np.float64))
    if t == ScalarType.Utf8:
        return Arr("utf8", np.full(size, v, dtype=object))
    # Int8/16, UInt*: arrays the arithmetic/compare ladders do not accept
    return Arr(f"other:{t.name}", np.full(size, v, dtype=object))


# ---------------------------------------------------------------------------
# expression evaluation (operators.rs:13-62)
# ---------------------------------------------------------------------------
def evaluate_expr(expr: PhysicalExpr, batch: pa.RecordBatch) -> Arr:
    if isinstance(expr, PhysicalExpr.Column):
        if expr.index >= batch.num_columns or expr.index < 0:
            raise ExecutionError(f"Column index {expr.index} out of bounds")  # :16-21
        return from_arrow(batch.column(expr.index))
    if isinstance(expr, PhysicalExpr.Literal):
        return create_literal_array(expr.value, batch.num_rows)
    if isinstance(expr, PhysicalExpr.BinaryExpr):
        l = evaluate_expr(expr.left, batch)
        r = evaluate_expr(expr.right, batch)
        return evaluate_binary_op(l, expr.op, r)
    if isinstance(expr, PhysicalExpr.UnaryExpr):
        return evaluate_unary_op(evaluate_expr(expr.expr, batch), expr.op)
    if isinstance(expr, PhysicalExpr.Unsupported):
        raise ExecutionError(
            UNSUPPORTED_EXPR_MESSAGES.get(expr.variant, f"{expr.variant} is not on the accelerated path")
        )
    raise TypeError(expr)


def _both_valid(l: Arr, r: Arr) -> Optional[np.ndarray]:
    if l.valid is None and r.valid is None:
        return None
    return l.validity() & r.validity()


_ARITH_NAME = {
    BinaryOp.Add: "addition",
    BinaryOp.Subtract: "subtraction",
    BinaryOp.Multiply: "multiplication",
    BinaryOp.Divide: "division",
}
_ARITH_SYM = {BinaryOp.Add: "+", BinaryOp.Subtract: "-", BinaryOp.Multiply: "*", BinaryOp.Divide: "/"}


def _checked_int_arith(l: Arr, op: BinaryOp, r: Arr) -> Arr:
    """arrow-rs numeric::{add,sub,mul,div} on integers are CHECKED: overflow and
    division by zero are errors, evaluated on valid slots only (operators.rs:389-500)."""
    bits = 64 if l.dtype == "i64" else 32
    lo, hi = (I64_MIN, I64_MAX) if bits == 64 else (I32_MIN, I32_MAX)
    valid = _both_valid(l, r)
    vmask = np.ones(len(l), dtype=bool) if valid is None else valid
    a = l.values.astype(np.int64)
    b = r.values.astype(np.int64)
    sym = _ARITH_SYM[op]
    if op == BinaryOp.Divide:
        zero = vmask & (b == 0)
        if zero.any():
            raise ArrowError("Divide by zero error")
        ovf = vmask & (a == lo) & (b == -1)
        if ovf.any():
            i = int(np.flatnonzero(ovf)[0])
            raise ArrowError(f"Compute error: Overflow happened on: {int(a[i])} {sym} {int(b[i])}")
        bb = np.where(b == 0, 1, b)
        bb = np.where((a == lo) & (bb == -1), 1, bb)
        res = _trunc_div(a, bb)  # Rust integer division truncates toward zero
    else:
        # exact via Python ints where a 64-bit wrap is possible, numpy elsewhere
        if bits == 32:
            if op == BinaryOp.Add:
                res = a + b
            elif op == BinaryOp.Subtract:
                res = a - b
            else:
                res = a * b
            ovf = vmask & ((res < lo) | (res > hi))
        else:
            with np.errstate(over="ignore"):
                if op == BinaryOp.Add:
                    res = a + b
                    ovf = ((a ^ res) & (b ^ res)) < 0
                elif op == BinaryOp.Subtract:
                    res = a - b
                    ovf = ((a ^ b) & (a ^ res)) < 0
                else:
                    res = a * b
                    small = (np.abs(a) < 2**31) & (np.abs(b) < 2**31) & (a != I64_MIN) & (b != I64_MIN)
                    ovf = np.zeros(len(a), dtype=bool)
                    for i in np.flatnonzero(~small):
                        p = int(a[i]) * int(b[i])
                        ovf[i] = p < lo or p > hi
            ovf = vmask & ovf
        if ovf.any():
            i = int(np.flatnonzero(ovf)[0])
            raise ArrowError(f"Compute error: Overflow happened on: {int(a[i])} {sym} {int(b[i])}")
    return Arr(l.dtype, res.astype(_NP[l.dtype]), valid)


def _float_arith(l: Arr, op: BinaryOp, r: Arr) -> Arr:
    valid = _both_valid(l, r)
    a, b = l.values, r.values
    with np.errstate(all="ignore"):
        if op == BinaryOp.Add:
            res = a + b
        elif op == BinaryOp.Subtract:
            res = a - b
        elif op == BinaryOp.Multiply:
            res = a * b
        else:
            res = a / b  # IEEE: x/0 = +-inf, 0/0 = NaN, no error
    return Arr(l.dtype, res.astype(_NP[l.dtype]), valid)


def modulo_op(l: Arr, r: Arr) -> Arr:
    """operators.rs:711-743: Int64/Int32 only, r == 0 gives NULL, Rust `%`
    (sign of the dividend)."""
    if l.dtype == r.dtype and l.dtype in ("i64", "i32"):
        a = l.values.astype(np.int64)
        b = r.values.astype(np.int64)
        both = l.validity() & r.validity()
        lo = I64_MIN if l.dtype == "i64" else I32_MIN
        # Rust panics (process abort with panic=abort) on MIN % -1; we surface it as an error
        bad = both & (a == lo) & (b == -1)
        if bad.any():
            raise ArrowError("Compute error: Overflow happened on: remainder MIN % -1")
        nz = b != 0
        bb = np.where(nz, b, 1)
        bb = np.where((a == lo) & (bb == -1), 1, bb)
        res = a - bb * _trunc_div(a, bb)
        valid = both & nz
        return Arr(l.dtype, res.astype(_NP[l.dtype]), valid)
    raise ExecutionError("Modulo operation requires integer arrays")


def _trunc_div(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    q = np.floor_divide(a, b)
    rem = a - q * b
    adj = (rem != 0) & ((a < 0) != (b < 0))
    return q + adj.astype(np.int64)


def cast_to_float64(x: Arr) -> Arr:
    """operators.rs:678-709"""
    if x.dtype in ("f32", "i64", "i32"):
        return Arr("f64", x.values.astype(np.float64), x.valid)
    return x


def coerce_numeric_types(l: Arr, r: Arr) -> Tuple[Arr, Arr]:
    """operators.rs:616-675"""
    if l.dtype == r.dtype:
        return l, r
    if l.dtype == "f64" and r.dtype in ("i64", "i32"):
        return l, Arr("f64", r.values.astype(np.float64), r.valid)
    if r.dtype == "f64" and l.dtype in ("i64", "i32"):
        return Arr("f64", l.values.astype(np.float64), l.valid), r
    if l.dtype == "f32" or r.dtype == "f32":
        return cast_to_float64(l), cast_to_float64(r)
    if l.dtype == "i64" and r.dtype == "i32":
        return l, Arr("i64", r.values.astype(np.int64), r.valid)
    if r.dtype == "i64" and l.dtype == "i32":
        return Arr("i64", l.values.astype(np.int64), l.valid), r
    return l, r


_CMP_SYM = {
    BinaryOp.Equal: "==",
    BinaryOp.NotEqual: "!=",
    BinaryOp.Less: "<",
    BinaryOp.LessEqual: "<=",
    BinaryOp.Greater: ">",
    BinaryOp.GreaterEqual: ">=",
}


_ARROW_NAMES = {"null": "Null", "bool": "Boolean", "i32": "Int32", "i64": "Int64", "f32": "Float32", "f64": "Float64", "utf8": "Utf8"}


def _arrow_name(dt: str) -> str:
    """arrow-rs `DataType` Display name, as it appears in the ArrowError text"""
    return _ARROW_NAMES.get(dt, dt)


def _compare(l: Arr, op: BinaryOp, r: Arr) -> Arr:
    """arrow-rs cmp::{eq,neq,lt,lt_eq,gt,gt_eq}: result null iff either side null;
    floats by totalOrder; mismatching types are an error."""
    l, r = coerce_numeric_types(l, r)
    if l.dtype != r.dtype or l.dtype.startswith("other"):
        raise ArrowError(
            f"Invalid argument error: Invalid comparison operation: {_arrow_name(l.dtype)} {_CMP_SYM[op]} {_arrow_name(r.dtype)}"
        )
    n = len(l)
    valid = _both_valid(l, r)
    if l.dtype == "null":
        # arrow-rs compares NullArrays to an all-null boolean array
        return Arr("bool", np.zeros(n, dtype=bool), np.zeros(n, dtype=bool))
    if l.dtype in ("f64", "f32"):
        a, b = total_order_key(l.values), total_order_key(r.values)
    elif l.dtype == "utf8":
        vm = np.ones(n, dtype=bool) if valid is None else valid
        a = np.array([(x.encode() if vm[i] else b"") for i, x in enumerate(l.values)], dtype=object)
        b = np.array([(x.encode() if vm[i] else b"") for i, x in enumerate(r.values)], dtype=object)
    else:
        a, b = l.values, r.values
    if op == BinaryOp.Equal:
        res = a == b
    elif op == BinaryOp.NotEqual:
        res = a != b
    elif op == BinaryOp.Less:
        res = a < b
    elif op == BinaryOp.LessEqual:
        res = a <= b
    elif op == BinaryOp.Greater:
        res = a > b
    else:
        res = a >= b
    return Arr("bool", np.asarray(res, dtype=bool), valid)


def evaluate_binary_op(l: Arr, op: BinaryOp, r: Arr) -> Arr:
    """operators.rs:382-612"""
    if op in _ARITH_NAME:
        # same-type ladders only: Int64, Int32, Float64, Float32 (no coercion)
        if l.dtype == r.dtype and l.dtype in ("i64", "i32"):
            return _checked_int_arith(l, op, r)
        if l.dtype == r.dtype and l.dtype in ("f64", "f32"):
            return _float_arith(l, op, r)
        raise ExecutionError(f"Unsupported types for {_ARITH_NAME[op]}")
    if op == BinaryOp.Modulo:
        return modulo_op(l, r)
    if op in _CMP_SYM:
        return _compare(l, op, r)
    if op in (BinaryOp.And, BinaryOp.Or):
        name = "AND" if op == BinaryOp.And else "OR"
        if l.dtype != "bool" or r.dtype != "bool":
            raise ExecutionError(f"{name} requires boolean arrays")
        # arrow compute::and / or are NOT Kleene: null if either side is null (:552,:568)
        valid = _both_valid(l, r)
        res = (l.values & r.values) if op == BinaryOp.And else (l.values | r.values)
        return Arr("bool", res, valid)
    if op == BinaryOp.TsMatch:
        raise ExecutionError("@@ left operand must be tsvector (string)")
    raise TypeError(op)


def evaluate_unary_op(x: Arr, op: UnaryOp) -> Arr:
    """operators.rs:349-380"""
    if op == UnaryOp.Not:
        if x.dtype != "bool":
            raise ExecutionError("NOT operator requires boolean array")
        return Arr("bool", ~x.values, x.valid)
    if op == UnaryOp.CastFloat64:
        # extension used by the multi-GPU final aggregate (not a reference UnaryOp): `x as f64`, NULL stays NULL
        if x.dtype not in ("i64", "i32", "f64", "f32"):
            raise ExecutionError("Unsupported type for cast to Float64")
        return Arr("f64", x.values.astype(np.float64), x.valid)
    if x.dtype in ("i64", "i32"):
        with np.errstate(over="ignore"):
            return Arr(x.dtype, (-x.values).astype(_NP[x.dtype]), x.valid)  # release build: wraps
    if x.dtype in ("f64", "f32"):
        return Arr(x.dtype, -x.values, x.valid)
    raise ExecutionError("Unsupported type for negation")


# ---------------------------------------------------------------------------
# aggregates (operators.rs:745-848)
# ---------------------------------------------------------------------------
def _wrap_i64(v: int) -> int:
    return ((v + 2**63) % 2**64) - 2**63


def _wrap_i32(v: int) -> int:
    return ((v + 2**31) % 2**32) - 2**31


def evaluate_aggregate(func: AggregateFunction, x: Arr) -> ScalarValue:
    n = len(x)
    vm = x.validity() if x.dtype != "null" else np.zeros(n, dtype=bool)
    nn = int(vm.sum())
    if func == AggregateFunction.Count:
        return ScalarValue.Int64(n - x.null_count())  # :747-750
    name = {AggregateFunction.Sum: "SUM", AggregateFunction.Avg: "AVG", AggregateFunction.Min: "MIN", AggregateFunction.Max: "MAX"}[func]
    if x.dtype not in ("i64", "i32", "f64", "f32"):
        raise ExecutionError(f"Unsupported type for {name}")
    vals = x.values[vm]

    def isum():
        # arrow sum on integers wraps; Int32 sums wrap in i32 and are then widened (:755-757)
        s = int(vals.astype(object).sum()) if nn else 0
        return _wrap_i64(s) if x.dtype == "i64" else _wrap_i32(s)

    def fsum():
        if x.dtype == "f64":
            return float(np.sum(vals, dtype=np.float64)) if nn else 0.0
        # Float32: arrow sums in f32, then `as f64`
        return float(np.sum(vals, dtype=np.float32)) if nn else 0.0

    if func == AggregateFunction.Sum:
        if x.dtype in ("i64", "i32"):
            return ScalarValue.Int64(isum() if nn else None)
        return ScalarValue.Float64(fsum() if nn else None)
    if func == AggregateFunction.Avg:
        if nn == 0:
            return ScalarValue.Float64(None)
        s = float(isum()) if x.dtype in ("i64", "i32") else fsum()
        return ScalarValue.Float64(s / float(nn))
    # Min / Max: type preserving, floats by totalOrder (NaN greatest, -0 < +0)
    ctor = {"i64": ScalarValue.Int64, "i32": ScalarValue.Int32, "f64": ScalarValue.Float64, "f32": ScalarValue.Float32}[x.dtype]
    if nn == 0:
        return ctor(None)
    if x.dtype in ("f64", "f32"):
        keys = total_order_key(vals)
        i = int(np.argmin(keys)) if func == AggregateFunction.Min else int(np.argmax(keys))
        return ctor(float(vals[i]))
    return ctor(int(vals.min()) if func == AggregateFunction.Min else int(vals.max()))


def _scalar_to_column(v: ScalarValue) -> pa.Array:
    """executor.rs:214-288 scalar_value_to_array"""
    from query_engine_b200.physical_plan import scalar_arrow_type

    if v.type == ScalarType.Null:
        return pa.nulls(1)
    return pa.array([v.value], type=scalar_arrow_type(v.type))


def scalar_values_to_batch(values: Sequence[ScalarValue]) -> List[pa.RecordBatch]:
    """executor.rs:192-212: one 1-row batch, fields col_0.., nullable"""
    if not values:
        return []
    arrays = [_scalar_to_column(v) for v in values]
    fields = [pa.field(f"col_{i}", a.type, True) for i, a in enumerate(arrays)]
    return [pa.RecordBatch.from_arrays(arrays, schema=pa.schema(fields))]


# ---------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------
def execute_filter(batches: Sequence[pa.RecordBatch], predicate: PhysicalExpr) -> List[pa.RecordBatch]:
    """executor.rs:131-155"""
    out = []
    for batch in batches:
        m = evaluate_expr(predicate, batch)
        if m.dtype != "bool":
            raise ExecutionError("Filter predicate must return boolean")
        keep = m.values & m.validity()  # arrow filter: null mask slot == drop
        fb = batch.filter(pa.array(keep))
        if fb.num_rows > 0:
            out.append(fb)
    return out


def _concat(arrs: Sequence[Arr]) -> Arr:
    dt = arrs[0].dtype
    for a in arrs:
        if a.dtype != dt:
            raise ArrowError("Invalid argument error: It is not possible to concatenate arrays of different data types.")
    vals = np.concatenate([a.values for a in arrs]) if arrs else np.zeros(0)
    valid = None
    if any(a.valid is not None for a in arrs):
        valid = np.concatenate([a.validity() for a in arrs])
    return Arr(dt, vals, valid)


def group_field_name(e: PhysicalExpr, i: int) -> str:
    return e.name if isinstance(e, PhysicalExpr.Column) else f"group_{i}"


def _group_key_of(a: Arr, i: int):
    if a.dtype == "null" or (a.valid is not None and not a.valid[i]):
        return None
    v = a.values[i]
    if a.dtype == "f64":
        return ("f", int(np.float64(v).view(np.int64)))
    if a.dtype == "f32":
        return ("f", int(np.float32(v).view(np.int32)))
    if a.dtype == "bool":
        return bool(v)
    if a.dtype in ("i64", "i32"):
        return int(v)
    return v


def execute_aggregate(
    batches: Sequence[pa.RecordBatch],
    group_exprs: Sequence[PhysicalExpr],
    aggr_exprs: Sequence[AggregateExpr],
    reference_asis: bool = False,
) -> List[pa.RecordBatch]:
    """executor.rs:157-190 (scalar path, Tier 1) + the Tier-2 grouped completion."""
    if not aggr_exprs:
        return []  # :163-165
    if not group_exprs:
        results = []
        for ae in aggr_exprs:
            vals = [evaluate_expr(ae.expr, b) for b in batches]
            if vals:  # :178  zero batches -> aggregate skipped
                results.append(evaluate_aggregate(ae.func, _concat(vals)))
        return scalar_values_to_batch(results)
    if reference_asis:
        return []  # :189
    if not batches:
        return []
    # ---- Tier 2: GROUP BY ----
    gcols = [_concat([evaluate_expr(g, b) for b in batches]) for g in group_exprs]
    acols = [_concat([evaluate_expr(a.expr, b) for b in batches]) for a in aggr_exprs]
    n = len(gcols[0])
    if n == 0:
        return []
    order: dict = {}
    gid = np.empty(n, dtype=np.int64)
    first_row = []
    for i in range(n):
        k = tuple(_group_key_of(g, i) for g in gcols)
        j = order.get(k)
        if j is None:
            j = len(order)
            order[k] = j
            first_row.append(i)
        gid[i] = j
    ng = len(order)
    first_row = np.array(first_row, dtype=np.int64)
    out_arrays = []
    fields = []
    for gi, g in enumerate(gcols):
        sub = Arr(g.dtype, g.values[first_row], None if g.valid is None else g.valid[first_row])
        arr = to_arrow(sub)
        out_arrays.append(arr)
        fields.append(pa.field(group_field_name(group_exprs[gi], gi), arr.type, True))
    sort_idx = np.argsort(gid, kind="stable")
    bounds = np.searchsorted(gid[sort_idx], np.arange(ng + 1))
    for ai, (ae, ac) in enumerate(zip(aggr_exprs, acols)):
        scalars = []
        for g in range(ng):
            rows = sort_idx[bounds[g] : bounds[g + 1]]
            sub = Arr(ac.dtype, ac.values[rows], None if ac.valid is None else ac.valid[rows])
            scalars.append(evaluate_aggregate(ae.func, sub))
        from query_engine_b200.physical_plan import scalar_arrow_type

        t = scalars[0].type
        arr = pa.array([s.value for s in scalars], type=scalar_arrow_type(t))
        out_arrays.append(arr)
        fields.append(pa.field(f"col_{ai}", arr.type, True))
    return [pa.RecordBatch.from_arrays(out_arrays, schema=pa.schema(fields))]


def _equi_keys(on: PhysicalExpr, n_left: int) -> Optional[List[Tuple[int, int]]]:
    """`on` as Column = Column (or an AND chain of those) across the two sides.
    Column indices address the merged left++right schema (planner.rs:148-157)."""
    if isinstance(on, PhysicalExpr.BinaryExpr) and on.op == BinaryOp.And:
        l = _equi_keys(on.left, n_left)
        r = _equi_keys(on.right, n_left)
        if l is None or r is None:
            return None
        return l + r
    if isinstance(on, PhysicalExpr.BinaryExpr) and on.op == BinaryOp.Equal:
        a, b = on.left, on.right
        if isinstance(a, PhysicalExpr.Column) and isinstance(b, PhysicalExpr.Column):
            if a.index < n_left <= b.index:
                return [(a.index, b.index - n_left)]
            if b.index < n_left <= a.index:
                return [(b.index, a.index - n_left)]
    return None


def _cartesian(left_batches, right_batches, keep_empty: bool) -> List[pa.RecordBatch]:
    """executor.rs:363-399,500-540 join_batches: left index repeated, right cycled"""
    out = []
    for lb in left_batches:
        for rb in right_batches:
            L, R = lb.num_rows, rb.num_rows
            li = pa.array(np.repeat(np.arange(L, dtype=np.uint32), R))
            ri = pa.array(np.tile(np.arange(R, dtype=np.uint32), L))
            cols = [c.take(li) for c in lb.columns] + [c.take(ri) for c in rb.columns]
            schema = pa.schema(list(lb.schema) + list(rb.schema))
            jb = pa.RecordBatch.from_arrays(cols, schema=schema)
            if keep_empty or jb.num_rows > 0:
                out.append(jb)
    return out


def execute_join(
    left_batches: Sequence[pa.RecordBatch],
    right_batches: Sequence[pa.RecordBatch],
    join_type: JoinType,
    on: Optional[PhysicalExpr],
    reference_asis: bool = False,
) -> List[pa.RecordBatch]:
    """executor.rs:343-361 + Tier-2 equi-join completion: INNER / LEFT (docs/SQL_REFERENCE.md:215-229) and, by the same
    documented SQL semantics, RIGHT (every right row; NULL left columns where nothing matched) and FULL (both) - the
    reference's execute_right_join / execute_full_join (executor.rs:401-435) are Cartesian stubs like the others."""
    if not left_batches or not right_batches:
        return []  # :350-352 (even for LEFT)
    n_left = left_batches[0].num_columns
    keys = None if on is None else _equi_keys(on, n_left)
    if reference_asis or keys is None or join_type not in (JoinType.Inner, JoinType.Left, JoinType.Right, JoinType.Full):
        return _cartesian(left_batches, right_batches, keep_empty=(join_type != JoinType.Inner))
    lt = pa.Table.from_batches(list(left_batches)).combine_chunks()
    rt = pa.Table.from_batches(list(right_batches)).combine_chunks()
    lk = [from_arrow(lt.column(i)) for i, _ in keys]
    rk = [from_arrow(rt.column(j)) for _, j in keys]
    for a, b in zip(lk, rk):
        ca, cb = coerce_numeric_types(a, b)
        if ca.dtype != cb.dtype:
            raise ArrowError(f"Invalid argument error: Invalid comparison operation: {a.dtype} == {b.dtype}")
    lk = [coerce_numeric_types(a, b)[0] for a, b in zip(lk, rk)]
    rk = [coerce_numeric_types(a, b)[1] for a, b in zip([from_arrow(lt.column(i)) for i, _ in keys], rk)]
    table: dict = {}
    for j in range(rt.num_rows):
        k = tuple(_group_key_of(c, j) for c in rk)
        if any(x is None for x in k):
            continue  # null keys never match
        table.setdefault(k, []).append(j)
    li: List[Optional[int]] = []
    ri: List[Optional[int]] = []
    matched_right = np.zeros(rt.num_rows, dtype=bool)
    for i in range(lt.num_rows):
        k = tuple(_group_key_of(c, i) for c in lk)
        m = None if any(x is None for x in k) else table.get(k)
        if m:
            for j in m:
                li.append(i)
                ri.append(j)
                matched_right[j] = True
        elif join_type in (JoinType.Left, JoinType.Full):
            li.append(i)
            ri.append(None)
    if join_type in (JoinType.Right, JoinType.Full):
        for j in np.flatnonzero(~matched_right):
            li.append(None)
            ri.append(int(j))
    if not li:
        return []
    lidx = pa.array(li, type=pa.int64())
    ridx = pa.array(ri, type=pa.int64())
    cols = [lt.column(c).combine_chunks().take(lidx) for c in range(lt.num_columns)]
    cols += [rt.column(c).combine_chunks().take(ridx) for c in range(rt.num_columns)]
    cols = [c.combine_chunks() if isinstance(c, pa.ChunkedArray) else c for c in cols]
    rfields = [f.with_nullable(True) if join_type in (JoinType.Left, JoinType.Full) else f for f in rt.schema]
    lfields = [f.with_nullable(True) if join_type in (JoinType.Right, JoinType.Full) else f for f in lt.schema]
    schema = pa.schema(lfields + rfields)
    return [pa.RecordBatch.from_arrays(cols, schema=schema)]


# ---------------------------------------------------------------------------
# plan walker (executor.rs:23-91) for the arms on the hot path
# ---------------------------------------------------------------------------
def execute(plan: PhysicalPlan, reference_asis: bool = False) -> List[pa.RecordBatch]:
    if isinstance(plan, PhysicalPlan.Scan):
        return plan.source.scan()
    if isinstance(plan, PhysicalPlan.Filter):
        return execute_filter(execute(plan.input, reference_asis), plan.predicate)
    if isinstance(plan, PhysicalPlan.HashAggregate):
        return execute_aggregate(execute(plan.input, reference_asis), plan.group_exprs, plan.aggr_exprs, reference_asis)
    if isinstance(plan, PhysicalPlan.HashJoin):
        return execute_join(
            execute(plan.left, reference_asis), execute(plan.right, reference_asis), plan.join_type, plan.on, reference_asis
        )
    if isinstance(plan, PhysicalPlan.Projection):
        out = []
        for b in execute(plan.input, reference_asis):  # executor.rs:93-129
            arrays = [to_arrow(evaluate_expr(e, b)) for e in plan.exprs]
            if arrays:
                out.append(pa.RecordBatch.from_arrays(arrays, schema=plan.schema))
        return out
    if isinstance(plan, PhysicalPlan.Sort):
        return execute(plan.input, reference_asis)  # no-op, executor.rs:290-297
    if isinstance(plan, PhysicalPlan.SubqueryScan):
        return execute(plan.subquery, reference_asis)
    if isinstance(plan, PhysicalPlan.Limit):
        # executor.rs:299-341
        out, total, rem = [], 0, plan.skip
        for b in execute(plan.input, reference_asis):
            if rem >= b.num_rows:
                rem -= b.num_rows
                continue
            start, rem = rem, 0
            end = b.num_rows if plan.fetch is None else min(start + max(plan.fetch - total, 0), b.num_rows)
            if start < end:
                s = b.slice(start, end - start)
                total += s.num_rows
                out.append(s)
                if plan.fetch is not None and total >= plan.fetch:
                    break
        return out
    raise TypeError(plan)


# ---------------------------------------------------------------------------
# exchange: Partitioner::partition_by_hash (query-distributed/src/partition.rs:151-212,292-316)
# ---------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _rotl(x, b):
    return ((x << np.uint64(b)) | (x >> np.uint64(64 - b)))


def siphash_stream(words: Sequence[np.ndarray], nbytes_each: Sequence[int], c_rounds=1, d_rounds=3, k0=0, k1=0) -> np.ndarray:
    """Vectorised SipHash-c-d over rows.  `words[i]` is an array (one element per
    row) of the little-endian integer written by the i-th `Hasher::write` call,
    `nbytes_each[i]` its width (8 for i64, 4 for i32).  Rust's DefaultHasher is
    SipHash-1-3 with k0 = k1 = 0 over the concatenated byte stream."""
    n = len(words[0]) if words else 0
    old = np.seterr(over="ignore")
    try:
        v0 = np.full(n, k0 ^ 0x736F6D6570736575, dtype=np.uint64)
        v1 = np.full(n, k1 ^ 0x646F72616E646F6D, dtype=np.uint64)
        v2 = np.full(n, k0 ^ 0x6C7967656E657261, dtype=np.uint64)
        v3 = np.full(n, k1 ^ 0x7465646279746573, dtype=np.uint64)

        def rounds(v0, v1, v2, v3, k):
            for _ in range(k):
                v0 = v0 + v1
                v1 = _rotl(v1, 13)
                v1 ^= v0
                v0 = _rotl(v0, 32)
                v2 = v2 + v3
                v3 = _rotl(v3, 16)
                v3 ^= v2
                v0 = v0 + v3
                v3 = _rotl(v3, 21)
                v3 ^= v0
                v2 = v2 + v1
                v1 = _rotl(v1, 17)
                v1 ^= v2
                v2 = _rotl(v2, 32)
            return v0, v1, v2, v3

        # assemble the byte stream into 64-bit little-endian message words
        tail = np.zeros(n, dtype=np.uint64)
        ntail = 0
        total = 0
        for w, nb in zip(words, nbytes_each):
            w = w.astype(np.uint64) & np.uint64((1 << (8 * nb)) - 1 if nb < 8 else _M64)
            total += nb
            tail = tail | (w << np.uint64(8 * ntail)) if ntail < 8 else tail
            if ntail + nb >= 8:
                m = tail
                v3 ^= m
                v0, v1, v2, v3 = rounds(v0, v1, v2, v3, c_rounds)
                v0 ^= m
                used = 8 - ntail
                rest = nb - used
                tail = (w >> np.uint64(8 * used)) if (rest > 0 and used < 8) else np.zeros(n, dtype=np.uint64)
                ntail = rest
            else:
                ntail += nb
        b = tail | np.uint64((total & 0xFF) << 56)
        v3 ^= b
        v0, v1, v2, v3 = rounds(v0, v1, v2, v3, c_rounds)
        v0 ^= b
        v2 ^= np.uint64(0xFF)
        v0, v1, v2, v3 = rounds(v0, v1, v2, v3, d_rounds)
        return v0 ^ v1 ^ v2 ^ v3
    finally:
        np.seterr(**old)


def siphash_bytes(data: bytes, c_rounds=2, d_rounds=4, k0=0, k1=0) -> int:
    """Scalar SipHash over a byte string (used to pin the skeleton against the
    official SipHash-2-4 vectors, and for Utf8 keys)."""

    def rotl(x, b):
        return ((x << b) | (x >> (64 - b))) & _M64

    v0 = k0 ^ 0x736F6D6570736575
    v1 = k1 ^ 0x646F72616E646F6D
    v2 = k0 ^ 0x6C7967656E657261
    v3 = k1 ^ 0x7465646279746573

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & _M64
        v1 = rotl(v1, 13) ^ v0
        v0 = rotl(v0, 32)
        v2 = (v2 + v3) & _M64
        v3 = rotl(v3, 16) ^ v2
        v0 = (v0 + v3) & _M64
        v3 = rotl(v3, 21) ^ v0
        v2 = (v2 + v1) & _M64
        v1 = rotl(v1, 17) ^ v2
        v2 = rotl(v2, 32)
        return v0, v1, v2, v3

    n = len(data)
    for off in range(0, n - n % 8, 8):
        m = int.from_bytes(data[off : off + 8], "little")
        v3 ^= m
        for _ in range(c_rounds):
            v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    b = (n & 0xFF) << 56 | int.from_bytes(data[n - n % 8 :], "little")
    v3 ^= b
    for _ in range(c_rounds):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= b
    v2 ^= 0xFF
    for _ in range(d_rounds):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return v0 ^ v1 ^ v2 ^ v3


def compute_row_hash(batch: pa.RecordBatch, col_indices: Sequence[int]) -> np.ndarray:
    """partition.rs:292-316: DefaultHasher over the non-null Int64 / Utf8 / Int32
    key values of the row; other column types contribute nothing."""
    n = batch.num_rows
    cols = [from_arrow(batch.column(i)) for i in col_indices]
    simple = all(c.dtype in ("i64", "i32") and c.valid is None for c in cols)
    if simple:
        words = [c.values.astype(np.int64).view(np.uint64) if c.dtype == "i64" else c.values.astype(np.int32).view(np.uint32).astype(np.uint64) for c in cols]
        return siphash_stream(words, [8 if c.dtype == "i64" else 4 for c in cols])
    out = np.zeros(n, dtype=np.uint64)
    for r in range(n):
        buf = b""
        for c in cols:
            if c.dtype == "null" or (c.valid is not None and not c.valid[r]):
                continue
            if c.dtype == "i64":
                buf += int(c.values[r]).to_bytes(8, "little", signed=True)
            elif c.dtype == "i32":
                buf += int(c.values[r]).to_bytes(4, "little", signed=True)
            elif c.dtype == "utf8":
                buf += c.values[r].encode() + b"\xff"  # `str::hash` = bytes + 0xff
        out[r] = siphash_bytes(buf, 1, 3)
    return out


def partition_by_hash(
    batches: Sequence[pa.RecordBatch], key_columns: Sequence[str], num_partitions: int
) -> List[List[pa.RecordBatch]]:
    """partition.rs:151-212.  Returns partitions[p] = list of batches (one per
    non-empty (input batch, partition) pair, input order, row order kept)."""
    parts: List[List[pa.RecordBatch]] = [[] for _ in range(num_partitions)]
    for batch in batches:
        names = batch.schema.names
        idx = [names.index(k) for k in key_columns if k in names]
        if not idx:
            raise QueryError("PartitionError", "No key columns found in batch")
        h = compute_row_hash(batch, idx)
        pidx = (h % np.uint64(num_partitions)).astype(np.int64)
        for p in range(num_partitions):
            rows = np.flatnonzero(pidx == p)
            if len(rows) == 0:
                continue
            parts[p].append(batch.take(pa.array(rows.astype(np.uint32))))
    return parts


def concat_merge(partitions: Sequence[Sequence[pa.RecordBatch]]) -> List[pa.RecordBatch]:
    """query-distributed/src/operators.rs:136-138"""
    return [b for p in partitions for b in p]


# ---------------------------------------------------------------------------
# comparison helpers used by the parity tests
# ---------------------------------------------------------------------------
def batches_to_table(batches: Sequence[pa.RecordBatch]) -> Optional[pa.Table]:
    if not batches:
        return None
    return pa.Table.from_batches(list(batches))
