"""Host-side mirror of the reference's physical-plan operator API.

This is the drop-in boundary of the hot path: the same descriptor types the
reference's planner lowering produces and `QueryExecutor::execute` consumes
(reference: crates/query-executor/src/physical_plan.rs:8-157,
crates/query-planner/src/logical_plan.rs:147-161 for ScalarValue,
crates/query-parser/src/ast.rs:101-107 for JoinType).  Only the variants that
sit on the accelerated path are given bodies; the others exist so that a plan
built by the reference's lowering can be represented and rejected with the
reference's own error text.

Nothing in this file touches CUDA; it is plain descriptors.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import pyarrow as pa


class QueryError(Exception):
    """Mirror of query_core::QueryError (crates/query-core/src/error.rs:4-50).

    `kind` is the Rust enum variant name; operator failures are
    ExecutionError, arrow kernel failures are ArrowError.
    """

    def __init__(self, kind: str, message: str):
        super().__init__(f"{kind}: {message}")
        self.kind = kind
        self.message = message


def ExecutionError(msg: str) -> QueryError:  # noqa: N802 - mirrors the Rust variant
    return QueryError("ExecutionError", msg)


def ArrowError(msg: str) -> QueryError:  # noqa: N802
    return QueryError("ArrowError", msg)


# --------------------------------------------------------------------------
# ScalarValue  (query-planner/src/logical_plan.rs:147-161)
# --------------------------------------------------------------------------
class ScalarType(enum.IntEnum):
    """Type tag of a ScalarValue; numbering is the C-ABI's qgpu_dtype."""

    Null = 0
    Boolean = 1
    Int32 = 2
    Int64 = 3
    Float32 = 4
    Float64 = 5
    Utf8 = 6
    Int8 = 7
    Int16 = 8
    UInt8 = 9
    UInt16 = 10
    UInt32 = 11
    UInt64 = 12


@dataclass(frozen=True)
class ScalarValue:
    """`ScalarValue::Int64(Some(5))` is `ScalarValue.Int64(5)`; `Int64(None)` is
    `ScalarValue.Int64(None)`; `ScalarValue::Null` is `ScalarValue.Null()`."""

    type: ScalarType
    value: Any = None

    # constructors named after the Rust variants
    @staticmethod
    def Null() -> "ScalarValue":  # noqa: N802
        return ScalarValue(ScalarType.Null, None)

    @staticmethod
    def Boolean(v: Optional[bool]) -> "ScalarValue":  # noqa: N802
        return ScalarValue(ScalarType.Boolean, v)

    @staticmethod
    def Int8(v):  # noqa: N802
        return ScalarValue(ScalarType.Int8, v)

    @staticmethod
    def Int16(v):  # noqa: N802
        return ScalarValue(ScalarType.Int16, v)

    @staticmethod
    def Int32(v: Optional[int]) -> "ScalarValue":  # noqa: N802
        return ScalarValue(ScalarType.Int32, v)

    @staticmethod
    def Int64(v: Optional[int]) -> "ScalarValue":  # noqa: N802
        return ScalarValue(ScalarType.Int64, v)

    @staticmethod
    def UInt8(v):  # noqa: N802
        return ScalarValue(ScalarType.UInt8, v)

    @staticmethod
    def UInt16(v):  # noqa: N802
        return ScalarValue(ScalarType.UInt16, v)

    @staticmethod
    def UInt32(v):  # noqa: N802
        return ScalarValue(ScalarType.UInt32, v)

    @staticmethod
    def UInt64(v):  # noqa: N802
        return ScalarValue(ScalarType.UInt64, v)

    @staticmethod
    def Float32(v: Optional[float]) -> "ScalarValue":  # noqa: N802
        return ScalarValue(ScalarType.Float32, v)

    @staticmethod
    def Float64(v: Optional[float]) -> "ScalarValue":  # noqa: N802
        return ScalarValue(ScalarType.Float64, v)

    @staticmethod
    def Utf8(v: Optional[str]) -> "ScalarValue":  # noqa: N802
        return ScalarValue(ScalarType.Utf8, v)

    @staticmethod
    def from_sql_number(text: str) -> "ScalarValue":
        """Literal typing of the planner: a '.' makes Float64, otherwise Int64
        (crates/query-planner/src/planner.rs:669-686)."""
        if "." in text:
            return ScalarValue.Float64(float(text))
        return ScalarValue.Int64(int(text))


_ARROW_OF_SCALAR = {
    ScalarType.Null: pa.null(),
    ScalarType.Boolean: pa.bool_(),
    ScalarType.Int8: pa.int8(),
    ScalarType.Int16: pa.int16(),
    ScalarType.Int32: pa.int32(),
    ScalarType.Int64: pa.int64(),
    ScalarType.UInt8: pa.uint8(),
    ScalarType.UInt16: pa.uint16(),
    ScalarType.UInt32: pa.uint32(),
    ScalarType.UInt64: pa.uint64(),
    ScalarType.Float32: pa.float32(),
    ScalarType.Float64: pa.float64(),
    ScalarType.Utf8: pa.string(),
}


def scalar_arrow_type(t: ScalarType) -> pa.DataType:
    return _ARROW_OF_SCALAR[t]


# --------------------------------------------------------------------------
# operators / functions (physical_plan.rs:119-157)
# --------------------------------------------------------------------------
class BinaryOp(enum.IntEnum):
    Add = 0
    Subtract = 1
    Multiply = 2
    Divide = 3
    Modulo = 4
    Equal = 5
    NotEqual = 6
    Less = 7
    LessEqual = 8
    Greater = 9
    GreaterEqual = 10
    And = 11
    Or = 12
    TsMatch = 13  # full-text search: string path, not accelerated


class UnaryOp(enum.IntEnum):
    Not = 0
    Minus = 1
    # extension, not in the reference's enum: `x as f64`; only the multi-GPU final-aggregate plan emits it
    CastFloat64 = 2


class AggregateFunction(enum.IntEnum):
    Count = 0
    Sum = 1
    Avg = 2
    Min = 3
    Max = 4


class JoinType(enum.IntEnum):
    """query_parser::JoinType (crates/query-parser/src/ast.rs:101-107)."""

    Inner = 0
    Left = 1
    Right = 2
    Full = 3
    Cross = 4


# --------------------------------------------------------------------------
# PhysicalExpr (physical_plan.rs:74-117)
# --------------------------------------------------------------------------
class PhysicalExpr:
    """Namespace + base class; variants are the nested dataclasses."""

    # populated below


@dataclass(frozen=True)
class _Column(PhysicalExpr):
    name: str
    index: int


@dataclass(frozen=True)
class _Literal(PhysicalExpr):
    value: ScalarValue


@dataclass(frozen=True)
class _BinaryExpr(PhysicalExpr):
    left: PhysicalExpr
    op: BinaryOp
    right: PhysicalExpr


@dataclass(frozen=True)
class _UnaryExpr(PhysicalExpr):
    op: UnaryOp
    expr: PhysicalExpr


@dataclass(frozen=True)
class _Unsupported(PhysicalExpr):
    """ScalarSubquery / InSubquery / Exists / WindowFunction / ScalarFunction:
    represented so the reference's error text can be reproduced
    (operators.rs:34-60)."""

    variant: str


PhysicalExpr.Column = _Column
PhysicalExpr.Literal = _Literal
PhysicalExpr.BinaryExpr = _BinaryExpr
PhysicalExpr.UnaryExpr = _UnaryExpr
PhysicalExpr.Unsupported = _Unsupported

UNSUPPORTED_EXPR_MESSAGES = {
    "ScalarSubquery": "Scalar subquery execution not yet implemented in synchronous context",
    "InSubquery": "IN subquery execution not yet implemented in synchronous context",
    "Exists": "EXISTS subquery execution not yet implemented in synchronous context",
    "WindowFunction": "Window function execution not yet implemented",
}


def col(name: str, index: int) -> PhysicalExpr:
    return PhysicalExpr.Column(name, index)


def lit(v) -> PhysicalExpr:
    """Literal typed the way the planner types SQL literals."""
    if isinstance(v, ScalarValue):
        return PhysicalExpr.Literal(v)
    if isinstance(v, bool):
        return PhysicalExpr.Literal(ScalarValue.Boolean(v))
    if isinstance(v, int):
        return PhysicalExpr.Literal(ScalarValue.Int64(v))
    if isinstance(v, float):
        return PhysicalExpr.Literal(ScalarValue.Float64(v))
    if isinstance(v, str):
        return PhysicalExpr.Literal(ScalarValue.Utf8(v))
    if v is None:
        return PhysicalExpr.Literal(ScalarValue.Null())
    raise TypeError(f"cannot make a literal from {type(v)}")


def binary(left: PhysicalExpr, op: BinaryOp, right: PhysicalExpr) -> PhysicalExpr:
    return PhysicalExpr.BinaryExpr(left, op, right)


def and_all(terms: Sequence[PhysicalExpr]) -> PhysicalExpr:
    """Left-deep AND chain, the shape the parser produces for `a AND b AND c`."""
    out = terms[0]
    for t in terms[1:]:
        out = PhysicalExpr.BinaryExpr(out, BinaryOp.And, t)
    return out


@dataclass(frozen=True)
class AggregateExpr:
    """physical_plan.rs:144-148"""

    func: AggregateFunction
    expr: PhysicalExpr


# --------------------------------------------------------------------------
# DataSource + PhysicalPlan (physical_plan.rs:8-72)
# --------------------------------------------------------------------------
class DataSource:
    """`trait DataSource { fn scan(&self) -> Result<Vec<RecordBatch>>; fn schema(&self) -> &Schema; }`"""

    def scan(self) -> List[pa.RecordBatch]:
        raise NotImplementedError

    def schema(self) -> pa.Schema:
        raise NotImplementedError


class MemoryDataSource(DataSource):
    """Mirror of query-storage MemoryDataSource (crates/query-storage/src/memory.rs:16,305-313):
    returns exactly the batches it was given."""

    def __init__(self, schema: pa.Schema, batches: Sequence[pa.RecordBatch]):
        self._schema = schema
        self._batches = list(batches)

    def scan(self) -> List[pa.RecordBatch]:
        return list(self._batches)

    def schema(self) -> pa.Schema:
        return self._schema


class PhysicalPlan:
    """Namespace + base class; variants are the nested dataclasses."""


@dataclass(frozen=True)
class _Scan(PhysicalPlan):
    source: DataSource
    schema: Optional[pa.Schema] = None


@dataclass(frozen=True)
class _Projection(PhysicalPlan):
    input: PhysicalPlan
    exprs: Sequence[PhysicalExpr]
    schema: pa.Schema


@dataclass(frozen=True)
class _Filter(PhysicalPlan):
    input: PhysicalPlan
    predicate: PhysicalExpr


@dataclass(frozen=True)
class _HashJoin(PhysicalPlan):
    left: PhysicalPlan
    right: PhysicalPlan
    join_type: JoinType
    on: Optional[PhysicalExpr] = None


@dataclass(frozen=True)
class _HashAggregate(PhysicalPlan):
    input: PhysicalPlan
    group_exprs: Sequence[PhysicalExpr]
    aggr_exprs: Sequence[AggregateExpr]


@dataclass(frozen=True)
class _Sort(PhysicalPlan):
    input: PhysicalPlan
    exprs: Sequence[PhysicalExpr] = field(default_factory=tuple)
    ascending: Sequence[bool] = field(default_factory=tuple)


@dataclass(frozen=True)
class _Limit(PhysicalPlan):
    input: PhysicalPlan
    skip: int = 0
    fetch: Optional[int] = None


@dataclass(frozen=True)
class _SubqueryScan(PhysicalPlan):
    subquery: PhysicalPlan
    schema: Optional[pa.Schema] = None


PhysicalPlan.Scan = _Scan
PhysicalPlan.Projection = _Projection
PhysicalPlan.Filter = _Filter
PhysicalPlan.HashJoin = _HashJoin
PhysicalPlan.HashAggregate = _HashAggregate
PhysicalPlan.Sort = _Sort
PhysicalPlan.Limit = _Limit
PhysicalPlan.SubqueryScan = _SubqueryScan
