"""query_engine_b200 — B200-native (sm_100a) filter / hash-join / hash-aggregate / exchange operators behind
query-engine's physical-plan API.  The product is `lib/libqgpu.so` (C ABI in include/qgpu.h); this package is the
host-side mirror of the reference's operator interface that drives it.  There is no CPU fallback."""
from .physical_plan import (  # noqa: F401
    AggregateExpr,
    AggregateFunction,
    ArrowError,
    BinaryOp,
    DataSource,
    ExecutionError,
    JoinType,
    MemoryDataSource,
    PhysicalExpr,
    PhysicalPlan,
    QueryError,
    ScalarType,
    ScalarValue,
    UnaryOp,
    and_all,
    binary,
    col,
    lit,
)

__all__ = [n for n in dir() if not n.startswith("_")]
