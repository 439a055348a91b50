"""`QueryExecutor` — the reference's executor API over libqgpu.

Mirrors `QueryExecutor::execute(&PhysicalPlan) -> Result<Vec<RecordBatch>>`
(reference: crates/query-executor/src/executor.rs:12-91).  The recursion is the reference's (child first, full
materialisation between operators) except that the materialised intermediate is a device-resident table and that
a `Filter` directly below a `HashAggregate` is fused into the aggregate's scan (one pass over the columns, no
compacted intermediate) — the result is what the unfused plan gives, including the reference's
"no surviving batch => Ok(vec![])" rule.

All compute happens in libqgpu's CUDA kernels.  There is no CPU path here: without the library or a GPU,
constructing the executor raises.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import pyarrow as pa

from . import _ffi
from .physical_plan import DataSource, ExecutionError, JoinType, PhysicalExpr, PhysicalPlan


class GpuDataSource(DataSource):
    """A `DataSource` whose batches already live in HBM.  `scan()` still returns host
    batches for any CPU consumer; `QueryExecutor` recognises the type and uses the resident table directly."""

    def __init__(self, ctx: _ffi.Context, batches: Sequence[pa.RecordBatch], schema: Optional[pa.Schema] = None):
        self._schema = schema if schema is not None else batches[0].schema
        self.ctx = ctx
        self.table = ctx.import_batches(list(batches), self._schema)

    def scan(self) -> List[pa.RecordBatch]:
        return self.table.to_batches()

    def schema(self) -> pa.Schema:
        return self._schema


class QueryExecutor:
    """Stateless like the reference's unit struct, apart from owning the device context."""

    def __init__(self, device: int = 0, ctx: Optional[_ffi.Context] = None):
        self.ctx = ctx if ctx is not None else _ffi.Context(device)

    # -- public API ---------------------------------------------------------------------------
    def execute(self, plan: PhysicalPlan) -> List[pa.RecordBatch]:
        t = self.execute_device(plan)
        return [] if t is None else t.to_batches()

    def execute_device(self, plan: PhysicalPlan) -> Optional[_ffi.Table]:
        """Same as `execute` but leaves the result in HBM (None == `Ok(vec![])`)."""
        return self._execute_plan(plan)

    # -- recursion (executor.rs:23-91) ---------------------------------------------------------
    def _execute_plan(self, plan: PhysicalPlan) -> Optional[_ffi.Table]:
        if isinstance(plan, PhysicalPlan.Scan):
            return self._scan(plan.source)
        if isinstance(plan, PhysicalPlan.Filter):
            inp = self._execute_plan(plan.input)
            return self.execute_filter(inp, plan.predicate)
        if isinstance(plan, PhysicalPlan.HashAggregate):
            child = plan.input
            if isinstance(child, PhysicalPlan.Filter):
                inp = self._execute_plan(child.input)
                return self.execute_aggregate(inp, plan.group_exprs, plan.aggr_exprs, predicate=child.predicate)
            inp = self._execute_plan(child)
            return self.execute_aggregate(inp, plan.group_exprs, plan.aggr_exprs)
        if isinstance(plan, PhysicalPlan.HashJoin):
            left = self._execute_plan(plan.left)
            right = self._execute_plan(plan.right)
            return self.execute_join(left, right, plan.join_type, plan.on)
        if isinstance(plan, PhysicalPlan.Projection):
            inp = self._execute_plan(plan.input)
            return self.execute_projection(inp, plan.exprs, plan.schema)
        if isinstance(plan, PhysicalPlan.Sort):
            return self._execute_plan(plan.input)  # the reference's execute_sort is a pass-through (executor.rs:290-297)
        if isinstance(plan, PhysicalPlan.SubqueryScan):
            return self._execute_plan(plan.subquery)  # executor.rs:71-74
        if isinstance(plan, PhysicalPlan.Limit):
            raise _ffi.UnsupportedOnGpu("Limit is not on the accelerated path (SURVEY.md 8f rank 3)")
        raise ExecutionError(f"unknown plan node {type(plan).__name__}")

    def _scan(self, source: DataSource) -> Optional[_ffi.Table]:
        if isinstance(source, GpuDataSource):
            return source.table
        batches = source.scan()
        if not batches:
            return None
        return self.ctx.import_batches(batches, batches[0].schema)

    # -- the three seams (same argument meaning as the Rust private methods) ---------------------
    def execute_filter(self, batches: Optional[_ffi.Table], predicate: PhysicalExpr) -> Optional[_ffi.Table]:
        """executor.rs:131-155"""
        if batches is None:
            return None
        return self.ctx.execute_filter(batches, predicate)

    def execute_aggregate(self, batches, group_exprs, aggr_exprs, predicate=None) -> Optional[_ffi.Table]:
        """executor.rs:157-190 (+ the Filter below it when `predicate` is given)"""
        if not aggr_exprs:
            return None  # :163-165
        if batches is None:
            return None
        return self.ctx.execute_aggregate(batches, list(group_exprs), list(aggr_exprs), predicate)

    def execute_join(self, left, right, join_type: JoinType, on: Optional[PhysicalExpr]) -> Optional[_ffi.Table]:
        """executor.rs:343-361"""
        if left is None or right is None:
            return None  # :350-352, even for LEFT
        return self.ctx.execute_join(left, right, int(join_type), on)

    def execute_projection(self, batches, exprs, schema: pa.Schema) -> Optional[_ffi.Table]:
        """executor.rs:93-129; output fields are named by `schema`"""
        if batches is None or not exprs:
            return None
        names = [f.name for f in schema] if schema is not None else [f"col_{i}" for i in range(len(exprs))]
        return self.ctx.execute_projection(batches, list(exprs), names)
