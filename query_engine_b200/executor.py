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
from .physical_plan import AggregateExpr, AggregateFunction, DataSource, ExecutionError, JoinType, PhysicalExpr, PhysicalPlan, lit


def _column_refs(e: Optional[PhysicalExpr], out: set) -> None:
    if e is None:
        return
    if isinstance(e, PhysicalExpr.Column):
        out.add(e.index)
    elif isinstance(e, PhysicalExpr.BinaryExpr):
        _column_refs(e.left, out)
        _column_refs(e.right, out)
    elif isinstance(e, PhysicalExpr.UnaryExpr):
        _column_refs(e.expr, out)


def _remap_columns(e: Optional[PhysicalExpr], remap: dict) -> Optional[PhysicalExpr]:
    if e is None:
        return None
    if isinstance(e, PhysicalExpr.Column):
        return PhysicalExpr.Column(e.name, remap[e.index])
    if isinstance(e, PhysicalExpr.BinaryExpr):
        return PhysicalExpr.BinaryExpr(_remap_columns(e.left, remap), e.op, _remap_columns(e.right, remap))
    if isinstance(e, PhysicalExpr.UnaryExpr):
        return PhysicalExpr.UnaryExpr(e.op, _remap_columns(e.expr, remap))
    return e


def prune_for_aggregate(batches: Sequence[pa.RecordBatch], group_exprs, aggr_exprs, predicate):
    """The projection the reference's planner never pushes under an aggregate (its `ProjectionPushdown` is a no-op,
    query-planner/src/optimizer.rs:62-69): keep only the columns the predicate / group keys / aggregate inputs read -
    host -> device bytes are the bound of the end-to-end path.  Column indices are remapped onto the kept columns;
    COUNT(col) over a column without nulls needs no data at all and becomes COUNT(1) (operators.rs:747-750:
    len - null_count).  Anything unusual (an out-of-range index, which must raise the reference's error) keeps every
    column.  Returns (batches, group_exprs, aggr_exprs, predicate) to run instead."""
    ncols = batches[0].num_columns
    refs: set = set()
    _column_refs(predicate, refs)
    for g in group_exprs:
        _column_refs(g, refs)
    aggs = list(aggr_exprs)
    for i, a in enumerate(aggs):
        e = a.expr
        if (a.func == AggregateFunction.Count and isinstance(e, PhysicalExpr.Column) and 0 <= e.index < ncols
                and e.index not in refs and all(b.column(e.index).null_count == 0 for b in batches)):
            aggs[i] = AggregateExpr(a.func, lit(1))
        else:
            _column_refs(e, refs)
    if any(i < 0 or i >= ncols for i in refs) or len(refs) >= ncols:
        return list(batches), list(group_exprs), list(aggr_exprs), predicate
    keep = sorted(refs) or [0]  # (at least one column carries the row count)
    remap = {old: new for new, old in enumerate(keep)}
    pruned = [b.select(keep) for b in batches]
    return (pruned, [_remap_columns(g, remap) for g in group_exprs],
            [AggregateExpr(a.func, _remap_columns(a.expr, remap)) for a in aggs], _remap_columns(predicate, remap))


class GpuDataSource(DataSource):
    """A `DataSource` whose batches already live in HBM.  `scan()` still returns host
    batches for any CPU consumer; `QueryExecutor` recognises the type and uses the resident table directly."""

    def __init__(self, ctx: _ffi.Context, batches: Sequence[pa.RecordBatch], schema: Optional[pa.Schema] = None):
        self._schema = schema if schema is not None else batches[0].schema
        self.ctx = ctx
        self.table = ctx.import_batches(list(batches), self._schema)

    @classmethod
    def from_table(cls, ctx: _ffi.Context, table: _ffi.Table, schema: pa.Schema) -> "GpuDataSource":
        src = cls.__new__(cls)
        src._schema, src.ctx, src.table = schema, ctx, table
        return src

    def scan(self) -> List[pa.RecordBatch]:
        return self.table.to_batches()

    def schema(self) -> pa.Schema:
        return self._schema


class GpuCatalog:
    """Device-resident tables by name: what `MemoryDataSource` plus the front-ends' table registry
    are to the reference (crates/query-storage/src/memory.rs:16-43, 305-313; CSV / Parquet registration
    crates/query-storage/src/csv.rs:24-38, query-cli `.load`).  A table is parsed on the host once (pyarrow's readers:
    storage I/O is out of scope), uploaded once - from pinned memory - and every query over it starts in HBM.
    `source(name)` gives the `DataSource` a `PhysicalPlan.Scan` takes."""

    def __init__(self, ctx: _ffi.Context):
        self.ctx = ctx
        self._schemas = {}

    def register_batches(self, name: str, batches: Sequence[pa.RecordBatch], schema: Optional[pa.Schema] = None) -> "GpuDataSource":
        src = GpuDataSource(self.ctx, list(batches), schema)
        self.ctx.catalog_put(name, src.table)
        self._schemas[name] = src.schema()
        return src

    def register_csv(self, name: str, path: str, **read_options) -> "GpuDataSource":
        import pyarrow.csv as pacsv

        # arrow-csv of the reference treats only the empty string as null
        conv = pacsv.ConvertOptions(null_values=[""], strings_can_be_null=True, **read_options)
        return self.register_batches(name, pacsv.read_csv(path, convert_options=conv).combine_chunks().to_batches())

    def register_parquet(self, name: str, path: str, columns: Optional[Sequence[str]] = None) -> "GpuDataSource":
        import pyarrow.parquet as pq

        return self.register_batches(name, pq.read_table(path, columns=columns).combine_chunks().to_batches())

    def source(self, name: str) -> "GpuDataSource":
        t = self.ctx.catalog_get(name)
        if t is None:
            raise ExecutionError(f"Table '{name}' not found")
        return GpuDataSource.from_table(self.ctx, t, self._schemas[name])

    def drop(self, name: str):
        self.ctx.catalog_drop(name)
        self._schemas.pop(name, None)


class QueryExecutor:
    """Stateless like the reference's unit struct, apart from owning the device context."""

    def __init__(self, device: int = 0, ctx: Optional[_ffi.Context] = None, tier2_sort: bool = False):
        self.ctx = ctx if ctx is not None else _ffi.Context(device)
        self.h2d_bytes = 0  # bytes of host batches handed to the library for upload so far
        # The reference's execute_sort is a pass-through (executor.rs:290-297) and that is the default here too.
        # tier2_sort=True runs the documented ORDER BY instead: a device lexsort; directly below a
        # Limit it becomes a Top-K.  Sort expressions must be plain columns; NULLs sort last ascending, first descending
        # (the default of arrow's SortOptions the reference would get).
        self.tier2_sort = tier2_sort

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
            child, pred = plan.input, None
            if isinstance(child, PhysicalPlan.Filter):
                child, pred = child.input, child.predicate
            if isinstance(child, PhysicalPlan.Scan) and not isinstance(child.source, GpuDataSource) and plan.aggr_exprs:
                return self._aggregate_over_host_scan(child.source, plan.group_exprs, plan.aggr_exprs, pred)
            inp = self._execute_plan(child)
            return self.execute_aggregate(inp, plan.group_exprs, plan.aggr_exprs, predicate=pred)
        if isinstance(plan, PhysicalPlan.HashJoin):
            left = self._execute_plan(plan.left)
            right = self._execute_plan(plan.right)
            return self.execute_join(left, right, plan.join_type, plan.on)
        if isinstance(plan, PhysicalPlan.Projection):
            inp = self._execute_plan(plan.input)
            return self.execute_projection(inp, plan.exprs, plan.schema)
        if isinstance(plan, PhysicalPlan.Sort):
            inp = self._execute_plan(plan.input)
            if not self.tier2_sort or inp is None:
                return inp  # the reference's execute_sort is a pass-through (executor.rs:290-297)
            return self.execute_sort(inp, plan.exprs, plan.ascending)
        if isinstance(plan, PhysicalPlan.SubqueryScan):
            return self._execute_plan(plan.subquery)  # executor.rs:71-74
        if isinstance(plan, PhysicalPlan.Limit):
            if self.tier2_sort and isinstance(plan.input, PhysicalPlan.Sort) and plan.fetch is not None:
                inp = self._execute_plan(plan.input.input)  # ORDER BY .. LIMIT: Top-K, only skip + fetch rows are gathered
                if inp is None:
                    return None
                top = self.execute_sort(inp, plan.input.exprs, plan.input.ascending, fetch=plan.skip + plan.fetch)
                return self.execute_limit(top, plan.skip, plan.fetch)
            inp = self._execute_plan(plan.input)
            return self.execute_limit(inp, plan.skip, plan.fetch)
        raise ExecutionError(f"unknown plan node {type(plan).__name__}")

    def _scan(self, source: DataSource) -> Optional[_ffi.Table]:
        if isinstance(source, GpuDataSource):
            return source.table
        batches = source.scan()
        if not batches:
            return None
        self.h2d_bytes += sum(b.nbytes for b in batches)
        return self.ctx.import_batches(batches, batches[0].schema)

    def _aggregate_over_host_scan(self, source: DataSource, group_exprs, aggr_exprs, predicate) -> Optional[_ffi.Table]:
        """HashAggregate [over Filter] over a Scan of HOST batches: only the columns the expressions read are uploaded
        (`prune_for_aggregate`)."""
        batches = source.scan()
        if not batches:
            return None
        pruned, groups, aggs, pred = prune_for_aggregate(batches, group_exprs, aggr_exprs, predicate)
        self.h2d_bytes += sum(b.nbytes for b in pruned)
        inp = self.ctx.import_batches(pruned, pruned[0].schema)
        return self.execute_aggregate(inp, groups, aggs, predicate=pred)

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

    def execute_sort(self, batches: _ffi.Table, exprs, ascending, fetch: Optional[int] = None) -> Optional[_ffi.Table]:
        """Tier-2 ORDER BY (the reference's execute_sort, executor.rs:290-297, returns its input)"""
        names = []
        for e in exprs:
            if not isinstance(e, PhysicalExpr.Column):
                raise _ffi.UnsupportedOnGpu("ORDER BY on a computed expression is not on the accelerated path")
            names.append(e.name)
        asc = list(ascending) + [True] * (len(names) - len(ascending))
        return self.ctx.sort(batches, names, asc, [not a for a in asc], fetch)

    def execute_limit(self, batches: Optional[_ffi.Table], skip: int, fetch: Optional[int]) -> Optional[_ffi.Table]:
        """executor.rs:299-341: skip `skip` rows, keep at most `fetch`; output batches are slices of the input batches"""
        if batches is None:
            return None
        return self.ctx.execute_limit(batches, skip, fetch)

    def execute_projection(self, batches, exprs, schema: pa.Schema) -> Optional[_ffi.Table]:
        """executor.rs:93-129; output fields are named by `schema`"""
        if batches is None or not exprs:
            return None
        names = [f.name for f in schema] if schema is not None else [f"col_{i}" for i in range(len(exprs))]
        return self.ctx.execute_projection(batches, list(exprs), names)
