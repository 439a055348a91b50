"""In-box multi-GPU execution: one process per GPU, NCCL over NVLink 5 / NVSwitch for the exchanges.

Mirrors the stage topology the reference's `DistributedPlanner` describes (it never executes it —
crates/query-distributed/src/planner.rs:200-249, executor.rs:212-263):

  Aggregate  ->  PartialAggregate on every rank's row shard  ->  gather  ->  FinalAggregate        (planner.rs:200-226)
                 (many groups: the library exchanges ROWS by key over NVLink instead and every rank returns its shard
                  of the groups - the result is then partitioned by key, concatenation over ranks finishes it)
  Join       ->  hash-shuffle BOTH inputs on the join key    ->  local HashJoin on every rank       (planner.rs:228-249)
                 (one library call when eligible: the radix join's scatter writes into the owning GPU over NVLink)
  Filter / Projection -> stay on the rank's shard; the result is the concatenation over ranks
                 (Merge::Concat, query-distributed/src/operators.rs:136-138)

The exchange itself (`Exchange::execute`, operators.rs:58) is `Context.shuffle_hash` / `Context.allgather`: device
partitioning kernels + grouped ncclSend/ncclRecv in libqgpu (csrc/shuffle.cu).  This module only rewrites plans; it is
host logic and is covered by world_size-2 gloo tests that need no GPU (tests/test_host_logic.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import pyarrow as pa

from . import _ffi
from .executor import GpuDataSource, QueryExecutor, prune_for_aggregate
from .physical_plan import (
    AggregateExpr,
    AggregateFunction as F,
    BinaryOp,
    ExecutionError,
    JoinType,
    PhysicalExpr,
    PhysicalPlan,
    UnaryOp,
    col,
    lit,
)


@dataclass
class _FinalCol:
    """How output column i of the original HashAggregate is rebuilt from the gathered partial columns."""

    kind: str  # "sum" | "min" | "max" | "avg"
    a: int  # partial column (offset among the aggregate columns)
    b: int = -1  # avg: the partial count column


def split_aggregates(aggr_exprs: Sequence[AggregateExpr]) -> Tuple[List[AggregateExpr], List[_FinalCol], int]:
    """PartialAggregate / FinalAggregate split (planner.rs:200-226).

    COUNT(e) -> partial COUNT(e), final SUM;  SUM -> SUM, SUM;  MIN -> MIN, MIN;  MAX -> MAX, MAX;
    AVG(e) -> partial SUM(e) + COUNT(e), final SUM(s) as f64 / SUM(c) as f64 (operators.rs:770-808).
    A trailing COUNT(1) counts the rows that reached the aggregate on each rank (the reference returns Ok(vec![])
    when none did).  Returns (partial aggregates, recipe per original aggregate, index of the row counter)."""
    partial: List[AggregateExpr] = []
    recipe: List[_FinalCol] = []

    def add(func, expr) -> int:
        for i, p in enumerate(partial):
            if p.func == func and p.expr == expr:
                return i
        partial.append(AggregateExpr(func, expr))
        return len(partial) - 1

    for a in aggr_exprs:
        if a.func == F.Count:
            recipe.append(_FinalCol("sum", add(F.Count, a.expr)))
        elif a.func == F.Sum:
            recipe.append(_FinalCol("sum", add(F.Sum, a.expr)))
        elif a.func == F.Min:
            recipe.append(_FinalCol("min", add(F.Min, a.expr)))
        elif a.func == F.Max:
            recipe.append(_FinalCol("max", add(F.Max, a.expr)))
        elif a.func == F.Avg:
            recipe.append(_FinalCol("avg", add(F.Sum, a.expr), add(F.Count, a.expr)))
        else:
            raise ExecutionError(f"unknown aggregate function {a.func}")
    rows = add(F.Count, lit(1))
    return partial, recipe, rows


def final_plan_exprs(n_keys: int, recipe: Sequence[_FinalCol], rows_col: int):
    """(final aggregate exprs over the gathered table, projection exprs over the final aggregate's output).

    Gathered table columns: keys 0..n_keys-1, then the partial aggregate columns.  The final aggregate emits
    keys, then one merged column per partial column; the projection picks / combines them into the original
    HashAggregate's output order."""
    n_partial = max([rows_col] + [max(r.a, r.b) for r in recipe]) + 1
    kinds = ["sum"] * n_partial
    for r in recipe:
        if r.kind in ("min", "max"):
            kinds[r.a] = r.kind
    func_of = {"sum": F.Sum, "min": F.Min, "max": F.Max}
    final_aggs = [AggregateExpr(func_of[kinds[j]], col(f"col_{j}", n_keys + j)) for j in range(n_partial)]
    proj: List[PhysicalExpr] = [col(f"k{g}", g) for g in range(n_keys)]
    for r in recipe:
        if r.kind == "avg":
            s = PhysicalExpr.UnaryExpr(UnaryOp.CastFloat64, col("s", n_keys + r.a))
            c = PhysicalExpr.UnaryExpr(UnaryOp.CastFloat64, col("c", n_keys + r.b))
            proj.append(PhysicalExpr.BinaryExpr(s, BinaryOp.Divide, c))
        else:
            proj.append(col("v", n_keys + r.a))
    return final_aggs, proj


@dataclass
class SortColumn:
    """query_distributed::SortColumn (crates/query-distributed/src/operators.rs:97-104)"""

    name: str
    ascending: bool = True
    nulls_first: bool = False


class Merge:
    """query_distributed::Merge (crates/query-distributed/src/operators.rs:63-194): combines the partitions' outputs.

    `Merge.concat(schema)`  - the batches of all partitions in partition order (concat_merge, :136-138)
    `Merge.sorted(schema, sort_columns)` - all batches concatenated and sorted by the named columns into ONE batch
    (sorted_merge, :141-194): a device lexsort + take (csrc/sort_kernels.cu); no partition has rows -> no batch."""

    def __init__(self, schema: pa.Schema, strategy: str, sort_columns: Sequence[SortColumn] = ()):
        self.schema, self.strategy, self.sort_columns = schema, strategy, list(sort_columns)

    @classmethod
    def concat(cls, schema: pa.Schema) -> "Merge":
        return cls(schema, "concat")

    @classmethod
    def sorted(cls, schema: pa.Schema, sort_columns: Sequence[SortColumn]) -> "Merge":
        return cls(schema, "sorted", sort_columns)

    def execute(self, ctx: _ffi.Context, partitions: Sequence[Sequence[pa.RecordBatch]]) -> List[pa.RecordBatch]:
        batches = [b for part in partitions for b in part]
        if self.strategy == "concat" or not batches:
            return batches
        table = ctx.import_batches(batches, batches[0].schema)
        out = self.execute_device(ctx, table)
        return [] if out is None else out.to_batches()

    def execute_device(self, ctx: _ffi.Context, table: Optional[_ffi.Table]) -> Optional[_ffi.Table]:
        """the same on a device-resident concatenation of the partitions (e.g. the result of `Context.allgather`)"""
        if table is None or self.strategy == "concat":
            return table
        cols = self.sort_columns
        return ctx.sort(table, [c.name for c in cols], [c.ascending for c in cols], [c.nulls_first for c in cols])


class DistributedExecutor:
    """Runs a PhysicalPlan over row-sharded inputs: every rank calls `execute` with the plan over ITS shard."""

    def __init__(self, executor: QueryExecutor, plan_level: bool = False):
        self.ex = executor
        self.ctx = executor.ctx
        self.plan_level = plan_level

    # -- public ---------------------------------------------------------------------------------
    def execute(self, plan: PhysicalPlan) -> List[pa.RecordBatch]:
        t = self.execute_device(plan)
        return [] if t is None else t.to_batches()

    def execute_device(self, plan: PhysicalPlan) -> Optional[_ffi.Table]:
        if isinstance(plan, PhysicalPlan.HashAggregate):
            return self._aggregate(plan)
        if isinstance(plan, PhysicalPlan.HashJoin):
            return self._join(plan)
        if isinstance(plan, PhysicalPlan.Filter):
            return self.ex.execute_filter(self.execute_device(plan.input), plan.predicate)
        if isinstance(plan, PhysicalPlan.Projection):
            return self.ex.execute_projection(self.execute_device(plan.input), plan.exprs, plan.schema)
        if isinstance(plan, PhysicalPlan.Limit):
            return self._limit(plan)
        if isinstance(plan, (PhysicalPlan.Sort, PhysicalPlan.SubqueryScan)):
            return self.execute_device(plan.input if isinstance(plan, PhysicalPlan.Sort) else plan.subquery)
        return self.ex.execute_device(plan)

    # -- Limit over the concatenation of the ranks' shards (Merge::Concat order = rank order) -----------
    def _limit(self, plan) -> Optional[_ffi.Table]:
        """executor.rs:299-341 on a result that lives sharded: every rank learns how many rows the ranks before it hold
        (one all-gather of a single Int64) and keeps its slice of the global [skip, skip + fetch) window"""
        inp = self.execute_device(plan.input)
        n = 0 if inp is None else inp.num_rows
        mine = self.ctx.import_batches([pa.RecordBatch.from_arrays([pa.array([n], type=pa.int64())], names=["n"])])
        counts = pa.Table.from_batches(self.ctx.allgather(mine).to_batches()).column(0).to_pylist()
        before = sum(counts[: self.ctx.rank])
        lo = max(plan.skip, before)
        hi = before + n if plan.fetch is None else min(before + n, plan.skip + plan.fetch)
        if inp is None or hi <= lo:
            return None
        return self.ctx.execute_limit(inp, lo - before, hi - lo)

    # -- Scan that keeps the schema of an empty shard -----------------------------------------------
    def _input(self, plan: PhysicalPlan) -> Optional[_ffi.Table]:
        if isinstance(plan, PhysicalPlan.Scan) and not isinstance(plan.source, GpuDataSource):
            batches = plan.source.scan()
            schema = batches[0].schema if batches else plan.source.schema()
            self.ex.h2d_bytes += sum(b.nbytes for b in batches)
            return self.ctx.import_batches(batches, schema)  # zero batches still carry the column types
        return self.execute_device(plan)

    # -- Aggregate: partial -> all-gather -> final ----------------------------------------------------
    def _aggregate(self, plan) -> Optional[_ffi.Table]:
        """PartialAggregate(n) -> FinalAggregate.  Default: one library call per rank ("agg.merge_ranks") — the scan
        kernel aggregates the rank's shard, the raw accumulators (scalar) or (group key, accumulator row) pairs
        (grouped) are all-gathered over NCCL and merged on the device, every rank returns the final table.
        `plan_level=True` keeps the same topology as three operator calls (partial plan, all-gather, final plan +
        projection) — the form the reference's DistributedPlanner describes, used by the host-logic tests."""
        if not plan.aggr_exprs:
            return None
        child, pred = plan.input, None
        if isinstance(child, PhysicalPlan.Filter):
            child, pred = child.input, child.predicate
        if isinstance(child, PhysicalPlan.Scan) and not isinstance(child.source, GpuDataSource) and not self.plan_level:
            # host shard: upload only the columns the aggregate reads (projection pushdown, as on one GPU)
            batches = child.source.scan()
            if batches:
                pruned, groups, aggs, pred = prune_for_aggregate(batches, plan.group_exprs, plan.aggr_exprs, pred)
                self.ex.h2d_bytes += sum(b.nbytes for b in pruned)
                inp = self.ctx.import_batches(pruned, pruned[0].schema)
                plan = PhysicalPlan.HashAggregate(plan.input, groups, aggs)
            else:
                inp = self.ctx.import_batches([], child.source.schema())
        else:
            inp = self._input(child)
        if inp is None:
            raise ExecutionError("a rank has no input table for the distributed aggregate (its Scan must carry a schema)")
        if self.plan_level:
            return self._aggregate_plan_level(plan, inp, pred)
        self.ctx.set_option("agg.merge_ranks", 1)
        try:
            return self.ctx.execute_aggregate(inp, list(plan.group_exprs), list(plan.aggr_exprs), pred)
        finally:
            self.ctx.set_option("agg.merge_ranks", 0)

    def _aggregate_plan_level(self, plan, inp, pred) -> Optional[_ffi.Table]:
        n_keys = len(plan.group_exprs)
        partial, recipe, rows_col = split_aggregates(plan.aggr_exprs)
        self.ctx.set_option("agg.keep_empty", 1)
        try:
            local = self.ctx.execute_aggregate(inp, list(plan.group_exprs), partial, pred)
        finally:
            self.ctx.set_option("agg.keep_empty", 0)
        gathered = self.ctx.allgather(local)
        if gathered is None:
            return None
        final_aggs, proj = final_plan_exprs(n_keys, recipe, rows_col)
        keys = [col(f"k{g}", g) for g in range(n_keys)]
        merged = self.ctx.execute_aggregate(gathered, keys, final_aggs)
        if merged is None:
            return None
        if n_keys == 0:
            # the reference returns Ok(vec![]) when no row reached the aggregate (executor.rs:178, 196-198)
            rows = merged.to_batches()[0].column(rows_col)[0].as_py()
            if not rows:
                return None
        names = [getattr(g, "name", f"group_{i}") for i, g in enumerate(plan.group_exprs)]
        names += [f"col_{i}" for i in range(len(plan.aggr_exprs))]
        return self.ctx.execute_projection(merged, proj, names)

    # -- Join: shuffle both sides on the key, join locally ---------------------------------------------
    def _join(self, plan) -> Optional[_ffi.Table]:
        left = self._input(plan.left)
        right = self._input(plan.right)
        if left is None or right is None:
            raise ExecutionError("a rank has no input table for the distributed join (its Scan must carry a schema)")
        lkeys, rkeys = join_key_names(plan.on, left.num_columns)
        # fast path: the join's own radix scatter is the exchange (NVLink stores into the owners' arenas)
        done, joined = self.ctx.execute_join_exchange(left, right, int(plan.join_type), plan.on)
        if done:
            return joined
        left_s = self.ctx.shuffle_hash(left, lkeys)
        # LEFT: a rank that received probe rows but no build rows still emits them, with NULL right columns (what the
        # single-GPU join and the NVLink-exchange path do) - so its empty build shard stays a typed zero-row table
        keep = int(plan.join_type) == int(JoinType.Left)
        if keep:
            self.ctx.set_option("shuffle.keep_empty", 1)
        try:
            right_s = self.ctx.shuffle_hash(right, rkeys)
        finally:
            if keep:
                self.ctx.set_option("shuffle.keep_empty", 0)
        if left_s is None or right_s is None:
            return None  # this rank owns no rows of one side: its share of the result is empty
        return self.ctx.execute_join(left_s, right_s, int(plan.join_type), plan.on)


def join_key_names(on: Optional[PhysicalExpr], n_left: int) -> Tuple[List[str], List[str]]:
    """Key column names of `l.a = r.b [AND ...]`; indices address the merged left++right schema (planner.rs:148-157)."""
    if on is None:
        raise _ffi.UnsupportedOnGpu("join without ON keeps the reference's Cartesian CPU path")
    lk: List[str] = []
    rk: List[str] = []

    def walk(e):
        if isinstance(e, PhysicalExpr.BinaryExpr) and e.op == BinaryOp.And:
            walk(e.left)
            walk(e.right)
            return
        if isinstance(e, PhysicalExpr.BinaryExpr) and e.op == BinaryOp.Equal:
            a, b = e.left, e.right
            if isinstance(a, PhysicalExpr.Column) and isinstance(b, PhysicalExpr.Column):
                if a.index < n_left <= b.index:
                    lk.append(a.name)
                    rk.append(b.name)
                    return
                if b.index < n_left <= a.index:
                    lk.append(b.name)
                    rk.append(a.name)
                    return
        raise _ffi.UnsupportedOnGpu("ON must be Column = Column (AND ...) across the two sides to be accelerated")

    walk(on)
    return lk, rk
