"""CPU: host-side logic — plan flattening for the C ABI, synthetic data invariants, the partial/final aggregate split
(world_size 2 over gloo, no GPU)."""
import os
import socket

import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle
from query_engine_b200 import (
    AggregateExpr,
    AggregateFunction as F,
    BinaryOp as B,
    MemoryDataSource,
    PhysicalExpr,
    PhysicalPlan,
    ScalarValue,
    _ffi,
    binary,
    col,
    datagen,
    lit,
    workloads,
)
from query_engine_b200.distributed import final_plan_exprs, join_key_names, split_aggregates
from tests.helpers import assert_batches_equal


def test_flatten_is_postfix_and_keeps_literal_types():
    e = binary(binary(col("x", 1), B.Multiply, lit(2.5)), B.Greater, lit(ScalarValue.Int32(7)))
    f = _ffi.FlatExpr(e)
    kinds = [f.array[i].kind for i in range(f.c.n_nodes)]
    assert kinds == [_ffi.EXPR_COLUMN, _ffi.EXPR_LITERAL, _ffi.EXPR_BINARY, _ffi.EXPR_LITERAL, _ffi.EXPR_BINARY]
    assert f.array[0].op == 1 and f.array[0].str == b"x"
    assert f.array[1].op == 5 and f.array[1].value.f64 == 2.5
    assert f.array[2].op == int(B.Multiply) and f.array[4].op == int(B.Greater)
    assert f.array[3].op == 2 and f.array[3].value.i64 == 7
    n = _ffi.FlatExpr(lit(ScalarValue.Int64(None)))
    assert n.array[0].is_null == 1
    with pytest.raises(ValueError):
        _ffi.FlatExpr(lit(ScalarValue.Int8(1000)))
    assert ScalarValue.from_sql_number("25").type.name == "Int64" and ScalarValue.from_sql_number("0.05").type.name == "Float64"


def test_datagen_invariants_and_shards():
    sf = 0.004
    li, od = datagen.lineitem(sf), datagen.orders(sf)
    assert len(li["l_orderkey"]) == 24000 and len(od["o_orderkey"]) == 6000
    assert len(np.unique(od["o_orderkey"])) == 6000
    assert np.isin(li["l_orderkey"], od["o_orderkey"]).all()  # every probe row has exactly one build row
    groups = set(zip(li["l_returnflag"].tolist(), li["l_linestatus"].tolist()))
    assert groups == {(0, 0), (1, 0), (2, 0), (2, 1)}  # A/F, R/F, N/F, N/O
    m = ((li["l_shipdate"] >= datagen.Q6_DATE_LO) & (li["l_shipdate"] < datagen.Q6_DATE_HI) & (li["l_discount"] >= 0.05)
         & (li["l_discount"] <= 0.07) & (li["l_quantity"] < 24.0))
    assert 0.005 < m.mean() < 0.04  # Q6 selectivity ~ 1.8 %
    # shards: independent, sizes add up, lines stay with their own orders' shard
    tot = 0
    for r in range(4):
        ls, os_ = datagen.lineitem(sf, r, 4), datagen.orders(sf, r, 4)
        tot += len(ls["l_orderkey"])
        assert np.isin(ls["l_orderkey"], os_["o_orderkey"]).all()
    assert tot == 24000
    hc = datagen.highcard(50_000, 1000)
    assert len(np.unique(hc["k"])) <= 1000 and len(hc["k"]) == 50_000


def test_workload_plans_have_the_planner_shape():
    q6 = workloads.q6(None)
    assert isinstance(q6, PhysicalPlan.HashAggregate) and isinstance(q6.input, PhysicalPlan.Filter) and not q6.group_exprs
    q1 = workloads.q1(None)
    assert [g.name for g in q1.group_exprs] == ["l_returnflag", "l_linestatus"] and len(q1.aggr_exprs) == 8
    j = workloads.join(None, None)
    assert join_key_names(j.on, 2) == (["l_orderkey"], ["o_orderkey"])
    rev = binary(col("o_orderkey", 2), B.Equal, col("l_orderkey", 0))
    assert join_key_names(rev, 2) == (["l_orderkey"], ["o_orderkey"])
    with pytest.raises(Exception):
        join_key_names(binary(col("a", 0), B.Less, col("b", 2)), 2)


def test_split_aggregates_recipe():
    aggs = [AggregateExpr(F.Avg, col("v", 1)), AggregateExpr(F.Sum, col("v", 1)), AggregateExpr(F.Count, col("v", 1)),
            AggregateExpr(F.Min, col("w", 2))]
    partial, recipe, rows = split_aggregates(aggs)
    assert [(p.func, p.expr) for p in partial] == [(F.Sum, col("v", 1)), (F.Count, col("v", 1)), (F.Min, col("w", 2)), (F.Count, lit(1))]
    assert [(r.kind, r.a, r.b) for r in recipe] == [("avg", 0, 1), ("sum", 0, -1), ("sum", 1, -1), ("min", 2, -1)]
    assert rows == 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = 3000 + 500 * rank
        mine = pa.RecordBatch.from_arrays(
            [pa.array(rng.integers(0, 6, n, dtype=np.int64)), pa.array(np.round(rng.normal(0, 50, n), 2), mask=rng.random(n) < 0.2),
             pa.array(rng.integers(-100, 100, n, dtype=np.int32))], names=["g", "v", "w"])
        pred = binary(col("w", 2), B.Greater, lit(ScalarValue.Int32(-50)))
        aggs = [AggregateExpr(F.Avg, col("v", 1)), AggregateExpr(F.Sum, col("v", 1)), AggregateExpr(F.Count, col("v", 1)),
                AggregateExpr(F.Min, col("w", 2)), AggregateExpr(F.Max, col("v", 1)), AggregateExpr(F.Avg, col("w", 2))]
        results = {}
        for groups in ([], [col("g", 0)]):
            partial, recipe, rows_col = split_aggregates(aggs)
            src = PhysicalPlan.Filter(PhysicalPlan.Scan(MemoryDataSource(mine.schema, [mine])), pred)
            local = oracle.execute(PhysicalPlan.HashAggregate(src, groups, partial))
            # "all-gather" of the partial tables (the GPU path does this with NCCL; here gloo moves pickled batches)
            parts = [None] * world
            dist.all_gather_object(parts, [b.to_pydict() for b in local])
            gathered = [pa.RecordBatch.from_pydict(d, schema=local[0].schema) for p in parts for d in p]
            final_aggs, proj = final_plan_exprs(len(groups), recipe, rows_col)
            keys = [col(f"k{i}", i) for i in range(len(groups))]
            merged = oracle.execute(PhysicalPlan.HashAggregate(PhysicalPlan.Scan(MemoryDataSource(gathered[0].schema, gathered)), keys, final_aggs))
            names = [g.name for g in groups] + [f"col_{i}" for i in range(len(aggs))]
            out = []
            for b in merged:
                arrays = [oracle.to_arrow(oracle.evaluate_expr(e, b)) for e in proj]
                out.append(pa.RecordBatch.from_arrays(arrays, names=names))
            # expectation: the single-node oracle on the union of the shards
            allrows = [None] * world
            dist.all_gather_object(allrows, mine.to_pydict())
            union = [pa.RecordBatch.from_pydict(d, schema=mine.schema) for d in allrows]
            usrc = PhysicalPlan.Filter(PhysicalPlan.Scan(MemoryDataSource(mine.schema, union)), pred)
            want = oracle.execute(PhysicalPlan.HashAggregate(usrc, groups, aggs))
            k = len(groups)
            assert_batches_equal(out, want, ordered=False, approx_cols=[k + 0, k + 1, k + 5])
            results[len(groups)] = sum(b.num_rows for b in out)
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def test_partial_final_split_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = dict(q.get(timeout=10) for _ in range(2))
    assert got[0] == got[1] == {0: 1, 1: 6}


class _RecordingCtx:
    """stands in for the device context: records what QueryExecutor hands to the library"""

    def __init__(self):
        self.imported = None
        self.call = None

    def import_batches(self, batches, schema):
        self.imported = (list(batches), schema)
        return "table"

    def execute_aggregate(self, table, group_exprs, aggr_exprs, predicate=None):
        self.call = (table, list(group_exprs), list(aggr_exprs), predicate)
        return "result"


def test_projection_pushdown_under_aggregate_over_host_scan():
    """only the referenced columns are uploaded, indices are remapped, COUNT(col) of a null-free column needs no data;
    an out-of-range index keeps the full upload so that the library raises the reference's error"""
    import numpy as np
    import pyarrow as pa

    from query_engine_b200 import AggregateExpr, AggregateFunction as F, BinaryOp as B, MemoryDataSource, PhysicalPlan, binary, col, lit
    from query_engine_b200.executor import QueryExecutor
    from query_engine_b200.physical_plan import PhysicalExpr

    n = 10
    b = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.int64)), pa.array(np.arange(n) * 1.0), pa.array(np.arange(n) * 2.0),
         pa.array([None if i == 3 else i for i in range(n)], type=pa.int64()), pa.array(np.arange(n, dtype=np.int64) % 3)],
        names=["id", "x", "y", "maybe", "g"])
    src = MemoryDataSource(b.schema, [b])
    plan = PhysicalPlan.HashAggregate(
        PhysicalPlan.Filter(PhysicalPlan.Scan(src), binary(col("y", 2), B.Greater, lit(1.0))), [col("g", 4)],
        [AggregateExpr(F.Sum, binary(col("y", 2), B.Multiply, col("y", 2))), AggregateExpr(F.Count, col("id", 0)),
         AggregateExpr(F.Count, col("maybe", 3))])
    ctx = _RecordingCtx()
    ex = QueryExecutor(ctx=ctx)
    assert ex.execute_device(plan) == "result"
    batches, schema = ctx.imported
    assert schema.names == ["y", "maybe", "g"]  # id (null-free COUNT) and x (unreferenced) are not uploaded
    assert ex.h2d_bytes == sum(x.nbytes for x in batches)
    _, groups, aggs, pred = ctx.call
    assert groups == [col("g", 2)]
    assert pred == binary(col("y", 0), B.Greater, lit(1.0))
    assert aggs[0] == AggregateExpr(F.Sum, binary(col("y", 0), B.Multiply, col("y", 0)))
    assert aggs[1] == AggregateExpr(F.Count, lit(1))             # COUNT(id): no nulls -> COUNT(1)
    assert aggs[2] == AggregateExpr(F.Count, col("maybe", 1))    # has a null: the column travels
    assert isinstance(aggs[1].expr, PhysicalExpr.Literal)

    # an out-of-range column index: nothing is pruned, the library sees the original expressions
    bad = PhysicalPlan.HashAggregate(PhysicalPlan.Scan(src), [], [AggregateExpr(F.Sum, col("nope", 9))])
    ctx2 = _RecordingCtx()
    QueryExecutor(ctx=ctx2).execute_device(bad)
    assert ctx2.imported[1].names == b.schema.names
    assert ctx2.call[2] == [AggregateExpr(F.Sum, col("nope", 9))]


def test_bench_reads_measured_traffic_of_an_operator():
    """roofline.traffic comes from the committed ncu capture (profiles/traffic.json): a single kernel, or the sum over
    the launches of a multi-kernel operator; unknown kernels give None"""
    import json
    import sys
    from pathlib import Path

    ROOT = Path(__file__).resolve().parent.parent
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    t = json.load(open(ROOT / "profiles" / "traffic.json"))
    one = bench.measured_traffic("k_scan_filter_agg_jit")
    assert one == t["k_scan_filter_agg_jit"]["dram_bytes_per_launch"]
    members = [k for k in t if k.startswith(("k_pj_", "k_rscatter_", "k_region_")) and "groups" not in k]
    op = " + ".join(members + ["k_region_offsets"])  # (a helper launch without a capture adds nothing)
    assert members and bench.measured_traffic(op) == sum(t[k]["dram_bytes_per_launch"] for k in members)
    assert bench.measured_traffic("k_does_not_exist") is None
    assert bench.measured_traffic(None) is None
    # a main launch of the operator without a capture of its own: no figure rather than an understated sum
    assert bench.measured_traffic("k_scan_filter_agg_jit + k_compact_of_another_workload") is None
    # the Q6 scan must not move more than its algorithmic bytes (survivor-only loads): 4 columns x 8 B x 60 M rows
    assert one <= 1.02 * 1.92e9
