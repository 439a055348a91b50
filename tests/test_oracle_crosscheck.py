"""CPU: the oracle's restatement of arrow-rs semantics cross-checked against pyarrow.compute where Arrow C++ and
arrow-rs agree, and its documented divergences (totalOrder floats, non-Kleene AND/OR, checked ints)."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import oracle
from query_engine_b200 import (
    AggregateExpr,
    AggregateFunction as F,
    BinaryOp as B,
    JoinType,
    MemoryDataSource,
    PhysicalExpr,
    PhysicalPlan,
    QueryError,
    UnaryOp,
    binary,
    col,
    lit,
)
from tests.helpers import assert_batches_equal, random_table


def scan(batches):
    return PhysicalPlan.Scan(MemoryDataSource(batches[0].schema, batches))


@pytest.mark.parametrize("null_frac", [0.0, 0.3])
def test_filter_vs_pyarrow(null_frac):
    rng = np.random.default_rng(1)
    b = random_table(rng, 5000, [("a", "i64"), ("b", "f64"), ("c", "i32")], null_frac)
    pred = binary(binary(col("a", 0), B.Greater, lit(25)), B.And, binary(col("b", 1), B.LessEqual, lit(10.5)))
    got = oracle.execute_filter([b], pred)
    mask = pc.and_(pc.greater(b.column(0), 25), pc.less_equal(b.column(1), 10.5))  # null & x -> null for non-false... see below
    # non-Kleene AND == pyarrow `and_` (not and_kleene); null mask rows are dropped by both filters
    want = b.filter(mask.fill_null(False))
    assert_batches_equal(got, [want], ordered=True)


def test_non_kleene_and_or():
    b = pa.RecordBatch.from_arrays([pa.array([True, False, None, None]), pa.array([None, None, True, False])], names=["x", "y"])
    a = oracle.evaluate_expr(binary(col("x", 0), B.And, col("y", 1)), b)
    o = oracle.evaluate_expr(binary(col("x", 0), B.Or, col("y", 1)), b)
    assert oracle.to_arrow(a).to_pylist() == [None, None, None, None]  # NULL AND FALSE = NULL (arrow `and`, not Kleene)
    assert oracle.to_arrow(o).to_pylist() == [None, None, None, None]


def test_float_total_order_compare():
    nan, nzero = float("nan"), -0.0
    b = pa.RecordBatch.from_arrays([pa.array([nan, nzero, 1.0, nan]), pa.array([nan, 0.0, nan, 1.0])], names=["x", "y"])
    eq = oracle.to_arrow(oracle.evaluate_expr(binary(col("x", 0), B.Equal, col("y", 1)), b)).to_pylist()
    lt = oracle.to_arrow(oracle.evaluate_expr(binary(col("x", 0), B.Less, col("y", 1)), b)).to_pylist()
    assert eq == [True, False, False, False]  # NaN == NaN, -0.0 != +0.0 under totalOrder
    assert lt == [False, True, True, False]  # -0.0 < +0.0, 1.0 < NaN, NaN is greatest


def test_checked_integer_arithmetic_and_modulo():
    b = pa.RecordBatch.from_arrays([pa.array([2**62, 7, -7], type=pa.int64()), pa.array([4, 0, 2], type=pa.int64())], names=["a", "z"])
    with pytest.raises(QueryError, match="Overflow"):
        oracle.evaluate_expr(binary(col("a", 0), B.Multiply, col("z", 1)), b)
    with pytest.raises(QueryError, match="Divide by zero"):
        oracle.evaluate_expr(binary(col("a", 0), B.Divide, col("z", 1)), b)
    m = oracle.to_arrow(oracle.evaluate_expr(binary(col("a", 0), B.Modulo, col("z", 1)), b)).to_pylist()
    assert m == [0, None, -1]  # % 0 -> NULL; sign of the dividend
    neg = oracle.to_arrow(oracle.evaluate_expr(PhysicalExpr.UnaryExpr(UnaryOp.Minus, col("a", 0)), b)).to_pylist()
    assert neg == [-(2**62), -7, 7]
    with pytest.raises(QueryError, match="Unsupported types for multiplication"):
        oracle.evaluate_expr(binary(col("a", 0), B.Multiply, lit(1.5)), b)  # Int64 * Float64: no coercion in arithmetic


@pytest.mark.parametrize("null_frac", [0.0, 0.4])
def test_scalar_aggregates_vs_pyarrow(null_frac):
    rng = np.random.default_rng(2)
    b = random_table(rng, 4000, [("a", "i64"), ("b", "f64"), ("c", "i32")], null_frac)
    aggs = [AggregateExpr(f, col(n, i)) for i, n in enumerate(["a", "b", "c"]) for f in (F.Count, F.Sum, F.Avg, F.Min, F.Max)]
    out = oracle.execute(PhysicalPlan.HashAggregate(scan([b]), [], aggs))[0]
    k = 0
    for i in range(3):
        c = b.column(i)
        assert out.column(k)[0].as_py() == pc.count(c).as_py()
        s = pc.sum(c).as_py()
        assert out.column(k + 1)[0].as_py() == pytest.approx(s, rel=1e-12)
        assert out.column(k + 2)[0].as_py() == pytest.approx(pc.mean(c).as_py(), rel=1e-12)
        assert out.column(k + 3)[0].as_py() == pc.min(c).as_py()
        assert out.column(k + 4)[0].as_py() == pc.max(c).as_py()
        assert out.schema.field(k + 3).type == c.type  # MIN / MAX keep the input type (Int32 stays Int32)
        assert out.schema.field(k + 1).type == (pa.float64() if pa.types.is_floating(c.type) else pa.int64())
        k += 5


def test_group_by_vs_pyarrow():
    rng = np.random.default_rng(3)
    b = random_table(rng, 3000, [("g", "smallint"), ("s", "str"), ("v", "f64"), ("w", "i64")], 0.1)
    aggs = [AggregateExpr(F.Count, col("v", 2)), AggregateExpr(F.Sum, col("w", 3)), AggregateExpr(F.Min, col("v", 2)), AggregateExpr(F.Max, col("w", 3))]
    got = oracle.execute(PhysicalPlan.HashAggregate(scan([b]), [col("g", 0), col("s", 1)], aggs))
    t = pa.Table.from_batches([b]).group_by(["g", "s"], use_threads=False).aggregate([("v", "count"), ("w", "sum"), ("v", "min"), ("w", "max")])
    want = pa.table([t.column("g"), t.column("s"), t.column("v_count"), t.column("w_sum"), t.column("v_min"), t.column("w_max")],
                    names=["g", "s", "col_0", "col_1", "col_2", "col_3"])
    assert_batches_equal(got, want.to_batches(), ordered=False)


@pytest.mark.parametrize("join_type,pa_type", [(JoinType.Inner, "inner"), (JoinType.Left, "left outer")])
def test_join_vs_pyarrow(join_type, pa_type):
    rng = np.random.default_rng(4)
    left = random_table(rng, 800, [("lk", "key"), ("lv", "f64")], 0.1)
    right = random_table(rng, 300, [("rk", "key"), ("rv", "i64")], 0.1)
    plan = PhysicalPlan.HashJoin(scan([left]), scan([right]), join_type, binary(col("lk", 0), B.Equal, col("rk", 2)))
    got = oracle.execute(plan)
    want = pa.Table.from_batches([left]).join(pa.Table.from_batches([right]), keys="lk", right_keys="rk", join_type=pa_type,
                                              coalesce_keys=False).select(["lk", "lv", "rk", "rv"])
    assert_batches_equal(got, want.to_batches(), ordered=False)
    # the reference as shipped: Cartesian product that ignores `on` (executor.rs:363-399, 500-540)
    asis = oracle.execute(plan, reference_asis=True)
    assert sum(x.num_rows for x in asis) == 800 * 300
