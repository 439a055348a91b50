"""GPU parity on the value edge cases SURVEY.md 8(a) spells out: IEEE totalOrder comparisons / MIN / MAX over NaN, -0.0,
+-inf and subnormals (arrow-rs cmp and min/max, operators.rs:509-538, 809-847), `i64 as f64` above 2^53
(coerce_numeric_types, :616-709), wrapping unary minus and wrapping integer SUM (:349-380, 751-769), checked arithmetic at
the ends of the Int64 / Int32 ranges with the reference's error TEXT, float group keys, the reference's own sample data
(data/employees.csv, BASELINE config 1) and the Q6 / Q1 plans at SF0.1 - all through libqgpu's C ABI against the oracle.
"""
import json
from pathlib import Path

import numpy as np
import pyarrow as pa
import pyarrow.csv as pacsv
import pytest

from oracle import oracle
from query_engine_b200 import (
    AggregateExpr,
    AggregateFunction as F,
    BinaryOp as B,
    MemoryDataSource,
    PhysicalExpr,
    PhysicalPlan,
    QueryError,
    ScalarValue,
    UnaryOp,
    binary,
    col,
    lit,
)
from tests.helpers import assert_batches_equal

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
I64_MIN, I64_MAX = np.iinfo(np.int64).min, np.iinfo(np.int64).max
I32_MIN, I32_MAX = np.iinfo(np.int32).min, np.iinfo(np.int32).max


def scan(batches):
    return PhysicalPlan.Scan(MemoryDataSource(batches[0].schema, batches))


def special_floats():
    """every class of IEEE value, several times, in a shuffled column (f64 and the same values as f32 where they fit)"""
    neg_nan = np.frombuffer(np.uint64(0xFFF8000000000001).tobytes(), dtype=np.float64)[0]
    payload_nan = np.frombuffer(np.uint64(0x7FF8000000000123).tobytes(), dtype=np.float64)[0]
    vals = [np.nan, neg_nan, payload_nan, 0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324, 2.2250738585072014e-308, 1.0, -1.0, 1e308,
            -1e308, 3.5, -3.5, 9007199254740993.0, -9007199254740992.0]
    rng = np.random.default_rng(5)
    a = np.array(vals * 40, dtype=np.float64)
    b = a.copy()
    rng.shuffle(a)
    rng.shuffle(b)
    g = rng.integers(0, 3, len(a), dtype=np.int64)
    with np.errstate(over="ignore"):
        f32 = a.astype(np.float32)  # (1e308 becomes +inf: fine, one more inf)
    return a, b, f32, g


FLOAT_PREDICATES = {
    "lt_zero": lambda: binary(col("x", 0), B.Less, lit(0.0)),
    "ge_neg_zero": lambda: binary(col("x", 0), B.GreaterEqual, lit(-0.0)),
    "eq_zero": lambda: binary(col("x", 0), B.Equal, lit(0.0)),
    "eq_nan": lambda: binary(col("x", 0), B.Equal, lit(float("nan"))),
    "gt_inf": lambda: binary(col("x", 0), B.Greater, lit(float("inf"))),
    "le_neg_inf": lambda: binary(col("x", 0), B.LessEqual, lit(float("-inf"))),
    "col_eq_col": lambda: binary(col("x", 0), B.Equal, col("x", 0)),
    "col_lt_col": lambda: binary(col("x", 0), B.Less, col("y", 1)),
    "col_ne_col": lambda: binary(col("x", 0), B.NotEqual, col("y", 1)),
    "f32_vs_f64": lambda: binary(col("z", 2), B.LessEqual, col("x", 0)),
    "f32_lt_lit": lambda: binary(col("z", 2), B.Less, lit(ScalarValue.Float32(0.0))),
    "int_vs_float": lambda: binary(col("g", 3), B.Less, col("x", 0)),
    "neg": lambda: binary(PhysicalExpr.UnaryExpr(UnaryOp.Minus, col("x", 0)), B.Greater, lit(0.0)),
    "div_by_zero": lambda: binary(binary(col("x", 0), B.Divide, lit(0.0)), B.Greater, lit(1.0)),
    "sub_inf": lambda: binary(binary(col("x", 0), B.Subtract, col("x", 0)), B.Equal, lit(0.0)),
}


def float_batches(null_frac=0.0):
    a, b, f32, g = special_floats()
    mask = None
    if null_frac:
        mask = np.random.default_rng(9).random(len(a)) < null_frac
    batch = pa.RecordBatch.from_arrays([pa.array(a, mask=mask), pa.array(b), pa.array(f32, mask=mask), pa.array(g)], names=["x", "y", "z", "g"])
    return [batch.slice(0, 300), batch.slice(300)]


@pytest.mark.parametrize("name", sorted(FLOAT_PREDICATES))
@pytest.mark.parametrize("null_frac", [0.0, 0.25])
def test_total_order_comparisons(executor, name, null_frac):
    batches = float_batches(null_frac)
    plan = PhysicalPlan.Filter(scan(batches), FLOAT_PREDICATES[name]())
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True, same_batching=True)


@pytest.mark.parametrize("grouped", [False, True])
@pytest.mark.parametrize("force_global", [0, 1])
def test_min_max_total_order(executor, ctx, grouped, force_global):
    """MIN / MAX order floats by totalOrder (-NaN < -inf < .. < -0 < +0 < .. < +inf < +NaN) and return the exact bits"""
    batches = float_batches(0.25)
    aggs = [AggregateExpr(F.Min, col("x", 0)), AggregateExpr(F.Max, col("x", 0)), AggregateExpr(F.Min, col("z", 2)),
            AggregateExpr(F.Max, col("z", 2)), AggregateExpr(F.Count, col("x", 0)), AggregateExpr(F.Min, col("y", 1)),
            AggregateExpr(F.Max, col("y", 1))]
    plan = PhysicalPlan.HashAggregate(scan(batches), [col("g", 3)] if grouped else [], aggs)
    ctx.set_option("group.force_global", force_global)
    try:
        got = executor.execute(plan)
    finally:
        ctx.set_option("group.force_global", 0)
    assert_batches_equal(got, oracle.execute(plan), ordered=False)


def test_float_group_keys(executor):
    """float GROUP BY keys group by value bits: -0.0 and +0.0, and NaNs of different payloads, are different groups"""
    batches = float_batches(0.25)
    plan = PhysicalPlan.HashAggregate(scan(batches), [col("x", 0)], [AggregateExpr(F.Count, col("y", 1)), AggregateExpr(F.Sum, col("g", 3))])
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=False)


def int_batches():
    big = np.array([I64_MAX, I64_MIN, I64_MAX - 1, I64_MIN + 1, 2**53, 2**53 + 1, -(2**53) - 1, 2**62, -(2**62), 0, 1, -1] * 25, dtype=np.int64)
    rng = np.random.default_rng(3)
    rng.shuffle(big)
    small = rng.integers(-3, 4, len(big), dtype=np.int64)
    i32 = np.array([I32_MAX, I32_MIN, I32_MAX - 1, I32_MIN + 1, 0, 7] * 50, dtype=np.int32)
    f = (2.0 ** 53) + rng.integers(-2, 3, len(big)).astype(np.float64)
    mask = rng.random(len(big)) < 0.2
    d = rng.integers(0, 4, len(big), dtype=np.int64)  # divisors 0..3 (0: modulo gives NULL, division raises)
    b = pa.RecordBatch.from_arrays([pa.array(big), pa.array(small), pa.array(i32), pa.array(f), pa.array(big, mask=mask), pa.array(d)],
                                   names=["a", "s", "c", "f", "an", "d"])
    return [b.slice(0, 128), b.slice(128)]


INT_PREDICATES = {
    # Int64 vs Float64 compares cast the integer `as f64` (lossy above 2^53) - coerce_numeric_types
    "i64_vs_f64_lossy": lambda: binary(col("a", 0), B.Equal, col("f", 3)),
    "i64_lt_f64_lossy": lambda: binary(col("a", 0), B.Less, col("f", 3)),
    "i64_vs_float_lit": lambda: binary(col("a", 0), B.GreaterEqual, lit(9007199254740992.0)),
    "i32_vs_i64": lambda: binary(col("c", 2), B.Less, col("a", 0)),
    "neg_wraps_min": lambda: binary(PhysicalExpr.UnaryExpr(UnaryOp.Minus, col("a", 0)), B.Less, lit(0)),   # -MIN == MIN
    "neg_wraps_i32": lambda: binary(PhysicalExpr.UnaryExpr(UnaryOp.Minus, col("c", 2)), B.Equal, col("c", 2)),
    "mod_extremes": lambda: binary(binary(col("a", 0), B.Modulo, lit(7)), B.Equal, lit(-1)),
    "mod_by_col_zero_is_null": lambda: binary(binary(col("a", 0), B.Modulo, col("d", 5)), B.Equal, lit(0)),
    "add_small_no_overflow": lambda: binary(binary(col("s", 1), B.Add, lit(I64_MAX - 3)), B.Greater, lit(0)),
    "eq_min": lambda: binary(col("a", 0), B.Equal, lit(I64_MIN)),
}


@pytest.mark.parametrize("name", sorted(INT_PREDICATES))
def test_integer_extremes(executor, name):
    batches = int_batches()
    plan = PhysicalPlan.Filter(scan(batches), INT_PREDICATES[name]())
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True, same_batching=True)


@pytest.mark.parametrize("grouped", [False, True])
def test_wrapping_integer_sums(executor, grouped):
    """arrow's integer `sum` wraps (operators.rs:751-769); AVG is computed from the wrapped sum; Int32 sums widen to Int64"""
    batches = int_batches()
    aggs = [AggregateExpr(F.Sum, col("a", 0)), AggregateExpr(F.Avg, col("a", 0)), AggregateExpr(F.Min, col("a", 0)), AggregateExpr(F.Max, col("a", 0)),
            AggregateExpr(F.Sum, col("c", 2)), AggregateExpr(F.Avg, col("c", 2)), AggregateExpr(F.Min, col("c", 2)), AggregateExpr(F.Max, col("c", 2)),
            AggregateExpr(F.Sum, col("an", 4)), AggregateExpr(F.Count, col("an", 4)), AggregateExpr(F.Avg, col("an", 4))]
    plan = PhysicalPlan.HashAggregate(scan(batches), [col("s", 1)] if grouped else [], aggs)
    off = 1 if grouped else 0
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=False, approx_cols=[off + 1, off + 5, off + 10])


CHECKED_ERRORS = [
    # (expression, expected substring of the ArrowError text: the operands of the FIRST failing row in evaluation order)
    ("mul", lambda: binary(col("a", 0), B.Multiply, lit(2))),
    ("add", lambda: binary(col("a", 0), B.Add, col("a", 0))),
    ("sub", lambda: binary(lit(0), B.Subtract, col("a", 0))),          # 0 - MIN overflows
    ("div_min_by_minus_one", lambda: binary(col("a", 0), B.Divide, lit(-1))),
    ("div_zero", lambda: binary(col("a", 0), B.Divide, col("d", 5))),
    ("mod_min_by_minus_one", lambda: binary(col("a", 0), B.Modulo, lit(-1))),
    ("i32_mul", lambda: binary(col("c", 2), B.Multiply, col("c", 2))),
    ("i32_add", lambda: binary(col("c", 2), B.Add, lit(ScalarValue.Int32(1)))),
]


@pytest.mark.parametrize("name,make", CHECKED_ERRORS, ids=[n for n, _ in CHECKED_ERRORS])
def test_checked_arithmetic_error_text(executor, name, make):
    """arrow-rs numeric::{add,sub,mul,div} are checked: the query fails, and the message names the first failing pair of
    operands in evaluation order (first batch, first row) exactly like the reference's ArrowError"""
    batches = int_batches()
    plan = PhysicalPlan.Filter(scan(batches), binary(make(), B.Greater, lit(0) if not name.startswith("i32") else lit(ScalarValue.Int32(0))))
    with pytest.raises(QueryError) as want:
        oracle.execute(plan)
    with pytest.raises(QueryError) as got:
        executor.execute(plan)
    assert str(got.value) == str(want.value)
    # the same expression under a fused aggregate and as a projection
    agg = PhysicalPlan.HashAggregate(scan(batches), [], [AggregateExpr(F.Max, make())])
    with pytest.raises(QueryError) as want:
        oracle.execute(agg)
    with pytest.raises(QueryError) as got:
        executor.execute(agg)
    assert str(got.value) == str(want.value)


def test_errors_only_count_on_valid_and_selected_rows(executor):
    """an overflow on a NULL slot, or on a row the Filter below already dropped, is not an error"""
    a = pa.array([I64_MAX, 5, I64_MAX, 7], type=pa.int64(), mask=np.array([True, False, False, False]))
    keep = pa.array([1, 1, 0, 1], type=pa.int64())
    b = pa.RecordBatch.from_arrays([a, keep], names=["a", "keep"])
    ok = PhysicalPlan.HashAggregate(PhysicalPlan.Filter(scan([b]), binary(col("keep", 1), B.Equal, lit(1))), [],
                                    [AggregateExpr(F.Sum, binary(col("a", 0), B.Add, lit(1)))])
    assert_batches_equal(executor.execute(ok), oracle.execute(ok), ordered=True)
    bad = PhysicalPlan.HashAggregate(scan([b]), [], [AggregateExpr(F.Sum, binary(col("a", 0), B.Add, lit(1)))])
    with pytest.raises(QueryError) as want:
        oracle.execute(bad)
    with pytest.raises(QueryError) as got:
        executor.execute(bad)
    assert str(got.value) == str(want.value)


# ---- BASELINE config 1: the reference's own sample data through libqgpu ------------------------------------------
def employees():
    # employees.csv the way the reference's front-ends read it: dept_id is Utf8 because one row holds the text NULL;
    # arrow-csv treats only the empty string as null
    t = pacsv.read_csv(GOLDEN / "employees.csv", convert_options=pacsv.ConvertOptions(
        column_types={"id": pa.int64(), "name": pa.string(), "age": pa.int64(), "salary": pa.int64(), "dept_id": pa.string()},
        null_values=[""], strings_can_be_null=True))
    return t.to_batches()


def test_config1_employees_group_by_dept(executor):
    """SELECT dept_id, COUNT(*), SUM(salary), AVG(salary) FROM employees WHERE age > 25 GROUP BY dept_id
    (COUNT(*) == COUNT(id): id is never null) against the hand-derived golden of SURVEY 8c and against the oracle"""
    batches = employees()
    names = batches[0].schema.names
    c = lambda n: col(n, names.index(n))  # noqa: E731
    plan = PhysicalPlan.HashAggregate(
        PhysicalPlan.Filter(scan(batches), binary(c("age"), B.Greater, lit(25))), [c("dept_id")],
        [AggregateExpr(F.Count, c("id")), AggregateExpr(F.Sum, c("salary")), AggregateExpr(F.Avg, c("salary"))])
    got = executor.execute(plan)
    assert_batches_equal(got, oracle.execute(plan), ordered=False, approx_cols=[3])
    golden = json.load(open(GOLDEN / "employees_golden.json"))
    rows = sorted(zip(*[c.to_pylist() for c in pa.Table.from_batches(got).columns]))
    want = sorted(tuple(r) for r in golden["tier2_group_by"]["rows"])
    assert [tuple(r) for r in rows] == want
    assert [f.type for f in got[0].schema] == [pa.string(), pa.int64(), pa.int64(), pa.float64()]


def test_config1_employees_scalar(executor):
    batches = employees()
    names = batches[0].schema.names
    c = lambda n: col(n, names.index(n))  # noqa: E731
    plan = PhysicalPlan.HashAggregate(
        PhysicalPlan.Filter(scan(batches), binary(c("age"), B.Greater, lit(25))), [],
        [AggregateExpr(F.Count, c("id")), AggregateExpr(F.Sum, c("salary")), AggregateExpr(F.Avg, c("salary")),
         AggregateExpr(F.Min, c("age")), AggregateExpr(F.Max, c("age"))])
    got = executor.execute(plan)
    assert_batches_equal(got, oracle.execute(plan), ordered=True, approx_cols=[2])
    golden = json.load(open(GOLDEN / "employees_golden.json"))
    assert [c[0].as_py() for c in got[0].columns] == golden["tier1_scalar"]["rows"][0]


# ---- the benchmark plans at SF0.1 against the oracle (smoke() runs them at SF0.002) --------------------------------
@pytest.mark.parametrize("query", ["q6", "q1"])
def test_benchmark_plans_sf01(executor, query):
    from query_engine_b200 import datagen, workloads

    li = datagen.lineitem(0.1)
    cols = workloads.Q6_COLS if query == "q6" else workloads.Q1_COLS
    batches = datagen.to_batches(li, 65536, cols)
    src = MemoryDataSource(batches[0].schema, batches)
    plan = workloads.q6(src, cols) if query == "q6" else workloads.q1(src, cols)
    approx = [0] if query == "q6" else [2, 3, 4, 5, 6, 7, 8]
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=False, approx_cols=approx)


# ---- joins beyond one fixed-width key: composite keys, Utf8 keys, RIGHT / FULL -----------------------------------------
def _csv(name, types):
    t = pacsv.read_csv(GOLDEN / name, convert_options=pacsv.ConvertOptions(column_types=types, null_values=[""], strings_can_be_null=True))
    return t.to_batches()


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
def test_employees_join_departments_on_utf8_key(ctx, join_type):
    """the reference's own sample data: employees.dept_id = departments.dept_id - a Utf8 key (one employee holds the
    text "NULL", which matches nothing; one department has no employee)"""
    from query_engine_b200 import JoinType
    from query_engine_b200.executor import QueryExecutor

    emp = employees()
    dep = _csv("departments.csv", {"dept_id": pa.string(), "dept_name": pa.string(), "location": pa.string()})
    on = binary(col("dept_id", emp[0].schema.names.index("dept_id")), B.Equal, col("dept_id", emp[0].num_columns + 0))
    plan = PhysicalPlan.HashJoin(scan(emp), scan(dep), getattr(JoinType, join_type), on)
    got = QueryExecutor(ctx=ctx).execute(plan)
    assert_batches_equal(got, oracle.execute(plan), ordered=False)


def _join_tables(seed, nl, nr, null_frac):
    rng = np.random.default_rng(seed)

    def side(n, prefix):
        k1 = rng.integers(0, 40, n, dtype=np.int64)
        k2 = np.array([f"k{x}" for x in rng.integers(0, 6, n)], dtype=object)
        k3 = rng.integers(0, 3, n).astype(np.int32)
        mask = rng.random(n) < null_frac if null_frac else None
        return pa.RecordBatch.from_arrays(
            [pa.array(k1, mask=mask), pa.array(list(k2), type=pa.string()), pa.array(k3), pa.array(rng.normal(size=n)),
             pa.array([f"{prefix}{i}" for i in range(n)], type=pa.string())],
            names=[f"{prefix}k1", f"{prefix}k2", f"{prefix}k3", f"{prefix}v", f"{prefix}s"])

    return side(nl, "l"), side(nr, "r")


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
@pytest.mark.parametrize("keys", ["int_and_utf8", "three_keys", "single_int_dups"])
@pytest.mark.parametrize("null_frac", [0.0, 0.15])
def test_general_joins_match_oracle(ctx, join_type, keys, null_frac):
    """composite keys (AND chain of Column = Column, an Int32 column against an Int64 one), Utf8 keys, duplicate keys
    on both sides, NULL key components (never match) - INNER / LEFT / RIGHT / FULL against the oracle's Tier-2 rule"""
    from query_engine_b200 import JoinType, and_all
    from query_engine_b200.executor import QueryExecutor

    l, r = _join_tables(17, 1500, 700, null_frac)
    nlc = l.num_columns
    eq = lambda i, j: binary(col(l.schema.names[i], i), B.Equal, col(r.schema.names[j], nlc + j))  # noqa: E731
    on = {"int_and_utf8": and_all([eq(0, 0), eq(1, 1)]), "three_keys": and_all([eq(0, 0), eq(1, 1), eq(2, 2)]), "single_int_dups": eq(0, 0)}[keys]
    plan = PhysicalPlan.HashJoin(scan([l.slice(0, 1000), l.slice(1000)]), scan([r]), getattr(JoinType, join_type), on)
    got = QueryExecutor(ctx=ctx).execute(plan)
    assert_batches_equal(got, oracle.execute(plan), ordered=False)


# ---- device sort: Merge::sorted_merge (query-distributed/src/operators.rs:141-194) and Top-K ---------------------------
def test_sorted_merge_reference_vector(ctx):
    """the reference's own (only value-asserted) test of the merge operators, operators.rs:343-374: partitions [[3, 1]] and
    [[4, 2]] merged by id ascending give ONE batch 1, 2, 3, 4"""
    from query_engine_b200.distributed import Merge, SortColumn

    def batch(ids):
        return pa.RecordBatch.from_arrays([pa.array(ids, type=pa.int64())], names=["id"])

    merge = Merge.sorted(batch([0]).schema, [SortColumn("id", ascending=True, nulls_first=False)])
    out = merge.execute(ctx, [[batch([3, 1])], [batch([4, 2])]])
    assert len(out) == 1 and out[0].column(0).to_pylist() == [1, 2, 3, 4]
    assert Merge.sorted(batch([0]).schema, [SortColumn("id", True, False)]).execute(ctx, [[], []]) == []
    concat = Merge.concat(batch([0]).schema).execute(ctx, [[batch([1, 2])], [batch([3, 4])]])
    assert sum(b.num_rows for b in concat) == 4 and [b.column(0).to_pylist() for b in concat] == [[1, 2], [3, 4]]


@pytest.mark.parametrize("keys", [[("a", True, False)], [("a", False, True)], [("g", True, True), ("b", False, False)],
                                  [("e", True, False), ("c", True, True), ("d", False, False)]])
def test_sort_matches_arrow_lexsort(ctx, keys):
    """lexsort by one or more columns (Int64 / Float64 with NaN and +-inf / Int32 / Float32 / Boolean), ascending and
    descending, nulls first and last, against arrow's own sort (floats by totalOrder like arrow-rs' lexsort)"""
    rng = np.random.default_rng(41)
    n = 30_000
    b = np.round(rng.normal(0, 100, n), 1)
    b[::97] = np.nan
    b[1::131] = np.inf
    b[2::173] = -np.inf
    mask = rng.random(n) < 0.1
    t = pa.RecordBatch.from_arrays(
        [pa.array(rng.integers(-500, 500, n, dtype=np.int64), mask=mask), pa.array(b, mask=rng.random(n) < 0.1),
         pa.array(rng.integers(-50, 50, n).astype(np.int32), mask=mask), pa.array(rng.normal(0, 3, n).astype(np.float32)),
         pa.array(rng.integers(0, 2, n).astype(bool), mask=rng.random(n) < 0.1), pa.array(rng.integers(0, 4, n, dtype=np.int64), mask=mask),
         pa.array(np.arange(n, dtype=np.int64))],
        names=["a", "b", "c", "d", "e", "g", "row"])
    table = ctx.import_batches([t.slice(0, 10_000), t.slice(10_000)], t.schema)
    names = [k for k, _, _ in keys]
    got = ctx.sort(table, names, [asc for _, asc, _ in keys], [nf for _, _, nf in keys]).to_batches()
    assert len(got) == 1 and got[0].num_rows == n
    g = pa.Table.from_batches(got)
    # the key tuples must come in lexicographic order: NULLs on the requested side, NaN greatest (totalOrder), ties free
    import math

    def rank(v, asc, nf):
        if v is None:
            return (0 if nf else 2, 0, 0.0)
        nan = isinstance(v, float) and math.isnan(v)
        x = 0.0 if nan else float(v)
        return (1, (1 if nan else 0) if asc else (0 if nan else 1), x if asc else -x)

    def tuples(tab):
        cols = [tab.column(k).to_pylist() for k, _, _ in keys]
        return [tuple(rank(c[i], asc, nf) for c, (_, asc, nf) in zip(cols, keys)) for i in range(tab.num_rows)]

    assert tuples(g) == sorted(tuples(pa.Table.from_batches([t])))
    # the permutation property: every input row appears exactly once
    assert sorted(g.column("row").to_pylist()) == list(range(n))


def test_top_k_and_tier2_sort_node(ctx):
    """ORDER BY .. LIMIT k as one call (the gather touches k rows); the executor's Sort node stays the reference's
    pass-through unless the Tier-2 completion is asked for"""
    from query_engine_b200.executor import QueryExecutor

    rng = np.random.default_rng(43)
    v = rng.permutation(100_000).astype(np.int64)
    b = pa.RecordBatch.from_arrays([pa.array(v), pa.array(v * 0.5)], names=["k", "x"])
    top = ctx.sort(ctx.import_batches([b], b.schema), ["k"], [False], [False], fetch=10).to_batches()
    assert top[0].column(0).to_pylist() == list(range(99_999, 99_989, -1))
    plan = PhysicalPlan.Limit(PhysicalPlan.Sort(scan([b]), [col("k", 0)], [True]), 0, 5)
    asis = QueryExecutor(ctx=ctx).execute(plan)
    assert_batches_equal(asis, oracle.execute(plan), ordered=True)  # pass-through: the first five input rows
    done = QueryExecutor(ctx=ctx, tier2_sort=True).execute(plan)
    assert done[0].column(0).to_pylist() == [0, 1, 2, 3, 4]


def test_resident_catalog_csv_and_parquet(ctx, tmp_path):
    """tables registered once (CSV / Parquet ingest -> HBM), looked up by name, scanned by several
    queries without another upload; dropping a name keeps handed-out sources alive"""
    import pyarrow.parquet as pq

    from query_engine_b200.executor import GpuCatalog, QueryExecutor

    cat = GpuCatalog(ctx)
    cat.register_csv("employees", str(GOLDEN / "employees.csv"), column_types={"dept_id": pa.string()})
    t = pa.table({"k": pa.array(np.arange(5000, dtype=np.int64) % 7), "v": pa.array(np.arange(5000) * 0.5)})
    pq.write_table(t, tmp_path / "t.parquet")
    cat.register_parquet("t", str(tmp_path / "t.parquet"))
    ex = QueryExecutor(ctx=ctx)
    before = ex.h2d_bytes
    src = cat.source("employees")
    names = src.schema().names
    c = lambda n: col(n, names.index(n))  # noqa: E731
    for _ in range(2):
        plan = PhysicalPlan.HashAggregate(PhysicalPlan.Filter(PhysicalPlan.Scan(src), binary(c("age"), B.Greater, lit(25))), [],
                                          [AggregateExpr(F.Count, c("id")), AggregateExpr(F.Sum, c("salary"))])
        got = ex.execute(plan)
        assert [x[0].as_py() for x in got[0].columns] == [5, 428000]
    plan = PhysicalPlan.HashAggregate(PhysicalPlan.Scan(cat.source("t")), [col("k", 0)], [AggregateExpr(F.Sum, col("v", 1))])
    want = oracle.execute(PhysicalPlan.HashAggregate(scan(t.to_batches()), [col("k", 0)], [AggregateExpr(F.Sum, col("v", 1))]))
    assert_batches_equal(ex.execute(plan), want, ordered=False, approx_cols=[1])
    assert ex.h2d_bytes == before  # nothing crossed PCIe again
    cat.drop("t")
    with pytest.raises(QueryError, match="Table 't' not found"):
        cat.source("t")
    assert ex.execute(PhysicalPlan.Limit(PhysicalPlan.Scan(src), 0, 3))[0].num_rows == 3


def test_concurrent_queries_from_four_threads(ctx):
    """SURVEY 8b threading: pgwire calls the executor from several connections at once (backend.rs:214-217, 569).  Four
    threads run different plans - over one shared resident table and over tables of their own - on ONE context at the same
    time (ctypes releases the GIL inside libqgpu; every thread gets a lane = stream + block cache); each result must be
    what the same plan gives alone."""
    import threading

    from query_engine_b200 import JoinType, datagen, workloads
    from query_engine_b200.executor import GpuCatalog, QueryExecutor

    li = datagen.lineitem(0.05)
    od = datagen.orders(0.05)
    cat = GpuCatalog(ctx)
    cat.register_batches("lineitem", datagen.to_batches(li, 32768, workloads.LINEITEM_COLS))
    names = workloads.LINEITEM_COLS

    def plans(tid):
        shared = cat.source("lineitem")
        own_l = MemoryDataSource(None, datagen.to_batches(li, 16384, workloads.JOIN_L_COLS))
        own_o = MemoryDataSource(None, datagen.to_batches(od, 16384, workloads.JOIN_O_COLS))
        return [
            ("q6", workloads.q6(shared, names), [0]),
            ("q1", workloads.q1(shared, names), [2, 3, 4, 5, 6, 7, 8]),
            ("filter", PhysicalPlan.Limit(workloads.filter_only(shared, names, "q6"), tid, 500 + tid), []),
            ("join", workloads.join(own_l, own_o, JoinType.Left if tid % 2 else JoinType.Inner), []),
        ]

    ex = QueryExecutor(ctx=ctx)
    expected = {}
    for tid in range(4):
        for name, plan, _ in plans(tid):
            expected[(tid, name)] = ex.execute(plan)
    errors = []

    def work(tid):
        try:
            mine = QueryExecutor(ctx=ctx)
            for rep in range(3):
                for name, plan, approx in plans(tid):
                    got = mine.execute(plan)
                    assert_batches_equal(got, expected[(tid, name)], ordered=name == "filter", approx_cols=approx)
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    cat.drop("lineitem")
