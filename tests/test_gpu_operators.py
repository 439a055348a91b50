"""GPU parity tests: libqgpu (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bit-exact for integer / key / count / copied values and row sets; 1e-9 relative for f64 SUM / AVG (north_star).
"""
import numpy as np
import pyarrow as pa
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
    ScalarValue,
    UnaryOp,
    and_all,
    binary,
    col,
    lit,
)
from tests.helpers import F32_SUM_RTOL, F64_RTOL, assert_batches_equal, random_table, split_batches

pytestmark = pytest.mark.gpu

SPEC = [("a", "i64"), ("b", "f64"), ("c", "i32"), ("d", "f32"), ("e", "bool"), ("g", "smallint"), ("s", "str")]


def scan(batches):
    return PhysicalPlan.Scan(MemoryDataSource(batches[0].schema, batches))


def make_batches(seed, n, null_frac, sizes=None):
    rng = np.random.default_rng(seed)
    b = random_table(rng, n, SPEC, null_frac)
    if sizes is None:
        sizes = [n]
    return split_batches(b, sizes)


PREDICATES = {
    "cmp_lit": binary(col("a", 0), B.Greater, lit(25)),
    "cmp_float_int_coerce": binary(col("b", 1), B.LessEqual, lit(3)),
    "i32_vs_i64": binary(col("c", 2), B.Less, col("a", 0)),
    "f32_vs_f64": binary(col("d", 3), B.GreaterEqual, col("b", 1)),
    "and_chain": and_all([binary(col("a", 0), B.GreaterEqual, lit(-500)), binary(col("b", 1), B.Less, lit(50.5)),
                          binary(col("c", 2), B.NotEqual, lit(ScalarValue.Int32(7)))]),
    "or": binary(binary(col("a", 0), B.Less, lit(-900)), B.Or, binary(col("b", 1), B.Greater, lit(150.0))),
    "not_bool_col": PhysicalExpr.UnaryExpr(UnaryOp.Not, col("e", 4)),
    "bool_col": col("e", 4),
    "arith": binary(binary(binary(col("a", 0), B.Multiply, lit(3)), B.Add, col("a", 0)), B.Greater, lit(100)),
    "nested_both_sides": binary(binary(col("b", 1), B.Multiply, col("b", 1)), B.Greater,
                                binary(binary(col("b", 1), B.Add, lit(1.5)), B.Multiply, lit(80.0))),
    "modulo": binary(binary(col("a", 0), B.Modulo, col("g", 5)), B.Equal, lit(1)),
    "neg": binary(PhysicalExpr.UnaryExpr(UnaryOp.Minus, col("a", 0)), B.Greater, lit(10)),
    "null_literal": binary(col("a", 0), B.Greater, lit(ScalarValue.Int64(None))),
    "and_with_or_inside": binary(binary(col("e", 4), B.Or, binary(col("a", 0), B.Greater, lit(0))), B.And,
                                 binary(col("b", 1), B.Less, lit(0.0))),
    "div_float": binary(binary(col("b", 1), B.Divide, col("b", 1)), B.Equal, lit(1.0)),
}


@pytest.mark.parametrize("name", sorted(PREDICATES))
@pytest.mark.parametrize("null_frac", [0.0, 0.2])
def test_filter_matches_oracle(executor, name, null_frac):
    batches = make_batches(11, 5000, null_frac, sizes=[1024, 1024, 1, 0, 2951])
    pred = PREDICATES[name]
    if name == "null_literal":
        # Int64 vs NullArray: arrow-rs refuses the comparison; both sides must raise the same ArrowError
        with pytest.raises(QueryError) as e_want:
            oracle.execute_filter(batches, pred)
        with pytest.raises(QueryError) as e_got:
            executor.execute(PhysicalPlan.Filter(scan(batches), pred))
        assert str(e_got.value) == str(e_want.value) == "ArrowError: Invalid argument error: Invalid comparison operation: Int64 > Null"
        return
    want = oracle.execute_filter(batches, pred)
    got = executor.execute(PhysicalPlan.Filter(scan(batches), pred))
    assert_batches_equal(got, want, ordered=True, same_batching=True)


def test_filter_drops_everything(executor):
    batches = make_batches(3, 2000, 0.0, sizes=[1000, 1000])
    got = executor.execute(PhysicalPlan.Filter(scan(batches), binary(col("a", 0), B.Greater, lit(10**6))))
    assert got == []


def test_filter_large_ragged(executor):
    batches = make_batches(5, 300_001, 0.05, sizes=[100_000, 8192, 191_809])
    pred = PREDICATES["and_chain"]
    assert_batches_equal(executor.execute(PhysicalPlan.Filter(scan(batches), pred)), oracle.execute_filter(batches, pred),
                         ordered=True, same_batching=True)


ALL_AGGS = [
    AggregateExpr(F.Count, col("a", 0)),
    AggregateExpr(F.Sum, col("a", 0)),
    AggregateExpr(F.Avg, col("a", 0)),
    AggregateExpr(F.Min, col("a", 0)),
    AggregateExpr(F.Max, col("a", 0)),
    AggregateExpr(F.Sum, col("b", 1)),
    AggregateExpr(F.Avg, col("b", 1)),
    AggregateExpr(F.Min, col("b", 1)),
    AggregateExpr(F.Max, col("b", 1)),
    AggregateExpr(F.Sum, col("c", 2)),
    AggregateExpr(F.Avg, col("c", 2)),
    AggregateExpr(F.Min, col("c", 2)),
    AggregateExpr(F.Max, col("c", 2)),
    AggregateExpr(F.Sum, col("d", 3)),
    AggregateExpr(F.Avg, col("d", 3)),
    AggregateExpr(F.Min, col("d", 3)),
    AggregateExpr(F.Max, col("d", 3)),
    AggregateExpr(F.Count, col("s", 6)),
    AggregateExpr(F.Sum, binary(col("b", 1), B.Multiply, binary(lit(1.0), B.Subtract, col("b", 1)))),
]
F32_AGGS = {13, 14}  # SUM / AVG over the Float32 column


def approx(offset=0):
    return {offset + i: (F32_SUM_RTOL if i in F32_AGGS else F64_RTOL) for i, a in enumerate(ALL_AGGS) if a.func in (F.Sum, F.Avg)}


@pytest.mark.parametrize("null_frac", [0.0, 0.3])
@pytest.mark.parametrize("with_filter", [False, True])
def test_scalar_aggregates(executor, null_frac, with_filter):
    batches = make_batches(21, 20_000, null_frac, sizes=[8192, 8192, 3616])
    inp = scan(batches)
    if with_filter:
        inp = PhysicalPlan.Filter(inp, binary(col("a", 0), B.Greater, lit(0)))
    plan = PhysicalPlan.HashAggregate(inp, [], ALL_AGGS)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True, approx_cols=approx())


def test_scalar_aggregate_all_null_and_empty(executor):
    b = pa.RecordBatch.from_arrays([pa.array([None, None, None], type=pa.int64()), pa.array([1.0, None, 2.0])], names=["a", "b"])
    aggs = [AggregateExpr(f, col("a", 0)) for f in (F.Count, F.Sum, F.Avg, F.Min, F.Max)]
    plan = PhysicalPlan.HashAggregate(scan([b]), [], aggs)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True)
    # filter that drops every row: the unfused plan feeds zero batches -> Ok(vec![])
    plan2 = PhysicalPlan.HashAggregate(PhysicalPlan.Filter(scan([b]), binary(col("b", 1), B.Greater, lit(5.0))), [], aggs)
    assert oracle.execute(plan2) == []
    assert executor.execute(plan2) == []


@pytest.mark.parametrize("null_frac", [0.0, 0.25])
@pytest.mark.parametrize("force_global", [0, 1])
@pytest.mark.parametrize("keys", [["g"], ["g", "e"], ["s"], ["g", "s"], ["c"]])
def test_group_by(executor, ctx, null_frac, force_global, keys):
    batches = make_batches(31, 6000, null_frac, sizes=[4096, 1904])
    names = [n for n, _ in SPEC]
    groups = [col(k, names.index(k)) for k in keys]
    plan = PhysicalPlan.HashAggregate(scan(batches), groups, ALL_AGGS)
    ctx.set_option("group.force_global", force_global)
    try:
        got = executor.execute(plan)
    finally:
        ctx.set_option("group.force_global", 0)
    want = oracle.execute(plan)
    k = len(keys)
    assert_batches_equal(got, want, ordered=False, approx_cols=approx(k))


@pytest.mark.parametrize("capacity", [1 << 22, 1 << 10])
def test_group_by_with_filter_high_cardinality(executor, ctx, capacity):
    """capacity 1 << 10: the first HBM table is far too small for the 50 000 groups - the pass must give up quickly
    (bounded probe chains, every CTA stops) and be redone with larger tables until the groups fit"""
    rng = np.random.default_rng(41)
    n = 200_000
    b = pa.RecordBatch.from_arrays(
        [pa.array(rng.integers(0, 50_000, n, dtype=np.int64)), pa.array(rng.uniform(0, 1e6, n)),
         pa.array(rng.integers(-10**6, 10**6, n, dtype=np.int64))], names=["k", "v", "w"])
    aggs = [AggregateExpr(F.Min, col("v", 1)), AggregateExpr(F.Max, col("v", 1)), AggregateExpr(F.Sum, col("v", 1)),
            AggregateExpr(F.Sum, col("w", 2)), AggregateExpr(F.Count, col("k", 0))]
    plan = PhysicalPlan.HashAggregate(PhysicalPlan.Filter(scan([b]), binary(col("w", 2), B.Greater, lit(-500_000))),
                                      [col("k", 0)], aggs)
    ctx.set_option("group.capacity", capacity)
    try:
        got = executor.execute(plan)
    finally:
        ctx.set_option("group.capacity", 1 << 22)
    # vectorised check (the oracle's Tier-2 loop is pure Python): pandas groupby on the filtered rows
    import pandas as pd

    df = b.to_pandas()
    df = df[df.w > -500_000]
    gb = df.groupby("k", sort=True)
    want = pa.table({"k": gb.v.min().index.values, "col_0": gb.v.min().values, "col_1": gb.v.max().values,
                     "col_2": gb.v.sum().values, "col_3": gb.w.sum().values, "col_4": gb.k.count().values.astype(np.int64)})
    assert_batches_equal(got, want.to_batches(), ordered=False, approx_cols=[3])


@pytest.mark.parametrize("join_type", [JoinType.Inner, JoinType.Left])
@pytest.mark.parametrize("null_frac", [0.0, 0.2])
def test_join(executor, join_type, null_frac):
    rng = np.random.default_rng(51)
    left = split_batches(random_table(rng, 3000, [("lk", "key"), ("lv", "f64"), ("ls", "str")], null_frac), [1000, 2000])
    right = split_batches(random_table(rng, 700, [("rk", "key"), ("rv", "i64"), ("rb", "bool")], null_frac), [300, 400])
    # keys drawn from different ranges: some probe rows miss, some build keys repeat
    on = binary(col("lk", 0), B.Equal, col("rk", 3))
    plan = PhysicalPlan.HashJoin(scan(left), scan(right), join_type, on)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=False)


def test_join_i32_i64_keys_and_reversed_on(executor):
    lk = pa.array(np.arange(100, dtype=np.int32))
    rk = pa.array(np.arange(50, 150, dtype=np.int64))
    left = [pa.RecordBatch.from_arrays([lk, pa.array(np.arange(100) * 1.5)], names=["a", "x"])]
    right = [pa.RecordBatch.from_arrays([rk, pa.array(np.arange(100) * 2)], names=["b", "y"])]
    on = binary(col("b", 2), B.Equal, col("a", 0))
    plan = PhysicalPlan.HashJoin(scan(left), scan(right), JoinType.Inner, on)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=False)


def test_exchange_hash_matches_siphash_oracle(ctx):
    rng = np.random.default_rng(61)
    b = random_table(rng, 4000, [("id", "i64"), ("c", "i32"), ("s", "str"), ("v", "f64")], 0.1)
    for keys in (["id"], ["id", "c"], ["s"], ["s", "id", "missing"]):
        want = oracle.partition_by_hash([b], keys, 5)
        t = ctx.import_batches([b])
        parts = ctx.exchange_hash(t, keys, 5)
        assert len(parts) == 5
        for p in range(5):
            got = [] if parts[p] is None else parts[p].to_batches()
            assert_batches_equal(got, want[p], ordered=True)
    with pytest.raises(QueryError, match="No key columns found in batch"):
        ctx.exchange_hash(ctx.import_batches([b]), ["nope"], 3)


def test_checked_arithmetic_errors(executor):
    b = pa.RecordBatch.from_arrays([pa.array([1, 2**62, 5], type=pa.int64()), pa.array([0, 1, 2], type=pa.int64())], names=["a", "z"])
    with pytest.raises(QueryError, match="Overflow"):
        executor.execute(PhysicalPlan.Filter(scan([b]), binary(binary(col("a", 0), B.Multiply, lit(4)), B.Greater, lit(0))))
    with pytest.raises(QueryError, match="Divide by zero"):
        executor.execute(PhysicalPlan.Filter(scan([b]), binary(binary(col("a", 0), B.Divide, col("z", 1)), B.Greater, lit(0))))
    # modulo by zero is NULL, not an error (operators.rs:711-743)
    plan = PhysicalPlan.Filter(scan([b]), binary(binary(col("a", 0), B.Modulo, col("z", 1)), B.Equal, lit(0)))
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True)
    # the same errors from the oracle
    with pytest.raises(QueryError):
        oracle.execute(PhysicalPlan.Filter(scan([b]), binary(binary(col("a", 0), B.Multiply, lit(4)), B.Greater, lit(0))))


def test_type_errors_carry_reference_text(executor):
    b = pa.RecordBatch.from_arrays([pa.array([1, 2], type=pa.int64()), pa.array([1.0, 2.0])], names=["a", "b"])
    for expr, text in [
        (binary(binary(col("a", 0), B.Multiply, col("b", 1)), B.Greater, lit(0)), "Unsupported types for multiplication"),
        (col("a", 0), "Filter predicate must return boolean"),
        (binary(col("a", 5), B.Greater, lit(0)), "out of bounds"),
    ]:
        with pytest.raises(QueryError) as e1:
            executor.execute(PhysicalPlan.Filter(scan([b]), expr))
        with pytest.raises(QueryError) as e2:
            oracle.execute(PhysicalPlan.Filter(scan([b]), expr))
        assert text in str(e1.value) and text in str(e2.value)


def test_projection(executor):
    batches = make_batches(71, 3000, 0.1, sizes=[1000, 2000])
    exprs = [col("a", 0), binary(col("b", 1), B.Multiply, lit(2.0)), binary(col("a", 0), B.Greater, lit(0)),
             binary(col("c", 2), B.Add, col("c", 2))]
    schema = pa.schema([("a", pa.int64()), ("b2", pa.float64()), ("pos", pa.bool_()), ("c2", pa.int32())])
    plan = PhysicalPlan.Projection(scan(batches), exprs, schema)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True, same_batching=True)


@pytest.mark.parametrize("join_type", [JoinType.Inner, JoinType.Left])
@pytest.mark.parametrize("dups", [False, True])
def test_join_fixed_width_paths(ctx, join_type, dups):
    """fixed-width, non-null tables on the generic (flat table) path: fused emit (unique build keys) vs index-pair emit
    (duplicate build keys); i32 probe key against an i64 build key"""
    from query_engine_b200.executor import QueryExecutor

    rng = np.random.default_rng(77)
    nl, nr = 40_000, 9_000
    rk = rng.permutation(np.arange(0, 3 * nr, 3, dtype=np.int64))  # unique
    if dups:
        rk[: nr // 10] = rk[nr // 10 : 2 * (nr // 10)]
    lk = rng.integers(0, 3 * nr, nl).astype(np.int32)  # two thirds of the probes miss
    left = [pa.RecordBatch.from_arrays([pa.array(lk), pa.array(rng.normal(size=nl)), pa.array(np.arange(nl, dtype=np.int64))],
                                       names=["lk", "lv", "lrow"])]
    right = [pa.RecordBatch.from_arrays([pa.array(rk), pa.array(rng.normal(size=nr).astype(np.float32)), pa.array(np.arange(nr, dtype=np.int64))],
                                        names=["rk", "rv", "rrow"])]
    plan = PhysicalPlan.HashJoin(scan(left), scan(right), join_type, binary(col("lk", 0), B.Equal, col("rk", 3)))
    got = QueryExecutor(ctx=ctx).execute(plan)
    lt, rt = pa.Table.from_batches(left), pa.Table.from_batches(right)
    lt64 = lt.set_column(0, "lk", lt.column("lk").cast(pa.int64()))  # Acero wants equal key types; ours coerces like the reference
    want = lt64.join(rt, keys="lk", right_keys="rk", join_type="inner" if join_type == JoinType.Inner else "left outer",
                     coalesce_keys=False).select(["lk", "lv", "lrow", "rk", "rv", "rrow"])
    want = want.set_column(0, "lk", want.column("lk").cast(pa.int32()))
    assert_batches_equal(got, want.to_batches(), ordered=False)


def _acero_join(left, right, join_type, lcols, rcols):
    lt, rt = pa.Table.from_batches(left), pa.Table.from_batches(right)
    want = lt.join(rt, keys="lk", right_keys="rk", join_type="inner" if join_type == JoinType.Inner else "left outer",
                   coalesce_keys=False)
    return want.select(lcols + rcols)


@pytest.mark.parametrize("join_type", [JoinType.Inner, JoinType.Left])
@pytest.mark.parametrize("parts", [1, 4, 64])
@pytest.mark.parametrize("dense", [0, 1])
@pytest.mark.parametrize("case", ["inline_payload", "gathered_payload", "key_only", "duplicates", "skew", "marker_key", "marker_value"])
@pytest.mark.parametrize("widths", ["mixed32", "all64"])
def test_join_radix_partitioned(ctx, join_type, parts, dense, case, widths):
    """the radix-partitioned join: region scatter, region tables (direct-addressed for a dense key range, two-slot
    buckets otherwise), probe + emit - and its device-side bail-outs (duplicate build keys, an overflowing region, a key
    equal to the empty marker) which must land on the generic path with the same rows.  `mixed32` tables carry 4-byte
    columns and take the first generation (pjoin_kernels.cu: region-major copies of both sides, compacting probe);
    `all64` tables take the second (rscatter_kernels.cu + pjoin2_kernels.cu: the probe side is scattered straight into
    the output at exact positions, ring of region tables; INNER joins with unmatched probe rows are compacted once).
    Compared with Acero (SQL equi-join, the oracle's Tier-2 rule)."""
    from query_engine_b200.executor import QueryExecutor

    rng = np.random.default_rng(91)
    nl, nr = 50_021, 9_003  # not multiples of any tile
    rk = rng.permutation(np.arange(-nr, 2 * nr, 3, dtype=np.int64))[:nr]  # unique, negative keys included
    lk = rng.integers(-nr, 2 * nr, nl).astype(np.int64)  # two thirds of the probes miss
    if case == "duplicates":
        rk[: nr // 10] = rk[nr // 10: 2 * (nr // 10)]
    if case == "skew":
        lk[: nl // 2] = rk[0]  # half of the probe side lands in one region
    if case == "marker_key":
        rk[5] = -1
        lk[:100] = -1
    lk[-50:] = 10 * nr + np.arange(50)  # probe keys far outside the build key range
    small_int = np.int32 if widths == "mixed32" else np.int64
    small_float = np.float32 if widths == "mixed32" else np.float64
    lcols = [("lk", pa.array(lk)), ("lv", pa.array(rng.normal(size=nl))), ("l32", pa.array(rng.integers(0, 1000, nl).astype(small_int)))]
    rcols = [("rk", pa.array(rk))]
    if case != "key_only":
        rv = rng.normal(size=nr)
        if case == "marker_value":
            rv[7] = np.frombuffer(np.uint64(2**64 - 1).tobytes(), dtype=np.float64)[0]  # the all-ones NaN = the empty marker
        rcols.insert(0, ("rv", pa.array(rv)))  # the key is NOT the first build column
    if case in ("gathered_payload", "duplicates"):
        rcols.append(("r32", pa.array(rng.normal(size=nr).astype(small_float))))
    left = [pa.RecordBatch.from_arrays([a for _, a in lcols], names=[n for n, _ in lcols])]
    right = [pa.RecordBatch.from_arrays([a for _, a in rcols], names=[n for n, _ in rcols])]
    rki = [n for n, _ in rcols].index("rk")
    plan = PhysicalPlan.HashJoin(scan(left), scan(right), join_type, binary(col("lk", 0), B.Equal, col("rk", len(lcols) + rki)))
    ctx.set_option("join.radix_min_build_rows", 1)
    ctx.set_option("join.parts", parts)
    ctx.set_option("join.dense", dense)
    try:
        got = QueryExecutor(ctx=ctx).execute(plan)
    finally:
        ctx.set_option("join.radix_min_build_rows", 1 << 19)
        ctx.set_option("join.parts", 0)
        ctx.set_option("join.dense", 1)
    want = _acero_join(left, right, join_type, [n for n, _ in lcols], [n for n, _ in rcols])
    assert_batches_equal(got, want.to_batches(), ordered=False)


def test_join_radix_no_match_is_empty(ctx):
    from query_engine_b200.executor import QueryExecutor

    left = [pa.RecordBatch.from_arrays([pa.array(np.arange(5000, dtype=np.int64)), pa.array(np.arange(5000) * 0.5)], names=["lk", "lv"])]
    right = [pa.RecordBatch.from_arrays([pa.array(np.arange(10_000, 12_000, dtype=np.int64)), pa.array(np.arange(2000) * 2.0)], names=["rk", "rv"])]
    plan = PhysicalPlan.HashJoin(scan(left), scan(right), JoinType.Inner, binary(col("lk", 0), B.Equal, col("rk", 2)))
    ctx.set_option("join.radix_min_build_rows", 1)
    try:
        assert QueryExecutor(ctx=ctx).execute(plan) == []
    finally:
        ctx.set_option("join.radix_min_build_rows", 1 << 19)


def test_join_sf1_default_path_properties(ctx):
    """SF1 (6 M probe rows, 1.5 M build rows) on the DEFAULT options: the radix-partitioned join with its real region
    geometry.  Checked through size-independent properties: one output row per probe row, equal keys on both sides,
    the multiset of probe rows is the input (checksums), every payload is its order's."""
    from query_engine_b200 import datagen, workloads
    from query_engine_b200.executor import GpuDataSource, QueryExecutor

    li, od = datagen.lineitem(1.0), datagen.orders(1.0)
    L = GpuDataSource(ctx, [datagen.to_batch({n: li[n] for n in workloads.JOIN_L_COLS})])
    O = GpuDataSource(ctx, [datagen.to_batch({n: od[n] for n in workloads.JOIN_O_COLS})])
    for dense in (1, 0):
        ctx.set_option("join.dense", dense)
        try:
            t = pa.Table.from_batches(QueryExecutor(ctx=ctx).execute(workloads.join(L, O)))
        finally:
            ctx.set_option("join.dense", 1)
        assert t.num_rows == len(li["l_orderkey"])
        lk = t.column(0).to_numpy()
        assert np.array_equal(lk, t.column(2).to_numpy())
        assert int(lk.sum()) == int(li["l_orderkey"].sum())
        # (l_orderkey, l_extendedprice) pairs: same multiset as the input
        got = np.sort(lk.astype(np.float64) * 1e7 + t.column(1).to_numpy())
        want = np.sort(li["l_orderkey"].astype(np.float64) * 1e7 + li["l_extendedprice"])
        assert np.array_equal(got, want)
        order = np.argsort(od["o_orderkey"], kind="stable")
        pos = np.searchsorted(od["o_orderkey"][order], lk)
        assert np.array_equal(od["o_totalprice"][order][pos], t.column(3).to_numpy())


@pytest.mark.parametrize("gen2", [1, 0])
def test_join_sf1_missing_orders(ctx, gen2):
    """SF1 with a tenth of the orders deleted (BASELINE config 4's LEFT variant) on the default geometry, both join
    generations: INNER keeps exactly the lines whose order exists (second generation: the output is compacted by the
    match flags), LEFT keeps every line and NULLs the right side of the orphans"""
    from query_engine_b200 import datagen, workloads
    from query_engine_b200.executor import GpuDataSource, QueryExecutor

    li, od = datagen.lineitem(1.0), datagen.orders(1.0)
    keep = (od["o_orderkey"] // 8) % 10 != 0
    od = {k: v[keep] for k, v in od.items()}
    L = GpuDataSource(ctx, [datagen.to_batch({n: li[n] for n in workloads.JOIN_L_COLS})])
    O = GpuDataSource(ctx, [datagen.to_batch({n: od[n] for n in workloads.JOIN_O_COLS})])
    has_order = np.isin(li["l_orderkey"], od["o_orderkey"])
    order = np.argsort(od["o_orderkey"], kind="stable")
    ctx.set_option("join.gen2", gen2)
    try:
        for dense in (1, 0):
            ctx.set_option("join.dense", dense)
            for jt in (JoinType.Inner, JoinType.Left):
                t = pa.Table.from_batches(QueryExecutor(ctx=ctx).execute(workloads.join(L, O, jt)))
                lk = t.column(0).to_numpy()
                rk = t.column(2)
                matched = np.asarray(rk.is_valid()) if jt == JoinType.Left else np.ones(t.num_rows, dtype=bool)
                if jt == JoinType.Inner:
                    assert t.num_rows == int(has_order.sum()) and rk.null_count == 0
                    assert int(lk.sum()) == int(li["l_orderkey"][has_order].sum())
                else:
                    assert t.num_rows == len(has_order) and rk.null_count == int((~has_order).sum())
                    assert int(lk.sum()) == int(li["l_orderkey"].sum())
                    assert t.column(3).null_count == rk.null_count
                    assert not np.isin(lk[~matched], od["o_orderkey"]).any()
                rkv = rk.to_numpy(zero_copy_only=False)
                assert np.array_equal(lk[matched], rkv[matched].astype(np.int64))
                pos = np.searchsorted(od["o_orderkey"][order], lk[matched])
                assert np.array_equal(od["o_totalprice"][order][pos], t.column(3).to_numpy(zero_copy_only=False)[matched])
    finally:
        ctx.set_option("join.gen2", 1)
        ctx.set_option("join.dense", 1)


def test_high_cardinality_group_by_partitioned_properties(executor, ctx):
    """5 M rows over 500 k keys on the default options: the rows are radix-scattered on the group hash and aggregated
    sub-table by sub-table; exact for counts / integer sums / min / max, 1e-9 for the f64 sum"""
    from query_engine_b200 import datagen, workloads
    from query_engine_b200.executor import GpuDataSource

    tb = datagen.highcard(5_000_000, 500_000)
    src = GpuDataSource(ctx, [datagen.to_batch(tb)])
    t = pa.Table.from_batches(executor.execute(workloads.highcard(src)))
    uniq, inv = np.unique(tb["k"], return_inverse=True)
    assert t.num_rows == len(uniq)
    o = np.argsort(t.column(0).to_numpy())
    assert np.array_equal(t.column(0).to_numpy()[o], uniq)
    assert np.array_equal(t.column(5).to_numpy()[o], np.bincount(inv, minlength=len(uniq)))
    sw = np.zeros(len(uniq), dtype=np.int64)
    np.add.at(sw, inv, tb["w"])
    assert np.array_equal(t.column(4).to_numpy()[o], sw)
    mn = np.full(len(uniq), np.inf)
    np.minimum.at(mn, inv, tb["v"])
    mx = np.full(len(uniq), -np.inf)
    np.maximum.at(mx, inv, tb["v"])
    assert np.array_equal(t.column(1).to_numpy()[o], mn) and np.array_equal(t.column(2).to_numpy()[o], mx)
    sv = np.bincount(inv, weights=tb["v"], minlength=len(uniq))
    assert np.allclose(t.column(3).to_numpy()[o], sv, rtol=F64_RTOL, atol=0.0)


def _highcard_reference(tb):
    import pandas as pd

    df = pd.DataFrame(tb)
    gb = df.groupby("k", sort=True)
    return pa.table({"k": gb.v.min().index.values.astype(np.int64), "col_0": gb.v.min().values, "col_1": gb.v.max().values,
                     "col_2": gb.v.sum().values, "col_3": gb.w.sum().values.astype(np.int64),
                     "col_4": gb.k.count().values.astype(np.int64)})


RADIX_GROUP_CASES = {
    # name: (rows, keys, options) - every case must give the same groups whichever path ends up producing them
    "two_levels": (300_000, 20_000, {"group.parts2": 16}),
    "one_level": (200_000, 15_000, {"group.parts2": 1, "group.part_slots": 1024}),
    "wide_first_level": (300_000, 20_000, {"group.bits1": 8, "group.parts2": 4}),
    "output_arrays_grow": (300_000, 20_000, {"group.parts2": 16, "group.out_cap": 1000}),
    "table_full_falls_back": (300_000, 100_000, {"group.parts2": 2, "group.part_slots": 256}),
}


@pytest.mark.parametrize("case", sorted(RADIX_GROUP_CASES))
def test_group_by_radix_partitioned(executor, ctx, case):
    """ops_groupby.cu on small inputs: two radix levels + shared-memory tables, every geometry and bail-out forced
    through options; -1 (the table's empty marker) is one of the keys"""
    from query_engine_b200 import datagen, workloads
    from query_engine_b200.executor import GpuDataSource

    rows, nkeys, opts = RADIX_GROUP_CASES[case]
    tb = datagen.highcard(rows, nkeys)
    tb["k"][::1000] = -1
    tb["k"][1::5000] = np.iinfo(np.int64).min
    ctx.set_option("group.partition_min_rows", 1)
    for k, v in opts.items():
        ctx.set_option(k, v)
    try:
        got = executor.execute(workloads.highcard(GpuDataSource(ctx, [datagen.to_batch(tb)])))
    finally:
        ctx.set_option("group.partition_min_rows", 1 << 22)
        for k in opts:
            ctx.set_option(k, 0)
    assert_batches_equal(got, _highcard_reference(tb).to_batches(), ordered=False, approx_cols=[3], check_names=False)


def test_group_by_radix_partitioned_skew_and_predicate(executor, ctx):
    """a key that holds half of the rows overflows its region (the scatter raises its flag, the HBM-table path takes
    over); a predicate and computed aggregate inputs go through the partitioned pass like through any other scan"""
    from query_engine_b200 import datagen
    from query_engine_b200.executor import GpuDataSource

    tb = datagen.highcard(400_000, 30_000)
    names = ["k", "v", "w"]
    pred = binary(col("w", 2), B.Greater, lit(-500_000))
    aggs = [AggregateExpr(F.Sum, binary(col("v", 1), B.Multiply, lit(2.0))), AggregateExpr(F.Max, col("w", 2)),
            AggregateExpr(F.Count, col("v", 1)), AggregateExpr(F.Avg, col("w", 2))]
    for skew in (False, True):
        t = {n: tb[n].copy() for n in names}
        if skew:
            t["k"][::2] = 42
        b = datagen.to_batch(t)
        plan = lambda src: PhysicalPlan.HashAggregate(PhysicalPlan.Filter(PhysicalPlan.Scan(src), pred), [col("k", 0)], aggs)  # noqa: E731
        ctx.set_option("group.partition_min_rows", 1)
        ctx.set_option("group.parts2", 8)
        try:
            got = executor.execute(plan(GpuDataSource(ctx, [b])))
        finally:
            ctx.set_option("group.partition_min_rows", 1 << 22)
            ctx.set_option("group.parts2", 0)
        want = oracle.execute(plan(MemoryDataSource(b.schema, [b])))
        assert_batches_equal(got, want, ordered=False, approx_cols=[1, 4])


@pytest.mark.parametrize("skip,fetch", [(0, 10), (5, None), (1000, 1500), (2047, 2), (4999, 5), (5000, 3), (0, 0), (7, 10**9), (10**9, 5)])
def test_limit_matches_oracle(executor, skip, fetch):
    """Limit (executor.rs:299-341): batch-wise slicing; also above a Filter (the shape `SELECT .. WHERE .. LIMIT n`)"""
    batches = make_batches(21, 5000, 0.1, sizes=[1024, 1024, 1, 0, 2951])
    plan = PhysicalPlan.Limit(scan(batches), skip, fetch)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True, same_batching=True)
    plan = PhysicalPlan.Limit(PhysicalPlan.Filter(scan(batches), PREDICATES["and_chain"]), skip, fetch)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=True, same_batching=True)


@pytest.mark.parametrize("null_frac", [0.0, 0.2])
@pytest.mark.parametrize("with_filter", [False, True])
def test_group_by_computed_expressions(executor, null_frac, with_filter):
    """physical_plan.rs:35-39: group_exprs are arbitrary PhysicalExprs - a computed key (named group_i), mixed with a plain
    column key, under an optional fused Filter"""
    batches = make_batches(31, 20_000, null_frac, sizes=[8192, 8192, 3616])
    groups = [binary(col("a", 0), B.Modulo, lit(7)), col("g", 5), binary(col("c", 2), B.Greater, lit(ScalarValue.Int32(0)))]
    aggs = [AggregateExpr(F.Count, col("a", 0)), AggregateExpr(F.Sum, col("b", 1)), AggregateExpr(F.Min, col("d", 3))]
    inp = PhysicalPlan.Filter(scan(batches), PREDICATES["and_chain"]) if with_filter else scan(batches)
    plan = PhysicalPlan.HashAggregate(inp, groups, aggs)
    assert_batches_equal(executor.execute(plan), oracle.execute(plan), ordered=False, approx_cols=[4])


def test_group_by_computed_expression_error_only_on_filtered_rows(executor):
    """the group expression is evaluated on the filtered batches: an overflow on a row the Filter drops is no error"""
    b = pa.RecordBatch.from_arrays([pa.array([2**62, 1, 2, 3], type=pa.int64()), pa.array([0, 1, 1, 1], type=pa.int64())], names=["a", "keep"])
    grp = [binary(col("a", 0), B.Multiply, lit(4))]
    ok = PhysicalPlan.HashAggregate(PhysicalPlan.Filter(scan([b]), binary(col("keep", 1), B.Equal, lit(1))), grp, [AggregateExpr(F.Count, col("a", 0))])
    assert_batches_equal(executor.execute(ok), oracle.execute(ok), ordered=False)
    bad = PhysicalPlan.HashAggregate(scan([b]), grp, [AggregateExpr(F.Count, col("a", 0))])
    with pytest.raises(QueryError) as want:
        oracle.execute(bad)
    with pytest.raises(QueryError) as got:
        executor.execute(bad)
    assert str(got.value) == str(want.value)
