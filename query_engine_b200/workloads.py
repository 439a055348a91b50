"""The benchmark queries of BASELINE.json `configs`, as the PhysicalPlan trees the reference's lowering
(crates/query-pgwire/src/backend.rs:614-724) would hand to `QueryExecutor::execute`.

Literals are typed the way the planner types SQL literals ('.' => Float64, else Int64;
crates/query-planner/src/planner.rs:669-686) — note `1.0 - l_discount`, not `1 - l_discount`: Int64 - Float64 is an
error in the reference (operators.rs:384-507).
"""
from __future__ import annotations

from .datagen import Q1_DATE_MAX, Q6_DATE_HI, Q6_DATE_LO
from .physical_plan import (
    AggregateExpr,
    AggregateFunction as F,
    BinaryOp as B,
    DataSource,
    JoinType,
    PhysicalPlan,
    and_all,
    binary,
    col,
    lit,
)

LINEITEM_COLS = [
    "l_orderkey",
    "l_quantity",
    "l_extendedprice",
    "l_discount",
    "l_tax",
    "l_shipdate",
    "l_returnflag",
    "l_linestatus",
]
Q6_COLS = ["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_shipdate"]
Q1_COLS = ["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate", "l_returnflag", "l_linestatus"]

# algorithmic bytes per input row: each referenced column read once, 8 B each
Q6_BYTES_PER_ROW = 4 * 8   # shipdate, discount, quantity, extendedprice (COUNT(l_orderkey) of a non-null column reads nothing)
Q1_BYTES_PER_ROW = 7 * 8   # shipdate, returnflag, linestatus, quantity, extendedprice, discount, tax
JOIN_BYTES_PER_PROBE_ROW_SF = 52  # (16*(nL+nR) + 32*n_out)/nL at the TPC-H 4:1 ratio
HIGHCARD_BYTES_PER_ROW = 3 * 8


def _c(names, n):
    return col(n, names.index(n))


def q6(source: DataSource, names=Q6_COLS) -> PhysicalPlan:
    """SELECT SUM(l_extendedprice * l_discount), COUNT(l_orderkey) FROM lineitem
    WHERE l_shipdate >= 8766 AND l_shipdate < 9131 AND l_discount >= 0.05 AND l_discount <= 0.07 AND l_quantity < 24.0"""
    pred = and_all(
        [
            binary(_c(names, "l_shipdate"), B.GreaterEqual, lit(Q6_DATE_LO)),
            binary(_c(names, "l_shipdate"), B.Less, lit(Q6_DATE_HI)),
            binary(_c(names, "l_discount"), B.GreaterEqual, lit(0.05)),
            binary(_c(names, "l_discount"), B.LessEqual, lit(0.07)),
            binary(_c(names, "l_quantity"), B.Less, lit(24.0)),
        ]
    )
    aggs = [
        AggregateExpr(F.Sum, binary(_c(names, "l_extendedprice"), B.Multiply, _c(names, "l_discount"))),
        AggregateExpr(F.Count, _c(names, "l_orderkey")),
    ]
    return PhysicalPlan.HashAggregate(PhysicalPlan.Filter(PhysicalPlan.Scan(source), pred), [], aggs)


def q1(source: DataSource, names=Q1_COLS) -> PhysicalPlan:
    """SELECT l_returnflag, l_linestatus, SUM(l_quantity), SUM(l_extendedprice),
    SUM(l_extendedprice*(1.0-l_discount)), SUM(l_extendedprice*(1.0-l_discount)*(1.0+l_tax)),
    AVG(l_quantity), AVG(l_extendedprice), AVG(l_discount), COUNT(l_orderkey)
    FROM lineitem WHERE l_shipdate <= 10471 GROUP BY l_returnflag, l_linestatus"""
    pred = binary(_c(names, "l_shipdate"), B.LessEqual, lit(Q1_DATE_MAX))
    price, disc, tax, qty = (_c(names, n) for n in ("l_extendedprice", "l_discount", "l_tax", "l_quantity"))
    disc_price = binary(price, B.Multiply, binary(lit(1.0), B.Subtract, disc))
    charge = binary(disc_price, B.Multiply, binary(lit(1.0), B.Add, tax))
    aggs = [
        AggregateExpr(F.Sum, qty),
        AggregateExpr(F.Sum, price),
        AggregateExpr(F.Sum, disc_price),
        AggregateExpr(F.Sum, charge),
        AggregateExpr(F.Avg, qty),
        AggregateExpr(F.Avg, price),
        AggregateExpr(F.Avg, disc),
        AggregateExpr(F.Count, _c(names, "l_orderkey")),
    ]
    groups = [_c(names, "l_returnflag"), _c(names, "l_linestatus")]
    return PhysicalPlan.HashAggregate(PhysicalPlan.Filter(PhysicalPlan.Scan(source), pred), groups, aggs)


JOIN_L_COLS = ["l_orderkey", "l_extendedprice"]
JOIN_O_COLS = ["o_orderkey", "o_totalprice"]


def join(lineitem_src: DataSource, orders_src: DataSource, join_type: JoinType = JoinType.Inner) -> PhysicalPlan:
    """lineitem(l_orderkey, l_extendedprice) JOIN orders(o_orderkey, o_totalprice) ON l_orderkey = o_orderkey;
    left = lineitem (probe), right = orders (build).  `on` indexes the merged left++right schema (planner.rs:148-157)."""
    on = binary(col("l_orderkey", 0), B.Equal, col("o_orderkey", len(JOIN_L_COLS) + 0))
    return PhysicalPlan.HashJoin(PhysicalPlan.Scan(lineitem_src), PhysicalPlan.Scan(orders_src), join_type, on)


HIGHCARD_COLS = ["k", "v", "w"]


def highcard(source: DataSource) -> PhysicalPlan:
    """SELECT k, MIN(v), MAX(v), SUM(v), SUM(w), COUNT(k) FROM t GROUP BY k"""
    k, v, w = (col(n, i) for i, n in enumerate(HIGHCARD_COLS))
    aggs = [
        AggregateExpr(F.Min, v),
        AggregateExpr(F.Max, v),
        AggregateExpr(F.Sum, v),
        AggregateExpr(F.Sum, w),
        AggregateExpr(F.Count, k),
    ]
    return PhysicalPlan.HashAggregate(PhysicalPlan.Scan(source), [k], aggs)


def filter_only(source: DataSource, names, which: str = "q6") -> PhysicalPlan:
    """The materialising Filter alone (executor.rs:131-155: every column of the surviving rows is copied) with the Q6
    predicate (~1.9 % of the rows survive) or the Q1 predicate (~98 %)."""
    plan = q6(source, names) if which == "q6" else q1(source, names)
    return plan.input  # Filter(Scan)
