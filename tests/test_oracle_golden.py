"""CPU: the oracle against the committed golden vectors (tests/golden/, see make_golden.py for provenance)."""
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
    PhysicalPlan,
    binary,
    col,
    lit,
)
from tests.helpers import assert_batches_equal

GOLDEN = Path(__file__).resolve().parent / "golden"


def employees_batches():
    """employees.csv the way the reference's front-ends read it: dept_id is Utf8 because one row holds the text NULL;
    arrow-csv treats only the empty string as null."""
    t = pacsv.read_csv(GOLDEN / "employees.csv", convert_options=pacsv.ConvertOptions(
        column_types={"id": pa.int64(), "name": pa.string(), "age": pa.int64(), "salary": pa.int64(), "dept_id": pa.string()},
        null_values=[""], strings_can_be_null=True))
    return t.to_batches()


def golden():
    return json.loads((GOLDEN / "employees_golden.json").read_text())


def golden_table(entry):
    types = {"int64": pa.int64(), "double": pa.float64(), "string": pa.string()}
    cols = list(zip(*entry["rows"]))
    return pa.table([pa.array(c, type=types[t]) for c, t in zip(cols, entry["types"])], names=entry["columns"])


def scan():
    b = employees_batches()
    return PhysicalPlan.Scan(MemoryDataSource(b[0].schema, b))


AGE_GT_25 = binary(col("age", 2), B.Greater, lit(25))


def tier1_plan():
    aggs = [AggregateExpr(F.Count, col("id", 0)), AggregateExpr(F.Sum, col("salary", 3)), AggregateExpr(F.Avg, col("salary", 3)),
            AggregateExpr(F.Min, col("age", 2)), AggregateExpr(F.Max, col("age", 2))]
    return PhysicalPlan.HashAggregate(PhysicalPlan.Filter(scan(), AGE_GT_25), [], aggs)


def tier2_plan():
    aggs = [AggregateExpr(F.Count, col("id", 0)), AggregateExpr(F.Sum, col("salary", 3)), AggregateExpr(F.Avg, col("salary", 3))]
    return PhysicalPlan.HashAggregate(PhysicalPlan.Filter(scan(), AGE_GT_25), [col("dept_id", 4)], aggs)


def test_filter_golden():
    out = oracle.execute(PhysicalPlan.Filter(scan(), AGE_GT_25))
    assert pa.Table.from_batches(out).column("id").to_pylist() == golden()["filter_age_gt_25"]["ids"]


def test_tier1_scalar_golden():
    g = golden()["tier1_scalar"]
    assert_batches_equal(oracle.execute(tier1_plan()), golden_table(g).to_batches(), ordered=True)


def test_tier2_group_by_golden_and_asis():
    g = golden()["tier2_group_by"]
    assert_batches_equal(oracle.execute(tier2_plan()), golden_table(g).to_batches(), ordered=False)
    # what the Rust literally returns for a grouped aggregate: Ok(vec![]) (executor.rs:189)
    assert oracle.execute(tier2_plan(), reference_asis=True) == g["reference_asis_rows"]


def test_siphash13_matches_cpython():
    """Rust DefaultHasher = SipHash-1-3 with key (0, 0); CPython's siphash13 with PYTHONHASHSEED=0 is the same
    function, implemented independently."""
    g = json.loads((GOLDEN / "siphash13_golden.json").read_text())
    for hx, want in g["vectors"]:
        data = bytes.fromhex(hx)
        assert oracle.siphash_bytes(data, 1, 3) == want, hx
    # the vectorised form used for Int64 / Int32 keys agrees with the byte form
    vals = np.array([1, -1, 42, 2**40 + 7], dtype=np.int64)
    got = oracle.siphash_stream([vals.view(np.uint64)], [8])
    for v, h in zip(vals, got):
        assert int(h) == oracle.siphash_bytes(int(v).to_bytes(8, "little", signed=True), 1, 3)


def test_siphash24_reference_vectors():
    """First entries of the official SipHash-2-4 test vectors (key 00..0f, message 00..n-1): pins the skeleton."""
    k0, k1 = 0x0706050403020100, 0x0F0E0D0C0B0A0908
    want = [0x726FDB47DD0E0E31, 0x74F839C593DC67FD, 0x0D6C8009D9A94F5A, 0x85676696D7FB7E2D]
    for n, w in enumerate(want):
        assert oracle.siphash_bytes(bytes(range(n)), 2, 4, k0, k1) == w


def test_exchange_row_conservation_like_the_reference_tests():
    """query-distributed/src/partition.rs:401-415 and operators.rs:307-316 assert only partition count and row
    conservation; same assertions on the oracle, plus: a row's partition is a function of its key."""
    b = pa.RecordBatch.from_arrays([pa.array([1, 2, 3, 4, 5], type=pa.int64()), pa.array(list("abcde"))], names=["id", "name"])
    parts = oracle.partition_by_hash([b], ["id"], 3)
    assert len(parts) == 3
    assert sum(x.num_rows for p in parts for x in p) == 5
    parts4 = oracle.partition_by_hash([b], ["id"], 4)
    assert len(parts4) == 4 and sum(x.num_rows for p in parts4 for x in p) == 5
    merged = oracle.concat_merge(parts)
    assert sum(x.num_rows for x in merged) == 5
    with pytest.raises(Exception, match="No key columns found in batch"):
        oracle.partition_by_hash([b], ["missing"], 3)
