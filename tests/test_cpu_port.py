"""CPU: the C++ port used as the timed CPU baseline (oracle/cpu_ref.cpp) agrees with the numpy oracle."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle
from query_engine_b200 import MemoryDataSource, PhysicalPlan, _ffi, datagen, workloads

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def port():
    subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    lib = C.CDLL(str(ROOT / "oracle" / "libcpuref.so"))
    lib.cpuref_last_error.restype = C.c_char_p
    lib.cpuref_filter_aggregate.restype = C.c_int32
    lib.cpuref_filter_aggregate.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cpuref_hash_join.restype = C.c_int64
    lib.cpuref_hash_join.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    return lib


def run_port(lib, cols, names, plan, batch_rows, threads=1, cap=4096):
    arrs = [np.ascontiguousarray(cols[n]) for n in names]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    types = (C.c_int32 * len(arrs))(*[3 if a.dtype == np.int64 else 5 for a in arrs])
    pred = plan.input.predicate if isinstance(plan.input, PhysicalPlan.Filter) else None
    fp = _ffi.FlatExpr(pred) if pred is not None else None
    g, a = _ffi.FlatExprArray(plan.group_exprs), _ffi.FlatAggArray(plan.aggr_exprs)
    keys = np.zeros(cap * max(1, g.n), dtype=np.uint64)
    vals = np.zeros(cap * a.n, dtype=np.uint64)
    valid = np.zeros(cap * a.n, dtype=np.uint8)
    ng = C.c_int64()
    rc = lib.cpuref_filter_aggregate(C.addressof(ptrs), C.addressof(types), len(arrs), len(arrs[0]), batch_rows,
                                     C.addressof(fp.c) if fp else None, C.addressof(g.array), g.n, C.addressof(a.array), a.n, threads,
                                     cap, keys.ctypes.data, vals.ctypes.data, valid.ctypes.data, C.addressof(ng))
    assert rc == 0, lib.cpuref_last_error()
    n = ng.value
    return n, keys[: max(n, 0) * max(1, g.n)].reshape(-1, max(1, g.n)), vals[: max(n, 0) * a.n].reshape(-1, a.n), valid


@pytest.mark.parametrize("threads", [1, 4])
def test_q6_port_matches_oracle(port, threads):
    li = datagen.lineitem(0.005)
    plan = workloads.q6(MemoryDataSource(None, datagen.to_batches(li, 1024, workloads.Q6_COLS)))
    want = oracle.execute(plan)[0]
    n, _, vals, valid = run_port(port, li, workloads.Q6_COLS, plan, 1024, threads)
    assert n == 1 and valid[0] == 1 and valid[1] == 1
    assert vals[0, 1].view(np.int64) == want.column(1)[0].as_py()
    assert vals[0, 0:1].view(np.float64)[0] == pytest.approx(want.column(0)[0].as_py(), rel=1e-12)


def test_q1_port_matches_oracle(port):
    li = datagen.lineitem(0.005)
    plan = workloads.q1(MemoryDataSource(None, datagen.to_batches(li, 8192, workloads.Q1_COLS)))
    want = pa.Table.from_batches(oracle.execute(plan))
    n, keys, vals, _ = run_port(port, li, workloads.Q1_COLS, plan, 8192)
    assert n == want.num_rows == 4
    got = {(int(k[0]), int(k[1])): v for k, v in zip(keys.view(np.int64), vals)}
    for r in range(4):
        v = got[(want.column(0)[r].as_py(), want.column(1)[r].as_py())]
        assert v[7:8].view(np.int64)[0] == want.column(9)[r].as_py()  # COUNT bit-exact
        for j in range(7):
            assert v[j : j + 1].view(np.float64)[0] == pytest.approx(want.column(2 + j)[r].as_py(), rel=1e-9)


def test_filter_that_keeps_nothing_is_ok_vec_empty(port):
    li = datagen.lineitem(0.001)
    from query_engine_b200 import AggregateExpr, AggregateFunction as F, BinaryOp as B, binary, col, lit

    plan = PhysicalPlan.HashAggregate(PhysicalPlan.Filter(PhysicalPlan.Scan(None), binary(col("l_quantity", 1), B.Greater, lit(1e9))), [],
                                      [AggregateExpr(F.Count, col("l_orderkey", 0))])
    n, _, _, _ = run_port(port, li, workloads.Q6_COLS, plan, 1024)
    assert n == -1


@pytest.mark.parametrize("join_type", [0, 1])
def test_join_port_matches_oracle(port, join_type):
    rng = np.random.default_rng(9)
    lk = rng.integers(0, 400, 2000, dtype=np.int64)
    rk = rng.integers(100, 500, 500, dtype=np.int64)
    cap = 20000
    ol, orr = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64)
    n = port.cpuref_hash_join(lk.ctypes.data, len(lk), rk.ctypes.data, len(rk), join_type, cap, ol.ctypes.data, orr.ctypes.data)
    from query_engine_b200 import BinaryOp as B, JoinType, binary, col

    left = [pa.RecordBatch.from_arrays([pa.array(lk), pa.array(np.arange(len(lk)))], names=["lk", "li"])]
    right = [pa.RecordBatch.from_arrays([pa.array(rk), pa.array(np.arange(len(rk)))], names=["rk", "ri"])]
    plan = PhysicalPlan.HashJoin(PhysicalPlan.Scan(MemoryDataSource(None, left)), PhysicalPlan.Scan(MemoryDataSource(None, right)),
                                 JoinType.Inner if join_type == 0 else JoinType.Left, binary(col("lk", 0), B.Equal, col("rk", 2)))
    want = pa.Table.from_batches(oracle.execute(plan))
    assert n == want.num_rows
    got = sorted(zip(ol[:n].tolist(), [None if x < 0 else x for x in orr[:n].tolist()]), key=lambda t: (t[0], -1 if t[1] is None else t[1]))
    exp = sorted(zip(want.column("li").to_pylist(), want.column("ri").to_pylist()), key=lambda t: (t[0], -1 if t[1] is None else t[1]))
    assert got == exp
