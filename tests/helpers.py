"""Comparison helpers of the parity tests: GPU result vs oracle result."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import pyarrow as pa

F64_RTOL = 1e-9  # north_star: 1e-9 relative for f64 SUM/AVG; everything else bit-exact
# SUM/AVG over a Float32 column: arrow-rs accumulates in f32 in an unspecified lane order, so the reference's own
# result carries ~n*2^-24 relative error; libqgpu accumulates in f64 and rounds once to f32.  Compared at 1e-4.
# (rtol, atol): the absolute term covers cancellation in small groups (|values| <= ~400 in the test tables).
F32_SUM_RTOL = (1e-4, 2e-2)


def table_of(batches: Sequence[pa.RecordBatch]) -> Optional[pa.Table]:
    batches = list(batches)
    if not batches:
        return None
    return pa.Table.from_batches(batches)


def _key(v):
    # total order over None / numbers / strings / NaN for order normalisation
    if v is None:
        return (0, 0)
    if isinstance(v, float):
        if math.isnan(v):
            return (3, 0)
        return (1, v)
    if isinstance(v, (int, bool)):
        return (1, v)
    return (2, v)


def rows_of(t: pa.Table) -> List[tuple]:
    cols = [c.to_pylist() for c in t.columns]
    return list(zip(*cols)) if cols else []


def _same_float(a: float, b: float, rtol) -> bool:
    atol = 0.0
    if isinstance(rtol, tuple):
        rtol, atol = rtol
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    if rtol == 0.0 and atol == 0.0:
        return np.float64(a).tobytes() == np.float64(b).tobytes()
    return abs(a - b) <= rtol * max(abs(a), abs(b)) + atol


def assert_tables_equal(got: Optional[pa.Table], want: Optional[pa.Table], *, ordered: bool, approx_cols: Sequence[int] = (),
                        check_names: bool = True):
    """Row-set equality.  Columns listed in `approx_cols` are f64 SUM/AVG outputs compared to F64_RTOL; all other
    columns (ints, keys, counts, copied floats) must match exactly (floats bitwise)."""
    if want is None or want.num_rows == 0:
        assert got is None or got.num_rows == 0, f"expected no rows, got {got.num_rows}"
        return
    assert got is not None, "expected rows, got an empty result"
    assert got.num_columns == want.num_columns, (got.schema, want.schema)
    for i in range(want.num_columns):
        assert got.schema.field(i).type == want.schema.field(i).type, (i, got.schema, want.schema)
        if check_names:
            assert got.schema.field(i).name == want.schema.field(i).name, (i, got.schema.names, want.schema.names)
    assert got.num_rows == want.num_rows, (got.num_rows, want.num_rows)
    g, w = rows_of(got), rows_of(want)
    if not ordered:
        exact = [i for i in range(want.num_columns) if i not in approx_cols]
        keyf = lambda r: tuple(_key(r[i]) for i in exact) + tuple(_key(r[i]) for i in approx_cols)  # noqa: E731
        g, w = sorted(g, key=keyf), sorted(w, key=keyf)
    for ri, (a, b) in enumerate(zip(g, w)):
        for ci, (x, y) in enumerate(zip(a, b)):
            if x is None or y is None:
                assert x is None and y is None, f"row {ri} col {ci}: {x!r} vs {y!r}"
            elif isinstance(y, float):
                rtol = 0.0
                if ci in approx_cols:
                    rtol = approx_cols[ci] if isinstance(approx_cols, dict) else F64_RTOL
                assert _same_float(x, y, rtol), f"row {ri} col {ci}: {x!r} vs {y!r}"
            else:
                assert x == y, f"row {ri} col {ci}: {x!r} vs {y!r}"


def assert_batches_equal(got: Sequence[pa.RecordBatch], want: Sequence[pa.RecordBatch], *, ordered=True, approx_cols=(),
                         same_batching=False, check_names=True):
    got, want = list(got), list(want)
    if same_batching:
        assert [b.num_rows for b in got] == [b.num_rows for b in want]
    assert_tables_equal(table_of(got), table_of(want), ordered=ordered, approx_cols=approx_cols, check_names=check_names)


def random_table(rng: np.random.Generator, n: int, spec, null_frac: float = 0.0) -> pa.RecordBatch:
    """spec: list of (name, kind) with kind in i64 / i32 / f64 / f32 / bool / str / smallint (few distinct) / key."""
    arrays, names = [], []
    for name, kind in spec:
        if kind == "i64":
            v = rng.integers(-1000, 1000, n, dtype=np.int64)
        elif kind == "i32":
            v = rng.integers(-1000, 1000, n, dtype=np.int32)
        elif kind == "f64":
            v = np.round(rng.normal(0, 100, n), 3)
        elif kind == "f32":
            v = np.round(rng.normal(0, 100, n), 2).astype(np.float32)
        elif kind == "bool":
            v = rng.integers(0, 2, n).astype(bool)
        elif kind == "smallint":
            v = rng.integers(0, 5, n, dtype=np.int64)
        elif kind == "key":
            v = rng.integers(0, max(2, n // 3), n, dtype=np.int64)
        elif kind == "str":
            v = np.array([f"s{x}" for x in rng.integers(0, 7, n)], dtype=object)
        else:
            raise ValueError(kind)
        mask = None
        if null_frac > 0:
            mask = rng.random(n) < null_frac
        arrays.append(pa.array(v, mask=mask) if kind != "str" else pa.array(list(v), type=pa.string(), mask=mask))
        names.append(name)
    return pa.RecordBatch.from_arrays(arrays, names=names)


def split_batches(b: pa.RecordBatch, sizes: Sequence[int]) -> List[pa.RecordBatch]:
    out, o = [], 0
    for s in sizes:
        out.append(b.slice(o, s))
        o += s
    assert o == b.num_rows
    return out
