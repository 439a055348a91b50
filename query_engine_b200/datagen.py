"""Synthetic TPC-H-shaped inputs of the benchmark configs (BASELINE.json `configs`).

All numeric columns are Int64 / Float64 (8 bytes), non-null.  Generation is numpy `Generator(PCG64(seed))`; a shard
(`shard`, `n_shards`) is the slice of the table one rank of a multi-GPU run owns: orders are split by canonical
order index, and a lineitem shard holds the lines of its own orders, so shards can be generated independently and
their union is a valid (every line has exactly one parent order) instance of the schema.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import pyarrow as pa

ORDERS_PER_SF = 1_500_000
LINEITEM_PER_SF = 6_000_000

# Q6 / Q1 constants (days since epoch)
Q6_DATE_LO = 8766  # 1994-01-01
Q6_DATE_HI = 9131  # 1995-01-01
Q1_DATE_MAX = 10471  # 1998-09-02
LINESTATUS_CUTOFF = 9298  # 1995-06-17


def _order_range(sf: float, shard: int, n_shards: int) -> Tuple[int, int]:
    n_o = int(round(ORDERS_PER_SF * sf))
    lo = n_o * shard // n_shards
    hi = n_o * (shard + 1) // n_shards
    return lo, hi


def _orderkey(i: np.ndarray) -> np.ndarray:
    """TPC-H sparse order keys: 8 used out of every 32."""
    return (i // 8) * 32 + i % 8 + 1


def orders(sf: float, shard: int = 0, n_shards: int = 1, seed: int = 1001) -> Dict[str, np.ndarray]:
    lo, hi = _order_range(sf, shard, n_shards)
    rng = np.random.Generator(np.random.PCG64([seed, shard, n_shards]))
    i = np.arange(lo, hi, dtype=np.int64)
    key = _orderkey(i)
    rng.shuffle(key)
    n = hi - lo
    return {
        "o_orderkey": key,
        "o_custkey": rng.integers(1, max(2, int(150_000 * sf) + 1), n, dtype=np.int64),
        "o_totalprice": np.round(rng.uniform(900.0, 500000.0, n), 2),
    }


def lineitem(sf: float, shard: int = 0, n_shards: int = 1, seed: int = 1002) -> Dict[str, np.ndarray]:
    lo, hi = _order_range(sf, shard, n_shards)
    n_total = int(round(LINEITEM_PER_SF * sf))
    n = n_total * (shard + 1) // n_shards - n_total * shard // n_shards
    rng = np.random.Generator(np.random.PCG64([seed, shard, n_shards]))
    n_o = hi - lo
    if n_o <= 0:
        raise ValueError("scale factor too small for this shard count")
    per = rng.integers(1, 8, n_o, dtype=np.int64)  # 1..7 lines per order
    idx = np.repeat(np.arange(lo, hi, dtype=np.int64), per)
    if len(idx) >= n:
        idx = idx[:n]
    else:
        idx = np.concatenate([idx, rng.integers(lo, hi, n - len(idx), dtype=np.int64)])
    rng.shuffle(idx)
    shipdate = rng.integers(8036, 10562, n, dtype=np.int64)
    linestatus = (shipdate > LINESTATUS_CUTOFF).astype(np.int64)  # 0 = F, 1 = O
    receipt = shipdate + rng.integers(1, 31, n, dtype=np.int64)
    returnflag = np.where(receipt > LINESTATUS_CUTOFF, 2, rng.integers(0, 2, n, dtype=np.int64)).astype(np.int64)
    # keep exactly the four TPC-H groups A/F, R/F, N/F, N/O: a line shipped after the cutoff is always N
    returnflag = np.where(linestatus == 1, 2, returnflag).astype(np.int64)
    return {
        "l_orderkey": _orderkey(idx),
        "l_quantity": rng.integers(1, 51, n).astype(np.float64),
        "l_extendedprice": np.round(rng.uniform(901.0, 104950.0, n), 2),
        "l_discount": rng.integers(0, 11, n).astype(np.float64) / 100.0,
        "l_tax": rng.integers(0, 9, n).astype(np.float64) / 100.0,
        "l_shipdate": shipdate,
        "l_returnflag": returnflag,
        "l_linestatus": linestatus,
    }


def highcard(n_rows: int, n_keys: int, shard: int = 0, n_shards: int = 1, seed: int = 1005, with_index: bool = False) -> Dict[str, np.ndarray]:
    """Config 5: `n_keys` distinct int64 keys (splitmix64 of the key index), rows drawn uniformly.  `with_index` also
    returns the key index of every row ("j": lets a checker count distinct keys with one bincount)."""
    rng = np.random.Generator(np.random.PCG64([seed, shard, n_shards]))
    n = n_rows * (shard + 1) // n_shards - n_rows * shard // n_shards
    j = rng.integers(0, n_keys, n, dtype=np.int64).astype(np.uint64)
    t = {
        "k": splitmix64(j).view(np.int64),
        "v": rng.uniform(0.0, 1e6, n),
        "w": rng.integers(-1_000_000, 1_000_001, n, dtype=np.int64),
    }
    if with_index:
        t["j"] = j.astype(np.int64)
    return t


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def to_batch(cols: Dict[str, np.ndarray], names=None) -> pa.RecordBatch:
    names = list(cols) if names is None else list(names)
    return pa.RecordBatch.from_arrays([pa.array(cols[n]) for n in names], names=names)


def to_batches(cols: Dict[str, np.ndarray], batch_rows: int, names=None):
    """The same table cut into `batch_rows`-row RecordBatches (zero-copy slices)."""
    b = to_batch(cols, names)
    return [b.slice(o, min(batch_rows, b.num_rows - o)) for o in range(0, b.num_rows, batch_rows)] or [b]
