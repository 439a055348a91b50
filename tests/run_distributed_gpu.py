"""Launched under torchrun (one rank per GPU) by tests/test_distributed_gpu.py and by hand on a multi-GPU box.

Every rank generates its shard, rank 0 additionally holds the union; the distributed result (aggregate: identical on
all ranks; join: concatenation over ranks) must equal the oracle's result on the union."""
import os
import sys
from pathlib import Path

import numpy as np
import pyarrow as pa

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import oracle  # noqa: E402
from query_engine_b200 import MemoryDataSource, _ffi, datagen, workloads  # noqa: E402
from query_engine_b200.distributed import DistributedExecutor  # noqa: E402
from query_engine_b200.executor import QueryExecutor  # noqa: E402
from tests.helpers import assert_batches_equal  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _ffi.Context(local)
    ids = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)
    dex = DistributedExecutor(QueryExecutor(ctx=ctx))
    dex_plan = DistributedExecutor(QueryExecutor(ctx=ctx), plan_level=True)
    sf = 0.01

    def shard(gen, names):
        mine = datagen.to_batches({n: gen(sf, rank, world)[n] for n in names}, 4096)
        union = [b for r in range(world) for b in datagen.to_batches({n: gen(sf, r, world)[n] for n in names}, 4096)]
        return mine, union

    # Q6 (scalar), Q1 (grouped) on both scan-kernel forms
    for jit in (0, 1):
        ctx.set_option("scan.jit", jit)
        for name, cols, mk, approx in (("q6", workloads.Q6_COLS, workloads.q6, [0]),
                                       ("q1", workloads.Q1_COLS, workloads.q1, [2, 3, 4, 5, 6, 7, 8])):
            mine, union = shard(datagen.lineitem, cols)
            want = oracle.execute(mk(MemoryDataSource(union[0].schema, union), cols))
            for d in (dex, dex_plan):  # fused in-library merge, and the three-operator plan-level form
                got = d.execute(mk(MemoryDataSource(mine[0].schema, mine), cols))
                assert_batches_equal(got, want, ordered=False, approx_cols=approx)
            if rank == 0:
                print(f"rank0: distributed {name} (jit={jit}) matches the oracle on the union", flush=True)
    ctx.set_option("scan.jit", 2)

    # a filter that keeps nothing on any rank -> Ok(vec![]) ; and one that keeps rows on one rank only
    from query_engine_b200 import AggregateExpr, AggregateFunction as F, BinaryOp as B, PhysicalPlan, binary, col, lit

    mine, union = shard(datagen.lineitem, workloads.Q6_COLS)
    src = MemoryDataSource(mine[0].schema, mine)
    none = PhysicalPlan.HashAggregate(PhysicalPlan.Filter(PhysicalPlan.Scan(src), binary(col("l_quantity", 1), B.Greater, lit(1e9))), [],
                                      [AggregateExpr(F.Count, col("l_orderkey", 0)), AggregateExpr(F.Avg, col("l_quantity", 1))])
    assert dex.execute(none) == []
    kmax = max(int(pa.Table.from_batches(union).column("l_orderkey").to_numpy().max()), 1)
    one = PhysicalPlan.HashAggregate(
        PhysicalPlan.Filter(PhysicalPlan.Scan(src), binary(col("l_orderkey", 0), B.Equal, lit(kmax))), [col("l_shipdate", 4)],
        [AggregateExpr(F.Count, col("l_orderkey", 0)), AggregateExpr(F.Avg, col("l_quantity", 1)), AggregateExpr(F.Min, col("l_discount", 3))])
    usrc = MemoryDataSource(union[0].schema, union)
    want = oracle.execute(PhysicalPlan.HashAggregate(
        PhysicalPlan.Filter(PhysicalPlan.Scan(usrc), one.input.predicate), one.group_exprs, one.aggr_exprs))
    assert_batches_equal(dex.execute(one), want, ordered=False, approx_cols=[2])

    # many groups (l_shipdate: ~2 500 distinct values): the shared-memory tier overflows, the fixed-size-block merge of
    # small group sets must send every rank down the general all-gather + re-group path; and 40 groups (quantity < 41
    # of 50 values): more than a CTA's directory, fewer than a block holds on most ranks
    for gcol, gidx, pr in (("l_shipdate", 4, None), ("l_quantity", 1, binary(col("l_quantity", 1), B.Less, lit(41.0)))):
        def mk(s):
            inp = PhysicalPlan.Scan(s) if pr is None else PhysicalPlan.Filter(PhysicalPlan.Scan(s), pr)
            return PhysicalPlan.HashAggregate(inp, [col(gcol, gidx)], [AggregateExpr(F.Count, col("l_orderkey", 0)),
                                                                       AggregateExpr(F.Sum, col("l_extendedprice", 2)),
                                                                       AggregateExpr(F.Max, col("l_discount", 3))])
        got = dex.execute(mk(src))
        assert_batches_equal(got, oracle.execute(mk(usrc)), ordered=False, approx_cols=[2])
    if rank == 0:
        print("rank0: distributed GROUP BY with 2 500 and with 40 groups matches the oracle on the union", flush=True)

    # join: shuffle both sides, local join, concatenate over ranks
    lmine, lunion = shard(datagen.lineitem, workloads.JOIN_L_COLS)
    omine, ounion = shard(datagen.orders, workloads.JOIN_O_COLS)
    got = dex.execute(workloads.join(MemoryDataSource(lmine[0].schema, lmine), MemoryDataSource(omine[0].schema, omine)))
    parts = [None] * world
    dist.all_gather_object(parts, [b.to_pydict() for b in got])
    if rank == 0:
        allb = [pa.RecordBatch.from_pydict(d, schema=got[0].schema if got else None) for p in parts for d in p]
        # vectorised expectation (the oracle's join is a Python loop): every lineitem row pairs with its order
        lt = pa.Table.from_batches(lunion)
        ot = pa.Table.from_batches(ounion)
        j = lt.join(ot, keys="l_orderkey", right_keys="o_orderkey", join_type="inner", coalesce_keys=False)
        j = j.select(["l_orderkey", "l_extendedprice", "o_orderkey", "o_totalprice"])
        assert_batches_equal(allb, j.to_batches(), ordered=False)
        print("rank0: distributed join matches the union join", flush=True)
    # the same join through the NVLink-scatter path (the radix join's scatter writes into the owning GPU's exchange
    # arena): dense and hashed table layouts, INNER and LEFT with a tenth of the orders missing
    from query_engine_b200 import JoinType

    def drop_tenth(batches):
        out = []
        for b in batches:
            keep = pa.array((b.column(0).to_numpy() // 8) % 10 != 0)
            out.append(b.filter(keep))
        return [b for b in out if b.num_rows]

    omine_d, ounion_d = drop_tenth(omine), drop_tenth(ounion)
    ctx.set_option("join.radix_min_build_rows", 1)
    ctx.set_option("exchange.arena_round", 1 << 16)  # fine-grained arena sizes: the calls below make the arena GROW
    for dense in (1, 0):
        ctx.set_option("join.dense", dense)
        for jt, how in ((JoinType.Inner, "inner"), (JoinType.Left, "left outer")):
            plan = workloads.join(MemoryDataSource(lmine[0].schema, lmine), MemoryDataSource(omine_d[0].schema, omine_d), jt)
            done, _ = ctx.execute_join_exchange(dex._input(plan.left), dex._input(plan.right), int(jt), plan.on)
            assert done, "the NVLink-scatter join must accept this shape"
            got = dex.execute(plan)
            parts = [None] * world
            dist.all_gather_object(parts, [b.to_pydict() for b in got])
            if rank == 0:
                allb = [pa.RecordBatch.from_pydict(d, schema=got[0].schema) for p in parts for d in p]
                j = pa.Table.from_batches(lunion).join(pa.Table.from_batches(ounion_d), keys="l_orderkey", right_keys="o_orderkey",
                                                       join_type=how, coalesce_keys=False)
                j = j.select(["l_orderkey", "l_extendedprice", "o_orderkey", "o_totalprice"])
                assert_batches_equal(allb, j.to_batches(), ordered=False)
                print(f"rank0: NVLink-scatter join (dense={dense}, {how}) matches the union join", flush=True)
    ctx.set_option("join.dense", 1)
    ctx.set_option("join.radix_min_build_rows", 1 << 19)

    # LEFT join through the NCCL-shuffle path ("join.exchange" = 0) with a build side of two rows in all: most ranks
    # receive probe rows but no build row and must still emit them with NULL right columns
    ctx.set_option("join.exchange", 0)
    try:
        tiny = [b.slice(0, 2) for b in ounion[:1]]
        tiny_mine = tiny if rank == 0 else []
        plan = workloads.join(MemoryDataSource(lmine[0].schema, lmine), MemoryDataSource(ounion[0].schema, tiny_mine), JoinType.Left)
        got = dex.execute(plan)
        parts = [None] * world
        dist.all_gather_object(parts, [b.to_pydict() for b in got])
        if rank == 0:
            allb = [pa.RecordBatch.from_pydict(d, schema=got[0].schema) for p in parts for d in p]
            j = pa.Table.from_batches(lunion).join(pa.Table.from_batches(tiny), keys="l_orderkey", right_keys="o_orderkey",
                                                   join_type="left outer", coalesce_keys=False)
            j = j.select(["l_orderkey", "l_extendedprice", "o_orderkey", "o_totalprice"])
            assert_batches_equal(allb, j.to_batches(), ordered=False)
            print("rank0: LEFT join over the NCCL shuffle with a 2-row build side matches the union join", flush=True)
    finally:
        ctx.set_option("join.exchange", 1)

    # high-cardinality GROUP BY: the rows are exchanged by key over NVLink and every rank returns ITS shard of the
    # groups (the concatenation over ranks is the result); rank 1's shard is empty in the second round
    import pandas as pd

    for empty_rank in (None, 1):
        n = 0 if rank == empty_rank else 60_000
        t = datagen.highcard(60_000 * world, 25_000, rank, world)
        t = {k: v[:n] for k, v in t.items()}
        mine_b = datagen.to_batch(t)
        plan = workloads.highcard(MemoryDataSource(mine_b.schema, [mine_b]))
        ctx.set_option("group.partition_min_rows", 1)
        try:
            got = dex.execute(plan)
        finally:
            ctx.set_option("group.partition_min_rows", 1 << 22)
        parts = [None] * world
        dist.all_gather_object(parts, ([b.to_pydict() for b in got], {k: v.tolist() for k, v in t.items()}))
        if rank == 0:
            allb = [pa.RecordBatch.from_pydict(d) for p, _ in parts for d in p]
            dtypes = {"k": np.int64, "v": np.float64, "w": np.int64}  # (an empty shard must not turn the keys into floats)
            df = pd.concat([pd.DataFrame({k: np.asarray(v, dtype=dtypes[k]) for k, v in tt.items()}) for _, tt in parts])
            gb = df.groupby("k", sort=True)
            want = pa.table({"k": gb.v.min().index.values.astype(np.int64), "col_0": gb.v.min().values, "col_1": gb.v.max().values,
                             "col_2": gb.v.sum().values, "col_3": gb.w.sum().values.astype(np.int64),
                             "col_4": gb.k.count().values.astype(np.int64)})
            keys = pa.Table.from_batches(allb).column(0).to_pylist()
            assert len(keys) == len(set(keys)), "a key was returned by two ranks"
            assert_batches_equal(allb, want.to_batches(), ordered=False, approx_cols=[3], check_names=False)
            print(f"rank0: key-exchanged GROUP BY (empty rank: {empty_rank}) matches pandas on the union", flush=True)
    # Limit over the sharded result of a Filter: the global [skip, skip + fetch) window of the rank-ordered concatenation
    mine, union = shard(datagen.lineitem, workloads.Q6_COLS)
    pred = binary(col("l_quantity", 1), B.Less, lit(10.0))
    for skip, fetch in ((0, 7), (5, 4000), (3000, None), (10**9, 5)):
        got = dex.execute(PhysicalPlan.Limit(PhysicalPlan.Filter(PhysicalPlan.Scan(MemoryDataSource(mine[0].schema, mine)), pred), skip, fetch))
        parts = [None] * world
        dist.all_gather_object(parts, [b.to_pydict() for b in got])
        if rank == 0:
            allb = [pa.RecordBatch.from_pydict(d, schema=mine[0].schema) for p in parts for d in p]
            want = oracle.execute(PhysicalPlan.Limit(PhysicalPlan.Filter(PhysicalPlan.Scan(MemoryDataSource(union[0].schema, union)), pred), skip, fetch))
            assert_batches_equal(allb, want, ordered=True)
    if rank == 0:
        print("rank0: distributed Limit matches the oracle on the union", flush=True)
    dist.barrier()
    if rank == 0:
        print("DISTRIBUTED PARITY OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
