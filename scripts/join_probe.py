"""dev probe: join step time vs kernel sum under different options"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from query_engine_b200 import _ffi, datagen, workloads
from query_engine_b200.executor import GpuDataSource, QueryExecutor
ctx = _ffi.Context(0); ex = QueryExecutor(ctx=ctx)
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10
li = datagen.lineitem(sf); od = datagen.orders(sf)
L = GpuDataSource(ctx, [datagen.to_batch({n: li[n] for n in workloads.JOIN_L_COLS})])
O = GpuDataSource(ctx, [datagen.to_batch({n: od[n] for n in workloads.JOIN_O_COLS})])
plan = workloads.join(L, O)
once = len(sys.argv) > 2
if once:
    r = ex.execute_device(plan); del r
    ctx.profile_reset(); ctx.profile_enable(True)
    r = ex.execute_device(plan); del r
    ctx.synchronize()
    print({k: round(v, 3) for k, v in ctx.profile_entries()})
    sys.exit(0)
import json, os
OPTS = json.loads(os.environ.get("JOIN_PROBE_OPTS", '[{"join.radix": 0}, {"join.radix": 1}]'))
for opts in OPTS:
    for k, v in opts.items(): ctx.set_option(k, v)
    for _ in range(3): r = ex.execute_device(plan); del r
    ctx.profile_reset(); ctx.profile_enable(True); ctx.synchronize()
    t0 = time.perf_counter(); ctx.timer_start()
    for _ in range(5): r = ex.execute_device(plan); del r
    ms = ctx.timer_stop(); wall = (time.perf_counter() - t0) * 1e3
    ctx.profile_enable(False)
    ctx.set_option("debug.host_timing", 1)
    ctx.set_option("debug.join_timing", 1)
    t1 = time.perf_counter(); r = ex.execute_device(plan); t2 = time.perf_counter(); del r; t3 = time.perf_counter()
    print("   one step host: call %.3f ms, release %.3f ms" % ((t2 - t1) * 1e3, (t3 - t2) * 1e3))
    ctx.set_option("debug.host_timing", 0)
    ctx.set_option("debug.join_timing", 0)
    agg = {}
    for nm, t in ctx.profile_entries(): agg.setdefault(nm, []).append(t)
    print(opts, "event ms/step %.3f wall %.3f kernel-sum %.3f" % (ms / 5, wall / 5, sum(sum(v) for v in agg.values()) / 5))
    print("   ", {k: round(sum(v) / 5, 3) for k, v in agg.items()})
