"""dev probe: Q1 kernel time under different scan options (rows per thread, CTAs per SM)"""
import json, os, sys
sys.path.insert(0, ".")
from query_engine_b200 import _ffi, datagen, workloads
from query_engine_b200.executor import GpuDataSource, QueryExecutor
ctx = _ffi.Context(0); ex = QueryExecutor(ctx=ctx)
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10
li = datagen.lineitem(sf)
which = os.environ.get("PROBE_WORKLOAD", "q1")
cols = workloads.Q1_COLS if which == "q1" else workloads.Q6_COLS
src = GpuDataSource(ctx, [datagen.to_batch({n: li[n] for n in cols})])
plan = (workloads.q1 if which == "q1" else workloads.q6)(src, cols)
OPTS = json.loads(os.environ.get("PROBE_OPTS", '[{}]'))
for opts in OPTS:
    for k, v in opts.items(): ctx.set_option(k, v)
    for _ in range(3): r = ex.execute_device(plan); del r
    ctx.profile_reset(); ctx.profile_enable(True); ctx.synchronize(); ctx.timer_start()
    for _ in range(10): r = ex.execute_device(plan); del r
    ms = ctx.timer_stop(); ctx.profile_enable(False)
    agg = {}
    for nm, t in ctx.profile_entries(): agg.setdefault(nm, []).append(t)
    print(opts, "ms/step %.3f" % (ms / 10), {k: round(sum(v) / 10, 3) for k, v in agg.items() if sum(v) / 10 > 0.02})
# where the host spends a step ("debug.host_timing"), and the wall time of a step with the profiler's events off
import time
ctx.set_option("debug.host_timing", 1)
r = ex.execute_device(plan); del r
ctx.set_option("debug.host_timing", 0)
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(50): r = ex.execute_device(plan); del r
ctx.synchronize()
print("wall ms/step over 50 steps, no events: %.3f" % ((time.perf_counter() - t0) * 1e3 / 50))
