#!/usr/bin/env python
"""bench.py — rows/s and HBM-roofline fraction of the relational hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload q6|q1|join|highcard]
                    [--sf 10] [--also q1,join]

Headline workload (config.workload) is BASELINE.json configs[1]: TPC-H Q6-shape filter + SUM over synthetic
lineitem SF10 (60 M rows, int64 / f64 columns).  A "step" is one execution of the query's physical plan
(`QueryExecutor.execute_device`) over the HBM-resident table; `e2e` is the same plan through `QueryExecutor.execute`
with HOST Arrow batches (pinned), i.e. H2D of the inputs and D2H of the result inside the timed region.
Under torchrun (N > 1) every rank owns an SF-sized shard (weak scaling); partial aggregates are all-gathered over NCCL
and merged on the GPU (PartialAggregate(n) -> FinalAggregate, query-distributed/src/planner.rs:200-226); the join
shuffles both inputs by key hash over NCCL first.

`--impl reference` times the CPU port of the reference's executor (oracle/cpu_ref.cpp; the Rust reference cannot be
built in this image) on the box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "rows_per_sec"
UNIT = "rows/s"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# workloads: data (host numpy), plan builder, algorithmic bytes, independent numpy check
# ------------------------------------------------------------------------------------------------
def make_workload(name, sf, shard, n_shards):
    from query_engine_b200 import datagen, workloads

    w = {"name": name}
    if name in ("q6", "q1"):
        li = datagen.lineitem(sf, shard, n_shards)
        names = workloads.Q6_COLS if name == "q6" else workloads.Q1_COLS
        w["tables"] = {"lineitem": {n: li[n] for n in names}}
        w["rows"] = len(li["l_orderkey"])
        w["bytes_per_row"] = workloads.Q6_BYTES_PER_ROW if name == "q6" else workloads.Q1_BYTES_PER_ROW
        w["plan"] = (lambda src: workloads.q6(src["lineitem"], names)) if name == "q6" else (lambda src: workloads.q1(src["lineitem"], names))
        w["desc"] = (
            f"TPC-H Q6-shape filter + SUM over synthetic lineitem SF{sf:g}" if name == "q6"
            else f"TPC-H Q1-shape filter + 4-group hash aggregate over synthetic lineitem SF{sf:g}"
        )
    elif name == "join":
        li = datagen.lineitem(sf, shard, n_shards)
        od = datagen.orders(sf, shard, n_shards)
        w["tables"] = {"lineitem": {n: li[n] for n in workloads.JOIN_L_COLS}, "orders": {n: od[n] for n in workloads.JOIN_O_COLS}}
        w["rows"] = len(li["l_orderkey"])  # probe rows
        n_l, n_r = len(li["l_orderkey"]), len(od["o_orderkey"])
        w["bytes_per_row"] = (16 * (n_l + n_r) + 32 * n_l) / n_l
        w["plan"] = lambda src: workloads.join(src["lineitem"], src["orders"])
        w["desc"] = f"synthetic lineitem JOIN orders INNER hash join on orderkey SF{sf:g}"
        w["operator_kernels"] = "k_pj_"  # the radix join is ONE operator of three launches
    elif name == "highcard":
        n_rows, n_keys = int(100_000_000 * sf / 10), int(10_000_000 * sf / 10)
        t = datagen.highcard(n_rows, n_keys, shard, n_shards)
        w["tables"] = {"t": t}
        w["rows"] = len(t["k"])
        w["bytes_per_row"] = workloads.HIGHCARD_BYTES_PER_ROW
        w["plan"] = lambda src: workloads.highcard(src["t"])
        w["desc"] = f"high-cardinality GROUP BY ({n_keys} distinct int64 keys, {n_rows} rows) MIN/MAX/SUM"
    else:
        raise SystemExit(f"unknown workload {name}")
    return w


def numpy_check(w, result_batches):
    """Independent closed-form check of the GPU result (plain numpy; not the oracle)."""
    import pyarrow as pa

    name = w["name"]
    t = pa.Table.from_batches(result_batches)
    if name == "q6":
        li = w["tables"]["lineitem"]
        from query_engine_b200.datagen import Q6_DATE_HI, Q6_DATE_LO

        m = ((li["l_shipdate"] >= Q6_DATE_LO) & (li["l_shipdate"] < Q6_DATE_HI) & (li["l_discount"] >= 0.05)
             & (li["l_discount"] <= 0.07) & (li["l_quantity"] < 24.0))
        s = float(np.sum(li["l_extendedprice"][m] * li["l_discount"][m]))
        got_s, got_c = t.column(0)[0].as_py(), t.column(1)[0].as_py()
        return got_c == int(m.sum()) and abs(got_s - s) <= 1e-9 * abs(s)
    if name == "q1":
        li = w["tables"]["lineitem"]
        from query_engine_b200.datagen import Q1_DATE_MAX

        m = li["l_shipdate"] <= Q1_DATE_MAX
        ok = t.num_rows == 4
        for r in range(t.num_rows):
            rf, ls = t.column(0)[r].as_py(), t.column(1)[r].as_py()
            g = m & (li["l_returnflag"] == rf) & (li["l_linestatus"] == ls)
            cnt = int(g.sum())
            sq = float(li["l_quantity"][g].sum())
            dp = float((li["l_extendedprice"][g] * (1.0 - li["l_discount"][g])).sum())
            ok &= t.column(9)[r].as_py() == cnt
            ok &= abs(t.column(2)[r].as_py() - sq) <= 1e-9 * abs(sq)
            ok &= abs(t.column(4)[r].as_py() - dp) <= 1e-9 * abs(dp)
        return bool(ok)
    if name == "join":
        li, od = w["tables"]["lineitem"], w["tables"]["orders"]
        if t.num_rows != len(li["l_orderkey"]):
            return False
        # checksum of checksums: every output row must pair equal keys, and the multiset of left rows is the input
        lk = t.column(0).to_numpy()
        ok = bool(np.array_equal(lk, t.column(2).to_numpy()))
        ok &= int(lk.sum()) == int(li["l_orderkey"].sum())
        order = np.argsort(od["o_orderkey"], kind="stable")
        pos = np.searchsorted(od["o_orderkey"][order], lk)
        ok &= bool(np.array_equal(od["o_totalprice"][order][pos], t.column(3).to_numpy()))
        return bool(ok)
    if name == "highcard":
        tb = w["tables"]["t"]
        uniq, inv = np.unique(tb["k"], return_inverse=True)
        if t.num_rows != len(uniq):
            return False
        o = np.argsort(t.column(0).to_numpy())
        cnt = np.bincount(inv, minlength=len(uniq))
        sw = np.zeros(len(uniq), dtype=np.int64)
        np.add.at(sw, inv, tb["w"])
        return bool(np.array_equal(t.column(5).to_numpy()[o], cnt) and np.array_equal(t.column(4).to_numpy()[o], sw))
    return True


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x0000000000000008: "hw_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000004: "sw_power_cap",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)  # the recipe's 200 ms; NVML queries stall kernel launches for milliseconds, keep them rare

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU port (oracle/cpu_ref.cpp) — baseline only
# ------------------------------------------------------------------------------------------------
def cpu_port_time(w, sample_rows, threads, batch_rows=8192, repeats=1):
    """Times the C++ port of the reference's executor on the first `sample_rows` rows; returns rows/s."""
    import subprocess

    from query_engine_b200 import _ffi
    from query_engine_b200.physical_plan import PhysicalPlan

    so = ROOT / "oracle" / "libcpuref.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.cpuref_last_error.restype = C.c_char_p
    name = w["name"]
    if name == "join":
        li, od = w["tables"]["lineitem"], w["tables"]["orders"]
        n = min(sample_rows, len(li["l_orderkey"]))
        lk = np.ascontiguousarray(li["l_orderkey"][:n])
        rk = np.ascontiguousarray(od["o_orderkey"])
        ol, orr = np.empty(n + 16, dtype=np.int64), np.empty(n + 16, dtype=np.int64)
        lib.cpuref_hash_join.restype = C.c_int64
        lib.cpuref_hash_join.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
        lib.cpuref_take8.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        out = np.empty(n + 16, dtype=np.uint64)
        t0 = time.perf_counter()
        cnt = lib.cpuref_hash_join(lk.ctypes.data, n, rk.ctypes.data, len(rk), 0, n + 16, ol.ctypes.data, orr.ctypes.data)
        for src, idx in ((li["l_orderkey"], ol), (li["l_extendedprice"], ol), (od["o_orderkey"], orr), (od["o_totalprice"], orr)):
            lib.cpuref_take8(np.ascontiguousarray(src).ctypes.data, idx.ctypes.data, cnt, out.ctypes.data)
        dt = time.perf_counter() - t0
        return n / dt, n
    tab = next(iter(w["tables"].values()))
    names = list(tab)
    n = min(sample_rows, len(tab[names[0]]))
    cols = [np.ascontiguousarray(tab[c][:n]) for c in names]
    ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    types = (C.c_int32 * len(cols))(*[3 if c.dtype == np.int64 else 5 for c in cols])
    plan = w["plan"]({k: None for k in w["tables"]})
    assert isinstance(plan, PhysicalPlan.HashAggregate)
    pred = plan.input.predicate if isinstance(plan.input, PhysicalPlan.Filter) else None
    fp = _ffi.FlatExpr(pred) if pred is not None else None
    g = _ffi.FlatExprArray(plan.group_exprs)
    a = _ffi.FlatAggArray(plan.aggr_exprs)
    cap = 1 << 24 if name == "highcard" else 1 << 12
    keys = np.empty(cap * max(1, g.n), dtype=np.uint64)
    vals = np.empty(cap * a.n, dtype=np.uint64)
    valid = np.empty(cap * a.n, dtype=np.uint8)
    ng = C.c_int64()
    lib.cpuref_filter_aggregate.restype = C.c_int32
    lib.cpuref_filter_aggregate.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rc = lib.cpuref_filter_aggregate(C.addressof(ptrs), C.addressof(types), len(cols), n, batch_rows,
                                         C.addressof(fp.c) if fp else None, C.addressof(g.array), g.n, C.addressof(a.array), a.n,
                                         threads, cap, keys.ctypes.data, vals.ctypes.data, valid.ctypes.data, C.addressof(ng))
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError(lib.cpuref_last_error().decode())
        best = dt if best is None else min(best, dt)
    return n / best, n


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, args.sf_ref, 0, 1)
    cores = os.cpu_count() or 1
    sample = min(w["rows"], args.ref_rows)
    for _ in range(args.warmup):
        cpu_port_time(w, min(sample, 500_000), 1)
    t0 = time.perf_counter()
    vals = [cpu_port_time(w, sample, 1)[0] for _ in range(args.steps)]
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v1 = float(np.median(vals))
    vN = cpu_port_time(w, sample, cores)[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": v1, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64/int64", "data": "synthetic",
        "config": {"workload": w["desc"], "sf": args.sf_ref, "sample_rows": sample, "batch_rows": 8192},
        "cpu_baseline": {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"first {sample} rows of the workload in 8192-row batches, single thread like the reference's executor",
                         "all_cores": {"value": vN, "cores": cores, "note": "same port with its per-batch work spread over all host threads (the reference has no such path)"}},
        "e2e": {"value": v1, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json, written by scripts/summarize_profiles.py), or None"""
    try:
        t = json.load(open(ROOT / "profiles" / "traffic.json"))
    except Exception:
        return None
    if kernel is None:
        return None
    parts = [t.get(k.strip()) for k in kernel.split("+")]
    if any(p is None for p in parts):
        return None
    return sum(p["dram_bytes_per_launch"] for p in parts)


def run_ours(args):
    # NCCL / torch may print banners on stdout; the contract is ONE JSON line there, so park fd 1 on stderr meanwhile
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def _run_ours(args):
    import pyarrow as pa

    from query_engine_b200 import MemoryDataSource, _ffi, datagen
    from query_engine_b200.executor import GpuDataSource, QueryExecutor

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _ffi.Context(local_rank)
    ex = QueryExecutor(ctx=ctx)
    dex = None
    if world > 1:
        import torch

        from query_engine_b200.distributed import DistributedExecutor

        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
        dex = DistributedExecutor(ex)

    def barrier():
        ctx.synchronize()
        if dist is not None:
            import torch

            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = load_peaks()
    results = {}
    names = [args.workload] + [x for x in args.also.split(",") if x and x != args.workload]
    for wi, name in enumerate(names):
        head = wi == 0
        # weak scaling: every rank owns a full SF-sized shard (the job is SF x world)
        w = make_workload(name, args.sf * world, rank, world)
        steps = args.steps if head else max(3, args.steps // 4)
        warm = args.warmup
        host_batches = {t: [datagen.to_batch(cols)] for t, cols in w["tables"].items()}
        h2d_bytes = sum(a.nbytes for cols in w["tables"].values() for a in cols.values())
        # pin the host columns (north_star: Arrow buffers consumed from pinned host memory)
        for cols in w["tables"].values():
            for a in cols.values():
                ctx.host_register(a.ctypes.data, a.nbytes)
        resident = {t: GpuDataSource(ctx, b) for t, b in host_batches.items()}
        plan = w["plan"](resident)
        run = (lambda: dex.execute_device(plan)) if dex is not None else (lambda: ex.execute_device(plan))

        out = run()
        res_batches = out.to_batches() if out is not None else []
        checked = numpy_check(w, res_batches) if (world == 1 and not args.no_check) else None
        del out
        for _ in range(warm):
            run()
        sampler = ClockSampler(local_rank)
        ctx.profile_reset()
        ctx.profile_enable(True)
        barrier()
        sampler.start()
        ctx.timer_start()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = run()
            del r
        ms = ctx.timer_stop()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        clocks = sampler.stop()
        ctx.profile_enable(False)
        launches = ctx.profile_count()
        entries = ctx.profile_entries()
        ms = max_over_ranks(ms)
        total_rows = sum_over_ranks(float(w["rows"]))
        value = total_rows * steps / (ms / 1e3)

        # dominant kernel (by summed device time over the timed region)
        agg = {}
        for nm, t in entries:
            a = agg.setdefault(nm, [0.0, 0])
            a[0] += t
            a[1] += 1
        kern_total = sum(v[0] for v in agg.values())
        top = max(agg.items(), key=lambda kv: kv[1][0]) if agg else (None, [0.0, 1])
        top_ms = top[1][0] / max(1, top[1][1])
        op_prefix = w.get("operator_kernels")
        if op_prefix and agg:
            # one operator = several launches that only make sense together (the radix join: two scatter passes + the
            # build/probe kernel): its algorithmic bytes are set against the SUM of those launches, not the longest one
            members = {k: v for k, v in agg.items() if k.startswith(op_prefix)}
            if members:
                top = (" + ".join(sorted(members)), [sum(v[0] for v in members.values()), steps])
                top_ms = top[1][0] / max(1, steps)
        # algorithmic bytes of one step on this rank / the summed kernel time of one step
        alg_bytes = w["bytes_per_row"] * w["rows"]
        step_kernel_ms = kern_total / steps if steps else 0.0
        roofline = {
            "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
            "kernel": top[0], "kernel_ms": top_ms, "kernel_share_of_step": (top[1][0] / kern_total) if kern_total else None,
            "algorithmic_bytes_per_launch": alg_bytes,
            "achieved": alg_bytes / (top_ms / 1e3) / 1e9 if top_ms else None,
            "traffic": measured_traffic(top[0]),
        }
        roofline["frac"] = roofline["achieved"] / peak if roofline["achieved"] else None
        roofline["step_achieved"] = alg_bytes / (ms / steps / 1e3) / 1e9
        roofline["step_frac"] = roofline["step_achieved"] / peak
        roofline["kernels"] = {k: {"ms_avg": v[0] / v[1], "launches_per_step": v[1] / steps} for k, v in sorted(agg.items())}

        # end to end: host Arrow batches -> H2D -> plan -> D2H of the result, through the public API
        e2e_steps = max(2, min(steps, args.e2e_steps))
        host_src = {t: MemoryDataSource(b[0].schema, b) for t, b in host_batches.items()}
        host_plan = w["plan"](host_src)
        run_host = (lambda: dex.execute(host_plan)) if dex is not None else (lambda: ex.execute(host_plan))
        # warm-up the way the loop runs (the previous result is still alive while the next one is produced), so both
        # sets of pinned result buffers exist before the clock starts (pinning 2 GB takes ~0.5 s, once)
        rb = run_host()
        rb = run_host()
        barrier()
        uploaded0 = ex.h2d_bytes
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(e2e_steps):
            rb = run_host()
            d2h = sum(b.nbytes for b in rb)
        ctx.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        barrier()
        # bytes actually handed to the library for upload (an aggregate over a host scan uploads only the columns it
        # reads); the multi-GPU executor imports whole shards and does not count, so fall back to the table size
        uploaded = (ex.h2d_bytes - uploaded0) // e2e_steps
        if uploaded > 0:
            h2d_bytes = int(uploaded)
        e2e = {"value": total_rows * e2e_steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps}

        results[name] = {
            "desc": w["desc"], "value": value, "ms_per_step": ms / steps, "wall_ms_per_step": wall_ms / steps, "steps": steps,
            "rows_per_step": total_rows, "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "checked_against_numpy": checked,
        }
        if head and rank == 0 and world == 1 and not args.no_cpu:
            v, n = cpu_port_time(w, args.cpu_rows, 1)
            results[name]["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"first {n} rows of the same workload, 8192-row batches, single thread (the reference's executor is single-threaded)",
            }
        for cols in w["tables"].values():
            for a in cols.values():
                ctx.host_unregister(a.ctypes.data)
        del resident, plan, host_plan, host_src, host_batches, w

    if rank == 0:
        h = results[names[0]]
        line = {
            "metric": METRIC, "value": h["value"], "unit": UNIT, "n_gpus": world, "steps": h["steps"], "warmup": args.warmup,
            "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64/int64", "data": "synthetic",
            "config": {"workload": h["desc"], "sf_per_gpu": args.sf, "rows_per_gpu": int(h["rows_per_step"] / world),
                       "parallelism": f"row-shard x{world}", "l2": "inputs (GBs) exceed the 126 MB L2; no flush needed"},
            "roofline": {k: h["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "kernel_ms",
                                                         "kernel_share_of_step", "peak_source", "step_achieved", "step_frac", "kernels",
                                                         "algorithmic_bytes_per_launch")},
            "e2e": h["e2e"], "gpu_launches": h["gpu_launches"], "clocks": h["clocks"],
            "checked_against_numpy": h["checked_against_numpy"],
        }
        if "cpu_baseline" in h:
            line["cpu_baseline"] = h["cpu_baseline"]
        if len(names) > 1:
            line["also"] = {n: {k: results[n][k] for k in ("desc", "value", "ms_per_step", "steps", "roofline", "e2e", "checked_against_numpy")}
                            for n in names[1:]}
    else:
        line = None
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="q6")
    ap.add_argument("--also", default="q1,join", help="further workloads measured after the headline one (comma list)")
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--sf-ref", type=float, default=1.0, help="scale factor the reference arm generates (it times a sample)")
    ap.add_argument("--ref-rows", type=int, default=2_000_000)
    ap.add_argument("--cpu-rows", type=int, default=4_000_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
