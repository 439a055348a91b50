#!/usr/bin/env python
"""bench.py — rows/s and HBM-roofline fraction of the relational hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload q6|q1|join|highcard|filter_q6|filter_q1] [--sf 10] [--also q1,join,highcard,filter_q6,filter_q1]

Headline workload (config.workload) is BASELINE.json configs[1]: TPC-H Q6-shape filter + SUM over synthetic
lineitem SF10 (60 M rows, int64 / f64 columns).  A "step" is one execution of the query's physical plan
(`QueryExecutor.execute_device`) over the HBM-resident table; `e2e` is the same plan through `QueryExecutor.execute`
with HOST Arrow batches (pinned), i.e. H2D of the inputs and D2H of the result inside the timed region.
Under torchrun (N > 1) every rank owns an SF-sized shard (weak scaling); partial aggregates are all-gathered over NCCL
and merged on the GPU (PartialAggregate(n) -> FinalAggregate, query-distributed/src/planner.rs:200-226); the join and
the high-cardinality GROUP BY exchange rows by key over NVLink.  The join's shards are dealt round-robin across the
ranks before the run, so a rank's rows carry keys of the whole key range and (N-1)/N of them cross NVLink.  At N > 1
every result is verified inside the run through all-reduced invariants (row counts, key / payload checksums, per-group
counts and sums) computed independently with numpy on every rank's input.

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
def redeal(cols, dist, world):
    """Row i of this rank's shard moves to rank i % world (one all-to-all per column, before the timed region): every
    rank ends up with 1/world of every canonical shard, i.e. with keys from the whole key range."""
    import torch

    out = {}
    for name, a in cols.items():
        parts = [np.ascontiguousarray(a[r::world]) for r in range(world)]
        in_splits = [len(p) for p in parts]
        send = torch.from_numpy(np.concatenate(parts)).cuda()
        sizes = torch.tensor(in_splits, dtype=torch.int64, device="cuda")
        out_sizes = torch.empty_like(sizes)
        dist.all_to_all_single(out_sizes, sizes)
        out_splits = [int(x) for x in out_sizes.tolist()]
        recv = torch.empty(sum(out_splits), dtype=send.dtype, device="cuda")
        dist.all_to_all_single(recv, send, out_splits, in_splits)
        out[name] = recv.cpu().numpy()
        del send, recv
    torch.cuda.empty_cache()
    return out


def make_workload(name, sf, shard, n_shards, dist=None):
    from query_engine_b200 import datagen, workloads

    w = {"name": name, "expect": {}}
    if name in ("q6", "q1", "filter_q6", "filter_q1"):
        li = datagen.lineitem(sf, shard, n_shards)
        q6_like = name in ("q6", "filter_q6")
        names = workloads.Q6_COLS if q6_like else workloads.Q1_COLS
        w["tables"] = {"lineitem": {n: li[n] for n in names}}
        w["rows"] = len(li["l_orderkey"])
        if name == "q6":
            w["bytes_per_row"] = workloads.Q6_BYTES_PER_ROW
            w["plan"] = lambda src: workloads.q6(src["lineitem"], names)
            w["desc"] = f"TPC-H Q6-shape filter + SUM over synthetic lineitem SF{sf:g}"
        elif name == "q1":
            w["bytes_per_row"] = workloads.Q1_BYTES_PER_ROW
            w["plan"] = lambda src: workloads.q1(src["lineitem"], names)
            w["desc"] = f"TPC-H Q1-shape filter + 4-group hash aggregate over synthetic lineitem SF{sf:g}"
        else:
            # the operator-faithful materialising Filter: 8 * c * (n + k) bytes for c columns, k survivors
            from query_engine_b200.datagen import Q1_DATE_MAX, Q6_DATE_HI, Q6_DATE_LO

            if q6_like:
                m = ((li["l_shipdate"] >= Q6_DATE_LO) & (li["l_shipdate"] < Q6_DATE_HI) & (li["l_discount"] >= 0.05)
                     & (li["l_discount"] <= 0.07) & (li["l_quantity"] < 24.0))
            else:
                m = li["l_shipdate"] <= Q1_DATE_MAX
            k = int(m.sum())
            w["expect"] = {"k": k, "key_sum": int(li["l_orderkey"][m].sum())}
            w["bytes_per_row"] = 8 * len(names) * (w["rows"] + k) / w["rows"]
            w["plan"] = lambda src: workloads.filter_only(src["lineitem"], names, "q6" if q6_like else "q1")
            # one operator = the predicate scan (selection bitmap), the prefix sums and the compaction of every column
            w["operator_kernels"] = ("k_scan_filter_mask", "k_compact", "k_scan_block_sums", "k_scan_sums", "k_scan_apply", "k_prefix_at")
            w["desc"] = (f"materialising Filter alone, {'Q6' if q6_like else 'Q1'} predicate ({100.0 * k / w['rows']:.1f} % survive), "
                         f"{len(names)} columns of synthetic lineitem SF{sf:g}")
    elif name == "join":
        li = datagen.lineitem(sf, shard, n_shards)
        od = datagen.orders(sf, shard, n_shards)
        lt = {n: li[n] for n in workloads.JOIN_L_COLS}
        ot = {n: od[n] for n in workloads.JOIN_O_COLS}
        # what the join must produce, from the canonical (still co-located) shards: every line pairs with its order
        order = np.argsort(ot["o_orderkey"], kind="stable")
        pos = np.searchsorted(ot["o_orderkey"][order], lt["l_orderkey"])
        w["expect"] = {"rows": len(lt["l_orderkey"]), "key_sum": int(lt["l_orderkey"].sum()),
                       "price_sum": float(ot["o_totalprice"][order][pos].sum()), "ext_sum": float(lt["l_extendedprice"].sum())}
        if dist is not None and n_shards > 1:
            lt, ot = redeal(lt, dist, n_shards), redeal(ot, dist, n_shards)
        w["tables"] = {"lineitem": lt, "orders": ot}
        w["rows"] = len(lt["l_orderkey"])  # probe rows
        n_l, n_r = len(lt["l_orderkey"]), len(ot["o_orderkey"])
        w["bytes_per_row"] = (16 * (n_l + n_r) + 32 * n_l) / n_l
        w["plan"] = lambda src: workloads.join(src["lineitem"], src["orders"])
        w["desc"] = f"synthetic lineitem JOIN orders INNER hash join on orderkey SF{sf:g}" + (
            ", shards dealt round-robin over the ranks (keys not co-located)" if n_shards > 1 else "")
        w["operator_kernels"] = ("k_pj_", "k_rscatter_", "k_region_")  # the radix join is ONE operator of several launches
        w["exchange_bytes"] = 16 * (n_l + n_r) * (n_shards - 1) / max(1, n_shards)  # leave this GPU in a hash exchange
    elif name == "highcard":
        # BASELINE config 5: 10 M distinct keys in all; 100 M rows on one GPU, 125 M rows per GPU otherwise (1 B on 8)
        n_keys = int(10_000_000 * sf / 10)
        per_gpu = int((100_000_000 if n_shards == 1 else 125_000_000) * sf / 10)
        t = datagen.highcard(per_gpu * n_shards, n_keys, shard, n_shards, with_index=True)
        j = t.pop("j")
        w["expect"]["j"] = j
        with np.errstate(over="ignore"):
            w["expect"] |= {"rows": len(t["k"]), "w_sum": int(t["w"].sum()), "v_sum": float(t["v"].sum()), "v_min": float(t["v"].min()),
                           "v_max": float(t["v"].max()), "key_sum": int(t["k"].astype(np.uint64).sum(dtype=np.uint64).view(np.int64)),
                           "seen": np.bincount(j, minlength=n_keys) > 0}
        w["tables"] = {"t": t}
        w["rows"] = len(t["k"])
        w["bytes_per_row"] = workloads.HIGHCARD_BYTES_PER_ROW
        w["plan"] = lambda src: workloads.highcard(src["t"])
        w["desc"] = f"high-cardinality GROUP BY ({n_keys} distinct int64 keys, {per_gpu * n_shards} rows) MIN/MAX/SUM"
        w["operator_kernels"] = ("k_rscatter_groups", "k_scan_filter_group_part", "k_pj_scatter_groups", "k_scan_filter_group_hbm",
                                 "k_group_", "k_agg_finalize", "k_gather", "k_merge_groups")
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
    if name in ("filter_q6", "filter_q1"):
        return bool(t.num_rows == w["expect"]["k"] and int(t.column(0).to_numpy().sum()) == w["expect"]["key_sum"])
    if name == "highcard":
        # per key index j (the key is splitmix64(j)): row count, SUM(w) (exact in f64: |sum| < 2^53), MIN / MAX of v
        from query_engine_b200.datagen import splitmix64

        tb, j = w["tables"]["t"], w["expect"]["j"]
        n_keys = len(w["expect"]["seen"])
        cnt = np.bincount(j, minlength=n_keys)
        sw = np.bincount(j, weights=tb["w"].astype(np.float64), minlength=n_keys)
        seen = np.flatnonzero(cnt > 0)
        if t.num_rows != len(seen):
            return False
        keys = splitmix64(seen.astype(np.uint64)).view(np.int64)
        oe, og = np.argsort(keys), np.argsort(t.column(0).to_numpy())
        ok = np.array_equal(t.column(0).to_numpy()[og], keys[oe])
        ok &= np.array_equal(t.column(5).to_numpy()[og], cnt[seen][oe])
        ok &= np.array_equal(t.column(4).to_numpy()[og], sw[seen][oe].astype(np.int64))
        sv = np.bincount(j, weights=tb["v"], minlength=n_keys)[seen][oe]
        ok &= bool(np.allclose(t.column(3).to_numpy()[og], sv, rtol=1e-9, atol=0.0))
        ok &= float(t.column(1).to_numpy().min()) == w["expect"]["v_min"] and float(t.column(2).to_numpy().max()) == w["expect"]["v_max"]
        return bool(ok)
    return True


def distributed_check(w, result_batches, dist):
    """N > 1: the result is verified through all-reduced invariants.  Aggregates (q6, q1) come back merged and identical
    on every rank: compared with all-reduced per-rank numpy partials.  The join and the high-cardinality GROUP BY come
    back sharded: row / group counts, key and payload checksums and extreme values are all-reduced over the ranks'
    shards of the result and compared with the same quantities all-reduced over the ranks' inputs."""
    import pyarrow as pa
    import torch

    def allsum(vals, dtype=torch.float64):
        x = torch.tensor(vals, dtype=dtype, device="cuda")
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
        return x.cpu().numpy()

    def close(a, b):
        return abs(a - b) <= 1e-9 * max(abs(a), abs(b))

    name = w["name"]
    t = pa.Table.from_batches(result_batches) if result_batches else None
    if name == "q6":
        li = w["tables"]["lineitem"]
        from query_engine_b200.datagen import Q6_DATE_HI, Q6_DATE_LO

        m = ((li["l_shipdate"] >= Q6_DATE_LO) & (li["l_shipdate"] < Q6_DATE_HI) & (li["l_discount"] >= 0.05)
             & (li["l_discount"] <= 0.07) & (li["l_quantity"] < 24.0))
        cnt = int(allsum([int(m.sum())], torch.int64)[0])
        ssum = float(allsum([float(np.sum(li["l_extendedprice"][m] * li["l_discount"][m]))])[0])
        return bool(t is not None and t.column(1)[0].as_py() == cnt and close(t.column(0)[0].as_py(), ssum))
    if name == "q1":
        li = w["tables"]["lineitem"]
        from query_engine_b200.datagen import Q1_DATE_MAX

        m = li["l_shipdate"] <= Q1_DATE_MAX
        groups = [(0, 0), (1, 0), (2, 0), (2, 1)]
        cnts, sums = [], []
        for rf, ls in groups:
            g = m & (li["l_returnflag"] == rf) & (li["l_linestatus"] == ls)
            cnts.append(int(g.sum()))
            sums += [float(li["l_quantity"][g].sum()), float((li["l_extendedprice"][g] * (1.0 - li["l_discount"][g])).sum())]
        cnts, sums = allsum(cnts, torch.int64), allsum(sums)
        ok = t is not None and t.num_rows == sum(1 for c in cnts if c > 0)
        for r in range(t.num_rows if ok else 0):
            gi = groups.index((t.column(0)[r].as_py(), t.column(1)[r].as_py()))
            ok &= t.column(9)[r].as_py() == int(cnts[gi])
            ok &= close(t.column(2)[r].as_py(), float(sums[2 * gi])) and close(t.column(4)[r].as_py(), float(sums[2 * gi + 1]))
        return bool(ok)
    if name == "join":
        e = w["expect"]
        if t is not None and t.num_rows:
            lk = t.column(0).to_numpy()
            local_ok = int(np.array_equal(lk, t.column(2).to_numpy()))
            got = [t.num_rows, int(lk.sum())], [float(t.column(3).to_numpy().sum()), float(t.column(1).to_numpy().sum())]
        else:
            local_ok, got = 1, ([0, 0], [0.0, 0.0])
        gi = allsum(got[0] + [local_ok], torch.int64)
        gf = allsum(got[1])
        wi = allsum([e["rows"], e["key_sum"]], torch.int64)
        wf = allsum([e["price_sum"], e["ext_sum"]])
        world = dist.get_world_size()
        return bool(gi[0] == wi[0] and gi[1] == wi[1] and gi[2] == world and close(float(gf[0]), float(wf[0])) and close(float(gf[1]), float(wf[1])))
    if name == "highcard":
        e = w["expect"]
        seen = torch.from_numpy(e["seen"].astype(np.int32)).cuda()
        dist.all_reduce(seen, op=dist.ReduceOp.SUM)
        distinct = int((seen > 0).sum().item())
        if t is not None and t.num_rows:
            k, cnt = t.column(0).to_numpy(), t.column(5).to_numpy()
            with np.errstate(over="ignore"):
                ks = int((k.astype(np.uint64) * cnt.astype(np.uint64)).sum(dtype=np.uint64).view(np.int64))
            gi = [t.num_rows, int(cnt.sum()), int(t.column(4).to_numpy().sum()), ks & 0x7fffffffffff]
            gf = [float(t.column(3).to_numpy().sum())]
            mn, mx = float(t.column(1).to_numpy().min()), float(t.column(2).to_numpy().max())
        else:
            gi, gf, mn, mx = [0, 0, 0, 0], [0.0], float("inf"), float("-inf")
        gi = allsum(gi, torch.int64)
        gf = allsum(gf)
        wi = allsum([e["rows"], e["w_sum"], e["key_sum"] & 0x7fffffffffff], torch.int64)
        wf = allsum([e["v_sum"]])
        ext = torch.tensor([-mn, mx, -e["v_min"], e["v_max"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(ext, op=dist.ReduceOp.MAX)
        ext = ext.cpu().numpy()
        mask = (1 << 47) - 1  # key checksums are compared modulo 2^47 (sums of 47-bit residues stay exact in int64)
        return bool(gi[0] == distinct and gi[1] == wi[0] and gi[2] == wi[1] and (int(gi[3]) & mask) == (int(wi[2]) & mask)
                    and close(float(gf[0]), float(wf[0])) and ext[0] == ext[2] and ext[1] == ext[3])
    if name in ("filter_q6", "filter_q1"):
        e = w["expect"]
        got = [t.num_rows, int(t.column(0).to_numpy().sum())] if t is not None else [0, 0]
        g = allsum(got, torch.int64)
        x = allsum([e["k"], e["key_sum"]], torch.int64)
        return bool(g[0] == x[0] and g[1] == x[1])
    return None


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
    if isinstance(plan, PhysicalPlan.Filter):  # the materialising Filter alone: no aggregates
        pred, group_exprs, aggr_exprs = plan.predicate, [], []
    else:
        assert isinstance(plan, PhysicalPlan.HashAggregate)
        pred = plan.input.predicate if isinstance(plan.input, PhysicalPlan.Filter) else None
        group_exprs, aggr_exprs = plan.group_exprs, plan.aggr_exprs
    fp = _ffi.FlatExpr(pred) if pred is not None else None
    g = _ffi.FlatExprArray(group_exprs)
    a = _ffi.FlatAggArray(aggr_exprs)
    cap = 1 << 24 if name == "highcard" else 1 << 12
    keys = np.empty(cap * max(1, g.n), dtype=np.uint64)
    vals = np.empty(cap * max(1, a.n), dtype=np.uint64)
    valid = np.empty(cap * max(1, a.n), dtype=np.uint8)
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
    """The reference's CPU path for the SAME config as our arm (BASELINE configs[1]: Q6 over the 60 M-row SF10 table):
    oracle/cpu_ref.cpp, the C++ port of the reference's executor structure (the Rust reference cannot be built here).
    `value` is the single-threaded figure - the reference's executor is single-threaded (sequential loops,
    executor.rs:138,170-176; no rayon) - and the same port spread over every host core is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, args.sf, 0, 1)
    cores = os.cpu_count() or 1
    # one step = the whole table unless --ref-rows bounds it (a step of Q6 at SF10 takes ~0.6 s on one core)
    sample = w["rows"] if args.ref_rows <= 0 else min(w["rows"], args.ref_rows)
    for _ in range(min(args.warmup, 2)):
        cpu_port_time(w, min(sample, 2_000_000), 1)
    t0 = time.perf_counter()
    vals = [cpu_port_time(w, sample, 1)[0] for _ in range(args.steps)]
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v1 = float(np.median(vals))
    vN = cpu_port_time(w, sample, cores)[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": v1, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64/int64", "data": "synthetic",
        "config": {"workload": w["desc"], "sf_per_gpu": args.sf, "rows_per_gpu": w["rows"], "sample_rows": sample, "batch_rows": 8192},
        "cpu_baseline": {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{sample} of the workload's {w['rows']} rows in 8192-row batches, single thread like the reference's executor",
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
    names = [k.strip() for k in kernel.split("+")]
    parts = [t.get(k) for k in names]
    found = [p for p in parts if p is not None]
    if not found:
        return None
    # an operator's small helper launches - offsets, prefix sums, finalize - are not captured: their traffic is kilobytes.
    # Every other launch of the operator must have its own capture, or the sum would understate the operator's traffic
    # (the two Filter workloads share kernel names but not kernels: only the Q6-selectivity one was captured).
    helpers = ("k_prefix_at", "k_scan_apply", "k_scan_block_sums", "k_scan_sums", "k_region_offsets", "k_agg_finalize", "k_group_", "k_gather",
               "k_key_minmax", "k_scan_filter_group_smem")
    for name, p in zip(names, parts):
        if p is None and not name.startswith(helpers):
            return None
    return sum(p["dram_bytes_per_launch"] for p in found)


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
    for kv in args.opt:
        k, v = kv.split("=", 1)
        ctx.set_option(k, int(v))
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
        w = make_workload(name, args.sf * world if name != "highcard" else args.sf, rank, world, dist)
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
        if args.no_check:
            checked = None
        elif world == 1:
            checked = numpy_check(w, res_batches)
        else:
            checked = distributed_check(w, res_batches, dist)
        del out, res_batches
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
            # one operator = several launches that only make sense together (the radix join: histogram + two scatter
            # passes + the build/probe kernel): its algorithmic bytes are set against the SUM of those launches
            members = {k: v for k, v in agg.items() if k.startswith(tuple(op_prefix))}
            if members:
                top = (" + ".join(sorted(members)), [sum(v[0] for v in members.values()), steps])
                top_ms = top[1][0] / max(1, steps)
        # algorithmic bytes of one step on this rank / the summed kernel time of one step
        alg_bytes = w["bytes_per_row"] * w["rows"]
        traffic = measured_traffic(top[0]) if world == 1 else None
        roofline = {
            "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
            "kernel": top[0], "kernel_ms": top_ms, "kernel_share_of_step": (top[1][0] / kern_total) if kern_total else None,
            "algorithmic_bytes_per_launch": alg_bytes,
            "achieved": alg_bytes / (top_ms / 1e3) / 1e9 if top_ms else None,
            "traffic": traffic,
            "traffic_note": ("dram bytes of one launch from the committed ncu --set full capture (profiles/traffic.json)" if traffic
                             else ("not measured at N > 1" if world > 1 else "no capture of this kernel set in profiles/traffic.json")),
        }
        roofline["frac"] = roofline["achieved"] / peak if roofline["achieved"] else None
        roofline["step_achieved"] = alg_bytes / (ms / steps / 1e3) / 1e9
        roofline["step_frac"] = roofline["step_achieved"] / peak
        kernels = {k: {"ms_avg": v[0] / v[1], "launches_per_step": v[1] / steps} for k, v in sorted(agg.items())}
        nvlink = None
        if world > 1 and w.get("exchange_bytes"):
            # bytes that leave this GPU in the exchange / the time of the kernels that move them (max over ranks)
            x_ms = max_over_ranks(sum(v["ms_avg"] * v["launches_per_step"] for k, v in kernels.items() if "nvlink" in k or k.startswith("nccl_all_to_all")))
            if x_ms > 0:
                gbs = w["exchange_bytes"] / (x_ms / 1e3) / 1e9
                nvlink = {"bytes_out_per_gpu": int(w["exchange_bytes"]), "exchange_ms": x_ms, "GB/s": gbs, "frac_of_900": gbs / 900.0}

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
        # bytes actually handed to the library for upload (an aggregate over a host scan uploads only the columns it reads)
        uploaded = (ex.h2d_bytes - uploaded0) // e2e_steps
        if uploaded > 0:
            h2d_bytes = int(uploaded)
        e2e = {"value": total_rows * e2e_steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps}
        del rb

        results[name] = {
            "desc": w["desc"], "value": value, "ms_per_step": ms / steps, "wall_ms_per_step": wall_ms / steps, "steps": steps,
            "rows_per_step": total_rows, "roofline": roofline, "kernels": kernels, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "checked": checked, "nvlink": nvlink,
        }
        if rank == 0 and world == 1 and not args.no_cpu:
            # the reference's CPU path beside it, on a bounded sample of the same workload: one thread (the reference's
            # executor is single-threaded) and every host core
            cpu_rows = args.cpu_rows if head else max(500_000, args.cpu_rows // 4)
            v, n = cpu_port_time(w, cpu_rows, 1)
            cores = os.cpu_count() or 1
            vn, _ = cpu_port_time(w, cpu_rows, cores) if name != "join" else (None, n)
            results[name]["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"first {n} rows of the same workload, 8192-row batches, single thread (the reference's executor is single-threaded)",
            }
            if vn is not None:
                results[name]["cpu_baseline"]["all_cores"] = {"value": vn, "cores": cores}
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
            "roofline": {k: h["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_note", "kernel", "kernel_ms",
                                                         "kernel_share_of_step", "peak_source", "step_achieved", "step_frac",
                                                         "algorithmic_bytes_per_launch")},
            "e2e": h["e2e"], "gpu_launches": h["gpu_launches"], "clocks": h["clocks"],
            "checked": h["checked"],
        }
        if "cpu_baseline" in h:
            line["cpu_baseline"] = h["cpu_baseline"]
        if len(names) > 1:
            # compact (and LAST on the line, so that a tail of the output still holds it): one short block per workload;
            # the per-kernel tables go to stderr and to gpurun_out/bench_detail_n<N>.json
            def short(r):
                o = {"value": float(f"{r['value']:.4g}"), "ms": round(r["ms_per_step"], 4), "frac": round(r["roofline"]["frac"] or 0.0, 3),
                     "step_frac": round(r["roofline"]["step_frac"], 3), "kernel_ms": round(r["roofline"]["kernel_ms"], 4),
                     "e2e": float(f"{r['e2e']['value']:.4g}"), "checked": r["checked"]}
                if r["roofline"]["traffic"]:
                    o["traffic_x"] = round(r["roofline"]["traffic"] / r["roofline"]["algorithmic_bytes_per_launch"], 2)
                if r.get("nvlink"):
                    o["nvlink_frac"] = round(r["nvlink"]["frac_of_900"], 3)
                if "cpu_baseline" in r:
                    o["cpu1"] = float(f"{r['cpu_baseline']['value']:.3g}")
                return o

            line["also"] = {n: short(results[n]) for n in names[1:]}
        detail = {n: {k: results[n][k] for k in ("desc", "value", "ms_per_step", "wall_ms_per_step", "steps", "roofline", "kernels", "e2e",
                                                 "checked", "nvlink", "gpu_launches") if k in results[n]} | (
                      {"cpu_baseline": results[n]["cpu_baseline"]} if "cpu_baseline" in results[n] else {}) for n in names}
        sys.stderr.write("bench detail: " + json.dumps(detail) + "\n")
        try:
            (ROOT / "gpurun_out").mkdir(exist_ok=True)
            (ROOT / "gpurun_out" / f"bench_detail_n{world}.json").write_text(json.dumps(detail, indent=1))
        except OSError:
            pass
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
    ap.add_argument("--also", default="q1,join,highcard,filter_q6,filter_q1",
                    help="further workloads measured after the headline one (comma list)")
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--ref-rows", type=int, default=0, help="reference arm: rows per step (0 = the whole table, the same config as ours)")
    ap.add_argument("--cpu-rows", type=int, default=4_000_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="library option (qgpu_set_option), repeatable")
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
