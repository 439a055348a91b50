"""Turns the scratch ncu outputs of one gpurun call into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [--rep gpurun_out/prof_<tag>.ncu-rep] [--launches gpurun_out/launches_<tag>.csv]

Writes
    profiles/<tag>_launches.csv      the ncu launch list (gpu__time_duration.sum per launch), copied as is
    profiles/<tag>_launch_share.md   per-kernel totals and share of the profiled command
    profiles/<tag>_ncu_full.csv      one row per captured launch of the `--set full` report: duration, DRAM bytes,
                                     L2 hit rate, achieved occupancy, registers, throughput percentages
    profiles/<tag>_bench.json        the bench line of the same call (when present)
"""
from __future__ import annotations

import csv
import shutil
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",  # active lanes per executed instruction (32 = no divergence)
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(tag: str, path: Path) -> None:
    if not path.exists():
        return
    OUT.mkdir(exist_ok=True)
    shutil.copy(path, OUT / f"{tag}_launches.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for r in rows:
        name, ns = r[4], float(r[-1])
        agg.setdefault(name, []).append(ns)
    total = sum(sum(v) for v in agg.values())
    lines = [f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)", "",
             "| kernel | launches | total ms | mean us | share |", "|---|---:|---:|---:|---:|"]
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{name}` | {len(v)} | {sum(v) / 1e6:.3f} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / total:.1%} |")
    lines.append("")
    lines.append(f"total {total / 1e6:.3f} ms over {len(rows)} launches")
    (OUT / f"{tag}_launch_share.md").write_text("\n".join(lines) + "\n")


_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
_TIME_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def full(tag: str, reps) -> None:
    """one CSV over all captures of the tag; ncu picks a unit per column PER REPORT, so byte and time columns are
    normalised (bytes, microseconds) before the rows of different reports are put under one header"""
    reps = [r for r in reps if r.exists()]
    if not reps:
        return
    OUT.mkdir(exist_ok=True)
    wrote_header = False
    with open(OUT / f"{tag}_ncu_full.csv", "w", newline="") as f:
        w = csv.writer(f)
        for rep in reps:
            raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(raw.splitlines()))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in FULL_METRICS if m in hdr]

            def norm(i):
                if units[i] in _BYTES:
                    return "byte", _BYTES[units[i]]
                if units[i] in _TIME_US:
                    return "us", _TIME_US[units[i]]
                return units[i], None

            if not wrote_header:
                w.writerow([hdr[i] + (f" [{norm(i)[0]}]" if units[i] else "") for i in cols] + ["capture"])
                wrote_header = True
            for r in rows[2:]:
                vals = []
                for i in cols:
                    scale = norm(i)[1]
                    vals.append(r[i] if scale is None or i == cols[0] else repr(float(r[i].replace(",", "")) * scale))
                w.writerow(vals + [rep.name])


def traffic(tag: str) -> None:
    """profiles/traffic.json: DRAM bytes (read + write) per launch of the kernels bench.py names, from the `--set full`
    capture of this tag; bench.py copies the dominant kernel's figure into `roofline.traffic`."""
    import json

    path = OUT / f"{tag}_ncu_full.csv"
    if not path.exists():
        return
    rows = list(csv.reader(open(path)))
    hdr = rows[0]

    def col(prefix):
        for i, h in enumerate(hdr):
            if h.startswith(prefix):
                return i, 1.0  # full() normalised the byte columns to bytes
        return None, 1.0

    ir, sr = col("dram__bytes_read.sum")
    iw, sw = col("dram__bytes_write.sum")
    if ir is None or iw is None:
        return
    per = {}
    icap = hdr.index("capture") if "capture" in hdr else None
    for r in rows[1:]:
        name = r[0]
        cap = r[icap] if icap is not None else ""
        total = float(r[ir]) * sr + float(r[iw]) * sw
        keys = []
        if "qgpu_jit_scan" in name:
            if total < 0.2e9:
                continue  # the shared-memory GROUP BY tier giving up on a many-groups input after a tile or two
            if "highcard" in cap:
                keys = ["k_scan_filter_group_part_jit"]
            elif "filter" in cap:
                keys = ["k_scan_filter_mask_jit"]
            else:
                keys = ["k_scan_filter_agg_jit" if total < 2.6e9 else "k_scan_filter_group_smem_jit"]
        elif "k_pj_join2<" in name:
            keys = ["k_pj_join2_dense" if "<(bool)1>" in name or "<1>" in name else "k_pj_join2_hashed"]
        elif "k_compact_sparse" in name:
            keys = ["k_compact_sparse"]
        elif "k_compact" in name:
            keys = ["k_compact"]
        elif "k_region_hist" in name:
            keys = ["k_region_hist"]
        elif "k_rscatter" in name:
            if "highcard" in cap:
                keys = ["k_rscatter_groups_l1", "k_rscatter_groups_l2"]  # (same kernel, same bytes: two radix levels)
            else:
                keys = ["k_rscatter_probe" if total > 1.0e9 else "k_rscatter_build"]
        elif "k_pj_join<1>" in name:
            keys = ["k_pj_join_dense"]
        elif "k_pj_join<0>" in name:
            keys = ["k_pj_join_hashed"]
        elif "k_pj_scatter" in name:
            keys = ["k_pj_scatter_probe" if total > 1.0e9 else "k_pj_scatter_build"]
        for key in keys:
            per.setdefault(key, []).append(total)
    out = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v), "source": f"profiles/{tag}_ncu_full.csv"} for k, v in per.items()}
    (OUT / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")


def main() -> None:
    tag = sys.argv[1]
    args = sys.argv[2:]
    reps = sorted((ROOT / "gpurun_out").glob(f"prof_{tag}*.ncu-rep"))
    lst = ROOT / "gpurun_out" / f"launches_{tag}.csv"
    if "--rep" in args:
        reps = [Path(args[args.index("--rep") + 1])]
    if "--launches" in args:
        lst = Path(args[args.index("--launches") + 1])
    launches(tag, lst)
    full(tag, reps)
    traffic(tag)
    bench = ROOT / "gpurun_out" / f"bench_{tag}.json"
    if bench.exists() and bench.stat().st_size:
        shutil.copy(bench, OUT / f"{tag}_bench.json")
    print(sorted(p.name for p in OUT.glob(f"{tag}_*")))


if __name__ == "__main__":
    main()
