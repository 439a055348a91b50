#!/bin/bash
# barrier-free directory of the shared-memory GROUP BY tier (Q1): parity first (under a timeout: a spin lock is new), then Q1 timing
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "group or agg or q1 or config1 or benchmark or high" 2>&1 | tail -4 ) | tee gpurun_out/r2w_pytest.log
PROBE_WORKLOAD=q1 timeout 300 python scripts/q1_probe.py 10 2>&1 | tail -5 | tee gpurun_out/r2w_q1.log
timeout 600 python bench.py --workload highcard --also= --steps 6 --warmup 2 --no-cpu --e2e-steps 1 > gpurun_out/r2w_hc.json 2> gpurun_out/r2w_hc.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_detail_n1.json"))["highcard"]
print("highcard ms/step", round(d["ms_per_step"],3), "checked", d.get("checked"), {k: round(v["ms_avg"],3) for k,v in d["kernels"].items() if v["ms_avg"]>0.05})
PY
