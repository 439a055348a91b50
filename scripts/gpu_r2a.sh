#!/bin/bash
# round 2, first GPU call: the new radix scatter / join / GROUP BY paths - parity subset, then per-kernel timings
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_operators.py -q -x -k "join or group or high" 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.log
timeout 600 python bench.py --workload join --also highcard --steps 20 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2a_bench.json"))
    for name, blk in [("join", d)] + list(d.get("also", {}).items()):
        r = blk["roofline"]
        print(name, "ms/step", round(blk["ms_per_step"], 3), "frac", round(r["frac"], 3), "step_frac", round(r["step_frac"], 3), "checked", blk.get("checked_against_numpy"))
        for k, v in r["kernels"].items():
            print("   ", k, round(v["ms_avg"], 4), v["launches_per_step"])
except Exception as e:
    print("no bench json:", e)
PY
