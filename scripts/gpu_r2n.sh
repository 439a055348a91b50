#!/bin/bash
# join2 knobs one at a time; GROUP BY partition pass (atomic sink + preload) through the bench
set -u
mkdir -p gpurun_out
export JOIN_PROBE_OPTS='[{"join.alias_key":1},{"join.release_red":1},{"join.release_red":0,"join.clear_last":1},{"join.clear_last":0,"join.clear_kb":256},{"join.clear_kb":32},{"join.clear_kb":64,"join.release_red":1,"join.ring":4}]'
timeout 600 python scripts/join_probe.py 10 2>&1 | grep -E "^\{|^    \{" | tee gpurun_out/r2n_join.log | cut -c1-300
( timeout 900 python -m pytest tests -m gpu -q -x -k "high or group" 2>&1 | tail -3 ) | tee gpurun_out/r2n_pytest.log
timeout 600 python bench.py --workload highcard --also= --steps 8 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2n_highcard.json 2> gpurun_out/r2n_highcard.err; echo "highcard rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_detail_n1.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "checked", blk.get("checked"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
