#!/bin/bash
# warp-specialised join kernel + deferred-combine GROUP BY sink: parity subset, then timings
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x -k "join or group or high or smoke or agg" 2>&1 | tail -8 ) | tee gpurun_out/r2l_pytest.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -4
export JOIN_PROBE_OPTS='[{"join.warp_specialised":0,"join.alias_key":0},{"join.warp_specialised":1,"join.alias_key":0},{"join.warp_specialised":1,"join.alias_key":1},{"join.warp_specialised":1,"join.stages":4},{"join.warp_specialised":1,"join.stages":3,"join.dense":0},{"join.warp_specialised":1,"join.stages":4,"join.dense":0},{"join.warp_specialised":0,"join.dense":0}]'
timeout 600 python scripts/join_probe.py 10 2>&1 | grep -E "^\{|^    \{" | tee gpurun_out/r2l_join.log | cut -c1-300
timeout 600 python bench.py --workload highcard --also= --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2l_highcard.json 2> gpurun_out/r2l_highcard.err; echo "highcard rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_detail_n1.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "checked", blk.get("checked"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
