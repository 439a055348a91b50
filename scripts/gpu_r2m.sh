#!/bin/bash
# join2 tweaks (item order, release RED, 256 KB clears) + GROUP BY partition pass variants (sink x preload)
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x -k "join or high or group" 2>&1 | tail -4 ) | tee gpurun_out/r2m_pytest.log
export JOIN_PROBE_OPTS='[{"join.alias_key":1},{"join.alias_key":0},{"join.alias_key":1,"join.ring":4},{"join.alias_key":1,"join.dense":0}]'
timeout 600 python scripts/join_probe.py 10 2>&1 | grep -E "^\{|^    \{" | tee gpurun_out/r2m_join.log | cut -c1-300
for V in "1 1" "0 1" "1 0" "0 0"; do
  set -- $V
  timeout 600 python bench.py --workload highcard --also= --steps 8 --warmup 3 --no-cpu --e2e-steps 1 --opt group.defer_combine=$1 --opt group.preload=$2 > gpurun_out/r2m_highcard_$1$2.json 2> gpurun_out/r2m_highcard_$1$2.err; echo "highcard defer=$1 preload=$2 rc=$?"
  python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_detail_n1.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "checked", blk.get("checked"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
done
