#!/bin/bash
# round 2, third GPU call: parity subset again, then option sweeps of the scatter / join / GROUP BY kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_operators.py -q -x -k "join or group or high" 2>&1 | tail -5 | tee gpurun_out/r2c_pytest.log
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    for name, blk in d.items():
        r = blk["roofline"]
        print(" ", name, "ms/step", round(blk["ms_per_step"], 3), "kernel_ms", round(r["kernel_ms"], 3), "frac", round(r["frac"], 3), "checked", blk.get("checked"),
              {k: round(v["ms_avg"], 3) for k, v in blk["kernels"].items() if v["ms_avg"] > 0.02})
except Exception as e:
    print("no detail:", e)
PY
}
run() {
  echo "== $*"
  timeout 600 python bench.py --workload join --also highcard --steps 12 --warmup 3 --no-cpu --e2e-steps 2 "$@" > gpurun_out/r2c_last.json 2> gpurun_out/r2c_last.err || tail -3 gpurun_out/r2c_last.err
  show gpurun_out/bench_detail_n1.json
}
run
run --opt scatter.stages=1
run --opt scatter.stages=1 --opt join.region_bytes2=16777216 --opt join.ring=3
run --opt scatter.stages=1 --opt join.region_bytes2=4194304 --opt join.ring=6
run --opt scatter.stages=1 --opt group.part_rows=2048 
run --opt scatter.stages=1 --opt group.part_rows=4096 --opt scan.rows_per_thread=8
