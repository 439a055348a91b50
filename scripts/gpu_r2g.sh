#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -14 | tee gpurun_out/r2g_smoke.log
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_edge_cases.py -q -x -k "group or catalog or benchmark or scalar or employees" 2>&1 | tail -5 | tee gpurun_out/r2g_pytest.log
show() {
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_detail_n1.json"))
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
  rm -f gpurun_out/bench_detail_n1.json
  timeout 600 python bench.py --workload q1 --also q6 --steps 30 --warmup 5 --no-cpu --e2e-steps 2 "$@" > gpurun_out/r2g_last.json 2> gpurun_out/r2g_last.err || tail -3 gpurun_out/r2g_last.err
  show
}
run
run --opt scan.jit_min_blocks=3
run --opt scan.jit_min_blocks=4
run --opt scan.rows_per_thread=8
run --opt scan.rows_per_thread=8 --opt scan.jit_min_blocks=3
run --opt scan.eager_loads=1
