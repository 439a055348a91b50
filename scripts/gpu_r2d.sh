#!/bin/bash
set -u
mkdir -p gpurun_out
./scripts/microbench/rank_ops | tee gpurun_out/r2d_rank_ops.log
timeout 1500 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_operators.py -q -k "not test_join_radix_partitioned" 2>&1 | tail -40 | tee gpurun_out/r2d_pytest.log
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
  rm -f gpurun_out/bench_detail_n1.json
  timeout 600 python bench.py --workload join --also highcard --steps 12 --warmup 3 --no-cpu --e2e-steps 2 "$@" > gpurun_out/r2d_last.json 2> gpurun_out/r2d_last.err || tail -3 gpurun_out/r2d_last.err
  show gpurun_out/bench_detail_n1.json
}
run --opt join.region_bytes2=16777216 --opt join.ring=3
run --opt join.region_bytes2=16777216 --opt join.ring=3 --opt scatter.rank_atomic=0
run --opt join.region_bytes2=16777216 --opt join.ring=3 --opt scatter.plain_stores=1
run --opt join.region_bytes2=16777216 --opt join.ring=4 --opt scatter.stages=1
