#!/bin/bash
# GROUP BY partition pass with interleaved probing: parity, then geometry / occupancy variants
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "high or group" 2>&1 | tail -3 ) | tee gpurun_out/r2t_pytest.log
run() {
  timeout 600 python bench.py --workload highcard --also= --steps 6 --warmup 2 --no-cpu --e2e-steps 1 "$@" > gpurun_out/r2t_tmp.json 2> gpurun_out/r2t_tmp.err
  python - "$@" <<'PY'
import json, sys
d=json.load(open("gpurun_out/bench_detail_n1.json"))["highcard"]
print(" ".join(sys.argv[1:]) or "(defaults)", "| ms/step", round(d["ms_per_step"],3), "checked", d.get("checked"), {k: round(v["ms_avg"],3) for k,v in d["kernels"].items() if v["ms_avg"]>0.05})
PY
}
run
run --opt scan.jit_min_blocks=4
run --opt group.part_rows=2048
run --opt group.part_rows=2048 --opt scan.jit_min_blocks=4
run --opt scan.l2_prefetch=0
run --opt group.part_rows=2048 --opt scan.l2_prefetch=0
