#!/bin/bash
# 8 GPUs: the bench (Q6 headline + Q1 + join), results verified in-run through all-reduced invariants; bounded to fit the budget
set -u
mkdir -p gpurun_out
timeout 175 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --e2e-steps 2 --also q1,join > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err
echo "n8 rc=$?"; tail -2 gpurun_out/r2y_bench_n8.err | cut -c1-300
cp gpurun_out/bench_detail_n8.json gpurun_out/r2y_detail_n8.json 2>/dev/null
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2y_detail_n8.json"))
except Exception as e:
    print("no detail", e); raise SystemExit
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "value", f"{blk['value']:.4g}", "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "checked", blk.get("checked"), "e2e", f"{blk['e2e']['value']:.3g}", blk.get("nvlink"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
