#!/bin/bash
# 8 GPUs: distributed parity on 8 ranks, then the bench (all workloads, in-run invariant checks)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/run_distributed_gpu.py 2>&1 | grep -E "rank0|PARITY|Error|error|Traceback" | tail -20 | tee gpurun_out/r2i_dist8.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err
echo "n8 rc=$?"; tail -3 gpurun_out/r2i_bench_n8.err | cut -c1-300
cp gpurun_out/bench_detail_n8.json gpurun_out/r2i_detail_n8.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_detail_n8.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "value", f"{blk['value']:.4g}", "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "checked", blk.get("checked"), "e2e", f"{blk['e2e']['value']:.3g}", blk.get("nvlink"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
tail -c 1200 gpurun_out/r2i_bench_n8.json
