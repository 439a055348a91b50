#!/bin/bash
# 2 GPUs: distributed parity (incl. the one-round-trip merge of small group sets and its fall-back), then the bench at N=2
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/run_distributed_gpu.py 2>&1 | grep -E "rank0|PARITY|Error|error|Traceback|assert" | tail -24 | tee gpurun_out/r2o_dist2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 --also q1,join > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
echo "n2 rc=$?"; tail -3 gpurun_out/r2o_bench_n2.err | cut -c1-300
cp gpurun_out/bench_detail_n2.json gpurun_out/r2o_detail_n2.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2o_detail_n2.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "value", f"{blk['value']:.4g}", "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "checked", blk.get("checked"), "e2e", f"{blk['e2e']['value']:.3g}", blk.get("nvlink"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.01})
PY
