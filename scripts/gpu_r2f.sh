#!/bin/bash
# 2 GPUs: full single-GPU suite (new: sort, general joins, computed keys, feedback-directed scans), distributed parity,
# the bench at N=1 (all workloads) and N=2
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_n1.err; echo "n1 rc=$?"; tail -2 gpurun_out/r2f_n1.err | cut -c1-300
cp gpurun_out/bench_detail_n1.json gpurun_out/r2f_detail_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "n2 rc=$?"; tail -3 gpurun_out/r2f_bench_n2.err | cut -c1-300
cp gpurun_out/bench_detail_n2.json gpurun_out/r2f_detail_n2.json
python - <<'PY'
import json
for f in ("gpurun_out/r2f_detail_n1.json","gpurun_out/r2f_detail_n2.json"):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "missing", e); continue
    print(f)
    for name, blk in d.items():
        r=blk["roofline"]
        print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "step_frac", round(r["step_frac"],3),"checked", blk.get("checked"), "e2e", f"{blk['e2e']['value']:.3g}", blk.get("nvlink"), blk.get("cpu_baseline",{}).get("value"))
        print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
