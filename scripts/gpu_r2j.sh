#!/bin/bash
# baseline of the restored tree: smoke, GPU parity tests, 1-GPU bench with per-kernel table
set -u
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -14 | tee gpurun_out/r2j_smoke.log
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r2j_pytest.log
( time timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_n1.err ) 2>&1 | tail -4; echo "n1 rc=$?"
cp gpurun_out/bench_detail_n1.json gpurun_out/r2j_detail_n1.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j_detail_n1.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "step_frac", round(r["step_frac"],3),"checked", blk.get("checked"), "e2e", f"{blk['e2e']['value']:.3g}")
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
tail -c 1500 gpurun_out/r2j_bench_n1.json
