#!/bin/bash
# sparse compaction kernel + one-round-trip op_filter: parity, then the filter workloads
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "filter or limit or join or edge or config1 or sort" 2>&1 | tail -3 ) | tee gpurun_out/r2q_pytest.log
timeout 600 python bench.py --workload filter_q6 --also filter_q1 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2q_filter.json 2> gpurun_out/r2q_filter.err; echo "filter rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_detail_n1.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "checked", blk.get("checked"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.004})
PY
