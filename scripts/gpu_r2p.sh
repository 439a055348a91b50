#!/bin/bash
# k_compact with grouped column loads (parity + the two filter workloads); full ncu capture of the GROUP BY partition pass
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "filter or limit or join" 2>&1 | tail -3 ) | tee gpurun_out/r2p_pytest.log
timeout 600 python bench.py --workload filter_q6 --also filter_q1,join --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2p_filter.json 2> gpurun_out/r2p_filter.err; echo "filter rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_detail_n1.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "checked", blk.get("checked"))
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
WCMD="python bench.py --workload highcard --also= --steps 2 --warmup 2 --no-cpu --no-check --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qgpu_jit_scan' --launch-skip 4 -c 1 -f -o gpurun_out/prof_r2p_highcard_part $WCMD > gpurun_out/ncu_r2p.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2p.log
