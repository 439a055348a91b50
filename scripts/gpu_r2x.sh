#!/bin/bash
# round-2 final single-GPU evidence: smoke, all GPU tests, the default bench, ncu launch list + one full capture per workload
set -u
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -12 | tee gpurun_out/r2x_smoke.log
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r2x_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_n1.err; echo "n1 rc=$?"
cp gpurun_out/bench_detail_n1.json gpurun_out/r2x_detail_n1.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2x_detail_n1.json"))
for name, blk in d.items():
    r=blk["roofline"]
    print(" ", name, "ms/step", round(blk["ms_per_step"],3), "kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"] or 0,3), "step_frac", round(r["step_frac"],3),"checked", blk.get("checked"), "e2e", f"{blk['e2e']['value']:.3g}")
    print("     ", {k: round(v["ms_avg"],3) for k,v in blk["kernels"].items() if v["ms_avg"]>0.02})
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2x_bench_ref.json 2> gpurun_out/r2x_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2x_bench_ref.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 2"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02x.csv $CMD > gpurun_out/ncu_list_r02x.log 2>&1; echo "ncu list rc=$?"
for W in q6 q1 join highcard filter_q6; do
  WCMD="python bench.py --workload $W --also= --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 2"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qgpu_jit_scan|k_rscatter|k_pj_join2|k_region_hist|k_compact' --launch-skip 10 -c 6 -f -o gpurun_out/prof_r02x_$W $WCMD > gpurun_out/ncu_full_r02x_$W.log 2>&1
  echo "ncu full $W rc=$?"
done
