#!/bin/bash
# round-2 final single-GPU evidence, lean (the first attempt's reports exceeded the 64 MiB copy-back limit): default bench
# line, reference arm, ncu launch list of the bench command, one small full capture per workload
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_n1.err; echo "n1 rc=$?"
cp gpurun_out/bench_detail_n1.json gpurun_out/r2z_detail_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-check --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02z.csv $CMD > gpurun_out/ncu_list_r02z.log 2>&1; echo "ncu list rc=$?"
for W in q6 q1 join highcard filter_q6; do
  WCMD="python bench.py --workload $W --also= --steps 2 --warmup 3 --no-cpu --no-check --e2e-steps 1"
  timeout 600 ncu --set full --clock-control none -k regex:'qgpu_jit_scan|k_rscatter|k_pj_join2|k_region_hist|k_compact' --launch-skip 8 -c 5 -f -o gpurun_out/prof_r02z_$W $WCMD > gpurun_out/ncu_full_r02z_$W.log 2>&1
  echo "ncu full $W rc=$?"
done
du -sh gpurun_out
