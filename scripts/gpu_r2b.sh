#!/bin/bash
# round 2, second GPU call: shared-memory atomics microbench + ncu full captures (with source) of the new kernels
set -u
mkdir -p gpurun_out
./scripts/microbench/smem_atomics | tee gpurun_out/r2b_smem_atomics.log
for W in join highcard; do
  WCMD="python bench.py --workload $W --also= --steps 3 --warmup 3 --sf 10 --no-cpu --no-check --e2e-steps 2"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rscatter|k_pj_join2|k_region_hist|qgpu_jit_scan' --launch-skip 8 -c 6 -f -o gpurun_out/prof_r2b_$W $WCMD > gpurun_out/ncu_full_r2b_$W.log 2>&1
  echo "ncu full $W rc=$?"
done
ls -la gpurun_out/*.ncu-rep
