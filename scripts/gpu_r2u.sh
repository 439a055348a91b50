#!/bin/bash
# full ncu captures of the GROUP BY partition pass at two partition sizes (the second qgpu_jit_scan launch of a step)
set -u
mkdir -p gpurun_out
for V in default small; do
  OPTS=""
  if [ $V = small ]; then OPTS="--opt group.part_rows=2048"; fi
  WCMD="python bench.py --workload highcard --also= --steps 2 --warmup 2 --no-cpu --no-check --e2e-steps 1 --opt scan.jit_min_blocks=4 $OPTS"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qgpu_jit_scan' --launch-skip 4 -c 2 -f -o gpurun_out/prof_r2u_part_$V $WCMD > gpurun_out/ncu_r2u_$V.log 2>&1
  echo "ncu $V rc=$?"
done
