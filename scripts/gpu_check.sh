#!/bin/bash
# One GPU call: parity tests, smoke, bench, then the ncu launch list + one full capture of the top kernels.
# usage (under gpurun): bash scripts/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --steps 3 --warmup 3 --sf 10 --no-cpu --no-check --e2e-steps 2"
  $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
  echo "ncu list rc=$?"
  # one full capture per workload (the two JIT scan kernels share a name): a few steady-state launches of its kernels
  for W in q6 q1 join; do
    WCMD="python bench.py --workload $W --also= --steps 3 --warmup 3 --sf 10 --no-cpu --no-check --e2e-steps 2"
    ncu --set full --clock-control none --import-source on -k regex:'qgpu_jit_scan|k_pj_' --launch-skip 3 -c 3 -f -o gpurun_out/prof_${TAG}_$W $WCMD > gpurun_out/ncu_full_${TAG}_$W.log 2>&1
    echo "ncu full $W rc=$?"
  done
fi
