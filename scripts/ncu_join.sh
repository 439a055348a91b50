#!/bin/bash
# ncu capture of the join kernels at SF10 (one GPU)
set -u
TAG=${1:-join}
CMD="python scripts/join_probe.py 10 once"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_pj_join' -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/plain_$TAG.log
