#!/bin/bash
# join: option sweep (ring / region size), then one full ncu capture of k_pj_join2 with source
set -u
mkdir -p gpurun_out
export JOIN_PROBE_OPTS='[{"join.ring":3,"join.region_bytes2":16777216},{"join.ring":4,"join.region_bytes2":16777216},{"join.ring":4,"join.region_bytes2":8388608},{"join.ring":6,"join.region_bytes2":8388608},{"join.ring":8,"join.region_bytes2":4194304},{"join.ring":3,"join.region_bytes2":33554432},{"join.ring":2,"join.region_bytes2":33554432}]'
timeout 600 python scripts/join_probe.py 10 2>&1 | grep -v "^ *$" | tee gpurun_out/r2k_sweep.log | cut -c1-400
CMD="python scripts/join_probe.py 10 once"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pj_join2' --launch-skip 1 -c 1 -f -o gpurun_out/prof_r2k_join2 $CMD > gpurun_out/ncu_r2k.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2k.log
