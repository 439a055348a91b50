#!/bin/bash
# host-side cost of a Q6 / Q1 step
set -u
mkdir -p gpurun_out
PROBE_WORKLOAD=q6 timeout 600 python scripts/q1_probe.py 10 2>&1 | tail -12 | tee gpurun_out/r2r_q6.log
PROBE_WORKLOAD=q1 PROBE_OPTS='[{}, {"scan.jit_min_blocks": 3}, {"scan.jit_min_blocks": 2, "scan.eager_loads": 1}]' timeout 600 python scripts/q1_probe.py 10 2>&1 | tail -14 | tee gpurun_out/r2r_q1.log
