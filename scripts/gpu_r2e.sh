#!/bin/bash
# 2 GPUs: distributed parity script, then the bench with its in-run invariant checks
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_distributed_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/r2e_pytest.log
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_edge_cases.py -q -x -k "join or general or employees" 2>&1 | tail -8 | tee gpurun_out/r2e_pytest1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
echo "bench rc=$?"; tail -5 gpurun_out/r2e_bench_n2.err | cut -c1-600; cat gpurun_out/r2e_bench_n2.json | cut -c1-3000
timeout 600 python bench.py --workload join --also highcard --steps 12 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_n1.err; tail -2 gpurun_out/r2e_n1.err | cut -c1-1500
