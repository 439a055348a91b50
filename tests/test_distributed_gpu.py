"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): run with
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_distributed_gpu.py
The pytest wrapper launches that script when at least two devices are visible."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_rank_aggregate_and_join_match_single_gpu_oracle():
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29731", str(ROOT / "tests" / "run_distributed_gpu.py")],
        capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DISTRIBUTED PARITY OK" in r.stdout
