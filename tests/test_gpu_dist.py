"""Sharded pool across 2 GPUs vs the same pool on one GPU: bit-identical particles and log-evidence.
Needs two devices; the single-GPU box skips it (run by hand: gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("log2n", [11, 14])
def test_sharded_sweep_equals_single_gpu(log2n):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dist_worker.py"), str(log2n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
