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


@pytest.mark.gpu
@pytest.mark.parametrize("world,log2n", [(2, 11), (2, 13), (4, 12), (8, 11)])
def test_sharded_sweep_ranks_on_one_gpu(world, log2n):
    """The same parity on a ONE-GPU box: `world` rank processes, all on device 0, each with its own arena mapped into
    the others over CUDA IPC exactly as across GPUs (gloo is the host plumbing; NCCL refuses two ranks on a device).
    The ranks' kernels are time-sliced, so every flag wait really waits for another context.  The sharded sweeps
    (HMM and linear-Gaussian, fp32 and fp64, odd T, several sweeps per handle) must equal the single-handle sweep of
    the same global pool bit for bit."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    env = dict(os.environ, SB_DIST_SAME_GPU="1", SB_DIST_SWEEPS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", str(29620 + world + log2n), os.path.join(ROOT, "tests", "dist_worker.py"), str(log2n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world,log2n,T", [(2, 11, 9), (2, 13, 16), (4, 12, 7), (8, 11, 5)])
def test_sharded_pmcmc_ranks_on_one_gpu(world, log2n, T):
    """Iterated conditional SMC on a SHARDED pool (src/pmcmc.jl:59-66, 83-94 across GPUs; retained entry in the last
    slot of the last rank, history rows sharded, the backward pass's counts and the retained lineage hopping between
    ranks through peer atomics) == sb_pmcmc_run on one handle: counts, retained trajectory, log-evidence and the
    history rows, bit for bit.  Rank processes share device 0 (CUDA IPC as across GPUs)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    env = dict(os.environ, SB_DIST_SAME_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + world + log2n), os.path.join(ROOT, "tests", "dist_pmcmc_worker.py"), str(log2n), str(T), "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_PMCMC_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_sharded_pmcmc_two_gpus():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tests", "dist_pmcmc_worker.py"), "14", "25", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_PMCMC_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
