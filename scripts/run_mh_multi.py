#!/usr/bin/env python
"""BASELINE config 5 across GPUs: independent MH chain sets, one per GPU, no communication on the data path;
the per-rank histograms are summed at the end (one all_reduce of 2 integers).

  torchrun --nproc-per-node N scripts/run_mh_multi.py [chains_per_gpu] [steps]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import stochy_jl_b200 as sb
    from gpu_util import cfg5_bnet
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    sites, factors, query = cfg5_bnet()
    with sb.Engine(device=local, precision="fp64", seed=1) as eng:
        eng.load(sb.BayesNet(sites, factors, query))
        eng.mh_counts(chains, 10, chain0=rank * chains)                  # warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        counts = eng.mh_counts(chains, steps, chain0=rank * chains)     # global chain ids: disjoint Philox streams
        ms = eng.stats().device_ms
    t = torch.tensor([float(counts[0]), float(counts[1]), ms], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        c0, c1 = t[0].item(), t[1].item()
        print(json.dumps({"config": 5, "n_gpus": world, "chains_per_gpu": chains, "steps": steps,
                          "chain_steps_per_s": world * chains * steps / (tmax[2].item() * 1e-3),
                          "P(R|W)": c1 / (c0 + c1), "exact": 0.7079276773296245}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
