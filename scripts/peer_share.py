#!/usr/bin/env python
"""Peer-load share of a sharded sweep, counted from the ancestors themselves (the driver's NVLink byte counters are not
exposed on this pool: `nvidia-smi nvlink -gt d` prints N/A, and ncu cannot replay a kernel that sits behind a cross-GPU
flag barrier): sweep ONE pool of 2^22 particles with history on one GPU, cut it into G equal shards, and count the
children whose parent lives in another shard (state read over NVLink) and the children tiles whose candidate parent
tiles start in another shard (weights by TMA over NVLink).  The results do not depend on G, so this IS what G GPUs do."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stochy_jl_b200 as sb
from gpu_util import cfg3_tables
T, N = 24, 1 << 22
A, logF, pi = cfg3_tables(T=T)
os.environ["SB_NO_RESIDENT"] = "1"
with sb.Engine(precision="fp32", resample="systematic", seed=1) as eng:
    eng.load(sb.DiscreteHMM(A, logF, pi=pi))
    eng.smc_sweep(N, it=0, history=True)
    anc = np.stack([eng.history(t, N)[1] for t in range(T - 1)]).astype(np.int64)   # [T-1, N] ancestors drawn after step t
child = np.arange(N)
for G in [2, 4, 8]:
    nl = N // G
    cross = (anc // nl != child // nl).mean()
    # a children tile (2048) whose FIRST parent is in another shard pulls its first weight tile from a peer
    first = anc[:, ::2048]
    tcross = (first // nl != child[::2048] // nl).mean()
    disp = np.abs(anc - child).max()
    print("G=%d: %.4f %% of the children have their parent on another GPU, %.3f %% of the children tiles start in a peer's tile; "
          "largest |parent - child| = %d particles (%.2f %% of a shard)" % (G, 100 * cross, 100 * tcross, disp, 100.0 * disp / nl), flush=True)
