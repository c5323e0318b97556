#!/usr/bin/env python
"""Is the TILED sweep (one step kernel + one tile scan per step) bound by kernel launches at the pool sizes the resident
sweep kernel does not take (> ~1.2 M particles)?  Device time per step (CUDA events around the sweep) against the host's
enqueue time per step (wall clock of the call minus the final synchronisation is not separable, so: wall per step of a
call whose device work is known) for 2^21 .. 2^24 particles, and for 2^20 with the resident kernel switched off."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stochy_jl_b200 as sb
from gpu_util import cfg3_tables
T = 400
A, logF, pi = cfg3_tables(T=T)
for log2n, nores in [(20, "1"), (20, None), (21, None), (22, None), (23, None), (24, None)]:
    if nores:
        os.environ["SB_NO_RESIDENT"] = nores
    else:
        os.environ.pop("SB_NO_RESIDENT", None)
    with sb.Engine(precision="fp32", resample="systematic", seed=1) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        eng.smc_sweep(1 << log2n, it=0)
        t0 = time.perf_counter()
        eng.smc_sweep(1 << log2n, it=1)
        wall = time.perf_counter() - t0
        st = eng.stats()
    print("N=2^%d %s: device %.2f us/step, wall %.2f us/step, launches %d" % (
        log2n, "tiled (resident off)" if nores else ("resident" if st.kernel_launches < T else "tiled"),
        st.device_ms * 1e3 / T, wall * 1e6 / T, st.kernel_launches), flush=True)
