import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import stochy_jl_b200 as sb
from gpu_util import cfg3_tables
A, logF, pi = cfg3_tables(T=7)
for N in [70001, 1 << 19, 1 << 20]:
    with sb.Engine(precision="fp32", resample="systematic", seed=21) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        lz = eng.smc_sweep(N, it=3, history=True)
        print(N, lz, eng.stats().kernel_launches, flush=True)
