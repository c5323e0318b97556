#!/usr/bin/env python
"""Diagnostics: phase stamps of the resident sweep kernel (SB_RES_TRACE=1) at BASELINE cfg3's size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stochy_jl_b200 as sb
from gpu_util import cfg3_tables
A, logF, pi = cfg3_tables(T=40)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
with sb.Engine(precision="fp32", resample="systematic", seed=1) as eng:
    eng.load(sb.DiscreteHMM(A, logF, pi=pi))
    eng.smc_sweep(N, it=0)
    eng.smc_sweep(N, it=1)
    print("device ms", eng.stats().device_ms)
