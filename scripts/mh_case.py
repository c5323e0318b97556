#!/usr/bin/env python
"""BASELINE cfg5 on one GPU (65536 chains x 10^4 steps of single-site MH on the sprinkler net): the case the MH ncu
summary under profiles/ is captured on.  Prints chain-steps/s and the pooled estimate of P(R | W)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stochy_jl_b200 as sb
from gpu_util import cfg5_bnet
sites, factors, query = cfg5_bnet()
chains, steps = 65536, 10000
with sb.Engine(precision="fp64", seed=3) as eng:
    eng.load(sb.BayesNet(sites, factors, query))
    for _ in range(2):
        c = eng.mh_counts(chains, steps)
        st = eng.stats()
    print({"chain_steps_per_s": chains * steps / (st.device_ms * 1e-3), "ms": st.device_ms, "P(R|W)": c[1] / c.sum()})
