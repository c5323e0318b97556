#!/usr/bin/env python
"""Diagnostics: where BASELINE cfg3 (K=16, T=1000, N=2^20) spends its time -- plain sweeps with and without history,
resident (one cooperative launch) vs tiled kernels, and the PMCMC iteration (sweep + backward counting)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stochy_jl_b200 as sb
from gpu_util import cfg3_tables
A, logF, pi = cfg3_tables(T=1000)
N = 1 << 20
with sb.Engine(precision="fp32", resample="systematic", seed=1) as eng:
    eng.load(sb.DiscreteHMM(A, logF, pi=pi))
    out = {}
    for hist in [False, True]:
        eng.smc_sweep(N, it=0, history=hist)
        eng.smc_sweep(N, it=1, history=hist)
        st = eng.stats()
        out["sweep_ms_hist%d" % hist] = st.device_ms
        out["launches_hist%d" % hist] = st.kernel_launches
    eng.pmcmc_counts(1, N, want_joint=False)
    eng.pmcmc_counts(3, N, want_joint=False)
    st = eng.stats()
    out["pmcmc_ms_per_iter"] = st.device_ms / 3
    out["us_per_step_pmcmc"] = st.device_ms / 3
    print(json.dumps(out))
