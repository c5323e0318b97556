#!/usr/bin/env python
"""Small multi-tile-per-CTA case for compute-sanitizer (racecheck / memcheck): one resident CTA per SM so that
every CTA of the persistent step kernel walks several tiles and reuses its shared-memory scratch.
  SB_CTAS_PER_SM=1 compute-sanitizer --tool racecheck python scripts/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import stochy_jl_b200 as sb
    from gpu_util import cfg3_tables, cfg4_data
    N = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
    A, logF, pi = cfg3_tables(T=3)
    with sb.Engine(precision="fp32", resample="systematic", seed=2) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        print("hmm logZ", eng.smc_sweep(N, it=0))
        marg, _ = eng.pmcmc_counts(2, 3000, want_joint=False)
        print("pmcmc counts", int(marg.sum()))
    d = cfg4_data(T=3)
    with sb.Engine(precision="fp32", seed=2) as eng:
        eng.load(sb.LinearGaussian(**d))
        print("lg logZ", eng.smc_sweep(N, it=0))
    w = np.random.default_rng(0).random(100000)
    with sb.Engine(precision="fp64") as eng:
        print("anc", eng.resample_systematic(w, 0.3, 100000)[:3])


if __name__ == "__main__":
    main()
