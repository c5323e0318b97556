#!/usr/bin/env python
"""Runs the five BASELINE.json configs on one GPU at full size and prints one JSON line per config
(timings are CUDA-event device times reported by the library; correctness figures beside them).

  python scripts/run_configs.py [--only 3] [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Exact:
    """Closed forms the configs are checked against, from the PRODUCT's own Enumerate on the GPU (sb_enum_hmm: exhaustive
    joint / scaled forward-backward, itself tested against the oracle in tests/) and a scalar Kalman filter -- this script
    does not touch oracle/."""

    def __init__(self, sb):
        self.sb = sb

    def _enum(self, A, logF, want_joint, **kw):
        with self.sb.Engine(precision="fp64") as e:
            e.load(self.sb.DiscreteHMM(A, logF, **kw))
            return e.enum_hmm(want_joint=want_joint)

    def hmm_enumerate_joint(self, A, logF, **kw):
        joint, _, lz = self._enum(A, logF, True, **kw)
        return float(np.exp(lz)), joint / joint.sum()

    def hmm_forward(self, A, logF, **kw):
        return self._enum(A, logF, False, **kw)[2], None

    @staticmethod
    def kalman_loglik(a, q, c, r, m0, P0, y):
        m, P, ll = m0, P0, 0.0
        for t, yt in enumerate(y):
            if t:
                m, P = a * m, a * a * P + q
            S = c * c * P + r
            ll += -0.5 * (np.log(2 * np.pi * S) + (yt - c * m) ** 2 / S)
            G = P * c / S
            m, P = m + G * (yt - c * m), (1 - G * c) * P
        return float(ll), None


def tv(p, q):
    keys = set(p) | set(q)
    return 0.5 * sum(abs(p.get(k, 0.0) - q.get(k, 0.0)) for k in keys)


def cfg1(sb, enumref, quick):
    """examples/hmm.jl, enum() -> pmcmc(1000, 100): reference numerics (literal mode), vs exact enumeration."""
    m = sb.hmm_example(T=3)
    t0 = time.perf_counter()
    d = sb.pmcmc(m, 200 if quick else 1000, 100, seed=1)
    wall = time.perf_counter() - t0
    Z, probs = enumref.hmm_enumerate_joint(m.A, m.logF, x0=m.x0, has_factor=m.has_factor)
    exact = {}
    for code in np.nonzero(probs)[0]:
        xs = tuple((int(code) // m.K ** t) % m.K for t in range(m.T))
        v = m.value(xs)
        exact[v] = exact.get(v, 0.0) + float(probs[code])
    got = {k: d.prob(k) for k in d.support()}
    return {"config": 1, "what": "examples/hmm.jl PMCMC 100 particles x 1000 iterations (literal reference numerics)",
            "tv_vs_enumerate": tv(got, exact), "wall_s": wall}


def cfg2(sb, enumref, quick):
    from gpu_util import cfg2_tables
    A, logF, pi = cfg2_tables()
    iters = 200 if quick else 1000
    out = {"config": 2, "what": "HMM T=10 K=3, PMCMC 100 particles x %d iterations vs Enumerate (3^10 paths)" % iters}
    Z, probs = enumref.hmm_enumerate_joint(A, logF, pi=pi)
    for mode in ["literal", "systematic"]:
        with sb.Engine(precision="fp64", resample=mode, seed=1) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            marg, joint = eng.pmcmc_counts(iters, 100, want_joint=True)
            st = eng.stats()
        j = joint / joint.sum()
        out["tv_joint_" + mode] = float(0.5 * np.abs(j - probs).sum())
        out["device_ms_" + mode] = st.device_ms
    return out


def cfg3(sb, enumref, quick):
    from gpu_util import cfg3_tables
    T = 200 if quick else 1000
    A, logF, pi = cfg3_tables(T=T)
    N = 1 << 20
    exact, _ = enumref.hmm_forward(A, logF, pi=pi)
    with sb.Engine(precision="fp32", resample="systematic", seed=1) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        eng.pmcmc_counts(1, N, want_joint=False)                       # warm-up (allocations)
        sweeps = 3 if quick else 10
        marg, _ = eng.pmcmc_counts(1 + sweeps, N, want_joint=False)    # 1 plain + `sweeps` conditional sweeps, history on
        st = eng.stats()
    # log Z-hat of the first (unconditional) sweep against the forward algorithm over 32 seeds
    lzs = []
    for seed in range(4 if quick else 32):
        with sb.Engine(precision="fp32", resample="systematic", seed=100 + seed) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            lzs.append(eng.smc_sweep(N, it=0))
    lzs = np.array(lzs)
    sd = float(lzs.std(ddof=1))
    return {"config": 3, "what": "HMM T=%d K=16 N=2^20 conditional SMC with history, %d sweeps" % (T, 1 + sweeps),
            "particle_steps_per_s": st.particle_steps / (st.device_ms * 1e-3), "device_ms": st.device_ms,
            "logZ_first": st.logZ_first, "logZ_forward": exact, "abs_err": abs(st.logZ_first - exact),
            "logZ_seeds": len(lzs), "logZ_mean": float(lzs.mean()), "logZ_sd": sd,
            "delta_over_se": float(abs(lzs.mean() - exact) / (sd / np.sqrt(len(lzs))))}


def cfg4(sb, enumref, quick):
    from gpu_util import cfg4_data
    d = cfg4_data(T=100 if quick else 500)
    exact, _ = enumref.kalman_loglik(d["a"], d["q"], d["c"], d["r"], d["m0"], d["P0"], d["y"])
    out = {"config": 4, "what": "linear-Gaussian T=%d N=2^24 plain SMC, systematic resampling every step" % len(d["y"]),
           "logZ_kalman": exact}
    N = 1 << 24
    # |log Z-hat - Kalman| over 16 seeds (fp32)
    lzs = []
    for seed in range(2 if quick else 16):
        with sb.Engine(precision="fp32", seed=200 + seed) as eng:
            eng.load(sb.LinearGaussian(**d))
            lzs.append(eng.smc_sweep(N, it=0))
    lzs = np.array(lzs)
    out["logZ_seeds"] = len(lzs); out["logZ_mean_fp32"] = float(lzs.mean()); out["logZ_sd_fp32"] = float(lzs.std(ddof=1))
    out["delta_over_se_fp32"] = float(abs(lzs.mean() - exact) / (lzs.std(ddof=1) / np.sqrt(len(lzs))))
    for prec in ["fp32", "fp64"]:
        with sb.Engine(precision=prec, seed=3) as eng:
            eng.load(sb.LinearGaussian(**d))
            eng.smc_sweep(N, it=0)
            lz = eng.smc_sweep(N, it=1)
            st = eng.stats()
        out["logZ_" + prec] = lz
        out["particle_steps_per_s_" + prec] = st.particle_steps / (st.device_ms * 1e-3)
    return out


def cfg5(sb, enumref, quick):
    from gpu_util import cfg5_bnet
    sites, factors, query = cfg5_bnet()
    chains, steps = 65536, 1000 if quick else 10000
    with sb.Engine(precision="fp64", seed=1) as eng:
        eng.load(sb.BayesNet(sites, factors, query))
        eng.mh_counts(chains, 10)
        counts = eng.mh_counts(chains, steps)
        st = eng.stats()
    return {"config": 5, "what": "batched MH, %d chains x %d steps, Bayes net C,S,R | W" % (chains, steps),
            "chain_steps_per_s": chains * steps / (st.device_ms * 1e-3), "P(R|W)": float(counts[1] / counts.sum()),
            "exact": 0.7079276773296245, "accept_rate": st.accepts / (chains * steps)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import stochy_jl_b200 as sb
    enumref = Exact(sb)
    for i, f in enumerate([cfg1, cfg2, cfg3, cfg4, cfg5], 1):
        if args.only and args.only != i:
            continue
        print(json.dumps(f(sb, enumref, args.quick)), flush=True)


if __name__ == "__main__":
    main()
