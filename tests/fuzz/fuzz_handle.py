#!/usr/bin/env python
"""Randomised hunt for state that leaks between calls on ONE handle: a long random sequence of operations -- load a
random discrete model, plain sweeps with and without history, PMCMC runs of random pool sizes (one-tile chain kernel,
resident sweep kernel, tiled kernels, cooperative backward pass), resampling hooks, a sweep that ends in the zero-mass
error, log-sum-exp -- every result checked against the oracle (bit for bit where the work is integer).
Test tooling: python tests/fuzz/fuzz_handle.py [ops] [seed]."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import stochy_jl_b200 as sb
from oracle import oracle as orc
orc.build()
ops = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
done = {}
SIZES = [7, 100, 2047, 2048, 5000, 70001, 300000, 1048577, 2100000]


def rand_model():
    K = int(rng.choice([1, 2, 5, 16, 17, 64])); T = int(rng.integers(1, 7))
    A = rng.random((K, K)) ** int(rng.choice([1, 4])); A[rng.random((K, K)) < 0.2] = 0.0
    A[np.arange(K), np.arange(K)] += 1e-3; A /= A.sum(axis=1, keepdims=True)
    pi = rng.random(K) + 1e-3; pi /= pi.sum()
    logF = rng.normal(size=(T, K)) * float(rng.choice([0.5, 3.0, 30.0]))
    hf = None
    if T > 2 and rng.random() < 0.3:
        hf = (rng.random(T) < 0.6).astype(np.uint8); hf[-1] = 1
    return A, logF, pi, hf


for prec in ["fp32", "fp64"]:
    seed = int(rng.integers(1, 1 << 30))
    with sb.Engine(precision=prec, resample="systematic", seed=seed) as eng:
        oo = orc.opts(orc.SYSTEMATIC, fp32=(prec == "fp32"), seed=seed)
        model = None
        for c in range(ops // 2):
            op = str(rng.choice(["load", "sweep", "sweep_hist", "pmcmc", "hook", "zero", "lse", "lg"], p=[.15, .2, .15, .25, .1, .05, .05, .05]))
            desc = dict(prec=prec, step=c, op=op)
            try:
                if op == "lg":                                       # another model family on the same handle, then back to a discrete one
                    Tl = int(rng.integers(1, 6)); y = rng.normal(size=Tl) * 2
                    par = dict(a=float(rng.uniform(-1, 1)), q=float(rng.uniform(.2, 2)), c=float(rng.uniform(.5, 2)), r=float(rng.uniform(.2, 2)),
                               m0=float(rng.normal()), P0=float(rng.uniform(.2, 2)), y=y)
                    Nl = int(rng.choice(SIZES[:7]))
                    eng.load(sb.LinearGaussian(**par))
                    lz = eng.smc_sweep(Nl, it=1)
                    ref = orc.lg_sweep(orc.Lg(par["a"], par["q"], par["c"], par["r"], par["m0"], par["P0"], y), oo, Nl, it=1)["logZ"]
                    # fp32: libm and CUDA's sqrt / sincos / log approximations differ at 1e-6, every weight moves a little and a
                    # per-cent of the children change parent: the two sweeps are different Monte-Carlo runs of the same
                    # filter.  What must hold on a reused handle: the SAME result as on a fresh handle, bit for bit.
                    ok = abs(lz - ref) <= 1e-6 * max(1.0, abs(ref)) if prec == "fp64" else abs(lz - ref) < 1.0
                    with sb.Engine(precision=prec, resample="systematic", seed=seed) as fresh:
                        fresh.load(sb.LinearGaussian(**par))
                        lz_fresh = fresh.smc_sweep(Nl, it=1)
                    ok = ok and lz_fresh == lz
                    if not ok:
                        desc["where"] = "lg logZ %r (fresh handle %r) vs oracle %r (N=%d, %r)" % (lz, lz_fresh, ref, Nl, {k: (v if k != "y" else v.tolist()) for k, v in par.items()})
                    model = None
                    done[op] = done.get(op, 0) + 1
                    if not ok:
                        bad += 1; print("MISMATCH", desc, flush=True)
                    continue
                if op == "load" or model is None:
                    model = rand_model()
                    A, logF, pi, hf = model
                    eng.load(sb.DiscreteHMM(A, logF, pi=pi, has_factor=hf))
                    om = orc.Hmm(A, logF, pi=pi, has_factor=hf)
                    if op == "load":
                        continue
                N = int(rng.choice(SIZES)); it = int(rng.integers(0, 5))
                desc.update(N=N, K=model[0].shape[0], T=model[1].shape[0], hf=None if model[3] is None else model[3].tolist())
                ok = True
                if op in ("sweep", "sweep_hist"):
                    ref = orc.hmm_sweep(om, oo, N, it=it, want_hist=(op == "sweep_hist"))
                    lz = eng.smc_sweep(N, it=it, history=(op == "sweep_hist"))
                    ok = abs(lz - ref["logZ"]) <= 1e-9 * max(1.0, abs(ref["logZ"]))
                    if ok and op == "sweep_hist":
                        for t in range(om.T):
                            x, a, _ = eng.history(t, N)
                            ok = ok and np.array_equal(x, ref["x"][t, :N]) and np.array_equal(a, ref["anc"][t, :N])
                elif op == "pmcmc":
                    N = min(N, 1048577); iters = int(rng.integers(1, 4))
                    ref = orc.hmm_pmcmc(om, oo, iters, N, want_joint=False)
                    marg, _ = eng.pmcmc_counts(iters, N, want_joint=False)
                    ok = np.array_equal(marg, ref["marg"])
                elif op == "hook":
                    n = int(rng.choice(SIZES)); w = rng.random(n) ** 3; u = float(rng.random())
                    ok = np.array_equal(eng.resample_systematic(w, u, n), orc.resample_systematic(w, u, n))
                elif op == "lse":
                    lw = rng.normal(size=int(rng.choice(SIZES))) * 10
                    ok = abs(eng.logsumexp(lw) - orc.logsumexp(lw)) < 1e-9 * max(1.0, abs(orc.logsumexp(lw)))
                elif op == "zero":                                   # a sweep that must fail -- and must leave the handle usable
                    K = model[0].shape[0]
                    lf = model[1].copy(); lf[-1, :] = -np.inf
                    eng.load(sb.DiscreteHMM(model[0], lf, pi=model[2], has_factor=model[3]))
                    try:
                        eng.smc_sweep(N, it=it); ok = False; desc["where"] = "no zero-mass error"
                    except sb.StochyError:
                        pass
                    eng.load(sb.DiscreteHMM(model[0], model[1], pi=model[2], has_factor=model[3]))
            except (orc.OracleError, sb.StochyError) as e:
                ok = False; desc["where"] = "raised %r" % str(e)
            done[op] = done.get(op, 0) + 1
            if not ok:
                bad += 1; print("MISMATCH", desc, flush=True)
print("fuzz handle: %d ops %r, %d mismatches" % (ops, done, bad), flush=True)
sys.exit(1 if bad else 0)
