#!/usr/bin/env python
"""Randomised parity hunt: random discrete models (K, T, zeros / tiny probabilities, hard factors, steps without a
factor, fixed or drawn first state), random pool sizes around tile boundaries, every resampling mode and precision,
plain and conditional sweeps (PMCMC counts) -- GPU against the oracle, bit for bit.  Uses the oracle: test tooling, run
by hand (python tests/fuzz/fuzz_sweep.py [cases] [seed]); prints the failing configuration and exits 1 on a mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import stochy_jl_b200 as sb
from oracle import oracle as orc
orc.build()
MODES = {"systematic": orc.SYSTEMATIC, "multinomial": 1, "literal": 2}
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for c in range(cases):
    K = int(rng.choice([1, 2, 3, 5, 16, 17, 37, 64]))
    T = int(rng.integers(1, 9))
    N = int(rng.choice([1, 7, 100, 2047, 2048, 2049, 4095, 4096, 5000, 20000, 70001]))
    if os.environ.get("FUZZ_BIG"):                                   # resident sweep kernel / many tiles per CTA of the tiled kernels
        N = int(rng.choice([300000, 1048575, 1048577, 1212415, 1212416, 2100000, 4200001]))
    mode = str(rng.choice(["systematic", "systematic", "multinomial", "literal"]))
    if mode == "literal":
        N = min(N, 3000)
    if mode == "multinomial":
        N = min(N, 300000)
    prec = str(rng.choice(["fp32", "fp64"]))
    A = rng.random((K, K)) ** int(rng.choice([1, 3, 6]))
    A[rng.random((K, K)) < rng.choice([0.0, 0.2, 0.5])] = 0.0
    A[np.arange(K), np.arange(K)] += 1e-3
    A /= A.sum(axis=1, keepdims=True)
    pi = rng.random(K) ** 4 + 1e-6; pi /= pi.sum()
    logF = rng.normal(size=(T, K)) * float(rng.choice([0.5, 3.0, 30.0]))
    if K > 2 and rng.random() < 0.5:
        logF[:, int(rng.integers(K))] = -np.inf                      # a state with no weight at all
    hf = None
    if T > 2 and rng.random() < 0.3:
        hf = (rng.random(T) < 0.6).astype(np.uint8); hf[-1] = 1
    x0 = int(rng.integers(K)) if rng.random() < 0.3 else -1
    seed = int(rng.integers(1, 1 << 30))
    iters = int(rng.integers(1, 4))
    desc = dict(case=c, K=K, T=T, N=N, mode=mode, prec=prec, x0=x0, hf=None if hf is None else hf.tolist(), seed=seed, iters=iters)
    if os.environ.get("FUZZ_ONLY") and int(os.environ["FUZZ_ONLY"]) != c:
        continue
    if os.environ.get("FUZZ_ONLY"):                                  # verbose replay of one case
        print(desc, "logF", logF.tolist(), flush=True)
        for small in ["1", None]:
            if small: os.environ["SB_NO_SMALL"] = small
            else: os.environ.pop("SB_NO_SMALL", None)
            for what in ["sweep", "sweep_hist", "pmcmc"]:
                try:
                    with sb.Engine(precision=prec, resample=mode, seed=seed) as eng:
                        eng.load(sb.DiscreteHMM(A, logF, x0=x0, pi=pi, has_factor=hf))
                        if what == "sweep": r = eng.smc_sweep(N, it=0)
                        elif what == "sweep_hist": r = eng.smc_sweep(N, it=0, history=True)
                        else: r = eng.pmcmc_counts(iters, N, want_joint=False)[0].sum()
                    print("  SB_NO_SMALL=%s %s ok %r" % (small, what, r), flush=True)
                except sb.StochyError as e:
                    print("  SB_NO_SMALL=%s %s RAISED %s" % (small, what, e), flush=True)
    try:
        om = orc.Hmm(A, logF, x0=x0, pi=pi, has_factor=hf)
        oo = orc.opts(MODES[mode], fp32=(prec == "fp32"), seed=seed)
        ref = orc.hmm_sweep(om, oo, N, it=0)
        refc = orc.hmm_pmcmc(om, oo, iters, N, want_joint=False)
        ok = True
        for small in (["1", None] if N + 1 <= 2048 else [None]):      # one-tile pools: the tiled kernels too, not only the chain kernel
            if small:
                os.environ["SB_NO_SMALL"] = small
            else:
                os.environ.pop("SB_NO_SMALL", None)
            with sb.Engine(precision=prec, resample=mode, seed=seed) as eng:
                eng.load(sb.DiscreteHMM(A, logF, x0=x0, pi=pi, has_factor=hf))
                lz = eng.smc_sweep(N, it=0, history=True)
                for t in range(T):
                    x, a, _ = eng.history(t, N)
                    if not (np.array_equal(x, ref["x"][t, :N]) and np.array_equal(a, ref["anc"][t, :N])):
                        ok = False; desc["where"] = "history row %d (SB_NO_SMALL=%s)" % (t, small); break
                # literal mode: the oracle sums exp(score) in fp64 while the GPU's log-evidence (an extension there: the
                # reference has none) comes from the 27-bit weight grid, whose floor biases a step by ~1e-7 relative
                tol = 1e-5 if mode == "literal" else 1e-9
                if ok and not (abs(lz - ref["logZ"]) <= tol * max(1.0, abs(ref["logZ"]))):
                    ok = False; desc["where"] = "logZ %r vs %r" % (lz, ref["logZ"])
                if ok:
                    marg, _ = eng.pmcmc_counts(iters, N, want_joint=False)
                    if not np.array_equal(marg, refc["marg"]):
                        ok = False; desc["where"] = "pmcmc counts (SB_NO_SMALL=%s)" % small
            if not ok:
                break
    except sb.StochyError as e:
        ok = False; desc["where"] = "GPU raised %r, the oracle did not" % str(e)
    except orc.OracleError as e:                                      # zero mass etc.: the GPU must raise the same condition
        try:
            with sb.Engine(precision=prec, resample=mode, seed=seed) as eng:
                eng.load(sb.DiscreteHMM(A, logF, x0=x0, pi=pi, has_factor=hf))
                eng.smc_sweep(N, it=0)
            ok = False; desc["where"] = "oracle raised %s, GPU did not" % e
        except sb.StochyError:
            ok = True
    if not ok:
        bad += 1
        print("MISMATCH", desc, flush=True)
print("fuzz: %d cases, %d mismatches" % (cases, bad), flush=True)
sys.exit(1 if bad else 0)
