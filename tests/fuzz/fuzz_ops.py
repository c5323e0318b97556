#!/usr/bin/env python
"""Randomised check of op-list programs (sb_model_ops) against the oracle's restatement: random scalar programs (normal /
beta / gamma / uniform / flip / keep draws, normal / bernoulli / poisson / zero scores, 0..4 observations per statement)
and Dirichlet programs, random pool sizes.  Transcendental functions differ in the last ulps between libm and CUDA, so a
weight can differ by 1e-15 relative and -- rarely -- move a resampling boundary; the check is therefore: the first step
(no parents) agrees to 1e-11 / 1e-9 always, the log-evidence of the whole sweep agrees to 1e-6 in at least 90 % of the cases
and to 0.3 always.  Test tooling: python tests/fuzz/fuzz_ops.py [cases] [seed]."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import stochy_jl_b200 as sb
import ops_util as ou
from oracle import oracle as orc
orc.build()
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad, close = 0, 0


def rand_scalar_program():
    T = int(rng.integers(1, 7)); steps = []
    kind = str(rng.choice(["unit", "real", "pos"]))                 # latent lives in (0,1), on the line, or on the positive half-line
    for t in range(T):
        ys_n = int(rng.integers(0, 5))
        if kind == "unit":
            draw = (ou.D_BETA, (float(rng.uniform(.3, 5)), float(rng.uniform(.3, 5)))) if (t == 0 or rng.random() < .2) else \
                   ((ou.D_UNIFORM, (0.05, 0.95)) if rng.random() < .2 else (ou.D_KEEP, ()))
            score = (ou.S_BERNOULLI, (0.9, 0.05), [float(v) for v in rng.integers(0, 2, ys_n)]) if ys_n else (ou.S_ZERO, (), [])
        elif kind == "pos":
            draw = (ou.D_GAMMA, (float(rng.uniform(.3, 6)), float(rng.uniform(.2, 3)))) if (t == 0 or rng.random() < .2) else (ou.D_KEEP, ())
            score = (ou.S_POISSON, (1.0, 0.1), [float(v) for v in rng.poisson(3.0, ys_n)]) if ys_n else (ou.S_ZERO, (), [])
        else:
            draw = (ou.D_NORMAL, ((0.0 if t == 0 else float(rng.uniform(-1, 1))), float(rng.normal()), float(rng.uniform(.2, 2))))
            if t > 0 and rng.random() < .15:
                draw = (ou.D_FLIP, (0.0, 0.3))
            score = (ou.S_NORMAL, (float(rng.uniform(.5, 2)), float(rng.normal()), float(rng.uniform(.3, 2))), list(rng.normal(size=ys_n))) if ys_n else (ou.S_ZERO, (), [])
        steps.append((draw[0], draw[1], score[0], score[1], score[2]))
    return steps


for c in range(cases):
    if rng.random() < 0.25:
        D = int(rng.integers(2, 9))
        steps = ou.dirichlet_categorical(alpha=tuple(float(a) for a in rng.uniform(.3, 4, D)), batches=int(rng.integers(1, 5)), per=int(rng.integers(1, 5)), seed=int(rng.integers(1 << 20)))[0] if D == 4 else None
        if steps is None:
            al = tuple(float(a) for a in rng.uniform(.3, 4, D)); B = int(rng.integers(1, 5))
            steps = [(ou.D_DIRICHLET if t == 0 else ou.D_KEEP, al if t == 0 else (), ou.S_CATEGORICAL, (), [float(v) for v in rng.integers(1, D + 1, int(rng.integers(1, 5)))]) for t in range(B)]
    else:
        steps = rand_scalar_program()
    N = int(rng.choice([100, 2048, 5000, 70001])); seed = int(rng.integers(1, 1 << 30)); it = int(rng.integers(0, 4))
    desc = dict(case=c, N=N, T=len(steps), draws=[s[0] for s in steps], scores=[s[2] for s in steps], seed=seed)
    try:
        m = orc.Ops(steps); o = orc.opts(orc.SYSTEMATIC, seed=seed)
        with sb.Engine(precision="fp64", seed=seed, store_logw=True) as eng:
            eng.load(ou.to_program(sb, steps[:1]))
            eng.smc_sweep(N, it=it)
            x0, lw0 = eng.particles(N), eng.logw_row(0, N)
        xr, lwr = orc.ops_step(m, o, N, 0, it, None)
        ok = np.allclose(x0, xr, rtol=1e-11, atol=1e-11) and np.allclose(lw0, lwr, rtol=1e-9, atol=1e-9)
        if not ok:
            desc["where"] = "first step: x %g lw %g" % (np.abs(x0 - xr).max(), np.nanmax(np.abs(lw0 - lwr)))
        else:
            with sb.Engine(precision="fp64", seed=seed) as eng:
                eng.load(ou.to_program(sb, steps))
                lz = eng.smc_sweep(N, it=it)
            ref = orc.ops_sweep(m, o, N, it=it)["logZ"]
            d = abs(lz - ref)
            close += d <= 1e-6 * max(1.0, abs(ref))
            if not d <= 0.3:
                ok = False; desc["where"] = "logZ %r vs %r" % (lz, ref)
    except (orc.OracleError, sb.StochyError) as e:
        ok = False; desc["where"] = "raised %r" % str(e)
    if not ok:
        bad += 1; print("MISMATCH", desc, flush=True)
print("fuzz ops: %d cases, %d mismatches, log-evidence within 1e-6 in %d" % (cases, bad, close), flush=True)
sys.exit(1 if bad or close < 0.9 * (cases - bad) else 0)
