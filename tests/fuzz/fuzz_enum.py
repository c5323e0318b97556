#!/usr/bin/env python
"""Randomised check of Enumerate on the GPU (sb_enum_hmm: exhaustive joint + scaled forward-backward marginals) against
the oracle's restatement of src/enumerate.jl:36-85 on random small discrete models (zeros in the rows, hard factors,
steps without a factor, fixed first state).  Test tooling: python tests/fuzz/fuzz_enum.py [cases] [seed]."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import stochy_jl_b200 as sb
from oracle import enumerate_ref as er
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
with sb.Engine(precision="fp64") as eng:
    for c in range(cases):
        K = int(rng.integers(1, 6)); T = int(rng.integers(1, 8))
        A = rng.random((K, K)) ** 2; A[rng.random((K, K)) < 0.25] = 0.0
        A[np.arange(K), np.arange(K)] += 1e-2; A /= A.sum(axis=1, keepdims=True)
        pi = rng.random(K) + 1e-3; pi /= pi.sum()
        logF = rng.normal(size=(T, K)) * float(rng.choice([0.5, 3.0]))
        if K > 1 and rng.random() < 0.4:
            logF[int(rng.integers(T)), int(rng.integers(K))] = -np.inf
        hf = None
        if T > 1 and rng.random() < 0.3:
            hf = (rng.random(T) < 0.6).astype(np.uint8); hf[-1] = 1
        x0 = int(rng.integers(K)) if rng.random() < 0.3 else -1
        desc = dict(case=c, K=K, T=T, x0=x0, hf=None if hf is None else hf.tolist())
        try:
            with np.errstate(all="ignore"):
                Z, probs = er.hmm_enumerate_joint(A, logF, x0=x0, pi=pi, has_factor=hf)
            if not Z > 0.0:                                          # every path has zero mass: Discrete(dict) raises (src/erp.jl:48)
                raise ZeroDivisionError("zero mass")
        except Exception as e:                                       # zero mass
            try:
                eng.load(sb.DiscreteHMM(A, logF, x0=x0, pi=pi, has_factor=hf)); eng.enum_hmm()
                bad += 1; print("MISMATCH", desc, "oracle raised %r, GPU did not" % str(e), flush=True)
            except sb.StochyError:
                pass
            continue
        try:
            eng.load(sb.DiscreteHMM(A, logF, x0=x0, pi=pi, has_factor=hf))
            joint, marg, lz = eng.enum_hmm()
            j = joint / joint.sum()
            codes = np.arange(K ** T)
            m_ref = np.zeros((T, K))
            for t in range(T):
                np.add.at(m_ref[t], (codes // K ** t) % K, probs)
            ok = np.allclose(j, probs, rtol=1e-10, atol=1e-13) and abs(lz - np.log(Z)) <= 1e-10 * max(1.0, abs(np.log(Z))) \
                and np.allclose(marg, m_ref, rtol=1e-9, atol=1e-12)
            if not ok:
                desc["where"] = "joint %g marg %g logZ %g vs %g" % (np.abs(j - probs).max(), np.abs(marg - m_ref).max(), lz, np.log(Z))
        except sb.StochyError as e:
            ok = False; desc["where"] = "GPU raised %r" % str(e)
        if not ok:
            bad += 1; print("MISMATCH", desc, flush=True)
print("fuzz enum: %d cases, %d mismatches" % (cases, bad), flush=True)
sys.exit(1 if bad else 0)
