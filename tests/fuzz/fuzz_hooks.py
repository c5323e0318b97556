#!/usr/bin/env python
"""Randomised parity hunt for the resampling hooks (quantise -> tile scan -> pull / search): random pool sizes and child
counts around tile boundaries, weight kinds (linear fp64, log fp64, log fp32), dynamic ranges from flat to 10^300, exact
zeros / -inf, a single survivor, offsets u at 0, just below 1 and random -- GPU against the oracle, bit for bit.  Test
tooling (uses the oracle), run by hand: python tests/fuzz/fuzz_hooks.py [cases] [seed]."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import stochy_jl_b200 as sb
from oracle import oracle as orc
orc.build()
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
sizes = [1, 2, 3, 31, 255, 256, 257, 2047, 2048, 2049, 4096, 6000, 16385, 100003, 300000, 1500000, 3000001]
bad = 0
with sb.Engine(precision="fp64") as eng:
    for c in range(cases):
        n = int(rng.choice(sizes)); m = int(rng.choice(sizes)) if rng.random() < 0.4 else n
        kind = int(rng.integers(3))
        spread = float(rng.choice([0.0, 1.0, 20.0, 300.0]))
        lw = rng.normal(size=n) * spread + float(rng.choice([0.0, -500.0, 300.0]))
        if kind == 2:
            lw = np.clip(lw, -80.0, 80.0)
        zero = rng.random(n) < float(rng.choice([0.0, 0.3, 0.999]))
        if zero.all():
            zero[int(rng.integers(n))] = False
        if kind == 0:
            w = np.exp(np.clip(lw, -700, 700)); w[zero] = 0.0
        else:
            w = lw.copy(); w[zero] = -np.inf
        u = float(rng.choice([0.0, 0.9999999999, rng.random()]))
        desc = dict(case=c, n=n, m=m, kind=kind, spread=spread, u=u, zeros=int(zero.sum()))
        try:
            ref = orc.resample_systematic(w, u, m, kind=kind)
            got = eng.resample_systematic(w, u, m, kind=kind)
            ok = np.array_equal(got, ref)
            if ok and n <= 300000 and m <= 300000:
                r = rng.random(m)
                ok = np.array_equal(eng.resample_multinomial(w, r, kind=kind), orc.resample_multinomial(w, r, kind=kind))
                if not ok: desc["where"] = "multinomial"
            elif not ok:
                desc["where"] = "systematic: %d of %d ancestors differ" % (int((got != ref).sum()), m)
        except (orc.OracleError, sb.StochyError) as e:
            ok = False; desc["where"] = "raised %r" % str(e)
        if not ok:
            bad += 1
            print("MISMATCH", desc, flush=True)
print("fuzz hooks: %d cases, %d mismatches" % (cases, bad), flush=True)
sys.exit(1 if bad else 0)
