#!/usr/bin/env python
"""Randomised parity hunt for batched single-site MH (src/mh.jl:47-111): random binary Bayes nets (1..12 sites, up to
four parents, 0..3 factor statements, hard factors, probabilities at 0 and 1), query sets of 1..4 sites (one site: the
per-chain counter path; several: the warp-aggregated histogram), chain counts that are not multiples of a warp or a CTA,
with and without the per-(site, state) tables -- counts and accept totals against the oracle, bit for bit.
Test tooling (uses the oracle): python tests/fuzz/fuzz_mh.py [cases] [seed]."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import stochy_jl_b200 as sb
from oracle import oracle as orc
orc.build()
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for c in range(cases):
    D = int(rng.integers(1, 13))
    sites, factors = [], []
    for d in range(D):
        npar = min(d, int(rng.integers(0, 5)))
        par = rng.choice(d, size=npar, replace=False) if npar else []
        mask = int(sum(1 << int(p) for p in par))
        p = rng.uniform(0.02, 0.98, size=1 << npar)
        if rng.random() < 0.15:
            p[int(rng.integers(len(p)))] = float(rng.choice([0.0, 1.0]))
        sites.append((mask, p.tolist()))
    for f in range(int(rng.integers(0, 4))):
        npar = min(D, int(rng.integers(1, 4)))
        par = rng.choice(D, size=npar, replace=False)
        mask = int(sum(1 << int(p) for p in par))
        tab = rng.normal(size=1 << npar) * float(rng.choice([0.5, 3.0]))
        if rng.random() < 0.3 and len(tab) > 1:
            tab[int(rng.integers(len(tab)))] = -np.inf
        factors.append((mask, tab.tolist()))
    query = sorted(set(int(q) for q in rng.choice(D, size=min(D, int(rng.integers(1, 5))), replace=False)))
    chains = int(rng.choice([1, 31, 33, 100, 127, 129, 1000, 5000]))
    steps = int(rng.choice([0, 1, 17, 150]))
    chain0 = int(rng.choice([0, 3, 1 << 20]))
    seed = int(rng.integers(1, 1 << 30))
    desc = dict(case=c, D=D, F=len(factors), query=query, chains=chains, steps=steps, chain0=chain0, seed=seed)
    try:
        net = sb.BayesNet(sites, factors, query)
        ob = orc.Bnet([m for m, _ in sites], [t for _, t in sites], [m for m, _ in factors], [t for _, t in factors], net.query_mask)
        ref, racc = orc.bnet_mh(ob, seed, chains, steps, chain0=chain0)
        ok = True
        for notab in [None, "1"]:
            if notab: os.environ["SB_MH_NOTAB"] = notab
            else: os.environ.pop("SB_MH_NOTAB", None)
            with sb.Engine(precision="fp64", seed=seed) as eng:
                eng.load(net)
                got = eng.mh_counts(chains, steps, chain0=chain0)
                if not (np.array_equal(got, ref) and eng.stats().accepts == racc):
                    ok = False; desc["where"] = "SB_MH_NOTAB=%s got %s ref %s accepts %d vs %d" % (notab, got[:4], ref[:4], eng.stats().accepts, racc)
                    break
    except (orc.OracleError, sb.StochyError) as e:
        ok = False; desc["where"] = "raised %r" % str(e)
    if not ok:
        bad += 1
        print("MISMATCH", desc, flush=True)
print("fuzz mh: %d cases, %d mismatches" % (cases, bad), flush=True)
sys.exit(1 if bad else 0)
