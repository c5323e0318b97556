#!/usr/bin/env python
"""Race hunt at the bench size: the full sweep (discrete HMM K=16, T steps, 2^24 particles, history off) must give the
same final particle set and log-evidence bit for bit run to run and for any number of resident CTAs per SM (the
persistent kernels then walk the tiles in a different order and their split barriers interleave differently).

  python scripts/stress_determinism.py [T] [log2n]          (one B200; prints one line per run and PASS / FAIL)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import stochy_jl_b200 as sb
    from gpu_util import cfg3_tables
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
    A, logF, pi = cfg3_tables(T=T, K=16)
    outs = []
    for ctas in [None, None, "2", "4", None]:
        if ctas:
            os.environ["SB_CTAS_PER_SM"] = ctas
        else:
            os.environ.pop("SB_CTAS_PER_SM", None)
        with sb.Engine(precision="fp32", resample="systematic", seed=17) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            lz = eng.smc_sweep(N, it=0)
            x = eng.particles(N)
        h = hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()[:16]
        outs.append((lz, h))
        print("CTAs per SM %-7s logZ %.12f  particles sha256 %s" % (ctas or "default", lz, h), flush=True)
    ok = all(o == outs[0] for o in outs)
    print("PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
