#!/usr/bin/env python
"""CPU baselines of SURVEY §8(d), from the oracle port (the reference itself is Julia 0.3, REQUIRE:1, not installable):

  * the LITERAL reference resampler -- N draws, each an O(N) inverse-CDF scan (src/rand.jl:2-13 via src/pmcmc.jl:51) --
    timed single-threaded (the reference is single-threaded: global `ctx`, src/Stochy.jl:31) at N in {100, 2^10, 2^12,
    2^14}, its N^2 law fitted and extrapolated to 2^20 and 2^24;
  * the same ancestors by binary search on the same sequential CDF (or_resample_literal_fast), O(N log N), up to 2^22;
  * cfg1: examples/hmm.jl-sized PMCMC, 100 particles x 1000 iterations, literal numerics, one thread;
  * the systematic-grid SMC sweep of the bench workload (K=16), one thread and all threads, 2^20 particles x 16 steps.

  python tests/cpu_baselines.py > profiles/rNN_cpu_baselines.jsonl        (CPU only; ~1 minute)
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def best(f, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    from oracle import oracle as orc
    from gpu_util import cfg2_tables, cfg3_tables
    orc.build()
    rng = np.random.default_rng(0)
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = lambda d: print(json.dumps(d), flush=True)
    out({"host": cpu, "logical_cpus": os.cpu_count(), "reference": "Julia 0.3 (REQUIRE:1) not installable; timings are the C port (oracle/)"})
    ctypes.CDLL("libgomp.so.1").omp_set_num_threads(1)
    # literal resampler, one thread
    lit = []
    for n in [100, 1 << 10, 1 << 12, 1 << 14]:
        w, r = rng.random(n), rng.random(n)
        t = best(lambda: orc.resample_literal(w, r), reps=5 if n <= 4096 else 2)
        lit.append((n, t))
        out({"what": "literal resample (src/rand.jl:2-13 x N draws), 1 thread", "N": n, "seconds": t, "draws_per_s": n / t})
    c = lit[-1][1] / lit[-1][0] ** 2                       # seconds per (draw x pool entry), from the largest size
    for n in [1 << 20, 1 << 24]:
        out({"what": "literal resample extrapolated (t = c N^2, c from N=2^14)", "N": n, "seconds": c * n * n, "draws_per_s": 1.0 / (c * n)})
    for n in [1 << 14, 1 << 20, 1 << 22]:
        w, r = rng.random(n), rng.random(n)
        t = best(lambda: orc.resample_literal(w, r, fast=True), reps=3)
        out({"what": "same ancestors by binary search on the sequential CDF, 1 thread", "N": n, "seconds": t, "draws_per_s": n / t})
    # cfg1-sized PMCMC with the reference numerics
    A, logF, pi = cfg2_tables()
    m = orc.Hmm(A, logF, pi=pi)
    t = best(lambda: orc.hmm_pmcmc(m, orc.opts(orc.LITERAL, seed=1), 1000, 100), reps=2)
    out({"what": "PMCMC 100 particles x 1000 iterations, T=10 K=3, literal numerics, 1 thread", "seconds": t,
         "particle_steps_per_s": 100 * 1000 * 10 / t})
    # the bench workload's sweep on the grid (what bench.py's cpu_baseline times), 1 thread and all threads
    A, logF, pi = cfg3_tables(T=16, K=16)
    m = orc.Hmm(A, logF, pi=pi)
    N = 1 << 20
    for threads in [1, os.cpu_count()]:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(threads)      # the OpenMP runtime the oracle library is linked to
        t = best(lambda: orc.hmm_sweep(m, orc.opts(orc.SYSTEMATIC, fp32=True, seed=1), N, want_hist=False), reps=2)
        out({"what": "SMC sweep K=16, 2^20 particles x 16 steps, systematic on the grid (oracle port)", "threads_requested": threads,
             "seconds": t, "particle_steps_per_s": N * 16 / t})


if __name__ == "__main__":
    main()
