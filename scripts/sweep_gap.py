import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", ".")); sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "tests"))
import stochy_jl_b200 as sb
from gpu_util import cfg3_tables
for T in [250, 1000]:
    A, logF, pi = cfg3_tables(T=T)
    with sb.Engine(precision="fp32", resample="systematic", seed=1) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        for i in range(3): eng.smc_sweep(1 << 24, it=i)
        dev = []
        eng.timer_start(); t0 = time.perf_counter()
        for i in range(5):
            eng.smc_sweep(1 << 24, it=3 + i); dev.append(eng.stats().device_ms)
        tot = eng.timer_stop(); wall = time.perf_counter() - t0
    print("T=%d: per-sweep device ms %s  | timer over 5 sweeps / 5 = %.3f ms, wall/5 = %.3f ms, gap per sweep = %.3f ms" % (
        T, ["%.3f" % d for d in dev], tot / 5, wall * 1e3 / 5, tot / 5 - sum(dev) / 5))
