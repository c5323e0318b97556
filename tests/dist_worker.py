"""Multi-GPU parity worker (run under torchrun, one rank per GPU; not collected by pytest).

Every rank shards one global pool of world x N_local particles, runs the global SMC sweep, and
rank 0 checks that (a) the log-evidence is identical on every rank and (b) the concatenated final
particle sets and the log-evidence are BIT-IDENTICAL to a single-GPU sweep of the same global pool
(same seed): the result must not depend on how many GPUs the pool is spread over."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import stochy_jl_b200 as sb
    from gpu_util import cfg3_tables, cfg4_data
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # SB_DIST_SAME_GPU=1: every rank on device 0 (a one-GPU box): the ranks are separate processes whose arenas
    # are mapped into each other over CUDA IPC exactly as across GPUs; the host plumbing is gloo (NCCL refuses two
    # ranks on one device).  The kernels of the ranks are time-sliced, the flag barrier still holds.
    same = os.environ.get("SB_DIST_SAME_GPU") == "1"
    if same:
        local = 0
    torch.cuda.set_device(local)
    if same:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdev = "cpu" if same else "cuda"
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    NL = 1 << log2n
    fails = []
    cases = []
    A, logF, pi = cfg3_tables(T=25)
    cases.append(("hmm fp32", lambda: sb.DiscreteHMM(A, logF, pi=pi), "fp32", np.int32))
    cases.append(("hmm fp64", lambda: sb.DiscreteHMM(A, logF, pi=pi), "fp64", np.int32))
    d = cfg4_data(T=20)
    cases.append(("lg fp32", lambda: sb.LinearGaussian(**d), "fp32", np.float64))
    cases.append(("lg fp64", lambda: sb.LinearGaussian(**d), "fp64", np.float64))
    # an op-list program (Beta draw, then kept; Bernoulli observe statements with several observations; a Normal random walk
    # at the end so that later steps draw too): the peer-addressed instantiation of the op-list step kernel
    ys = (np.random.default_rng(3).random((6, 3)) < 0.35).astype(float)
    prog = [((("beta", 2.0, 3.0) if t == 0 else ("keep",)), ("bernoulli", 1.0, 0.0, list(ys[t]))) for t in range(6)]
    prog += [(("normal", 1.0, 0.0, 0.05), ("normal", 1.0, 0.0, 0.5, [0.3, 0.4])) for _ in range(5)]
    cases.append(("ops fp64", lambda: sb.OpsProgram(prog), "fp64", np.float64))
    # SB_DIST_FUZZ=n: n more random discrete models (same on every rank: seeded), K from 1 to 64 (head markers that carry the
    # state in 4 or 6 bits, or none), wide score ranges, zeros in the transition rows, odd and even T
    frng = np.random.default_rng(int(os.environ.get("SB_DIST_FUZZ_SEED", "1")))
    for i in range(int(os.environ.get("SB_DIST_FUZZ", "0"))):
        K = int(frng.choice([1, 2, 5, 16, 17, 40, 64])); T = int(frng.integers(1, 12))
        Af = frng.random((K, K)) ** int(frng.choice([1, 4])); Af[frng.random((K, K)) < 0.3] = 0.0
        Af[np.arange(K), np.arange(K)] += 1e-3; Af /= Af.sum(axis=1, keepdims=True)
        pif = frng.random(K) + 1e-3; pif /= pif.sum()
        lf = frng.normal(size=(T, K)) * float(frng.choice([0.5, 5.0, 40.0]))
        prec = str(frng.choice(["fp32", "fp64"]))
        cases.append(("fuzz%d K=%d T=%d %s" % (i, K, T, prec), (lambda Af=Af, lf=lf, pif=pif: sb.DiscreteHMM(Af, lf, pi=pif)), prec, np.int32))
    for name, mk, prec, dt in cases:
        eng = sb.Engine(device=local, precision=prec, resample="systematic", seed=11)
        eng.load(mk())
        eng.shard(NL)
        nsw = int(os.environ.get("SB_DIST_SWEEPS", "2"))
        lz = [eng.smc_sweep(NL, it=i) for i in range(nsw)]          # several sweeps, no barrier between them: the flag epochs keep
                                                                    # counting and (odd T) the starting pool slot alternates
        x = torch.from_numpy(eng.particles(NL).astype(np.float64)).to(tdev)
        allx = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(allx, x)
        lzt = torch.tensor(lz, dtype=torch.float64, device=tdev)
        alllz = [torch.empty_like(lzt) for _ in range(world)]
        dist.all_gather(alllz, lzt)
        dist.barrier()
        eng.close()
        if rank == 0:
            for r in range(world):
                if not torch.equal(alllz[r], alllz[0]):
                    fails.append("%s: logZ differs between rank 0 and rank %d" % (name, r))
            ref = sb.Engine(device=local, precision=prec, resample="systematic", seed=11)
            ref.load(mk())
            lz_ref = [ref.smc_sweep(NL * world, it=i) for i in range(nsw)]
            x_ref = ref.particles(NL * world).astype(np.float64)
            ref.close()
            got = torch.cat(allx).cpu().numpy()
            if lz_ref != lz:
                fails.append("%s: logZ %r != single-GPU %r" % (name, lz, lz_ref))
            if not np.array_equal(got, x_ref):
                fails.append("%s: %d of %d final particles differ from the single-GPU sweep"
                             % (name, int((got != x_ref).sum()), got.size))
            print("[dist] %-22s world=%d N_local=2^%d logZ=%.6f %s" % (name, world, log2n, lz[-1],
                                                                      "ok" if not fails else "FAIL"), flush=True)
        dist.barrier()
    if rank == 0:
        print("DIST_PARITY_OK" if not fails else "DIST_PARITY_FAIL\n" + "\n".join(fails), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
