"""Sharded conditional SMC / PMCMC parity worker (run under torchrun; not collected by pytest).

src/pmcmc.jl:59-66, 83-94 across GPUs: every rank shards one pool of world x N_local entries (N = world x N_local - 1
particles + the retained one, which lives in the last entry of the last GPU), runs `iters` PMCMC iterations, and rank 0
checks that the summed per-time counts, the retained trajectory, the log-evidence of the first and last sweep and the
history rows (states and GLOBAL ancestors) of the last sweep are BIT-IDENTICAL to sb_pmcmc_run on one handle with the
same particle count and seed."""
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
    from gpu_util import cfg3_tables
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    same = os.environ.get("SB_DIST_SAME_GPU") == "1"
    if same:
        local = 0
    torch.cuda.set_device(local)
    if same:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdev = "cpu" if same else "cuda"
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    NL = 1 << log2n
    N = NL * world - 1
    fails = []
    A, logF, pi = cfg3_tables(T=T)
    models = [("cfg3", A, logF, pi)]
    # SB_DIST_FUZZ=n: n more random discrete models (seeded: the same on every rank), K from 1 to 64, wide score ranges
    frng = np.random.default_rng(int(os.environ.get("SB_DIST_FUZZ_SEED", "1")))
    for i in range(int(os.environ.get("SB_DIST_FUZZ", "0"))):
        K = int(frng.choice([1, 2, 5, 16, 17, 40, 64])); Tf = int(frng.integers(1, T + 1))
        Af = frng.random((K, K)) ** int(frng.choice([1, 4])); Af[frng.random((K, K)) < 0.3] = 0.0
        Af[np.arange(K), np.arange(K)] += 1e-3; Af /= Af.sum(axis=1, keepdims=True)
        pif = frng.random(K) + 1e-3; pif /= pif.sum()
        models.append(("fuzz%d K=%d T=%d" % (i, K, Tf), Af, frng.normal(size=(Tf, K)) * float(frng.choice([0.5, 5.0, 40.0])), pif))
    for mname, A, logF, pi in models:
      T = logF.shape[0]
      for prec in (["fp32", "fp64"] if mname == "cfg3" else [str(frng.choice(["fp32", "fp64"]))]):
        for pick in ([-1, 0] if mname == "cfg3" else [int(frng.choice([-1, 0]))]):   # uniform pick (systematic default) / particle 1 as src/pmcmc.jl:89
            eng = sb.Engine(device=local, precision=prec, resample="systematic", seed=5, retained_pick=pick)
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            eng.shard(NL, history_rows=T)
            marg, _ = eng.pmcmc_counts(iters, N)
            st = eng.stats()
            xs = eng.retained()
            rows = []
            for t in range(T):
                x, a, _ = eng.history(t, NL)
                rows.append(np.stack([x, a]))
            mine = torch.from_numpy(np.stack(rows).astype(np.int64)).to(tdev)       # [T, 2, NL]
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            xst = torch.from_numpy(xs.astype(np.int64)).to(tdev)
            allxs = [torch.empty_like(xst) for _ in range(world)]
            dist.all_gather(allxs, xst)
            dist.barrier()
            eng.close()
            if rank == 0:
                ref = sb.Engine(device=local, precision=prec, resample="systematic", seed=5, retained_pick=pick)
                ref.load(sb.DiscreteHMM(A, logF, pi=pi))
                marg_ref, _ = ref.pmcmc_counts(iters, N, want_joint=False)
                st_ref = ref.stats()
                xs_ref = ref.retained()
                name = "pmcmc %s %s pick=%d" % (mname, prec, pick)
                if not np.array_equal(marg, marg_ref):
                    fails.append("%s: counts differ (%d cells)" % (name, int((marg != marg_ref).sum())))
                if marg.sum() != iters * N * T:
                    fails.append("%s: counts sum %d != iters*N*T %d" % (name, marg.sum(), iters * N * T))
                for r in range(world):
                    if not np.array_equal(allxs[r].cpu().numpy(), xs_ref):
                        fails.append("%s: retained trajectory on rank %d differs" % (name, r))
                if (st.logZ_first, st.logZ_last) != (st_ref.logZ_first, st_ref.logZ_last):
                    fails.append("%s: logZ %r != %r" % (name, (st.logZ_first, st.logZ_last), (st_ref.logZ_first, st_ref.logZ_last)))
                got = torch.cat(allr, dim=2).cpu().numpy()            # [T, 2, world*NL]
                for t in range(T):
                    x, a, _ = ref.history(t, N + 1)
                    if not np.array_equal(got[t, 0, :N + 1], x):
                        fails.append("%s: states of step %d differ" % (name, t)); break
                    if not np.array_equal(got[t, 1, :N], a[:N]):
                        fails.append("%s: ancestors of step %d differ" % (name, t)); break
                ref.close()
                print("[dist] %-34s world=%d N_local=2^%d T=%d logZ=%.6f sharded %.2f ms/iteration, one GPU %.2f %s" % (
                    name, world, log2n, T, st.logZ_last, st.device_ms / iters, st_ref.device_ms / iters,
                                                                                 "ok" if not fails else "FAIL"), flush=True)
            dist.barrier()
    if rank == 0:
        print("DIST_PMCMC_OK" if not fails else "DIST_PMCMC_FAIL\n" + "\n".join(fails), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
