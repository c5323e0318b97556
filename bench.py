#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the SMC sweep (BASELINE.json metric) on N B200s.

A "step" is one full SMC sweep of the bench workload: a discrete HMM (SURVEY §8(d) cfg3 model:
K=16 sticky states, 16 symbols, T=1000 observations simulated with default_rng(12345)) at
2^24 particles per GPU, systematic resampling at every factor statement, fp32 weights, history off.
One sweep = N*T = 1.68e10 particle-steps per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --workload lg [--precision fp64]               BASELINE cfg4 (linear-Gaussian, T=500); never the default
  python bench.py --impl reference ...                           the reference's CPU algorithm (oracle port)

Prints ONE JSON line (rank 0).  Besides the contract's keys the line carries correctness tokens the driver can
see: `check` (log-evidence and a 64-bit hash of the final particle set of sweep it=0), for N > 1
`parity_vs_single_gpu` (a reduced pool swept once sharded over the N GPUs and once on rank 0's GPU alone:
log-evidence and particle hash must be equal bit for bit) and `strong` (2^24 particles in total over the N GPUs),
and `mh_replicas` (BASELINE cfg5: 65536 chains x 10^4 steps per GPU, one chain set per GPU, no communication).
See DESIGN.md "Measurement" for how each field is derived.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
LOG2_N = 24
T_STEPS = 1000
T_LG = 500
K_STATES = 16
# fused step kernel, canonical layout of SURVEY §8 (S = W = 4 B): read w32 + read parent state + write state + write w32.
# The ancestor vector is never materialised, so SURVEY's unfused 24 B (2S+2W+2A) is reported beside it, not claimed.
BYTES_HMM = 16
BYTES_LG = {"fp32": 16, "fp64": 24}
SURVEY_BYTES = 24
RESAMPLE_GATHER_BYTES = 20          # SURVEY §8(d) "resample/gather path alone": W + 2A + 2S, every byte really moved


def workload_tables(T=T_STEPS):
    from gpu_util import cfg3_tables
    return cfg3_tables(T=T, K=K_STATES)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def digest64(a):
    """64-bit hash of an array's bytes (BLAKE2b-64, hex)."""
    return hashlib.blake2b(np.ascontiguousarray(a).tobytes(), digest_size=8).hexdigest()


def combine_digests(ds):
    """Hash of a sharded particle set = hash of the per-shard hashes in rank order."""
    return hashlib.blake2b("".join(ds).encode(), digest_size=8).hexdigest()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload_text(args):
    if args.workload == "lg":
        return ("linear-Gaussian state space (BASELINE cfg4: x0~N(0,1), x_t = 0.9 x_{t-1} + N(0,1), y_t = x_t + N(0,1)) "
                "T=%d, 2^%d particles per GPU, one plain SMC sweep per step (propagate + weight + normalise + systematic "
                "resample + gather every t), %s state and weights, history off" % (args.T, args.log2n, args.precision))
    return ("discrete HMM K=%d T=%d, 2^%d particles per GPU, one SMC sweep per step "
            "(propagate + weight + normalise + systematic resample + gather every t), "
            "history off" % (K_STATES, args.T, args.log2n))


# ----------------------------------------------------------------------------- CPU arms (oracle port)
def cpu_sample(budget_log2n=22, T=16):
    """Bounded sample of the same workload on the host cores with the oracle port: same model, same
    (systematic, weight-grid) resampling scheme and the same fp32 log-weights as the GPU arm,
    N = 2^22 particles x the first 16 steps (about 10-30 s of one core)."""
    from oracle import oracle as orc
    A, logF, pi = workload_tables(T=T)
    m = orc.Hmm(A, logF, pi=pi)
    o = orc.opts(orc.SYSTEMATIC, fp32=True, seed=1)
    N = 1 << budget_log2n
    cores = orc.lib().or_num_threads()
    t0 = time.perf_counter()
    orc.hmm_sweep(m, o, N, it=0, want_hist=False)
    dt = time.perf_counter() - t0
    return N * T / dt, cores, dt, ("discrete HMM K=16, N=2^%d particles x first %d steps, oracle port (C, OpenMP), systematic "
                                   "resampling on the weight grid, fp32 log-weights" % (budget_log2n, T))


def literal_extrapolation(n_probe=1 << 12):
    """The reference's OWN resampler (src/pmcmc.jl:51: N draws, each an O(N) scan of src/rand.jl:2-13) timed on one
    thread at a size it finishes, and its N^2 law carried to one resampling step of the bench pool."""
    from oracle import oracle as orc
    rng = np.random.default_rng(7)
    w = rng.random(n_probe)
    r = rng.random(n_probe)
    orc.resample_literal(w, r)
    t0 = time.perf_counter()
    orc.resample_literal(w, r)
    dt = time.perf_counter() - t0
    c = dt / (float(n_probe) ** 2)
    return {"probe_n": n_probe, "probe_s": dt, "s_per_draw_per_entry": c,
            "one_resampling_step_at_2^24_s": c * float(1 << 24) ** 2,
            "note": "literal reference resampler is O(N^2) per step; extrapolated, not run at the bench size"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is rank 0 alone and may use every host
    # core (libgomp reads the variable when the oracle library is loaded, which has not happened yet)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, dt, sample = cpu_sample(budget_log2n=21, T=8)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32",          # what this arm computes in: fp32 log-weights quantised onto the integer grid (as the GPU arm)
            "data": "synthetic",
            # the same workload as our arm's line (same keys), plus what the CPU arm actually runs per bench step
            "config": {"workload": workload_text(args),
                       "particles_per_gpu": 1 << args.log2n, "T": args.T, "K": K_STATES, "resample": "systematic",
                       "same_config": False,
                       "reference_arm": "C restatement of Stochy.jl src/pmcmc.jl + src/rand.jl on the host cores (Julia 0.3, "
                                        "REQUIRE:1, cannot run here).  It runs the SAME systematic weight-grid scheme as the GPU arm "
                                        "(a faster algorithm than the reference's literal O(N^2) multinomial resampler, whose "
                                        "extrapolated cost is in `literal_resampler`), and each bench step is a bounded sample of "
                                        "the workload: N=2^21 particles x the first 8 steps"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "literal_resampler": literal_extrapolation(),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- our arm
def make_model(sb, args):
    if args.workload == "lg":
        from gpu_util import cfg4_data
        return sb.LinearGaussian(**cfg4_data(T=args.T))
    A, logF, pi = workload_tables(T=args.T)
    return sb.DiscreteHMM(A, logF, pi=pi)


def exact_log_evidence(sb, args, local):
    """Closed-form log-evidence of the bench model, which log Z-hat estimates: the product's own Enumerate on the GPU
    (scaled forward-backward, sb_enum_hmm) for the discrete model; a scalar Kalman filter (below) for cfg4."""
    if args.workload == "lg":
        from gpu_util import cfg4_data
        d = cfg4_data(T=args.T)
        m, P, ll = d["m0"], d["P0"], 0.0
        for t, y in enumerate(d["y"]):
            if t:
                m, P = d["a"] * m, d["a"] * d["a"] * P + d["q"]
            S = d["c"] * d["c"] * P + d["r"]
            ll += -0.5 * (np.log(2 * np.pi * S) + (y - d["c"] * m) ** 2 / S)
            G = P * d["c"] / S
            m, P = m + G * (y - d["c"] * m), (1 - G * d["c"]) * P
        return float(ll)
    with sb.Engine(device=local, precision="fp64") as e:
        e.load(make_model(sb, args))
        return float(e.enum_hmm(want_joint=False)[2])


def run_ours(args):
    import torch
    import stochy_jl_b200 as sb
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = 1 << args.log2n
    T = args.T
    model = make_model(sb, args)
    dt_x = np.int32 if args.workload == "hmm" else np.float64
    # N > 1: ONE global pool of world x 2^24 particles sharded over the GPUs (weak scaling), resampled
    # globally at every step; every rank uses the same seed (the Philox counters are global particle ids)
    eng = sb.Engine(device=local, precision=args.precision, resample="systematic", seed=1)
    eng.load(model)
    if world > 1 or args.force_shard:
        eng.shard(N)                                # --force-shard: the sharded code path on one GPU (diagnostics)
    # live kernel timing inside the timed region: a CUDA-event pair around every 50th launch of the step kernel (20 samples per
    # sweep at T = 1000).  An event between the scan and the step kernel costs that step its programmatic overlap (~3 us), so the
    # sampled steps are kept few: every 10th step made the whole sweep 1.2 % slower than the unprofiled one (76.0 against 75.1 ms)
    eng.set_profile(max(1, T // 20))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_strings(s, n=16):
        if world == 1:
            return [s]
        t = torch.tensor(list(s.encode().ljust(n)), dtype=torch.uint8, device="cuda")
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [bytes(o.cpu().tolist()).decode().strip() for o in out]

    check = None
    for i in range(max(args.warmup, 1)):
        lz = eng.smc_sweep(N, it=i)
        if i == 0:                                  # correctness token of sweep it=0 (outside the timed region)
            ds = gather_strings(digest64(eng.particles(N)))
            check = {"it": 0, "logZ": lz, "particles_blake2b64": combine_digests(ds) if world > 1 else ds[0]}
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches, kern_ms, dev_ms = 0, [], 0.0
    t0 = time.perf_counter()
    eng.timer_start()                               # CUDA events on the library's own (launching) stream
    for i in range(args.steps):
        eng.smc_sweep(N, it=args.warmup + i)
        st = eng.stats()
        launches += st.kernel_launches
        kern_ms.append(st.step_kernel_ms)
    dev_ms = eng.timer_stop()                       # exactly K sweeps, including the gaps between them
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = all_max(dev_ms)                        # device time of exactly K sweeps; max over ranks
    value = world * N * T * args.steps / (dev_ms * 1e-3)

    # end to end through the public API with host buffers: model tables H2D, sweep, result D2H
    eng.set_profile(0)
    if args.workload == "hmm":
        h2d = model.A.nbytes + model.logF.nbytes * 2 + model.pi.nbytes + T
    else:
        h2d = model.y.nbytes + 6 * 8
    d2h = 8 + N * np.dtype(dt_x).itemsize
    tdt = torch.int32 if args.workload == "hmm" else torch.float64
    xpin = torch.empty(N, dtype=tdt).pin_memory().numpy()     # page-locked host buffer for the result
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.load(model)                               # host -> device copy of this step's inputs
        eng.smc_sweep(N, it=1000 + i)                 # log-evidence comes back to the host
        eng.particles(N, out=xpin)                    # device -> host read of the final particle set
    barrier()
    e2e_value = world * N * T * args.steps / all_max(time.perf_counter() - t0)

    extras = {}
    # ---- N > 1: the sharded sweep must equal the single-GPU sweep of the same global pool, bit for bit
    if world > 1:
        extras["parity_vs_single_gpu"] = parity_vs_single_gpu(sb, torch, dist, args, rank, world, local, gather_strings)
        extras["strong"] = strong_scaling(sb, torch, dist, args, rank, world, local, all_max)
        if args.workload == "hmm":                  # BASELINE cfg4 as written: linear-Gaussian, 2^24 particles IN TOTAL, Kalman-checked
            a4 = argparse.Namespace(**vars(args))
            a4.workload, a4.T, a4.precision = "lg", 500, "fp32"
            blk = strong_scaling(sb, torch, dist, a4, rank, world, local, all_max, sweeps=2)
            blk["workload"] = "linear-Gaussian state space (BASELINE cfg4), T=500, fp32, 2^24 particles in total over the GPUs"
            if rank == 0:
                blk["logZ_kalman"] = exact_log_evidence(sb, a4, local)
                blk["abs_err_vs_kalman"] = abs(blk["logZ_last"] - blk["logZ_kalman"])
            extras["strong_cfg4_lg"] = blk
    extras["mh_replicas"] = mh_replicas(sb, torch, dist, rank, world, local, all_max)
    if world > 1 and args.workload == "hmm":
        extras["pmcmc_sharded"] = pmcmc_sharded(sb, torch, dist, rank, world, local)
    if world == 1 and rank == 0 and args.workload == "hmm" and not args.force_shard:
        extras["cfg3_csmc"] = cfg3_csmc(sb, local)
        extras["cfg4_lg"] = cfg4_lg(sb, args, local)
        extras["ops_program"] = ops_program(sb, local)

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = float(np.mean(kern_ms))
        bpp = BYTES_HMM if args.workload == "hmm" else BYTES_LG[args.precision]
        ach = N * bpp / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        kname = ("k_step_hmm<PULL,%s>" if args.workload == "hmm" else "k_step_lg<PULL,%s>") % args.precision
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
                "config": {"workload": workload_text(args),
                           "particles_per_gpu": N, "T": T, "K": K_STATES if args.workload == "hmm" else None, "resample": "systematic",
                           "l2": "per-step working set %d MB > 126 MB L2 (inputs larger than L2)" % (bpp * N >> 20),
                           "parallelism": "1 GPU" if world == 1 else
                           "one global pool of %d x 2^%d particles sharded over %d GPUs, global systematic resampling, "
                           "parents read across GPUs by peer loads over NVLink (no NCCL on the data path)" % (world, args.log2n, world)},
                "wall_ms_per_step": wall * 1e3 / args.steps,
                "gpu_launches": int(launches),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": (ach / peak) if ach else None, "traffic": None,
                             "bytes_per_particle_step": bpp,
                             "kernel_ms_per_launch": k_ms, "peak_source": peak_src,
                             "achieved_survey_24B": (ach * SURVEY_BYTES / bpp) if (ach and args.workload == "hmm") else None},
                "check": check}
        try:
            exact = exact_log_evidence(sb, args, local)
            check["logZ_exact"] = exact                 # forward algorithm (GPU Enumerate) / Kalman filter, fp64
            check["abs_err"] = abs(check["logZ"] - exact)
        except Exception as ex:
            check["logZ_exact"] = None
            check["note"] = "closed form unavailable: %s" % ex
        line.update(extras)
        if world == 1:
            if args.workload == "hmm":
                line["roofline_resample_gather"] = resample_gather_path(sb, torch, eng, N, peak)
            v, cores, dt, sample = cpu_sample()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        else:
            line["cpu_baseline"] = None                  # timed on rank 0 at N = 1 only (see the N = 1 line)
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                key = "k_step_hmm_bytes_per_launch" if args.workload == "hmm" else "k_step_lg_%s_bytes_per_launch" % args.precision
                line["roofline"]["traffic"] = json.load(open(tr)).get(key)
            except Exception:
                pass
        emit(line)
    barrier()                                       # no rank unmaps its arena while a peer may still read it
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def parity_vs_single_gpu(sb, torch, dist, args, rank, world, local, gather_strings, log2n=21, T=8):
    """A reduced global pool (world x 2^21 particles, T = 8) swept sharded over all GPUs and, on rank 0's GPU alone,
    unsharded: log-evidence and the hash of the particle set must agree bit for bit.  Outside every timed region."""
    NL = 1 << log2n
    a2 = argparse.Namespace(**vars(args))
    a2.T = T
    m = make_model(sb, a2)
    e = sb.Engine(device=local, precision=args.precision, resample="systematic", seed=5)
    e.load(m)
    e.shard(NL)
    lz = e.smc_sweep(NL, it=0)
    ds = gather_strings(digest64(e.particles(NL)))
    lzs = gather_strings(repr(lz), n=32)
    dist.barrier()
    e.close()
    out = None
    if rank == 0:
        r = sb.Engine(device=local, precision=args.precision, resample="systematic", seed=5)
        r.load(m)
        lz1 = r.smc_sweep(NL * world, it=0)
        x1 = r.particles(NL * world)
        r.close()
        d1 = [digest64(x1[g * NL:(g + 1) * NL]) for g in range(world)]
        out = {"particles_per_gpu": NL, "T": T, "logZ_sharded": lz, "logZ_single_gpu": lz1,
               "hash_sharded": combine_digests(ds), "hash_single_gpu": combine_digests(d1),
               "logZ_equal_on_all_ranks": all(s == lzs[0] for s in lzs),
               "ok": bool(lz == lz1 and ds == d1 and all(s == lzs[0] for s in lzs))}
    dist.barrier()
    return out


def strong_scaling(sb, torch, dist, args, rank, world, local, all_max, log2_total=24, sweeps=3):
    """BASELINE cfg4's shape: a FIXED pool of 2^24 particles in total sharded over the N GPUs."""
    NL = (1 << log2_total) // world
    e = sb.Engine(device=local, precision=args.precision, resample="systematic", seed=1)
    e.load(make_model(sb, args))
    e.shard(NL)
    e.smc_sweep(NL, it=0)
    dist.barrier(); torch.cuda.synchronize()
    e.timer_start()
    for i in range(sweeps):
        lz = e.smc_sweep(NL, it=1 + i)
    ms = all_max(e.timer_stop())
    dist.barrier()
    e.close()
    return {"scaling": "strong", "particles_total": 1 << log2_total, "particles_per_gpu": NL, "sweeps": sweeps,
            "ms_per_sweep": ms / sweeps, "value": (1 << log2_total) * args.T * sweeps / (ms * 1e-3), "unit": UNIT, "logZ_last": lz}


def pmcmc_sharded(sb, torch, dist, rank, world, local, log2n=14, T=16, iters=2):
    """Iterated conditional SMC on ONE pool sharded over the GPUs (src/pmcmc.jl:59-66, 83-94: retained particle in the last
    entry of the last GPU, history rows sharded, backward pass with peer atomics) against sb_pmcmc_run on rank 0's GPU
    alone: per-time counts and retained trajectory must be equal bit for bit.  Outside the timed region."""
    NL = 1 << log2n
    N = NL * world - 1
    A, logF, pi = workload_tables(T)
    eng = sb.Engine(device=local, precision="fp32", resample="systematic", seed=9)
    eng.load(sb.DiscreteHMM(A, logF, pi=pi))
    eng.shard(NL, history_rows=T)
    marg, _ = eng.pmcmc_counts(iters, N)
    xs = eng.retained()
    dist.barrier()
    eng.close()
    out = None
    if rank == 0:
        ref = sb.Engine(device=local, precision="fp32", resample="systematic", seed=9)
        ref.load(sb.DiscreteHMM(A, logF, pi=pi))
        marg_ref, _ = ref.pmcmc_counts(iters, N, want_joint=False)
        xs_ref = ref.retained()
        ref.close()
        out = {"particles": N, "particles_per_gpu": NL, "T": T, "iterations": iters,
               "counts_blake2b64_sharded": digest64(marg), "counts_blake2b64_single_gpu": digest64(marg_ref),
               "counts_total": int(marg.sum()), "retained_equal": bool(np.array_equal(xs, xs_ref)),
               "ok": bool(np.array_equal(marg, marg_ref) and np.array_equal(xs, xs_ref))}
    dist.barrier()
    return out


def cfg3_csmc(sb, local, log2n=20, T=1000, iters=3):
    """BASELINE cfg3 as written: K = 16, T = 1000, N = 2^20, conditional SMC with history (src/pmcmc.jl:78-99): particle-steps
    per second of whole PMCMC iterations (sweep with ancestor rows + the backward pass that counts every particle's value
    and extracts the next retained trajectory), device time.  20 B per particle-step (16 + the ancestor row)."""
    N = 1 << log2n
    A, logF, pi = workload_tables(T)
    with sb.Engine(device=local, precision="fp32", resample="systematic", seed=1) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        eng.pmcmc_counts(1, N, want_joint=False)                       # warm-up (allocations)
        marg, _ = eng.pmcmc_counts(iters, N, want_joint=False)
        st = eng.stats()
    peak, _ = measured_peak()
    us_step = st.device_ms * 1e3 / (iters * T)
    return {"workload": "discrete HMM K=16 T=%d, N=2^%d, %d PMCMC iterations (conditional SMC with history + backward counting)" % (T, log2n, iters),
            "value": iters * N * T / (st.device_ms * 1e-3), "unit": UNIT, "ms_per_iteration": st.device_ms / iters,
            "us_per_step": us_step, "kernel_launches": int(st.kernel_launches),
            "frac_of_hbm_peak_at_20B": N * 20 / (us_step * 1e-6) / 1e9 / peak,
            "counts_total": int(marg.sum()), "counts_expected": iters * N * T, "logZ_first": st.logZ_first}


def cfg4_lg(sb, args, local, sweeps=2):
    """BASELINE cfg4 on one GPU: linear-Gaussian state space, T = 500, 2^24 particles, fp32, plain SMC with systematic
    resampling at every step; log-evidence against the Kalman filter."""
    a4 = argparse.Namespace(**vars(args))
    a4.workload, a4.T, a4.precision = "lg", 500, "fp32"
    N = 1 << 24
    with sb.Engine(device=local, precision="fp32", resample="systematic", seed=1) as eng:
        eng.load(make_model(sb, a4))
        eng.smc_sweep(N, it=0)
        eng.timer_start()
        for i in range(sweeps):
            lz = eng.smc_sweep(N, it=1 + i)
        ms = eng.timer_stop()
    exact = exact_log_evidence(sb, a4, local)
    return {"workload": "linear-Gaussian state space (BASELINE cfg4), T=500, fp32, 2^24 particles", "value": N * 500 * sweeps / (ms * 1e-3),
            "unit": UNIT, "ms_per_sweep": ms / sweeps, "logZ_last": lz, "logZ_kalman": exact, "abs_err_vs_kalman": abs(lz - exact)}


def ops_program(sb, local, log2n=20):
    """An op-list program inside the sweep (sb_model_ops): p = ~Beta(2, 3), then six observe(Bernoulli(p), y1, y2, y3)
    statements -- log-evidence estimate against the conjugate closed form."""
    import math
    a, b = 2.0, 3.0
    ys = (np.random.default_rng(3).random((6, 3)) < 0.35).astype(float)
    steps = [((("beta", a, b) if t == 0 else ("keep",)), ("bernoulli", 1.0, 0.0, list(ys[t]))) for t in range(len(ys))]
    s, f = ys.sum(), ys.size - ys.sum()
    lb = lambda x, y: math.lgamma(x) + math.lgamma(y) - math.lgamma(x + y)
    with sb.Engine(device=local, precision="fp64", seed=1) as eng:
        eng.load(sb.OpsProgram(steps))
        eng.smc_sweep(1 << log2n)
        lz = eng.smc_sweep(1 << log2n, it=1)
        st = eng.stats()
    return {"workload": "p ~ Beta(2,3); 6 x observe(Bernoulli(p), 3 observations), 2^%d particles, fp64" % log2n,
            "logZ": lz, "logZ_exact": lb(a + s, b + f) - lb(a, b), "particle_steps_per_s": (1 << log2n) * len(ys) / (st.device_ms * 1e-3)}


def mh_replicas(sb, torch, dist, rank, world, local, all_max, chains=65536, steps=10000):
    """BASELINE cfg5: batched single-site MH (src/mh.jl) on the Bayes net C,S,R | W, one independent chain set per GPU
    (global chain ids: disjoint Philox streams), no communication on the data path; histograms summed at the end."""
    from gpu_util import cfg5_bnet
    sites, factors, query = cfg5_bnet()
    with sb.Engine(device=local, precision="fp64", seed=1) as e:
        e.load(sb.BayesNet(sites, factors, query))
        e.mh_counts(chains, 10, chain0=rank * chains)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        counts = e.mh_counts(chains, steps, chain0=rank * chains)
        ms = e.stats().device_ms
    t = torch.tensor([float(counts[0]), float(counts[1])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = all_max(ms)
    return {"workload": "batched MH, Bayes net C,S,R | W, %d chains x %d steps per GPU, one chain set per GPU" % (chains, steps),
            "chain_steps_per_s": world * chains * steps / (ms * 1e-3), "unit": "chain-steps/s", "ms": ms,
            "P(R|W)": float(t[1].item() / (t[0].item() + t[1].item())), "exact": 0.7079276773296245}


def resample_gather_path(sb, torch, eng, N, peak, iters=20):
    """north_star's "resample/gather path" on its own (the unfused hooks a caller with explicit log-weights uses):
    quantise -> tile scan -> pull (ancestors written) -> gather, device-resident buffers, 20 algorithmic bytes per
    particle (W + 2A + 2S) -- the ancestor vector IS written and re-read here."""
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    lw = torch.randn(N, device="cuda", dtype=torch.float32, generator=g) * 2.0
    src = torch.randint(0, 1 << 30, (N,), device="cuda", dtype=torch.int32, generator=g)
    anc = torch.empty(N + 2048, device="cuda", dtype=torch.int32)
    dst = torch.empty(N, device="cuda", dtype=torch.int32)
    torch.cuda.synchronize()
    nat = sb._native
    def once(u):
        eng.dev_resample_systematic(lw.data_ptr(), nat.W_LOG_F32, N, u, N, anc.data_ptr())
        eng.dev_gather(src.data_ptr(), 4, N, anc.data_ptr(), N, dst.data_ptr())
    for i in range(3):
        once(0.1 + 0.01 * i)
    eng.sync()
    eng.timer_start()
    for i in range(iters):
        once(0.3 + 0.01 * i)
    ms = eng.timer_stop() / iters
    eng.sync()
    a = anc[:N]
    ok = bool((a[1:] >= a[:-1]).all().item() and int(a.min().item()) >= 0 and int(a.max().item()) < N
              and torch.equal(dst, src[a.long()]))
    ach = N * RESAMPLE_GATHER_BYTES / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernels": "k_quantise<f32 log> + k_tilescan + k_pull + k_gather", "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "bytes_per_particle": RESAMPLE_GATHER_BYTES, "ms_per_resample_gather": ms,
            "particles": N, "ancestors_sorted_in_range_and_gather_exact": ok,
            "l2": "lw, w32, anc, src, dst = 5 x 64 MB > 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hmm", choices=["hmm", "lg"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--log2n", type=int, default=LOG2_N)
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--force-shard", action="store_true", help="run the sharded (peer-addressed) kernels even on one GPU")
    args = ap.parse_args()
    if args.T is None:
        args.T = T_LG if args.workload == "lg" else T_STEPS
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on the first
    # collective when NCCL_DEBUG=VERSION): everything but the result line goes to stderr.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


_RESULT_FD = None


def emit(line):
    """The result line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


if __name__ == "__main__":
    main()
