#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the SMC sweep (BASELINE.json metric) on N B200s.

A "step" is one full SMC sweep of the bench workload: a discrete HMM (SURVEY §8(d) cfg3 model:
K=16 sticky states, 16 symbols, T=1000 observations simulated with default_rng(12345)) at
2^24 particles per GPU, systematic resampling at every factor statement, fp32 weights, history off.
One sweep = N*T = 1.68e10 particle-steps per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...                           the reference's CPU algorithm (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is derived.
"""
import argparse
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
K_STATES = 16
BYTES_PER_PARTICLE_STEP = 16        # fused step kernel: read w32 + read parent state + write state + write w32
SURVEY_BYTES = 24                   # SURVEY §8(d) unfused accounting 2S+2W+2A (reported beside it)


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


# ----------------------------------------------------------------------------- CPU arms (oracle port)
def cpu_sample(budget_log2n=22, T=16):
    """Bounded sample of the same workload on the host cores with the oracle port: same model,
    same resampling scheme, N = 2^22 particles x the first 16 steps (about 10-30 s of one core)."""
    from oracle import oracle as orc
    A, logF, pi = workload_tables(T=T)
    m = orc.Hmm(A, logF, pi=pi)
    o = orc.opts(orc.SYSTEMATIC, fp32=True, seed=1)
    N = 1 << budget_log2n
    cores = orc.lib().or_num_threads()
    t0 = time.perf_counter()
    orc.hmm_sweep(m, o, N, it=0, want_hist=False)
    dt = time.perf_counter() - t0
    return N * T / dt, cores, dt, "discrete HMM K=16, N=2^%d particles x first %d steps, oracle port (C, OpenMP)" % (budget_log2n, T)


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
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the same workload as our arm's line (same keys), plus what the CPU arm actually runs per bench step
            "config": {"workload": "discrete HMM K=%d T=%d, 2^%d particles per GPU, one SMC sweep per step "
                                   "(propagate + weight + normalise + systematic resample + gather every t), "
                                   "history off" % (K_STATES, args.T, args.log2n),
                       "particles_per_gpu": 1 << args.log2n, "T": args.T, "K": K_STATES, "resample": "systematic",
                       "reference_arm": "C restatement of Stochy.jl src/pmcmc.jl + src/rand.jl on the host cores (Julia 0.3, "
                                        "REQUIRE:1, cannot run here); each bench step is a bounded sample of the workload: "
                                        "N=2^21 particles x the first 8 steps"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import stochy_jl_b200 as sb
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = 1 << args.log2n
    T = args.T
    A, logF, pi = workload_tables(T=T)
    model = sb.DiscreteHMM(A, logF, pi=pi)
    # N > 1: ONE global pool of world x 2^24 particles sharded over the GPUs (weak scaling), resampled
    # globally at every step; every rank uses the same seed (the Philox counters are global particle ids)
    eng = sb.Engine(device=local, precision="fp32", resample="systematic", seed=1)
    eng.load(model)
    if world > 1 or args.force_shard:
        eng.shard(N)                                # --force-shard: the sharded code path on one GPU (diagnostics)
    eng.set_profile(max(1, T // 100))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        eng.smc_sweep(N, it=i)
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
    # device time of exactly K sweeps; max over ranks
    tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    dev_ms = float(tt.item())
    value = world * N * T * args.steps / (dev_ms * 1e-3)

    # end to end through the public API with host buffers: model tables H2D, sweep, result D2H
    eng.set_profile(0)
    h2d = model.A.nbytes + model.logF.nbytes * 2 + model.pi.nbytes + T
    d2h = 8 + N * 4
    xpin = torch.empty(N, dtype=torch.int32).pin_memory().numpy()     # page-locked host buffer for the result
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2 = eng                                       # same handle: buffers are already sized
        e2.load(model)                                 # host -> device copy of this step's inputs
        e2.smc_sweep(N, it=1000 + i)                   # log-evidence comes back to the host
        x = e2.particles(N, out=xpin)                  # device -> host read of the final particle set
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * N * T * args.steps / float(te.item())

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = float(np.mean(kern_ms))
        ach = N * BYTES_PER_PARTICLE_STEP / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "discrete HMM K=%d T=%d, 2^%d particles per GPU, one SMC sweep per step "
                                       "(propagate + weight + normalise + systematic resample + gather every t), "
                                       "history off" % (K_STATES, T, args.log2n),
                           "particles_per_gpu": N, "T": T, "K": K_STATES, "resample": "systematic",
                           "l2": "per-step working set 4 x N x 4 B = %d MB > 126 MB L2 (inputs larger than L2)" % (16 * N >> 20),
                           "parallelism": "1 GPU" if world == 1 else
                           "one global pool of %d x 2^%d particles sharded over %d GPUs, global systematic resampling, "
                           "parents read across GPUs by peer loads over NVLink (no NCCL on the data path)" % (world, args.log2n, world)},
                "wall_ms_per_step": wall * 1e3 / args.steps,
                "gpu_launches": int(launches),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": {"bound": "hbm", "kernel": "k_step_hmm<PULL,fp32>", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": (ach / peak) if ach else None, "traffic": None,
                             "bytes_per_particle_step": BYTES_PER_PARTICLE_STEP,
                             "kernel_ms_per_launch": k_ms, "peak_source": peak_src,
                             "achieved_survey_24B": (ach * SURVEY_BYTES / BYTES_PER_PARTICLE_STEP) if ach else None}}
        if world == 1:
            v, cores, dt, sample = cpu_sample()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                line["roofline"]["traffic"] = json.load(open(tr)).get("k_step_hmm_bytes_per_launch")
            except Exception:
                pass
        emit(line)
    barrier()                                       # no rank unmaps its arena while a peer may still read it
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2_N)
    ap.add_argument("--T", type=int, default=T_STEPS)
    ap.add_argument("--force-shard", action="store_true", help="run the sharded (peer-addressed) kernels even on one GPU")
    args = ap.parse_args()
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
