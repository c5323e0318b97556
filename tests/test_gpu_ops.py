"""GPU: op-list programs (sb_model_ops: Beta / Gamma / Normal / uniform / flip draws and Bernoulli / Normal / Poisson
observe statements inside the SMC sweep -- src/pmcmc.jl:32-46, src/Stochy.jl:55-56, src/erp.jl:1-30) against the oracle's
restatement step by step, and against closed-form evidence."""
import math

import numpy as np
import pytest

import ops_util as ou
from gpu_util import cfg4_data, need_gpu

pytestmark = pytest.mark.gpu


def sb_mod():
    need_gpu()
    import stochy_jl_b200 as sb
    return sb


def _run(sb, steps, N, seed, it=0, store_logw=True, precision="fp64"):
    with sb.Engine(precision=precision, seed=seed, store_logw=store_logw) as eng:
        eng.load(ou.to_program(sb, steps))
        lz = eng.smc_sweep(N, it=it)
        x = eng.particles(N)
        lw = [eng.logw_row(t, N) for t in range(len(steps))] if store_logw else None
    return lz, x, lw


@pytest.mark.parametrize("prog", ["beta_bernoulli", "gamma_poisson", "mix", "lg", "dirichlet"])
def test_ops_stepwise_parity(oracle, prog):
    """Every step's draw + score equals the oracle's given the SAME parents (libm vs CUDA transcendental functions
    differ in the last ulps: states 1e-11, scores 1e-9 relative); ancestors bit-exact on the GPU's own log-weights
    (the parents the GPU used are recomputed by the oracle's systematic resampler from those weights)."""
    sb = sb_mod()
    steps = {"beta_bernoulli": lambda: ou.beta_bernoulli()[0], "gamma_poisson": lambda: ou.gamma_poisson()[0],
             "mix": ou.uniform_flip_mix, "lg": lambda: ou.lg_as_ops(cfg4_data(T=5)),
             "dirichlet": lambda: ou.dirichlet_categorical()[0]}[prog]()
    N, seed, it = 6000, 9, 2
    o = oracle.opts(oracle.SYSTEMATIC, seed=seed)
    m = oracle.Ops(steps)
    x_prev = None
    for T in range(1, len(steps) + 1):
        lz, x_gpu, lw = _run(sb, steps[:T], N, seed, it=it)
        if T == 1:
            x_ref, lw_ref = oracle.ops_step(m, o, N, 0, it, None)
        else:
            u = oracle.uniform53(seed, 0, T - 2, it, 1, 0)
            anc = oracle.resample_systematic(lw[T - 2], u, N, kind=1)
            x_ref, lw_ref = oracle.ops_step(m, o, N, T - 1, it, x_prev[..., anc])
        assert np.allclose(x_gpu, x_ref, rtol=1e-11, atol=1e-11), (prog, T, np.abs(x_gpu - x_ref).max())
        assert np.allclose(lw[T - 1], lw_ref, rtol=1e-9, atol=1e-9), (prog, T, np.abs(lw[T - 1] - lw_ref).max())
        x_prev = x_gpu
    ref = oracle.ops_sweep(m, o, N, it=it)
    assert lz == pytest.approx(ref["logZ"], rel=1e-6, abs=1e-6)


@pytest.mark.parametrize("prog", ["beta_bernoulli", "gamma_poisson", "dirichlet_categorical"])
def test_ops_evidence_closed_form(prog):
    """Beta-Bernoulli / Gamma-Poisson programs with several observations per observe statement: the log-evidence
    estimate agrees with the conjugate closed form within 4 standard errors over 16 seeds."""
    sb = sb_mod()
    steps, exact = getattr(ou, prog)()
    vals = np.array([_run(sb, steps, 1 << 17, 300 + s, store_logw=False)[0] for s in range(16)])
    se = vals.std(ddof=1) / math.sqrt(len(vals))
    assert abs(vals.mean() - exact) < 4 * se + 1e-4, (vals.mean(), exact, se)
    assert vals.std(ddof=1) < 0.05


def test_ops_lg_vs_kalman_and_lg_kernel(enumref):
    """cfg4's model written as an op list: evidence vs the Kalman filter, and the same posterior mean as the dedicated
    linear-Gaussian kernel (different random streams: statistical agreement)."""
    sb = sb_mod()
    d = cfg4_data(T=60)
    exact, _ = enumref.kalman_loglik(d["a"], d["q"], d["c"], d["r"], d["m0"], d["P0"], d["y"])
    vals, means = [], []
    for s in range(12):
        lz, x, _ = _run(sb, ou.lg_as_ops(d), 1 << 16, 700 + s, store_logw=False)
        vals.append(lz); means.append(x.mean())
    vals = np.array(vals)
    se = vals.std(ddof=1) / math.sqrt(len(vals))
    assert abs(vals.mean() - exact) < 4 * se + 1e-4, (vals.mean(), exact, se)
    with sb.Engine(precision="fp64", seed=5) as eng:
        eng.load(sb.LinearGaussian(**d))
        eng.smc_sweep(1 << 18)
        ref_mean = eng.particles(1 << 18).mean()
    assert abs(np.mean(means) - ref_mean) < 0.02


def test_ops_fp32_handle_and_large_pool():
    """an fp32 handle still computes the program in fp64 (same result as an fp64 handle); a pool of 2^20 particles
    (persistent grid, many tiles per CTA) is reproducible"""
    sb = sb_mod()
    steps, exact = ou.beta_bernoulli()
    a = _run(sb, steps, 50000, 4, store_logw=False, precision="fp32")
    b = _run(sb, steps, 50000, 4, store_logw=False, precision="fp64")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    c = _run(sb, steps, 1 << 20, 4, store_logw=False)
    d = _run(sb, steps, 1 << 20, 4, store_logw=False)
    assert c[0] == d[0] and np.array_equal(c[1], d[1]) and abs(c[0] - exact) < 0.05


@pytest.mark.parametrize("prog", ["beta_bernoulli", "dirichlet_categorical"])
def test_ops_multinomial_resampling(prog):
    """the same programs under multinomial resampling on the weight grid (explicit ancestors: the other ancestor mode of
    the step kernel, scalar and vector latents): evidence against the closed form"""
    sb = sb_mod()
    steps, exact = getattr(ou, prog)()
    vals = []
    for s in range(8):
        with sb.Engine(precision="fp64", resample="multinomial", seed=40 + s) as eng:
            eng.load(ou.to_program(sb, steps))
            vals.append(eng.smc_sweep(1 << 16))
            x = eng.particles(1 << 16)
            assert np.all(np.isfinite(x)) and (x.ndim == 1 or np.allclose(x.sum(axis=0), 1.0, atol=1e-12))
    vals = np.array(vals)
    se = vals.std(ddof=1) / math.sqrt(len(vals))
    assert abs(vals.mean() - exact) < 4 * se + 1e-3, (vals.mean(), exact, se)


def test_ops_errors():
    sb = sb_mod()
    with sb.Engine(precision="fp64") as eng:
        with pytest.raises(sb.StochyError):                      # Beta(0, 1)
            eng.load(sb.OpsProgram([(("beta", 0.0, 1.0), ("bernoulli", 1.0, 0.0, [1.0]))]))
        with pytest.raises(sb.StochyError):                      # Bernoulli observation outside {0, 1}
            eng.load(sb.OpsProgram([(("beta", 1.0, 1.0), ("bernoulli", 1.0, 0.0, [2.0]))]))
        with pytest.raises(sb.StochyError):                      # the first statement must draw
            eng.load(sb.OpsProgram([(("keep",), ("zero",))]))
        with pytest.raises(sb.StochyError):                      # Categorical observation outside 1..D
            eng.load(sb.OpsProgram([(("dirichlet", 1.0, 1.0, 1.0), ("categorical", [4.0]))]))
        with pytest.raises(sb.StochyError):                      # a scalar draw inside a Dirichlet program
            eng.load(sb.OpsProgram([(("dirichlet", 1.0, 1.0), ("categorical", [1.0])), (("beta", 1.0, 1.0), ("zero",))]))
        with pytest.raises(sb.StochyError):                      # more than 8 components
            eng.load(sb.OpsProgram([(("dirichlet",) + (1.0,) * 9, ("zero",))]))
        # p outside (0, 1): log of a negative number -> the reference's NaN error
        eng.load(sb.OpsProgram([(("uniform", 0.0, 1.0), ("bernoulli", 3.0, 0.0, [1.0, 0.0]))]))
        with pytest.raises(sb.StochyError) as ei:
            eng.smc_sweep(4096)
        assert "nan" in str(ei.value).lower()
