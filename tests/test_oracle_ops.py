"""CPU: the oracle's restatement of op-list programs (sample / observe of src/pmcmc.jl:32-46, src/Stochy.jl:55-56 with
Beta / Gamma / Normal / flip / uniform ERPs, src/erp.jl:1-30) against closed-form evidence and scipy's logpdfs.
The reference holds no numeric test of these ERPs' samplers or densities (Distributions.jl 0.6.3, not in the tree):
the forms are the textbook ones, pinned here."""
import math

import numpy as np
import pytest

import ops_util as ou


def _mean_logZ(oracle, steps, N, seeds):
    m = oracle.Ops(steps)
    vals = np.array([oracle.ops_sweep(m, oracle.opts(oracle.SYSTEMATIC, seed=100 + s), N)["logZ"] for s in range(seeds)])
    return vals.mean(), vals.std(ddof=1) / math.sqrt(seeds)


def test_beta_bernoulli_evidence(oracle):
    steps, exact = ou.beta_bernoulli()
    mean, se = _mean_logZ(oracle, steps, 1 << 14, 12)
    assert abs(mean - exact) < 4 * se + 1e-3, (mean, exact, se)


def test_gamma_poisson_evidence(oracle):
    steps, exact = ou.gamma_poisson()
    mean, se = _mean_logZ(oracle, steps, 1 << 14, 12)
    assert abs(mean - exact) < 4 * se + 1e-3, (mean, exact, se)


def test_dirichlet_categorical_evidence(oracle):
    steps, exact = ou.dirichlet_categorical()
    mean, se = _mean_logZ(oracle, steps, 1 << 14, 12)
    assert abs(mean - exact) < 4 * se + 1e-3, (mean, exact, se)
    # the draw itself: rows on the simplex with the Dirichlet's means
    x, lw = oracle.ops_step(oracle.Ops(steps), oracle.opts(oracle.SYSTEMATIC, seed=2), 100000, 0, 0, None)
    a = np.array([0.8, 1.5, 2.0, 0.6])
    assert x.shape == (4, 100000) and np.allclose(x.sum(axis=0), 1.0, atol=1e-12)
    assert np.allclose(x.mean(axis=1), a / a.sum(), atol=5e-3)
    ys = [int(y) for y in steps[0][4]]
    assert np.allclose(lw, sum(np.log(x[y - 1]) for y in ys), rtol=1e-12, atol=1e-12)


def test_lg_as_ops_vs_kalman(oracle, enumref):
    from gpu_util import cfg4_data
    d = cfg4_data(T=30)
    exact, _ = enumref.kalman_loglik(d["a"], d["q"], d["c"], d["r"], d["m0"], d["P0"], d["y"])
    mean, se = _mean_logZ(oracle, ou.lg_as_ops(d), 1 << 14, 12)
    assert abs(mean - exact) < 4 * se + 1e-3, (mean, exact, se)


def test_step_scores_and_draw_moments(oracle):
    """one step from given parents: scores equal scipy's logpdf sums; Beta / Gamma draws have the right moments"""
    from scipy import stats
    N = 200000
    o = oracle.opts(oracle.SYSTEMATIC, seed=8)
    ys = [1.0, 0.0, 0.0, 1.0]
    x, lw = oracle.ops_step(oracle.Ops([(ou.D_BETA, (2.0, 5.0), ou.S_BERNOULLI, (1.0, 0.0), ys)]), o, N, 0, 0, None)
    assert np.allclose(lw, sum(stats.bernoulli.logpmf(int(y), x) for y in ys), rtol=1e-12, atol=1e-12)
    assert abs(x.mean() - 2 / 7) < 4 * x.std() / math.sqrt(N) and abs(x.var() - 10 / (49 * 8)) < 2e-3
    assert stats.kstest(x, stats.beta(2.0, 5.0).cdf).pvalue > 1e-3
    yp = [3.0, 7.0]
    x, lw = oracle.ops_step(oracle.Ops([(ou.D_GAMMA, (0.7, 2.0), ou.S_POISSON, (2.0, 0.5), yp)]), o, N, 0, 0, None)
    assert np.allclose(lw, sum(stats.poisson.logpmf(int(y), 2.0 * x + 0.5) for y in yp), rtol=1e-11, atol=1e-11)
    assert stats.kstest(x, stats.gamma(0.7, scale=2.0).cdf).pvalue > 1e-3
    xp = np.linspace(-2, 2, N)
    yn = [0.3, -1.1, 2.0]
    x, lw = oracle.ops_step(oracle.Ops([(ou.D_NORMAL, (0.0, 0.0, 1.0), ou.S_ZERO, (), []),
                                        (ou.D_NORMAL, (0.9, 0.1, 0.5), ou.S_NORMAL, (1.5, -0.2, 0.8), yn)]), o, N, 1, 0, xp)
    assert np.allclose(lw, sum(stats.norm.logpdf(y, 1.5 * x - 0.2, 0.8) for y in yn), rtol=1e-12, atol=1e-12)
    z = (x - (0.9 * xp + 0.1)) / 0.5
    assert abs(z.mean()) < 4 / math.sqrt(N) and abs(z.var() - 1) < 0.02
    x, _ = oracle.ops_step(oracle.Ops([(ou.D_NORMAL, (0.0, 0.0, 1.0), ou.S_ZERO, (), []),
                                       (ou.D_FLIP, (0.5, 0.25), ou.S_ZERO, (), [])]), o, N, 1, 0, np.ones(N))
    assert set(np.unique(x)) <= {0.0, 1.0} and abs(x.mean() - 0.75) < 4 * math.sqrt(0.75 * 0.25 / N)
    x, _ = oracle.ops_step(oracle.Ops([(ou.D_UNIFORM, (-1.0, 3.0), ou.S_ZERO, (), [])]), o, N, 0, 0, None)
    assert x.min() >= -1.0 and x.max() <= 3.0 and abs(x.mean() - 1.0) < 4 * (4 / math.sqrt(12)) / math.sqrt(N)
