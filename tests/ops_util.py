"""Op-list programs used by the CPU (oracle) and GPU tests, with their closed-form evidence."""
import math

import numpy as np

D_KEEP, D_NORMAL, D_BETA, D_GAMMA, D_UNIFORM, D_FLIP, D_DIRICHLET = range(7)
S_ZERO, S_NORMAL, S_BERNOULLI, S_POISSON, S_CATEGORICAL = range(5)
_DN = ["keep", "normal", "beta", "gamma", "uniform", "flip", "dirichlet"]
_SN = ["zero", "normal", "bernoulli", "poisson", "categorical"]


def beta_bernoulli(a=2.0, b=3.0, batches=6, per=3, seed=3):
    """p = ~Beta(a, b); then `batches` statements observe(Bernoulli(p), ys...) of `per` observations each
    (src/Stochy.jl:56: several observations per observe).  Evidence: B(a + s, b + f) / B(a, b)."""
    rng = np.random.default_rng(seed)
    ys = (rng.random((batches, per)) < 0.35).astype(float)
    steps = [(D_BETA if t == 0 else D_KEEP, (a, b) if t == 0 else (), S_BERNOULLI, (1.0, 0.0), list(ys[t])) for t in range(batches)]
    s, f = ys.sum(), ys.size - ys.sum()
    lb = lambda x, y: math.lgamma(x) + math.lgamma(y) - math.lgamma(x + y)
    return steps, lb(a + s, b + f) - lb(a, b)


def gamma_poisson(k=3.0, theta=1.5, batches=5, per=2, seed=4):
    """lam = ~Gamma(k, theta); observe(Poisson(lam), ys...) per batch.  Evidence: negative-binomial product."""
    rng = np.random.default_rng(seed)
    ys = rng.poisson(4.0, size=(batches, per)).astype(float)
    steps = [(D_GAMMA if t == 0 else D_KEEP, (k, theta) if t == 0 else (), S_POISSON, (1.0, 0.0), list(ys[t])) for t in range(batches)]
    S, n = ys.sum(), ys.size
    ev = (math.lgamma(k + S) - math.lgamma(k) - sum(math.lgamma(y + 1.0) for y in ys.ravel())
          + S * math.log(theta) - (k + S) * math.log(1.0 + n * theta))
    return steps, ev


def dirichlet_categorical(alpha=(0.8, 1.5, 2.0, 0.6), batches=5, per=4, seed=5):
    """theta = ~Dir(alpha); observe(Categorical(theta), ys...) per batch (ys 1-based).  Evidence: Dirichlet-multinomial
    Gamma(a0) / Gamma(a0 + n) * prod_k Gamma(a_k + n_k) / Gamma(a_k)."""
    rng = np.random.default_rng(seed)
    D = len(alpha)
    ys = rng.choice(D, size=(batches, per), p=[0.1, 0.5, 0.3, 0.1][:D]) + 1
    steps = [(D_DIRICHLET if t == 0 else D_KEEP, tuple(alpha) if t == 0 else (), S_CATEGORICAL, (), [float(y) for y in ys[t]]) for t in range(batches)]
    a0, n = sum(alpha), ys.size
    ev = math.lgamma(a0) - math.lgamma(a0 + n)
    for k in range(D):
        nk = int((ys == k + 1).sum())
        ev += math.lgamma(alpha[k] + nk) - math.lgamma(alpha[k])
    return steps, ev


def lg_as_ops(d):
    """BASELINE cfg4's linear-Gaussian model written as an op list (checked against the Kalman filter)."""
    steps = []
    for t, y in enumerate(d["y"]):
        draw = (D_NORMAL, (0.0, d["m0"], math.sqrt(d["P0"]))) if t == 0 else (D_NORMAL, (d["a"], 0.0, math.sqrt(d["q"])))
        steps.append((draw[0], draw[1], S_NORMAL, (d["c"], 0.0, math.sqrt(d["r"])), [float(y)]))
    return steps


def uniform_flip_mix(T=5, seed=6):
    """exercises uniform / flip / normal draws and the zero score: u ~ Uniform(.2,.8); z = flip(u); x ~ Normal(2 z - 1, .5) ..."""
    rng = np.random.default_rng(seed)
    steps = [(D_UNIFORM, (0.2, 0.8), S_BERNOULLI, (1.0, 0.0), [1.0, 0.0, 1.0]),
             (D_FLIP, (1.0, 0.0), S_ZERO, (), []),
             (D_NORMAL, (2.0, -1.0, 0.5), S_NORMAL, (1.0, 0.0, 1.0), list(rng.normal(size=2)))]
    for t in range(3, T):
        steps.append((D_NORMAL, (0.8, 0.0, 0.7), S_NORMAL, (1.0, 0.5, 1.2), list(rng.normal(size=1 + t % 2))))
    return steps


def to_program(sb, steps):
    """the product's descriptor (names) from the integer-coded step list"""
    out = []
    for draw, dp, score, sp, ys in steps:
        d = (_DN[draw],) + tuple(dp)
        s = (_SN[score],) + tuple(sp) + (list(ys),) if score != S_ZERO else ("zero",)
        out.append((d, s))
    return sb.OpsProgram(out)
