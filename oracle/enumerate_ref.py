"""Pure-Python restatement of the reference's exact-inference oracle and result type.

TEST INFRASTRUCTURE ONLY (see oracle/stochy_oracle.h).  Follows, relative to /root/reference/:
  src/enumerate.jl:36-47   sample = enqueue every support value with score + logpdf; factor adds
  src/enumerate.jl:63-85   exit accumulates exp(score) per return value; result = Discrete(returns)
  src/erp.jl:39-56         Discrete: normalise by z, z == 0 -> "Distribution has zero probability mass."
  src/erp.jl:100-121       hellingerdistance, kl
  src/erp.jl:22-30         flip -> (false,true)[Bernoulli+1]; randominteger(k) = Categorical(k), 1-based
Python has no CPS transform, so continuations are replaced by replaying the model with a forced
prefix of choices; the sequence of floating-point additions to the score is the same as the
reference's (program order along each path).

Closed forms used as independent cross-checks (no reference line): forward algorithm for the
discrete HMM, Kalman filter for the linear-Gaussian model.
"""
import math
from collections import deque

import numpy as np


# ---------------------------------------------------------------- ERPs (src/erp.jl:1-30)
class Bernoulli:
    def __init__(self, p=0.5):
        self.p = p

    def support(self):
        return (0, 1)

    def logpdf(self, x):
        q = self.p if x == 1 else 1.0 - self.p
        return math.log(q) if q > 0 else -math.inf


class Categorical:
    """Categorical(k) = uniform on 1..k; Categorical(ps) = 1-based index distribution."""

    def __init__(self, arg):
        self.ps = [1.0 / arg] * arg if isinstance(arg, int) else list(arg)

    def support(self):
        return tuple(range(1, len(self.ps) + 1))

    def logpdf(self, x):
        q = self.ps[x - 1]
        return math.log(q) if q > 0 else -math.inf


class Discrete:
    """src/erp.jl:39-56.  Also the result type of every inference entry point."""

    def __init__(self, d):
        self.x = list(d.keys())
        p = [float(v) for v in d.values()]
        z = sum(p)
        if z == 0:
            raise ZeroDivisionError("Distribution has zero probability mass.")   # erp.jl:48
        self.dict = dict(d)
        self.z = z
        self.p = [v / z for v in p] if z != 1.0 else p

    def support(self):
        return self.x

    def logpdf(self, x):
        return math.log(self.dict[x] / self.z)      # erp.jl:56

    def prob(self, x):
        return self.dict.get(x, 0.0) / self.z


def score(erp, x):
    return erp.logpdf(x)


# ---------------------------------------------------------------- Enumerate (src/enumerate.jl)
class _Abort(Exception):
    pass


class EnumCtx:
    def __init__(self, prefix, frontier):
        self.prefix = prefix
        self.frontier = frontier
        self.i = 0
        self.score = 0.0

    def sample(self, erp):
        if self.i < len(self.prefix):
            v = self.prefix[self.i]
            self.i += 1
            self.score = self.score + score(erp, v)            # enumerate.jl:38 (score carried in the queue entry)
            return v
        for v in erp.support():                                 # enumerate.jl:37-39
            self.frontier.append(self.prefix + (v,))
        raise _Abort()                                          # enumerate.jl:40 runnext

    def factor(self, s):
        self.score += s                                         # enumerate.jl:45

    def observe(self, erp, *xs):
        self.factor(sum(score(erp, x) for x in xs))             # src/Stochy.jl:55-56: ONE factor

    def flip(self, p=0.5):
        return (False, True)[self.sample(Bernoulli(p))]         # erp.jl:22-26

    def randominteger(self, k):
        return self.sample(Categorical(k))                      # erp.jl:28-30


def enum(model, maxexec=0, order="bfs"):
    """enum(comp, maxexec) src/enumerate.jl:57-85 (breadth-first default, :57,60)."""
    frontier = deque([()])
    returns = {}
    n = 0
    while frontier:
        prefix = frontier.popleft() if order == "bfs" else frontier.pop()
        ctx = EnumCtx(prefix, frontier)
        try:
            v = model(ctx)
        except _Abort:
            continue
        returns[v] = returns.get(v, 0.0) + math.exp(ctx.score)   # enumerate.jl:72
        n += 1
        if maxexec > 0 and n == maxexec:
            break
    return Discrete(returns)                                     # enumerate.jl:84


def hellingerdistance(p, q):
    """src/erp.jl:100-111; q's support must be a subset of p's."""
    ps, qs = list(p.support()), list(q.support())
    if not set(qs) <= set(ps):
        raise AssertionError("support mismatch")
    acc = 0.0
    for x in ps:
        px = math.exp(score(p, x))
        qx = math.exp(score(q, x)) if x in qs else 0.0
        acc += (math.sqrt(px) - math.sqrt(qx)) ** 2
    return math.sqrt(acc / 2.0)


def kl(p, q):
    """src/erp.jl:113-119."""
    return sum(math.exp(score(p, x)) * (score(p, x) - score(q, x)) for x in p.support())


def total_variation(p, q):
    keys = set(p) | set(q)
    return 0.5 * sum(abs(p.get(k, 0.0) - q.get(k, 0.0)) for k in keys)


# ---------------------------------------------------------------- closed forms
def hmm_forward(A, logF, x0=-1, pi=None, has_factor=None):
    """log Z and per-time smoothing marginals of the discrete HMM by forward-backward (fp64)."""
    A = np.asarray(A, dtype=np.float64)
    F = np.exp(np.asarray(logF, dtype=np.float64))
    T, K = F.shape
    if has_factor is not None:
        F = np.where(np.asarray(has_factor, dtype=bool)[:, None], F, 1.0)
    p0 = A[x0] if x0 >= 0 else (np.full(K, 1.0 / K) if pi is None else np.asarray(pi, dtype=np.float64))
    alpha = np.zeros((T, K))
    c = np.zeros(T)
    a = p0 * F[0]
    c[0] = a.sum()
    alpha[0] = a / c[0]
    for t in range(1, T):
        a = (alpha[t - 1] @ A) * F[t]
        c[t] = a.sum()
        alpha[t] = a / c[t]
    beta = np.ones((T, K))
    for t in range(T - 2, -1, -1):
        beta[t] = (A @ (F[t + 1] * beta[t + 1])) / c[t + 1]
    post = alpha * beta
    post /= post.sum(axis=1, keepdims=True)
    return float(np.log(c).sum()), post


def hmm_enumerate_joint(A, logF, x0=-1, pi=None, has_factor=None):
    """Exact joint over trajectories by brute force = what enum() returns for the hmm2.jl form.
    Returns (Z, probs[K**T]) with little-endian base-K trajectory codes (digit t = x_t)."""
    A = np.asarray(A, dtype=np.float64)
    logF = np.asarray(logF, dtype=np.float64)
    T, K = logF.shape
    p0 = A[x0] if x0 >= 0 else (np.full(K, 1.0 / K) if pi is None else np.asarray(pi, dtype=np.float64))
    hf = np.ones(T, dtype=bool) if has_factor is None else np.asarray(has_factor, dtype=bool)
    w = np.zeros(K ** T)
    # vectorised over all K^T codes
    codes = np.arange(K ** T)
    digits = [(codes // K ** t) % K for t in range(T)]
    lp = np.log(p0)[digits[0]] + np.where(hf[0], logF[0][digits[0]], 0.0)
    with np.errstate(divide="ignore"):
        logA = np.log(A)
    for t in range(1, T):
        lp = lp + logA[digits[t - 1], digits[t]] + (logF[t][digits[t]] if hf[t] else 0.0)
    w = np.exp(lp)
    Z = w.sum()
    return float(Z), w / Z


def kalman_loglik(a, q, c, r, m0, P0, y):
    """log p(y_{1:T}) of x_1~N(m0,P0), x_t = a x_{t-1} + N(0,q), y_t = c x_t + N(0,r) (variances)."""
    m, P = m0, P0
    ll = 0.0
    means = []
    for t, yt in enumerate(y):
        if t > 0:
            m, P = a * m, a * a * P + q
        S = c * c * P + r
        v = yt - c * m
        ll += -0.5 * (math.log(2.0 * math.pi * S) + v * v / S)
        Kg = P * c / S
        m, P = m + Kg * v, (1.0 - Kg * c) * P
        means.append(m)
    return ll, np.array(means)
