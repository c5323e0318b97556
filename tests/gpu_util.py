import numpy as np
import pytest


def need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def dyadic_weights(rng, n, bits=10, frac=16):
    k = rng.integers(0, 2 ** bits, size=n).astype(np.int64)
    s = int(k.sum())
    target = 1 << int(np.ceil(np.log2(max(s, 2))))
    k[-1] += target - s
    return k.astype(np.float64) / 2.0 ** frac


CFG2 = dict(
    A=np.array([[.7, .2, .1], [.2, .6, .2], [.1, .3, .6]]),
    B=np.array([[.8, .1, .1], [.1, .8, .1], [.1, .1, .8]]),
    y=[0, 0, 1, 2, 2, 1, 0, 0, 2, 1],
)


def cfg2_tables():
    A, B, y = CFG2["A"], CFG2["B"], CFG2["y"]
    return A, np.log(B[:, y].T), np.full(3, 1.0 / 3.0)


def cfg3_tables(T=1000, K=16, seed=12345):
    """sticky 16-state HMM, 16 symbols, y simulated with default_rng(12345)."""
    A = np.full((K, K), 0.01)
    np.fill_diagonal(A, 0.85)
    B = np.full((K, K), 0.02)
    np.fill_diagonal(B, 0.70)
    rng = np.random.default_rng(seed)
    x = rng.integers(0, K)
    ys = []
    for _ in range(T):
        x = rng.choice(K, p=A[x])
        ys.append(rng.choice(K, p=B[x]))
    return A, np.log(B[:, ys].T), np.full(K, 1.0 / K)


def cfg4_data(T=500, seed=12345):
    """x0~N(0,1); x_t = 0.9 x_{t-1} + N(0,1); y_t = x_t + N(0,1)."""
    rng = np.random.default_rng(seed)
    x = rng.normal()
    ys = []
    for t in range(T):
        if t:
            x = 0.9 * x + rng.normal()
        ys.append(x + rng.normal())
    return dict(a=0.9, q=1.0, c=1.0, r=1.0, m0=0.0, P0=1.0, y=np.array(ys))


def cfg5_bnet():
    """C~flip(.5); S~flip(C?.1:.5); R~flip(C?.8:.2); observe(Bernoulli(pW[S][R]), true)."""
    with np.errstate(divide="ignore"):
        logf = np.log(np.array([0.0, 0.9, 0.9, 0.99]))       # index = S + 2R
    sites = [(0, [0.5]), (1, [0.5, 0.1]), (1, [0.2, 0.8])]
    factors = [(0b110, logf)]
    return sites, factors, [2]
