"""The weight grid (the deterministic fixed-point CDF both the oracle and the CUDA path use) is an
EXTENSION; these tests tie it back to the reference's literal resampler (src/rand.jl:2-13 called
from src/pmcmc.jl:48-52) and check its own invariants."""
import numpy as np
import pytest


def dyadic_weights(rng, n, bits=10, frac=16):
    """weights k_i / 2^frac whose sum is a power of two: every partial sum is exact in fp64, so any
    summation order reproduces the reference's sequential CDF bit for bit."""
    k = rng.integers(0, 2 ** bits, size=n).astype(np.int64)
    s = int(k.sum())
    target = 1 << int(np.ceil(np.log2(max(s, 2))))
    k[-1] += target - s          # last weight absorbs the remainder (may be large; still exact)
    w = k.astype(np.float64) / 2.0 ** frac
    return w


@pytest.mark.parametrize("n", [1, 2, 5, 100, 2048, 2049, 5000, 70000])
def test_grid_multinomial_equals_literal_on_dyadic(oracle, n):
    rng = np.random.default_rng(n)
    w = dyadic_weights(rng, n)
    r = rng.random(3000)
    lit = oracle.resample_literal(w, r)                 # O(n*m) literal loop
    fast = oracle.resample_literal(w, r, fast=True)
    grid = oracle.resample_multinomial(w, r, kind=0)
    assert np.array_equal(lit, fast)
    assert np.array_equal(lit, grid)


def test_literal_fast_equals_literal_general(oracle):
    rng = np.random.default_rng(7)
    for n in [3, 101, 4097]:
        w = rng.random(n) ** 4
        w[rng.integers(0, n, size=n // 5)] = 0.0
        r = rng.random(2000)
        assert np.array_equal(oracle.resample_literal(w, r), oracle.resample_literal(w, r, fast=True))


def test_grid_vs_literal_mismatch_rate_general(oracle):
    # on general weights the grid and the sequentially rounded fp64 CDF differ only for uniforms
    # that fall within rounding distance of a boundary
    rng = np.random.default_rng(11)
    n, m = 1 << 14, 1 << 16
    w = rng.random(n)
    r = rng.random(m)
    a = oracle.resample_literal(w, r, fast=True)
    b = oracle.resample_multinomial(w, r, kind=0)
    assert np.mean(a != b) < 1e-4
    assert np.max(np.abs(a - b)) <= 1


def test_systematic_properties(oracle):
    rng = np.random.default_rng(5)
    for n, m in [(1, 7), (10, 10), (2048, 2048), (5000, 3000), (3000, 5000), (1 << 16, 1 << 16)]:
        w = rng.random(n) ** 3
        u = rng.random()
        anc = oracle.resample_systematic(w, u, m, kind=0)
        assert anc.min() >= 0 and anc.max() < n
        assert np.all(np.diff(anc) >= 0)                       # sorted
        cnt = np.bincount(anc, minlength=n)
        expect = m * w / w.sum()
        assert np.all(np.abs(cnt - expect) < 1.0 + 1e-6)       # |offspring - m p_i| < 1
    # zero weights never get offspring; a single heavy particle takes everything
    w = np.zeros(5000)
    w[1234] = 3.0
    assert np.all(oracle.resample_systematic(w, 0.3, 777, kind=0) == 1234)


def test_systematic_exact_on_dyadic(oracle):
    # equal dyadic weights, m == n: child c has threshold (c+u)/n, parent = c for every u in [0,1)
    n = 4096
    w = np.full(n, 0.25)
    for u in [0.0, 0.5, 0.999999]:
        assert np.array_equal(oracle.resample_systematic(w, u, n, kind=0), np.arange(n))


def test_log_kinds_agree_with_linear(oracle):
    rng = np.random.default_rng(3)
    n = 10000
    lw = rng.normal(size=n) * 3 - 50.0
    u = 0.37
    a1 = oracle.resample_systematic(lw, u, n, kind=1)
    a0 = oracle.resample_systematic(np.exp(lw - lw.max()), u, n, kind=0)
    assert np.mean(a1 != a0) < 5e-3
    a2 = oracle.resample_systematic(lw.astype(np.float32), u, n, kind=2)
    assert np.mean(a2 != a1) < 2e-2
    # log-weights far below fp64 exp underflow are fine on the grid (D2/H3 extension)
    a3 = oracle.resample_systematic(lw - 5000.0, u, n, kind=1)
    assert np.mean(a3 != a1) < 5e-3


def test_grid_errors(oracle):
    with pytest.raises(oracle.OracleError) as e:
        oracle.resample_systematic(np.array([1.0, np.nan]), 0.5, 2, kind=0)
    assert e.value.rc == oracle.ENAN
    with pytest.raises(oracle.OracleError) as e:
        oracle.resample_systematic(np.zeros(10), 0.5, 2, kind=0)
    assert e.value.rc == oracle.EZEROMASS
    with pytest.raises(oracle.OracleError) as e:
        oracle.resample_systematic(np.full(10, -np.inf), 0.5, 2, kind=1)
    assert e.value.rc == oracle.EZEROMASS
    with pytest.raises(oracle.OracleError):
        oracle.resample_literal(np.array([0.0, 0.0]), np.array([0.5]))      # 0/0 -> NaN (D2)


def test_q31_accuracy(oracle):
    L = oracle.lib()
    for d in np.linspace(-28.5, 0.0, 2001):
        q = L.or_q31_f64(float(d))
        assert abs(q - 2.0 ** (d + 27)) <= 1.0 + 2.0 ** (d + 27) * 1e-14
        qf = L.or_q31_f32(float(np.float32(d)))
        assert abs(qf - 2.0 ** (float(np.float32(d)) + 27)) <= 1.0 + 2.0 ** (d + 27) * 1e-6
    assert L.or_q31_f64(0.0) in (2 ** 27 - 1, 2 ** 27)
    assert L.or_q31_f64(-40.0) == 0
    assert L.or_q31_f64(float("-inf")) == 0
