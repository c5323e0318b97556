"""Pins the CPU oracle against every known-answer test the reference holds for this path
 and against the hand-derived goldens of SURVEY.md §4."""
import math

import numpy as np
import pytest


# ---- test/inferencetests.jl:3-15 : Enumerate KAT
def test_enum_three_flips(enumref):
    def model(c):
        a, b, cc = c.flip(0.5), c.flip(0.5), c.flip(0.5)
        c.factor(0 if (a or b) else -1)
        return int(a) + int(b) + int(cc)
    d = enumref.enum(model)
    for v, p in [(0, 0.054615886286517), (1, 0.351538628762172), (2, 0.445384113713482), (3, 0.148461371237827)]:
        assert math.exp(d.logpdf(v)) == pytest.approx(p, abs=1e-14)


# ---- test/inferencetests.jl:17-20 : zero mass error
def test_enum_zero_mass(enumref):
    def model(c):
        c.factor(-math.inf)
        return True
    with pytest.raises(ZeroDivisionError):
        enumref.enum(model)


# ---- test/memtests.jl:7-11 analogue: P(true)=0.7 regardless of exploration order
@pytest.mark.parametrize("order", ["bfs", "dfs"])
def test_enum_orders(enumref, order):
    def model(c):
        g = c.flip(0.7)
        c.flip(0.6)
        return g
    d = enumref.enum(model, order=order)
    assert math.exp(d.logpdf(True)) == pytest.approx(0.7, abs=1e-14)
    assert math.exp(d.logpdf(False)) == pytest.approx(0.3, abs=1e-14)


# ---- test/randtests.jl:1-2 : NaN error in the categorical draw
def test_rand_nan(oracle):
    with pytest.raises(oracle.OracleError) as e:
        oracle.rand_categorical([math.nan], 0.5)
    assert "unexpected nan in rand" in str(e.value)
    with pytest.raises(oracle.OracleError):
        oracle.rand_categorical([math.nan, math.nan], 0.5)


# ---- src/rand.jl:6-12 : strict <, last index fallback
def test_rand_semantics(oracle):
    ps = [0.25, 0.25, 0.5]
    assert oracle.rand_categorical(ps, 0.0) == 0
    assert oracle.rand_categorical(ps, 0.2499999) == 0
    assert oracle.rand_categorical(ps, 0.25) == 1          # r < acc is strict
    assert oracle.rand_categorical(ps, 0.5) == 2
    assert oracle.rand_categorical(ps, 0.999999) == 2
    assert oracle.rand_categorical([1.0], 0.3) == 0
    # last entry is never added to acc: un-normalised tails fall through to n
    assert oracle.rand_categorical([0.1, 0.1, 0.1], 0.9) == 2


# ---- test/erptests.jl:1-10 : Discrete
def test_discrete(enumref):
    hist = {0: 2, 1: 3, 2: 5}
    erp = enumref.Discrete(hist)
    assert sorted(erp.support()) == [0, 1, 2]
    for x in range(3):
        assert erp.logpdf(x) == pytest.approx(math.log(hist[x] / 10), abs=1e-15)


# ---- test/erptests.jl:12-17 : Hellinger KAT
def test_hellinger(enumref):
    p = enumref.Categorical([0.25, 0.25, 0.5])
    q = enumref.Discrete({1: 2, 3: 3})
    assert enumref.hellingerdistance(p, q) == pytest.approx(0.36885, abs=1e-5)
    with pytest.raises(AssertionError):
        enumref.hellingerdistance(p, enumref.Discrete({4: 1}))


# ---- examples/hmm.jl:4-32 golden
def test_enum_hmm_example(enumref):
    def model(c):
        states, obs = [True], []
        for _ in range(3):
            s = c.flip(0.7) if states[-1] else c.flip(0.3)
            o = c.flip(0.9) if s else c.flip(0.1)
            states.append(s)
            obs.append(o)
        c.factor(0 if obs == [False, False, False] else -math.inf)
        return tuple(states[1:])
    d = enumref.enum(model)
    assert d.z == pytest.approx(0.12916, abs=1e-12)
    F, T = False, True
    gold = {(F, F, F): 0.8296918550634872, (T, F, F): 0.09218798389594302, (F, F, T): 0.039509135955404145,
            (F, T, F): 0.01693248683803035, (T, T, F): 0.010243109321771441, (F, T, T): 0.004389903995044904,
            (T, F, T): 0.004389903995044904, (T, T, T): 0.0026556209352740765}
    for k, v in gold.items():
        assert d.prob(k) == pytest.approx(v, rel=1e-12)


HMM2_A = [[0.9, 0.1], [0.1, 0.9]]
HMM2_LOGF = [[-1.0, 0.0]] * 4


# ---- examples/hmm2.jl:5-24 golden, three ways
def test_enum_hmm2_example(enumref):
    def model(c):
        states = [False]
        for o in [True] * 4:
            s = c.flip(0.9 if states[-1] else 0.1)
            c.factor(0 if s == o else -1)
            states.append(s)
        return tuple(states)
    d = enumref.enum(model)
    assert d.z == pytest.approx(0.13253212222378208, rel=1e-13)
    F, T = False, True
    gold = {(F, T, T, T, T): 0.5500553282992594, (F, F, T, T, T): 0.20235404678810578,
            (F, F, F, F, F): 0.09067153285758026, (F, F, F, T, T): 0.07444189365118824,
            (F, F, F, F, T): 0.027385642236143076, (F, T, T, T, F): 0.02248378297645619}
    for k, v in gold.items():
        assert d.prob(k) == pytest.approx(v, rel=1e-12)
    assert len(d.support()) == 16
    # the table-form descriptor of the same model agrees (forward algorithm + brute-force joint)
    logZ, post = enumref.hmm_forward(HMM2_A, HMM2_LOGF, x0=0)
    assert logZ == pytest.approx(-2.0209302310601185, rel=1e-13)
    Z, joint = enumref.hmm_enumerate_joint(HMM2_A, HMM2_LOGF, x0=0)
    assert Z == pytest.approx(0.13253212222378208, rel=1e-13)
    for k, v in gold.items():
        code = sum(int(b) << t for t, b in enumerate(k[1:]))
        assert joint[code] == pytest.approx(v, rel=1e-12)
    # smoothing marginals from the joint == forward-backward
    for t in range(4):
        pt = sum(joint[c] for c in range(16) if (c >> t) & 1)
        assert post[t, 1] == pytest.approx(pt, rel=1e-12)


def test_philox_kat(oracle):
    # Random123 known-answer vectors for philox4x32-10
    assert list(oracle.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_kalman_against_dense_gaussian(enumref):
    # log p(y) from the Kalman recursion == log-density of the joint Gaussian of y
    rng = np.random.default_rng(0)
    T, a, q, c, r, m0, P0 = 6, 0.9, 1.0, 1.0, 1.0, 0.0, 1.0
    y = rng.normal(size=T)
    ll, _ = enumref.kalman_loglik(a, q, c, r, m0, P0, y)
    # build cov of x then y
    Sx = np.zeros((T, T))
    var = np.zeros(T)
    var[0] = P0
    for t in range(1, T):
        var[t] = a * a * var[t - 1] + q
    for i in range(T):
        for j in range(T):
            lo, hi = min(i, j), max(i, j)
            Sx[i, j] = var[lo] * a ** (hi - lo)
    Sy = c * c * Sx + r * np.eye(T)
    mu = np.array([c * m0 * a ** t for t in range(T)])
    d = y - mu
    ref = -0.5 * (T * math.log(2 * math.pi) + np.linalg.slogdet(Sy)[1] + d @ np.linalg.solve(Sy, d))
    assert ll == pytest.approx(ref, rel=1e-12)
