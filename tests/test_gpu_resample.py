"""GPU parity of the resampling chain (quantise -> tile scan -> pull / search -> gather) against the
CPU oracle, through the C ABI.  Bar: bit-exact ancestor indices (integer work)."""
import numpy as np
import pytest

from gpu_util import dyadic_weights, need_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    need_gpu()
    import stochy_jl_b200 as sb
    e = sb.Engine(precision="fp64")
    yield e
    e.close()


SIZES = [1, 2, 7, 2047, 2048, 2049, 4096, 5000, 100003, (1 << 20) + 3]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_systematic_bitexact(eng, oracle, n, kind):
    rng = np.random.default_rng(n * 3 + kind)
    if kind == 0:
        w = rng.random(n) ** 3
    else:
        w = rng.normal(size=n) * 4 - 300.0
    for u in (0.0, rng.random(), 0.9999999):
        got = eng.resample_systematic(w, u, n, kind=kind)
        ref = oracle.resample_systematic(w, u, n, kind=kind)
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("n,m", [(10, 100000), (100000, 10), (5000, 3000), (3000, 5000), (2049, 2048), (1, 5000),
                                 (3000000, 2500001), (2000000, 3100000)])       # the last two: persistent pull grid (many tiles per CTA)
def test_systematic_ragged(eng, oracle, n, m):
    rng = np.random.default_rng(n + m)
    w = rng.random(n)
    got = eng.resample_systematic(w, 0.61, m)
    assert np.array_equal(got, oracle.resample_systematic(w, 0.61, m))


def test_systematic_degenerate(eng, oracle):
    rng = np.random.default_rng(1)
    n = 300000
    # one heavy particle; a heavy particle in the last (ragged) tile; many zeros; 600 orders of magnitude
    cases = []
    w = np.zeros(n); w[123456] = 2.5; cases.append((w, 0))
    w = np.zeros(n); w[n - 1] = 1e-300; cases.append((w, 0))
    w = rng.random(n); w[rng.random(n) < 0.95] = 0.0; cases.append((w, 0))
    w = np.full(n, 1e-12); w[::4099] = 1.0; cases.append((w, 0))                 # one child per ~2 tiles
    lw = rng.normal(size=n) * 200.0; cases.append((lw, 1))
    lw = np.full(n, -np.inf); lw[77] = -1e4; lw[n - 5] = -1e4 - 3; cases.append((lw, 1))
    # parents with hundreds of children each (one parent covers several 256-slot warp boundaries of a children
    # tile), alone in their thread, two in one thread, and with light neighbours that still own a child
    w = np.full(n, 1e-9); w[::300] = 1.0; cases.append((w, 0))
    w = np.full(n, 1e-9); w[5::997] = 1.0; w[6::997] = 0.7; cases.append((w, 0))
    w = np.full(n, 0.004); w[3::640] = 1.0; w[7::640] = 0.45; cases.append((w, 0))
    w = rng.random(n) ** 40; cases.append((w, 0))
    for w, kind in cases:
        got = eng.resample_systematic(w, 0.37, n, kind=kind)
        assert np.array_equal(got, oracle.resample_systematic(w, 0.37, n, kind=kind))


def test_systematic_retained_pool(eng, oracle):
    # the N+1 pool of conditional SMC (src/pmcmc.jl:59-66): draw N from N+1
    rng = np.random.default_rng(2)
    for N in (100, 2048, 4096, 65536):
        w = rng.random(N + 1)
        got = eng.resample_systematic(w, 0.2, N)
        assert np.array_equal(got, oracle.resample_systematic(w, 0.2, N))
        assert got.max() <= N


@pytest.mark.parametrize("n", [1, 5, 2048, 2049, 70001])
def test_multinomial_bitexact_and_reference(eng, oracle, n):
    rng = np.random.default_rng(n)
    r = rng.random(5000)
    w = rng.random(n) ** 2
    assert np.array_equal(eng.resample_multinomial(w, r), oracle.resample_multinomial(w, r))
    lw = np.log(w + 1e-30)
    assert np.array_equal(eng.resample_multinomial(lw, r, kind=1), oracle.resample_multinomial(lw, r, kind=1))
    # on exactly representable weights the grid reproduces the reference's literal algorithm
    wd = dyadic_weights(rng, n)
    lit = oracle.resample_literal(wd, r)
    assert np.array_equal(eng.resample_multinomial(wd, r), lit)
    assert np.array_equal(eng.resample_literal(wd, r), lit)


@pytest.mark.parametrize("n", [1, 2, 100, 101, 2048, 5000, 66000])
def test_literal_bitexact(eng, oracle, n):
    """src/rand.jl:2-13 + src/pmcmc.jl:48-52 on arbitrary weights: sequential fp64 numerics on the GPU."""
    rng = np.random.default_rng(n)
    w = rng.random(n) ** 5
    w[rng.integers(0, n, size=n // 7)] = 0.0
    if w.sum() == 0:
        w[0] = 1.0
    r = np.concatenate([rng.random(4000), [0.0, 0.5, 1.0 - 2 ** -53]])
    assert np.array_equal(eng.resample_literal(w, r), oracle.resample_literal(w, r, fast=(n > 5000)))


def test_error_conventions(eng):
    import stochy_jl_b200 as sb
    # test/randtests.jl:1-2
    for w in ([np.nan], [np.nan, np.nan]):
        with pytest.raises(sb.StochyError) as e:
            eng.resample_literal(np.array(w), np.array([0.5]))
        assert "unexpected nan in rand" in str(e.value)
    with pytest.raises(sb.StochyError) as e:
        eng.resample_literal(np.zeros(4), np.array([0.5]))          # 0/0 (D2)
    assert e.value.code == sb._native.ENAN
    with pytest.raises(sb.StochyError) as e:
        eng.resample_systematic(np.array([1.0, np.nan, 2.0]), 0.5, 3)
    assert e.value.code == sb._native.ENAN
    with pytest.raises(sb.StochyError) as e:
        eng.resample_systematic(np.zeros(5000), 0.5, 10)
    assert e.value.code == sb._native.EZEROMASS
    with pytest.raises(sb.StochyError) as e:
        eng.resample_systematic(np.full(10, -np.inf), 0.5, 10, kind=1)
    assert e.value.code == sb._native.EZEROMASS
    with pytest.raises(sb.StochyError) as e:
        eng.resample_systematic(np.array([0.0, np.inf]), 0.5, 2, kind=1)
    assert e.value.code == sb._native.ENAN
    # the handle stays usable after an error (ctx restored, test/inferencetests.jl:46-53)
    assert list(eng.resample_systematic(np.array([1.0, 1.0]), 0.5, 2)) == [0, 1]


@pytest.mark.parametrize("n", [1, 31, 8192, 100000, 3000001])
def test_logsumexp(eng, oracle, n):
    rng = np.random.default_rng(n)
    lw = rng.normal(size=n) * 30 - 1000.0
    ref = oracle.logsumexp(lw)
    assert eng.logsumexp(lw) == pytest.approx(ref, rel=1e-12)
    assert eng.logsumexp(lw.astype(np.float32), kind=2) == pytest.approx(oracle.logsumexp(lw.astype(np.float32).astype(np.float64)), rel=1e-12)
    lw[:] = -np.inf
    assert eng.logsumexp(lw) == -np.inf


def test_gather(eng):
    rng = np.random.default_rng(4)
    for dt in (np.int32, np.float32, np.float64):
        src = (rng.random(100001) * 1000).astype(dt)
        anc = rng.integers(0, len(src), size=77777).astype(np.int32)
        assert np.array_equal(eng.gather(src, anc), src[anc])


def test_full_size_properties(eng, oracle):
    """BASELINE size (2^24): properties that do not need the oracle to finish in seconds, plus
    bit-exact parity against the OpenMP oracle (it does finish in seconds)."""
    n = 1 << 24
    rng = np.random.default_rng(24)
    lw = (rng.standard_normal(n) * 2.0).astype(np.float32)
    u = 0.123456789
    anc = eng.resample_systematic(lw, u, n, kind=2)
    assert anc.min() >= 0 and anc.max() < n
    assert np.all(np.diff(anc) >= 0)                               # sortedness
    w = np.exp(lw.astype(np.float64))
    cnt = np.bincount(anc, minlength=n)
    assert np.all(np.abs(cnt - n * w / w.sum()) < 1.0 + 1e-3)      # |offspring - N p_i| < 1
    assert np.array_equal(anc, oracle.resample_systematic(lw, u, n, kind=2))
    # idempotence: resampling the (uniform) resampled population with equal weights is the identity
    assert np.array_equal(eng.resample_systematic(np.ones(n), u, n), np.arange(n))
