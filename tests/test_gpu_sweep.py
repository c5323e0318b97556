"""GPU parity of the full SMC sweep / PMCMC / MH against the CPU oracle on the same Philox stream
(bit-exact for the discrete model), and against Enumerate / forward algorithm / Kalman closed forms."""
import numpy as np
import pytest

from gpu_util import cfg2_tables, cfg3_tables, cfg4_data, cfg5_bnet, need_gpu

pytestmark = pytest.mark.gpu

MODES = {"systematic": 0, "multinomial": 1, "literal": 2}


def sb_mod():
    need_gpu()
    import stochy_jl_b200 as sb
    return sb


@pytest.mark.parametrize("resample", ["systematic", "multinomial", "literal"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("N", [100, 2048, 5000])
def test_hmm_sweep_bitexact(oracle, resample, precision, N):
    sb = sb_mod()
    A, logF, pi = cfg2_tables()
    T = logF.shape[0]
    with sb.Engine(precision=precision, resample=resample, seed=11, store_logw=True) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        lz = eng.smc_sweep(N, it=3, history=True)
        ref = oracle.hmm_sweep(oracle.Hmm(A, logF, pi=pi), oracle.opts(MODES[resample], fp32=(precision == "fp32"), seed=11), N, it=3)
        for t in range(T):
            x, a, lw = eng.history(t, N, want_lw=True)
            assert np.array_equal(x, ref["x"][t, :N]), t
            assert np.array_equal(a, ref["anc"][t, :N]), t
            assert np.array_equal(lw, ref["lw"][t, :N]), t
        # log-evidence: 1e-5 relative is the north-star bar; the grid is far tighter
        assert lz == pytest.approx(ref["logZ"], rel=1e-9 if resample != "literal" else 1e-6)


def test_hmm2_fixed_x0_and_partial_factors(oracle):
    sb = sb_mod()
    # examples/hmm2.jl form (fixed initial state) and the examples/hmm.jl form (one terminal hard factor)
    for model in (sb.hmm2_example(), sb.hmm_example(3)):
        for resample in ("systematic", "multinomial", "literal"):
            with sb.Engine(precision="fp64", resample=resample, seed=5) as eng:
                eng.load(model)
                N = 3000
                lz = eng.smc_sweep(N, it=0, history=True)
                om = oracle.Hmm(model.A, model.logF, x0=model.x0, has_factor=model.has_factor)
                ref = oracle.hmm_sweep(om, oracle.opts(MODES[resample], seed=5), N)
                for t in range(model.T):
                    x, a, _ = eng.history(t, N)
                    assert np.array_equal(x, ref["x"][t, :N])
                    assert np.array_equal(a, ref["anc"][t, :N])
                assert lz == pytest.approx(ref["logZ"], rel=1e-6, abs=1e-9)


@pytest.mark.parametrize("resample", ["systematic", "multinomial", "literal"])
@pytest.mark.parametrize("N", [100, 2048])
def test_pmcmc_counts_bitexact(oracle, resample, N):
    """Conditional SMC with the retained particle appended as pool entry N+1 (src/pmcmc.jl:59-66),
    retained = particle 1 (src/pmcmc.jl:89), ALL N values counted (src/pmcmc.jl:90-92)."""
    sb = sb_mod()
    A, logF, pi = cfg2_tables()
    iters = 20
    with sb.Engine(precision="fp64", resample=resample, seed=3) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        marg, joint = eng.pmcmc_counts(iters, N)
        xs = eng.retained()
    ref = oracle.hmm_pmcmc(oracle.Hmm(A, logF, pi=pi), oracle.opts(MODES[resample], seed=3), iters, N)
    assert marg.sum() == iters * N * logF.shape[0]
    assert np.array_equal(marg, ref["marg"])
    assert np.array_equal(joint, ref["joint"])
    assert xs.min() >= 0 and xs.max() < 3


def test_cfg2_pmcmc_vs_enumerate(enumref):
    """BASELINE config 2: T=10, K=3 PMCMC 100 particles x 1000 iterations in the reference's own
    numerics vs Stochy's exhaustive Enumerate.  Stated bound: TV <= 0.02 on per-time
    marginals; joint TV over 59049 outcomes reported (bound 0.15: 1e5 samples on ~6e4 cells)."""
    sb = sb_mod()
    A, logF, pi = cfg2_tables()
    Z, joint_exact = enumref.hmm_enumerate_joint(A, logF, pi=pi)
    _, post = enumref.hmm_forward(A, logF, pi=pi)
    with sb.Engine(precision="fp64", resample="literal", seed=1) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        marg, joint = eng.pmcmc_counts(1000, 100)
    pm = marg / marg.sum(axis=1, keepdims=True)
    tv_t = 0.5 * np.abs(pm - post).sum(axis=1)
    assert tv_t.max() <= 0.02, tv_t
    tv_joint = 0.5 * np.abs(joint / joint.sum() - joint_exact).sum()
    assert tv_joint <= 0.15, tv_joint


def test_hmm2_api_vs_enumerate(enumref):
    """examples/hmm2.jl through the reference-shaped API: smc(100) / pmcmc(1000,100) vs enum()."""
    sb = sb_mod()
    m = sb.hmm2_example()
    Z, joint_exact = enumref.hmm_enumerate_joint(m.A, m.logF, x0=0)
    exact = {}
    for code, p in enumerate(joint_exact):
        exact[m.value(tuple((code >> t) & 1 for t in range(4)))] = p
    d = sb.pmcmc(m, 1000, 100, seed=2)
    est = {k: d.prob(k) for k in d.support()}
    assert enumref.total_variation(est, exact) <= 0.02
    d1 = sb.smc(m, 100, seed=2)
    assert set(d1.support()) <= set(exact)
    # test/inferencetests.jl:36-44 smoke: flip + soft factor, support == {true,false}
    flipm = sb.DiscreteHMM([[0.5, 0.5], [0.5, 0.5]], [[-1.0, 0.0]], x0=0, value=lambda xs: bool(xs[0]))
    d2 = sb.pmcmc(flipm, 5, 5)
    assert set(d2.support()) <= {True, False}
    d3 = sb.pmcmc(flipm, 50, 50)
    assert set(d3.support()) == {True, False}
    assert d3.prob(True) == pytest.approx(1 / (1 + np.exp(-1.0)), abs=0.05)


def test_zero_mass_and_hard_factor():
    sb = sb_mod()
    # test/memtests.jl:52-68 spirit: -Inf factors with a surviving particle must work ...
    A = [[0.5, 0.5], [0.5, 0.5]]
    ok = sb.DiscreteHMM(A, [[0.0, -np.inf], [0.0, 0.0]], x0=0)
    with sb.Engine(precision="fp64", seed=1) as eng:
        eng.load(ok)
        eng.smc_sweep(1000, history=True)
        x0, a0, _ = eng.history(0, 1000)
        assert np.any(x0 == 1) and np.all(x0[a0] == 0)
    # ... and an all -Inf factor is the reference's zero-mass / NaN error (test/inferencetests.jl:17-20)
    bad = sb.DiscreteHMM(A, [[-np.inf, -np.inf]], x0=0)
    for resample in ("systematic", "literal"):
        with sb.Engine(precision="fp64", resample=resample) as eng:
            eng.load(bad)
            with pytest.raises(sb.StochyError):
                eng.smc_sweep(100)


def test_cfg3_logZ_vs_forward(enumref):
    """BASELINE config 3 model (K=16, sticky) at reduced T: log Z-hat of unconditional sweeps vs the
    forward algorithm, |mean - exact| within 4 standard errors over 16 seeds."""
    sb = sb_mod()
    A, logF, pi = cfg3_tables(T=200)
    exact, _ = enumref.hmm_forward(A, logF, pi=pi)
    vals = []
    for seed in range(16):
        with sb.Engine(precision="fp32", resample="systematic", seed=100 + seed) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            vals.append(eng.smc_sweep(1 << 16))
    vals = np.array(vals)
    se = vals.std(ddof=1) / np.sqrt(len(vals))
    assert abs(vals.mean() - exact) < 4 * se + 1e-3, (vals.mean(), exact, se)
    assert vals.std(ddof=1) < 0.3


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_lg_stepwise_parity(oracle, precision):
    """Linear-Gaussian model: every step's propagate+weight is checked against the oracle given the
    SAME parents (tolerance: transcendental functions differ in the last ulp between libm and CUDA);
    ancestors are checked bit-exactly by feeding the GPU's own log-weights to the oracle's grid."""
    sb = sb_mod()
    d = cfg4_data(T=12)
    N = 6000
    fp32 = precision == "fp32"
    m = oracle.Lg(d["a"], d["q"], d["c"], d["r"], d["m0"], d["P0"], d["y"])
    o = oracle.opts(oracle.SYSTEMATIC, fp32=fp32, seed=9)
    tol = 2e-5 if fp32 else 1e-11
    for T in range(1, 5):
        dd = dict(d); dd["y"] = d["y"][:T]
        with sb.Engine(precision=precision, seed=9, store_logw=True) as eng:
            eng.load(sb.LinearGaussian(**dd))
            eng.smc_sweep(N, it=2)
            x_gpu = eng.particles(N)
            lw_gpu = eng.logw_row(T - 1, N)
            lw_prev = eng.logw_row(T - 2, N) if T > 1 else None
        if T == 1:
            x_ref, lw_ref = oracle.lg_step(m, o, N, 0, 2, None)
        else:
            # parents as the GPU saw them: previous-step particles come from a (T-1)-step run
            dd2 = dict(d); dd2["y"] = d["y"][:T - 1]
            with sb.Engine(precision=precision, seed=9, store_logw=True) as e2:
                e2.load(sb.LinearGaussian(**dd2))
                e2.smc_sweep(N, it=2)
                x_prev = e2.particles(N)
            u = oracle.uniform53(9, 0, T - 2, 2, 1, 0)
            anc = oracle.resample_systematic(lw_prev.astype(np.float32) if fp32 else lw_prev, u, N, kind=2 if fp32 else 1)
            x_ref, lw_ref = oracle.lg_step(m, o, N, T - 1, 2, x_prev[anc])
        assert np.allclose(x_gpu, x_ref, rtol=tol, atol=tol)
        assert np.allclose(lw_gpu, lw_ref, rtol=1e-5 if not fp32 else 2e-4, atol=tol * 10)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_lg_logZ_vs_kalman(enumref, precision):
    """BASELINE config 4 model at reduced size: |log Z-hat - log Z_Kalman| within 4 sigma over 16 seeds."""
    sb = sb_mod()
    d = cfg4_data(T=100)
    exact, _ = enumref.kalman_loglik(d["a"], d["q"], d["c"], d["r"], d["m0"], d["P0"], d["y"])
    vals = []
    for seed in range(16):
        with sb.Engine(precision=precision, seed=500 + seed) as eng:
            eng.load(sb.LinearGaussian(**d))
            vals.append(eng.smc_sweep(1 << 17))
    vals = np.array(vals)
    se = vals.std(ddof=1) / np.sqrt(len(vals))
    assert abs(vals.mean() - exact) < 4 * se + (2e-3 if precision == "fp32" else 1e-4), (vals.mean(), exact, se)
    assert vals.std(ddof=1) < 0.05


def test_mh_bitexact_and_posterior(oracle, enumref):
    """BASELINE config 5 model: batched MH counts equal the oracle's chain for chain; pooled estimate
    of P(R | W) matches the Enumerate posterior 0.7079."""
    sb = sb_mod()
    sites, factors, query = cfg5_bnet()
    net = sb.BayesNet(sites, factors, query)
    ob = oracle.Bnet([m for m, _ in sites], [c for _, c in sites], [m for m, _ in factors], [t for _, t in factors], net.query_mask)
    with sb.Engine(precision="fp64", seed=21) as eng:
        eng.load(net)
        got = eng.mh_counts(4096, 200, chain0=5)
        acc = eng.stats().accepts
    ref, racc = oracle.bnet_mh(ob, 21, 4096, 200, chain0=5)
    assert np.array_equal(got, ref)
    assert acc == racc

    def model(c):
        C_ = c.flip(0.5)
        S = c.flip(0.1 if C_ else 0.5)
        R = c.flip(0.8 if C_ else 0.2)
        pw = {(False, False): 0.0, (True, False): 0.9, (False, True): 0.9, (True, True): 0.99}[(S, R)]
        c.observe(enumref.Bernoulli(pw), 1)
        return R
    exact = enumref.enum(model).prob(True)
    assert exact == pytest.approx(0.7079, abs=5e-4)
    with sb.Engine(precision="fp64", seed=22) as eng:
        eng.load(net)
        cnt = eng.mh_counts(65536, 2000)
    assert cnt[1] / cnt.sum() == pytest.approx(exact, abs=3e-3)
    d = sb.mh(net, 500, chains=2048, seed=4)                       # reference-shaped API
    assert set(d.support()) == {True, False}
    assert d.prob(True) == pytest.approx(exact, abs=0.02)


def test_erp_vocabulary():
    """flip / randominteger / Normal / Beta / Gamma / Dirichlet (src/erp.jl:1-30): moments of the
    samplers, log-densities against scipy.stats."""
    sb = sb_mod()
    from scipy import stats
    n = 400000
    with sb.Engine(precision="fp64") as eng:
        nat = sb._native
        f = eng.erp_sample(nat.ERP_FLIP, [0.3], n, seed=1)
        assert set(np.unique(f)) == {0.0, 1.0} and f.mean() == pytest.approx(0.3, abs=4e-3)
        r = eng.erp_sample(nat.ERP_RANDOMINTEGER, [7], n, seed=2)
        assert r.min() == 1 and r.max() == 7                                   # 1-based, src/erp.jl:28-30
        assert np.bincount(r.astype(int))[1:] / n == pytest.approx(np.full(7, 1 / 7), abs=4e-3)
        z = eng.erp_sample(nat.ERP_NORMAL, [1.5, 2.0], n, seed=3)
        assert stats.kstest(z, stats.norm(1.5, 2.0).cdf).pvalue > 1e-3
        b = eng.erp_sample(nat.ERP_BETA, [2.0, 5.0], n, seed=4)
        assert stats.kstest(b, stats.beta(2.0, 5.0).cdf).pvalue > 1e-3
        b2 = eng.erp_sample(nat.ERP_BETA, [0.5, 0.5], n, seed=5)
        assert stats.kstest(b2, stats.beta(0.5, 0.5).cdf).pvalue > 1e-3
        g = eng.erp_sample(nat.ERP_GAMMA, [3.0, 2.0], n, seed=6)
        assert stats.kstest(g, stats.gamma(3.0, scale=2.0).cdf).pvalue > 1e-3
        alpha = np.array([0.7, 2.0, 3.5])
        dsamp = eng.dirichlet_sample(alpha, n, seed=7)
        assert np.allclose(dsamp.sum(axis=1), 1.0)
        assert dsamp.mean(axis=0) == pytest.approx(alpha / alpha.sum(), abs=3e-3)
        x = np.linspace(0.01, 0.99, 99)
        assert np.allclose(eng.erp_logpdf(nat.ERP_BETA, [2.0, 5.0], x), stats.beta.logpdf(x, 2.0, 5.0), rtol=1e-12, atol=1e-12)
        assert np.allclose(eng.erp_logpdf(nat.ERP_NORMAL, [1.5, 2.0], x), stats.norm.logpdf(x, 1.5, 2.0), rtol=1e-12)
        assert np.allclose(eng.erp_logpdf(nat.ERP_GAMMA, [3.0, 2.0], x), stats.gamma.logpdf(x, 3.0, scale=2.0), rtol=1e-12, atol=1e-12)
        assert np.allclose(eng.erp_logpdf(nat.ERP_FLIP, [0.5], np.array([0.0, 1.0])), np.log(0.5))
        assert np.allclose(eng.erp_logpdf(nat.ERP_FLIP, [0.7], np.array([0.0, 1.0])), [np.log1p(-0.7), np.log(0.7)])
        xs = dsamp[:1000]
        assert np.allclose(eng.dirichlet_logpdf(alpha, xs), [stats.dirichlet.logpdf(v / v.sum(), alpha) for v in xs], rtol=1e-9, atol=1e-9)


def test_full_size_sweep_properties(oracle, enumref):
    """BASELINE size (2^24 particles): size-independent properties of a whole sweep.
    (a) every ancestor row is sorted, in range, and its offspring counts are within one of N p_i
        computed from that step's states; (b) the log-evidence is within 0.02 of the forward algorithm;
    (c) the final particle set equals the one of a second handle with the same seed (determinism) and
        differs under another seed; (d) the oracle agrees bit for bit at 2^20 (it finishes in seconds there)."""
    sb = sb_mod()
    T, N = 6, 1 << 24
    A, logF, pi = cfg3_tables(T=T)
    exact, _ = enumref.hmm_forward(A, logF, pi=pi)
    with sb.Engine(precision="fp32", resample="systematic", seed=5) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        lz = eng.smc_sweep(N, it=0, history=True)
        assert abs(lz - exact) < 0.02, (lz, exact)
        for t in range(T):
            x, a, _ = eng.history(t, N)
            assert x.min() >= 0 and x.max() < 16
            assert a.min() >= 0 and a.max() < N and np.all(np.diff(a) >= 0)
            w = np.exp(logF[t])[x]
            cnt = np.bincount(a, minlength=N)
            assert np.all(np.abs(cnt - N * w / w.sum()) < 1.0 + 1e-2), "offspring counts off at step %d" % t
        xa = eng.particles(N)
    with sb.Engine(precision="fp32", resample="systematic", seed=5) as e2:
        e2.load(sb.DiscreteHMM(A, logF, pi=pi))
        assert e2.smc_sweep(N, it=0) == lz
        assert np.array_equal(e2.particles(N), xa)
    with sb.Engine(precision="fp32", resample="systematic", seed=6) as e3:
        e3.load(sb.DiscreteHMM(A, logF, pi=pi))
        e3.smc_sweep(N, it=0)
        assert not np.array_equal(e3.particles(N), xa)
    M = 1 << 20
    with sb.Engine(precision="fp32", resample="systematic", seed=5) as e4:
        e4.load(sb.DiscreteHMM(A, logF, pi=pi))
        lz4 = e4.smc_sweep(M, it=0)
        ref = oracle.hmm_sweep(oracle.Hmm(A, logF, pi=pi), oracle.opts(oracle.SYSTEMATIC, fp32=True, seed=5), M)
        assert np.array_equal(e4.particles(M), ref["x"][T - 1, :M])
        assert lz4 == pytest.approx(ref["logZ"], rel=1e-9)


def test_full_size_lg_vs_kalman(enumref):
    """BASELINE config 4 at full particle count (2^24, reduced T): log-evidence against the Kalman filter.
    At this N the Monte-Carlo error is ~1e-3 per sweep; fp32 weights add a rounding floor."""
    sb = sb_mod()
    d = cfg4_data(T=50)
    exact, _ = enumref.kalman_loglik(d["a"], d["q"], d["c"], d["r"], d["m0"], d["P0"], d["y"])
    for prec, tol in [("fp32", 2e-2), ("fp64", 1e-2)]:
        with sb.Engine(precision=prec, seed=77) as eng:
            eng.load(sb.LinearGaussian(**d))
            lz = eng.smc_sweep(1 << 24)
        assert abs(lz - exact) < tol, (prec, lz, exact)


@pytest.mark.parametrize("K,N,T", [(1, 50, 3), (2, 1, 4), (37, 3000, 5), (64, 2500, 4), (64, 2048, 1), (5, 2047, 6), (5, 2049, 6)])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_hmm_edge_models_bitexact(oracle, K, N, T, precision):
    """Edge shapes of the discrete model: one state, one particle, one step, the largest K (64 KB of lookup words:
    opt-in shared memory), rows with zero and tiny probabilities (several thresholds inside one lookup bucket: the
    scanning instantiation of the step kernel), pools that end exactly at / just past a tile boundary."""
    sb = sb_mod()
    rng = np.random.default_rng(1000 + K + N)
    A = rng.random((K, K)) ** 6                      # a few large entries, many tiny ones
    A[rng.random((K, K)) < 0.2] = 0.0                # exact zeros
    A[np.arange(K), np.arange(K)] += 1e-3            # no empty row
    A /= A.sum(axis=1, keepdims=True)
    pi = rng.random(K) ** 4 + 1e-6
    pi /= pi.sum()
    logF = rng.normal(size=(T, K)) * 3.0
    if K > 2:
        logF[:, 0] = -np.inf                         # a state with no weight at all
    fp32 = precision == "fp32"
    ref = oracle.hmm_sweep(oracle.Hmm(A, logF, pi=pi), oracle.opts(oracle.SYSTEMATIC, fp32=fp32, seed=21), N)
    with sb.Engine(precision=precision, resample="systematic", seed=21) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        lz = eng.smc_sweep(N, it=0, history=True)
        for t in range(T):
            x, a, _ = eng.history(t, N)
            assert np.array_equal(x, ref["x"][t, :N]), "states differ at step %d" % t
            assert np.array_equal(a, ref["anc"][t, :N]), "ancestors differ at step %d" % t
    assert lz == pytest.approx(ref["logZ"], rel=1e-9, abs=1e-12)


@pytest.mark.parametrize("K,N", [(1, 4095), (1, 2047), (5, 30000)])
def test_steps_without_factor_and_wide_score_range(oracle, monkeypatch, K, N):
    """Steps without a factor statement (no resampling, src/pmcmc.jl:36-46 is simply not reached) between factor steps
    whose scores differ by dozens of powers of two: the tile exponents of a factor-less step must not leak into the next
    scan's running maximum (found by tests/fuzz/fuzz_sweep.py: the stale exponent shifted the next pool's sums to zero and
    the sweep raised 'zero probability mass').  Tiled kernels (the one-tile chain kernel off) against the oracle."""
    sb = sb_mod()
    monkeypatch.setenv("SB_NO_SMALL", "1")
    rng = np.random.default_rng(7 + K)
    T = 6
    A = rng.random((K, K)) + 0.1
    A /= A.sum(axis=1, keepdims=True)
    pi = np.full(K, 1.0 / K)
    logF = rng.normal(size=(T, K)) * 0.5 + np.array([9.2, -2.0, 34.2, -14.4, -25.9, -31.9])[:, None]
    hf = np.array([1, 0, 0, 1, 0, 1], dtype=np.uint8)
    for prec in ["fp32", "fp64"]:
        ref = oracle.hmm_sweep(oracle.Hmm(A, logF, pi=pi, has_factor=hf), oracle.opts(oracle.SYSTEMATIC, fp32=(prec == "fp32"), seed=3), N)
        with sb.Engine(precision=prec, resample="systematic", seed=3) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi, has_factor=hf))
            lz = eng.smc_sweep(N, it=0, history=True)
            for t in range(T):
                x, a, _ = eng.history(t, N)
                assert np.array_equal(x, ref["x"][t, :N]) and np.array_equal(a, ref["anc"][t, :N]), (prec, t)
        assert lz == pytest.approx(ref["logZ"], rel=1e-9)


def test_result_independent_of_grid(monkeypatch):
    """The persistent step kernel walks tiles in an order that depends on the number of resident CTAs; the result
    must not.  One CTA per SM at 2^25 particles gives every CTA 110 tiles (the per-CTA table of candidate parent
    tiles is refilled every 64), the default grid gives 22: final particles and log-evidence must be identical."""
    sb = sb_mod()
    A, logF, pi = cfg3_tables(T=24)
    N = 1 << 25
    outs = []
    for ctas in ["1", "3", None, None]:      # the default grid twice: run-to-run identical too (split barriers, no races)
        if ctas:
            monkeypatch.setenv("SB_CTAS_PER_SM", ctas)
        else:
            monkeypatch.delenv("SB_CTAS_PER_SM", raising=False)
        with sb.Engine(precision="fp32", resample="systematic", seed=9) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            lz = eng.smc_sweep(N, it=0)
            outs.append((lz, eng.particles(N)))
    for o in outs[1:]:
        assert outs[0][0] == o[0]
        assert np.array_equal(outs[0][1], o[1])


@pytest.mark.parametrize("log2n", [25, 26, 27])
def test_large_pool_logZ_vs_forward(enumref, log2n):
    """Pools beyond 2^24 particles use the tile-scan instantiations that keep 2, 4 and 8 tile records per thread.  Their
    record loads once sat ABOVE the grid-dependency wait in the compiled code (ld.global.nc hoisted by ptxas: the scan
    read tile sums the step kernel had not written yet; log-evidence off by ~0.015 at 2^26, run-to-run different).
    The Monte-Carlo error at these sizes is far below the tolerance; two runs must agree bit for bit."""
    sb = sb_mod()
    A, logF, pi = cfg3_tables(T=24)
    exact, _ = enumref.hmm_forward(A, logF, pi=pi)
    vals = []
    for _ in range(2):
        with sb.Engine(precision="fp32", resample="systematic", seed=9) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            vals.append(eng.smc_sweep(1 << log2n, it=0))
    assert vals[0] == vals[1]
    assert abs(vals[0] - exact) < 4e-3, (vals[0], exact)


CHAIN_KEY = 0x9E3779B97F4A7C15


@pytest.mark.parametrize("resample", ["systematic", "multinomial", "literal"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_pmcmc_chain_set_bitexact(oracle, resample, precision):
    """A set of independent PMCMC chains, one CTA per chain with the whole iteration loop on chip: the pooled
    counts equal the sum of the oracle's runs with the chains' Philox keys (seed + c * 0x9E37...), bit for bit,
    in all three numerics; chain 0 alone equals the tiled kernels (SB_NO_SMALL)."""
    sb = sb_mod()
    A, logF, pi = cfg2_tables()
    iters, N, chains, seed = 6, 37, 5, 12
    with sb.Engine(precision=precision, resample=resample, seed=seed) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        marg, joint = eng.pmcmc_chains_counts(chains, iters, N)
        lz0 = eng.stats().logZ_first
    m = oracle.Hmm(A, logF, pi=pi)
    rm = np.zeros_like(marg)
    rj = np.zeros_like(joint)
    for c in range(chains):
        o = oracle.opts(MODES[resample], fp32=(precision == "fp32"), seed=(seed + c * CHAIN_KEY) % (1 << 64))
        ref = oracle.hmm_pmcmc(m, o, iters, N)
        rm += ref["marg"]
        rj += ref["joint"]
        if c == 0:
            assert lz0 == pytest.approx(ref["logZ_first"], rel=1e-9 if resample != "literal" else 1e-6)
    assert np.array_equal(marg, rm)
    assert np.array_equal(joint, rj)


@pytest.mark.parametrize("resample", ["systematic", "multinomial", "literal"])
def test_pmcmc_one_tile_route_equals_tiled_kernels(monkeypatch, resample):
    """sb_pmcmc_run sends pools that fit one tile to the one-CTA kernel; the tiled kernels (forced with
    SB_NO_SMALL=1) must give identical counts, retained particle and log-evidence."""
    sb = sb_mod()
    A, logF, pi = cfg2_tables()
    outs = []
    for no_small in ["1", None]:
        if no_small:
            monkeypatch.setenv("SB_NO_SMALL", no_small)
        else:
            monkeypatch.delenv("SB_NO_SMALL", raising=False)
        with sb.Engine(precision="fp64", resample=resample, seed=4) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            marg, joint = eng.pmcmc_counts(8, 100)
            outs.append((marg, joint, eng.retained(), eng.stats().logZ_first, eng.stats().kernel_launches))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])
    assert outs[0][3] == pytest.approx(outs[1][3], rel=1e-12)
    assert outs[1][4] == 1 and outs[0][4] >= 8 * 3           # one launch instead of several per iteration (hundreds when the sweeps run step by step)


@pytest.mark.parametrize("D", [1, 5, 8, 10])
def test_mh_random_nets_bitexact(oracle, monkeypatch, D):
    """Random binary Bayes nets (up to three parents per site, two factor statements, one of them hard): the batched
    chains equal the oracle's chain for chain, with the per-(site, state) tables (D <= 8) and without (SB_MH_NOTAB)."""
    sb = sb_mod()
    rng = np.random.default_rng(40 + D)
    sites, factors = [], []
    for d in range(D):
        npar = min(d, int(rng.integers(0, 4)))
        par = rng.choice(d, size=npar, replace=False) if npar else []
        mask = int(sum(1 << int(p) for p in par))
        sites.append((mask, rng.uniform(0.05, 0.95, size=1 << npar).tolist()))
    for f in range(2):
        npar = min(D, 2)
        par = rng.choice(D, size=npar, replace=False)
        mask = int(sum(1 << int(p) for p in par))
        tab = rng.normal(size=1 << npar)
        if f == 1 and D > 1:
            tab[0] = -np.inf                                    # a hard factor: -Inf scores must reject, NaN ratios too
        factors.append((mask, tab.tolist()))
    query = sorted(set(int(q) for q in rng.choice(D, size=min(D, 2), replace=False)))
    net = sb.BayesNet(sites, factors, query)
    ob = oracle.Bnet([m for m, _ in sites], [c for _, c in sites], [m for m, _ in factors], [t for _, t in factors], net.query_mask)
    ref, racc = oracle.bnet_mh(ob, 9, 1000, 150, chain0=3)
    for notab in [None, "1"]:
        if notab:
            monkeypatch.setenv("SB_MH_NOTAB", notab)
        else:
            monkeypatch.delenv("SB_MH_NOTAB", raising=False)
        with sb.Engine(precision="fp64", seed=9) as eng:
            eng.load(net)
            got = eng.mh_counts(1000, 150, chain0=3)
            assert np.array_equal(got, ref), "counts differ (SB_MH_NOTAB=%s)" % notab
            assert eng.stats().accepts == racc


def test_handle_reuse_across_sizes_and_models():
    """One handle, many calls: growing and shrinking pools, discrete and linear-Gaussian models, hooks in between.
    Every result must equal the result of a fresh handle (buffers regrown, cached launch configurations, state left
    behind by the previous call must not leak into the next)."""
    sb = sb_mod()
    A, logF, pi = cfg3_tables(T=5)
    d = cfg4_data(T=5)
    hmm, lg = sb.DiscreteHMM(A, logF, pi=pi), sb.LinearGaussian(**d)
    plan = [("hmm", 100), ("hmm", 5000), ("lg", 70000), ("hmm", (1 << 20) + 5), ("hook", 30001), ("lg", 17),
            ("hmm", 3000), ("pmcmc", 300), ("hmm", 2048), ("lg", 1 << 21), ("pmcmc", 2500), ("hmm", 17)]
    rng = np.random.default_rng(3)
    w = rng.random(30001)

    def run(eng, what, n):
        if what == "hmm":
            eng.load(hmm)
            return eng.smc_sweep(n, it=1), eng.particles(n)
        if what == "lg":
            eng.load(lg)
            return eng.smc_sweep(n, it=1), eng.particles(n)
        if what == "hook":
            return 0.0, eng.resample_systematic(w, 0.77, n)
        eng.load(hmm)
        marg, _ = eng.pmcmc_counts(3, n, want_joint=False)
        return eng.stats().logZ_last, marg

    with sb.Engine(precision="fp32", resample="systematic", seed=8) as shared:
        for what, n in plan:
            got = run(shared, what, n)
            with sb.Engine(precision="fp32", resample="systematic", seed=8) as fresh:
                ref = run(fresh, what, n)
            assert got[0] == ref[0], (what, n)
            assert np.array_equal(got[1], ref[1]), (what, n)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_lg_outlier_takes_the_waiting_path(monkeypatch, enumref, precision):
    """The linear-Gaussian step kernel skips the wait for the tile exponent when a warp's own maximum already reaches
    the density's mode.  An outlying observation (40 standard deviations) puts every particle far below the mode, so
    every warp has to wait for the other warps' maxima; with several tiles per CTA the split barrier's phases are
    exercised too.  The result must not depend on the grid.  (No particle reaches an observation that far out, so
    the estimate of the evidence falls short of the Kalman filter's by a wide margin - it can only be bounded above.)"""
    sb = sb_mod()
    d = cfg4_data(T=6)
    d["y"] = d["y"].copy()
    d["y"][3] = 40.0
    exact, _ = enumref.kalman_loglik(d["a"], d["q"], d["c"], d["r"], d["m0"], d["P0"], d["y"])
    N = 1 << 22
    outs = []
    for ctas in ["1", None]:
        if ctas:
            monkeypatch.setenv("SB_CTAS_PER_SM", ctas)
        else:
            monkeypatch.delenv("SB_CTAS_PER_SM", raising=False)
        with sb.Engine(precision=precision, seed=4) as eng:
            eng.load(sb.LinearGaussian(**d))
            lz = eng.smc_sweep(N)
            outs.append((lz, eng.particles(N)))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.isfinite(outs[0][0]) and outs[0][0] < exact + 1.0, (outs[0][0], exact)


def test_discrete_erp_hooks(oracle):
    """Discrete(x, p) as a model primitive (src/erp.jl:39-59): draws are bit-exact against the oracle's rand(ps)
    (src/rand.jl:2-13) on the same Philox uniforms, after the reference's normalisation p /= z (only when z != 1);
    score = log(dict[x] / z); zero mass and NaN raise the reference's errors."""
    sb = sb_mod()
    from stochy_jl_b200 import _native as nat
    rng = np.random.default_rng(3)
    with sb.Engine(precision="fp64") as eng:
        for p in ([0.2, 0.5, 0.3], [3.0, 1.0, 0.0, 4.0, 2.0], [1.0], [0.0, 0.0, 7.0], list(rng.random(37)), [0.25, 0.25, 0.5]):
            p = np.array(p, dtype=np.float64)
            z = 0.0
            for v in p:
                z += v
            ps = p / z if z != 1.0 else p
            n = 2500
            got = eng.discrete_sample(p, n, seed=8)
            ref = np.array([oracle.rand_categorical(ps, oracle.uniform53(8, i, 0, 0, 5, 0)) + 1 for i in range(n)])
            assert np.array_equal(got, ref)
            freq = np.bincount(got, minlength=len(p) + 1)[1:] / n
            assert np.abs(freq - p / z).max() < 0.05
            idx = np.arange(0, len(p) + 2)
            with np.errstate(divide="ignore"):
                want = np.concatenate([[-np.inf], np.log(p / z), [-np.inf]])
            assert np.array_equal(eng.discrete_logpdf(p, idx), want)
        with pytest.raises(sb.StochyError) as ei:
            eng.discrete_sample([0.0, 0.0], 10)
        assert ei.value.code == nat.EZEROMASS and "zero probability mass" in str(ei.value)
        with pytest.raises(sb.StochyError) as ei:
            eng.discrete_sample([0.3, float("nan"), 0.2], 1000)
        assert ei.value.code == nat.ENAN and "unexpected nan in rand" in str(ei.value)
        assert np.array_equal(eng.discrete_sample([0.5, 0.5], 4, seed=1), eng.discrete_sample([0.5, 0.5], 4, seed=1))


@pytest.mark.parametrize("log2n,T", [(16, 8), (20, 8), (20, 40)])
def test_tiled_conditional_smc_with_history_full_size(oracle, log2n, T):
    """BASELINE cfg3's shape at its full particle count (K = 16 sticky HMM, N = 2^20, conditional SMC with
    history; T reduced to 8 so the oracle finishes in seconds): iteration 0 is a plain sweep, iteration 1 a
    CONDITIONAL sweep whose pool of N+1 carries the retained particle (src/pmcmc.jl:59-66).  Every history
    row (states and ancestors, including the retained slot N), the next retained trajectory (src/pmcmc.jl:89)
    and the per-time value counts over all N particles of both iterations (src/pmcmc.jl:90-92) must equal
    the oracle's, bit for bit.  T = 40 at 2^20: the backward pass switches from its counting phase to the walks of the
    live lineages part-way (k_count_back_all), on the resident sweep kernel's history."""
    sb = sb_mod()
    N = 1 << log2n
    A, logF, pi = cfg3_tables(T=T)
    om = oracle.Hmm(A, logF, pi=pi)
    oo = oracle.opts(oracle.SYSTEMATIC, fp32=True, seed=9)
    ref0 = oracle.hmm_sweep(om, oo, N, it=0)
    ref1 = oracle.hmm_sweep(om, oo, N, it=1, xstar=ref0["xstar"])
    refc = oracle.hmm_pmcmc(om, oo, 2, N, want_joint=False)
    with sb.Engine(precision="fp32", resample="systematic", seed=9) as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        marg, _ = eng.pmcmc_counts(2, N, want_joint=False)
        st = eng.stats()
        assert np.array_equal(marg, refc["marg"])
        assert marg.sum() == 2 * N * T
        for t in range(T):
            x, a, _ = eng.history(t, N + 1)
            assert np.array_equal(x, ref1["x"][t, :N + 1]), t          # slot N = the retained particle's state
            assert np.array_equal(a[:N], ref1["anc"][t, :N]), t
        assert np.array_equal(eng.retained(), ref1["xstar"])
        assert st.logZ_first == pytest.approx(ref0["logZ"], rel=1e-9)
        assert st.logZ_last == pytest.approx(ref1["logZ"], rel=1e-9)


def test_cfg1_hmm_example_pmcmc_vs_enumerate():
    """BASELINE cfg1: examples/hmm.jl with line 27's enum() replaced by pmcmc(1000, 100) (iterations first,
    src/pmcmc.jl:78), reference numerics; total variation against exhaustive enumeration (the product's own
    enum on the GPU, and the hand-derived goldens of SURVEY section 4)."""
    import json
    import os
    sb = sb_mod()
    m = sb.hmm_example(3)
    d = sb.pmcmc(m, 1000, 100, seed=1)
    exact = sb.enum(m)
    keys = set(d.support()) | set(exact.support())
    tv = 0.5 * sum(abs(d.prob(k) - exact.prob(k)) for k in keys)
    assert tv <= 0.02, tv
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))["hmm_example_enum"]["p"]
    for name, p in kat.items():
        key = tuple(c == "T" for c in name)
        assert exact.prob(key) == pytest.approx(p, rel=1e-10)
        assert abs(d.prob(key) - p) <= 0.02


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("N", [5000, 70001, 1 << 20])
def test_resident_sweep_equals_tiled_kernels(monkeypatch, precision, N):
    """Pools whose tiles are all resident at once run a whole sweep as ONE cooperative launch (sb_resident.cuh).  The
    tiled kernels (forced with SB_NO_RESIDENT=1: one step kernel + one tile scan per step) must give identical
    states and ancestors at every step, log-evidence, PMCMC counts and retained trajectory -- plain and conditional
    sweeps (src/pmcmc.jl:59-66, 83-94)."""
    sb = sb_mod()
    A, logF, pi = cfg3_tables(T=7)
    outs = []
    for no_res in ["1", None]:
        if no_res:
            monkeypatch.setenv("SB_NO_RESIDENT", no_res)
        else:
            monkeypatch.delenv("SB_NO_RESIDENT", raising=False)
        with sb.Engine(precision=precision, resample="systematic", seed=21) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            lz = eng.smc_sweep(N, it=3, history=True)
            launches = eng.stats().kernel_launches
            rows = [eng.history(t, N) for t in range(7)]
            lz2 = eng.smc_sweep(N, it=4)
            parts = eng.particles(N)
            marg, _ = eng.pmcmc_counts(3, N, want_joint=False)
            outs.append((lz, launches, rows, lz2, parts, marg, eng.retained(), eng.stats().logZ_last))
    a, b = outs
    assert a[0] == b[0] and a[3] == b[3] and a[7] == b[7]
    assert b[1] == 1 and a[1] >= 14                                   # one launch instead of 2 per step
    for t in range(7):
        assert np.array_equal(a[2][t][0], b[2][t][0]), "states differ at step %d" % t
        assert np.array_equal(a[2][t][1], b[2][t][1]), "ancestors differ at step %d" % t
    assert np.array_equal(a[4], b[4])
    assert np.array_equal(a[5], b[5]) and np.array_equal(a[6], b[6])
