"""Committed golden fixtures (tests/golden/, generator: make_golden.py).

CPU: the oracle reproduces the reference's own known answers (reference_kats.json: numbers copied from the
reference's tests, file:line in each entry) and its own frozen vectors (oracle_vectors.json).
GPU: the CUDA path, through the C ABI, reproduces the same frozen vectors bit for bit."""
import json
import math
import os

import numpy as np
import pytest

from gpu_util import cfg2_tables, cfg5_bnet, need_gpu

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(HERE, "reference_kats.json")))
VEC = json.load(open(os.path.join(HERE, "oracle_vectors.json")))


def _bits(xs):
    return "".join("T" if x else "F" for x in xs)


# ------------------------------------------------------------------ CPU: oracle vs the reference's numbers
def test_kat_three_flips(enumref):
    def model(c):
        a, b, cc = c.flip(0.5), c.flip(0.5), c.flip(0.5)
        c.factor(0 if (a or b) else -1)
        return int(a) + int(b) + int(cc)
    d = enumref.enum(model)
    for v, p in KATS["enum_three_flips"]["p"].items():
        assert math.exp(d.logpdf(int(v))) == pytest.approx(p, abs=1e-14)


def test_kat_hellinger(enumref):
    k = KATS["hellinger"]
    p = enumref.Discrete({i + 1: v for i, v in enumerate(k["p"])})
    q = enumref.Discrete({int(a): b for a, b in k["q_counts"].items()})
    assert enumref.hellingerdistance(p, q) == pytest.approx(k["value"], abs=k["tol"])


def test_kat_errors(oracle):
    for inp in KATS["rand_nan"]["inputs"]:
        with pytest.raises(oracle.OracleError) as e:
            oracle.rand_categorical([math.nan] * len(inp), 0.5)
        assert KATS["rand_nan"]["error"] in str(e.value)


def test_kat_hmm2_host_descriptor(enumref):
    """examples/hmm2.jl through the host descriptor's tables == the hand enumeration."""
    import stochy_jl_b200 as sb
    m = sb.hmm2_example()
    Z, probs = enumref.hmm_enumerate_joint(m.A, m.logF, x0=m.x0)
    k = KATS["hmm2_example_enum"]
    assert Z == pytest.approx(k["Z"], rel=1e-12) and math.log(Z) == pytest.approx(k["logZ"], rel=1e-12)
    K = m.K
    for key, p in k["p"].items():
        xs = [1 if ch == "T" else 0 for ch in key[1:]]           # the fixed initial state is the first letter
        code = sum(x * K ** t for t, x in enumerate(xs))
        assert probs[code] == pytest.approx(p, rel=1e-12)


def test_kat_hmm_host_descriptor(enumref):
    """examples/hmm.jl (one terminal hard factor) through the augmented-state descriptor."""
    import stochy_jl_b200 as sb
    m = sb.hmm_example(T=3)
    Z, probs = enumref.hmm_enumerate_joint(m.A, m.logF, x0=m.x0, has_factor=m.has_factor)
    k = KATS["hmm_example_enum"]
    assert Z == pytest.approx(k["Z"], rel=1e-12)
    agg = {}
    for code in np.nonzero(probs)[0]:
        xs = tuple((int(code) // m.K ** t) % m.K for t in range(m.T))
        key = _bits(m.value(xs))
        agg[key] = agg.get(key, 0.0) + probs[code]
    for key, p in k["p"].items():
        assert agg[key] == pytest.approx(p, rel=1e-12)


def test_oracle_frozen_vectors(oracle):
    v = VEC["literal"]
    assert oracle.resample_literal(v["w"], v["r"]).tolist() == v["anc"]
    v = VEC["literal_dyadic"]
    assert oracle.resample_literal(v["w"], v["r"]).tolist() == v["anc"]
    assert oracle.resample_multinomial(v["w"], v["r"], kind=0).tolist() == v["anc"]      # grid == literal on dyadic weights
    v = VEC["systematic_log_f64"]
    assert oracle.resample_systematic(np.array(v["lw"]), v["u"], v["m"], kind=1).tolist() == v["anc"]
    v = VEC["multinomial_log_f64"]
    assert oracle.resample_multinomial(np.array(v["lw"]), v["r"], kind=1).tolist() == v["anc"]
    A, logF, pi = cfg2_tables()
    m = oracle.Hmm(A, logF, pi=pi)
    for mode, name in [(oracle.LITERAL, "literal"), (oracle.SYSTEMATIC, "systematic")]:
        g = VEC["pmcmc_cfg2_" + name]
        out = oracle.hmm_pmcmc(m, oracle.opts(mode, seed=g["seed"]), g["iters"], g["N"], want_joint=False)
        assert out["marg"].tolist() == g["marg"]
    sites, factors, query = cfg5_bnet()
    b = oracle.Bnet([s[0] for s in sites], [s[1] for s in sites], [f[0] for f in factors], [f[1] for f in factors],
                    sum(1 << q for q in query))
    g = VEC["mh_cfg5"]
    counts, acc = oracle.bnet_mh(b, g["seed"], g["chains"], g["steps"])
    assert counts.tolist() == g["counts"] and acc == g["accepts"]


# ------------------------------------------------------------------ GPU: CUDA path vs the frozen vectors
@pytest.mark.gpu
def test_gpu_frozen_resampling():
    need_gpu()
    import stochy_jl_b200 as sb
    with sb.Engine(precision="fp64") as eng:
        v = VEC["literal"]
        assert eng.resample_literal(v["w"], v["r"]).tolist() == v["anc"]
        v = VEC["literal_dyadic"]
        assert eng.resample_literal(v["w"], v["r"]).tolist() == v["anc"]
        assert eng.resample_multinomial(v["w"], v["r"]).tolist() == v["anc"]
        v = VEC["systematic_log_f64"]
        assert eng.resample_systematic(np.array(v["lw"]), v["u"], v["m"], kind=sb._native.W_LOG_F64).tolist() == v["anc"]
        v = VEC["multinomial_log_f64"]
        assert eng.resample_multinomial(np.array(v["lw"]), v["r"], kind=sb._native.W_LOG_F64).tolist() == v["anc"]


@pytest.mark.gpu
def test_gpu_frozen_pmcmc_and_mh():
    need_gpu()
    import stochy_jl_b200 as sb
    A, logF, pi = cfg2_tables()
    for name in ["literal", "systematic"]:
        g = VEC["pmcmc_cfg2_" + name]
        with sb.Engine(precision="fp64", resample=name, seed=g["seed"]) as eng:
            eng.load(sb.DiscreteHMM(A, logF, pi=pi))
            marg, _ = eng.pmcmc_counts(g["iters"], g["N"], want_joint=False)
            assert marg.tolist() == g["marg"]
            # literal mode: the oracle sums plain exp weights, the library reads the grid's tile scan
            assert eng.stats().logZ_first == pytest.approx(g["logZ_first"], rel=1e-9 if name != "literal" else 1e-6)
    sites, factors, query = cfg5_bnet()
    g = VEC["mh_cfg5"]
    with sb.Engine(precision="fp64", seed=g["seed"]) as eng:
        eng.load(sb.BayesNet(sites, factors, query))
        counts = eng.mh_counts(g["chains"], g["steps"])
        assert counts.tolist() == g["counts"] and eng.stats().accepts == g["accepts"]


@pytest.mark.gpu
def test_gpu_enumerate_matches_reference_kats(enumref):
    """enum() on the GPU for the examples/hmm.jl and hmm2.jl descriptors == the hand enumerations, and the
    exhaustive joint / forward-backward marginals == the CPU restatement of src/enumerate.jl on cfg2."""
    need_gpu()
    import stochy_jl_b200 as sb
    d = sb.enum(sb.hmm2_example())
    for key, p in KATS["hmm2_example_enum"]["p"].items():
        v = (False,) + tuple(ch == "T" for ch in key[1:])
        assert d.prob(v) == pytest.approx(p, rel=1e-12)
    assert len(d.support()) == 16
    d = sb.enum(sb.hmm_example(T=3))
    for key, p in KATS["hmm_example_enum"]["p"].items():
        assert d.prob(tuple(ch == "T" for ch in key)) == pytest.approx(p, rel=1e-12)
    A, logF, pi = cfg2_tables()
    Z, probs = enumref.hmm_enumerate_joint(A, logF, pi=pi)
    with sb.Engine(precision="fp64") as eng:
        eng.load(sb.DiscreteHMM(A, logF, pi=pi))
        joint, marg, lz = eng.enum_hmm()
    assert lz == pytest.approx(math.log(Z), rel=1e-12)
    assert np.allclose(joint / joint.sum(), probs, rtol=1e-11, atol=0)
    K, T = 3, 10
    codes = np.arange(K ** T)
    for t in range(T):
        ref = np.bincount((codes // K ** t) % K, weights=probs, minlength=K)
        assert np.allclose(marg[t], ref, rtol=1e-10)
    # zero mass: error("Distribution has zero probability mass.")  (test/inferencetests.jl:17-20)
    with sb.Engine(precision="fp64") as eng:
        eng.load(sb.DiscreteHMM([[0.5, 0.5], [0.5, 0.5]], [[-np.inf, -np.inf]], pi=[0.5, 0.5]))
        with pytest.raises(sb.StochyError) as e:
            eng.enum_hmm()
        assert KATS["zero_mass"]["error"] in str(e.value)
