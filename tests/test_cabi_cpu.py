"""CPU-side checks of the drop-in boundary: the shared library builds, loads without a GPU, exports
every symbol include/stochy_b200.h declares, and refuses to compute without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sb():
    import stochy_jl_b200 as sb
    sb.build()
    return sb


def test_header_symbols_exported(sb):
    hdr = open(os.path.join(ROOT, "include", "stochy_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sb_handle"}
    L = C.CDLL(sb._native.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(sb._native.EXPORTS) <= declared


def test_version_and_defaults(sb):
    L = sb.lib()
    assert L.sb_version() == 100
    o = sb._native.Options()
    L.sb_options_default(C.byref(o))
    assert (o.precision, o.resample_mode, o.retained_pick, o.seed) == (sb._native.FP32, sb._native.SYSTEMATIC, -1, 1)


def test_no_cpu_fallback(sb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sb.StochyError) as e:
        sb.Engine()
    assert e.value.code == sb._native.ECUDA
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(sb.StochyError):
        sb.pmcmc(sb.hmm2_example(), 5, 5)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "stochy.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "stochy_oracle" not in txt, f
    # the oracle is test infrastructure: outside tests/ only bench.py (CPU legs) and __graft_entry__.py (build + smoke) use it
    for dp in [os.path.join(ROOT, "scripts"), os.path.join(ROOT, "julia"), os.path.join(ROOT, "include")]:
        for f in os.listdir(dp):
            if f.endswith((".py", ".sh", ".jl", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "stochy_oracle" not in txt, f
    assert "oracle" not in open(os.path.join(ROOT, "stochy_jl_b200.py")).read()


def test_host_api_rejects_general_programs(sb):
    # a closure (general CPS program) is not a fixed-structure descriptor
    import torch
    if torch.cuda.is_available():
        with pytest.raises(sb.StochyError) as e:
            sb.pmcmc(lambda: None, 1, 1)
        assert e.value.code == sb._native.EUNSUPPORTED


def test_discrete_result_type(sb):
    # test/erptests.jl:1-10
    import math
    hist = {0: 2, 1: 3, 2: 5}
    erp = sb.Discrete(hist)
    assert sorted(sb.support(erp)) == [0, 1, 2]
    for x in range(3):
        assert sb.score(erp, x) == pytest.approx(math.log(hist[x] / 10))
    with pytest.raises(sb.StochyError) as e:
        sb.Discrete({1: 0})
    assert "zero probability mass" in str(e.value)


def test_hmm_example_descriptor_matches_enumerate(sb, enumref):
    # the augmented-state table form of examples/hmm.jl has the same exact posterior (SURVEY §4 goldens)
    m = sb.hmm_example(3)
    Z, joint = enumref.hmm_enumerate_joint(m.A, m.logF, x0=m.x0, has_factor=m.has_factor)
    assert Z == pytest.approx(0.12916, abs=1e-12)
    post = {}
    for code, p in enumerate(joint):
        if p > 0:
            xs = tuple((code // 4 ** t) % 4 for t in range(3))
            v = m.value(xs)
            post[v] = post.get(v, 0.0) + p
    F, T = False, True
    assert post[(F, F, F)] == pytest.approx(0.8296918550634872, rel=1e-12)
    assert post[(T, F, F)] == pytest.approx(0.09218798389594302, rel=1e-12)
    assert post[(T, T, T)] == pytest.approx(0.0026556209352740765, rel=1e-12)


def test_product_hellinger_kat(sb):
    """test/erptests.jl:12-17 against the PRODUCT's hellingerdistance (src/erp.jl:100-111), not the oracle's."""
    import json, os
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))["hellinger"]
    p = sb.Categorical(kat["p"])
    q = sb.Discrete({int(k): v for k, v in kat["q_counts"].items()})
    assert sb.hellingerdistance(p, q) == pytest.approx(kat["value"], abs=kat["tol"])
    assert sb.hellingerdistance(q, q) == 0.0
    with pytest.raises(AssertionError):                     # @assert issubset(qsupp, psupp), src/erp.jl:103
        sb.hellingerdistance(p, sb.Discrete({4: 1}))


def test_product_kl_and_asserts(sb):
    """kl(p, q) = sum_x p(x) (log p(x) - log q(x)) over a COMMON support (src/erp.jl:113-118)."""
    import math
    p = sb.Discrete({"a": 1, "b": 3})
    q = sb.Discrete({"b": 1, "a": 1})
    want = 0.25 * math.log(0.25 / 0.5) + 0.75 * math.log(0.75 / 0.5)
    assert sb.kl(p, q) == pytest.approx(want, rel=1e-14)
    assert sb.kl(p, p) == 0.0
    c = sb.Categorical([0.5, 0.5])
    assert sb.kl(sb.Discrete({1: 1, 2: 3}), c) == pytest.approx(want, rel=1e-14)
    with pytest.raises(AssertionError):                     # @assert length(psupp) == length(qsupp)
        sb.kl(p, sb.Discrete({"a": 1}))
    with pytest.raises(AssertionError):                     # @assert x in qsupp
        sb.kl(p, sb.Discrete({"a": 1, "c": 1}))


def test_product_discrete_show_order(sb):
    """show(io, ::Discrete) (src/erp.jl:80-96): rows sorted by descending probability, keys padded to one width."""
    d = sb.Discrete({"x": 1, "long": 6, "yy": 3})
    rows = repr(d).split("\n")
    assert [r.split("|")[0].strip() for r in rows] == ["'long'", "'yy'", "'x'"]
    assert len({r.index("|") for r in rows}) == 1
    assert [float(r.split("|")[1]) for r in rows] == pytest.approx([0.6, 0.3, 0.1])


def test_no_global_load_hoisted_above_grid_dependency_wait(sb):
    """Kernels launched with programmatic stream serialisation read what their predecessor wrote only behind
    `griddepcontrol.wait` (SASS: ACQBULK).  ptxas once moved the tile scan's `ld.global.nc` record loads in front of it
    (wrong log-evidence beyond 2^25 particles; DESIGN.md section 4, profiles/r02_notes.md).  Static check of the built
    library: in front of ACQBULK the tile scans may load only the host-drawn offset, the linear-Gaussian step kernels and
    the pull kernels nothing, the discrete step kernels only their model tables (15 or 30 128-bit table loads + the scores)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", sb._native.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    pre, has, name = {}, set(), None
    seen = False
    for line in out.splitlines():
        if "Function :" in line:
            name, seen = line.split("Function :")[1].strip(), False
        elif name and "ACQBULK" in line:
            seen = True
            has.add(name)
        elif name and not seen and re.search(r"\bLDG\b|LDG\.", line):
            pre[name] = pre.get(name, 0) + 1
    assert any("k_tilescan" in k for k in has) and any("k_step_hmm" in k for k in has), "no ACQBULK found: check the parser"
    for k in sorted(has):
        n = pre.get(k, 0)
        if "k_tilescan" in k:
            assert n <= 1, (k, n)
        elif "k_step_hmm" in k:
            assert n <= 31, (k, n)
        else:
            assert n == 0, (k, n)
