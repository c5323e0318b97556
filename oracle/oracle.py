"""ctypes loader for the CPU oracle (oracle/stochy_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libstochy_oracle.so")

OK, ENAN, EZEROMASS, EINVAL = 0, 1, 2, 3
SYSTEMATIC, MULTINOMIAL, LITERAL = 0, 1, 2
TILE = 2048


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("stochy_oracle.c", "stochy_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class OrHmm(C.Structure):
    _fields_ = [("K", C.c_int32), ("T", C.c_int32), ("x0", C.c_int32),
                ("pi", C.c_void_p), ("A", C.c_void_p), ("logF", C.c_void_p), ("has_factor", C.c_void_p)]


class OrLg(C.Structure):
    _fields_ = [("T", C.c_int32), ("a", C.c_double), ("q", C.c_double), ("c", C.c_double),
                ("r", C.c_double), ("m0", C.c_double), ("P0", C.c_double), ("y", C.c_void_p)]


class OrOpts(C.Structure):
    _fields_ = [("resample_mode", C.c_int32), ("fp32", C.c_int32), ("retained_pick", C.c_int32),
                ("seed", C.c_uint64)]


class OrBnet(C.Structure):
    _fields_ = [("D", C.c_int32), ("F", C.c_int32), ("site_parents", C.c_void_p), ("site_cpt", C.c_void_p),
                ("fac_parents", C.c_void_p), ("fac_logf", C.c_void_p), ("cpt_stride", C.c_int32),
                ("query_mask", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.or_uniform53.restype = C.c_double
        L.or_uniform53.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.or_rand_categorical.restype = C.c_int64
        L.or_rand_categorical.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.POINTER(C.c_int)]
        L.or_q31_f64.restype = C.c_uint32
        L.or_q31_f64.argtypes = [C.c_double]
        L.or_q31_f32.restype = C.c_uint32
        L.or_q31_f32.argtypes = [C.c_float]
        L.or_logsumexp.restype = C.c_double
        L.or_logsumexp.argtypes = [C.c_void_p, C.c_int64]
        L.or_bnet_score.restype = C.c_double
        L.or_bnet_score.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.or_resample_systematic.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_int64, C.c_void_p]
        L.or_resample_multinomial.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.or_resample_literal.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.or_resample_literal_fast.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.or_hmm_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_hmm_pmcmc.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_lg_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_lg_step.restype = None
        L.or_lg_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_uint32, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
        L.or_ops_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.or_ops_step.restype = None
        L.or_ops_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_uint32, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
        L.or_bnet_mh.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.or_philox4x32_10.restype = None
        L.or_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, rc):
        self.rc = rc
        msg = {ENAN: "unexpected nan in rand", EZEROMASS: "Distribution has zero probability mass.",
               EINVAL: "invalid argument"}.get(rc, "rc=%d" % rc)
        super().__init__(msg)


def _chk(rc):
    if rc != OK:
        raise OracleError(rc)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def philox(ctr, key):
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().or_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def uniform53(seed, idx, t, it, purpose, which):
    return lib().or_uniform53(seed, idx, t, it, purpose, which)


def rand_categorical(ps, r):
    """src/rand.jl:1-15 (0-based result)."""
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    err = C.c_int(0)
    i = lib().or_rand_categorical(_p(ps), len(ps), float(r), C.byref(err))
    _chk(err.value)
    return int(i)


def resample_literal(w, r, fast=False):
    w = np.ascontiguousarray(w, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    anc = np.empty(len(r), dtype=np.int64)
    f = lib().or_resample_literal_fast if fast else lib().or_resample_literal
    _chk(f(_p(w), len(w), _p(r), len(r), _p(anc)))
    return anc


def _grid_data(data, kind):
    if kind == 2:
        return np.ascontiguousarray(data, dtype=np.float32)
    return np.ascontiguousarray(data, dtype=np.float64)


def resample_systematic(data, u, m, kind=0):
    """kind 0: linear fp64 weights; 1: fp64 log-weights; 2: fp32 log-weights."""
    d = _grid_data(data, kind)
    anc = np.empty(m, dtype=np.int64)
    _chk(lib().or_resample_systematic(_p(d), kind, len(d), float(u), m, _p(anc)))
    return anc


def resample_multinomial(data, r, kind=0):
    d = _grid_data(data, kind)
    r = np.ascontiguousarray(r, dtype=np.float64)
    anc = np.empty(len(r), dtype=np.int64)
    _chk(lib().or_resample_multinomial(_p(d), kind, len(d), _p(r), len(r), _p(anc)))
    return anc


def logsumexp(lw):
    lw = np.ascontiguousarray(lw, dtype=np.float64)
    return lib().or_logsumexp(_p(lw), len(lw))


class Hmm:
    def __init__(self, A, logF, x0=-1, pi=None, has_factor=None):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.K = self.A.shape[0]
        self.logF = np.ascontiguousarray(logF, dtype=np.float64)
        self.T = self.logF.shape[0]
        self.x0 = int(x0)
        self.pi = np.ascontiguousarray(pi if pi is not None else np.full(self.K, 1.0 / self.K), dtype=np.float64)
        self.has_factor = None if has_factor is None else np.ascontiguousarray(has_factor, dtype=np.uint8)
        self.c = OrHmm(self.K, self.T, self.x0, _p(self.pi), _p(self.A), _p(self.logF), _p(self.has_factor))


class Lg:
    def __init__(self, a, q, c, r, m0, P0, y):
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.T = len(self.y)
        self.a, self.q, self.cc, self.r, self.m0, self.P0 = a, q, c, r, m0, P0
        self.c = OrLg(self.T, a, q, c, r, m0, P0, _p(self.y))


def opts(resample_mode=SYSTEMATIC, fp32=False, retained_pick=None, seed=1):
    if retained_pick is None:
        retained_pick = 1 if resample_mode == SYSTEMATIC else 0
    return OrOpts(resample_mode, int(fp32), int(retained_pick), seed)


def hmm_sweep(m, o, N, it=0, xstar=None, want_hist=True):
    S = N + 1
    xh = np.empty((m.T, S), dtype=np.int32) if want_hist else None
    ah = np.empty((m.T, S), dtype=np.int32) if want_hist else None
    lh = np.empty((m.T, S), dtype=np.float64) if want_hist else None
    lz = C.c_double(0)
    xs_out = np.empty(m.T, dtype=np.int32)
    xs = None if xstar is None else np.ascontiguousarray(xstar, dtype=np.int32)
    _chk(lib().or_hmm_sweep(C.byref(m.c), C.byref(o), N, it, _p(xs), _p(xh), _p(ah), _p(lh),
                            C.byref(lz), _p(xs_out)))
    return dict(x=xh, anc=ah, lw=lh, logZ=lz.value, xstar=xs_out)


def hmm_pmcmc(m, o, iters, N, want_joint=True):
    marg = np.zeros((m.T, m.K), dtype=np.int64)
    joint = np.zeros(m.K ** m.T, dtype=np.int64) if want_joint else None
    lz = C.c_double(0)
    _chk(lib().or_hmm_pmcmc(C.byref(m.c), C.byref(o), iters, N, _p(marg), _p(joint), C.byref(lz)))
    return dict(marg=marg, joint=joint, logZ_first=lz.value)


def lg_sweep(m, o, N, it=0):
    x = np.empty(N, dtype=np.float64)
    lw = np.empty(N, dtype=np.float64)
    lz = C.c_double(0)
    _chk(lib().or_lg_sweep(C.byref(m.c), C.byref(o), N, it, _p(x), C.byref(lz), _p(lw)))
    return dict(x=x, lw=lw, logZ=lz.value)


def lg_step(m, o, N, t, it, x_prev_gathered):
    xp = np.ascontiguousarray(x_prev_gathered, dtype=np.float64) if x_prev_gathered is not None else np.zeros(N)
    x = np.empty(N, dtype=np.float64)
    lw = np.empty(N, dtype=np.float64)
    lib().or_lg_step(C.byref(m.c), C.byref(o), N, t, it, _p(xp), _p(x), _p(lw))
    return x, lw


class OrOpStep(C.Structure):
    _fields_ = [("draw", C.c_int32), ("score", C.c_int32), ("dp", C.c_double * 4), ("sp", C.c_double * 4),
                ("obs_off", C.c_int32), ("obs_n", C.c_int32), ("par_off", C.c_int32), ("par_n", C.c_int32)]


class OrOps(C.Structure):
    _fields_ = [("T", C.c_int32), ("D", C.c_int32), ("steps", C.c_void_p), ("obs", C.c_void_p)]


class Ops:
    """steps: list of (draw, dp, score, sp, ys) with the integer codes of include/stochy_b200.h (SB_DRAW_*, SB_SCORE_*)."""

    def __init__(self, steps):
        self.T = len(steps)
        self.steps = (OrOpStep * self.T)()
        obs = []
        self.D = 1
        for t, (draw, dp, score, sp, ys) in enumerate(steps):
            st = self.steps[t]
            st.draw, st.score = int(draw), int(score)
            if int(draw) == 6:                                    # Dirichlet: dp = alpha (vector parameter, parked in obs[])
                st.par_off, st.par_n = len(obs), len(dp)
                obs.extend(float(v) for v in dp)
                self.D = max(self.D, len(dp))
            else:
                for i, v in enumerate(dp):
                    st.dp[i] = float(v)
            for i, v in enumerate(sp):
                st.sp[i] = float(v)
            st.obs_off, st.obs_n = len(obs), len(ys)
            obs.extend(float(y) for y in ys)
        self.obs = np.ascontiguousarray(obs if obs else [0.0], dtype=np.float64)
        self.c = OrOps(self.T, self.D, C.cast(self.steps, C.c_void_p), _p(self.obs))


def ops_sweep(m, o, N, it=0):
    x = np.empty(N * m.D, dtype=np.float64)
    lw = np.empty(N, dtype=np.float64)
    lz = C.c_double(0)
    _chk(lib().or_ops_sweep(C.byref(m.c), C.byref(o), N, it, _p(x), C.byref(lz), _p(lw)))
    return dict(x=x if m.D == 1 else x.reshape(m.D, N), lw=lw, logZ=lz.value)


def ops_step(m, o, N, t, it, x_prev_gathered):
    """x_prev_gathered: [N] (scalar programs) or [D, N] planes (Dirichlet programs); returns x of the same shape"""
    xp = np.ascontiguousarray(x_prev_gathered, dtype=np.float64) if x_prev_gathered is not None else np.zeros(N * m.D)
    x = np.empty(N * m.D, dtype=np.float64)
    lw = np.empty(N, dtype=np.float64)
    lib().or_ops_step(C.byref(m.c), C.byref(o), N, t, it, _p(xp), _p(x), _p(lw))
    return (x if m.D == 1 else x.reshape(m.D, N)), lw


class Bnet:
    def __init__(self, site_parents, site_cpt, fac_parents, fac_logf, query_mask):
        self.D = len(site_parents)
        self.F = len(fac_parents)
        stride = max([len(c) for c in site_cpt] + [len(f) for f in fac_logf] + [1])
        self.stride = stride
        self.site_parents = np.asarray(site_parents, dtype=np.uint32)
        self.fac_parents = np.asarray(fac_parents, dtype=np.uint32)
        self.site_cpt = np.zeros((self.D, stride), dtype=np.float64)
        for d, c in enumerate(site_cpt):
            self.site_cpt[d, :len(c)] = c
        self.fac_logf = np.zeros((max(self.F, 1), stride), dtype=np.float64)
        for f, c in enumerate(fac_logf):
            self.fac_logf[f, :len(c)] = c
        self.query_mask = int(query_mask)
        self.c = OrBnet(self.D, self.F, _p(self.site_parents), _p(self.site_cpt), _p(self.fac_parents),
                        _p(self.fac_logf), stride, self.query_mask)


def bnet_mh(b, seed, chains, steps, chain0=0):
    nq = bin(b.query_mask).count("1")
    counts = np.zeros(1 << nq, dtype=np.int64)
    acc = C.c_int64(0)
    _chk(lib().or_bnet_mh(C.byref(b.c), seed, chains, steps, chain0, _p(counts), C.byref(acc)))
    return counts, acc.value


def bnet_score(b, state):
    cs = np.zeros(32, dtype=np.float64)
    s = lib().or_bnet_score(C.byref(b.c), state, _p(cs))
    return s, cs[:b.D]
