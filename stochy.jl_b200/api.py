"""Host-side mirror of the Stochy.jl inference interface for the accelerated path.

Reference interface (paths relative to the Stochy.jl tree):
  pmcmc(comp, numiterations, numparticles)  src/pmcmc.jl:78-99   (iterations FIRST)
  smc(comp, n) = pmcmc(comp, 1, n)          src/pmcmc.jl:101-103
  mh(comp, numsteps=10)                     src/mh.jl:47-93
  Discrete / support / score                src/erp.jl:39-56
  hellingerdistance, kl                     src/erp.jl:100-121
Every entry point returns a Discrete over the program's return values, exactly like the
reference.  `comp` must be a fixed-structure model descriptor (DiscreteHMM, LinearGaussian,
BayesNet); a general closure is rejected with StochyError -- there is no CPU fallback.
Errors of the reference (`error("unexpected nan in rand")`, `error("Distribution has zero
probability mass.")`) surface as StochyError with the same text; the native handle is always
released, mirroring the reference's `try ... finally ctx = ctxold`.
"""
import ctypes as C
import math

import numpy as np

from . import _native as nat


class StochyError(RuntimeError):
    """Julia ErrorException analogue; .code holds the SB_* status."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ result type (src/erp.jl:39-56)
class Discrete:
    def __init__(self, d):
        self.x = list(d.keys())
        p = [float(v) for v in d.values()]
        z = sum(p)
        if z == 0:
            raise StochyError(nat.EZEROMASS, "Distribution has zero probability mass.")
        self.dict = dict(d)
        self.z = z
        self.p = [v / z for v in p] if z != 1.0 else p

    def support(self):
        return self.x

    def score(self, x):
        return math.log(self.dict[x] / self.z)

    def prob(self, x):
        return self.dict.get(x, 0) / self.z

    def __repr__(self):                      # src/erp.jl:80-96: descending probability
        rows = sorted(zip(self.p, map(repr, self.x)), reverse=True)
        w = max(len(k) for _, k in rows)
        return "\n".join("%s | %s" % (k.ljust(w), p) for p, k in rows)


class Categorical:
    """Categorical(p) as the reference sees it through Distributions.jl (src/erp.jl:1-16): support 1..k (1-based,
    like randominteger, src/erp.jl:28-30), score = logpdf.  Host-side value type for hellingerdistance / kl."""

    def __init__(self, p):
        self.p = [float(v) for v in p]

    def support(self):
        return list(range(1, len(self.p) + 1))

    def score(self, x):
        return math.log(self.p[x - 1]) if 1 <= x <= len(self.p) and self.p[x - 1] > 0 else -math.inf


def support(d):
    return d.support()


def score(d, x):
    return d.score(x)


def hellingerdistance(p, q):
    ps, qs = list(p.support()), list(q.support())
    if not set(qs) <= set(ps):
        raise AssertionError("issubset(qsupp, psupp)")
    acc = 0.0
    for x in ps:
        px = math.exp(p.score(x))
        qx = math.exp(q.score(x)) if x in qs else 0.0
        acc += (math.sqrt(px) - math.sqrt(qx)) ** 2
    return math.sqrt(acc / 2.0)


def kl(p, q):
    """kl(p, q) (src/erp.jl:113-118): both supports must hold the same values (the reference's two asserts)."""
    ps, qs = list(p.support()), list(q.support())
    if len(ps) != len(qs):
        raise AssertionError("length(psupp) == length(qsupp)")
    for x in ps:
        if x not in qs:
            raise AssertionError("x in qsupp")
    return sum(math.exp(p.score(x)) * (p.score(x) - q.score(x)) for x in ps)


# ------------------------------------------------------------------ model descriptors
class DiscreteHMM:
    """Fixed per-step structure in the examples/hmm2.jl form (see sb_model_discrete_hmm).

    value: callable mapping the tuple of sampled states (x_1..x_T) to the program's return value
    (default: the tuple itself).  examples/hmm2.jl returns the states with the fixed initial
    state prepended and as Bools -- see `hmm2_example`.
    """

    def __init__(self, A, logF, x0=-1, pi=None, has_factor=None, value=None):
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.K = int(self.A.shape[0])
        self.logF = np.ascontiguousarray(logF, dtype=np.float64)
        self.T = int(self.logF.shape[0])
        if self.A.shape != (self.K, self.K) or self.logF.shape != (self.T, self.K):
            raise StochyError(nat.EINVAL, "DiscreteHMM: A must be KxK and logF TxK")
        self.x0 = int(x0)
        self.pi = None if pi is None else np.ascontiguousarray(pi, dtype=np.float64)
        if self.x0 < 0 and self.pi is None:
            self.pi = np.full(self.K, 1.0 / self.K)
        self.has_factor = None if has_factor is None else np.ascontiguousarray(has_factor, dtype=np.uint8)
        self.value = value

    def load(self, eng):
        eng._chk(nat.lib().sb_model_discrete_hmm(eng.h, self.K, self.T, self.x0, _ptr(self.pi), _ptr(self.A),
                                                 _ptr(self.logF), _ptr(self.has_factor)))


class LinearGaussian:
    """x_1 ~ Normal(m0, sqrt(P0)); x_t ~ Normal(a x_{t-1}, sqrt(q)); observe(Normal(c x_t, sqrt(r)), y_t)."""

    def __init__(self, a, q, c, r, m0, P0, y):
        self.a, self.q, self.c, self.r, self.m0, self.P0 = float(a), float(q), float(c), float(r), float(m0), float(P0)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.T = len(self.y)

    def load(self, eng):
        eng._chk(nat.lib().sb_model_linear_gaussian(eng.h, self.T, self.a, self.q, self.c, self.r, self.m0,
                                                    self.P0, _ptr(self.y)))


class OpsProgram:
    """Fixed-structure ERP program with ONE real latent per particle (sb_model_ops): a list of steps, each
    `(draw, score)` where
      draw  = ("normal", a, b, sd)   x ~ Normal(a x_prev + b, sd)         | ("beta", a, b) | ("gamma", shape, scale)
              | ("uniform", lo, hi) | ("flip", a, b)  x = flip(a x_prev + b) | ("keep",)
              | ("dirichlet", alpha_1, ..., alpha_D)  VECTOR latent theta ~ Dirichlet(alpha), D <= 8
      score = ("normal", c, d, sd, ys)  observe(Normal(c x + d, sd), ys...) | ("bernoulli", c, d, ys)
              | ("poisson", c, d, ys) | ("zero",)  factor(0.0)
              | ("categorical", ys)  observe(Categorical(theta), ys...), ys in 1..D (vector latent)
    -- sample / observe of src/pmcmc.jl:32-46 and src/Stochy.jl:55-56 for programs such as
       p = ~Beta(2, 3); observe(Bernoulli(p), true, false, true); observe(Bernoulli(p), false, ...)."""
    DRAWS = {"keep": 0, "normal": 1, "beta": 2, "gamma": 3, "uniform": 4, "flip": 5, "dirichlet": 6}
    SCORES = {"zero": 0, "normal": 1, "bernoulli": 2, "poisson": 3, "categorical": 4}

    def __init__(self, steps):
        self.T = len(steps)
        self.K = 0
        self.D = 1
        self.steps = (nat.OpStep * self.T)()
        obs = []
        for t, (draw, score) in enumerate(steps):
            st = self.steps[t]
            st.draw = self.DRAWS[draw[0]]
            if draw[0] == "dirichlet":                            # vector parameter: parked in the observation array
                st.par_off, st.par_n = len(obs), len(draw) - 1
                obs.extend(float(v) for v in draw[1:])
                self.D = max(self.D, len(draw) - 1)
            else:
                for i, v in enumerate(draw[1:]):
                    st.dp[i] = float(v)
            st.score = self.SCORES[score[0]]
            ys = []
            if score[0] != "zero":
                *par, ys = score[1:]
                for i, v in enumerate(par):
                    st.sp[i] = float(v)
            st.obs_off, st.obs_n = len(obs), len(ys)
            obs.extend(float(y) for y in ys)
        self.obs = np.ascontiguousarray(obs if obs else [0.0], dtype=np.float64)
        self.n_obs = len(obs)

    def load(self, eng):
        eng._chk(nat.lib().sb_model_ops(eng.h, self.T, C.cast(self.steps, C.c_void_p), _ptr(self.obs), self.n_obs))


class BayesNet:
    """Binary Bayes net: sites[d] = (parent_mask, cpt) with cpt[i] = P(site=1 | packed parent bits i);
    factors[f] = (parent_mask, logf table).  query: iterable of site indices whose values are returned."""

    def __init__(self, sites, factors, query):
        self.D, self.F = len(sites), len(factors)
        self.stride = max([len(c) for _, c in sites] + [len(t) for _, t in factors] + [1])
        self.site_parents = np.array([m for m, _ in sites], dtype=np.uint32)
        self.site_cpt = np.zeros((self.D, self.stride))
        for d, (_, c) in enumerate(sites):
            self.site_cpt[d, :len(c)] = c
        self.fac_parents = np.array([m for m, _ in factors] or [0], dtype=np.uint32)
        self.fac_logf = np.zeros((max(self.F, 1), self.stride))
        for f, (_, t) in enumerate(factors):
            self.fac_logf[f, :len(t)] = t
        self.query = sorted(int(q) for q in query)
        self.query_mask = sum(1 << q for q in self.query)

    def load(self, eng):
        eng._chk(nat.lib().sb_model_bayesnet(eng.h, self.D, self.F, _ptr(self.site_parents), _ptr(self.site_cpt),
                                             _ptr(self.fac_parents), _ptr(self.fac_logf), self.stride,
                                             self.query_mask))


def hmm2_example(observations=(True, True, True, True)):
    """examples/hmm2.jl:5-24 as a descriptor: init false; state ~ flip(prev ? .9 : .1);
    factor(state == obs ? 0 : -1); returns the Bool list including the initial state."""
    A = [[0.9, 0.1], [0.1, 0.9]]
    logF = [[0.0 if bool(k) == bool(o) else -1.0 for k in (0, 1)] for o in observations]
    return DiscreteHMM(A, logF, x0=0, value=lambda xs: (False,) + tuple(bool(x) for x in xs))


def hmm_example(T=3):
    """examples/hmm.jl:4-32 with enum() -> pmcmc: hidden chain true -> flip(.7/.3), sampled
    observations flip(.9/.1), ONE terminal hard factor (all observations false).  Augmented state
    z = 2*ok + s (ok = all observations so far were false) keeps the per-step structure fixed."""
    K = 4
    A = np.zeros((K, K))
    for ok in (0, 1):
        for s in (0, 1):
            for s2 in (0, 1):
                ps = (0.7 if s else 0.3) if s2 else (0.3 if s else 0.7)
                po_false = 0.1 if s2 else 0.9
                A[2 * ok + s, 2 * (ok & 1) + s2] += ps * (po_false if ok else 0.0)
                A[2 * ok + s, 0 + s2] += ps * ((1 - po_false) if ok else 1.0)
    logF = np.zeros((T, K))
    logF[T - 1, 0:2] = -np.inf
    hf = np.zeros(T, dtype=np.uint8)
    hf[T - 1] = 1
    return DiscreteHMM(A, logF, x0=3, has_factor=hf, value=lambda xs: tuple(bool(x & 1) for x in xs))


# ------------------------------------------------------------------ engine (one native handle)
class Engine:
    def __init__(self, device=0, precision="fp32", resample="systematic", seed=1, store_logw=False,
                 retained_pick=-1):
        L = nat.lib()
        o = nat.Options()
        L.sb_options_default(C.byref(o))
        o.device = device
        o.precision = {"fp32": nat.FP32, "fp64": nat.FP64}[precision]
        o.resample_mode = {"systematic": nat.SYSTEMATIC, "multinomial": nat.MULTINOMIAL, "literal": nat.LITERAL,
                           "reference": nat.LITERAL}[resample]
        o.seed = seed
        o.store_logw = int(store_logw)
        o.retained_pick = retained_pick
        self.opts = o
        self.h = C.c_void_p()
        rc = L.sb_create(C.byref(o), C.byref(self.h))
        if rc != nat.OK:
            raise StochyError(rc, "sb_create failed: %s" % L.sb_last_error(None).decode())
        self.model = None

    def close(self):
        """Release the native handle.  A sharded handle first waits at a group barrier: a peer's last tile scan may
        still be reading this rank's arena, which sb_destroy frees (every rank must therefore call close())."""
        if self.h:
            grp = getattr(self, "_shard_group", False)
            if grp is not False:
                try:
                    from . import dist as sbdist
                    sbdist.barrier(grp)
                except Exception:
                    pass
                self._shard_group = False
            nat.lib().sb_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != nat.OK:
            raise StochyError(rc, nat.lib().sb_last_error(self.h).decode())

    def stats(self):
        s = nat.Stats()
        self._chk(nat.lib().sb_get_stats(self.h, C.byref(s)))
        return s

    def timer_start(self):
        self._chk(nat.lib().sb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double(0)
        self._chk(nat.lib().sb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def set_profile(self, every):
        self._chk(nat.lib().sb_set_profile(self.h, every))

    def load(self, model):
        if not hasattr(model, "load"):
            raise StochyError(nat.EUNSUPPORTED,
                              "model is not a fixed-structure descriptor (general CPS programs are not "
                              "supported on the GPU path and there is no CPU fallback)")
        model.load(self)
        self.model = model
        return self

    # ---- multi-GPU: ONE pool sharded over the ranks of a torch.distributed group (one process per GPU)
    def shard(self, N_local, group=None, history_rows=0):
        """Make this handle rank `dist.get_rank()`'s share of a global pool of world x N_local particles.
        torch.distributed is only the plumbing (the 64-byte CUDA-IPC handles travel through one
        all_gather); afterwards smc_sweep(N_local) sweeps the GLOBAL pool, with resampled particles
        read across GPUs by the step kernel itself (NVLink peer loads).
        history_rows = T reserves the state history inside the shared allocation: needed for conditional SMC /
        pmcmc_counts on the sharded pool (src/pmcmc.jl:59-66, 83-94 across GPUs)."""
        from . import dist as sbdist
        rank, world = sbdist.rank_world(group)
        self._chk(nat.lib().sb_dist_init(self.h, rank, world))
        if history_rows:
            self._chk(nat.lib().sb_dist_reserve_history(self.h, int(history_rows)))
        mine = (C.c_uint8 * 64)()
        self._chk(nat.lib().sb_dist_ipc_export(self.h, N_local, mine))
        allh = sbdist.exchange_handles(bytes(mine), group)          # rank order
        buf = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(allh))
        self._chk(nat.lib().sb_dist_ipc_import(self.h, buf))
        sbdist.barrier(group)                                       # every arena is mapped everywhere
        self.rank, self.world, self.N_local = rank, world, N_local
        self._shard_group = group                                   # close() meets the peers at a barrier first
        return self

    # ---- inference
    def smc_sweep(self, N, it=0, history=False):
        lz = C.c_double(0)
        self._chk(nat.lib().sb_smc_sweep(self.h, N, it, int(history), C.byref(lz)))
        return lz.value

    def pmcmc_counts(self, iters, N, want_joint=True):
        """src/pmcmc.jl:78-99.  On a sharded handle (shard(..., history_rows=T)): N = world x N_local - 1 GLOBAL
        particles, every rank calls it, and the ranks' shares of the counts are summed over the group."""
        m = self.model
        marg = np.zeros((m.T, m.K), dtype=np.int64)
        sharded = getattr(self, "_shard_group", False) is not False
        joint = np.zeros(m.K ** m.T, dtype=np.int64) if (want_joint and not sharded) else None
        self._chk(nat.lib().sb_pmcmc_run(self.h, iters, N, _ptr(marg), _ptr(joint)))
        if sharded:
            from . import dist as sbdist
            marg = sbdist.sum_over_ranks(marg, self._shard_group)
        return marg, joint

    def pmcmc_chains_counts(self, chains, iters, N, chain0=0, want_joint=True):
        """`chains` independent PMCMC chains (one CTA each, N <= 2047); pooled counts."""
        m = self.model
        marg = np.zeros((m.T, m.K), dtype=np.int64)
        joint = np.zeros(m.K ** m.T, dtype=np.int64) if want_joint else None
        self._chk(nat.lib().sb_pmcmc_chains(self.h, chains, iters, N, chain0, _ptr(marg), _ptr(joint)))
        return marg, joint

    def enum_hmm(self, want_joint=True):
        """Exhaustive enumeration of the loaded discrete model: (unnormalised joint masses or None,
        exact per-time marginals [T, K], log total mass)."""
        m = self.model
        joint = np.zeros(m.K ** m.T, dtype=np.float64) if want_joint else None
        marg = np.zeros((m.T, m.K), dtype=np.float64)
        lz = C.c_double(0)
        self._chk(nat.lib().sb_enum_hmm(self.h, _ptr(joint), _ptr(marg), C.byref(lz)))
        return joint, marg, lz.value

    def mh_counts(self, chains, steps, chain0=0):
        nv = 1 << bin(self.model.query_mask).count("1")
        counts = np.zeros(nv, dtype=np.int64)
        self._chk(nat.lib().sb_mh_run(self.h, chains, steps, chain0, _ptr(counts)))
        return counts

    # ---- hooks
    @staticmethod
    def _wdata(w, kind):
        return np.ascontiguousarray(w, dtype=np.float32 if kind == nat.W_LOG_F32 else np.float64)

    def resample_systematic(self, w, u, m, kind=nat.W_LINEAR_F64):
        d = self._wdata(w, kind)
        anc = np.empty(m, dtype=np.int32)
        self._chk(nat.lib().sb_resample_systematic(self.h, _ptr(d), kind, len(d), float(u), m, _ptr(anc)))
        return anc

    def resample_multinomial(self, w, r, kind=nat.W_LINEAR_F64):
        d = self._wdata(w, kind)
        r = np.ascontiguousarray(r, dtype=np.float64)
        anc = np.empty(len(r), dtype=np.int32)
        self._chk(nat.lib().sb_resample_multinomial(self.h, _ptr(d), kind, len(d), _ptr(r), len(r), _ptr(anc)))
        return anc

    def resample_literal(self, w, r):
        w = np.ascontiguousarray(w, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        anc = np.empty(len(r), dtype=np.int32)
        self._chk(nat.lib().sb_resample_literal(self.h, _ptr(w), len(w), _ptr(r), len(r), _ptr(anc)))
        return anc

    def logsumexp(self, lw, kind=nat.W_LOG_F64):
        d = self._wdata(lw, kind)
        out = C.c_double(0)
        self._chk(nat.lib().sb_logsumexp(self.h, _ptr(d), kind, len(d), C.byref(out)))
        return out.value

    def gather(self, src, anc):
        src = np.ascontiguousarray(src)
        anc = np.ascontiguousarray(anc, dtype=np.int32)
        dst = np.empty(len(anc), dtype=src.dtype)
        self._chk(nat.lib().sb_gather(self.h, _ptr(src), src.dtype.itemsize, len(src), _ptr(anc), len(anc), _ptr(dst)))
        return dst

    # ---- the same operations on DEVICE-resident buffers (raw device addresses, e.g. torch tensors' data_ptr());
    # asynchronous on the handle's stream: call sync() before reading the results
    def dev_resample_systematic(self, d_weights, kind, n, u, m, d_anc):
        self._chk(nat.lib().sb_dev_resample_systematic(self.h, C.c_void_p(d_weights), kind, n, float(u), m, C.c_void_p(d_anc)))

    def dev_gather(self, d_src, elem_bytes, n, d_anc, m, d_dst):
        self._chk(nat.lib().sb_dev_gather(self.h, C.c_void_p(d_src), elem_bytes, n, C.c_void_p(d_anc), m, C.c_void_p(d_dst)))

    def sync(self):
        self._chk(nat.lib().sb_sync(self.h))

    def erp_sample(self, erp, params, n, seed=1):
        p = np.zeros(2)
        p[:len(params)] = params
        out = np.empty(n)
        self._chk(nat.lib().sb_erp_sample(self.h, erp, _ptr(p), n, seed, _ptr(out)))
        return out

    def erp_logpdf(self, erp, params, x):
        p = np.zeros(2)
        p[:len(params)] = params
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x))
        self._chk(nat.lib().sb_erp_logpdf(self.h, erp, _ptr(p), _ptr(x), len(x), _ptr(out)))
        return out

    def dirichlet_sample(self, alpha, n, seed=1):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        out = np.empty((n, len(a)))
        self._chk(nat.lib().sb_erp_sample_dirichlet(self.h, _ptr(a), len(a), n, seed, _ptr(out)))
        return out

    def dirichlet_logpdf(self, alpha, x):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x))
        self._chk(nat.lib().sb_erp_logpdf_dirichlet(self.h, _ptr(a), len(a), _ptr(x), len(x), _ptr(out)))
        return out

    def discrete_sample(self, p, n, seed=1):
        """rand(Discrete(x, p)) (src/erp.jl:52): 1-based positions in the support, n draws."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.empty(n, dtype=np.int32)
        self._chk(nat.lib().sb_erp_sample_discrete(self.h, _ptr(p), len(p), n, seed, _ptr(out)))
        return out

    def discrete_logpdf(self, p, index):
        """score(Discrete(x, p), x[index]) = log(p[index] / sum(p)) (src/erp.jl:56), index 1-based."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        idx = np.ascontiguousarray(index, dtype=np.int32)
        out = np.empty(len(idx))
        self._chk(nat.lib().sb_erp_logpdf_discrete(self.h, _ptr(p), len(p), _ptr(idx), len(idx), _ptr(out)))
        return out

    # ---- state access
    def history(self, t, count, want_lw=False):
        x = np.empty(count, dtype=np.int32)
        a = np.empty(count, dtype=np.int32)
        lw = np.empty(count, dtype=np.float64) if want_lw else None
        self._chk(nat.lib().sb_get_history(self.h, t, count, _ptr(x), _ptr(a), _ptr(lw)))
        return x, a, lw

    def logw_row(self, t, count):
        lw = np.empty(count, dtype=np.float64)
        self._chk(nat.lib().sb_get_history(self.h, t, count, None, None, _ptr(lw)))
        return lw

    def retained(self):
        xs = np.empty(self.model.T, dtype=np.int32)
        self._chk(nat.lib().sb_get_retained(self.h, _ptr(xs)))
        return xs

    def particles(self, count, out=None):
        """Final particle states of the last sweep.  out: optional destination array (e.g. page-locked memory,
        which the device-to-host copy then reaches at full link speed)."""
        dt = np.int32 if isinstance(self.model, DiscreteHMM) else np.float64
        D = getattr(self.model, "D", 1)                             # Dirichlet programs: D planes of `count` (shape [D, count])
        x = np.empty(count * D, dtype=dt) if out is None else out
        if x.dtype != dt or x.size < count * D or not x.flags["C_CONTIGUOUS"]:
            raise StochyError(nat.EINVAL, "particles: out must be a contiguous %s array of at least %d entries" % (np.dtype(dt).name, count * D))
        self._chk(nat.lib().sb_get_particles(self.h, count, _ptr(x)))
        return x if D == 1 or out is not None else x.reshape(D, count)


# ------------------------------------------------------------------ the reference's entry points
def _decode_joint(model, joint):
    K, T = model.K, model.T
    out = {}
    for code in np.nonzero(joint)[0]:
        c = int(code)
        xs = tuple((c // K ** t) % K for t in range(T))
        v = model.value(xs) if model.value else xs
        out[v] = out.get(v, 0) + int(joint[code])
    return out


def pmcmc(comp, numiterations, numparticles, chains=1, **engine_kw):
    """pmcmc(comp, numiterations, numparticles) -> Discrete(counts)   (src/pmcmc.jl:78-99).
    Default numerics are the reference's (`resample="literal"`: plain exp weights, multinomial
    draws on the sequential fp64 CDF, N+1 pool with the retained particle, retained = particle 1).
    chains > 1: that many independent chains side by side (numparticles <= 2047), counts pooled."""
    engine_kw.setdefault("resample", "literal")
    engine_kw.setdefault("precision", "fp64")
    eng = Engine(**engine_kw)
    try:
        eng.load(comp)
        if not isinstance(comp, DiscreteHMM):
            raise StochyError(nat.EUNSUPPORTED, "pmcmc: the return value of this model has no finite encoding; "
                                                "use Engine.smc_sweep for its log-evidence")
        if chains > 1:
            _, joint = eng.pmcmc_chains_counts(chains, numiterations, numparticles, want_joint=True)
        else:
            _, joint = eng.pmcmc_counts(numiterations, numparticles, want_joint=True)
        return Discrete(_decode_joint(comp, joint))
    finally:
        eng.close()                      # `finally ctx = ctxold` (src/pmcmc.jl:95-97)


def enum(comp, **engine_kw):
    """enum(comp) -> Discrete(returns)   (src/enumerate.jl:57-85): exact posterior over the program's return
    values by exhaustive exploration, for the fixed-structure discrete descriptor (K^T <= 2^26 paths)."""
    engine_kw.setdefault("precision", "fp64")
    eng = Engine(**engine_kw)
    try:
        eng.load(comp)
        if not isinstance(comp, DiscreteHMM):
            raise StochyError(nat.EUNSUPPORTED, "enum: only the fixed-structure discrete model is enumerated on the GPU path")
        joint, _, _ = eng.enum_hmm(want_joint=True)
        K, T = comp.K, comp.T
        out = {}
        for code in np.nonzero(joint)[0]:
            c = int(code)
            xs = tuple((c // K ** t) % K for t in range(T))
            v = comp.value(xs) if comp.value else xs
            out[v] = out.get(v, 0.0) + float(joint[code])             # returns[value] += exp(score)  (enumerate.jl:72)
        return Discrete(out)
    finally:
        eng.close()


def smc(comp, n, **engine_kw):
    """smc(comp, n) = pmcmc(comp, 1, n)   (src/pmcmc.jl:101-103)."""
    return pmcmc(comp, 1, n, **engine_kw)


def mh(comp, numsteps=10, chains=1, **engine_kw):
    """mh(comp, numsteps) -> Discrete(counts) (src/mh.jl:47-93), batched over `chains` independent
    chains whose per-step values are pooled."""
    engine_kw.setdefault("precision", "fp64")
    eng = Engine(**engine_kw)
    try:
        eng.load(comp)
        if not isinstance(comp, BayesNet):
            raise StochyError(nat.EUNSUPPORTED, "mh: only fixed-structure binary Bayes nets run on the GPU path")
        counts = eng.mh_counts(chains, numsteps)
        nq = len(comp.query)
        out = {}
        for v in np.nonzero(counts)[0]:
            bits = tuple(bool((int(v) >> i) & 1) for i in range(nq))
            out[bits[0] if nq == 1 else bits] = int(counts[v])
        return Discrete(out)
    finally:
        eng.close()
