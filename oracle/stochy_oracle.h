/*
 * stochy_oracle.h -- CPU ORACLE for the Stochy.jl SMC/PMCMC/MH hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libstochy_b200.so) never links, includes or calls anything in oracle/.
 *
 * It is a plain-C fp64 restatement of the reference algorithm (Julia 0.3, cannot run here:
 * no julia binary, REQUIRE:1 pins 0.3) following, line by line:
 *   src/rand.jl:1-15        inverse-CDF categorical draw (strict `<`, last-index fallback, NaN error)
 *   src/pmcmc.jl:48-66      multinomial resampling, plain exp weights, N+1 pool with retained particle
 *   src/pmcmc.jl:36-46,68-99 sweep ordering, retained = particle 1, count ALL N return values
 *   src/mh.jl:22-43,61-111  single-site MH and log_mh_ratio
 *   src/erp.jl:22-30,44-56  flip / randominteger conventions, Discrete normalisation
 * plus the documented EXTENSIONS the north star asks for (no reference line exists for them):
 *   the fixed-point "weight grid", systematic resampling on it, log-evidence estimate.
 *
 * PARITY PIN: the literal routines are pinned by the reference's own known-answer tests
 * (test/randtests.jl:1-2, test/inferencetests.jl:3-20, test/erptests.jl:1-17, test/memtests.jl:7-11)
 * through tests/test_oracle_*.py.  Samplers' random STREAMS are unpinned (Julia dSFMT +
 * Distributions.jl 0.6.3 are not in the tree): the oracle and the CUDA path share a Philox4x32-10
 * stream defined here instead, so GPU-vs-oracle parity is exact given that stream.
 */
#ifndef STOCHY_ORACLE_H
#define STOCHY_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_OK         0
#define OR_ENAN       1   /* "unexpected nan in rand"   src/rand.jl:15 */
#define OR_EZEROMASS  2   /* "Distribution has zero probability mass." src/erp.jl:48 */
#define OR_EINVAL     3

#define OR_TILE       2048
#define OR_QBITS      27

/* resample modes (same numbering as include/stochy_b200.h) */
#define OR_RESAMPLE_SYSTEMATIC   0
#define OR_RESAMPLE_MULTINOMIAL  1
#define OR_RESAMPLE_LITERAL      2

/* ---- Philox4x32-10 ---- */
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the stream both sides use: purpose 0=propagate 1=resample 2=init 3=retained-pick 4=mh */
double or_uniform53(uint64_t seed, uint64_t idx, uint32_t t, uint32_t iter, uint32_t purpose, int which);

/* ---- src/rand.jl:1-15 ---- */
int64_t or_rand_categorical(const double* ps, int64_t n, double r, int* err);

/* ---- src/pmcmc.jl:48-52 literal: ps = w/sum(w); m draws, each a linear scan ---- */
int or_resample_literal(const double* w, int64_t n, const double* r, int64_t m, int64_t* anc);
/* identical ancestors, O(n + m log n): sequential running sum stored, then searched */
int or_resample_literal_fast(const double* w, int64_t n, const double* r, int64_t m, int64_t* anc);

/* ---- weight grid (extension; see DESIGN.md "weight grid") ---- */
uint32_t or_q31_f64(double d);           /* floor(2^d * 2^OR_QBITS), d <= 0, deterministic IEEE-only */
uint32_t or_q31_f32(float d);
typedef struct {
    int64_t   n, n_tiles;
    int32_t*  e;        /* [n_tiles] tile exponents (OR_E_EMPTY if tile has no mass) */
    uint64_t* Q;        /* [n_tiles] tile sums at tile scale */
    uint64_t* P;        /* [n_tiles+1] exclusive prefix of rescaled tile sums */
    uint64_t* C;        /* [n] inclusive global CDF on the grid (exact shifts; multinomial) */
    uint64_t* cl;       /* [n] tile-local inclusive prefix of w32 (systematic rescales these) */
    int32_t   E, sh0;   /* sh0 = LS, the left shift applied before the per-tile right shift */
    uint64_t  total;
} or_grid;
#define OR_E_EMPTY (-0x40000000)
/* kind: 0 = linear fp64 weights, 1 = fp64 log-weights, 2 = fp32 log-weights (data points to float) */
int  or_grid_quantise(const void* data, int kind, int64_t n, uint32_t* w32, int32_t* e_out);
int  or_grid_from_w32(const uint32_t* w32, const int32_t* e, int64_t n, or_grid* g);
int  or_grid_build(const void* data, int kind, int64_t n, or_grid* g);
void or_grid_free(or_grid* g);
double or_grid_log_mean_weight(const or_grid* g);  /* log( sum w / n ) for kinds 1,2 */
int  or_grid_systematic(const or_grid* g, double u, int64_t m, int64_t* anc);
int  or_grid_multinomial(const or_grid* g, const double* r, int64_t m, int64_t* anc);
/* convenience: build + resample + free */
int  or_resample_systematic(const void* data, int kind, int64_t n, double u, int64_t m, int64_t* anc);
int  or_resample_multinomial(const void* data, int kind, int64_t n, const double* r, int64_t m, int64_t* anc);

/* ---- log-sum-exp (extension D2/H3) ---- */
double or_logsumexp(const double* lw, int64_t n);

/* ---- models ---- */
typedef struct {
    int32_t K, T;
    int32_t x0;               /* >=0: fixed initial state (examples/hmm2.jl:18); <0: x_1 ~ pi */
    const double* pi;         /* [K] */
    const double* A;          /* [K*K] row-major transition probabilities */
    const double* logF;       /* [T*K] factor score of state k at step t (src/pmcmc.jl:36-37) */
    const uint8_t* has_factor;/* [T] or NULL (= all 1): steps with no factor do not resample */
} or_hmm;

typedef struct {
    int32_t T;
    double a, q, c, r, m0, P0;
    const double* y;          /* [T] */
} or_lg;

typedef struct {
    int32_t resample_mode;    /* OR_RESAMPLE_* */
    int32_t fp32;             /* 1: fp32 log-weights / fp32 LG state */
    int32_t retained_pick;    /* 0: particle index 0 (src/pmcmc.jl:89); 1: uniform Philox pick (H4) */
    uint64_t seed;
} or_opts;

/* One SMC sweep (one PMCMC iteration, src/pmcmc.jl:83-94).
 * xstar: retained trajectory [T] or NULL.  Outputs (any may be NULL):
 *   x_hist [T*(N+1)], anc_hist [T*(N+1)] (row t = ancestors drawn at the resample after step t),
 *   lw_hist [T*(N+1)], logZ (sum over factor steps of log mean weight), xstar_out [T] */
int or_hmm_sweep(const or_hmm* m, const or_opts* o, int64_t N, uint32_t iter, const int32_t* xstar,
                 int32_t* x_hist, int32_t* anc_hist, double* lw_hist, double* logZ, int32_t* xstar_out);

/* pmcmc(comp, numiterations, numparticles) src/pmcmc.jl:78-99 for the discrete HMM.
 * marg [T*K] += per-time value counts over ALL N particles every iteration (D10);
 * joint [K^T] (or NULL) += joint trajectory counts (little-endian base-K code, digit t = x_t). */
int or_hmm_pmcmc(const or_hmm* m, const or_opts* o, int64_t iters, int64_t N,
                 int64_t* marg, int64_t* joint, double* logZ_first);

int or_lg_sweep(const or_lg* m, const or_opts* o, int64_t N, uint32_t iter,
                double* x_last, double* logZ, double* lw_last);
/* one propagate+weight step from given parents (stepwise parity): */
void or_lg_step(const or_lg* m, const or_opts* o, int64_t N, uint32_t t, uint32_t iter,
                const double* x_prev_gathered, double* x_out, double* lw_out);

/* ---- op-list program: one real latent per particle, per step a draw statement and an observe statement
 * (src/pmcmc.jl:32-46 for sample/factor; observe(erp, xs...) = factor(sum of scores) src/Stochy.jl:55-56;
 * ERPs src/erp.jl:1-30 = Distributions.jl's: logpdf forms textbook, unpinned by the reference's tests).
 * Codes and parameter meaning as sb_op_step in include/stochy_b200.h. ---- */
typedef struct {
    int32_t draw, score;
    double  dp[4], sp[4];
    int32_t obs_off, obs_n;
    int32_t par_off, par_n;   /* Dirichlet's alpha: obs[par_off .. par_off + par_n) */
} or_op_step;
typedef struct {
    int32_t T;
    int32_t D;                /* latent dimension: 1 (scalar programs) or the Dirichlet's (<= 8); x arrays are D planes of N */
    const or_op_step* steps;  /* [T] */
    const double* obs;
} or_ops;
/* one propagate+weight step from given parents */
void or_ops_step(const or_ops* m, const or_opts* o, int64_t N, uint32_t t, uint32_t iter,
                 const double* x_prev_gathered, double* x_out, double* lw_out);
int or_ops_sweep(const or_ops* m, const or_opts* o, int64_t N, uint32_t iter,
                 double* x_last, double* logZ, double* lw_last);

/* ---- Bayes net single-site MH (src/mh.jl:47-111) ---- */
typedef struct {
    int32_t D;                /* sampled binary sites, topological order */
    int32_t F;                /* factor statements */
    const uint32_t* site_parents; /* [D] bitmask of earlier sites */
    const double*   site_cpt;     /* [D*cpt_stride] P(site=1 | parent bits packed in site order) */
    const uint32_t* fac_parents;  /* [F] */
    const double*   fac_logf;     /* [F*cpt_stride] */
    int32_t cpt_stride;
    uint32_t query_mask;      /* value = bits of the sites in the mask, packed */
} or_bnet;
/* counts[2^popcount(query_mask)] += per-step values over all chains (src/mh.jl:87) */
int or_bnet_mh(const or_bnet* b, uint64_t seed, int64_t chains, int64_t steps, int64_t chain0,
               int64_t* counts, int64_t* accepts);
double or_bnet_score(const or_bnet* b, uint32_t state, double* choicescores);

int or_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
