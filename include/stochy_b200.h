/*
 * stochy_b200.h -- C ABI of libstochy_b200.so: the B200-native replacement for the data-parallel
 * hot path of Stochy.jl (the SMC sweep inside pmcmc, and batched single-site MH).
 *
 * Plain C: pointers, sizes, POD structs.  No exceptions, no torch / CUDA types in any signature.
 * Every function returns an int32 status (SB_OK == 0) unless stated otherwise; the text of the
 * last failure is available from sb_last_error().  All pointers are HOST pointers unless the
 * parameter name starts with `d_` (device pointer on the handle's device).  The caller owns every
 * buffer it passes; the library copies what it needs before returning and owns all device memory
 * behind the opaque handle.  Calls are synchronous (they return after the handle's stream has
 * drained).  A handle is not thread-safe; distinct handles may be used from distinct threads.
 *
 * Reference interfaces replaced (paths relative to the Stochy.jl tree):
 *   sb_pmcmc_run        <- pmcmc(store,k,address,comp,numiterations,numparticles)  src/pmcmc.jl:78-99
 *   sb_pmcmc_chains     <- many independent pmcmc(...) calls (one per chain)                       idem
 *   sb_smc_sweep        <- smc(store,k,a,comp,n) = pmcmc(...,1,n)                  src/pmcmc.jl:101-103
 *   sb_mh_run           <- mh(store,k,address,comp,numsteps)                       src/mh.jl:47-93
 *   sb_enum_hmm         <- enum(store,k,address,comp) for the discrete descriptor  src/enumerate.jl:57-85
 *   sb_model_*          <- the @pp model closure `comp` restricted to fixed per-step structure
 *                          (examples/hmm.jl:4-25, examples/hmm2.jl:5-16); sample/factor/observe
 *                          dispatch src/Stochy.jl:47-56, src/pmcmc.jl:32-46
 *   sb_resample_literal <- resample(particles, weights, n) + rand(ps)              src/pmcmc.jl:48-52, src/rand.jl:1-15
 *   sb_resample_multinomial / sb_resample_systematic / sb_logsumexp
 *                       <- same call site, north-star extensions (scalable, deterministic)
 *   sb_erp_sample / sb_erp_logpdf <- rand(e) / score(e,x)                          src/pmcmc.jl:33, src/Stochy.jl:8, src/erp.jl:1-30
 *   error codes         <- error("unexpected nan in rand") src/rand.jl:15;
 *                          error("Distribution has zero probability mass.") src/erp.jl:48
 */
#ifndef STOCHY_B200_H
#define STOCHY_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_VERSION 100

/* status codes */
#define SB_OK           0
#define SB_ENAN         1   /* NaN / +Inf weight: "unexpected nan in rand" (src/rand.jl:15) */
#define SB_EZEROMASS    2   /* all weights zero: "Distribution has zero probability mass." (src/erp.jl:48) */
#define SB_EINVAL       3
#define SB_EUNSUPPORTED 4   /* model is not of fixed per-step structure; there is NO CPU fallback */
#define SB_ECUDA        5
#define SB_EPEER        6   /* multi-GPU: a peer GPU never signalled (flag barrier timed out) */
#define SB_ENOMEM       7

/* precision of log-weights and of the linear-Gaussian state */
#define SB_FP32 0
#define SB_FP64 1

/* resampling */
#define SB_RESAMPLE_SYSTEMATIC  0  /* one uniform per step, deterministic integer CDF ("weight grid") */
#define SB_RESAMPLE_MULTINOMIAL 1  /* N i.i.d. uniforms on the same grid (src/pmcmc.jl:51 semantics) */
#define SB_RESAMPLE_LITERAL     2  /* reference numerics: plain exp, sequential fp64 CDF (src/rand.jl:2-13) */

/* weight kinds accepted by the resampling hooks */
#define SB_W_LINEAR_F64 0
#define SB_W_LOG_F64    1
#define SB_W_LOG_F32    2

/* ERP vocabulary (src/erp.jl:1-30) */
#define SB_ERP_FLIP           0  /* params: p            values {0,1}            */
#define SB_ERP_RANDOMINTEGER  1  /* params: k            values 1..k (1-based)   */
#define SB_ERP_NORMAL         2  /* params: mu, sigma                            */
#define SB_ERP_BETA           3  /* params: a, b                                 */
#define SB_ERP_GAMMA          4  /* params: shape, scale (building block of Dirichlet) */
#define SB_ERP_UNIFORM        5  /* params: lo, hi                               */

typedef struct sb_handle sb_handle;

typedef struct {
    int32_t  device;         /* CUDA device ordinal */
    int32_t  precision;      /* SB_FP32 | SB_FP64 */
    int32_t  resample_mode;  /* SB_RESAMPLE_* */
    int32_t  retained_pick;  /* -1 auto (index 0 for multinomial/literal as src/pmcmc.jl:89, uniform for systematic); 0; 1 */
    int32_t  store_logw;     /* 1: also keep per-step log-weights (debug / parity; costs W bytes per particle-step) */
    int32_t  use_graph;      /* reserved, must be 0.  Pools whose tiles are all resident at once run a sweep as ONE persistent
                              * cooperative launch instead (no option needed); a CUDA-graph replay of the tiled launch sequence
                              * would save nothing: the host enqueues faster than the GPU runs at every size */
    uint64_t seed;           /* Philox4x32-10 key */
} sb_options;

typedef struct {
    int64_t kernel_launches;   /* kernels launched by the last run call */
    double  device_ms;         /* CUDA-event time of the last run call's device work */
    double  step_kernel_ms;    /* device time spent in the fused step kernel (when timed) */
    int64_t particle_steps;    /* N * T * sweeps of the last run call */
    double  logZ_first;        /* log-evidence estimate of the first (unconditional) sweep */
    double  logZ_last;
    int64_t accepts;           /* MH */
} sb_stats;

int32_t     sb_version(void);
void        sb_options_default(sb_options* o);
int32_t     sb_create(const sb_options* o, sb_handle** out);
void        sb_destroy(sb_handle* h);
const char* sb_last_error(const sb_handle* h);   /* valid until the next call on h; h may be NULL */
int32_t     sb_get_stats(const sb_handle* h, sb_stats* out);
/* every > 0: bracket every `every`-th launch of the fused step kernel (t >= 1, at most 128 per run
 * call) with a CUDA-event pair on the handle's stream; sb_stats.step_kernel_ms is then the AVERAGE
 * duration of one launch in the last run call.  0 switches it off. */
int32_t     sb_set_profile(sb_handle* h, int32_t every);

/* ---- models (tables are copied; fixed per-step structure only) ---- */

/* Discrete state-space model in the examples/hmm2.jl form:
 *   x_1 ~ Categorical(A[x0,:]) if x0 >= 0 else Categorical(pi);  x_t ~ Categorical(A[x_{t-1},:]);
 *   factor(logF[t][x_t]) after each step t with has_factor[t] != 0 (NULL = every step).
 * Categorical draws follow src/rand.jl:2-13 on the row's sequential running sum. K <= 64. */
int32_t sb_model_discrete_hmm(sb_handle* h, int32_t K, int32_t T, int32_t x0, const double* pi,
                              const double* A, const double* logF, const uint8_t* has_factor);

/* x_1 ~ Normal(m0, sqrt(P0)); x_t ~ Normal(a x_{t-1}, sqrt(q)); observe(Normal(c x_t, sqrt(r)), y_t) */
int32_t sb_model_linear_gaussian(sb_handle* h, int32_t T, double a, double q, double c, double r,
                                 double m0, double P0, const double* y);

/* Fixed-structure ERP program with ONE real-valued latent per particle ("op list"): step t draws x_t from an ERP
 * whose parameters may depend (affinely) on x_{t-1}, then conditions on that step's observations with
 * observe(erp, ys...) = factor(sum of score(erp, y)) (src/Stochy.jl:55-56; ERPs: src/erp.jl:1-30 -- every
 * Distributions.jl distribution is an ERP there; these are the ones of the north star: flip, gaussian, beta, gamma,
 * uniform; poisson as a second observation family).  Replaces sample(...)/factor(...) of src/pmcmc.jl:32-46 for such
 * programs, e.g.  p = ~Beta(a, b); observe(Bernoulli(p), ys_1...); observe(Bernoulli(p), ys_2...); ...
 * A program is either scalar (all steps draw / keep ONE real latent) or a Dirichlet program: a latent VECTOR theta of
 * D <= 8 components drawn by SB_DRAW_DIRICHLET, kept by SB_DRAW_KEEP, scored by SB_SCORE_CATEGORICAL / SB_SCORE_ZERO
 * (theta = ~Dir(alpha); observe(Categorical(theta), ys...); ...); sb_get_particles then returns D planes of `count`.
 * Computed in fp64 whatever the handle's precision.  Rejection samplers (Beta, Gamma: Marsaglia-Tsang) draw from the
 * particle's own Philox counter stream (particle, draw number, step, iteration). */
#define SB_DRAW_KEEP      0   /* x_t = x_{t-1} (a statement-free step of the program between two observe calls) */
#define SB_DRAW_NORMAL    1   /* x_t ~ Normal(dp[0] x_{t-1} + dp[1], dp[2])   (standard deviation) */
#define SB_DRAW_BETA      2   /* x_t ~ Beta(dp[0], dp[1]) */
#define SB_DRAW_GAMMA     3   /* x_t ~ Gamma(shape dp[0], scale dp[1]) */
#define SB_DRAW_UNIFORM   4   /* x_t ~ Uniform(dp[0], dp[1]) */
#define SB_DRAW_FLIP      5   /* x_t = flip(dp[0] x_{t-1} + dp[1]) as 0.0 / 1.0 */
#define SB_DRAW_DIRICHLET 6   /* VECTOR latent: theta_t ~ Dirichlet(par[par_off .. par_off + par_n)), 2 <= par_n <= 8 components */
#define SB_SCORE_ZERO      0  /* factor(0.0): the step still resamples */
#define SB_SCORE_NORMAL    1  /* observe(Normal(sp[0] x_t + sp[1], sp[2]), ys...) */
#define SB_SCORE_BERNOULLI 2  /* observe(Bernoulli(sp[0] x_t + sp[1]), ys...), ys in {0, 1} */
#define SB_SCORE_POISSON   3  /* observe(Poisson(sp[0] x_t + sp[1]), ys...), ys non-negative integers */
#define SB_SCORE_CATEGORICAL 4 /* VECTOR latent: observe(Categorical(theta_t), ys...), ys in 1..D (1-based as the reference's randominteger) */
#define SB_OPS_MAXD 8
typedef struct sb_op_step {
    int32_t draw, score;
    double  dp[4], sp[4];
    int32_t obs_off, obs_n;   /* this step's observations: obs[obs_off .. obs_off + obs_n) */
    int32_t par_off, par_n;   /* vector-valued parameters of the draw (Dirichlet's alpha): obs[par_off .. par_off + par_n) */
} sb_op_step;
int32_t sb_model_ops(sb_handle* h, int32_t T, const sb_op_step* steps, const double* obs, int64_t n_obs);

/* Binary Bayes net for MH: D sampled flip sites in topological order, site d ~ flip(cpt[d][parents]),
 * F factor statements factor(logf[f][parents]); parents are bitmasks over sites, table index = the
 * parent bits packed in ascending site order; both tables have row stride cpt_stride.
 * The returned value of the program is the packed bits of query_mask. D <= 16. */
int32_t sb_model_bayesnet(sb_handle* h, int32_t D, int32_t F, const uint32_t* site_parents,
                          const double* site_cpt, const uint32_t* fac_parents, const double* fac_logf,
                          int32_t cpt_stride, uint32_t query_mask);

/* ---- inference entry points ---- */

/* smc(comp, n): one unconditional sweep with iteration index `iter` of the Philox stream.
 * out_logZ: sum over factor steps of log(mean weight).  History is kept only if the model is
 * discrete and a later sb_get_history / value query needs it (store_history != 0). */
int32_t sb_smc_sweep(sb_handle* h, int64_t N, uint32_t iter, int32_t store_history, double* out_logZ);

/* pmcmc(comp, numiterations, numparticles) for the discrete model (NOTE the reference's argument
 * order: iterations first, src/pmcmc.jl:78).  marg[T*K] and joint[K^T] (either may be NULL) are
 * ADDED to: per-time and joint trajectory counts over ALL N particles of every iteration
 * (src/pmcmc.jl:90-92).  joint uses little-endian base-K codes (digit t = x_t) and needs K^T <= 2^26. */
int32_t sb_pmcmc_run(sb_handle* h, int64_t numiterations, int64_t numparticles,
                     int64_t* marg, int64_t* joint);

/* A SET of independent PMCMC chains (global ids chain0 ..) of numparticles each, whose pool (numparticles + 1)
 * fits one tile of 2048: one CTA per chain runs the chain's whole iteration loop (src/pmcmc.jl:83-94) in one
 * kernel launch.  Chain c uses the Philox key seed + c * 0x9E3779B97F4A7C15; chain 0 reproduces sb_pmcmc_run.
 * marg / joint are ADDED to with the pooled counts of all chains.  sb_pmcmc_run itself takes this route for
 * numparticles <= 2047. */
int32_t sb_pmcmc_chains(sb_handle* h, int64_t chains, int64_t numiterations, int64_t numparticles, int64_t chain0,
                        int64_t* marg, int64_t* joint);

/* enum(comp) for the discrete model (src/enumerate.jl:57-85): exhaustive exploration.
 * joint[K^T] (or NULL): UNNORMALISED mass exp(score) of every complete execution path, little-endian base-K
 * trajectory codes -- the `returns` dictionary of enumerate.jl:66-72 before Discrete() normalises it; needs
 * K^T <= 2^26.  marg[T*K] (or NULL): the same posterior marginalised per time step, computed exactly in
 * O(T K^2) by forward-backward, for any T.  out_logZ = log of the total mass.  All-zero mass -> SB_EZEROMASS. */
int32_t sb_enum_hmm(sb_handle* h, double* joint, double* marg, double* out_logZ);

/* mh(comp, numsteps) batched: `chains` independent chains (global ids chain0..), counts[2^|query|]
 * is ADDED to with the value after every step of every chain (src/mh.jl:87). */
int32_t sb_mh_run(sb_handle* h, int64_t chains, int64_t steps, int64_t chain0, int64_t* counts);

/* ---- resampling / reduction hooks (host buffers; exact parity targets) ---- */
int32_t sb_resample_systematic(sb_handle* h, const void* weights, int32_t kind, int64_t n,
                               double u, int64_t m, int32_t* anc_out);
int32_t sb_resample_multinomial(sb_handle* h, const void* weights, int32_t kind, int64_t n,
                                const double* r, int64_t m, int32_t* anc_out);
int32_t sb_resample_literal(sb_handle* h, const double* w, int64_t n, const double* r, int64_t m,
                            int32_t* anc_out);
int32_t sb_logsumexp(sb_handle* h, const void* lw, int32_t kind /* SB_W_LOG_F64 | SB_W_LOG_F32 */,
                     int64_t n, double* out);
/* dst[i] = src[anc[i]] for elem_bytes in {4, 8} */
int32_t sb_gather(sb_handle* h, const void* src, int32_t elem_bytes, int64_t n, const int32_t* anc,
                  int64_t m, void* dst);

/* same operations on DEVICE-resident buffers (what the bench times as `value`);
 * d_tmp handling is internal.  sb_dev_* do not synchronise; call sb_sync. */
int32_t sb_dev_resample_systematic(sb_handle* h, const void* d_weights, int32_t kind, int64_t n,
                                   double u, int64_t m, int32_t* d_anc);
int32_t sb_dev_gather(sb_handle* h, const void* d_src, int32_t elem_bytes, int64_t n,
                      const int32_t* d_anc, int64_t m, void* d_dst);
int32_t sb_sync(sb_handle* h);
/* event-timed region on the handle's stream: sb_timer_start ... sb_timer_stop -> milliseconds */
int32_t sb_timer_start(sb_handle* h);
int32_t sb_timer_stop(sb_handle* h, double* ms);

/* ---- ERP vocabulary hooks ---- */
int32_t sb_erp_sample(sb_handle* h, int32_t erp, const double* params, int64_t n, uint64_t seed, double* out);
int32_t sb_erp_logpdf(sb_handle* h, int32_t erp, const double* params, const double* x, int64_t n, double* out);
/* Dirichlet(alpha[k]): out[n*k] rows on the simplex */
int32_t sb_erp_sample_dirichlet(sb_handle* h, const double* alpha, int32_t k, int64_t n, uint64_t seed, double* out);
int32_t sb_erp_logpdf_dirichlet(sb_handle* h, const double* alpha, int32_t k, const double* x, int64_t n, double* out);

/* Discrete(x, p) as a model primitive (src/erp.jl:39-59): p[k] = values(dict), possibly un-normalised (divided by
 * z = sum(p) when z != 1, :46-49; z == 0 -> SB_EZEROMASS, :47).  rand(erp) = x[rand(p)] with the inverse-CDF scan of
 * src/rand.jl:2-13 (NaN -> SB_ENAN): out_index[n] are 1-based positions in the support x.  score(erp, x) =
 * log(dict[x] / z) (:56): index[n] 1-based, -inf outside 1..k. */
int32_t sb_erp_sample_discrete(sb_handle* h, const double* p, int32_t k, int64_t n, uint64_t seed, int32_t* out_index);
int32_t sb_erp_logpdf_discrete(sb_handle* h, const double* p, int32_t k, const int32_t* index, int64_t n, double* out);

/* ---- state access (debug / parity / results) ---- */
/* rows of the last sweep's history at step t: x (int32 for discrete, double for LG), anc (ancestors
 * drawn AFTER step t), lw (needs store_logw).  Any out pointer may be NULL.  `count` entries each. */
int32_t sb_get_history(sb_handle* h, int32_t t, int64_t count, void* x_out, int32_t* anc_out, double* lw_out);
int32_t sb_get_retained(sb_handle* h, int32_t* xstar_out /* [T] */);
/* final particle states of the last sweep (LG: double[N]; discrete: int32[N]) */
int32_t sb_get_particles(sb_handle* h, int64_t count, void* x_out);

/* ---- multi-GPU: ONE particle pool sharded over `world` GPUs, one process per GPU ----
 * sb_dist_init, then sb_dist_ipc_export (allocates this rank's share of the pool, N_local particles, a
 * power of two, and returns its 64-byte CUDA-IPC handle), exchange the handles over the host's own
 * plumbing, sb_dist_ipc_import with all of them in rank order.  Afterwards sb_smc_sweep(h, N_local, ...)
 * runs a GLOBAL sweep over world x N_local particles on every rank (same seed on every rank): parent
 * weights / states on other GPUs are read over NVLink by the fused step kernel itself, tile records
 * are exchanged through peer memory behind a flag barrier.  out_logZ is identical on every rank.
 * All ranks must make the same sequence of sweep calls, and must not destroy their handle before every
 * rank has finished (barrier first).
 *
 * Conditional SMC / PMCMC on the sharded pool (src/pmcmc.jl:59-66, 83-94 across GPUs): call
 * sb_dist_reserve_history(h, T) between sb_dist_init and sb_dist_ipc_export (the state history rows then live in
 * the exported allocation: the parents of step t are read from row t-1 of the owning GPU), load the discrete model
 * (T steps at most), and call sb_pmcmc_run(h, iters, world x N_local - 1, marg, NULL) on every rank: the pool of
 * numparticles + 1 entries fills the shards exactly, the retained particle is the last entry of the last GPU,
 * ancestor rows hold GLOBAL indices, the backward pass (counts of every particle's value, the next retained
 * trajectory) follows the lineages across GPUs with peer atomics behind the same flag barrier.  `marg` receives
 * THIS rank's share of the counts (the particles it owns at each step): the host sums the ranks.  Bit-identical
 * to the single-GPU sb_pmcmc_run of the same particle count and seed. */
int32_t sb_dist_init(sb_handle* h, int32_t rank, int32_t world);
int32_t sb_dist_reserve_history(sb_handle* h, int32_t rows);
int32_t sb_dist_ipc_export(sb_handle* h, int64_t N_local, uint8_t out_handle[64]);
int32_t sb_dist_ipc_import(sb_handle* h, const uint8_t* all_handles /* [world][64] */);

#ifdef __cplusplus
}
#endif
#endif
