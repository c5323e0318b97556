/*
 * stochy_oracle.c -- CPU ORACLE (test infrastructure; see stochy_oracle.h header comment).
 *
 * Plain C, fp64, compiled with -ffp-contract=off so that every a*b+c below is two IEEE
 * roundings unless fma() is written explicitly.  Reference citations are relative to
 * /root/reference/.
 */
#define _GNU_SOURCE
#include "stochy_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int or_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011, published constants).           */
/* The reference uses Julia's global dSFMT (src/rand.jl:4) which is   */
/* not in the tree: stream parity is unpinned.           */
/* ------------------------------------------------------------------ */
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void stream4(uint64_t seed, uint64_t idx, uint32_t t, uint32_t iter, uint32_t purpose,
                           uint32_t out[4]) {
    uint32_t ctr[4] = { (uint32_t)idx, (uint32_t)(idx >> 32), t, (iter << 4) | purpose };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    or_philox4x32_10(ctr, key, out);
}
static inline double u53_from(uint32_t a, uint32_t b) {
    uint64_t m = ((uint64_t)a << 21) | (uint64_t)(b >> 11);
    return (double)m * 0x1p-53;
}
static inline uint64_t r53_from(uint32_t a, uint32_t b) {
    return ((uint64_t)a << 21) | (uint64_t)(b >> 11);
}
double or_uniform53(uint64_t seed, uint64_t idx, uint32_t t, uint32_t iter, uint32_t purpose, int which) {
    uint32_t o[4];
    stream4(seed, idx, t, iter, purpose, o);
    return u53_from(o[2 * which], o[2 * which + 1]);
}
#define P_PROP 0u
#define P_RESAMPLE 1u
#define P_INIT 2u
#define P_PICK 3u
#define P_MH 4u

/* ------------------------------------------------------------------ */
/* src/rand.jl:1-15                                                   */
/* ------------------------------------------------------------------ */
int64_t or_rand_categorical(const double* ps, int64_t n, double r, int* err) {
    double acc = 0.0;                       /* rand.jl:3 */
    for (int64_t i = 0; i < n - 1; ++i) {   /* rand.jl:6  for i in 1:n-1 */
        acc += ps[i];                       /* rand.jl:7 */
        if (isnan(acc)) { *err = OR_ENAN; return -1; }   /* rand.jl:8 */
        if (r < acc) return i;              /* rand.jl:9 (strict) */
    }
    if (n == 1 && isnan(ps[0])) { *err = OR_ENAN; return -1; }  /* rand.jl:11 */
    return n - 1;                           /* rand.jl:12 (0-based here) */
}

/* src/pmcmc.jl:48-52.  `sum` association order is unpinned (Julia 0.3 Base, not in tree):
 * the oracle DEFINES it as the sequential left-to-right fp64 sum. */
int or_resample_literal(const double* w, int64_t n, const double* r, int64_t m, int64_t* anc) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += w[i];
    double* ps = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) ps[i] = w[i] / s;       /* pmcmc.jl:50 */
    int err = OR_OK;
    for (int64_t j = 0; j < m && err == OR_OK; ++j)         /* pmcmc.jl:51, one rand() per draw */
        anc[j] = or_rand_categorical(ps, n, r[j], &err);
    free(ps);
    return err;
}

int or_resample_literal_fast(const double* w, int64_t n, const double* r, int64_t m, int64_t* anc) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += w[i];
    double* acc = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double a = 0.0;
    int nan_at = -1;
    for (int64_t i = 0; i < n; ++i) {
        a += w[i] / s;
        acc[i] = a;
        if (isnan(a) && nan_at < 0 && i < n - 1) nan_at = 1;
    }
    if (n == 1 && isnan(w[0] / s)) { free(acc); return OR_ENAN; }
    if (nan_at >= 0) { free(acc); return OR_ENAN; }  /* a NaN partial sum is always reached */
    #pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m; ++j) {
        /* first i in [0,n-1) with r < acc[i]; acc is non-decreasing for non-negative weights */
        int64_t lo = 0, hi = n - 1;
        double rj = r[j];
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (rj < acc[mid]) hi = mid; else lo = mid + 1;
        }
        anc[j] = lo;
    }
    free(acc);
    return OR_OK;
}

/* ------------------------------------------------------------------ */
/* Weight grid (EXTENSION -- no reference line; DESIGN.md §"weight    */
/* grid").  All operations are IEEE basic ops / explicit fma so the   */
/* CUDA path reproduces them bit for bit.                             */
/* ------------------------------------------------------------------ */
static inline double pow2i(int k) {       /* 2^k for -1022 <= k <= 1023 */
    union { uint64_t u; double d; } v;
    v.u = (uint64_t)(1023 + k) << 52;
    return v.d;
}

uint32_t or_q31_f64(double d) {
    if (!(d > -(double)(OR_QBITS + 2))) return 0;   /* also catches -inf and NaN */
    double k = floor(d);
    double r = d - k;                     /* exact, in [0,1) */
    double t = (r - 0.5) * 0x1.62e42fefa39efp-1;   /* ln 2 */
    /* degree-11 Taylor of e^t, |t| <= 0.347, Horner with explicit fma */
    double p = 1.0 / 39916800.0;
    p = fma(p, t, 1.0 / 3628800.0);
    p = fma(p, t, 1.0 / 362880.0);
    p = fma(p, t, 1.0 / 40320.0);
    p = fma(p, t, 1.0 / 5040.0);
    p = fma(p, t, 1.0 / 720.0);
    p = fma(p, t, 1.0 / 120.0);
    p = fma(p, t, 1.0 / 24.0);
    p = fma(p, t, 1.0 / 6.0);
    p = fma(p, t, 0.5);
    p = fma(p, t, 1.0);
    p = fma(p, t, 1.0);
    p = p * 0x1.6a09e667f3bcdp+0;         /* sqrt 2 */
    return (uint32_t)(uint64_t)(p * pow2i(OR_QBITS + (int)k));
}

uint32_t or_q31_f32(float d) {
    if (!(d > -(float)(OR_QBITS + 2))) return 0;
    float k = floorf(d);
    float r = d - k;
    float t = (r - 0.5f) * 0x1.62e430p-1f;
    float p = 1.0f / 5040.0f;
    p = fmaf(p, t, 1.0f / 720.0f);
    p = fmaf(p, t, 1.0f / 120.0f);
    p = fmaf(p, t, 1.0f / 24.0f);
    p = fmaf(p, t, 1.0f / 6.0f);
    p = fmaf(p, t, 0.5f);
    p = fmaf(p, t, 1.0f);
    p = fmaf(p, t, 1.0f);
    p = p * 0x1.6a09e6p+0f;
    return (uint32_t)(uint64_t)((double)p * pow2i(OR_QBITS + (int)k));
}

static inline int ceil_log2_i64(int64_t v) {
    int l = 0;
    while (((int64_t)1 << l) < v) ++l;
    return l;
}

#define LOG2E_D 0x1.71547652b82fep+0
#define LOG2E_F 0x1.715476p+0f

/* Stage 1 of the grid: per tile of OR_TILE entries, a power-of-two exponent e_b >= every weight
 * and fixed-point weights w32_i = floor(w_i / 2^e_b * 2^OR_QBITS)  (OR_QBITS = 27). */
int or_grid_quantise(const void* data, int kind, int64_t n, uint32_t* w32, int32_t* e_out) {
    int64_t nt = (n + OR_TILE - 1) / OR_TILE;
    int bad = 0;
    const double* wd = (const double*)data;
    const float* wf = (const float*)data;
    #pragma omp parallel for schedule(static) reduction(|:bad)
    for (int64_t b = 0; b < nt; ++b) {
        int64_t i0 = b * OR_TILE, i1 = i0 + OR_TILE < n ? i0 + OR_TILE : n;
        int32_t e = OR_E_EMPTY;
        if (kind == 0) {
            for (int64_t i = i0; i < i1; ++i) {
                double w = wd[i];
                if (!(w >= 0.0) || isinf(w)) { bad = 1; continue; }
                if (w > 0.0) { int ex; frexp(w, &ex); if (ex > e) e = ex; }
            }
        } else if (kind == 1) {
            double mx = -INFINITY;
            for (int64_t i = i0; i < i1; ++i) {
                double l2 = wd[i] * LOG2E_D;
                if (isnan(l2) || l2 == INFINITY) { bad = 1; continue; }
                if (l2 > mx) mx = l2;
            }
            if (mx > -INFINITY) e = (int32_t)ceil(mx);
        } else {
            float mx = -INFINITY;
            for (int64_t i = i0; i < i1; ++i) {
                float l2 = wf[i] * LOG2E_F;
                if (isnan(l2) || l2 == INFINITY) { bad = 1; continue; }
                if (l2 > mx) mx = l2;
            }
            if (mx > -INFINITY) e = (int32_t)ceilf(mx);
        }
        e_out[b] = e;
        for (int64_t i = i0; i < i1; ++i) {
            uint32_t q = 0;
            if (e != OR_E_EMPTY && !bad) {
                if (kind == 0)      q = (uint32_t)(uint64_t)scalbn(wd[i], OR_QBITS - e);
                else if (kind == 1) q = or_q31_f64(wd[i] * LOG2E_D - (double)e);
                else                q = or_q31_f32(wf[i] * LOG2E_F - (float)e);
            }
            w32[i] = q;
        }
    }
    return bad ? OR_ENAN : OR_OK;
}

/* Stage 2: exact integer CDF.  Tile-local inclusive sums c_i (<= 2^38), global exponent
 * E = max e_b, left shift LS = 24 - ceil_log2(n_tiles), C_i = P_b + ((c_i << LS) >> (E - e_b)). */
int or_grid_from_w32(const uint32_t* w32, const int32_t* e, int64_t n, or_grid* g) {
    memset(g, 0, sizeof(*g));
    if (n <= 0) return OR_EINVAL;
    int64_t nt = (n + OR_TILE - 1) / OR_TILE;
    if (nt > ((int64_t)1 << 20)) return OR_EINVAL;
    g->n = n; g->n_tiles = nt;
    g->e = (int32_t*)malloc(sizeof(int32_t) * (size_t)nt);
    g->Q = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nt);
    g->P = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nt + 1));
    g->C = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    g->cl = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    memcpy(g->e, e, sizeof(int32_t) * (size_t)nt);
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nt; ++b) {
        int64_t i0 = b * OR_TILE, i1 = i0 + OR_TILE < n ? i0 + OR_TILE : n;
        uint64_t c = 0;
        for (int64_t i = i0; i < i1; ++i) { c += w32[i]; g->C[i] = c; g->cl[i] = c; }
        g->Q[b] = c;
    }
    int32_t E = OR_E_EMPTY;
    for (int64_t b = 0; b < nt; ++b) if (g->Q[b] != 0 && g->e[b] > E) E = g->e[b];
    if (E == OR_E_EMPTY) { or_grid_free(g); return OR_EZEROMASS; }
    int LS = 24 - ceil_log2_i64(nt);
    g->E = E; g->sh0 = LS;
    uint64_t acc = 0;
    for (int64_t b = 0; b < nt; ++b) {
        g->P[b] = acc;
        if (g->Q[b] != 0) {
            int64_t sh = (int64_t)E - g->e[b];
            acc += sh >= 63 ? 0 : ((g->Q[b] << LS) >> sh);
        }
    }
    g->P[nt] = acc; g->total = acc;
    if (acc == 0) { or_grid_free(g); return OR_EZEROMASS; }
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nt; ++b) {
        int64_t i0 = b * OR_TILE, i1 = i0 + OR_TILE < n ? i0 + OR_TILE : n;
        int64_t sh = g->Q[b] == 0 ? 63 : (int64_t)E - g->e[b];
        for (int64_t i = i0; i < i1; ++i)
            g->C[i] = g->P[b] + (sh >= 63 ? 0 : ((g->C[i] << LS) >> sh));
    }
    return OR_OK;
}

int or_grid_build(const void* data, int kind, int64_t n, or_grid* g) {
    memset(g, 0, sizeof(*g));
    if (n <= 0) return OR_EINVAL;
    int64_t nt = (n + OR_TILE - 1) / OR_TILE;
    uint32_t* w32 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    int32_t* e = (int32_t*)malloc(sizeof(int32_t) * (size_t)nt);
    int rc = or_grid_quantise(data, kind, n, w32, e);
    if (rc == OR_OK) rc = or_grid_from_w32(w32, e, n, g);
    free(w32); free(e);
    return rc;
}

void or_grid_free(or_grid* g) {
    free(g->e); free(g->Q); free(g->P); free(g->C); free(g->cl);
    memset(g, 0, sizeof(*g));
}

/* log(mean weight) = log(total) + (E - QBITS - LS) ln2 - log n   (log-weight kinds) */
double or_grid_log_mean_weight(const or_grid* g) {
    return log((double)g->total) + (double)(g->E - OR_QBITS - g->sh0) * 0x1.62e42fefa39efp-1
           - log((double)g->n);
}

/* ---- systematic resampling on the grid (EXTENSION; strict `<` and last-index fallback by
 * analogy with src/rand.jl:9,12).  Integer-only per particle:
 *   plan:  s = min(36, 62 - ceil_log2(m)) sub-child bits; the CDF is rescaled so that its total is
 *          ~ m * 2^s:  ratio = m * 2^(s+LS) / total = f * 2^x,  R = ceil(f * 2^26),  r0 = 26 - x;
 *   per particle (tile b, raw tile-local prefix c <= 2^38):  delta = (c * R) >> (r0 + E - e_b)
 *          (c * R < 2^64: one 64-bit multiply-accumulate per particle on the GPU);
 *   thresholds: child k sits at U + k * 2^s with U = floor(u * (Tt - (m-1) 2^s)), Tt = rescaled total;
 *   F(C) = #children strictly below C = C <= U ? 0 : min(m, ((C - U - 1) >> s) + 1).             */
typedef struct { uint32_t R; int r0, s; } sys_plan;

static int sys_make_plan(const or_grid* g, int64_t m, sys_plan* p) {
    int s = 62 - ceil_log2_i64(m);
    if (s > 36) s = 36;
    double ratio = ldexp((double)m, s + g->sh0) / (double)g->total;
    int x;
    double f = frexp(ratio, &x);
    uint64_t R = (uint64_t)ceil(f * 67108864.0);
    if (R >= 67108864ull) { R = 33554432ull; x += 1; }
    int r0 = 26 - x;
    if (r0 < 0) { s += r0; r0 = 0; }
    if (s < 0) return OR_EINVAL;
    p->R = (uint32_t)R; p->r0 = r0; p->s = s;
    return OR_OK;
}
static inline uint64_t sys_delta(uint64_t c, uint32_t R, int rb) {
    return rb >= 64 ? 0 : (c * (uint64_t)R) >> rb;
}
static inline int64_t sys_F(uint64_t C, uint64_t U, int s, int64_t m) {
    if (C <= U) return 0;
    uint64_t k = ((C - U - 1) >> s) + 1;
    return k >= (uint64_t)m ? m : (int64_t)k;
}

int or_grid_systematic(const or_grid* g, double u, int64_t m, int64_t* anc) {
    sys_plan p;
    int rc = sys_make_plan(g, m, &p);
    if (rc) return rc;
    const int64_t nt = g->n_tiles;
    uint64_t* Pt = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nt + 1));
    uint64_t acc = 0;
    for (int64_t b = 0; b < nt; ++b) {
        Pt[b] = acc;
        if (g->Q[b] != 0) acc += sys_delta(g->Q[b], p.R, p.r0 + (g->E - g->e[b]));
    }
    Pt[nt] = acc;
    const uint64_t Tt = acc, lastk = (uint64_t)(m - 1) << p.s;
    if (Tt <= lastk) { free(Pt); return OR_EZEROMASS; }
    const uint64_t span = Tt - lastk;
    const uint64_t R53 = (uint64_t)(u * 0x1p53);
    const uint64_t U = (uint64_t)(((unsigned __int128)R53 * span) >> 53);
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nt; ++b) {
        if (g->Q[b] == 0) continue;
        int64_t i0 = b * OR_TILE, i1 = i0 + OR_TILE < g->n ? i0 + OR_TILE : g->n;
        const int rb = p.r0 + (g->E - g->e[b]);
        int64_t lo = sys_F(Pt[b], U, p.s, m);
        for (int64_t i = i0; i < i1; ++i) {
            uint64_t C = Pt[b] + sys_delta(g->cl[i], p.R, rb);
            if (b == nt - 1 && i == i1 - 1) C = Tt;          /* (identical by construction) */
            int64_t hi = sys_F(C, U, p.s, m);
            for (int64_t c = lo; c < hi; ++c) anc[c] = i;
            lo = hi;
        }
    }
    free(Pt);
    return OR_OK;
}

int or_grid_multinomial(const or_grid* g, const double* r, int64_t m, int64_t* anc) {
    #pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m; ++j) {
        uint64_t R = (uint64_t)(r[j] * 0x1p53);                 /* r is a 53-bit uniform */
        uint64_t thr = (uint64_t)(((unsigned __int128)R * g->total) >> 53);
        int64_t lo = 0, hi = g->n - 1;                          /* first i with thr < C[i] */
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (thr < g->C[mid]) hi = mid; else lo = mid + 1;
        }
        anc[j] = lo;
    }
    return OR_OK;
}

int or_resample_systematic(const void* data, int kind, int64_t n, double u, int64_t m, int64_t* anc) {
    or_grid g; int rc = or_grid_build(data, kind, n, &g);
    if (rc) return rc;
    rc = or_grid_systematic(&g, u, m, anc);
    or_grid_free(&g);
    return rc;
}
int or_resample_multinomial(const void* data, int kind, int64_t n, const double* r, int64_t m, int64_t* anc) {
    or_grid g; int rc = or_grid_build(data, kind, n, &g);
    if (rc) return rc;
    rc = or_grid_multinomial(&g, r, m, anc);
    or_grid_free(&g);
    return rc;
}

double or_logsumexp(const double* lw, int64_t n) {
    double mx = -INFINITY;
    for (int64_t i = 0; i < n; ++i) if (lw[i] > mx) mx = lw[i];
    if (mx == -INFINITY) return -INFINITY;
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += exp(lw[i] - mx);
    return mx + log(s);
}

/* ------------------------------------------------------------------ */
/* Discrete HMM sweep: src/pmcmc.jl:32-46 (sample / factor),          */
/* :48-66 (resample), :68-76 (exit), examples/hmm2.jl:5-24 model form */
/* ------------------------------------------------------------------ */
static void cum_rows(const double* p, int rows, int K, double* cum) {
    for (int r = 0; r < rows; ++r) {
        double acc = 0.0;                       /* same running sum as rand.jl:3-7 */
        for (int k = 0; k < K; ++k) { acc += p[r * K + k]; cum[r * K + k] = acc; }
    }
}
/* u = R * 2^-32 with a 32-bit R: `u < cum[k]` (rand.jl:9) is then EXACTLY `R < ceil(cum[k] * 2^32)`;
 * thresholds are stored saturated to 32 bits (differs only for R = 2^32-1 against cum >= 1 - 2^-32). */
static inline uint32_t cat_threshold(double cum) {
    double t = ceil(cum * 4294967296.0);
    if (!(t > 0.0)) return 0u;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
static inline int32_t cat_draw(const double* cum, int K, uint32_t R) {
    for (int k = 0; k < K - 1; ++k) if (R < cat_threshold(cum[k])) return k;   /* rand.jl:9 */
    return K - 1;                                                              /* rand.jl:12 */
}
static inline uint32_t prop_u32(uint64_t seed, uint64_t i, uint32_t t, uint32_t iter) {
    uint32_t o[4];
    stream4(seed, i >> 2, t, iter, P_PROP, o);
    return o[i & 3];
}

static int resample_pool(const or_opts* o, const double* lw, int64_t n, int64_t N, uint32_t t, uint32_t iter,
                         int64_t* anc, double* logmean) {
    int rc;
    if (o->resample_mode == OR_RESAMPLE_LITERAL) {
        /* pmcmc.jl:55,64: weights = exp(score), no max subtraction (D2) */
        double* w = (double*)malloc(sizeof(double) * (size_t)n);
        double* r = (double*)malloc(sizeof(double) * (size_t)N);
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) { w[i] = exp(lw[i]); s += w[i]; }
        for (int64_t j = 0; j < N; ++j) r[j] = or_uniform53(o->seed, (uint64_t)j >> 1, t, iter, P_RESAMPLE, (int)(j & 1));
        rc = or_resample_literal_fast(w, n, r, N, anc);
        if (logmean) *logmean = log(s / (double)n);
        free(w); free(r);
        return rc;
    }
    or_grid g;
    if (o->fp32) {
        float* lf = (float*)malloc(sizeof(float) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) lf[i] = (float)lw[i];
        rc = or_grid_build(lf, 2, n, &g);
        free(lf);
    } else {
        rc = or_grid_build(lw, 1, n, &g);
    }
    if (rc) return rc;
    if (logmean) *logmean = or_grid_log_mean_weight(&g);
    if (o->resample_mode == OR_RESAMPLE_SYSTEMATIC) {
        double u = or_uniform53(o->seed, 0, t, iter, P_RESAMPLE, 0);
        rc = or_grid_systematic(&g, u, N, anc);
    } else {
        double* r = (double*)malloc(sizeof(double) * (size_t)N);
        for (int64_t j = 0; j < N; ++j) r[j] = or_uniform53(o->seed, (uint64_t)j >> 1, t, iter, P_RESAMPLE, (int)(j & 1));
        rc = or_grid_multinomial(&g, r, N, anc);
        free(r);
    }
    or_grid_free(&g);
    return rc;
}

int or_hmm_sweep(const or_hmm* m, const or_opts* o, int64_t N, uint32_t iter, const int32_t* xstar,
                 int32_t* x_hist, int32_t* anc_hist, double* lw_hist, double* logZ, int32_t* xstar_out) {
    const int K = m->K, T = m->T;
    const int64_t S = N + 1;                    /* row stride: slot N is the retained particle */
    int own_x = 0, own_a = 0;
    if (!x_hist)   { x_hist = (int32_t*)malloc(sizeof(int32_t) * (size_t)(S * T)); own_x = 1; }
    if (!anc_hist) { anc_hist = (int32_t*)malloc(sizeof(int32_t) * (size_t)(S * T)); own_a = 1; }
    double* cumA = (double*)malloc(sizeof(double) * (size_t)(K * K));
    double* cumpi = (double*)malloc(sizeof(double) * (size_t)K);
    cum_rows(m->A, K, K, cumA);
    if (m->x0 < 0) cum_rows(m->pi, 1, K, cumpi);
    double* lw = (double*)malloc(sizeof(double) * (size_t)S);
    int64_t* anc = (int64_t*)malloc(sizeof(int64_t) * (size_t)S);
    double lz = 0.0;
    int rc = OR_OK;
    for (int t = 0; t < T && rc == OR_OK; ++t) {
        int32_t* xt = x_hist + (int64_t)t * S;
        const int32_t* xp = t ? x_hist + (int64_t)(t - 1) * S : NULL;
        const int32_t* ap = t ? anc_hist + (int64_t)(t - 1) * S : NULL;
        /* pmcmc.jl:32-34: every particle draws from the prior; particle order is irrelevant
         * because the stream is counter-based (idx = particle pair, word = parity). */
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            uint32_t u = prop_u32(o->seed, (uint64_t)i, (uint32_t)t, iter);
            int32_t x;
            if (t == 0) x = m->x0 >= 0 ? cat_draw(cumA + m->x0 * K, K, u) : cat_draw(cumpi, K, u);
            else        x = cat_draw(cumA + xp[ap[i]] * K, K, u);
            xt[i] = x;
            lw[i] = m->logF[t * K + x];                     /* pmcmc.jl:37 Step(..., score) */
        }
        int64_t n = N;
        if (xstar) {                                        /* pmcmc.jl:59-66: pool of N+1 */
            xt[N] = xstar[t];
            lw[N] = m->logF[t * K + xstar[t]];              /* retained.path[t+1].score */
            n = N + 1;
        } else {
            xt[N] = -1; lw[N] = -INFINITY;
        }
        if (lw_hist) memcpy(lw_hist + (int64_t)t * S, lw, sizeof(double) * (size_t)S);
        int32_t* at = anc_hist + (int64_t)t * S;
        if (!m->has_factor || m->has_factor[t]) {
            double lm = 0.0;
            rc = resample_pool(o, lw, n, N, (uint32_t)t, iter, anc, &lm);   /* pmcmc.jl:42 */
            if (rc) break;
            lz += lm;
            for (int64_t i = 0; i < N; ++i) at[i] = (int32_t)anc[i];
        } else {
            for (int64_t i = 0; i < N; ++i) at[i] = (int32_t)i;
        }
        at[N] = (int32_t)N;                                 /* the retained slot descends from itself */
    }
    if (rc == OR_OK && xstar_out) {
        /* pmcmc.jl:89: retained = particles[1]; H4: uniform pick when ancestors are sorted */
        int64_t c = 0;
        if (o->retained_pick) {
            double u = or_uniform53(o->seed, 0, (uint32_t)T, iter, P_PICK, 0);
            c = (int64_t)(u * (double)N);
            if (c >= N) c = N - 1;
        }
        int64_t j = anc_hist[(int64_t)(T - 1) * S + c];
        for (int t = T - 1; t >= 0; --t) {
            xstar_out[t] = x_hist[(int64_t)t * S + j];
            if (t) j = anc_hist[(int64_t)(t - 1) * S + j];
        }
    }
    if (logZ) *logZ = lz;
    free(cumA); free(cumpi); free(lw); free(anc);
    if (own_x) free(x_hist);
    if (own_a) free(anc_hist);
    return rc;
}

int or_hmm_pmcmc(const or_hmm* m, const or_opts* o, int64_t iters, int64_t N,
                 int64_t* marg, int64_t* joint, double* logZ_first) {
    const int K = m->K, T = m->T;
    const int64_t S = N + 1;
    int32_t* xh = (int32_t*)malloc(sizeof(int32_t) * (size_t)(S * T));
    int32_t* ah = (int32_t*)malloc(sizeof(int32_t) * (size_t)(S * T));
    int32_t* xs = (int32_t*)malloc(sizeof(int32_t) * (size_t)T);
    int32_t* xs2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)T);
    int rc = OR_OK;
    for (int64_t it = 0; it < iters; ++it) {                /* pmcmc.jl:83 */
        double lz;
        rc = or_hmm_sweep(m, o, N, (uint32_t)it, it ? xs : NULL, xh, ah, NULL, &lz, xs2);
        if (rc) break;
        if (it == 0 && logZ_first) *logZ_first = lz;
        memcpy(xs, xs2, sizeof(int32_t) * (size_t)T);       /* pmcmc.jl:89 */
        for (int64_t c = 0; c < N; ++c) {                   /* pmcmc.jl:90-92: ALL N values (D10) */
            int64_t j = ah[(int64_t)(T - 1) * S + c];
            int64_t code = 0, mul = 1;
            /* digits little-endian in t; build from the back */
            int64_t pw = 1;
            for (int t = 0; t < T - 1 && joint; ++t) pw *= K;
            mul = pw;
            for (int t = T - 1; t >= 0; --t) {
                int32_t x = xh[(int64_t)t * S + j];
                if (marg) marg[t * K + x] += 1;
                if (joint) { code += mul * x; mul /= K; }
                if (t) j = ah[(int64_t)(t - 1) * S + j];
            }
            if (joint) joint[code] += 1;
        }
    }
    free(xh); free(ah); free(xs); free(xs2);
    return rc;
}

/* ------------------------------------------------------------------ */
/* Linear-Gaussian state space (Normal ERP, src/erp.jl:1-18 auto-wrap; */
/* observe = one factor of logpdf, src/Stochy.jl:55-56).               */
/* Normal sampler / logpdf: Distributions.jl 0.6.3 not in tree ->      */
/* textbook Box-Muller and closed-form logpdf; parity unpinned.        */
/* ------------------------------------------------------------------ */
#define TWO_PI 6.283185307179586476925286766559
/* Box-Muller, both branches used: particles 2q and 2q+1 share (u1, u2) and take r cos(theta), r sin(theta).
 * fp64: one Philox block per pair (two 53-bit uniforms); fp32: one block per FOUR particles (four 24-bit
 * uniforms).  Distributions.jl 0.6.3's rand(Normal) is not in the reference tree: stream unpinned. */
static inline double normal_f64(uint64_t seed, uint64_t i, uint32_t t, uint32_t iter) {
    uint32_t o[4];
    stream4(seed, i >> 1, t, iter, P_PROP, o);
    double u1 = u53_from(o[0], o[1]), u2 = u53_from(o[2], o[3]);
    double r = sqrt(-2.0 * log(1.0 - u1));
    return (i & 1) ? r * sin(TWO_PI * u2) : r * cos(TWO_PI * u2);
}
static inline float normal_f32(uint64_t seed, uint64_t i, uint32_t t, uint32_t iter) {
    uint32_t o[4];
    stream4(seed, i >> 2, t, iter, P_PROP, o);
    int w = (int)((i >> 1) & 1) * 2;
    float u1 = (float)(o[w] >> 8) * 0x1p-24f, u2 = (float)(o[w + 1] >> 8) * 0x1p-24f;
    float r = sqrtf(-2.0f * logf(1.0f - u1));
    return (i & 1) ? r * sinf((float)TWO_PI * u2) : r * cosf((float)TWO_PI * u2);
}

void or_lg_step(const or_lg* m, const or_opts* o, int64_t N, uint32_t t, uint32_t iter,
                const double* xpg, double* x_out, double* lw_out) {
    const double lc = -0.5 * log(TWO_PI * m->r);
    if (!o->fp32) {
        const double sq = sqrt(m->q), sp = sqrt(m->P0), hr = -0.5 / m->r;
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            double z = normal_f64(o->seed, (uint64_t)i, t, iter);
            double x = t == 0 ? m->m0 + sp * z : m->a * xpg[i] + sq * z;
            double d = m->y[t] - m->c * x;
            x_out[i] = x;
            lw_out[i] = hr * d * d + lc;                             /* hr = -0.5 / r: the product's operation sequence */
        }
    } else {
        const float sq = sqrtf((float)m->q), sp = sqrtf((float)m->P0);
        const float a = (float)m->a, c = (float)m->c, y = (float)m->y[t];
        const float hr = -0.5f / (float)m->r, lcf = (float)lc, m0 = (float)m->m0;
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            float z = normal_f32(o->seed, (uint64_t)i, t, iter);
            float x = t == 0 ? m0 + sp * z : a * (float)xpg[i] + sq * z;
            float d = y - c * x;
            x_out[i] = (double)x;
            lw_out[i] = (double)(hr * d * d + lcf);
        }
    }
}

int or_lg_sweep(const or_lg* m, const or_opts* o, int64_t N, uint32_t iter,
                double* x_last, double* logZ, double* lw_last) {
    double* x = (double*)malloc(sizeof(double) * (size_t)N);
    double* xg = (double*)malloc(sizeof(double) * (size_t)N);
    double* lw = (double*)malloc(sizeof(double) * (size_t)N);
    int64_t* anc = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
    double lz = 0.0; int rc = OR_OK;
    for (int t = 0; t < m->T; ++t) {
        if (t) {
            #pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; ++i) xg[i] = x[anc[i]];   /* pmcmc.jl:51 deepcopy(particles[a]) */
        }
        or_lg_step(m, o, N, (uint32_t)t, iter, xg, x, lw);
        double lm;
        rc = resample_pool(o, lw, N, N, (uint32_t)t, iter, anc, &lm);
        if (rc) break;
        lz += lm;
    }
    if (rc == OR_OK) {
        if (x_last) memcpy(x_last, x, sizeof(double) * (size_t)N);
        if (lw_last) memcpy(lw_last, lw, sizeof(double) * (size_t)N);
        if (logZ) *logZ = lz;
    }
    free(x); free(xg); free(lw); free(anc);
    return rc;
}

/* ------------------------------------------------------------------ */
/* Op-list programs: a real latent, per step one draw statement         */
/* (sample = rand(e), src/pmcmc.jl:32-34) and one observe statement     */
/* (factor of the summed scores, src/Stochy.jl:55-56, src/pmcmc.jl:36). */
/* Random stream of particle i at step t: Philox counter                */
/* (i | (draw pair) << 32, t, iter, PROP), two 53-bit uniforms a block. */
/* Distributions.jl 0.6.3's samplers are not in the tree: Beta = ratio  */
/* of Gammas, Gamma = Marsaglia-Tsang, Normal = Box-Muller (cos branch).*/
/* ------------------------------------------------------------------ */
typedef struct { uint64_t seed, idx; uint32_t t, iter, n; } ops_rng;
static inline double ops_next(ops_rng* g) {
    uint32_t o[4];
    stream4(g->seed, g->idx | ((uint64_t)(g->n >> 1) << 32), g->t, g->iter, P_PROP, o);
    double u = (g->n & 1) ? u53_from(o[2], o[3]) : u53_from(o[0], o[1]);
    ++g->n;
    return u;
}
static inline double ops_normal(ops_rng* g) {
    double u1 = ops_next(g), u2 = ops_next(g);
    return sqrt(-2.0 * log(1.0 - u1)) * cos(TWO_PI * u2);
}
static double ops_gamma(ops_rng* g, double shape) {
    double boost = 1.0;
    if (shape < 1.0) { boost = pow(1.0 - ops_next(g), 1.0 / shape); shape += 1.0; }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
        const double z = ops_normal(g);
        double v = 1.0 + c * z;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = 1.0 - ops_next(g);
        if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) return boost * d * v;
    }
    return boost * d;
}
/* score(erp, y): logpdf of Distributions.jl (textbook forms) */
static inline double lp_normal(double mu, double sd, double y) { double z = (y - mu) / sd; return (-0.5 * z * z - log(sd)) - 0.9189385332046727418; }
static inline double lp_bernoulli(double p, double y) { return y != 0.0 ? log(p) : log(1.0 - p); }
static inline double lp_poisson(double lam, double y) { return (y * log(lam) - lam) - lgamma(y + 1.0); }

void or_ops_step(const or_ops* m, const or_opts* o, int64_t N, uint32_t t, uint32_t iter,
                 const double* xpg, double* x_out, double* lw_out) {
    const or_op_step* s = &m->steps[t];
    const int D = m->D;
    if (D > 1) {
        /* Dirichlet program: theta = ~Dir(alpha) (K gammas normalised), kept by later steps,
         * observe(Categorical(theta), ys...) = factor(sum of log theta[y]) (src/Stochy.jl:55-56) */
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            ops_rng g = {o->seed, (uint64_t)i, t, iter, 0u};
            double th[8];
            if (s->draw == 6) {
                double sum = 0.0;
                for (int d = 0; d < D; ++d) { th[d] = ops_gamma(&g, m->obs[s->par_off + d]); sum += th[d]; }
                for (int d = 0; d < D; ++d) th[d] /= sum;
            } else {
                for (int d = 0; d < D; ++d) th[d] = (t == 0 || !xpg) ? 0.0 : xpg[(size_t)d * N + i];
            }
            double lw = 0.0;
            if (s->score == 4)
                for (int k = 0; k < s->obs_n; ++k) lw += log(th[(int)m->obs[s->obs_off + k] - 1]);
            for (int d = 0; d < D; ++d) x_out[(size_t)d * N + i] = th[d];
            lw_out[i] = lw;
        }
        return;
    }
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        ops_rng g = {o->seed, (uint64_t)i, t, iter, 0u};
        const double xprev = (t == 0 || !xpg) ? 0.0 : xpg[i];
        double x = xprev;
        switch (s->draw) {                                           /* sample: rand(e) */
            case 1: x = (s->dp[0] * xprev + s->dp[1]) + s->dp[2] * ops_normal(&g); break;
            case 2: { double X = ops_gamma(&g, s->dp[0]), Y = ops_gamma(&g, s->dp[1]); x = X / (X + Y); } break;
            case 3: x = s->dp[1] * ops_gamma(&g, s->dp[0]); break;
            case 4: x = s->dp[0] + (s->dp[1] - s->dp[0]) * ops_next(&g); break;
            case 5: x = ops_next(&g) < (s->dp[0] * xprev + s->dp[1]) ? 1.0 : 0.0; break;
            default: break;
        }
        const double par = s->sp[0] * x + s->sp[1];
        double lw = 0.0;                                             /* observe(erp, ys...): sum([score(erp, y) for y in ys]) */
        for (int k = 0; k < s->obs_n; ++k) {
            const double y = m->obs[s->obs_off + k];
            if (s->score == 1) lw += lp_normal(par, s->sp[2], y);
            else if (s->score == 2) lw += lp_bernoulli(par, y);
            else if (s->score == 3) lw += lp_poisson(par, y);
        }
        x_out[i] = x;
        lw_out[i] = lw;
    }
}

int or_ops_sweep(const or_ops* m, const or_opts* o, int64_t N, uint32_t iter,
                 double* x_last, double* logZ, double* lw_last) {
    const size_t D = (size_t)m->D;
    double* x = (double*)malloc(sizeof(double) * (size_t)N * D);
    double* xg = (double*)malloc(sizeof(double) * (size_t)N * D);
    double* lw = (double*)malloc(sizeof(double) * (size_t)N);
    int64_t* anc = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
    double lz = 0.0; int rc = OR_OK;
    or_opts o64 = *o; o64.fp32 = 0;                                  /* op-list programs are computed in fp64 */
    for (int t = 0; t < m->T; ++t) {
        if (t) {
            #pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; ++i)                          /* pmcmc.jl:51 deepcopy(particles[a]) */
                for (size_t d = 0; d < D; ++d) xg[d * (size_t)N + i] = x[d * (size_t)N + anc[i]];
        }
        or_ops_step(m, &o64, N, (uint32_t)t, iter, xg, x, lw);
        double lm;
        rc = resample_pool(&o64, lw, N, N, (uint32_t)t, iter, anc, &lm);
        if (rc) break;
        lz += lm;
    }
    if (rc == OR_OK) {
        if (x_last) memcpy(x_last, x, sizeof(double) * (size_t)N * D);
        if (lw_last) memcpy(lw_last, lw, sizeof(double) * (size_t)N);
        if (logZ) *logZ = lz;
    }
    free(x); free(xg); free(lw); free(anc);
    return rc;
}

/* ------------------------------------------------------------------ */
/* Bayes-net single-site MH: src/mh.jl:22-43 (sample/factor),          */
/* :47-93 (mh loop), :95-111 (log_mh_ratio) specialised to a           */
/* fixed-structure trace (every address present in both traces, only   */
/* site j fresh) -- SURVEY §3.2.                                       */
/* ------------------------------------------------------------------ */
static inline uint32_t pack_bits(uint32_t state, uint32_t mask) {
    uint32_t idx = 0, pos = 0;
    while (mask) {
        uint32_t b = mask & (~mask + 1u);
        if (state & b) idx |= 1u << pos;
        ++pos; mask ^= b;
    }
    return idx;
}
static inline double site_p1(const or_bnet* b, int d, uint32_t state) {
    return b->site_cpt[d * b->cpt_stride + pack_bits(state, b->site_parents[d])];
}
/* logpdf(Bernoulli(p), x): log(p) / log(1-p)  (Distributions 0.6.3; pinned only at p=.5,.7 by
 * test/inferencetests.jl:3-15, test/memtests.jl:7-11) */
static inline double bern_lp(double p, int x) { return x ? log(p) : log(1.0 - p); }

double or_bnet_score(const or_bnet* b, uint32_t state, double* cs) {
    double s = 0.0;
    for (int d = 0; d < b->D; ++d) {
        double c = bern_lp(site_p1(b, d, state), (state >> d) & 1);   /* mh.jl:33 choicescore */
        if (cs) cs[d] = c;
        s += c;                                                        /* mh.jl:36 */
    }
    for (int f = 0; f < b->F; ++f)
        s += b->fac_logf[f * b->cpt_stride + pack_bits(state, b->fac_parents[f])];   /* mh.jl:41 */
    return s;
}

int or_bnet_mh(const or_bnet* b, uint64_t seed, int64_t chains, int64_t steps, int64_t chain0,
               int64_t* counts, int64_t* accepts) {
    const int D = b->D;
    int nq = 0; for (uint32_t mk = b->query_mask; mk; mk &= mk - 1) ++nq;
    const int64_t nv = (int64_t)1 << nq;
    int64_t acc_total = 0;
    #pragma omp parallel
    {
        int64_t* loc = (int64_t*)calloc((size_t)nv, sizeof(int64_t));
        int64_t acc_loc = 0;
        #pragma omp for schedule(static)
        for (int64_t ch = 0; ch < chains; ++ch) {
            uint64_t gch = (uint64_t)(chain0 + ch);
            /* mh.jl:57 initialise: run the program once from the prior */
            uint32_t state = 0;
            for (int d = 0; d < D; ++d) {
                uint32_t o[4];
                stream4(seed, gch, (uint32_t)d, 0, P_INIT, o);
                double u = (double)o[0] * 0x1p-32;
                if (u < site_p1(b, d, state)) state |= 1u << d;
            }
            double cs_old[32];
            double score_old = or_bnet_score(b, state, cs_old);      /* mh.jl:58 */
            for (int64_t s = 0; s < steps; ++s) {                    /* mh.jl:61 */
                uint32_t o[4];
                stream4(seed, gch, (uint32_t)s, (uint32_t)(s >> 32), P_MH, o);
                int j = (int)(((uint64_t)o[0] * (uint64_t)D) >> 32); /* mh.jl:62 rand(1:n) */
                double up = (double)o[1] * 0x1p-32, ua = (double)o[2] * 0x1p-32;
                /* mh.jl:70-76: forced fresh draw at site j given the unchanged prefix */
                uint32_t prefix = state & ((1u << j) - 1u);
                uint32_t prop = state & ~(1u << j);
                if (up < site_p1(b, j, prefix)) prop |= 1u << j;
                double cs_new[32];
                double score_new = or_bnet_score(b, prop, cs_new);   /* later sites reuse values, mh.jl:30-31 */
                /* mh.jl:95-111 with |old|=|new|=D: -log D cancels */
                double fw = cs_new[j], bw = cs_old[j];
                double lr = score_new - score_old + bw - fw;
                double a = exp(lr);
                int accept = !isnan(a) && ua < (a < 1.0 ? a : 1.0);  /* mh.jl:78; NaN compares false */
                if (accept) {                                        /* mh.jl:80-84 */
                    state = prop; score_old = score_new;
                    memcpy(cs_old, cs_new, sizeof(double) * (size_t)D);
                    ++acc_loc;
                }
                loc[pack_bits(state, b->query_mask)] += 1;           /* mh.jl:87 */
            }
        }
        #pragma omp critical
        {
            for (int64_t v = 0; v < nv; ++v) counts[v] += loc[v];
            acc_total += acc_loc;
        }
        free(loc);
    }
    if (accepts) *accepts = acc_total;
    return OR_OK;
}
