// sb_small.cuh -- whole PMCMC chains in ONE kernel launch (sm_100a).
//
// Stochy programs are usually run with a few hundred particles (examples/hmm2.jl:22 `smc(100)`,
// test/inferencetests.jl:36 `pmcmc(5,5)`).  At that size the tiled path is bound by kernel launches
// (about 45 per iteration), so when the pool of a chain (N particles + the retained one) fits ONE
// tile, one CTA runs the chain's entire `for i in 1:numiterations` loop (src/pmcmc.jl:83-94) on chip:
// propagate, weight, resample (all three numerics), history, the next retained particle and the
// value counts -- and `chains` independent chains run side by side, one CTA each (the "chain set"
// of a GPU).  Chain c uses the Philox key seed + c * 0x9E3779B97F4A7C15, so chain 0 reproduces the
// tiled path and the oracle bit for bit, and chain c the oracle run with that key.
//
// Reference lines: sample src/pmcmc.jl:32-34, factor :36-46, resample :48-66, pmcmcexit :68-76,
// pmcmc loop :78-99, rand(ps) src/rand.jl:1-15.
#pragma once
#include "sb_kernels.cuh"

#define SB_CHAIN_KEY 0x9E3779B97F4A7C15ull

struct SbSmallArgs {
    // model (device tables built by sb_model_discrete_hmm)
    int K, T;
    const uint32_t* tabA;            // [K << SB_LG_BUCKETS] lookup words of the transition rows
    const uint32_t* tab0;            // [1 << SB_LG_BUCKETS] row of the first draw (pi, or A[x0])
    const uint32_t* thrA;            // [K*K] thresholds (scanned only in buckets holding several)
    const uint32_t* thr0;            // [K]
    const double* logF;              // [T*K]
    const double* expF;              // [T*K] exp(logF): literal numerics (src/pmcmc.jl:55)
    const unsigned char* has_factor; // [T]
    // run
    int N;                           // particles per chain (N + 1 <= SB_TILE)
    int S;                           // row stride of the history (>= N + 1)
    long long iters;
    long long chain0;
    int mode;                        // SB_RESAMPLE_* (0 systematic, 1 multinomial on the grid, 2 literal)
    int pick;                        // retained particle: 0 = particle 1 (src/pmcmc.jl:89), 1 = uniform pick
    unsigned long long seed;
    // per-chain scratch in global memory (L2-resident at these sizes)
    int* x_hist;                     // [chains][T][S]
    int* anc_hist;                   // [chains][T][S]  row t = ancestors drawn after step t
    int* xstar;                      // [chains][2][T]
    // outputs
    unsigned long long* marg;        // [T*K] += per-time value counts of ALL N particles (src/pmcmc.jl:90-92), or nullptr
    unsigned long long* joint;       // [K^T] += trajectory counts, or nullptr
    double* logZ;                    // [chains][2]: log-evidence of the first and of the last sweep
    unsigned* err;
    int hist_smem;                   // 1: the chain's history rows (T * S states + ancestors) live in shared memory
    int marg_smem;                   // 1: per-time counts are accumulated in shared memory and flushed once per chain
};

struct SbSmallSmem {
    unsigned long long scan[SB_WARPS + 1];
    double redd[SB_WARPS];
    SbStepInfo si;
    unsigned long long Q;            // tile sum of the current pool
    int e;                           // tile exponent of the current pool
    double lits;                     // literal numerics: sum(weights)
    double lz;
};

// dynamic shared memory (SP = pool slots rounded up to the 2048 / 8 blocked layout of the CTA, i.e. S8 = 8 * ceil(S / 8)):
// tabA | tab0 | l2[K] | ef[K] | thrA | thr0 | q[K] | x[S8] | w[S8] | anc[S8] | big[(S8+8) * 8] | wlit[(S8+8) * 8] | marg[T*K] | hist[2*T*S]
__host__ __device__ inline size_t sb_small_smem_bytes(int K, int T, int S, int hist_smem, int marg_smem) {
    const size_t S8 = ((size_t)S + 7) & ~(size_t)7;
    return ((size_t)(K + 1) << SB_LG_BUCKETS) * 4 + (size_t)2 * K * 8 + (size_t)(K * K + K + K) * 4 + 3 * S8 * 4 + 2 * (S8 + 8) * 8 + 16 +
           (marg_smem ? (size_t)T * K * 4 : 0) + (hist_smem ? (size_t)2 * T * S * 4 : 0);
}

template <bool FP32>
__global__ void __launch_bounds__(SB_THREADS)
k_pmcmc_small(const __grid_constant__ SbSmallArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ SbSmallSmem sm;
    const int K = a.K, T = a.T, N = a.N, S = a.S;
    const int G = 1 << SB_LG_BUCKETS;
    uint32_t* s_tabA = reinterpret_cast<uint32_t*>(dyn);
    uint32_t* s_tab0 = s_tabA + (size_t)K * G;
    double* s_l2 = reinterpret_cast<double*>(s_tab0 + G);
    double* s_ef = s_l2 + K;                                                    // exp(logF[t]) (literal numerics)
    uint32_t* s_thrA = reinterpret_cast<uint32_t*>(s_ef + K);
    uint32_t* s_thr0 = s_thrA + K * K;
    uint32_t* s_q = s_thr0 + K;
    const int S8 = (S + 7) & ~7;
    int* s_x = reinterpret_cast<int*>(s_q + K);
    uint32_t* s_w = reinterpret_cast<uint32_t*>(s_x + S8);
    int* s_anc = reinterpret_cast<int*>(s_w + S8);
    unsigned char* p8 = reinterpret_cast<unsigned char*>(s_anc + S8);
    p8 += (8 - (reinterpret_cast<size_t>(p8) & 7)) & 7;
    unsigned long long* s_cdf = reinterpret_cast<unsigned long long*>(p8);      // grid CDF (multinomial) ...
    double* s_acc = reinterpret_cast<double*>(p8);                              // ... or the literal running sum
    double* s_wlit = reinterpret_cast<double*>(p8) + (S8 + 8);
    unsigned* s_marg = reinterpret_cast<unsigned*>(s_wlit + (S8 + 8));
    int* s_hist = reinterpret_cast<int*>(s_marg + (a.marg_smem ? T * K : 0));

    const int tid = threadIdx.x;
    const int base = tid * SB_ITEMS;                                 // blocked: this thread's 8 pool slots
    const unsigned long long seed = a.seed + (unsigned long long)(a.chain0 + blockIdx.x) * SB_CHAIN_KEY;
    int* xh = a.hist_smem ? s_hist : a.x_hist + (size_t)blockIdx.x * T * S;
    int* ah = a.hist_smem ? s_hist + (size_t)T * S : a.anc_hist + (size_t)blockIdx.x * T * S;
    int* xs = a.xstar + (size_t)blockIdx.x * 2 * T;
    if (a.marg_smem) for (int i = tid; i < T * K; i += SB_THREADS) s_marg[i] = 0u;

    for (int i = tid; i < K * G; i += SB_THREADS) s_tabA[i] = __ldg(a.tabA + i);
    for (int i = tid; i < G; i += SB_THREADS) s_tab0[i] = __ldg(a.tab0 + i);
    for (int i = tid; i < K * K; i += SB_THREADS) s_thrA[i] = __ldg(a.thrA + i);
    for (int i = tid; i < K; i += SB_THREADS) s_thr0[i] = __ldg(a.thr0 + i);
    __syncthreads();

    int cur = 0;                                                     // xstar buffer holding the retained trajectory
    for (long long it = 0; it < a.iters; ++it) {
        const bool cond = it > 0;                                    // src/pmcmc.jl:86: no retained particle in the first sweep
        const int n_pool = N + (cond ? 1 : 0);
        const unsigned iter = (unsigned)it;
        if (tid == 0) sm.lz = 0.0;
        // t runs to T inclusive: pass T only draws the ancestors after the last step (src/pmcmc.jl:42 at the last factor)
        for (int t = 0; t <= T; ++t) {
            // this step's scores: requested now, needed after the resampling below
            double lf = 0.0, ef = 0.0;
            if (t < T && tid < K) { lf = __ldg(a.logF + (size_t)t * K + tid); ef = __ldg(a.expF + (size_t)t * K + tid); }
            // ------------------------------------------------------------ (1) resample the pool of step t-1
            if (t > 0) {
                const int tp = t - 1;
                if (!a.has_factor[tp]) {
                    for (int i = tid; i < N; i += SB_THREADS) s_anc[i] = i;
                } else {
                    // the pool on the weight grid (one tile): log(mean weight) and the resampling plans
                    const int E = sm.e;
                    const unsigned long long Q = sm.Q;
                    const bool mass = E != SB_E_EMPTY && Q != 0ull;
                    const unsigned long long total = mass ? (Q << 24) : 0ull;       // LS = 24 for one tile
                    if (tid == 0 && mass)
                        sm.lz += log((double)total) + (double)(E - SB_QBITS - 24) * SB_LN2 - log((double)n_pool);
                    if (!mass && tid == 0) atomicOr(a.err, SB_ERRBIT_ZERO);
                    // tile-local inclusive prefix of the fixed-point weights
                    uint32_t w[SB_ITEMS];
                    unsigned tsum = 0;
#pragma unroll
                    for (int j = 0; j < SB_ITEMS; ++j) { w[j] = (base + j < n_pool) ? s_w[base + j] : 0u; tsum += w[j]; }
                    unsigned long long tot;
                    unsigned long long c = sb_block_excl_scan_u64((unsigned long long)tsum, sm.scan, &tot);
                    if (a.mode == SB_RESAMPLE_SYSTEMATIC) {
                        if (tid == 0) {
                            SbStepInfo si;
                            sb_plan_rescale(si, total, n_pool, N, E, 24, 1);
                            const unsigned long long Tt = (si.valid && Q != 0ull) ? sb_sys_delta(Q, si.R, si.r0) : 0ull;
                            const SbU4 o = sb_stream4(seed, 0ull, (unsigned)tp, iter, SB_P_RESAMPLE);
                            sb_plan_offset(si, Tt, sb_r53(o.x, o.y));
                            if (!si.valid) atomicOr(a.err, SB_ERRBIT_ZERO);
                            sm.si = si;
                        }
                        __syncthreads();
                        const SbStepInfo si = sm.si;
                        if (si.valid) {
                            // parent i owns the children whose thresholds lie in (C_{i-1}, C_i]: strict `<` of src/rand.jl:9
                            int lo = sb_sys_F(sb_sys_delta(c, si.R, si.r0), si.U, si.s, N);
#pragma unroll
                            for (int j = 0; j < SB_ITEMS; ++j) {
                                c += w[j];
                                const int hi = sb_sys_F(sb_sys_delta(c, si.R, si.r0), si.U, si.s, N);
                                for (int k = lo; k < hi; ++k) s_anc[k] = base + j;
                                lo = hi;
                            }
                        } else {
                            for (int i = tid; i < N; i += SB_THREADS) s_anc[i] = i;
                        }
                    } else if (a.mode == SB_RESAMPLE_MULTINOMIAL) {
#pragma unroll
                        for (int j = 0; j < SB_ITEMS; ++j) { c += w[j]; if (base + j < n_pool) s_cdf[base + j] = c << 24; }
                        __syncthreads();
                        for (int i = tid; i < N; i += SB_THREADS) {
                            const SbU4 o = sb_stream4(seed, (unsigned long long)(i >> 1), (unsigned)tp, iter, SB_P_RESAMPLE);
                            const unsigned long long R = (i & 1) ? sb_r53(o.z, o.w) : sb_r53(o.x, o.y);
                            const unsigned long long thr = (__umul64hi(R, total) << 11) | ((R * total) >> 53);
                            int lo = 0, hi = n_pool - 1;
                            while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (thr < s_cdf[mid]) hi = mid; else lo = mid + 1; }
                            s_anc[i] = lo;
                        }
                    } else {
                        // literal reference numerics: sequentially rounded sum and running sum (src/pmcmc.jl:50, src/rand.jl:6-9)
                        if (tid == 0) {
                            double s = 0.0;
                            for (int i = 0; i < n_pool; ++i) s = __dadd_rn(s, s_wlit[i]);      // sum(weights), one rounding per add
                            sm.lits = s;
                        }
                        __syncthreads();
                        {
                            const double s = sm.lits;
                            for (int i = tid; i < n_pool; i += SB_THREADS) s_acc[i] = __ddiv_rn(s_wlit[i], s);   // ps = weights / sum
                        }
                        __syncthreads();
                        if (tid == 0) {
                            double acc = 0.0;
                            bool bad = false;
                            for (int i = 0; i < n_pool; ++i) {
                                acc = __dadd_rn(acc, s_acc[i]);                                  // src/rand.jl:7
                                if (isnan(acc) && (i < n_pool - 1 || n_pool == 1)) bad = true;  // src/rand.jl:8,11
                                s_acc[i] = acc;
                            }
                            if (bad) atomicOr(a.err, SB_ERRBIT_NAN);
                        }
                        __syncthreads();
                        for (int i = tid; i < N; i += SB_THREADS) {
                            const SbU4 o = sb_stream4(seed, (unsigned long long)(i >> 1), (unsigned)tp, iter, SB_P_RESAMPLE);
                            const double r = (i & 1) ? sb_u53(o.z, o.w) : sb_u53(o.x, o.y);
                            int lo = 0, hi = n_pool - 1;                  // first i in [0, n-1) with r < acc[i], else n-1
                            while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (r < s_acc[mid]) hi = mid; else lo = mid + 1; }
                            s_anc[i] = lo;
                        }
                    }
                }
                __syncthreads();
                for (int i = tid; i <= N; i += SB_THREADS) ah[(size_t)tp * S + i] = i < N ? s_anc[i] : N;   // the retained slot descends from itself
                if (t == T) break;
            }
            // ------------------------------------------------------------ (2)+(3) gather parents, propagate (src/pmcmc.jl:32-34)
            int x[SB_ITEMS];
            uint32_t r8[SB_ITEMS];
            if (base < N) {                                               // one Philox block per four particles
#pragma unroll
                for (int p = 0; p < SB_ITEMS / 4; ++p) {
                    const SbU4 o = sb_stream4(seed, (unsigned long long)((base >> 2) + p), (unsigned)t, iter, SB_P_PROP);
                    r8[4 * p] = o.x; r8[4 * p + 1] = o.y; r8[4 * p + 2] = o.z; r8[4 * p + 3] = o.w;
                }
            }
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                const int i = base + j;
                int xv = 0;
                if (i < N) {
                    const uint32_t R = r8[j];
                    const uint32_t* tab = t == 0 ? s_tab0 : s_tabA + ((size_t)s_x[s_anc[i]] << SB_LG_BUCKETS);
                    const uint32_t g = tab[R >> (32 - SB_LG_BUCKETS)];
                    xv = (int)(g & 63u) + (((R << SB_LG_BUCKETS) > g) ? 1 : 0);
                    if (g & (1u << (SB_LG_BUCKETS - 1))) {               // several thresholds inside the bucket: scan (src/rand.jl:6-12)
                        const uint32_t* row = t == 0 ? s_thr0 : s_thrA + s_x[s_anc[i]] * K;
                        xv = (int)(g & 63u);
                        while (xv < K - 1 && R >= row[xv]) ++xv;
                    }
                } else if (i == N && cond) {
                    xv = xs[cur * T + t];                                // retained particle (src/pmcmc.jl:59-66)
                }
                x[j] = xv;
            }
            if (tid < K) {
                if (FP32) {
                    const float l2 = __fmul_rn((float)lf, SB_LOG2E_F);
                    if (isnan(l2) || l2 == INFINITY) atomicOr(a.err, SB_ERRBIT_NAN);
                    reinterpret_cast<float*>(s_l2)[tid] = l2;
                } else {
                    const double l2 = __dmul_rn(lf, SB_LOG2E_D);
                    if (isnan(l2) || l2 == INFINITY) atomicOr(a.err, SB_ERRBIT_NAN);
                    s_l2[tid] = l2;
                }
                s_ef[tid] = ef;
            }
            __syncthreads();                                             // every read of the old states is done; scores staged
            // ------------------------------------------------------------ (4)+(5) weight and quantise (one tile)
            int e;
            if (FP32) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) if (base + j < n_pool) mx = fmaxf(mx, reinterpret_cast<float*>(s_l2)[x[j]]);
                mx = sb_block_max_f32(mx, reinterpret_cast<float*>(sm.redd));
                e = mx > -INFINITY ? (int)ceilf(mx) : SB_E_EMPTY;
                if (tid < K) s_q[tid] = e == SB_E_EMPTY ? 0u : sb_q31_f32(__fsub_rn(reinterpret_cast<float*>(s_l2)[tid], (float)e));
            } else {
                double mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) if (base + j < n_pool) mx = fmax(mx, s_l2[x[j]]);
                mx = sb_block_max_f64(mx, sm.redd);
                e = mx > -INFINITY ? (int)ceil(mx) : SB_E_EMPTY;
                if (tid < K) s_q[tid] = e == SB_E_EMPTY ? 0u : sb_q31_f64(__dsub_rn(s_l2[tid], (double)e));
            }
            __syncthreads();
            unsigned long long tsum = 0;
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                const int i = base + j;
                const uint32_t q = i < n_pool ? s_q[x[j]] : 0u;
                tsum += q;
                if (i < n_pool) {
                    s_x[i] = x[j];
                    s_w[i] = q;
                    xh[(size_t)t * S + i] = x[j];                        // Step push (src/pmcmc.jl:37)
                    if (a.mode == SB_RESAMPLE_LITERAL) s_wlit[i] = s_ef[x[j]];
                }
            }
            const unsigned long long Q = sb_block_sum_u64(tsum, sm.scan);
            if (tid == 0) { sm.Q = Q; sm.e = e; }
            __syncthreads();
        }
        __syncthreads();                                                 // the last ancestor row is complete
        // ---------------------------------------------------------------- the next retained particle (src/pmcmc.jl:89)
        if (tid == 0) {
            int c = 0;
            if (a.pick) {
                const SbU4 o = sb_stream4(seed, 0ull, (unsigned)T, iter, SB_P_PICK);
                c = (int)__double2ll_rz(__dmul_rn(sb_u53(o.x, o.y), (double)N));
                if (c >= N) c = N - 1;
            }
            int j = ah[(size_t)(T - 1) * S + c];
            for (int t = T - 1; t >= 0; --t) {
                xs[(cur ^ 1) * T + t] = xh[(size_t)t * S + j];
                if (t) j = ah[(size_t)(t - 1) * S + j];
            }
            a.logZ[(size_t)blockIdx.x * 2 + 1] = sm.lz;
            if (it == 0) a.logZ[(size_t)blockIdx.x * 2] = sm.lz;
        }
        // ---------------------------------------------------------------- counts[p.value] += 1 for ALL particles (src/pmcmc.jl:90-92)
        for (int c = tid; c < N; c += SB_THREADS) {
            int j = ah[(size_t)(T - 1) * S + c];
            unsigned long long code = 0;
            for (int t = T - 1; t >= 0; --t) {
                const int xv = xh[(size_t)t * S + j];
                if (a.marg_smem) atomicAdd(s_marg + t * K + xv, 1u);
                else if (a.marg) atomicAdd(a.marg + (size_t)t * K + xv, 1ull);
                code = code * (unsigned long long)K + (unsigned long long)xv;
                if (t) j = ah[(size_t)(t - 1) * S + j];
            }
            if (a.joint) atomicAdd(a.joint + code, 1ull);
        }
        cur ^= 1;
        __syncthreads();                                                 // xstar of the next sweep is written; history may be overwritten
    }
    // the retained trajectory of the chain is now in buffer (iters & 1) of xstar
    if (a.marg_smem && a.marg)
        for (int i = tid; i < T * K; i += SB_THREADS) if (s_marg[i]) atomicAdd(a.marg + i, (unsigned long long)s_marg[i]);
}
