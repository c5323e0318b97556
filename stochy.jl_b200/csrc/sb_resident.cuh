// sb_resident.cuh -- a whole SMC sweep of the discrete model in ONE cooperative kernel launch (sm_100a).
//
// Pools whose tiles all fit on the GPU at once (one CTA per tile of 2048 particles, up to ~590 tiles = 1.2 M
// particles: BASELINE cfg3's 2^20) spend more time between the kernels of a sweep than inside them: at 2^20 a step
// moves 21 MB (3 us of HBM time) while the step kernel's ramp, the serial tile scan and the launch hand-overs cost
// 15 us.  Here the CTAs stay resident for the whole `for t` loop: each owns ONE tile of the pool, the barrier of
// `factor` (src/pmcmc.jl:36-46: every particle reaches the statement before anyone resamples) is one grid-wide
// barrier per step, and the normalisation behind it (the tile scan: src/pmcmc.jl:50) is computed REDUNDANTLY by every
// CTA from the n/2048 tile records (12 bytes per tile, L2 hits) into its own shared memory -- integer adds only, so
// every CTA gets the same bits as k_tilescan -- instead of by one serial kernel plus a second barrier.  The Philox
// block of the next step and its score tables are computed before the barrier (they depend on nothing).
//
// Same integers as the tiled path (k_step_hmm + k_tilescan): states, ancestors, tile records and the per-step
// SbStepInfo records are bit-identical (tests/test_gpu_sweep.py::test_resident_sweep_equals_tiled_kernels).
// Everything written inside the kernel and read by another CTA later is read with ordinary (coherent) loads or TMA
// behind the grid barrier -- never ld.global.nc.
//
// Replaces, per sweep: T x (sample -> rand(e) src/pmcmc.jl:32-34, factor :36-46, resample :48-66).
#pragma once
#include "sb_kernels.cuh"

#define SB_RES_MAXPER 4                     // tile records per thread of the redundant scan: pools of at most 1024 tiles

struct SbResArgs {
    // model tables (device; built by sb_model_discrete_hmm)
    const uint32_t* tabA;                   // [K << SB_LG_BUCKETS] lookup words of the transition rows
    const uint32_t* tab0;                   // [1 << SB_LG_BUCKETS] row of the first draw (pi, or A[x0])
    const uint32_t* thrA;                   // [K*K] thresholds (scanned only in buckets holding several)
    const uint32_t* thr0;                   // [K]
    const double* logF;                     // [T*K]
    int K, T;
    // the two pool slots (slot = t & 1)
    uint32_t* w32[2];
    int* e[2];
    unsigned long long* Q[2];
    unsigned long long* wp[2];
    int* xbuf;                              // rows of S entries: row(t) = history ? t : t & 1
    int* ancbuf;                            // history: row t-1 = ancestors drawn after step t-1; else row 0 = the final resample
    long long S;
    int history, final_anc;
    const int* xstar;                       // retained trajectory [T] (conditional SMC) or nullptr
    long long N;                            // particles
    int n_out;                              // N (+1 with the retained particle)
    int ntiles;                             // tiles_of(n_out) == gridDim.x
    const unsigned long long* r53;          // [T] host-drawn systematic offsets
    SbStepInfo* info;                       // [T] per-step records (log-evidence on the host)
    unsigned iter;
    SbRoundKeys rk;
    unsigned* err;
};

struct SbResSmem {
    SbPullSmem ps;
    SbStepInfo si;
    unsigned long long scan[SB_WARPS + 1];
    unsigned long long tot[SB_WARPS];
    unsigned red[SB_WARPS];
    int redi[SB_WARPS];
};

// grid-wide barrier: every global write of every CTA before it is visible to every thread after it
__device__ __forceinline__ void sb_grid_sync(cg::grid_group& g) { g.sync(); }

__device__ __forceinline__ void sb_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// The tile scan of k_tilescan, by ONE CTA for itself: the same integer formulas in the same roles, a block scan
// instead of the cluster scan.  Results: sm.si, s_Pt[0..nt], s_cs[0..nt] (child_start).  Thread r owns the `per`
// consecutive tiles from r * per.
__device__ __forceinline__ void sb_res_scan(const int* e_g, const unsigned long long* Q_g, int nt, long long n, long long m,
                                            unsigned long long R53, SbResSmem& sm, unsigned long long* s_Pt, int* s_cs,
                                            SbStepInfo* info_out, unsigned* err) {
    const int tid = threadIdx.x;
    const int per = (nt + SB_THREADS - 1) / SB_THREADS;          // <= SB_RES_MAXPER
    const int b0 = min(nt, tid * per), b1 = min(nt, b0 + per);
    unsigned long long qr[SB_RES_MAXPER];
    int er[SB_RES_MAXPER];
#pragma unroll
    for (int i = 0; i < SB_RES_MAXPER; ++i) {
        const int b = b0 + i;
        qr[i] = 0ull; er[i] = SB_E_EMPTY;
        if (b < b1) { qr[i] = __ldcg(Q_g + b); er[i] = __ldcg(e_g + b); }
    }
    int E = SB_E_EMPTY;
#pragma unroll
    for (int i = 0; i < SB_RES_MAXPER; ++i) if (qr[i] != 0ull) E = max(E, er[i]);
    E = sb_block_max_i32(E, sm.redi);
    const int LS = 24 - (nt <= 1 ? 0 : 32 - __clz(nt - 1));
    unsigned long long loc = 0;
#pragma unroll
    for (int i = 0; i < SB_RES_MAXPER; ++i)
        if (qr[i] != 0ull) { const int sh = E - er[i]; loc += sh >= 63 ? 0ull : ((qr[i] << LS) >> sh); }
    unsigned long long total;
    (void)sb_block_excl_scan_u64(loc, sm.scan, &total);
    SbStepInfo si;
    sb_plan_rescale(si, total, n, m, E, LS, nt);
    unsigned long long loc2 = 0;
    if (si.valid) {
#pragma unroll
        for (int i = 0; i < SB_RES_MAXPER; ++i) if (qr[i] != 0ull) loc2 += sb_sys_delta(qr[i], si.R, si.r0 + (E - er[i]));
    }
    unsigned long long Tt;
    unsigned long long run2 = sb_block_excl_scan_u64(loc2, sm.scan, &Tt);
    sb_plan_offset(si, Tt, R53);
    const int mi = (int)m;
#pragma unroll
    for (int i = 0; i < SB_RES_MAXPER; ++i) {
        const int b = b0 + i;
        if (b < b1) {
            s_Pt[b] = run2;
            s_cs[b] = si.valid ? sb_sys_F(run2, si.U, si.s, mi) : 0;
            if (si.valid && qr[i] != 0ull) run2 += sb_sys_delta(qr[i], si.R, si.r0 + (E - er[i]));
        }
    }
    if (tid == 0) {
        s_Pt[nt] = Tt; s_cs[nt] = mi;
        sm.si = si;
        if (info_out) { *info_out = si; if (!si.valid) atomicOr(err, SB_ERRBIT_ZERO); }
    }
    __syncthreads();
}

// candidate parent tiles of the children tile starting at c0: what k_tilescan's first_tile table holds, found by a
// binary search in the child_start array (first tile b with cs[b] <= c0 < cs[b+1])
__device__ __forceinline__ int sb_res_first_tile(const int* s_cs, int nt, int c) {
    int lo = 0, hi = nt;                                          // first j in [0, nt] with cs[j] > c  (cs[nt] = m > c)
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cs[mid] > c) hi = mid; else lo = mid + 1; }
    return lo - 1;
}

// ancestors of this CTA's children tile from the pool in slot `prev` (scan results in shared memory).  The marker
// array is all -1 on entry and on exit.  `use` = the number of pulls this CTA has done before (mbarrier parity).
template <bool S32>
__device__ __forceinline__ void sb_res_pull(const SbResArgs& a, int prev, int c0, int use, SbResSmem& sm,
                                            const unsigned long long* s_Pt, const int* s_cs, int anc[SB_ITEMS]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SbStepInfo& si = sm.si;
    const int nt = si.nt, m = (int)si.m;
    if (tid == 0) {
        int2 r = make_int2(0, -1);
        if (c0 < m) {
            r.x = sb_res_first_tile(s_cs, nt, c0);
            r.y = (c0 + SB_TILE < m) ? sb_res_first_tile(s_cs, nt, c0 + SB_TILE) : nt - 1;
        }
        sm.ps.ft[0] = r;
    }
    __syncthreads();
    const int2 rng = sm.ps.ft[0];
    const int bA = rng.x, bB = min(rng.x + 1, nt - 1);
    if (tid == 0) {
        sb_fence_proxy_async();                                   // the weights were written through the generic proxy (other CTAs)
        sb_mbar_expect_tx(&sm.ps.bar, 2u * SB_TILE * 4u);
        sb_bulk_g2s(sm.ps.w[0], a.w32[prev] + (size_t)bA * SB_TILE, SB_TILE * 4u, &sm.ps.bar);
        sb_bulk_g2s(sm.ps.w[1], a.w32[prev] + (size_t)bB * SB_TILE, SB_TILE * 4u, &sm.ps.bar);
    }
    if (warp == SB_META_WARP) {
        SbPullMeta& mt = sm.ps.meta[0];
        const int which = lane < 16 ? (lane >> 3) : (lane & 1);
        const int bsel = which ? bB : bA;
        if (lane < 16) mt.wp[which][lane & 7] = __ldcg(a.wp[prev] + (size_t)bsel * SB_WARPS + (lane & 7));
        else if (lane < 18) mt.Pt[which] = s_Pt[bsel];
        else if (lane < 20) mt.cs[which] = s_cs[bsel];
        else if (lane < 22) mt.ce[which] = s_cs[bsel + 1];
        else if (lane < 24) mt.e[which] = __ldcg(a.e[prev] + bsel);
        __syncwarp();
        sb_pull_seg<S32>(si, c0, rng, 0, sm.ps);
    }
    __syncthreads();
    sb_pull_ancestors<false, S32, true>(si, SbPool{a.w32[prev], nullptr, a.wp[prev], {nullptr}, 0, 0, 0, 31, 0, 0}, a.e[prev], s_Pt, s_cs,
                                       c0, rng, 0, (uint32_t)(use & 1), false, 0, sm.ps, anc);
}

template <bool FP32, bool MULTI>
__global__ void __launch_bounds__(SB_THREADS, 4)
k_sweep_hmm(const __grid_constant__ SbResArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ SbResSmem sm;
    const int K = a.K, T = a.T, nt = a.ntiles;
    const int G = 1 << SB_LG_BUCKETS;
    uint32_t* s_tabA = reinterpret_cast<uint32_t*>(dyn);                   // K * 256
    uint32_t* s_tab0 = s_tabA + (size_t)K * G;                             // 256
    double* s_l2 = reinterpret_cast<double*>(s_tab0 + G);                  // K (float view when FP32)
    unsigned long long* s_Pt = reinterpret_cast<unsigned long long*>(s_l2 + K);   // nt + 1
    int* s_cs = reinterpret_cast<int*>(s_Pt + (nt + 1));                   // nt + 1 (+ pad)
    uint32_t* s_key = reinterpret_cast<uint32_t*>(s_cs + ((nt + 2) & ~1)); // K
    uint32_t* s_qtab = s_key + K;                                          // K * K
    uint32_t* s_thrA = s_qtab + K * K;                                     // K * K (MULTI)
    uint32_t* s_thr0 = s_thrA + (MULTI ? K * K : 0);                       // K (MULTI)
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int N = (int)a.N;
    const int c0 = tile * SB_TILE, base = c0 + tid * SB_ITEMS;
    for (int i = tid; i < ((K * G) >> 2); i += SB_THREADS) reinterpret_cast<uint4*>(s_tabA)[i] = __ldg(reinterpret_cast<const uint4*>(a.tabA) + i);
    for (int i = tid; i < (G >> 2); i += SB_THREADS) reinterpret_cast<uint4*>(s_tab0)[i] = __ldg(reinterpret_cast<const uint4*>(a.tab0) + i);
    if (MULTI) {
        for (int i = tid; i < K * K; i += SB_THREADS) s_thrA[i] = __ldg(a.thrA + i);
        for (int i = tid; i < K; i += SB_THREADS) s_thr0[i] = __ldg(a.thr0 + i);
    }
    {
        const int4 neg = make_int4(-1, -1, -1, -1);
        reinterpret_cast<int4*>(sm.ps.marker)[tid * 2] = neg;
        reinterpret_cast<int4*>(sm.ps.marker)[tid * 2 + 1] = neg;
    }
    if (tid == 0) sb_mbar_init(&sm.ps.bar, 1);
    __syncthreads();
    int pulls = 0;
    SbU4 rnd[SB_ITEMS / 4];
#pragma unroll
    for (int p = 0; p < SB_ITEMS / 4; ++p) rnd[p] = sb_stream4_rk(a.rk, (unsigned long long)((base >> 2) + p), 0u, a.iter, SB_P_PROP);
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1, prev = cur ^ 1;
        int* xrow = a.xbuf + (size_t)(a.history ? t : cur) * a.S;
        const int* xprev = a.xbuf + (size_t)(a.history ? t - 1 : prev) * a.S;
        // ---- this step's scores: exponent keys and the K x K table of grid weights (as k_step_hmm's prologue)
        if (tid < K) {
            const double lf = __ldg(a.logF + (size_t)t * K + tid);
            int ce; bool live;
            if (FP32) {
                const float l2 = __fmul_rn((float)lf, SB_LOG2E_F);
                if (isnan(l2) || l2 == INFINITY) atomicOr(a.err, SB_ERRBIT_NAN);
                reinterpret_cast<float*>(s_l2)[tid] = l2;
                live = l2 > -INFINITY; ce = (int)ceilf(fminf(fmaxf(l2, -16777215.0f), 16777215.0f));
            } else {
                const double l2 = __dmul_rn(lf, SB_LOG2E_D);
                if (isnan(l2) || l2 == INFINITY) atomicOr(a.err, SB_ERRBIT_NAN);
                s_l2[tid] = l2;
                live = l2 > -INFINITY; ce = (int)ceil(fmin(fmax(l2, -16777215.0), 16777215.0));
            }
            s_key[tid] = live ? (((unsigned)(ce + SB_KEY_BIAS) << 6) | (unsigned)tid) : 0u;
        }
        __syncthreads();
        for (int i = tid; i < K * K; i += SB_THREADS) {
            const int ks = i / K, kx = i - ks * K;
            const unsigned key = s_key[ks];
            const int e = (int)(key >> 6) - SB_KEY_BIAS;
            uint32_t q = 0u;
            if (key) q = FP32 ? sb_q31_f32(__fsub_rn(reinterpret_cast<float*>(s_l2)[kx], (float)e)) : sb_q31_f64(__dsub_rn(s_l2[kx], (double)e));
            s_qtab[i] = q;
        }
        // ---- (1) ancestors: the pull from the pool of step t-1
        int anc[SB_ITEMS];
        if (t > 0) {
            if (sm.si.s >= 32) sb_res_pull<true>(a, prev, c0, pulls, sm, s_Pt, s_cs, anc);
            else               sb_res_pull<false>(a, prev, c0, pulls, sm, s_Pt, s_cs, anc);
            ++pulls;
            if (a.history) {
                int* arow = a.ancbuf + (size_t)(t - 1) * a.S + base;
                if (base + SB_ITEMS <= N) {
                    reinterpret_cast<int4*>(arow)[0] = make_int4(anc[0], anc[1], anc[2], anc[3]);
                    reinterpret_cast<int4*>(arow)[1] = make_int4(anc[4], anc[5], anc[6], anc[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < SB_ITEMS; ++j) {
                        const int i = base + j;
                        if (i < N) arow[j] = anc[j];
                        else if (i == N) arow[j] = N;                     // the retained slot descends from itself
                    }
                }
            }
        } else {
            __syncthreads();                                             // s_qtab complete (the pull has its own barriers)
        }
        // ---- (2) gather the parents' states (coherent loads: written by other CTAs in this launch)
        int xp[SB_ITEMS];
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) xp[j] = (t > 0 && base + j < N) ? __ldcg(xprev + anc[j]) : 0;
        // ---- (3) propagate (src/pmcmc.jl:32-34): the draws of k_step_hmm on the same Philox block
        int x[SB_ITEMS];
#pragma unroll
        for (int p = 0; p < SB_ITEMS / 4; ++p) {
            const uint32_t r4[4] = {rnd[p].x, rnd[p].y, rnd[p].z, rnd[p].w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int j = 4 * p + h;
                const uint32_t R = r4[h];
                const uint32_t* tb = t ? s_tabA + ((size_t)xp[j] << SB_LG_BUCKETS) : s_tab0;
                const uint32_t g = tb[R >> (32 - SB_LG_BUCKETS)];
                int lo = (int)(g & 63u) + (((R << SB_LG_BUCKETS) > g) ? 1 : 0);
                if (MULTI && (g & 128u)) {
                    const uint32_t* row = t ? s_thrA + xp[j] * K : s_thr0;
                    lo = (int)(g & 63u);
                    while (lo < K - 1 && R >= row[lo]) ++lo;
                }
                x[j] = lo;
            }
        }
        if (base + SB_ITEMS > N) {                                       // retained particle (src/pmcmc.jl:59-66) and padding
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                const int i = base + j;
                if (i >= N) x[j] = (i == N && a.xstar) ? __ldg(a.xstar + t) : 0;
            }
        }
        // ---- (4)+(5) weight and quantise
        unsigned key = 0u;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) if (base + j < a.n_out) key = max(key, s_key[x[j]]);
        key = __reduce_max_sync(0xffffffffu, key);
        if ((tid & 31) == 0) sm.red[tid >> 5] = key;
        __syncthreads();
        {
            const uint4 u = reinterpret_cast<const uint4*>(sm.red)[0], v = reinterpret_cast<const uint4*>(sm.red)[1];
            key = max(max(max(u.x, u.y), max(u.z, u.w)), max(max(v.x, v.y), max(v.z, v.w)));
        }
        const int e = key ? (int)(key >> 6) - SB_KEY_BIAS : SB_E_EMPTY;
        const uint32_t* qrow = s_qtab + (key & 63u) * K;
        uint32_t q[SB_ITEMS];
        unsigned tsum = 0;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) {
            q[j] = (key != 0u && base + j < a.n_out) ? qrow[x[j]] : 0u;
            tsum += q[j];
        }
        sb_store8_u32(reinterpret_cast<uint32_t*>(xrow) + base, reinterpret_cast<const uint32_t*>(x));
        sb_store8_u32(a.w32[cur] + base, q);
        sb_block_sum_wp(tsum, sm.tot, a.wp[cur] + (size_t)tile * SB_WARPS, a.Q[cur] + tile);
        if (tid == 0) a.e[cur][tile] = e;
        // ---- independent of the barrier: the next step's Philox block
        if (t + 1 < T) {
#pragma unroll
            for (int p = 0; p < SB_ITEMS / 4; ++p) rnd[p] = sb_stream4_rk(a.rk, (unsigned long long)((base >> 2) + p), (unsigned)(t + 1), a.iter, SB_P_PROP);
        }
        // ---- the barrier of `factor` (src/pmcmc.jl:36-46), then the normalisation (src/pmcmc.jl:50) by every CTA for itself
        sb_grid_sync(grid);
        sb_res_scan(a.e[cur], a.Q[cur], nt, (long long)a.n_out, a.N, __ldg(a.r53 + t), sm, s_Pt, s_cs,
                    tile == 0 ? a.info + t : nullptr, a.err);
    }
    if (a.final_anc && c0 < N) {                                         // the resample after the last step (src/pmcmc.jl:42)
        int anc[SB_ITEMS];
        const int prev = (T - 1) & 1;
        if (sm.si.s >= 32) sb_res_pull<true>(a, prev, c0, pulls, sm, s_Pt, s_cs, anc);
        else               sb_res_pull<false>(a, prev, c0, pulls, sm, s_Pt, s_cs, anc);
        int* arow = a.ancbuf + (size_t)(a.history ? T - 1 : 0) * a.S + base;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) if (base + j < N) arow[j] = anc[j];
    }
}

__host__ inline size_t sb_res_smem_bytes(int K, int nt, bool multi) {
    const size_t G = (size_t)1 << SB_LG_BUCKETS;
    return ((size_t)K * G + G) * 4 + (size_t)K * 8 + (size_t)(nt + 1) * 8 + (size_t)((nt + 2) & ~1) * 4 +
           (size_t)(K + K * K) * 4 + (multi ? (size_t)(K * K + K) * 4 : 0);
}
