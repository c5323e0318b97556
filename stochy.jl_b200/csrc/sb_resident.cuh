// sb_resident.cuh -- a whole SMC sweep of the discrete model in ONE cooperative kernel launch (sm_100a).
//
// Pools whose tiles all fit on the GPU at once (one CTA per tile of 2048 particles, up to ~590 tiles = 1.2 M
// particles: BASELINE cfg3's 2^20) spend more time between the kernels of a sweep than inside them: at 2^20 a step
// moves 21 MB (3 us of HBM time) while the step kernel's ramp, the serial tile scan and the launch hand-overs cost
// 15 us.  Here the CTAs stay resident for the whole `for t` loop: each owns ONE tile of the pool, the barrier of
// `factor` (src/pmcmc.jl:36-46: every particle reaches the statement before anyone resamples) is one grid-wide
// barrier per step, and the normalisation behind it (the tile scan: src/pmcmc.jl:50) is computed REDUNDANTLY by every CTA from those
// records into its own shared memory -- integer adds only, so every CTA gets the same bits as k_tilescan -- instead
// of by one serial kernel plus a second barrier.  The Philox block of the next step and its score tables are
// computed while the CTA waits for the others (they depend on nothing).
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
    ulonglong2* rec[2];                     // [ntiles] published tile records: (Q | epoch16 << 48, e | epoch32 << 32)
    int* xbuf;                              // rows of S entries: row(t) = history ? t : t & 1
    int* ancbuf;                            // history: row t-1 = ancestors drawn after step t-1; else row 0 = the final resample
    long long S;
    int history, final_anc;
    const int* xstar;                       // retained trajectory [T] (conditional SMC) or nullptr
    long long N;                            // particles
    int n_out;                              // N (+1 with the retained particle)
    int ntiles;                             // tiles_of(n_out) == gridDim.x
    int xb;                                 // > 0: head markers carry the parent's state in xb bits (no gather)
    unsigned epoch0;                        // the record of step t carries epoch0 + t + 1 (fresh for every launch)
    unsigned* bar;                          // grid barrier counter (sb_grid_barrier) and its value before this launch
    unsigned bar0;
    const unsigned long long* r53;          // [T] host-drawn systematic offsets
    SbStepInfo* info;                       // [T] per-step records (log-evidence on the host)
    unsigned iter;
    SbRoundKeys rk;
    unsigned* err;
    unsigned long long* dbg;                // diagnostics (SB_RES_TRACE): globaltimer stamps of CTA 0, 16 per step
};

#define SB_RES_STAMP(i) do { if (TRACE && a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && t < 64) { unsigned long long ts_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_)); a.dbg[t * 16 + (i)] = ts_; } } while (0)

struct SbResSmem {
    SbPullSmem ps;
    SbStepInfo si;
    unsigned long long scanA[SB_WARPS], scanB[SB_WARPS];
    unsigned long long tot[SB_WARPS];
    __align__(16) unsigned red[SB_WARPS];   // read back as two 16-byte vectors
    __align__(16) int redi[SB_WARPS];
    int bA, bLast;                          // candidate parent tiles of this CTA's children tile (found inside the scan)
    unsigned kmax, qthr;                    // as in SbHmmSmem
};

__device__ __forceinline__ void sb_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ ulonglong2 sb_ld_rec_volatile(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sb_st_rec_volatile(ulonglong2* p, unsigned long long x, unsigned long long y) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}

// The normalisation behind the barrier of `factor` (src/pmcmc.jl:50).  Every CTA publishes ONE 16-byte record per step
// -- its tile sum and exponent, both words stamped with the step's epoch (a stale record is an error, not a hang) --
// and, behind the grid barrier, reads ALL records (at most four per thread, one round trip to L2).
// Then the tile scan of k_tilescan by ONE CTA for itself: the same integer formulas in the same roles, block scans
// instead of the cluster scan, so every CTA gets the same bits.  Results: sm.si, s_Pt[0..nt], s_cs[0..nt]
// (child_start), and the candidate parent tiles of the children tile starting at c0 (what k_tilescan's first_tile
// table holds): sm.bA = the tile b with cs[b] <= c0 < cs[b+1], sm.bLast likewise for c0 + SB_TILE (or nt - 1).
// Thread r owns the `per` consecutive tiles from r * per.  Four CTA barriers.
// (One warp doing this scan for the CTA, the other seven parked at the barrier, was measured: the CTAs publish closer
// together -- max 4.4 us behind the first instead of 6.0: an SM holding four CTAs is issue-bound and the redundant scan
// is a third of its instructions -- but a lone warp issues one dependent instruction every 5-10 cycles: 10 us per scan
// against 2.9 us.  profiles/r02_notes.md.)
template <int PER>
__device__ __forceinline__ void sb_res_scan(const ulonglong2* rec, unsigned epoch, int nt, long long n, long long m,
                                            unsigned long long R53, int c0, SbResSmem& sm, unsigned long long* s_Pt, int* s_cs,
                                            SbStepInfo* info_out, unsigned* err) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = PER;                                          // tiles per thread (>= ceil(nt / SB_THREADS))
    const int b0 = min(nt, tid * per), b1 = min(nt, b0 + per);
    unsigned long long qr[PER];
    int er[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int b = b0 + i;
        qr[i] = 0ull; er[i] = SB_E_EMPTY;
        if (b < b1) {
            // behind the grid barrier every record of this step is there (polling the records instead of a barrier was
            // measured: 513 CTAs x 256 threads spinning on L2 slow the CTAs that still work -- 16.0 vs 14.5 us per step)
            const ulonglong2 v = sb_ld_rec_volatile(rec + b);
            if ((unsigned)(v.x >> 48) != (epoch & 0xffffu) || (unsigned)(v.y >> 32) != epoch) atomicOr(err, SB_ERRBIT_PEER);
            qr[i] = v.x & ((1ull << 48) - 1ull); er[i] = (int)(unsigned)v.y;
        }
    }
    if (tid == 0) { sm.bA = 0; sm.bLast = nt - 1; }
    int E = SB_E_EMPTY;
#pragma unroll
    for (int i = 0; i < PER; ++i) if (qr[i] != 0ull) E = max(E, er[i]);
    E = __reduce_max_sync(0xffffffffu, E);
    if (lane == 0) sm.redi[warp] = E;
    __syncthreads();                                              // (1)
    {
        const int4 u = reinterpret_cast<const int4*>(sm.redi)[0], v = reinterpret_cast<const int4*>(sm.redi)[1];
        E = max(max(max(u.x, u.y), max(u.z, u.w)), max(max(v.x, v.y), max(v.z, v.w)));
    }
    const int LS = 24 - (nt <= 1 ? 0 : 32 - __clz(nt - 1));
    unsigned long long loc = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i)
        if (qr[i] != 0ull) { const int sh = E - er[i]; loc += sh >= 63 ? 0ull : ((qr[i] << LS) >> sh); }
    unsigned long long inc = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = sb_shfl_up_u64(inc, d); if (lane >= d) inc += o; }
    if (lane == 31) sm.scanA[warp] = inc;
    __syncthreads();                                              // (2)
    unsigned long long total = 0;
#pragma unroll
    for (int w = 0; w < SB_WARPS; ++w) total += sm.scanA[w];
    SbStepInfo si;
    sb_plan_rescale(si, total, n, m, E, LS, nt);
    unsigned long long loc2 = 0;
    if (si.valid) {
#pragma unroll
        for (int i = 0; i < PER; ++i) if (qr[i] != 0ull) loc2 += sb_sys_delta(qr[i], si.R, si.r0 + (E - er[i]));
    }
    unsigned long long inc2 = loc2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = sb_shfl_up_u64(inc2, d); if (lane >= d) inc2 += o; }
    if (lane == 31) sm.scanB[warp] = inc2;
    __syncthreads();                                              // (3)
    unsigned long long Tt = 0, run2 = inc2 - loc2;
#pragma unroll
    for (int w = 0; w < SB_WARPS; ++w) { const unsigned long long v = sm.scanB[w]; Tt += v; if (w < warp) run2 += v; }
    sb_plan_offset(si, Tt, R53);
    const int mi = (int)m;
    int cs = si.valid ? sb_sys_F(run2, si.U, si.s, mi) : 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int b = b0 + i;
        if (b < b1) {
            s_Pt[b] = run2;
            s_cs[b] = cs;
            if (si.valid && qr[i] != 0ull) run2 += sb_sys_delta(qr[i], si.R, si.r0 + (E - er[i]));
            int ce = si.valid ? sb_sys_F(run2, si.U, si.s, mi) : 0;
            if (b == nt - 1) ce = mi;
            if (cs <= c0 && c0 < ce) sm.bA = b;
            if (c0 + SB_TILE < mi && cs <= c0 + SB_TILE && c0 + SB_TILE < ce) sm.bLast = b;
            cs = ce;
        }
    }
    if (tid == 0) {
        s_Pt[nt] = Tt; s_cs[nt] = mi;
        sm.si = si;
        if (info_out) { *info_out = si; if (!si.valid) atomicOr(err, SB_ERRBIT_ZERO); }
    }
    __syncthreads();                                              // (4)
}

// ancestors of this CTA's children tile from the pool in slot `prev` (scan results in shared memory).  The marker
// array is all -1 on entry and on exit.  `use` = the number of pulls this CTA has done before (mbarrier parity).
// CARRY: anc[] = (parent << xb) | parent state (sb_pull_tile), the parents read their states from xprev.
template <bool S32, bool CARRY>
__device__ __forceinline__ void sb_res_pull(const SbResArgs& a, int prev, const int* xprev, int c0, int use, SbResSmem& sm,
                                            const unsigned long long* s_Pt, const int* s_cs, int anc[SB_ITEMS]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SbStepInfo& si = sm.si;
    const int nt = si.nt, m = (int)si.m;
    const int2 rng = c0 < m ? make_int2(sm.bA, sm.bLast) : make_int2(0, -1);
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
    SbPool pool;
    pool.w32 = a.w32[prev]; pool.x = xprev; pool.wp = a.wp[prev];
    pool.off_w32 = pool.off_x = pool.off_wp = 0; pool.log2n = 31; pool.c_off = 0; pool.tile_off = 0;
    sb_pull_ancestors<false, S32, true, CARRY>(si, pool, a.e[prev], s_Pt, s_cs, c0, rng, 0, (uint32_t)(use & 1), false, 0, sm.ps, anc, a.xb);
}

// exponent keys and the K x K table of grid weights of step t (as k_step_hmm's prologue); ends with a CTA barrier
template <bool FP32>
__device__ __forceinline__ void sb_res_scores(const SbResArgs& a, int t, double* s_l2, uint32_t* s_key, uint32_t* s_qtab, SbResSmem& sm) {
    const int tid = threadIdx.x, K = a.K;
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
    unsigned km = 0u;
    for (int i = 0; i < K; ++i) km = max(km, s_key[i]);
    for (int i = tid; i < K * K; i += SB_THREADS) {
        const int ks = i / K, kx = i - ks * K;
        const unsigned key = s_key[ks];
        const int e = (int)(key >> 6) - SB_KEY_BIAS;
        uint32_t q = 0u;
        if (key) q = FP32 ? sb_q31_f32(__fsub_rn(reinterpret_cast<float*>(s_l2)[kx], (float)e)) : sb_q31_f64(__dsub_rn(s_l2[kx], (double)e));
        s_qtab[i] = q;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned qt = km == 0u ? 0xffffffffu : 0u;                // heaviest below-top state in the top row (see k_step_hmm)
        for (int i = 0; i < K; ++i) if ((s_key[i] >> 6) != (km >> 6)) qt = max(qt, s_qtab[(km & 63u) * K + i]);
        sm.kmax = km; sm.qthr = qt;
    }
}

// TRACE: the diagnostic instantiation (SB_RES_TRACE=1) with the phase stamps; without it they are not compiled in at all
// (predicated off they still cost an issue slot each: 3 % of the kernel's instructions)
template <bool FP32, bool MULTI, bool CARRY, bool TRACE = false>
__global__ void __launch_bounds__(SB_THREADS, 4)
k_sweep_hmm(const __grid_constant__ SbResArgs a) {
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
    const bool tail = c0 + SB_TILE > N;                                    // padding / the retained slot live in this tile
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
    int pulls = 0;
    SbU4 rnd[SB_ITEMS / 4];
#pragma unroll
    for (int p = 0; p < SB_ITEMS / 4; ++p) rnd[p] = sb_stream4_rk(a.rk, (unsigned long long)((base >> 2) + p), 0u, a.iter, SB_P_PROP);
    sb_res_scores<FP32>(a, 0, s_l2, s_key, s_qtab, sm);
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1, prev = cur ^ 1;
        SB_RES_STAMP(0);
        int* xrow = a.xbuf + (size_t)(a.history ? t : cur) * a.S;
        const int* xprev = a.xbuf + (size_t)(a.history ? t - 1 : prev) * a.S;
        // ---- (1) ancestors: the pull from the pool of step t-1
        int anc[SB_ITEMS];
        const int xsh = CARRY ? a.xb : 0;
        if (t > 0) {
            if (sm.si.s >= 32) sb_res_pull<true, CARRY>(a, prev, xprev, c0, pulls, sm, s_Pt, s_cs, anc);
            else               sb_res_pull<false, CARRY>(a, prev, xprev, c0, pulls, sm, s_Pt, s_cs, anc);
            ++pulls;
            if (a.history) {
                int* arow = a.ancbuf + (size_t)(t - 1) * a.S + base;
                if (!tail) {
                    reinterpret_cast<int4*>(arow)[0] = make_int4(anc[0] >> xsh, anc[1] >> xsh, anc[2] >> xsh, anc[3] >> xsh);
                    reinterpret_cast<int4*>(arow)[1] = make_int4(anc[4] >> xsh, anc[5] >> xsh, anc[6] >> xsh, anc[7] >> xsh);
                } else {
#pragma unroll
                    for (int j = 0; j < SB_ITEMS; ++j) {
                        const int i = base + j;
                        if (i < N) arow[j] = anc[j] >> xsh;
                        else if (i == N) arow[j] = N;                     // the retained slot descends from itself
                    }
                }
            }
        }
        SB_RES_STAMP(1);
        // ---- (2) the parents' states: carried by the head markers, or gathered (coherent loads: written by other
        // CTAs in this launch)
        int xp[SB_ITEMS];
        if (CARRY) {
            const int xm = (1 << xsh) - 1;
            if (t > 0 && !tail) {                                        // (CTA-uniform: every tile but the pool's last)
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) xp[j] = anc[j] & xm;
            } else {
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) xp[j] = (t > 0 && base + j < N) ? (anc[j] & xm) : 0;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) xp[j] = (t > 0 && base + j < N) ? __ldcg(xprev + anc[j]) : 0;
        }
        SB_RES_STAMP(2);
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
        if (tail) {                                                      // retained particle (src/pmcmc.jl:59-66) and padding
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                const int i = base + j;
                if (i >= N) x[j] = (i == N && a.xstar) ? __ldg(a.xstar + t) : 0;
            }
        }
        SB_RES_STAMP(3);
        // ---- (4)+(5) weight and quantise: the weights of the top row at once (see k_step_hmm), the exact keys
        // only where no particle of the warp proves the top exponent
        const unsigned kmax = sm.kmax;
        uint32_t q[SB_ITEMS];
        unsigned key = 0u;
        {
            const uint32_t* qtop = s_qtab + (kmax & 63u) * K;
            unsigned wm = 0u;
            if (!tail) {
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) { q[j] = qtop[x[j]]; wm = max(wm, q[j]); }
            } else {
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) { q[j] = base + j < a.n_out ? qtop[x[j]] : 0u; wm = max(wm, q[j]); }
            }
            wm = __reduce_max_sync(0xffffffffu, wm);
            if (wm > sm.qthr) key = kmax;
            else {
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) if (base + j < a.n_out) key = max(key, s_key[x[j]]);
                key = __reduce_max_sync(0xffffffffu, key);
            }
        }
        if ((tid & 31) == 0) sm.red[tid >> 5] = key;
        __syncthreads();
        {
            const uint4 u = reinterpret_cast<const uint4*>(sm.red)[0], v = reinterpret_cast<const uint4*>(sm.red)[1];
            key = max(max(max(u.x, u.y), max(u.z, u.w)), max(max(v.x, v.y), max(v.z, v.w)));
        }
        const int e = key ? (int)(key >> 6) - SB_KEY_BIAS : SB_E_EMPTY;
        if (key != kmax) {                                               // CTA-uniform: the tile exponent is below the top one
            const uint32_t* qrow = s_qtab + (key & 63u) * K;
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) q[j] = (key != 0u && base + j < a.n_out) ? qrow[x[j]] : 0u;
        }
        unsigned tsum = 0;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) tsum += q[j];
        sb_store8_u32(reinterpret_cast<uint32_t*>(xrow) + base, reinterpret_cast<const uint32_t*>(x));
        sb_store8_u32(a.w32[cur] + base, q);
        // tile sum and warp prefixes (one barrier), then the record: every thread's stores above are ordered before
        // thread 0's fence by the barrier, the fence before the record
        sb_warp_total(tsum, sm.tot);
        __syncthreads();
        if (tid <= SB_WARPS) {
            unsigned long long pre = 0;
#pragma unroll
            for (int w = 0; w < SB_WARPS; ++w) if (w < tid) pre += sm.tot[w];
            if (tid < SB_WARPS) a.wp[cur][(size_t)tile * SB_WARPS + tid] = pre;
            else a.Q[cur][tile] = pre;
        }
        if (tid == 0) a.e[cur][tile] = e;
        __syncthreads();
        const unsigned epoch = a.epoch0 + (unsigned)t + 1u;
        if (tid == 0) {
            unsigned long long Qt = 0;
#pragma unroll
            for (int w = 0; w < SB_WARPS; ++w) Qt += sm.tot[w];
            // the record, then the ARRIVE half of the grid barrier (release at GPU scope: the barrier above ordered every
            // thread's stores of this step before it; the records are only read behind the barrier, so no fence of
            // their own).  The wait half comes after the work below, which needs nothing from the other CTAs.
            sb_st_rec_volatile(a.rec[cur] + tile, Qt | ((unsigned long long)(epoch & 0xffffu) << 48),
                               (unsigned long long)(unsigned)e | ((unsigned long long)epoch << 32));
            sb_grid_arrive(a.bar);
            if (TRACE && a.dbg && (t == 10 || t == 11)) {                // diagnostics: when (and where) every CTA published step 10 / 11
                unsigned long long ts_; unsigned smid_;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_));
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
                a.dbg[1024 + (t - 10) * 2048 + tile * 2] = ts_; a.dbg[1024 + (t - 10) * 2048 + tile * 2 + 1] = smid_;
            }
        }
        SB_RES_STAMP(4);
        // ---- independent of the other CTAs, so it overlaps the wait for them: the next step's scores and Philox block
        if (t + 1 < T) {
            sb_res_scores<FP32>(a, t + 1, s_l2, s_key, s_qtab, sm);
#pragma unroll
            for (int p = 0; p < SB_ITEMS / 4; ++p) rnd[p] = sb_stream4_rk(a.rk, (unsigned long long)((base >> 2) + p), (unsigned)(t + 1), a.iter, SB_P_PROP);
        }
        SB_RES_STAMP(5);
        // ---- the barrier of `factor` (src/pmcmc.jl:36-46), wait half, + the normalisation (src/pmcmc.jl:50), by every CTA for itself
        if (tid == 0) sb_grid_wait(a.bar, a.bar0 + (unsigned)(t + 1) * gridDim.x);
        __syncthreads();
        SB_RES_STAMP(6);
        {
            // (records per thread as a compile-time bound: at 2^20 particles two, not SB_RES_MAXPER's four)
            SbStepInfo* io = tile == 0 ? a.info + t : nullptr;
            const unsigned long long R53 = __ldg(a.r53 + t);
            if (nt <= SB_THREADS)          sb_res_scan<1>(a.rec[cur], epoch, nt, (long long)a.n_out, a.N, R53, c0, sm, s_Pt, s_cs, io, a.err);
            else if (nt <= 2 * SB_THREADS) sb_res_scan<2>(a.rec[cur], epoch, nt, (long long)a.n_out, a.N, R53, c0, sm, s_Pt, s_cs, io, a.err);
            else                           sb_res_scan<SB_RES_MAXPER>(a.rec[cur], epoch, nt, (long long)a.n_out, a.N, R53, c0, sm, s_Pt, s_cs, io, a.err);
        }
        SB_RES_STAMP(7);
    }
    if (a.final_anc && c0 < N) {                                         // the resample after the last step (src/pmcmc.jl:42)
        int anc[SB_ITEMS];
        const int prev = (T - 1) & 1;
        const int* xlast = a.xbuf + (size_t)(a.history ? T - 1 : prev) * a.S;
        if (sm.si.s >= 32) sb_res_pull<true, false>(a, prev, xlast, c0, pulls, sm, s_Pt, s_cs, anc);
        else               sb_res_pull<false, false>(a, prev, xlast, c0, pulls, sm, s_Pt, s_cs, anc);
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
