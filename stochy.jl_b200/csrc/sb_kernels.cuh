// sb_kernels.cuh -- the kernels of libstochy_b200 (sm_100a).  See DESIGN.md for the per-kernel
// byte counts and rooflines.  Reference lines each kernel replaces are cited at the kernel.
#pragma once
#include "stochy_b200.h"
#include "sb_device.cuh"
#include <cfloat>
#include <cmath>
#include <type_traits>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// =====================================================================================
// k_quantise: host/device supplied weights -> (w32, tile exponent, tile sum).
// Replaces `weights = Float64[exp(p.path[end].score) ...]` (src/pmcmc.jl:55,64) for the hooks.
// KIND 0: linear fp64, 1: fp64 log-weights, 2: fp32 log-weights.   Bytes: read W, write 4.
// =====================================================================================
template <int KIND>
__global__ void __launch_bounds__(SB_THREADS)
k_quantise(const void* __restrict__ data, long long n, uint32_t* __restrict__ w32, int* __restrict__ e_out,
           unsigned long long* __restrict__ Q_out, unsigned long long* __restrict__ wp_out, int* __restrict__ emax,
           unsigned* __restrict__ err) {
    __shared__ unsigned long long s_u64[SB_WARPS + 1];
    __shared__ double s_f64[SB_WARPS];
    __shared__ int s_i32[SB_WARPS];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * SB_TILE + tid * SB_ITEMS;
    double vd[SB_ITEMS];
    float vf[SB_ITEMS];
    bool bad = false;
    int e = SB_E_EMPTY;
    const bool full = base + SB_ITEMS <= n;
    if (KIND == 2) {
        const float* p = (const float*)data;
        float mx = -INFINITY;
        float in[SB_ITEMS];
        if (full) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p + base));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p + base) + 1);
            in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w; in[4] = b.x; in[5] = b.y; in[6] = b.z; in[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) in[j] = (base + j < n) ? __ldg(p + base + j) : -INFINITY;
        }
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) {
            float lw = in[j];
            float l2 = __fmul_rn(lw, SB_LOG2E_F);
            if (isnan(l2) || l2 == INFINITY) { bad = true; l2 = -INFINITY; }
            vf[j] = l2;
            mx = fmaxf(mx, l2);
        }
        mx = sb_block_max_f32(mx, (float*)s_f64);
        if (mx > -INFINITY) e = (int)ceilf(mx);
    } else {
        const double* p = (const double*)data;
        double in[SB_ITEMS];
        if (full) {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; j += 2) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(p + base + j));
                in[j] = a.x; in[j + 1] = a.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) in[j] = (base + j < n) ? __ldg(p + base + j) : (KIND == 1 ? -INFINITY : 0.0);
        }
        if (KIND == 1) {
            double mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                double lw = in[j];
                double l2 = __dmul_rn(lw, SB_LOG2E_D);
                if (isnan(l2) || l2 == INFINITY) { bad = true; l2 = -INFINITY; }
                vd[j] = l2;
                mx = fmax(mx, l2);
            }
            mx = sb_block_max_f64(mx, s_f64);
            if (mx > -INFINITY) e = (int)ceil(mx);
        } else {
            int me = SB_E_EMPTY;
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                double w = in[j];
                if (!(w >= 0.0) || isinf(w)) { bad = true; w = 0.0; }
                vd[j] = w;
                if (w > 0.0) { int ex; frexp(w, &ex); me = max(me, ex); }
            }
            e = sb_block_max_i32(me, s_i32);
        }
    }
    uint32_t q[SB_ITEMS];
    unsigned tsum = 0;
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) {
        uint32_t v = 0;
        if (e != SB_E_EMPTY) {
            if (KIND == 0)      v = (uint32_t)__double2ull_rz(scalbn(vd[j], SB_QBITS - e));
            else if (KIND == 1) v = sb_q31_f64(__dsub_rn(vd[j], (double)e));
            else                v = sb_q31_f32(__fsub_rn(vf[j], (float)e));
        }
        q[j] = v;
        tsum += v;
    }
    sb_store8_u32(w32 + base, q);                 // w32 is padded to whole tiles
    sb_block_sum_wp(tsum, s_u64, wp_out + (size_t)blockIdx.x * SB_WARPS, Q_out + blockIdx.x);
    if (tid == 0) {
        e_out[blockIdx.x] = e;
        // e != EMPTY means some weight of the tile is positive: its grid value is at least 2^26, so the tile has mass
        if (e != SB_E_EMPTY) atomicMax(emax, e);
    }
    if (bad) atomicOr(err, SB_ERRBIT_NAN);
}

// =====================================================================================
// k_tilescan: exact integer exclusive scan of the rescaled tile sums; the systematic offset u;
// first child of every tile; log(mean weight) of the pool.  One CTA (latency-bound, O(n/2048)).
// Replaces `ps = weights/sum(weights)` (src/pmcmc.jl:50); the log-evidence term is the north-star
// extension.  Deterministic: integer adds only.
// =====================================================================================
#define SB_SCAN_THREADS 1024
#define SB_SCAN_CTAS    8      // portable cluster: 8 SMs x 1024 threads = one tile per thread at 2^24 particles
#define SB_SCAN_CTAS_MAX 16    // B200 allows 16 (opt-in): used for pools above 2^26 particles

// The scan runs as ONE thread-block cluster: block-local warp-shuffle scans, then the block totals
// are exchanged through distributed shared memory (a cluster barrier is ~0.5 us; a second kernel
// launch or a single-SM scan of 8192 tiles is 10-30 us).  All arithmetic is integer: any association
// order gives the same bits.  Every cluster-wide exchange has its own shared-memory slot, so the only
// cluster barriers are the ones that publish a slot (no guard barriers).
struct SbScanSmem {
    unsigned long long warp_tot[32];
    unsigned long long blk_tot[2];   // one slot per scan; read by the other CTAs of the cluster
    int blk_max;                     // idem
    int warp_max[32];
    unsigned long long base, total;  // broadcast inside the CTA
    int gmax;
};

__device__ __forceinline__ int sb_cluster_max(int v, SbScanSmem& sm, cg::cluster_group& cl) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = (int)cl.num_blocks();
#pragma unroll
    for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    if (lane == 0) sm.warp_max[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = sm.warp_max[lane];
#pragma unroll
        for (int d = 16; d; d >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, d));
        if (lane == 0) sm.blk_max = w;
    }
    cl.sync();
    if (threadIdx.x < 32) {
        int r = lane < nb ? *cl.map_shared_rank(&sm.blk_max, lane) : SB_E_EMPTY;
#pragma unroll
        for (int d = 16; d; d >>= 1) r = max(r, __shfl_xor_sync(0xffffffffu, r, d));
        if (lane == 0) sm.gmax = r;
    }
    __syncthreads();
    return sm.gmax;
}

// exclusive prefix of `loc` over all threads of the cluster (thread order = rank * 1024 + tid); `which`
// selects the exchange slot (each slot is used once per kernel)
__device__ __forceinline__ unsigned long long sb_cluster_excl_scan(unsigned long long loc, SbScanSmem& sm,
                                                                    cg::cluster_group& cl, int which, unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = (int)cl.num_blocks();
    unsigned long long inc = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = sb_shfl_up_u64(inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();                             // warp_tot / base / total of the previous scan are no longer read
    if (lane == 31) sm.warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long wt = sm.warp_tot[lane];
        unsigned long long wi = wt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long o = sb_shfl_up_u64(wi, d);
            if (lane >= d) wi += o;
        }
        sm.warp_tot[lane] = wi - wt;             // exclusive prefix of the warp totals
        if (lane == 31) sm.blk_tot[which] = wi;
    }
    cl.sync();
    if (threadIdx.x < 32) {
        const unsigned rank = cl.block_rank();
        unsigned long long bt = lane < nb ? *cl.map_shared_rank(&sm.blk_tot[which], lane) : 0ull;
        unsigned long long base = lane < (int)rank ? bt : 0ull, tot = bt;
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const uint32_t blo = __shfl_xor_sync(0xffffffffu, (uint32_t)base, d), bhi = __shfl_xor_sync(0xffffffffu, (uint32_t)(base >> 32), d);
            const uint32_t tlo = __shfl_xor_sync(0xffffffffu, (uint32_t)tot, d), thi = __shfl_xor_sync(0xffffffffu, (uint32_t)(tot >> 32), d);
            base += ((unsigned long long)bhi << 32) | blo;
            tot += ((unsigned long long)thi << 32) | tlo;
        }
        if (lane == 0) { sm.base = base; sm.total = tot; }
    }
    __syncthreads();
    *total = sm.total;
    return sm.base + sm.warp_tot[warp] + inc - loc;
}

// Sharded pool: GPU g's tile records (e, Q) live in section g of the record arrays of ITS arena (arrays are
// laid out by GLOBAL tile index in every arena).  The (replicated, deterministic) scan of every GPU reads
// section g from GPU g -- local HBM or NVLink peer loads, one round trip, all loads in flight together --
// behind a flag barrier, and fills the foreign sections of its own exponent array for the pull.
struct SbScanPeers {
    const char* base[SB_MAX_PEERS];
    unsigned long long off_flags;       // unsigned[SB_MAX_PEERS]: flags[g] = last epoch GPU g has finished
    unsigned long long off_e, off_Q;    // record arrays of the pool slot inside every arena
    int world, self;
    int log2t;                          // tiles per GPU = 2^log2t
    unsigned epoch;                     // wait until every flags[g] >= epoch
    int rec_slot;                       // >= 0: every GPU's summary record arrived with its flag (sb_signal_done), in this slot
};
#define SB_SCAN_MAXPER 8

// PER > 0: every thread keeps its PER consecutive tiles in registers (PER = 1: 2^24 particles, PER = 8:
// 2^27); PER == 0: any size, tile records re-read from global memory in every pass.
// emax (single GPU): the largest tile exponent, accumulated with atomicMax by the kernels that produced the
// pool; consumed and cleared here (saves one cluster-wide reduction).  Sharded pools take it from the GPUs' summary
// records (sp.rec_slot >= 0: they arrived with the flags, sb_signal_done) or, without records, reduce e.
// The cluster size is given at launch (8, or 16 for large pools).
template <int PER, bool PEERS>
__global__ void __launch_bounds__(SB_SCAN_THREADS)
k_tilescan(const int* __restrict__ e_g, const unsigned long long* __restrict__ Q_g, int nt, long long n, long long m,
           const unsigned long long* __restrict__ r53_ptr, SbStepInfo* __restrict__ info,
           unsigned long long* __restrict__ P, unsigned long long* __restrict__ Pt, int* __restrict__ child_start,
           int* __restrict__ first_tile, int want_exact, unsigned* __restrict__ err, int* __restrict__ emax,
           const __grid_constant__ SbScanPeers sp) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ SbScanSmem sm;
    const unsigned rank = cl.block_rank();
    const int nthreads = SB_SCAN_THREADS * (int)cl.num_blocks();
    const int gtid = (int)rank * SB_SCAN_THREADS + threadIdx.x;
    sb_grid_dep_launch();                                     // the next step kernel may start its prologue
    const unsigned long long R53 = __ldg(r53_ptr);            // host-drawn systematic offset (53-bit uniform)
    const int per = PER > 0 ? PER : (nt + nthreads - 1) / nthreads;
    sb_grid_dep_wait();                                       // the pool's tile records are complete
    const int b0 = min(nt, gtid * per), b1 = min(nt, b0 + per);
    if (PEERS) {
        // every peer has finished the step kernel that produced this pool (and therefore also every
        // read of the pool before it): the flags are raised over NVLink by the last CTA of that kernel
        if (threadIdx.x < sp.world) {
            const volatile unsigned* f = reinterpret_cast<const volatile unsigned*>(sp.base[sp.self] + sp.off_flags) + threadIdx.x;
            const long long t0 = clock64();
            while ((int)(*f - sp.epoch) < 0) {
                if (clock64() - t0 > 8000000000ll) { atomicOr(err, SB_ERRBIT_PEER); break; }   // ~4 s: a peer died
                __nanosleep(200);
            }
        }
        __syncthreads();
    }
    unsigned long long qr[PER > 0 ? PER : 1];
    int er[PER > 0 ? PER : 1];
    if (PER > 0) {
#pragma unroll
        for (int i = 0; i < (PER > 0 ? PER : 1); ++i) {
            const int b = b0 + i;
            qr[i] = 0ull; er[i] = SB_E_EMPTY;
            if (b < b1) {
                if (PEERS) {
                    const char* owner = sp.base[b >> sp.log2t];
                    qr[i] = __ldcv(reinterpret_cast<const unsigned long long*>(owner + sp.off_Q) + b);
                    er[i] = __ldcv(reinterpret_cast<const int*>(owner + sp.off_e) + b);
                } else { qr[i] = __ldcg(Q_g + b); er[i] = __ldcg(e_g + b); }   // NOT __ldg: see sb_grid_dep_wait
            }
        }
        if (PEERS) {                             // the pull reads exponents of any tile from the local array
            int* e_mine = reinterpret_cast<int*>(const_cast<char*>(sp.base[sp.self]) + sp.off_e);
#pragma unroll
            for (int i = 0; i < (PER > 0 ? PER : 1); ++i) {
                const int b = b0 + i;
                if (b < b1 && (b >> sp.log2t) != sp.self) e_mine[b] = er[i];
            }
        }
    }
#define SB_TS_FOR(body)                                                                        \
    if (PER > 0) { _Pragma("unroll") for (int i_ = 0; i_ < (PER > 0 ? PER : 1); ++i_) { const int b = b0 + i_; if (b < b1) { const unsigned long long qv = qr[i_]; const int ev = er[i_]; body } } } \
    else for (int b = b0; b < b1; ++b) { const unsigned long long qv = __ldcg(Q_g + b); const int ev = __ldcg(e_g + b); body }
    // global exponent
    int E = SB_E_EMPTY;
    bool spec = false;
    unsigned long long total = 0, run = 0;
    if (PEERS && sp.rec_slot >= 0) {
        // the GPUs' summaries: the exponent maximum without a reduction and, when it is the exponent the step kernels
        // assumed, the exact total without the first pass (sb_signal_done; integer adds: the same bits)
        const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(sp.base[sp.self] + sp.off_flags + SB_REC_OFF) +
                                        2 * (sp.rec_slot * SB_MAX_PEERS);
        spec = !want_exact;
        int et0 = SB_E_EMPTY;
        for (int g = 0; g < sp.world; ++g) {
            const unsigned long long w0 = __ldcv(rec + 2 * g);
            E = max(E, (int)(unsigned)w0);
            if (g == 0) et0 = (int)(unsigned)(w0 >> 32);
            spec = spec && (int)(unsigned)(w0 >> 32) == et0;
            total += __ldcv(rec + 2 * g + 1);
        }
        spec = spec && et0 == E && E != SB_E_EMPTY;
    } else if (PEERS) {
        SB_TS_FOR((void)b; if (qv != 0ull) E = max(E, ev);)
        E = sb_cluster_max(E, sm, cl);
    } else {
        E = __ldcg(emax);                        // cleared below, after the first cluster barrier
    }
    const int LS = 24 - (nt <= 1 ? 0 : 32 - __clz(nt - 1));  // LS = 24 - ceil_log2(nt)
    // exact CDF of tile sums (log-evidence; multinomial resampling)
    // single GPU: the step kernel of the discrete model has accumulated this very sum at the exponent it expected
    // (sb_spec_term); if that exponent is E, the first pass and its cluster scan are skipped
    if (!PEERS && !want_exact) {
        const int sE = __ldcg(emax + 1);
        const unsigned long long st = __ldcg(reinterpret_cast<const unsigned long long*>(emax + 2));
        spec = sE == E && E != SB_E_EMPTY;
        total = st;
    }
    if (!spec) {
        unsigned long long loc = 0;
        SB_TS_FOR((void)b; if (qv != 0ull) { const int sh = E - ev; loc += sh >= 63 ? 0ull : ((qv << LS) >> sh); })
        run = sb_cluster_excl_scan(loc, sm, cl, 0, &total);
    }
    if (want_exact) {
        SB_TS_FOR(P[b] = run; if (qv != 0ull) { const int sh = E - ev; run += sh >= 63 ? 0ull : ((qv << LS) >> sh); })
        if (gtid == 0) P[nt] = total;
    }
    // plan of the systematic rescale (every thread computes the same few integers)
    SbStepInfo si;
    sb_plan_rescale(si, total, n, m, E, LS, nt);
    // rescaled CDF of tile sums (systematic)
    unsigned long long loc2 = 0;
    if (si.valid) { SB_TS_FOR((void)b; if (qv != 0ull) loc2 += sb_sys_delta(qv, si.R, si.r0 + (E - ev));) }
    unsigned long long Tt;
    unsigned long long run2 = sb_cluster_excl_scan(loc2, sm, cl, 1, &Tt);
    if (!PEERS && gtid == 0) {                                // every thread of the cluster has read them (the barrier inside the scan above)
        emax[0] = SB_E_EMPTY; emax[1] = SB_E_EMPTY;
        *reinterpret_cast<unsigned long long*>(emax + 2) = 0ull;
    }
    sb_plan_offset(si, Tt, R53);
    if (gtid == 0) {
        if (!si.valid) atomicOr(err, SB_ERRBIT_ZERO);
        *info = si;
    }
    const int mi = (int)m;
    int cs = si.valid ? sb_sys_F(run2, si.U, si.s, mi) : 0;
    SB_TS_FOR(
        Pt[b] = run2;
        child_start[b] = cs;
        if (si.valid && qv != 0ull) run2 += sb_sys_delta(qv, si.R, si.r0 + (E - ev));
        int ce = si.valid ? sb_sys_F(run2, si.U, si.s, mi) : 0;
        if (b == nt - 1) ce = mi;
        /* children-tile boundaries k * SB_TILE in [cs, ce) start inside parent tile b */
        for (int k = (cs + SB_TILE - 1) / SB_TILE; (long long)k * SB_TILE < ce; ++k) first_tile[k] = b;
        cs = ce;
    )
    if (gtid == 0) { Pt[nt] = Tt; child_start[nt] = mi; }
    cl.sync();                                   // no CTA may exit while its shared memory can still be read
#undef SB_TS_FOR
}

// =====================================================================================
// k_pull: ancestors only (hooks, and the final resample of a sweep).
// Replaces the N calls of rand(ps) (src/rand.jl:2-13 via src/pmcmc.jl:51) by ONE merge of the
// sorted child thresholds with the integer CDF.  Bytes per child: read 4 (w32), write 4 (anc).
// =====================================================================================
template <bool PEERS>
__global__ void __launch_bounds__(SB_THREADS)
k_pull(const SbStepInfo* __restrict__ info, const __grid_constant__ SbPool pool, const int* __restrict__ e,
       const unsigned long long* __restrict__ Pt, const int* __restrict__ child_start,
       const int* __restrict__ first_tile, int* __restrict__ anc_out) {
    __shared__ SbPullSmem ps;
    const SbStepInfo si = *info;
    const int c0 = pool.c_off + blockIdx.x * SB_TILE;             // global child index
    if (c0 >= (int)si.m) return;
    if (threadIdx.x == 0) sb_mbar_init(&ps.bar, 1);
    const int2 rng = sb_pull_range(si, first_tile, c0);
    __syncthreads();
    sb_pull_first<PEERS>(si, pool, e, Pt, child_start, c0, rng, ps);
    int anc[SB_ITEMS];
    if (si.s >= 32) sb_pull_ancestors<PEERS, true>(si, pool, e, Pt, child_start, c0, rng, 0, 0u, false, 0, ps, anc);
    else            sb_pull_ancestors<PEERS, false>(si, pool, e, Pt, child_start, c0, rng, 0, 0u, false, 0, ps, anc);
    const int base = c0 + threadIdx.x * SB_ITEMS;
    int* out = anc_out + (size_t)blockIdx.x * SB_TILE + threadIdx.x * SB_ITEMS;      // local children row
    if (base + SB_ITEMS <= (int)si.m) {
        reinterpret_cast<int4*>(out)[0] = make_int4(anc[0], anc[1], anc[2], anc[3]);
        reinterpret_cast<int4*>(out)[1] = make_int4(anc[4], anc[5], anc[6], anc[7]);
    } else {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) if (base + j < (int)si.m) out[j] = anc[j];
    }
}

// The same as a PERSISTENT grid (pools of many tiles: the resampling hooks at bench size): every CTA walks its children
// tiles with the step kernels' pipeline -- candidate tiles looked up per chunk, the next tile's weights (TMA) and records
// (cp.async) requested while this tile's markers are filled, the next tile's boundary table built by one warp -- instead
// of paying the ramp (mbarrier init, first copies, boundary table, three dependent round trips) once per tile.
template <bool PEERS>
__global__ void __launch_bounds__(SB_THREADS, 5)
k_pull_persistent(const SbStepInfo* __restrict__ info, const __grid_constant__ SbPool pool, const int* __restrict__ e,
                  const unsigned long long* __restrict__ Pt, const int* __restrict__ child_start,
                  const int* __restrict__ first_tile, int* __restrict__ anc_out, int ntiles) {
    __shared__ SbPullSmem ps;
    const SbStepInfo si = *info;
    const int tid = threadIdx.x;
    if (tid == 0) sb_mbar_init(&ps.bar, 1);
    __syncthreads();
    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
        if ((k % SB_FT_CHUNK) == 0) sb_pull_ranges<PEERS>(si, pool, e, Pt, child_start, first_tile, tile, ntiles, k == 0, ps);
        else __syncthreads();                                    // the previous tile's markers are read and reset, this tile's boundary table is built
        const int kk = k % SB_FT_CHUNK;
        const int2 rng = ps.ft[kk];
        const bool more = tile + (int)gridDim.x < ntiles;
        const int nb = ps.ft[kk + 1].x;
        const int c0 = pool.c_off + tile * SB_TILE;              // global child index
        int anc[SB_ITEMS];
        if (si.s >= 32) sb_pull_ancestors<PEERS, true>(si, pool, e, Pt, child_start, c0, rng, k & 1, (uint32_t)(k & 1), more, nb, ps, anc);
        else            sb_pull_ancestors<PEERS, false>(si, pool, e, Pt, child_start, c0, rng, k & 1, (uint32_t)(k & 1), more, nb, ps, anc);
        if (more && (tid >> 5) == SB_META_WARP) {
            const int c0n = c0 + (int)gridDim.x * SB_TILE;
            if (si.s >= 32) sb_pull_seg<true>(si, c0n, ps.ft[kk + 1], (k + 1) & 1, ps);
            else            sb_pull_seg<false>(si, c0n, ps.ft[kk + 1], (k + 1) & 1, ps);
        }
        const int base = c0 + tid * SB_ITEMS;
        int* out = anc_out + (size_t)tile * SB_TILE + tid * SB_ITEMS;      // local children row
        if (base + SB_ITEMS <= (int)si.m) {
            reinterpret_cast<int4*>(out)[0] = make_int4(anc[0], anc[1], anc[2], anc[3]);
            reinterpret_cast<int4*>(out)[1] = make_int4(anc[4], anc[5], anc[6], anc[7]);
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) if (base + j < (int)si.m) out[j] = anc[j];
        }
    }
}

// =====================================================================================
// Multinomial on the grid: materialise the integer CDF, then one upper_bound per draw.
// Same semantics as src/rand.jl:2-13 (first i with r < acc_i, last index as fallback).
// =====================================================================================
__global__ void __launch_bounds__(SB_THREADS)
k_cdf(const SbStepInfo* __restrict__ info, const uint32_t* __restrict__ w32, const int* __restrict__ e,
      const unsigned long long* __restrict__ P, unsigned long long* __restrict__ cdf) {
    __shared__ unsigned long long s_scan[SB_WARPS + 1];
    const SbStepInfo si = *info;
    const int b = blockIdx.x, tid = threadIdx.x;
    uint32_t w[SB_ITEMS];
    sb_load8_u32(w32 + (size_t)b * SB_TILE + tid * SB_ITEMS, w);
    unsigned long long tsum = 0;
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) tsum += w[j];
    unsigned long long tot;
    unsigned long long c = sb_block_excl_scan_u64(tsum, s_scan, &tot);
    const int sh = tot ? si.E - e[b] : 63;
    const unsigned long long Pb = P[b];
    const long long base = (long long)b * SB_TILE + tid * SB_ITEMS;
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) {
        c += w[j];
        if (base + j < si.n) cdf[base + j] = Pb + (sh >= 63 ? 0ull : ((c << si.LS) >> sh));
    }
}

// r_in != nullptr: caller-supplied 53-bit uniforms (hook); else Philox stream (purpose RESAMPLE).
__global__ void __launch_bounds__(256)
k_multinomial_search(const SbStepInfo* __restrict__ info, const unsigned long long* __restrict__ cdf,
                     const double* __restrict__ r_in, unsigned long long seed, unsigned t, unsigned iter,
                     int* __restrict__ anc_out) {
    const SbStepInfo si = *info;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= si.m) return;
    unsigned long long R;
    if (r_in) R = __double2ull_rz(__dmul_rn(r_in[j], 0x1p53));
    else {
        const SbU4 o = sb_stream4(seed, (unsigned long long)j >> 1, t, iter, SB_P_RESAMPLE);
        R = (j & 1) ? sb_r53(o.z, o.w) : sb_r53(o.x, o.y);
    }
    // thr = floor(R * total / 2^53), exact in 128 bits
    const unsigned long long hi = __umul64hi(R, si.total), lo = R * si.total;
    const unsigned long long thr = (hi << 11) | (lo >> 53);
    long long a = 0, b = si.n - 1;
    while (a < b) {
        const long long mid = a + ((b - a) >> 1);
        if (thr < __ldg(cdf + mid)) b = mid; else a = mid + 1;
    }
    anc_out[j] = (int)a;
}

// =====================================================================================
// Literal reference numerics (src/pmcmc.jl:48-52 + src/rand.jl:2-13): sequentially rounded fp64
// sum, ps = w/sum, sequentially rounded running sum.  One CTA; thread 0 owns the two dependent
// add chains, the other threads only stage data through shared memory.  Parity path (small N).
// =====================================================================================
#define SB_LIT_CHUNK 2048
__global__ void __launch_bounds__(256)
k_literal_cdf(const double* __restrict__ w, long long n, double* __restrict__ acc_out, unsigned* __restrict__ err) {
    __shared__ double buf[SB_LIT_CHUNK];
    __shared__ double s_sum;
    const int tid = threadIdx.x;
    double s = 0.0;
    for (long long c0 = 0; c0 < n; c0 += SB_LIT_CHUNK) {
        const int len = (int)min((long long)SB_LIT_CHUNK, n - c0);
        for (int i = tid; i < len; i += blockDim.x) buf[i] = w[c0 + i];
        __syncthreads();
        if (tid == 0) for (int i = 0; i < len; ++i) s = __dadd_rn(s, buf[i]);
        __syncthreads();
    }
    if (tid == 0) s_sum = s;
    __syncthreads();
    s = s_sum;
    double a = 0.0;
    bool bad = false;
    for (long long c0 = 0; c0 < n; c0 += SB_LIT_CHUNK) {
        const int len = (int)min((long long)SB_LIT_CHUNK, n - c0);
        for (int i = tid; i < len; i += blockDim.x) buf[i] = __ddiv_rn(w[c0 + i], s);   // pmcmc.jl:50
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < len; ++i) {
                a = __dadd_rn(a, buf[i]);                                             // rand.jl:7
                if (isnan(a) && (c0 + i < n - 1 || n == 1)) bad = true;               // rand.jl:8,11
                buf[i] = a;
            }
        }
        __syncthreads();
        for (int i = tid; i < len; i += blockDim.x) acc_out[c0 + i] = buf[i];
        __syncthreads();
    }
    if (tid == 0 && bad) atomicOr(err, SB_ERRBIT_NAN);
}

__global__ void __launch_bounds__(256)
k_literal_search(const double* __restrict__ acc, long long n, const double* __restrict__ r_in, long long m,
                 unsigned long long seed, unsigned t, unsigned iter, int* __restrict__ anc_out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double r;
    if (r_in) r = r_in[j];
    else {
        const SbU4 o = sb_stream4(seed, (unsigned long long)j >> 1, t, iter, SB_P_RESAMPLE);
        r = (j & 1) ? sb_u53(o.z, o.w) : sb_u53(o.x, o.y);
    }
    long long a = 0, b = n - 1;                       // first i in [0, n-1) with r < acc[i], else n-1
    while (a < b) {
        const long long mid = a + ((b - a) >> 1);
        if (r < __ldg(acc + mid)) b = mid; else a = mid + 1;
    }
    anc_out[j] = (int)a;
}

// =====================================================================================
// k_gather: dst[i] = src[anc[i]]  (deepcopy(particles[...]) src/pmcmc.jl:51).
// Bytes per child: read 4 (anc) + S (state, monotone for systematic) + write S.
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
k_gather(const T* __restrict__ src, const int* __restrict__ anc, long long m, T* __restrict__ dst) {
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 + 4 <= m) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(anc + i0));
        T v0 = __ldg(src + a.x), v1 = __ldg(src + a.y), v2 = __ldg(src + a.z), v3 = __ldg(src + a.w);
        dst[i0] = v0; dst[i0 + 1] = v1; dst[i0 + 2] = v2; dst[i0 + 3] = v3;
    } else {
        for (long long i = i0; i < m; ++i) dst[i] = src[anc[i]];
    }
}
template <>
__global__ void __launch_bounds__(256)
k_gather<uint32_t>(const uint32_t* __restrict__ src, const int* __restrict__ anc, long long m, uint32_t* __restrict__ dst) {
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 + 4 <= m) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(anc + i0));
        uint4 v;
        v.x = __ldg(src + a.x); v.y = __ldg(src + a.y); v.z = __ldg(src + a.z); v.w = __ldg(src + a.w);
        *reinterpret_cast<uint4*>(dst + i0) = v;
    } else {
        for (long long i = i0; i < m; ++i) dst[i] = src[anc[i]];
    }
}

// =====================================================================================
// Log-sum-exp: warp-shuffle + shared-memory reduction of (max, sum exp(x - max)) pairs.
// Stabilised replacement for `sum(exp(score))` (src/pmcmc.jl:50,55); extension D2/H3.
// =====================================================================================
struct SbLse { double m, s; };
__device__ __forceinline__ SbLse sb_lse_combine(SbLse a, SbLse b) {
    if (a.m == -INFINITY) return b;
    if (b.m == -INFINITY) return a;
    if (a.m >= b.m) return SbLse{a.m, a.s + b.s * exp(b.m - a.m)};
    return SbLse{b.m, b.s + a.s * exp(a.m - b.m)};
}
__device__ __forceinline__ SbLse sb_lse_block(SbLse v, SbLse* sm) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        SbLse o;
        o.m = __shfl_xor_sync(0xffffffffu, v.m, d);
        o.s = __shfl_xor_sync(0xffffffffu, v.s, d);
        v = sb_lse_combine(v, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    SbLse r = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = sb_lse_combine(r, sm[w]);
    return r;
}
template <typename T>
__global__ void __launch_bounds__(256)
k_lse_partial(const T* __restrict__ lw, long long n, SbLse* __restrict__ partial, unsigned* __restrict__ err) {
    __shared__ SbLse sm[8];
    // two passes over this CTA's contiguous slab: max first, then the sum (slab stays in L1/L2)
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long i0 = (long long)blockIdx.x * per, i1 = min(n, i0 + per);
    double mx = -INFINITY;
    bool bad = false;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const double v = (double)__ldg(lw + i);
        if (isnan(v) || v == INFINITY) bad = true;
        mx = fmax(mx, v);
    }
    __shared__ double s_mx[8];
    mx = sb_block_max_f64(mx, s_mx);
    double s = 0.0;
    if (mx > -INFINITY)
        for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) s += exp((double)__ldg(lw + i) - mx);
    SbLse r = sb_lse_block(SbLse{mx > -INFINITY ? mx : -INFINITY, s}, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
    if (bad) atomicOr(err, SB_ERRBIT_NAN);
}
__global__ void __launch_bounds__(256)
k_lse_final(const SbLse* __restrict__ partial, int np, double* __restrict__ out) {
    __shared__ SbLse sm[8];
    SbLse v{-INFINITY, 0.0};
    for (int i = threadIdx.x; i < np; i += blockDim.x) v = sb_lse_combine(v, partial[i]);
    v = sb_lse_block(v, sm);
    if (threadIdx.x == 0) *out = v.m == -INFINITY ? -INFINITY : v.m + log(v.s);
}

// =====================================================================================
// The fused step kernels.  One CTA works through tiles of 2048 children (persistent grid, model
// tables staged into shared memory once per CTA).  Per tile:
//   (1) resample: pull this tile's ancestors from the previous pool (or explicit / identity),
//   (2) gather the parents' states,
//   (3) propagate through the model's ERP draw with a counter-based Philox stream,
//   (4) weight with the step's factor / observe score,
//   (5) quantise the weights onto the grid and emit the tile exponent, tile sum and warp prefixes.
// Replaces, per particle-step: sample -> rand(e) (src/pmcmc.jl:32-34), the model body,
// factor's Step push (src/pmcmc.jl:37), and resample + deepcopy (src/pmcmc.jl:42,48-66).
// Algorithmic bytes per particle-step: read w32 (4) + read parent state (S) + write state (S)
// + write w32 (4) [+ 4 for the ancestor row when history is kept].
//
// Synchronisation per children tile of k_step_hmm (pull instantiation): ONE CTA barrier (all head markers written)
// and ONE split-phase mbarrier (sm.bar2, one arrival per warp): a warp arrives once its exponent key is published
// -- warp SB_META_WARP also the next tile's boundary table -- and waits at the start of the next tile; the tile exponent itself
// needs the wait only when the warp's own key is below the model's largest key (sm.kmax).  What the split barrier
// lets run ahead is buffered by tile parity (sm.tot, the meta slots); everything else is rewritten only behind the
// next full barrier ("How k_step_hmm is built").  k_step_lg follows the same protocol; its shortcut is
// the mode of the observation density (no log2-weight exceeds log2e * lc).
// =====================================================================================
#define SB_ANC_NONE     0   // first step: no parents
#define SB_ANC_PULL     1   // systematic resampling fused in
#define SB_ANC_EXPLICIT 2   // ancestors already materialised (multinomial / literal)
#define SB_ANC_IDENTITY 3   // previous step had no factor statement

#define SB_LG_BUCKETS 8     // inverse-CDF lookup: 256 buckets per transition row
#define SB_KEY_BIAS (1 << 24)
#ifndef SB_ABLATE
#define SB_ABLATE 0            // diagnostics only: 1 = no Philox, 2 = no gather, 3 = no pull (identity ancestors)
#endif

struct SbHmmStepArgs {
    // previous pool
    SbPool pool;
    const int* e_prev;          // [global tiles]
    const unsigned long long* Pt;
    const int* child_start;
    const int* first_tile;
    const SbStepInfo* info_prev;
    const int* anc_in;          // local children row
    int* anc_out;               // history row of the resample after step t-1 (or nullptr)
    // this step's outputs (local indices)
    int* x_out;
    uint32_t* w32_out;
    SbRecOut rec;               // tile exponents and sums
    SbSignal sig;               // sharded: tell the peers when this GPU's share of the pool is complete
    unsigned long long* wp_out;
    double* lw_out;             // optional (store_logw)
    double* wlit_out;           // optional (literal mode): exp(score) from the host-built table
    // model tables (global memory; staged into shared memory)
    const uint32_t* tab;        // [rows * 256] inverse-CDF lookup words (see sb_model_discrete_hmm)
    const uint32_t* thr;        // [rows * K] ceil(2^32 * sequential running sum) of the transition rows
    const double* logF_t;       // [K]
    const double* expF_t;       // [K] or nullptr
    const int* xstar;           // retained trajectory [T] or nullptr
    long long N;                // global particle count
    int n_out;                  // N (+1 with a retained particle)
    int ntiles;                 // local tiles to produce
    int nt_pool;                // tiles of the whole pool (all GPUs): the scan's LS comes from it (sb_spec_add)
    int K, rows, t;
    int xb;                     // CARRY instantiations: state bits of a head marker ((parent << xb) | state)
    unsigned iter;
    unsigned long long seed;
    SbRoundKeys rk;             // Philox round keys of `seed`
    unsigned* err;
};

struct SbHmmSmem {
    SbPullSmem ps;
    double red[SB_WARPS];
    unsigned long long tot[2][SB_WARPS];   // warp totals of the tile just produced, by tile parity: with the split barrier a
                                           // fast warp may finish tile k+1 before the flushing warp has written out the totals of tile k
    unsigned long long bar2;    // split-phase barrier of the fused HMM kernel: one arrival per warp and children tile
    int e_tile[2];              // exponent of the tile just produced, by tile parity (speculative total: sb_spec_term)
    unsigned long long spec_acc;   // this CTA's share of the speculative total (one thread reads and writes it: no registers
                                   // of the other 255 are spent on it)
    unsigned kmax;              // largest exponent key any state of this step can have
    unsigned qthr;              // largest grid weight, in kmax's row of the weight table, of a state BELOW the top exponent:
                                // a particle weighing more proves that the tile exponent is the top one
};

// this CTA's share of the speculative exact total (sb_spec_term), by the one thread that holds a tile sum
__device__ __forceinline__ void sb_spec_add(SbHmmSmem& sm, unsigned long long Q, int parity, int ntiles) {
    const int E_top = (int)(sm.kmax >> 6) - SB_KEY_BIAS;
    const int LS = 24 - (ntiles <= 1 ? 0 : 32 - __clz(ntiles - 1));
    sm.spec_acc += sb_spec_term(Q, sm.e_tile[parity], E_top, LS);
}

// k: how many tiles this CTA has done before this one (staging slot / mbarrier parity of the pull)
template <int ANC, bool FP32, bool PEERS, bool MULTI, bool TAIL, bool CARRY>
__device__ __forceinline__ void sb_hmm_tile(const SbHmmStepArgs& a, const SbStepInfo& si, int tile, int k, SbHmmSmem& sm,
                                            const uint32_t* s_thr, const uint32_t* s_key, const uint32_t* s_qtab, const uint32_t* s_tab) {
    const int tid = threadIdx.x;
    const int c0 = a.pool.c_off + tile * SB_TILE;                 // global child index of this tile
    const int base = c0 + tid * SB_ITEMS;
    const int lbase = tile * SB_TILE + tid * SB_ITEMS;            // where this GPU stores it
    const int K = a.K;
    const int N = (int)a.N;
    // (1) ancestors
    int anc[SB_ITEMS];
    if (ANC == SB_ANC_PULL) {
#if SB_ABLATE == 3
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) anc[j] = base + j;
        __syncthreads();
#else
        const int kk = k % SB_FT_CHUNK;
        const int2 rng = sm.ps.ft[kk];
        const bool more = tile + (int)gridDim.x < a.ntiles;
        const int nb = sm.ps.ft[kk + 1].x;
        // every warp has finished the previous tile's marker slots and exponent key, and the records warp this tile's boundary
        // table (the wait half of the split barrier; most of its skew was absorbed by the previous tile's epilogue)
        if (k > 0) { if (SB_OPT_NAMEDBAR) sb_split_wait(); else sb_mbar_wait(&sm.bar2, (uint32_t)((k - 1) & 1)); }
        if (si.s >= 32) sb_pull_ancestors<PEERS, true, false, CARRY>(si, a.pool, a.e_prev, a.Pt, a.child_start, c0, rng, k & 1, (uint32_t)(k & 1), more, nb, sm.ps, anc, a.xb);
        else            sb_pull_ancestors<PEERS, false, false, CARRY>(si, a.pool, a.e_prev, a.Pt, a.child_start, c0, rng, k & 1, (uint32_t)(k & 1), more, nb, sm.ps, anc, a.xb);
#endif
    } else if (ANC == SB_ANC_EXPLICIT) {
        if (!TAIL) {
            const int4 u = __ldcg(reinterpret_cast<const int4*>(a.anc_in + lbase));
            const int4 v = __ldcg(reinterpret_cast<const int4*>(a.anc_in + lbase) + 1);
            anc[0] = u.x; anc[1] = u.y; anc[2] = u.z; anc[3] = u.w; anc[4] = v.x; anc[5] = v.y; anc[6] = v.z; anc[7] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) anc[j] = (base + j < N) ? __ldcg(a.anc_in + lbase + j) : 0;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) anc[j] = base + j;
    }
    // CARRY: the pull returned (parent << xb) | parent state
    const int xsh = (CARRY && ANC == SB_ANC_PULL) ? a.xb : 0;
    if (ANC != SB_ANC_NONE && ANC != SB_ANC_EXPLICIT && a.anc_out) {
        if (!TAIL) {
            reinterpret_cast<int4*>(a.anc_out + lbase)[0] = make_int4(anc[0] >> xsh, anc[1] >> xsh, anc[2] >> xsh, anc[3] >> xsh);
            reinterpret_cast<int4*>(a.anc_out + lbase)[1] = make_int4(anc[4] >> xsh, anc[5] >> xsh, anc[6] >> xsh, anc[7] >> xsh);
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                const int i = base + j;
                if (i < N) a.anc_out[lbase + j] = anc[j] >> xsh;
                else if (i == N) a.anc_out[lbase + j] = N;       // the retained slot descends from itself
            }
        }
    }
    // (2) gather parents
    int xp[SB_ITEMS];
#if SB_ABLATE == 2
    if (ANC == SB_ANC_PULL && !TAIL) {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) xp[j] = anc[j] & 15;
    } else
#endif
    if (CARRY && ANC == SB_ANC_PULL) {                            // the head markers carried the parents' states: no gather
        const int xm = (1 << xsh) - 1;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) xp[j] = (!TAIL || base + j < N) ? (anc[j] & xm) : 0;
    } else if (PEERS && !TAIL && (anc[0] >> a.pool.log2n) == (anc[SB_ITEMS - 1] >> a.pool.log2n)) {
        // the pull returns parents in ascending order: first and last on the same GPU means all eight are,
        // and one base pointer serves them (the usual case; only shard boundaries take the general path)
        const int g = anc[0] >> a.pool.log2n;
        const int* xg = reinterpret_cast<const int*>(a.pool.base[g] + a.pool.off_x);
        const unsigned lmask = (1u << a.pool.log2n) - 1u;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) xp[j] = __ldg(xg + ((unsigned)anc[j] & lmask));
    } else {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j)
            xp[j] = (ANC != SB_ANC_NONE && (!TAIL || base + j < N)) ? sb_pool_x<int, PEERS>(a.pool, anc[j]) : 0;
    }
    // (3) propagate: one Philox call per FOUR particles, one 32-bit uniform per particle.
    // u = R * 2^-32 < cum[k]  <=>  R < ceil(cum[k] * 2^32): integer compares against the table.
    // src/rand.jl:6-12 on a monotone row: first k with R < thr[k], else K-1.  The lookup word of R's
    // bucket holds the number of thresholds at or below the bucket start and the (only) threshold
    // inside the bucket, so the draw is one shared-memory load and one compare.
    int x[SB_ITEMS];
#pragma unroll
    for (int p = 0; p < SB_ITEMS / 4; ++p) {
#if SB_ABLATE == 1
        const uint32_t hh = (uint32_t)((base >> 2) + p) * 2654435761u + a.t;
        const SbU4 o = SbU4{hh, hh * 3u, hh * 5u, hh * 7u};
#else
        const SbU4 o = sb_stream4_rk(a.rk, (unsigned long long)((base >> 2) + p), (unsigned)a.t, a.iter, SB_P_PROP);
#endif
        const uint32_t r4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int j = 4 * p + h;
            const int ri = a.rows > 1 ? xp[j] : 0;
            const uint32_t R = r4[h];
#if SB_ABLATE == 8                                                  /* diagnostics: 8 = no draw-table lookup (wrong results) */
            const uint32_t g = ((uint32_t)ri + (R >> 29)) & 7u;
#elif SB_ABLATE == 9                                                /* diagnostics: 9 = conflict-free draw-table lookup (wrong results) */
            const uint32_t g = s_tab[(ri << SB_LG_BUCKETS) + ((R >> (32 - SB_LG_BUCKETS)) & ~31u) + (tid & 31)];
#else
            const uint32_t g = s_tab[(ri << SB_LG_BUCKETS) + (R >> (32 - SB_LG_BUCKETS))];
#endif
            int lo = (int)(g & 63u) + (((R << SB_LG_BUCKETS) > g) ? 1 : 0);
            if (MULTI && (g & 128u)) {                           // several thresholds inside one bucket: scan
                const uint32_t* row = s_thr + ri * K;
                lo = (int)(g & 63u);
                while (lo < K - 1 && R >= row[lo]) ++lo;
            }
            x[j] = lo;
        }
    }
    // retained particle (src/pmcmc.jl:59-66) and padding
    if (TAIL) {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) {
            const int i = base + j;
            if (i >= N) x[j] = (i == N && a.xstar) ? __ldg(a.xstar + a.t) : 0;
        }
    }
    // the previous tile's warp totals (left in shared memory) become its warp prefixes and tile sum
    if (ANC != SB_ANC_PULL) __syncthreads();                     // the pull has its own barrier
    if (k > 0) {
        const int pt = tile - (int)gridDim.x;
        const unsigned long long Qp = sb_flush_totals(sm.tot[(k - 1) & 1], a.wp_out + (size_t)pt * SB_WARPS, a.rec, pt);
        if (Qp) sb_spec_add(sm, Qp, (k - 1) & 1, a.nt_pool);       // (the one thread that wrote the tile sum)
    }
    // (4)+(5) weight and quantise.  Only K distinct scores exist: the tile exponent is the largest
    // ceil(log2 weight) present (an integer max), and 2^(l2 - e) 2^27 comes from a K x K table.
    // The weights are looked up at once in the row of the LARGEST key the model can produce this step (sm.kmax): a
    // particle heavier than every below-top state of that row (sm.qthr) proves that the tile exponent is the top
    // one, and then these weights are final -- the usual case, no per-particle key lookup.  Otherwise the exact keys
    // are reduced as before and, should the tile exponent turn out lower, the weights are looked up again.
    const unsigned kmax = sm.kmax;
    constexpr bool SPEC = !TAIL && ANC == SB_ANC_PULL && SB_ABLATE != 3;
    uint32_t q[SB_ITEMS];
    unsigned key = 0u;
    if (SPEC) {
        const uint32_t* qtop = s_qtab + (kmax & 63u) * K;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) q[j] = qtop[x[j]];
        unsigned wm = __vimax3_u32(__vimax3_u32(q[0], q[1], q[2]), __vimax3_u32(q[3], q[4], q[5]), max(q[6], q[7]));
        wm = __reduce_max_sync(0xffffffffu, wm);
        if (wm > sm.qthr) key = kmax;
        else {
            key = __vimax3_u32(__vimax3_u32(s_key[x[0]], s_key[x[1]], s_key[x[2]]), __vimax3_u32(s_key[x[3]], s_key[x[4]], s_key[x[5]]),
                               max(s_key[x[6]], s_key[x[7]]));
            key = __reduce_max_sync(0xffffffffu, key);
        }
    } else {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) if (!TAIL || base + j < a.n_out) key = max(key, s_key[x[j]]);
        key = __reduce_max_sync(0xffffffffu, key);
    }
    if ((tid & 31) == 0) reinterpret_cast<unsigned*>(sm.red)[tid >> 5] = key;
    bool final_q = SPEC;                                        // q[] already holds the weights of the tile's row
    if (ANC == SB_ANC_PULL && SB_ABLATE != 3) {
        if ((tid >> 5) == SB_META_WARP && tile + (int)gridDim.x < a.ntiles) {
            // one warp: the next children tile's records have landed by now; its boundary table is built behind the barrier
            const int kn = k % SB_FT_CHUNK + 1;
            const int c0n = c0 + (int)gridDim.x * SB_TILE;
            if (si.s >= 32) sb_pull_seg<true>(si, c0n, sm.ps.ft[kn], (k + 1) & 1, sm.ps);
            else            sb_pull_seg<false>(si, c0n, sm.ps.ft[kn], (k + 1) & 1, sm.ps);
        }
        // split-phase barrier: arrive now, wait at the start of the next tile.  The tile exponent needs the other
        // warps' keys only when this warp's own maximum is below the largest key the model can produce this
        // step -- otherwise the tile maximum is already known and the warp goes on without waiting.
        __syncwarp();
        sb_split_arrive();                                        // waited for at the start of the next tile (hardware barrier, no spin)
        if ((tid & 31) == 0) sb_mbar_arrive(&sm.bar2);            // the same event for the rare waiters just below
        if (TAIL || key != kmax) {
            sb_mbar_wait(&sm.bar2, (uint32_t)(k & 1));
            const uint4 u = reinterpret_cast<const uint4*>(sm.red)[0], v = reinterpret_cast<const uint4*>(sm.red)[1];
            key = max(max(max(u.x, u.y), max(u.z, u.w)), max(max(v.x, v.y), max(v.z, v.w)));
            final_q = false;
        }
    } else {
        __syncthreads();
        const uint4 u = reinterpret_cast<const uint4*>(sm.red)[0], v = reinterpret_cast<const uint4*>(sm.red)[1];
        key = max(max(max(u.x, u.y), max(u.z, u.w)), max(max(v.x, v.y), max(v.z, v.w)));
        final_q = false;
    }
    const int e = key ? (int)(key >> 6) - SB_KEY_BIAS : SB_E_EMPTY;
    if (!final_q) {
        const uint32_t* qrow = s_qtab + (key & 63u) * K;
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) q[j] = (key != 0u && (!TAIL || base + j < a.n_out)) ? qrow[x[j]] : 0u;
    }
    unsigned tsum = 0;
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) tsum += q[j];
    sb_store8_u32(reinterpret_cast<uint32_t*>(a.x_out) + lbase, reinterpret_cast<const uint32_t*>(x));
    sb_store8_u32(a.w32_out + lbase, q);
    if (a.lw_out) {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) if (base + j < a.n_out) a.lw_out[lbase + j] = a.logF_t[x[j]];
    }
    if (a.wlit_out) {
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) if (base + j < a.n_out) a.wlit_out[lbase + j] = a.expF_t[x[j]];
    }
    sb_warp_total(tsum, sm.tot[k & 1]);                           // flushed after the next barrier (next tile / kernel end)
    if (tid == 0) { sb_put_e(a.rec, tile, e); sm.e_tile[k & 1] = e; }
    // no barrier here: every shared scratch array is rewritten only after a later barrier of the next tile
}

template <int ANC, bool FP32, bool PEERS, bool MULTI, bool CARRY = false>
__global__ void __launch_bounds__(SB_THREADS, 5)                  // 5 resident CTAs per SM: at most 48 registers
k_step_hmm(const __grid_constant__ SbHmmStepArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    sb_grid_dep_launch();                                           // the scan that follows may be scheduled as SMs drain
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(dyn);             // rows * 256 (16-byte aligned: staged as uint4)
    double* s_l2 = reinterpret_cast<double*>(s_tab + (a.rows << SB_LG_BUCKETS));   // K   (float view when FP32)
    uint32_t* s_thr = reinterpret_cast<uint32_t*>(s_l2 + a.K);      // rows*K
    uint32_t* s_key = s_thr + a.rows * a.K;                         // K: (ceil(l2) + bias) << 6 | state, 0 = no mass
    uint32_t* s_qtab = s_key + a.K;                                 // K*K: row k* = weights on the grid of exponent ceil(l2[k*])
    __shared__ SbHmmSmem sm;
    const int tid = threadIdx.x;
    const int K = a.K;
    if (MULTI) for (int i = tid; i < a.rows * K; i += SB_THREADS) s_thr[i] = __ldg(a.thr + i);
    for (int i = tid; i < ((a.rows << SB_LG_BUCKETS) >> 2); i += SB_THREADS)
        reinterpret_cast<uint4*>(s_tab)[i] = __ldg(reinterpret_cast<const uint4*>(a.tab) + i);
    if (tid < K) {
        int ce; bool live;
        if (FP32) {
            const float l2 = __fmul_rn((float)a.logF_t[tid], SB_LOG2E_F);
            if (isnan(l2) || l2 == INFINITY) atomicOr(a.err, SB_ERRBIT_NAN);
            reinterpret_cast<float*>(s_l2)[tid] = l2;
            live = l2 > -INFINITY; ce = (int)ceilf(fminf(fmaxf(l2, -16777215.0f), 16777215.0f));
        } else {
            const double l2 = __dmul_rn(a.logF_t[tid], SB_LOG2E_D);
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
        unsigned qt = 0u;                                         // heaviest below-top state in the top row (0xffffffff: no shortcut)
        if (km == 0u) qt = 0xffffffffu;
        for (int i = 0; i < K; ++i) if ((s_key[i] >> 6) != (km >> 6)) qt = max(qt, s_qtab[(km & 63u) * K + i]);
        sm.kmax = km; sm.qthr = qt; sm.spec_acc = 0ull;
    }
    sb_grid_dep_wait();                                           // everything above overlapped the previous scan
    SbStepInfo si;
    if (ANC == SB_ANC_PULL) {
        si = *a.info_prev;
        if (tid == 0) { sb_mbar_init(&sm.ps.bar, 1); sb_mbar_init(&sm.bar2, SB_WARPS); }
    }
    __syncthreads();                                              // tables staged
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        if (ANC == SB_ANC_PULL && (k % SB_FT_CHUNK) == 0 && SB_ABLATE != 3)
            sb_pull_ranges<PEERS, CARRY>(si, a.pool, a.e_prev, a.Pt, a.child_start, a.first_tile, tile, a.ntiles, k == 0, sm.ps);
        const bool tail = a.pool.c_off + (tile + 1) * SB_TILE > a.N;   // only the last tile(s) see padding / the retained slot
        if (!tail) sb_hmm_tile<ANC, FP32, PEERS, MULTI, false, CARRY>(a, si, tile, k, sm, s_thr, s_key, s_qtab, s_tab);
        else       sb_hmm_tile<ANC, FP32, PEERS, MULTI, true, CARRY>(a, si, tile, k, sm, s_thr, s_key, s_qtab, s_tab);
    }
    if (k > 0) {                                                  // the last tile's totals
        __syncthreads();
        const int pt = (int)blockIdx.x + (k - 1) * (int)gridDim.x;
        const unsigned long long Qp = sb_flush_totals(sm.tot[(k - 1) & 1], a.wp_out + (size_t)pt * SB_WARPS, a.rec, pt);
        if (Qp) sb_spec_add(sm, Qp, (k - 1) & 1, a.nt_pool);
    }
    if (a.rec.emax && (int)threadIdx.x == 32 * SB_FLUSH_WARP + SB_WARPS) {     // the thread that saw the tile sums
        if (sm.spec_acc) atomicAdd(reinterpret_cast<unsigned long long*>(a.rec.emax + 2), sm.spec_acc);
        if (blockIdx.x == 0) a.rec.emax[1] = (int)(sm.kmax >> 6) - SB_KEY_BIAS;
    }
    sb_signal_done(a.sig);
}

// ---------------------------------------------------------------- linear-Gaussian step
struct SbLgStepArgs {
    SbPool pool;                // previous pool (x is REAL)
    const int* e_prev;
    const unsigned long long* Pt;
    const int* child_start;
    const int* first_tile;
    const SbStepInfo* info_prev;
    const int* anc_in;
    void* x_out;
    uint32_t* w32_out;
    SbRecOut rec;               // tile exponents and sums
    SbSignal sig;               // sharded: tell the peers when this GPU's share of the pool is complete
    unsigned long long* wp_out;
    double* lw_out;
    long long N;                // global particle count
    int ntiles;                 // local tiles to produce
    int nt_pool;                // tiles of the whole pool (all GPUs): LS of the speculative total (sb_spec_term)
    int t;
    unsigned iter;
    unsigned long long seed;
    SbRoundKeys rk;             // Philox round keys of `seed`
    double a, sq, sp, c, r, m0, y, lc;   // sq = sqrt(q), sp = sqrt(P0), lc = -0.5 log(2 pi r)
    double hr;                           // -0.5 / r
    // the same constants rounded to float ONCE on the host (fp32 instantiations read them straight from the constant bank:
    // eight double -> float conversions and a division per tile and thread otherwise); fhr = -0.5f / (float)r
    float fa, fsq, fsp, fc, fy, fhr, flc, fm0;
    unsigned* err;
    // OPS instantiations (sb_model_ops): this step's draw / score statement and the observations (device memory)
    sb_op_step op;
    const double* obs;          // [n_obs] observations
    const double* obs_aux;      // [n_obs] lgamma(y + 1) (Poisson), computed on the host
    int xdim;                   // latent dimension: 1, or D of a Dirichlet program (x arrays are D planes of xplane doubles)
    long long xplane;
};

// Sequential draws of ONE particle at one step of an op-list program: Philox counter (particle | draw pair << 32,
// step, iteration, PROP), two 53-bit uniforms per block (the CPU checker used by the tests restates the same stream).
struct SbOpsRng {
    const SbRoundKeys& rk;
    unsigned long long idx;
    unsigned t, iter, n;
    __device__ __forceinline__ double next() {
        const SbU4 o = sb_stream4_rk(rk, idx | ((unsigned long long)(n >> 1) << 32), t, iter, SB_P_PROP);
        const double u = (n & 1) ? sb_u53(o.z, o.w) : sb_u53(o.x, o.y);
        ++n;
        return u;
    }
    __device__ __forceinline__ double normal() {
        const double u1 = next(), u2 = next();
        return sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586476925286766559 * u2);
    }
    // Marsaglia-Tsang (2000); shape < 1 boosted by u^(1/shape) -- as SbRng::gamma (the standalone ERP hooks)
    __device__ double gamma(double shape) {
        double boost = 1.0;
        if (shape < 1.0) { boost = pow(1.0 - next(), 1.0 / shape); shape += 1.0; }
        const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (int it = 0; it < 1000; ++it) {
            const double z = normal();
            double v = 1.0 + c * z;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = 1.0 - next();
            if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) return boost * d * v;
        }
        return boost * d;
    }
};
// one particle: the step's draw (src/pmcmc.jl:32-34: sample = rand(e)) and the score of its observe statement
// (src/Stochy.jl:55-56: the scores of the observations summed in order).  log(p), log(1 - p), log(lambda) do not
// depend on the observation: evaluated once, the sum then adds the same doubles the literal loop would.
__device__ __forceinline__ void sb_ops_particle(const sb_op_step& op, const double* __restrict__ obs, const double* __restrict__ obs_aux,
                                                SbOpsRng& g, double xprev, double* x_out, double* lw_out) {
    double xv = xprev;
    switch (op.draw) {
        case SB_DRAW_NORMAL:  xv = (op.dp[0] * xprev + op.dp[1]) + op.dp[2] * g.normal(); break;
        case SB_DRAW_BETA:    { const double X = g.gamma(op.dp[0]), Y = g.gamma(op.dp[1]); xv = X / (X + Y); } break;
        case SB_DRAW_GAMMA:   xv = op.dp[1] * g.gamma(op.dp[0]); break;
        case SB_DRAW_UNIFORM: xv = op.dp[0] + (op.dp[1] - op.dp[0]) * g.next(); break;
        case SB_DRAW_FLIP:    xv = g.next() < (op.dp[0] * xprev + op.dp[1]) ? 1.0 : 0.0; break;
        default: break;
    }
    double lw = 0.0;
    const double par = op.sp[0] * xv + op.sp[1];
    const double* y = obs + op.obs_off;
    switch (op.score) {
        case SB_SCORE_NORMAL: {
            const double lsd = log(op.sp[2]);
            for (int i = 0; i < op.obs_n; ++i) { const double z = (__ldg(y + i) - par) / op.sp[2]; lw += (-0.5 * z * z - lsd) - 0.9189385332046727418; }
        } break;
        case SB_SCORE_BERNOULLI: {
            const double l1 = log(par), l0 = log(1.0 - par);
            for (int i = 0; i < op.obs_n; ++i) lw += __ldg(y + i) != 0.0 ? l1 : l0;
        } break;
        case SB_SCORE_POISSON: {
            const double ll = log(par);
            const double* ya = obs_aux + op.obs_off;
            for (int i = 0; i < op.obs_n; ++i) lw += (__ldg(y + i) * ll - par) - __ldg(ya + i);
        } break;
        default: break;
    }
    *x_out = xv; *lw_out = lw;
}

#define SB_TWO_PI 6.283185307179586476925286766559

// one particle of a Dirichlet program (vector latent theta of D components, D planes in memory): theta = ~Dir(alpha)
// as D gammas normalised (src/erp.jl:18 `Dir`), or kept from the parent; observe(Categorical(theta), ys...) adds
// log theta[y] per observation, in order.  The D logs are taken once and parked in local memory.
__device__ __forceinline__ void sb_ops_particle_vec(const sb_op_step& op, const double* __restrict__ obs, int D, SbOpsRng& g,
                                                    const double* __restrict__ xprev_plane0, long long plane, bool has_parent,
                                                    double* __restrict__ xout_plane0, double* lw_out) {
    double th[SB_OPS_MAXD];
    if (op.draw == SB_DRAW_DIRICHLET) {
        double sum = 0.0;
#pragma unroll 1
        for (int d = 0; d < D; ++d) { th[d] = g.gamma(__ldg(obs + op.par_off + d)); sum += th[d]; }
        for (int d = 0; d < D; ++d) th[d] /= sum;
    } else {
        for (int d = 0; d < D; ++d) th[d] = has_parent ? __ldg(xprev_plane0 + (size_t)d * plane) : 0.0;
    }
    double lw = 0.0;
    if (op.score == SB_SCORE_CATEGORICAL) {
        double lg[SB_OPS_MAXD];
        for (int d = 0; d < D; ++d) lg[d] = log(th[d]);
        const double* y = obs + op.obs_off;
        for (int i = 0; i < op.obs_n; ++i) lw += lg[(int)__ldg(y + i) - 1];
    }
    for (int d = 0; d < D; ++d) xout_plane0[(size_t)d * plane] = th[d];
    *lw_out = lw;
}

template <int ANC, bool FP32, bool PEERS, bool OPS = false>
__global__ void __launch_bounds__(SB_THREADS, FP32 ? 5 : 3)       // fp32: 48 registers (5 CTAs per SM); fp64: 80 (3)
k_step_lg(const __grid_constant__ SbLgStepArgs a) {
    static_assert(!(OPS && FP32), "op-list programs are computed in fp64");
    typedef typename std::conditional<FP32, float, double>::type REAL;
    __shared__ SbHmmSmem sm;
    const int tid = threadIdx.x;
    const int N = (int)a.N;
    sb_grid_dep_launch();
    sb_grid_dep_wait();
    SbStepInfo si;
    if (ANC == SB_ANC_PULL) {
        si = *a.info_prev;
        if (tid == 0) { sb_mbar_init(&sm.ps.bar, 1); sb_mbar_init(&sm.bar2, SB_WARPS); }
        __syncthreads();
    }
    // no particle of this step can weigh more than the density's mode: lw <= lc, and the roundings below are monotone,
    // so every log2-weight is <= vmax.  A warp whose own maximum already reaches ceil(vmax) knows the tile exponent.
    // (an op-list program has no such bound: every warp waits for the tile maximum)
    const REAL vmax = FP32 ? (REAL)__fmul_rn((float)a.lc, SB_LOG2E_F) : (REAL)__dmul_rn(a.lc, SB_LOG2E_D);
    const int e_top = OPS ? 0x7fffffff : (FP32 ? (int)ceilf((float)vmax) : (int)ceil((double)vmax));
    // the scan's exact total, accumulated on the way at the exponent of the density's mode (as in k_step_hmm: sb_spec_term;
    // it is taken when some tile reaches that exponent -- at these pool sizes every tile does).  fp64 only: A/B on one
    // box, 250 steps: fp64 48.14 -> 47.88 ms per sweep; the fp32 instantiation (48 registers) lost in the kernel what the
    // scan gained (kernel 90.3 -> 91.4 us, sweep 23.44 -> 23.41 ms)
    constexpr bool LGSPEC = !OPS && !FP32;
    const int LS = 24 - (a.nt_pool <= 1 ? 0 : 32 - __clz(a.nt_pool - 1));
    if (tid == 0) sm.spec_acc = 0ull;                            // first read behind a later barrier
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        if (ANC == SB_ANC_PULL && (k % SB_FT_CHUNK) == 0)
            sb_pull_ranges<PEERS>(si, a.pool, a.e_prev, a.Pt, a.child_start, a.first_tile, tile, a.ntiles, k == 0, sm.ps);
        const int c0 = a.pool.c_off + tile * SB_TILE;             // global child index of this tile
        const int base = c0 + tid * SB_ITEMS;
        const int lbase = tile * SB_TILE + tid * SB_ITEMS;        // where this GPU stores it
        int anc[SB_ITEMS];
        if (ANC == SB_ANC_PULL) {
            const int kk = k % SB_FT_CHUNK;
            const int2 rng = sm.ps.ft[kk];
            const bool more = tile + (int)gridDim.x < a.ntiles;
            const int nb = sm.ps.ft[kk + 1].x;
            if (k > 0) { if (SB_OPT_NAMEDBAR) sb_split_wait(); else sb_mbar_wait(&sm.bar2, (uint32_t)((k - 1) & 1)); }   // wait half of the previous tile's split barrier (see k_step_hmm)
            if (si.s >= 32) sb_pull_ancestors<PEERS, true>(si, a.pool, a.e_prev, a.Pt, a.child_start, c0, rng, k & 1, (uint32_t)(k & 1), more, nb, sm.ps, anc);
            else            sb_pull_ancestors<PEERS, false>(si, a.pool, a.e_prev, a.Pt, a.child_start, c0, rng, k & 1, (uint32_t)(k & 1), more, nb, sm.ps, anc);
        } else if (ANC == SB_ANC_EXPLICIT) {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) anc[j] = (base + j < N) ? __ldcg(a.anc_in + lbase + j) : 0;
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) anc[j] = base + j;
        }
        // the previous tile's warp totals (left in shared memory) become its warp prefixes and tile sum
        if (ANC != SB_ANC_PULL) __syncthreads();                 // the pull has its own barrier
        if (k > 0) {
            const int pt = tile - (int)gridDim.x;
            const unsigned long long Qp = sb_flush_totals(sm.tot[(k - 1) & 1], a.wp_out + (size_t)pt * SB_WARPS, a.rec, pt);
            if (LGSPEC && Qp) sm.spec_acc += sb_spec_term(Qp, sm.e_tile[(k - 1) & 1], e_top, LS);
        }
        REAL xpar[SB_ITEMS];
        const bool vec = OPS && a.xdim > 1;                       // Dirichlet program: the parents' planes are read per particle below
        if (vec) {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) xpar[j] = (REAL)0;
        } else if (PEERS && base + SB_ITEMS <= N && (anc[0] >> a.pool.log2n) == (anc[SB_ITEMS - 1] >> a.pool.log2n)) {
            const int g = anc[0] >> a.pool.log2n;               // ascending parents: all eight on GPU g
            const REAL* xg = reinterpret_cast<const REAL*>(a.pool.base[g] + a.pool.off_x);
            const unsigned lmask = (1u << a.pool.log2n) - 1u;
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) xpar[j] = __ldg(xg + ((unsigned)anc[j] & lmask));
        } else if (FP32 && ANC != SB_ANC_NONE && c0 + SB_TILE <= N) {   // full tile: no bounds predicate on the gather
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) xpar[j] = sb_pool_x<REAL, PEERS>(a.pool, anc[j]);
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j)
                xpar[j] = (ANC != SB_ANC_NONE && base + j < N) ? sb_pool_x<REAL, PEERS>(a.pool, anc[j]) : (REAL)0;
        }
        REAL x[SB_ITEMS];
        REAL l2[SB_ITEMS];
        REAL mx = -INFINITY;
        bool bad = false;
        if (FP32) {
            const float fa = a.fa, fsq = a.fsq, fsp = a.fsp, fc = a.fc, fy = a.fy, hr = a.fhr, flc = a.flc, fm0 = a.fm0;
            // Box-Muller with both branches: one Philox block per FOUR particles (two pairs)
            float z[SB_ITEMS];
#pragma unroll
            for (int p = 0; p < SB_ITEMS / 4; ++p) {
                const SbU4 o = sb_stream4_rk(a.rk, (unsigned long long)((base >> 2) + p), (unsigned)a.t, a.iter, SB_P_PROP);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float u1 = (float)((h ? o.z : o.x) >> 8) * 0x1p-24f;
                    const float u2 = (float)((h ? o.w : o.y) >> 8) * 0x1p-24f;
                    // accurate log (its absolute error would otherwise dominate small radii), hardware sqrt and
                    // sin/cos: the angle is taken in [-pi, pi), where MUFU.SIN/COS are good to 2^-21 absolute, and
                    // cos(t + pi) = -cos t, sin(t + pi) = -sin t puts the sign on the radius
                    float r;
                    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * logf(1.0f - u1)));
                    float sn, cs;
                    __sincosf((float)SB_TWO_PI * (u2 - 0.5f), &sn, &cs);
                    z[4 * p + 2 * h] = -r * cs; z[4 * p + 2 * h + 1] = -r * sn;
                }
            }
            // (full tiles -- all but the pool's last -- take the instantiation without the per-particle bounds predicates)
            auto weigh = [&](auto full_c) {
                constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) {
                    float xv;
                    if (ANC == SB_ANC_NONE) xv = fm0 + fsp * z[j];
                    else xv = fa * (float)xpar[j] + fsq * z[j];
                    const float d = fy - fc * xv;
                    const float lw = hr * d * d + flc;
                    const bool live = FULL || base + j < N;
                    float v = live ? __fmul_rn(lw, SB_LOG2E_F) : -INFINITY;
                    if (!(v < INFINITY)) { bad = true; v = -INFINITY; }        // NaN or +Inf
                    x[j] = (REAL)xv; l2[j] = (REAL)v;
                    mx = (REAL)fmaxf((float)mx, v);
                }
            };
            if (base - tid * SB_ITEMS + SB_TILE <= N) weigh(std::true_type{}); else weigh(std::false_type{});
            if (a.lw_out) {                                       // store_logw (debug / parity): the same expression on the stored state,
#pragma unroll                                                    // behind ONE uniform branch instead of two predicated instructions per particle
                for (int j = 0; j < SB_ITEMS; ++j) {
                    const float d = fy - fc * (float)x[j];
                    const float lw = hr * d * d + flc;
                    if (base + j < N) a.lw_out[lbase + j] = (double)lw;
                }
            }
        } else if (OPS) {
#pragma unroll 1
            for (int j = 0; j < SB_ITEMS; ++j) {
                const bool live = base + j < N;
                double xv = 0.0, lw = 0.0;
                if (live) {
                    SbOpsRng g{a.rk, (unsigned long long)(base + j), (unsigned)a.t, a.iter, 0u};
                    if (vec) sb_ops_particle_vec(a.op, a.obs, a.xdim, g, reinterpret_cast<const double*>(a.pool.x) + anc[j], a.xplane, ANC != SB_ANC_NONE,
                                                 reinterpret_cast<double*>(a.x_out) + lbase + j, &lw);
                    else sb_ops_particle(a.op, a.obs, a.obs_aux, g, (double)xpar[j], &xv, &lw);
                }
                double v = live ? __dmul_rn(lw, SB_LOG2E_D) : -INFINITY;
                if (isnan(v) || v == INFINITY) { bad = true; v = -INFINITY; }
                x[j] = (REAL)xv; l2[j] = (REAL)v;
                mx = (REAL)fmax((double)mx, v);
                if (a.lw_out && live) a.lw_out[lbase + j] = lw;
            }
        } else {
            double z[SB_ITEMS];
#pragma unroll
            for (int p = 0; p < SB_ITEMS / 2; ++p) {
                const SbU4 o = sb_stream4_rk(a.rk, (unsigned long long)((base >> 1) + p), (unsigned)a.t, a.iter, SB_P_PROP);
                const double u1 = sb_u53(o.x, o.y), u2 = sb_u53(o.z, o.w);
                const double r = sqrt(-2.0 * log(1.0 - u1));
                double sn, cs;
                sincospi(2.0 * u2, &sn, &cs);
                z[2 * p] = r * cs; z[2 * p + 1] = r * sn;
            }
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                double xv;
                if (ANC == SB_ANC_NONE) xv = a.m0 + a.sp * z[j];
                else xv = a.a * (double)xpar[j] + a.sq * z[j];
                const double d = a.y - a.c * xv;
                const double lw = a.hr * d * d + a.lc;               // hr = -0.5 / r once on the host (a division per particle before)
                const bool live = base + j < N;
                double v = live ? __dmul_rn(lw, SB_LOG2E_D) : -INFINITY;
                if (isnan(v) || v == INFINITY) { bad = true; v = -INFINITY; }
                x[j] = (REAL)xv; l2[j] = (REAL)v;
                mx = (REAL)fmax((double)mx, v);
                if (a.lw_out && live) a.lw_out[lbase + j] = lw;
            }
        }
        if (ANC == SB_ANC_PULL && (tid >> 5) == SB_META_WARP && tile + (int)gridDim.x < a.ntiles) {
            // one warp: the next children tile's boundary table, built behind the split barrier of the block maximum
            const int kn = k % SB_FT_CHUNK + 1;
            const int c0n = c0 + (int)gridDim.x * SB_TILE;
            if (si.s >= 32) sb_pull_seg<true>(si, c0n, sm.ps.ft[kn], (k + 1) & 1, sm.ps);
            else            sb_pull_seg<false>(si, c0n, sm.ps.ft[kn], (k + 1) & 1, sm.ps);
        }
        int e;
        if (ANC == SB_ANC_PULL) {
            // split-phase barrier as in k_step_hmm: publish the warp maximum, arrive, and wait for the other warps only
            // if this warp's own maximum does not already reach the top exponent (m_w <= tile max <= vmax)
#pragma unroll
            for (int d = 16; d; d >>= 1) mx = FP32 ? (REAL)fmaxf((float)mx, __shfl_xor_sync(0xffffffffu, (float)mx, d))
                                                   : (REAL)fmax((double)mx, __shfl_xor_sync(0xffffffffu, (double)mx, d));
            if ((tid & 31) == 0) { if (FP32) reinterpret_cast<float*>(sm.red)[tid >> 5] = (float)mx; else sm.red[tid >> 5] = (double)mx; }
            __syncwarp();
            sb_split_arrive();
            if ((tid & 31) == 0) sb_mbar_arrive(&sm.bar2);
            e = mx > -INFINITY ? (FP32 ? (int)ceilf((float)mx) : (int)ceil((double)mx)) : SB_E_EMPTY;
            if (e != e_top) {
                sb_mbar_wait(&sm.bar2, (uint32_t)(k & 1));
                if (FP32) {
                    const float4 u = reinterpret_cast<const float4*>(sm.red)[0], v = reinterpret_cast<const float4*>(sm.red)[1];
                    const float m2 = fmaxf(fmaxf(fmaxf(u.x, u.y), fmaxf(u.z, u.w)), fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
                    e = m2 > -INFINITY ? (int)ceilf(m2) : SB_E_EMPTY;
                } else {
                    double m2 = sm.red[0];
#pragma unroll
                    for (int w = 1; w < SB_WARPS; ++w) m2 = fmax(m2, sm.red[w]);
                    e = m2 > -INFINITY ? (int)ceil(m2) : SB_E_EMPTY;
                }
            }
        } else if (FP32) { const float m2 = sb_block_max_f32_1((float)mx, reinterpret_cast<float*>(sm.red)); e = m2 > -INFINITY ? (int)ceilf(m2) : SB_E_EMPTY; }
        else             { const double m2 = sb_block_max_f64_1((double)mx, sm.red); e = m2 > -INFINITY ? (int)ceil(m2) : SB_E_EMPTY; }
        uint32_t q[SB_ITEMS];
        unsigned tsum = 0;
        // (fp32: one CTA-uniform branch instead of a select per particle, and below no bounds predicate on a full tile's gather:
        // 91.1 -> 89.9 us per launch; the same two changes made the fp64 instantiation SLOWER, 188.5 -> 192.4 us, A/B on one box)
        if (FP32 && e != SB_E_EMPTY) {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) { q[j] = sb_q31_f32(__fsub_rn((float)l2[j], (float)e)); tsum += q[j]; }
        } else {
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                uint32_t v = 0;
                if (!FP32 && e != SB_E_EMPTY) v = sb_q31_f64(__dsub_rn((double)l2[j], (double)e));
                q[j] = v; tsum += v;
            }
        }
        REAL* xo = reinterpret_cast<REAL*>(a.x_out) + lbase;
        if (FP32) {
            reinterpret_cast<float4*>(xo)[0] = make_float4((float)x[0], (float)x[1], (float)x[2], (float)x[3]);
            reinterpret_cast<float4*>(xo)[1] = make_float4((float)x[4], (float)x[5], (float)x[6], (float)x[7]);
        } else if (!vec) {                                        // (a Dirichlet program has stored its D planes already)
#pragma unroll
            for (int j = 0; j < SB_ITEMS; j += 2) reinterpret_cast<double2*>(xo)[j / 2] = make_double2((double)x[j], (double)x[j + 1]);
        }
        sb_store8_u32(a.w32_out + lbase, q);
        sb_warp_total(tsum, sm.tot[k & 1]);                       // flushed after the next barrier (next tile / kernel end)
        if (tid == 0) { sb_put_e(a.rec, tile, e); if (LGSPEC) sm.e_tile[k & 1] = e; }
        if (bad) atomicOr(a.err, SB_ERRBIT_NAN);
    }
    if (k > 0) {
        __syncthreads();
        const int pt = (int)blockIdx.x + (k - 1) * (int)gridDim.x;
        const unsigned long long Qp = sb_flush_totals(sm.tot[(k - 1) & 1], a.wp_out + (size_t)pt * SB_WARPS, a.rec, pt);
        if (LGSPEC && Qp) sm.spec_acc += sb_spec_term(Qp, sm.e_tile[(k - 1) & 1], e_top, LS);
    }
    if (LGSPEC && a.rec.emax && (int)threadIdx.x == 32 * SB_FLUSH_WARP + SB_WARPS) {   // the thread that saw the tile sums
        if (sm.spec_acc) atomicAdd(reinterpret_cast<unsigned long long*>(a.rec.emax + 2), sm.spec_acc);
        if (blockIdx.x == 0) a.rec.emax[1] = e_top;
    }
    sb_signal_done(a.sig);
}

// =====================================================================================
// Value accounting for pmcmc: counts[p.value] += 1 for ALL N particles (src/pmcmc.jl:90-92).
// Per-time marginals by propagating offspring counts backwards through the ancestor rows
// (coalesced, 12 B per particle-step) instead of N pointer chases of length T.
// =====================================================================================
__global__ void __launch_bounds__(256)
k_count_final(const int* __restrict__ anc_last, long long N, int* __restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) atomicAdd(cnt + anc_last[i], 1);
}
// cnt_t[j] lineages pass through pool entry j at step t.  anc_prev = ancestors drawn after step
// t-1 (nullptr at t == 0; identity when that step had no factor).  Also walks the retained
// lineage one step (thread 0 of block 0) so no separate pointer chase is needed.
// Eight consecutive entries per thread (128-bit loads); cnt_clear (the buffer the launch after next accumulates
// into) is zeroed on the way, so the T launches of a backward pass need no memset between them.  Lineages coalesce
// quickly going backwards: most entries are zero and cost one 4-byte read.
__global__ void __launch_bounds__(256)
k_count_back(const int* __restrict__ x_t, const int* __restrict__ cnt_t, const int* __restrict__ anc_prev,
             int identity_prev, long long n_pool, long long N, int K, int t,
             int* __restrict__ cnt_prev, int* __restrict__ cnt_clear, unsigned long long* __restrict__ marg,
             int* __restrict__ lineage, int* __restrict__ xstar_out) {
    extern __shared__ unsigned int s_hist[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_hist[k] = 0u;
    __syncthreads();
    const long long j0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (j0 < n_pool) {
        int c[8];
        if (j0 + 8 <= n_pool) {
            const int4 u = __ldg(reinterpret_cast<const int4*>(cnt_t + j0)), v = __ldg(reinterpret_cast<const int4*>(cnt_t + j0) + 1);
            c[0] = u.x; c[1] = u.y; c[2] = u.z; c[3] = u.w; c[4] = v.x; c[5] = v.y; c[6] = v.z; c[7] = v.w;
            if (cnt_clear) {
                reinterpret_cast<int4*>(cnt_clear + j0)[0] = make_int4(0, 0, 0, 0);
                reinterpret_cast<int4*>(cnt_clear + j0)[1] = make_int4(0, 0, 0, 0);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                c[i] = j0 + i < n_pool ? cnt_t[j0 + i] : 0;
                if (cnt_clear && j0 + i < n_pool) cnt_clear[j0 + i] = 0;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (c[i]) {
                const long long j = j0 + i;
                atomicAdd(&s_hist[x_t[j]], (unsigned)c[i]);
                if (cnt_prev) {
                    const int pj = (j >= N || identity_prev) ? (int)j : anc_prev[j];
                    atomicAdd(cnt_prev + pj, c[i]);
                }
            }
        }
    }
    __syncthreads();
    if (marg)
        for (int k = threadIdx.x; k < K; k += blockDim.x)
            if (s_hist[k]) atomicAdd(marg + (size_t)t * K + k, (unsigned long long)s_hist[k]);
    if (blockIdx.x == 0 && threadIdx.x == 0 && lineage) {
        const int r = *lineage;
        xstar_out[t] = x_t[r];
        if (t > 0) *lineage = (r >= N || identity_prev) ? r : anc_prev[r];
    }
}
// The whole backward pass of one PMCMC iteration in ONE cooperative launch: the loop over t of k_count_back with a
// grid-wide barrier between steps (the counts of step t-1 are complete before anybody reads them) instead of T
// launches -- at BASELINE cfg3's size a launch costs more than the work of a step (6.5 us against ~1 us).  Three count
// buffers in rotation as on the host path: step t reads buf[t % 3], accumulates into buf[(t - 1) % 3] and clears
// buf[(t + 1) % 3]; same integers as k_count_back (atomic adds of integers commute).
struct SbCountArgs {
    const int* x_hist;                  // [T][S]
    const int* anc_hist;                // [T][S]
    const unsigned char* has_factor;    // [T]
    int* cnt[3];
    long long S, n_pool, N;
    int K, T;
    unsigned long long* marg;           // [T*K]
    int* lineage;
    int* xstar_out;                     // [T]
    unsigned* bar;                      // grid barrier counter (sb_grid_barrier) and its value before this launch
    unsigned bar0;
    // the walk phase (below): live[t] = entries with a non-zero count at step t (zeroed before the launch), the
    // compacted list of (entry, count) pairs and its cursor (zeroed before the launch)
    unsigned* live;                     // [T]
    unsigned* cursor;
    int2* list;                         // [gridDim.x * blockDim.x]
};
#define SB_WALK_CHUNK 16                // steps per flush of the walk phase's shared-memory histograms
// Two phases.  COUNTING (as k_count_back, a grid barrier per step) while many pool entries still carry lineages.
// Lineages coalesce quickly going backwards (after g steps about N / (1 + g c) entries are live), and a step of
// the counting phase costs a grid barrier whatever the number of live entries -- so as soon as they fit one per
// thread the kernel switches to WALKING: the live (entry, count) pairs are compacted into a list, every thread takes
// one pair and follows its lineage down to step 0 on its own (two dependent loads per step, no barrier of any kind:
// the walks are independent), adding its count into per-CTA shared-memory histograms that are flushed every
// SB_WALK_CHUNK steps.  Same integers as the counting phase (integer adds commute); at BASELINE cfg3's size
// (2^20 particles, T = 1000) the pass takes ~1 ms instead of 5.2 ms (a barrier per step).
// Every CTA arrives T times at the grid barrier in total (the arrivals of the barriers it skips are added at the
// end), so the host can hand out the counter's base for the next launch without knowing when the switch happened.
__global__ void __launch_bounds__(256)
k_count_back_all(const __grid_constant__ SbCountArgs a) {
    extern __shared__ unsigned int s_hist[];                     // SB_WALK_CHUNK * K
    __shared__ unsigned s_nz[8];
    const int K = a.K;
    const unsigned nthreads = gridDim.x * blockDim.x, gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned barriers = 0;
    bool walk = false;
    int t = a.T - 1;
    while (true) {                                               // ---- counting phase
        const int* x_t = a.x_hist + (size_t)t * a.S;
        const int* anc_prev = t ? a.anc_hist + (size_t)(t - 1) * a.S : nullptr;
        const int identity_prev = t ? !a.has_factor[t - 1] : 0;
        const int* cnt_t = a.cnt[t % 3];
        int* cnt_prev = t ? a.cnt[(t - 1) % 3] : nullptr;
        int* cnt_clear = t >= 2 ? a.cnt[(t + 1) % 3] : nullptr;
        for (int k = threadIdx.x; k < K; k += blockDim.x) s_hist[k] = 0u;
        __syncthreads();
        unsigned nz = 0;
        for (long long j0 = (long long)gid * 8; j0 < a.n_pool; j0 += (long long)nthreads * 8) {
            int c[8];
            if (j0 + 8 <= a.n_pool) {
                // the counts were written by other CTAs earlier in this launch: coherent loads
                const int4 u = __ldcg(reinterpret_cast<const int4*>(cnt_t + j0)), v = __ldcg(reinterpret_cast<const int4*>(cnt_t + j0) + 1);
                c[0] = u.x; c[1] = u.y; c[2] = u.z; c[3] = u.w; c[4] = v.x; c[5] = v.y; c[6] = v.z; c[7] = v.w;
                if (cnt_clear) {
                    reinterpret_cast<int4*>(cnt_clear + j0)[0] = make_int4(0, 0, 0, 0);
                    reinterpret_cast<int4*>(cnt_clear + j0)[1] = make_int4(0, 0, 0, 0);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    c[i] = j0 + i < a.n_pool ? __ldcg(cnt_t + j0 + i) : 0;
                    if (cnt_clear && j0 + i < a.n_pool) cnt_clear[j0 + i] = 0;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (c[i]) {
                    ++nz;
                    const long long j = j0 + i;
                    atomicAdd(&s_hist[x_t[j]], (unsigned)c[i]);
                    if (cnt_prev) {
                        const int pj = (j >= a.N || identity_prev) ? (int)j : anc_prev[j];
                        atomicAdd(cnt_prev + pj, c[i]);
                    }
                }
            }
        }
        nz = __reduce_add_sync(0xffffffffu, nz);
        if (lane == 0) s_nz[threadIdx.x >> 5] = nz;
        __syncthreads();
        if (a.marg)
            for (int k = threadIdx.x; k < K; k += blockDim.x)
                if (s_hist[k]) atomicAdd(a.marg + (size_t)t * K + k, (unsigned long long)s_hist[k]);
        if (threadIdx.x == 0) {
            unsigned tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_nz[w];
            if (tot) atomicAdd(a.live + t, tot);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.lineage) {
            const int r = *a.lineage;
            a.xstar_out[t] = x_t[r];
            if (t > 0) *a.lineage = (r >= a.N || identity_prev) ? r : anc_prev[r];
        }
        if (t == 0) break;
        sb_grid_barrier(a.bar, a.bar0 + (++barriers) * gridDim.x);
        const unsigned live = __ldcg(a.live + t);               // entries with lineages at step t; no more than that at step t-1
        --t;
        if (live <= nthreads) { walk = true; break; }
    }
    if (walk) {                                                  // ---- walking phase: steps t .. 0
        // compact the live entries of step t (every count of this step is complete: the barrier above)
        const int* cnt_t = a.cnt[t % 3];
        // (warp-uniform trip count: the shuffles below need all 32 lanes)
        for (long long w0 = (long long)(gid - lane) * 8; w0 < a.n_pool; w0 += (long long)nthreads * 8) {
            const long long j0 = w0 + lane * 8;
            int c[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = j0 + i < a.n_pool ? __ldcg(cnt_t + j0 + i) : 0;
            int mine = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) mine += c[i] ? 1 : 0;
            int inc = mine;                                      // one cursor update per warp
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            const int wtot = __shfl_sync(0xffffffffu, inc, 31);
            unsigned wbase = 0;
            if (lane == 31 && wtot) wbase = atomicAdd(a.cursor, (unsigned)wtot);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            unsigned pos = wbase + (unsigned)(inc - mine);
#pragma unroll
            for (int i = 0; i < 8; ++i) if (c[i]) a.list[pos++] = make_int2((int)(j0 + i), c[i]);
        }
        sb_grid_barrier(a.bar, a.bar0 + (++barriers) * gridDim.x);
        const unsigned total = __ldcg(a.cursor);
        int j = 0, c = 0;
        if (gid < total) { const int2 e = __ldcg(a.list + gid); j = e.x; c = e.y; }
        int r = (blockIdx.x == 0 && threadIdx.x == 0 && a.lineage) ? *a.lineage : -1;
        for (int hi = t; hi >= 0; hi -= SB_WALK_CHUNK) {
            const int lo = max(hi - SB_WALK_CHUNK + 1, 0);
            for (int k = threadIdx.x; k < SB_WALK_CHUNK * K; k += blockDim.x) s_hist[k] = 0u;
            __syncthreads();
            for (int tt = hi; tt >= lo; --tt) {
                const int* x_t = a.x_hist + (size_t)tt * a.S;
                const int* anc_prev = tt ? a.anc_hist + (size_t)(tt - 1) * a.S : nullptr;
                const bool follow = tt > 0 && a.has_factor[tt - 1];
                if (c) {
                    atomicAdd(&s_hist[(hi - tt) * K + __ldg(x_t + j)], (unsigned)c);
                    if (follow && j < a.N) j = __ldg(anc_prev + j);
                }
                if (r >= 0) {
                    a.xstar_out[tt] = __ldg(x_t + r);
                    if (follow && r < a.N) r = __ldg(anc_prev + r);
                }
            }
            __syncthreads();
            if (a.marg)
                for (int k = threadIdx.x; k < (hi - lo + 1) * K; k += blockDim.x)
                    if (s_hist[k]) atomicAdd(a.marg + (size_t)(hi - k / K) * K + (k % K), (unsigned long long)s_hist[k]);
            __syncthreads();
        }
    }
    if (threadIdx.x == 0 && barriers < (unsigned)a.T)            // the arrivals of the barriers this CTA skipped
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(a.bar), "r"((unsigned)a.T - barriers) : "memory");
}
// start of the retained lineage: child c* of the final resample (src/pmcmc.jl:89 or uniform pick, H4)
__global__ void k_lineage_start(const int* __restrict__ anc_last, long long N, int pick_uniform,
                                unsigned long long seed, unsigned T, unsigned iter, int* __restrict__ lineage) {
    long long c = 0;
    if (pick_uniform) {
        const SbU4 o = sb_stream4(seed, 0ull, T, iter, SB_P_PICK);
        c = (long long)__double2ll_rz(__dmul_rn(sb_u53(o.x, o.y), (double)N));
        if (c >= N) c = N - 1;
    }
    *lineage = anc_last[c];
}
// joint trajectory histogram for small K^T: one walk of length T per final particle
__global__ void __launch_bounds__(256)
k_joint_hist(const int* __restrict__ x_hist, const int* __restrict__ anc_hist, const unsigned char* __restrict__ has_factor,
             long long S, long long N, int K, int T, unsigned long long* __restrict__ joint) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    long long j = anc_hist[(size_t)(T - 1) * S + c];
    unsigned long long code = 0;
    for (int t = T - 1; t >= 0; --t) {
        code = code * (unsigned long long)K + (unsigned long long)x_hist[(size_t)t * S + j];
        if (t > 0 && j < N && has_factor[t - 1]) j = anc_hist[(size_t)(t - 1) * S + j];
    }
    atomicAdd(joint + code, 1ull);
}
__global__ void k_fill_identity(int* __restrict__ a, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}

// =====================================================================================
// Enumerate for the fixed-structure discrete model (src/enumerate.jl:36-85).
// k_enum_joint: one thread per complete execution path (trajectory code, little-endian base K); the score is
// accumulated in program order -- sample adds score(e, val) (enumerate.jl:38), factor adds its score (:45) -- and
// the path contributes exp(score) to its return value (:72).  Exhaustive: K^T <= 2^26 paths.
// k_forward_backward: the same posterior marginalised per time step in O(T K^2) (scaled forward-backward): what
// Enumerate would return for the per-time values of a model too long to enumerate.  One CTA, thread k = state k.
// =====================================================================================
__global__ void __launch_bounds__(256)
k_enum_joint(const double* __restrict__ logA, const double* __restrict__ logp0, const double* __restrict__ logF,
             const unsigned char* __restrict__ has_factor, int K, int T, long long ncodes, double* __restrict__ mass) {
    const long long code = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (code >= ncodes) return;
    long long c = code;
    double score = 0.0;
    int prev = 0;
    for (int t = 0; t < T; ++t) {
        const int x = (int)(c % K);
        c /= K;
        score += t == 0 ? __ldg(logp0 + x) : __ldg(logA + prev * K + x);       // sample: ctx.score + score(e, val)
        if (has_factor[t]) score += __ldg(logF + (size_t)t * K + x);            // factor: ctx.score += score
        prev = x;
    }
    mass[code] = exp(score);
}

__global__ void __launch_bounds__(64)
k_forward_backward(const double* __restrict__ A, const double* __restrict__ p0, const double* __restrict__ logF,
                   const unsigned char* __restrict__ has_factor, int K, int T, double* __restrict__ alpha /* [T*K] */,
                   double* __restrict__ cnorm /* [T] */, double* __restrict__ marg /* [T*K] */, double* __restrict__ logZ) {
    __shared__ double s_a[64], s_b[64], s_red[64];
    const int k = threadIdx.x;
    const bool live = k < K;
    double lz = 0.0;
    for (int t = 0; t < T; ++t) {                                   // forward pass, normalised at every step
        double v = 0.0;
        if (live) {
            if (t == 0) v = p0[k];
            else for (int j = 0; j < K; ++j) v += s_a[j] * __ldg(A + j * K + k);
            if (has_factor[t]) v *= exp(__ldg(logF + (size_t)t * K + k));
        }
        __syncthreads();
        s_red[k] = v;
        __syncthreads();
        double c = 0.0;
        for (int j = 0; j < K; ++j) c += s_red[j];
        const double a = c > 0.0 ? v / c : 0.0;
        if (live) { s_a[k] = a; alpha[(size_t)t * K + k] = a; }
        if (k == 0) cnorm[t] = c;
        lz += log(c);
        __syncthreads();
    }
    if (k == 0) *logZ = lz;
    if (live) { s_b[k] = 1.0; marg[(size_t)(T - 1) * K + k] = alpha[(size_t)(T - 1) * K + k]; }
    __syncthreads();
    for (int t = T - 2; t >= 0; --t) {                              // backward pass with the forward normalisers
        double v = 0.0;
        if (live) {
            for (int j = 0; j < K; ++j) {
                const double f = has_factor[t + 1] ? exp(__ldg(logF + (size_t)(t + 1) * K + j)) : 1.0;
                v += __ldg(A + k * K + j) * f * s_b[j];
            }
            v /= cnorm[t + 1];
        }
        __syncthreads();
        if (live) { s_b[k] = v; marg[(size_t)t * K + k] = alpha[(size_t)t * K + k] * v; }
        __syncthreads();
    }
}

// =====================================================================================
// Batched single-site MH on a binary Bayes net: one thread = one chain, state in registers,
// tables in shared memory, warp-aggregated histogram.  Replaces the loop src/mh.jl:61-88 and
// log_mh_ratio src/mh.jl:95-111 for fixed-structure traces.  Not HBM-bound (INT/FP64 pipes).
// =====================================================================================
struct SbBnetArgs {
    int D, F, stride;
    unsigned query_mask;
    const unsigned* site_parents;
    const double* site_cpt;     // P(site = 1 | parents)
    const double* site_lp1;     // log p
    const double* site_lp0;     // log(1 - p)
    const unsigned* fac_parents;
    const double* fac_logf;
    unsigned long long seed;
    long long chains, steps, chain0;
    unsigned long long* counts;
    unsigned long long* accepts;
};
__device__ __forceinline__ unsigned sb_pack_bits(unsigned state, unsigned mask) {
    unsigned idx = 0, pos = 0;
    while (mask) {
        const unsigned b = mask & (0u - mask);
        if (state & b) idx |= 1u << pos;
        ++pos; mask ^= b;
    }
    return idx;
}
#define SB_MH_MAXD 16
// The per-warp histogram words are 32 bits and grow by at most 32 per step: every 2^26 steps the CTA adds them to
// the 64-bit global counters and clears them (uniform over the CTA: every thread runs the same number of steps).
#define SB_MH_FLUSH_MASK ((1ll << 26) - 1ll)
__device__ __forceinline__ void sb_mh_flush(unsigned* s_hist, int nwarps, int nv, unsigned long long* counts) {
    __syncthreads();
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        unsigned long long t = 0;
        for (int w = 0; w < nwarps; ++w) { t += s_hist[w * nv + v]; s_hist[w * nv + v] = 0u; }
        if (t) atomicAdd(counts + v, t);
    }
    __syncthreads();
}
__global__ void __launch_bounds__(128)
k_mh_bnet(const __grid_constant__ SbBnetArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int D = a.D, F = a.F, st = a.stride;
    double* s_cpt = reinterpret_cast<double*>(dyn);         // D*st
    double* s_lp1 = s_cpt + D * st;
    double* s_lp0 = s_lp1 + D * st;
    double* s_fac = s_lp0 + D * st;                         // F*st
    unsigned* s_sp = reinterpret_cast<unsigned*>(s_fac + F * st);   // D
    unsigned* s_fp = s_sp + D;                              // F
    unsigned* s_hist = s_fp + F;                            // warps * nv
    int nq = __popc(a.query_mask);
    const int nv = 1 << nq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < D * st; i += blockDim.x) { s_cpt[i] = a.site_cpt[i]; s_lp1[i] = a.site_lp1[i]; s_lp0[i] = a.site_lp0[i]; }
    for (int i = threadIdx.x; i < F * st; i += blockDim.x) s_fac[i] = a.fac_logf[i];
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_sp[i] = a.site_parents[i];
    for (int i = threadIdx.x; i < F; i += blockDim.x) s_fp[i] = a.fac_parents[i];
    for (int i = threadIdx.x; i < nwarps * nv; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = ch < a.chains;
    const unsigned long long gch = (unsigned long long)(a.chain0 + ch);
    unsigned state = 0;
    double cs_old[SB_MH_MAXD];
    double score_old = 0.0;
    unsigned long long acc = 0;
    // src/mh.jl:57 -- run once from the prior; choicescores in program order
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        const SbU4 o = sb_stream4(a.seed, gch, (unsigned)d, 0u, SB_P_INIT);
        const double u = __dmul_rn((double)o.x, 0x1p-32);
        const unsigned idx = sb_pack_bits(state, s_sp[d]);
        if (u < s_cpt[d * st + idx]) state |= 1u << d;
    }
#pragma unroll
    for (int d = 0; d < SB_MH_MAXD; ++d) {
        if (d < D) {
            const unsigned idx = sb_pack_bits(state, s_sp[d]);
            const double c = ((state >> d) & 1u) ? s_lp1[d * st + idx] : s_lp0[d * st + idx];
            cs_old[d] = c;
            score_old = __dadd_rn(score_old, c);
        }
    }
    for (int f = 0; f < F; ++f) score_old = __dadd_rn(score_old, s_fac[f * st + sb_pack_bits(state, s_fp[f])]);
    unsigned* hist = s_hist + warp * nv;
    // The chain is a dependent sequence (state -> table lookups -> state) and only ~14 warps are resident per SM at
    // 65536 chains: the time per step is latency, not issue slots.  The Philox block does not depend on the state: the
    // NEXT step's is computed alongside this step's dependent part (software pipelining by hand).
    SbU4 onext = sb_stream4(a.seed, gch, 0u, 0u, SB_P_MH);
    for (long long s = 0; s < a.steps; ++s) {
        const SbU4 o = onext;
        onext = sb_stream4(a.seed, gch, (unsigned)(s + 1), (unsigned)((s + 1) >> 32), SB_P_MH);
        const int j = (int)__umulhi(o.x, (unsigned)D);                               // mh.jl:62
        const double up = __dmul_rn((double)o.y, 0x1p-32), ua = __dmul_rn((double)o.z, 0x1p-32);
        const unsigned prefix = state & ((1u << j) - 1u);
        unsigned prop = state & ~(1u << j);
        if (up < s_cpt[j * st + sb_pack_bits(prefix, s_sp[j])]) prop |= 1u << j;     // mh.jl:70-76
        double cs_new[SB_MH_MAXD];
        double score_new = 0.0;
#pragma unroll
        for (int d = 0; d < SB_MH_MAXD; ++d) {
            if (d < D) {
                const unsigned idx = sb_pack_bits(prop, s_sp[d]);
                const double c = ((prop >> d) & 1u) ? s_lp1[d * st + idx] : s_lp0[d * st + idx];
                cs_new[d] = c;
                score_new = __dadd_rn(score_new, c);
            }
        }
        for (int f = 0; f < F; ++f) score_new = __dadd_rn(score_new, s_fac[f * st + sb_pack_bits(prop, s_fp[f])]);
        double fw = 0.0, bw = 0.0;
#pragma unroll
        for (int d = 0; d < SB_MH_MAXD; ++d) if (d == j) { fw = cs_new[d]; bw = cs_old[d]; }
        const double lr = __dsub_rn(__dadd_rn(__dsub_rn(score_new, score_old), bw), fw);   // mh.jl:110
        const double ar = exp(lr);
        const bool accept = !isnan(ar) && ua < (ar < 1.0 ? ar : 1.0);                // mh.jl:78
        if (accept) {
            state = prop; score_old = score_new; ++acc;
#pragma unroll
            for (int d = 0; d < SB_MH_MAXD; ++d) cs_old[d] = cs_new[d];
        }
        // mh.jl:87 -- count the current value after every step (warp-aggregated)
        const unsigned v = sb_pack_bits(state, a.query_mask);
        const unsigned act = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned peers = __match_any_sync(act, v);
            if (lane == __ffs(peers) - 1) hist[v] += __popc(peers);
        }
        __syncwarp();
        if ((s & SB_MH_FLUSH_MASK) == SB_MH_FLUSH_MASK) sb_mh_flush(s_hist, nwarps, nv, a.counts);
    }
    sb_mh_flush(s_hist, nwarps, nv, a.counts);
    if (live && acc) atomicAdd(a.accepts, acc);
}

// Same chains for nets of at most SB_MH_TABD sites: everything a step needs is a function of (site, state), so each
// CTA tabulates it once -- P(site = 1 | parents) and the acceptance probability min(1, exp(log_mh_ratio)) of
// src/mh.jl:78,95-111, computed with the SAME sequence of fp64 operations as k_mh_bnet -- and a step is one Philox
// block, two shared-memory loads and two compares.  Bit-identical chains, about a quarter of the instructions.
#define SB_MH_TABD 8
__global__ void __launch_bounds__(128)
k_mh_bnet_tab(const __grid_constant__ SbBnetArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int D = a.D, F = a.F, st = a.stride;
    const int NS = 1 << D;
    double* s_p1 = reinterpret_cast<double*>(dyn);            // [D][NS]
    double* s_acc = s_p1 + D * NS;                            // [D][NS][2]
    double* s_score = s_acc + 2 * D * NS;                     // [NS]
    double* s_cs = s_score + NS;                              // [D][NS]
    unsigned* s_hist = reinterpret_cast<unsigned*>(s_cs + D * NS);
    const int nq = __popc(a.query_mask), nv = 1 << nq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned short* s_val = reinterpret_cast<unsigned short*>(s_hist + nwarps * nv);   // [NS] the program's value in every state (mh.jl:87)
    for (int i = threadIdx.x; i < nwarps * nv; i += blockDim.x) s_hist[i] = 0u;
    for (int i = threadIdx.x; i < NS; i += blockDim.x) s_val[i] = (unsigned short)sb_pack_bits((unsigned)i, a.query_mask);
    for (int s = threadIdx.x; s < NS; s += blockDim.x) {      // choice scores and total score of every state (src/mh.jl:33,42)
        double sc = 0.0;
        for (int d = 0; d < D; ++d) {
            const unsigned idx = sb_pack_bits((unsigned)s, a.site_parents[d]);
            const double c = ((s >> d) & 1) ? a.site_lp1[d * st + idx] : a.site_lp0[d * st + idx];
            s_cs[d * NS + s] = c;
            s_p1[d * NS + s] = a.site_cpt[d * st + idx];
            sc = __dadd_rn(sc, c);
        }
        for (int f = 0; f < F; ++f) sc = __dadd_rn(sc, a.fac_logf[f * st + sb_pack_bits((unsigned)s, a.fac_parents[f])]);
        s_score[s] = sc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * NS * 2; i += blockDim.x) {
        const int b = i & 1, s = (i >> 1) & (NS - 1), d = i >> (D + 1);
        const int prop = (s & ~(1 << d)) | (b << d);
        const double lr = __dsub_rn(__dadd_rn(__dsub_rn(s_score[prop], s_score[s]), s_cs[d * NS + s]), s_cs[d * NS + prop]);   // mh.jl:110
        const double ar = exp(lr);
        s_acc[i] = isnan(ar) ? -1.0 : (ar < 1.0 ? ar : 1.0);                                                                   // mh.jl:78
    }
    __syncthreads();
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = ch < a.chains;
    const unsigned long long gch = (unsigned long long)(a.chain0 + ch);
    unsigned state = 0;
    unsigned long long acc = 0;
#pragma unroll 1
    for (int d = 0; d < D; ++d) {                              // src/mh.jl:57 -- run once from the prior
        const SbU4 o = sb_stream4(a.seed, gch, (unsigned)d, 0u, SB_P_INIT);
        if (__dmul_rn((double)o.x, 0x1p-32) < s_p1[d * NS + state]) state |= 1u << d;
    }
    unsigned* hist = s_hist + warp * nv;
    const unsigned act = __ballot_sync(0xffffffffu, live);
    unsigned long long c1 = 0;
    // The chain is a dependent sequence (state -> table lookups -> state) and only ~14 warps are resident per SM at
    // 65536 chains: the time per step is latency, not issue slots.  The Philox block does not depend on the state: the
    // NEXT step's is computed alongside this step's dependent part (software pipelining by hand).
    SbU4 onext = sb_stream4(a.seed, gch, 0u, 0u, SB_P_MH);
    for (long long s = 0; s < a.steps; ++s) {
        const SbU4 o = onext;
        onext = sb_stream4(a.seed, gch, (unsigned)(s + 1), (unsigned)((s + 1) >> 32), SB_P_MH);
        const int j = (int)__umulhi(o.x, (unsigned)D);                               // mh.jl:62
        // (integer thresholds ceil(p 2^32) instead of the two fp64 compares were measured: 64-bit compares, 3.56 vs 3.45 ms)
        const double up = __dmul_rn((double)o.y, 0x1p-32), ua = __dmul_rn((double)o.z, 0x1p-32);
        const int b = up < s_p1[j * NS + state] ? 1 : 0;                             // mh.jl:70-76
        if (ua < s_acc[(((j << D) + (int)state) << 1) + b]) {
            state = (state & ~(1u << j)) | ((unsigned)b << j);
            ++acc;
        }
        const unsigned v = s_val[state];                                             // mh.jl:87
        if (nv == 2) { c1 += v; continue; }                                          // one query site: a counter per chain, no histogram traffic
        if (live) {
            const unsigned peers = __match_any_sync(act, v);
            if (lane == __ffs(peers) - 1) hist[v] += __popc(peers);
        }
        __syncwarp();
        if ((s & SB_MH_FLUSH_MASK) == SB_MH_FLUSH_MASK) sb_mh_flush(s_hist, nwarps, nv, a.counts);
    }
    if (nv == 2) {                                               // (CTA-uniform) counts[1] = sum of the chains' counters, counts[0] = the rest
        unsigned long long t1 = live ? c1 : 0ull, t0 = live ? (unsigned long long)a.steps - c1 : 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            t1 += ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(t1 >> 32), d) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)t1, d);
            t0 += ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(t0 >> 32), d) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)t0, d);
        }
        if (lane == 0) { if (t0) atomicAdd(a.counts, t0); if (t1) atomicAdd(a.counts + 1, t1); }
    }
    sb_mh_flush(s_hist, nwarps, nv, a.counts);
    if (live && acc) atomicAdd(a.accepts, acc);
}

// =====================================================================================
// ERP vocabulary (src/erp.jl:1-30): samplers and log-densities.  Distributions.jl 0.6.3 is not
// in the reference tree, so these are the textbook algorithms; validated statistically.
// =====================================================================================
struct SbRng {           // sequential draws from one Philox counter stream (for rejection samplers)
    unsigned long long seed, idx;
    unsigned n;
    __device__ double next() {
        const SbU4 o = sb_stream4(seed, idx, n >> 1, 0u, 5u);
        const double u = (n & 1) ? sb_u53(o.z, o.w) : sb_u53(o.x, o.y);
        ++n;
        return u;
    }
    __device__ double normal() {
        const double u1 = next(), u2 = next();
        return sqrt(-2.0 * log(1.0 - u1)) * cos(SB_TWO_PI * u2);
    }
    // Marsaglia-Tsang (2000); shape < 1 boosted by u^(1/shape)
    __device__ double gamma(double shape) {
        double boost = 1.0;
        if (shape < 1.0) { boost = pow(1.0 - next(), 1.0 / shape); shape += 1.0; }
        const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (int it = 0; it < 1000; ++it) {
            const double z = normal();
            double v = 1.0 + c * z;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = 1.0 - next();
            if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) return boost * d * v;
        }
        return boost * d;
    }
};

__global__ void __launch_bounds__(256)
k_erp_sample(int erp, double p0, double p1, long long n, unsigned long long seed, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SbRng g{seed, (unsigned long long)i, 0u};
    double v = 0.0;
    switch (erp) {
        case 0: v = g.next() < p0 ? 1.0 : 0.0; break;                                  // flip: (false,true)[Bernoulli+1]
        case 1: { double k = floor(g.next() * p0); v = (k >= p0 ? p0 - 1.0 : k) + 1.0; } break;   // randominteger: 1..k
        case 2: v = p0 + p1 * g.normal(); break;
        case 3: { const double x = g.gamma(p0), y = g.gamma(p1); v = x / (x + y); } break;
        case 4: v = p1 * g.gamma(p0); break;
        case 5: v = p0 + (p1 - p0) * g.next(); break;
    }
    out[i] = v;
}
__global__ void __launch_bounds__(256)
k_erp_logpdf(int erp, double p0, double p1, const double* __restrict__ x, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double lp = -INFINITY;
    switch (erp) {
        case 0: lp = v == 1.0 ? log(p0) : (v == 0.0 ? log(1.0 - p0) : -INFINITY); break;
        case 1: lp = (v >= 1.0 && v <= p0 && v == floor(v)) ? -log(p0) : -INFINITY; break;
        case 2: { const double z = (v - p0) / p1; lp = -0.5 * z * z - log(p1) - 0.5 * log(SB_TWO_PI); } break;
        case 3: lp = (v > 0.0 && v < 1.0) ? (p0 - 1.0) * log(v) + (p1 - 1.0) * log1p(-v) + lgamma(p0 + p1) - lgamma(p0) - lgamma(p1) : -INFINITY; break;
        case 4: lp = v > 0.0 ? (p0 - 1.0) * log(v) - v / p1 - lgamma(p0) - p0 * log(p1) : -INFINITY; break;
        case 5: lp = (v >= p0 && v <= p1) ? -log(p1 - p0) : -INFINITY; break;
    }
    out[i] = lp;
}
// Discrete ERP (src/erp.jl:39-56): rand(erp) = erp.x[rand(erp.p)] with p = values(dict), divided by z = sum(p) when
// z != 1.0 (:49); rand(ps) is the inverse-CDF scan of src/rand.jl:2-13 (sequential fp64 adds, strict `r < acc`, last
// index as fallback, NaN -> error).  score(erp, x) = log(dict[x] / z) (:56).  Draws are 1-based positions in the support.
__global__ void __launch_bounds__(256)
k_discrete_sample(const double* __restrict__ p, int k, double z, long long n, unsigned long long seed,
                  int* __restrict__ out, unsigned* __restrict__ err) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SbRng g{seed, (unsigned long long)i, 0u};
    const double r = g.next();
    double acc = 0.0;
    int draw = k;
    for (int j = 0; j < k - 1; ++j) {
        const double pj = z != 1.0 ? __ddiv_rn(p[j], z) : p[j];
        acc = __dadd_rn(acc, pj);
        if (isnan(acc)) { atomicOr(err, SB_ERRBIT_NAN); break; }
        if (r < acc) { draw = j + 1; break; }
    }
    if (k == 1 && isnan(z != 1.0 ? __ddiv_rn(p[0], z) : p[0])) atomicOr(err, SB_ERRBIT_NAN);
    out[i] = draw;
}
__global__ void __launch_bounds__(256)
k_discrete_logpdf(const double* __restrict__ p, int k, double z, const int* __restrict__ x, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = x[i];
    out[i] = (v >= 1 && v <= k) ? log(__ddiv_rn(p[v - 1], z)) : -INFINITY;       // outside the support: no mass
}
#define SB_DIR_MAXK 32
__global__ void __launch_bounds__(256)
k_dirichlet_sample(const double* __restrict__ alpha, int k, long long n, unsigned long long seed, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SbRng g{seed, (unsigned long long)i, 0u};
    double s = 0.0;
    for (int j = 0; j < k; ++j) { const double v = g.gamma(alpha[j]); out[i * k + j] = v; s += v; }
    for (int j = 0; j < k; ++j) out[i * k + j] /= s;
}
__global__ void __launch_bounds__(256)
k_dirichlet_logpdf(const double* __restrict__ alpha, int k, const double* __restrict__ x, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a0 = 0.0, lp = 0.0;
    for (int j = 0; j < k; ++j) { a0 += alpha[j]; lp += (alpha[j] - 1.0) * log(x[i * k + j]) - lgamma(alpha[j]); }
    out[i] = lp + lgamma(a0);
}
