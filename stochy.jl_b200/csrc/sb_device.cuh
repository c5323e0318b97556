// sb_device.cuh -- device-side building blocks of libstochy_b200 (sm_100a).
//
// Everything here is HBM-bound integer / byte work plus a little fp64; there is no dense
// contraction on this path, so no tensor cores (tcgen05) are used.  The rules that matter are
// coalesced 128-bit accesses, grids sized in tiles of SB_TILE particles, shared-memory staging
// of the per-tile scan, and keeping the bytes per particle-step at the algorithmic minimum.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define SB_TILE     2048          // particles per tile == per CTA
#define SB_THREADS  256
#define SB_ITEMS    8             // consecutive particles per thread (blocked arrangement)
#define SB_WARPS    (SB_THREADS / 32)
#define SB_QBITS    27            // fixed-point weight bits: tile prefix <= 2^38, times a 26-bit multiplier < 2^64
#define SB_E_EMPTY  (-0x40000000)

#define SB_ERRBIT_NAN   1u
#define SB_ERRBIT_ZERO  2u
#define SB_ERRBIT_PEER  4u          // a peer GPU never signalled (multi-GPU flag barrier timed out)

#define SB_P_PROP     0u
#define SB_P_RESAMPLE 1u
#define SB_P_INIT     2u
#define SB_P_PICK     3u
#define SB_P_MH       4u

#define SB_LN2     0x1.62e42fefa39efp-1
#define SB_LOG2E_D 0x1.71547652b82fep+0
#define SB_LOG2E_F 0x1.715476p+0f
#define SB_SQRT2_D 0x1.6a09e667f3bcdp+0
#define SB_SQRT2_F 0x1.6a09e6p+0f

// Written by k_tilescan for the pool that is about to be resampled; read by every consumer.
struct SbStepInfo {
    unsigned long long total;   // exact integer sum of the grid weights (<= 2^62): log-evidence, multinomial
    unsigned long long Tt;      // total of the rescaled CDF (~ m * 2^s): systematic
    unsigned long long U;       // systematic offset: child k sits at U + k * 2^s
    long long n;                // pool size
    long long m;                // children to draw
    unsigned R;                 // 32-bit multiplier of the rescale
    int r0;                     // right shift of the rescale for a tile at the global exponent
    int s;                      // sub-child bits
    int E;                      // global exponent = max tile exponent
    int LS;                     // left shift of the exact CDF
    int nt;                     // tiles in the pool
    int valid;
};

// Where a POOL (the particle set of one step) lives.
// PEERS == false (one GPU): the local pointers w32 / x / wp.
// PEERS == true  (particles sharded over G GPUs, each a power-of-two 2^log2n of them): pool index i
// lives on GPU i >> log2n at local index i & (2^log2n - 1).  Every GPU keeps its share in one
// arena with the same layout; base[g] is GPU g's arena (a CUDA-IPC peer mapping for g != self),
// read over NVLink by the same loads that read local HBM -- the migration of resampled particles
// IS the gather, there is no pack / send / unpack.
#define SB_MAX_PEERS 8
struct SbPool {
    const uint32_t* w32;                // [tiles * SB_TILE] fixed-point weights
    const void* x;                      // [tiles * SB_TILE] particle states
    const unsigned long long* wp;       // [tiles * SB_WARPS] exclusive prefix of the warp sums inside each tile
    const char* base[SB_MAX_PEERS];
    unsigned long long off_w32, off_x, off_wp;
    int log2n;                          // particles per GPU = 2^log2n
    int c_off;                          // global index of this GPU's first child
    int tile_off;                       // global index of this GPU's first tile
};
template <bool PEERS>
__device__ __forceinline__ const uint32_t* sb_pool_tile(const SbPool& p, int b) {
    if (!PEERS) return p.w32 + (size_t)b * SB_TILE;
    const int ts = p.log2n - 11;                                 // log2(SB_TILE) = 11
    return reinterpret_cast<const uint32_t*>(p.base[b >> ts] + p.off_w32) + (size_t)(b & ((1 << ts) - 1)) * SB_TILE;
}
template <bool PEERS>
__device__ __forceinline__ const unsigned long long* sb_pool_wp_ptr(const SbPool& p, int b) {
    if (!PEERS) return p.wp + (size_t)b * SB_WARPS;
    const int ts = p.log2n - 11;
    return reinterpret_cast<const unsigned long long*>(p.base[b >> ts] + p.off_wp) + (size_t)(b & ((1 << ts) - 1)) * SB_WARPS;
}
// the 2048 states of parent tile b (4-byte states: the discrete model), read by the parents themselves when the head
// markers carry the state (CARRY): coalesced 128-bit loads instead of the children's scattered gather
template <bool PEERS>
__device__ __forceinline__ const int* sb_pool_xtile(const SbPool& p, int b) {
    if (!PEERS) return reinterpret_cast<const int*>(p.x) + (size_t)b * SB_TILE;
    const int ts = p.log2n - 11;
    return reinterpret_cast<const int*>(p.base[b >> ts] + p.off_x) + (size_t)(b & ((1 << ts) - 1)) * SB_TILE;
}
template <typename T, bool PEERS>
__device__ __forceinline__ T sb_pool_x(const SbPool& p, int i) {
    if (!PEERS) return __ldg(reinterpret_cast<const T*>(p.x) + i);
    return __ldg(reinterpret_cast<const T*>(p.base[i >> p.log2n] + p.off_x) + (i & (int)((1u << p.log2n) - 1u)));
}

// ---------------------------------------------------------------- Philox4x32-10
struct SbU4 { uint32_t x, y, z, w; };

__device__ __forceinline__ SbU4 sb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                          uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0;
        uint32_t n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return SbU4{c0, c1, c2, c3};
}
// Same function with the ten round keys precomputed on the host (kernel parameters are constant-bank
// operands: the key schedule costs no instruction in the hot loop).
struct SbRoundKeys { uint32_t k[20]; };
__device__ __forceinline__ SbU4 sb_philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const SbRoundKeys& rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ rk.k[2 * r];
        uint32_t n2 = h0 ^ c3 ^ rk.k[2 * r + 1];
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    }
    return SbU4{c0, c1, c2, c3};
}
__device__ __forceinline__ SbU4 sb_stream4_rk(const SbRoundKeys& rk, unsigned long long idx, uint32_t t, uint32_t iter, uint32_t purpose) {
    return sb_philox_rk((uint32_t)idx, (uint32_t)(idx >> 32), t, (iter << 4) | purpose, rk);
}
__device__ __forceinline__ SbU4 sb_stream4(unsigned long long seed, unsigned long long idx, uint32_t t,
                                           uint32_t iter, uint32_t purpose) {
    return sb_philox((uint32_t)idx, (uint32_t)(idx >> 32), t, (iter << 4) | purpose,
                     (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ unsigned long long sb_r53(uint32_t a, uint32_t b) {
    return ((unsigned long long)a << 21) | (unsigned long long)(b >> 11);
}
__device__ __forceinline__ double sb_u53(uint32_t a, uint32_t b) {
    return __dmul_rn((double)sb_r53(a, b), 0x1p-53);
}

// ---------------------------------------------------------------- deterministic 2^d on the 31-bit grid
// Same IEEE operation sequence as the CPU checker used by the tests (which restates
// this definition independently); intrinsics forbid the compiler from contracting mul+add.
__device__ __forceinline__ uint32_t sb_q31_f64(double d) {
    if (!(d > -(double)(SB_QBITS + 2))) return 0u;
    double k = floor(d);
    double r = __dsub_rn(d, k);
    double t = __dmul_rn(__dsub_rn(r, 0.5), SB_LN2);
    double p = 1.0 / 39916800.0;
    p = __fma_rn(p, t, 1.0 / 3628800.0);
    p = __fma_rn(p, t, 1.0 / 362880.0);
    p = __fma_rn(p, t, 1.0 / 40320.0);
    p = __fma_rn(p, t, 1.0 / 5040.0);
    p = __fma_rn(p, t, 1.0 / 720.0);
    p = __fma_rn(p, t, 1.0 / 120.0);
    p = __fma_rn(p, t, 1.0 / 24.0);
    p = __fma_rn(p, t, 1.0 / 6.0);
    p = __fma_rn(p, t, 0.5);
    p = __fma_rn(p, t, 1.0);
    p = __fma_rn(p, t, 1.0);
    p = __dmul_rn(p, SB_SQRT2_D);
    double s = __longlong_as_double((long long)(1023 + SB_QBITS + (int)k) << 52);
    return (uint32_t)__double2ull_rz(__dmul_rn(p, s));
}
__device__ __forceinline__ uint32_t sb_q31_f32(float d) {
    if (!(d > -(float)(SB_QBITS + 2))) return 0u;
    float k = floorf(d);
    float r = __fsub_rn(d, k);
    float t = __fmul_rn(__fsub_rn(r, 0.5f), 0x1.62e430p-1f);
    float p = 1.0f / 5040.0f;
    p = __fmaf_rn(p, t, 1.0f / 720.0f);
    p = __fmaf_rn(p, t, 1.0f / 120.0f);
    p = __fmaf_rn(p, t, 1.0f / 24.0f);
    p = __fmaf_rn(p, t, 1.0f / 6.0f);
    p = __fmaf_rn(p, t, 0.5f);
    p = __fmaf_rn(p, t, 1.0f);
    p = __fmaf_rn(p, t, 1.0f);
    p = __fmul_rn(p, SB_SQRT2_F);
    // p * 2^(QBITS+k) truncated.  The scale is a power of two in the normal range (k > -29) and so is the product: the
    // multiplication is exact, and the conversion truncates -- the same integer as shifting the float's 24-bit significand
    // by hand (the oracle's C does exactly that; the hand-shifted form cost ten integer instructions per particle here).
    const float s = __int_as_float((127 + SB_QBITS + (int)k) << 23);
    return __float2uint_rz(__fmul_rn(p, s));
}

// Rescaled tile-local prefix: (c * R) >> rb with c <= 2^38 and a 26-bit R (product < 2^64).
__device__ __forceinline__ unsigned long long sb_sys_delta(unsigned long long c, unsigned R, int rb) {
    return rb >= 64 ? 0ull : (c * (unsigned long long)R) >> rb;
}
// Number of children k in [0,m) whose threshold U + k * 2^s lies strictly below C (strict `<` as
// in src/rand.jl:9).  Monotone in C.
__device__ __forceinline__ int sb_sys_F(unsigned long long C, unsigned long long U, int s, int m) {
    if (C <= U) return 0;
    const unsigned long long k = ((C - U - 1ull) >> s) + 1ull;
    return k >= (unsigned long long)m ? m : (int)k;
}

// Plan of the systematic rescale from the exact total of the pool (integers only after one fp64 division,
// whose result is turned into a 26-bit multiplier with an explicit ceiling: every GPU, the tile scan and the
// one-tile kernel get the same bits).  Child k's threshold is U + k * 2^s on the rescaled CDF.
__device__ __forceinline__ void sb_plan_rescale(SbStepInfo& si, unsigned long long total, long long n, long long m,
                                                int E, int LS, int nt) {
    si.total = total; si.n = n; si.m = m; si.E = E; si.LS = LS; si.nt = nt;
    si.Tt = 0ull; si.U = 0ull; si.R = 0u; si.r0 = 0; si.s = 0;
    si.valid = (E != SB_E_EMPTY && total != 0ull) ? 1 : 0;
    if (si.valid) {
        const int clm = m <= 1 ? 0 : 64 - __clzll(m - 1);
        int s = 62 - clm;
        if (s > 36) s = 36;
        const double ratio = __ddiv_rn(scalbn((double)m, s + LS), __ull2double_rn(total));
        const unsigned long long rb_ = (unsigned long long)__double_as_longlong(ratio);
        int x = (int)((rb_ >> 52) & 0x7ffull) - 1022;                      // ratio = f * 2^x, f in [0.5, 1)
        const unsigned long long mant = (rb_ & 0xfffffffffffffull) | (1ull << 52);   // f * 2^53
        unsigned long long R = (mant >> 27) + ((mant & ((1ull << 27) - 1ull)) ? 1ull : 0ull);   // ceil(f * 2^26)
        if (R >= 67108864ull) { R = 33554432ull; x += 1; }
        int r0 = 26 - x;
        if (r0 < 0) { s += r0; r0 = 0; }
        if (s < 0) si.valid = 0;
        si.R = (unsigned)R; si.r0 = r0; si.s = s < 0 ? 0 : s;
    }
}
// Tt: total of the rescaled CDF; R53: the step's 53-bit uniform.  The offset U is spread over the slack
// Tt - (m - 1) 2^s.  Clears `valid` when the pool has no mass.
__device__ __forceinline__ void sb_plan_offset(SbStepInfo& si, unsigned long long Tt, unsigned long long R53) {
    const unsigned long long lastk = (unsigned long long)(si.m - 1) << si.s;
    if (si.valid && Tt > lastk) {
        const unsigned long long span = Tt - lastk;
        si.U = (__umul64hi(R53, span) << 11) | ((R53 * span) >> 53);
        si.Tt = Tt;
    } else {
        si.valid = 0;
    }
}

// ---------------------------------------------------------------- block-level primitives (256 threads)
__device__ __forceinline__ unsigned long long sb_shfl_up_u64(unsigned long long v, int d) {
    uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
    uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
    return ((unsigned long long)hi << 32) | lo;
}

// exclusive prefix of one u64 per thread across the CTA; *total gets the CTA sum.  smem: SB_WARPS+1 u64
__device__ __forceinline__ unsigned long long sb_block_excl_scan_u64(unsigned long long v, unsigned long long* sm,
                                                                     unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = sb_shfl_up_u64(inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();                       // protects sm reuse across calls
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    unsigned long long base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SB_WARPS; ++w) {
        unsigned long long s = sm[w];
        if (w < warp) base += s;
        tot += s;
    }
    *total = tot;
    return base + inc - v;
}

__device__ __forceinline__ int sb_block_max_i32(int v, int* sm) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = sm[0];
#pragma unroll
    for (int w = 1; w < SB_WARPS; ++w) r = max(r, sm[w]);
    return r;
}
__device__ __forceinline__ double sb_block_max_f64(double v, double* sm) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
#pragma unroll
    for (int w = 1; w < SB_WARPS; ++w) r = fmax(r, sm[w]);
    return r;
}
__device__ __forceinline__ float sb_block_max_f32(float v, float* sm) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sm[0];
#pragma unroll
    for (int w = 1; w < SB_WARPS; ++w) r = fmaxf(r, sm[w]);
    return r;
}
__device__ __forceinline__ unsigned long long sb_block_sum_u64(unsigned long long v, unsigned long long* sm) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, d);
        uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), d);
        v += ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = 0;
#pragma unroll
    for (int w = 0; w < SB_WARPS; ++w) r += sm[w];
    return r;
}

// 128-bit streaming accesses (each thread owns 8 consecutive 4-byte items = two 16-byte vectors)
__device__ __forceinline__ void sb_load8_u32(const uint32_t* __restrict__ p, uint32_t v[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void sb_store8_u32(uint32_t* p, const uint32_t v[8]) {
    reinterpret_cast<uint4*>(p)[0] = make_uint4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

// ---------------------------------------------------------------- programmatic dependent launch
// The kernels of a sweep alternate step -> scan -> step -> ...  Each is launched with programmatic stream
// serialisation: it may become resident and run its prologue (stage model tables, build the weight table)
// while its predecessor is still running, and calls sb_grid_dep_wait() before it touches anything the
// predecessor wrote.  Both are no-ops for a kernel launched the ordinary way.
// CAUTION: ptxas moves ld.global.nc (__ldg) loads ABOVE griddepcontrol.wait when their address does not depend on
// anything loaded after it (seen in k_tilescan<2>: the tile records were read while the step kernel was still writing
// them, 2^25-particle pools).  Whatever the predecessor wrote and this kernel addresses by thread/tile index alone is
// therefore read with __ldcg; loads whose ADDRESS comes out of such a load (or out of shared memory filled after the
// wait) cannot be hoisted and may stay __ldg.
__device__ __forceinline__ void sb_grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void sb_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- asynchronous copies (TMA bulk, cp.async) and mbarrier
__device__ __forceinline__ uint32_t sb_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void sb_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sb_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void sb_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SB_DONE;\n"
        "bra SB_WAIT;\n"
        "SB_DONE:\n"
        "}\n" ::"r"(sb_smem_addr(bar)), "r"(parity) : "memory");
}
// Split-phase CTA barrier on the hardware named barriers 1..SB_WARPS (no spinning: a waiting warp is descheduled and
// costs no issue slots -- an mbarrier try_wait loop re-issues itself every few dozen cycles, which was a fifth of
// all instructions the step kernels executed).  MEASURED SLOWER than the spinning mbarrier (76.4 vs 73.9 us per launch
// of k_step_hmm: the spinning warp resumes sooner, and the issue slots it burns are free ones), so it is compiled out
// by default (SB_OPT_NAMEDBAR 0); kept for the record.  Barrier 1 + w belongs to warp w: every warp ARRIVES on all of them
// (non-blocking) once it has published what the others need, and later WAITS on its own -- which completes when all
// SB_WARPS warps have arrived.  One wait per arrive; a full CTA barrier separates consecutive arrives (callers'
// invariant), so phases cannot overlap.  Writes before the arrive are visible after the wait (bar.arrive / bar.sync).
#ifndef SB_OPT_NAMEDBAR
#define SB_OPT_NAMEDBAR 0       // 0: the split barrier waits on its mbarrier (try_wait spin) instead of a hardware named barrier
#endif
#ifndef SB_OPT_XPREFETCH
#define SB_OPT_XPREFETCH 1      // 0: no L2 prefetch of the parent tiles' states
#endif
__device__ __forceinline__ void sb_split_arrive() {
#if SB_OPT_NAMEDBAR
#pragma unroll
    for (int j = 0; j < SB_WARPS; ++j) asm volatile("bar.arrive %0, %1;" ::"r"(j + 1), "r"(32 * SB_WARPS + 32) : "memory");
#endif
}
__device__ __forceinline__ void sb_split_wait() {
    asm volatile("bar.sync %0, %1;" ::"r"((int)(threadIdx.x >> 5) + 1), "r"(32 * SB_WARPS + 32) : "memory");
}
// Grid-wide barrier of a cooperative launch (all CTAs resident): one release-add per CTA on a counter that only ever
// grows, one thread per CTA spinning on an acquire-load until the counter reaches `target` = (count before the
// launch) + (barriers so far + 1) * gridDim.x -- the host hands out the base, so nothing is reset between launches.
// (The library's own rather than cooperative_groups' grid.sync() so that it can be split into arrive / wait: sb_grid_arrive.)
__device__ __forceinline__ void sb_grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}
// The same barrier in two halves (thread 0 of the CTA calls both; the caller puts a __syncthreads() before the arrive
// -- every thread's writes are then ordered before the release -- and one after the wait): work that needs nothing
// from the other CTAs goes between the two and is off the critical path of the step.
__device__ __forceinline__ void sb_grid_arrive(unsigned* counter) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void sb_grid_wait(unsigned* counter, unsigned target) {
    unsigned v;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
}
// one thread: `bytes` (multiple of 16, 16-byte aligned both sides) from global memory -- local HBM or a
// peer GPU's arena over NVLink -- into shared memory; completion is counted on the mbarrier
__device__ __forceinline__ void sb_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sb_smem_addr(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(sb_smem_addr(bar)) : "memory");
}
// one thread: ask for `bytes` (multiple of 16) of global memory to be brought into L2
__device__ __forceinline__ void sb_bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sb_smem_addr(dst)), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void sb_cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sb_smem_addr(dst)), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void sb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void sb_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- the pull: ancestors of one children tile
#define SB_FT_CHUNK 64
struct __align__(16) SbPullMeta {           // records of the two candidate parent tiles A, B of a children tile
    unsigned long long wp[2][SB_WARPS];
    unsigned long long Pt[2];
    int cs[2], ce[2], e[2], pad[2];
};
struct __align__(16) SbPullSmem {
    int marker[SB_TILE];                     // head markers, then the children's parents
    uint32_t w[2][SB_TILE];                  // weights of parent tiles A and B (TMA bulk copies, one tile ahead)
    SbPullMeta meta[2];                      // double-buffered (cp.async, one tile ahead)
    unsigned long long bar;                  // mbarrier: the bytes of w have landed
    unsigned long long scan[SB_WARPS + 1];
    int2 ft[SB_FT_CHUNK + 1];                // (first, last) candidate parent tile of this CTA's next children tiles
    // per children tile, written by warp SB_META_WARP one tile ahead (sb_pull_seg): the rescale constants of the two
    // staged parent tiles and the child index (relative, clamped) at each of their 8 warp boundaries + end
    long long seg_baseb[2];
    int seg_rb[2];
    int seg_k[2][SB_WARPS + 1];
};

// children (relative to c0, clamped to [0, cw]) whose threshold lies strictly below the CDF value
// Ptb + (acc >> rb);  D = baseb + (acc >> rb) with baseb = Ptb - U + 2^s - 1 - (c0 << s).
// S32: s >= 32, so D >> s is the arithmetic shift of D's high word alone.
template <bool S32>
__device__ __forceinline__ int sb_pull_k(long long baseb, unsigned long long acc, int rb, int s, int cw) {
    const long long D = baseb + (long long)(acc >> rb);
    if (S32) {
        const int k = (int)(D >> 32) >> (s - 32);
        return __vimin_s32_relu(k, cw);                          // max(min(k, cw), 0) in one instruction
    }
    const long long k = D >> s;
    return k <= 0 ? 0 : (k >= (long long)cw ? cw : (int)k);
}

// 64-bit accumulate of a 32 x 32 -> 64 product in ONE instruction (IMAD.WIDE.U32 with a 64-bit addend)
__device__ __forceinline__ unsigned long long sb_mad_wide(uint32_t a, uint32_t b, unsigned long long c) {
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}

// One parent tile's contribution to the children [c0, c0 + cw): warp scan of its weights on top of
// the warp prefix the producer left, then every parent that owns children in range writes its
// index at its first child's slot ("head marker").  No block-level synchronisation.
// A warp whose 256 parents own no child in range (k1 <= k0: the child index at its prefix and at the next
// warp's, looked up by the caller) skips everything: a children tile overlaps about one tile's worth of
// parents out of the two staged.
// wsrc: the tile's weights (shared-memory stage or global memory), read only by warps that work.
//
// Inside a thread the CDF is evaluated exactly ONCE with the full 64-bit formula (at the thread's own
// exclusive prefix c):  D0 = baseb + (c R >> rb),  K = D0 >> s.  For the prefix c + a (a < 2^30: partial
// sums of the thread's 8 weights) the nested-floor identity gives
//     (baseb + ((c + a) R >> rb)) >> s  ==  K + ((Z0 + a R) >> (rb + s)),   Z0 = (D0 mod 2^s) 2^rb + (c R mod 2^rb),
// so each further parent costs one multiply-accumulate into Z and one 32-bit shift of Z's high word
// (rb + s in [32, 62], the ordinary case; other shifts take the literal formula).
// CARRY: a head marker is (parent index << xb) | parent state, so the child that takes the marker has its parent's
// state without a gather; the parent reads its own 8 states (xsrc: the tile's state array, global memory) with two
// coalesced 128-bit loads issued behind the warp scan.  Markers still ascend with the slot (the index dominates).
// XPRE: the caller loaded the states already (xpa, xpc: in flight since before the weights landed).
template <bool COH>
__device__ __forceinline__ int4 sb_ld_x4(const int* p, int i) {
    return COH ? __ldcg(reinterpret_cast<const int4*>(p) + i) : __ldg(reinterpret_cast<const int4*>(p) + i);
}
template <bool S32, bool CARRY, bool COH = false, bool XPRE = false>
__device__ __forceinline__ void sb_pull_tile(const SbStepInfo& si, int b, const uint32_t* __restrict__ wsrc,
                                             long long baseb, int rb, unsigned long long wbase, int cw, SbPullSmem& sm,
                                             const int* __restrict__ xsrc = nullptr, int xb = 0,
                                             int4 xpa = make_int4(0, 0, 0, 0), int4 xpc = make_int4(0, 0, 0, 0)) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int s = si.s;
    const unsigned R = si.R;
    uint32_t w[SB_ITEMS];
    {
        const uint4 u = reinterpret_cast<const uint4*>(wsrc)[tid * 2], v = reinterpret_cast<const uint4*>(wsrc)[tid * 2 + 1];
        w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w; w[4] = v.x; w[5] = v.y; w[6] = v.z; w[7] = v.w;
    }
    unsigned tsum = 0;                                           // 8 * 2^27 fits 32 bits
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) tsum += w[j];
    unsigned long long inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = sb_shfl_up_u64(inc, d);
        if (lane >= d) inc += o;
    }
    if (tsum == 0) return;
    int4 xa = xpa, xc = xpc;
    if (CARRY && !XPRE) {                                        // in flight during the CDF evaluation below
        // COH: the states were written earlier in this same launch (resident sweep kernel): coherent loads
        xa = sb_ld_x4<COH>(xsrc, tid * 2);
        xc = sb_ld_x4<COH>(xsrc, tid * 2 + 1);
    }
    const unsigned long long c = wbase + inc - tsum;            // tile-local exclusive prefix of this thread
    unsigned long long acc = c * (unsigned long long)R;
    const int base = b * SB_TILE + tid * SB_ITEMS;
    const int sh = rb + s;
    // hi[j]: children (relative, clamped) below the END of parent j; parent j owns the slots [hi[j-1], hi[j])
    int lo0, hi[SB_ITEMS];
    if (S32 && sh <= 62) {
        const long long D0 = baseb + (long long)(acc >> rb);
        const int K = min((int)(D0 >> 32) >> (s - 32), 1 << 29);   // children below this thread: < 2^28, no overflow
        unsigned long long Z = (((unsigned long long)D0 & ((1ull << s) - 1ull)) << rb) | (acc & ((1ull << rb) - 1ull));
        const int sh32 = sh - 32;
        lo0 = __vimin_s32_relu(K, cw);
        hi[SB_ITEMS - 1] = __vimin_s32_relu(K + (int)((uint32_t)(sb_mad_wide(tsum, R, Z) >> 32) >> sh32), cw);
        if (hi[SB_ITEMS - 1] <= lo0) return;                     // this thread's parents own no children in range
#pragma unroll
        for (int j = 0; j < SB_ITEMS - 1; ++j) {
            Z = sb_mad_wide(w[j], R, Z);
            hi[j] = __vimin_s32_relu(K + (int)((uint32_t)(Z >> 32) >> sh32), cw);
        }
    } else {
        lo0 = sb_pull_k<S32>(baseb, acc, rb, s, cw);
        hi[SB_ITEMS - 1] = sb_pull_k<S32>(baseb, acc + (unsigned long long)tsum * R, rb, s, cw);
        if (hi[SB_ITEMS - 1] <= lo0) return;
#pragma unroll
        for (int j = 0; j < SB_ITEMS - 1; ++j) {
            acc += (unsigned long long)w[j] * R;
            hi[j] = sb_pull_k<S32>(baseb, acc, rb, s, cw);
        }
    }
    // head markers; and the owner of the first 256-slot boundary above lo0 (the first slot of the next warp of
    // children) writes a marker there too, so the fill of the markers never has to look across warps:
    // it is parent number #{j : hi[j] <= bnd} of this thread
    const int bnd = (lo0 | 255) + 1;
    if (CARRY) {
        const int xs[SB_ITEMS] = {xa.x, xa.y, xa.z, xa.w, xc.x, xc.y, xc.z, xc.w};
        const int bsh = base << xb;
        int mk[SB_ITEMS];
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) mk[j] = bsh + (j << xb) + xs[j];
        int lo = lo0, bv = mk[0];
#pragma unroll
        for (int j = 0; j < SB_ITEMS; ++j) {
#if !(defined(SB_ABLATE) && (SB_ABLATE == 6 || SB_ABLATE == 7))     /* diagnostics: 6 = no head-marker stores (wrong results) */
            if (lo < hi[j]) sm.marker[lo] = mk[j];
#elif SB_ABLATE == 6
            if (lo < hi[j] && mk[j] == -12345) sm.marker[lo] = mk[j];
#endif
#if defined(SB_ABLATE) && SB_ABLATE == 7                            /* diagnostics: 7 = head markers at conflict-free slots (wrong results) */
            if (lo < hi[j]) sm.marker[(lo & 0x700) + j * 32 + lane] = mk[j];
#endif
            if (j < SB_ITEMS - 1 && hi[j] <= bnd) bv = mk[j + 1];  // hi ascends: the last assignment is parent #{j : hi[j] <= bnd}
            lo = hi[j];
        }
        if (hi[SB_ITEMS - 1] > bnd) {
            sm.marker[bnd] = bv;
            if (hi[SB_ITEMS - 1] > bnd + 256) {
                lo = lo0;
#pragma unroll
                for (int j = 0; j < SB_ITEMS; ++j) {
#pragma unroll 1
                    for (int q = (lo | 255) + 1; q < hi[j]; q += 256) sm.marker[q] = mk[j];
                    lo = hi[j];
                }
            }
        }
        return;
    }
    int lo = lo0, nbelow = 0;
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) {
        if (lo < hi[j]) sm.marker[lo] = base + j;
        if (j < SB_ITEMS - 1) nbelow += hi[j] <= bnd ? 1 : 0;
        lo = hi[j];
    }
    if (hi[SB_ITEMS - 1] > bnd) {
        sm.marker[bnd] = base + nbelow;
        if (hi[SB_ITEMS - 1] > bnd + 256) {                       // a parent with hundreds of children: every boundary it covers
            lo = lo0;
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
#pragma unroll 1
                for (int q = (lo | 255) + 1; q < hi[j]; q += 256) sm.marker[q] = base + j;
                lo = hi[j];
            }
        }
    }
}

// (first, last) candidate parent tile of the children tile that starts at global child c0
__device__ __forceinline__ int2 sb_pull_range(const SbStepInfo& si, const int* __restrict__ first_tile, int c0) {
    const int m = (int)si.m;
    if (c0 >= m) return make_int2(0, -1);                        // no children here: nothing is pulled
    const int cb = c0 / SB_TILE;
    // written by the scan this kernel depends on, at an address known before the wait: NOT __ldg (see sb_grid_dep_wait)
    return make_int2(__ldcg(first_tile + cb), (c0 + SB_TILE < m) ? __ldcg(first_tile + cb + 1) : si.nt - 1);
}

// Warps 0 and SB_META_WARP (all 32 lanes each): start the copies of the two candidate parent tiles b_lo, b_lo+1 of
// an upcoming children tile -- weights by TMA bulk copy (one lane of warp 0), tile records by 4/8-byte cp.async
// (warp SB_META_WARP, which later waits for them and builds the boundary table: sb_pull_seg) -- into
// sm.w / sm.meta[slot].  The per-tile serial duties are spread over three warps (the third flushes the tile totals).
// The pool may live on a peer GPU: the copies then run over NVLink.
#define SB_META_WARP  1
#define SB_FLUSH_WARP 2
template <bool PEERS, bool CARRY = false>
__device__ __forceinline__ void sb_pull_prefetch(const SbStepInfo& si, const SbPool& pool, const int* __restrict__ e,
                                                 const unsigned long long* __restrict__ Pt, const int* __restrict__ child_start,
                                                 int b_lo, int slot, SbPullSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bA = b_lo, bB = min(b_lo + 1, si.nt - 1);
    if (warp == 0) {
        if (lane == 0) {
            sb_mbar_expect_tx(&sm.bar, 2u * SB_TILE * 4u);
            sb_bulk_g2s(sm.w[0], sb_pool_tile<PEERS>(pool, bA), SB_TILE * 4u, &sm.bar);
            sb_bulk_g2s(sm.w[1], sb_pool_tile<PEERS>(pool, bB), SB_TILE * 4u, &sm.bar);
            if (CARRY && SB_OPT_XPREFETCH) {
                // the parents read their own states one tile from now (sb_pull_tile): have them in L2 by then
                // (local memory only: a peer's memory is not cached here)
                const int ts = pool.log2n - 11, self = PEERS ? (pool.tile_off >> ts) : 0;
                if (!PEERS || (bA >> ts) == self) sb_bulk_prefetch_l2(sb_pool_xtile<PEERS>(pool, bA), SB_TILE * 4u);
                if (bB != bA && (!PEERS || (bB >> ts) == self)) sb_bulk_prefetch_l2(sb_pool_xtile<PEERS>(pool, bB), SB_TILE * 4u);
            }
        }
        return;
    }
    // tile records: lanes 0-15 the 2 x 8 warp prefixes, 16-17 Pt (8 bytes each); 18-19 child_start, 20-21 child_start + 1,
    // 22-23 e (4 bytes each).  Source and destination are selected per lane (no divergent branches).
    SbPullMeta& mt = sm.meta[slot];
    const int which = lane < 16 ? (lane >> 3) : (lane & 1);
    const int bsel = which ? bB : bA;
    const unsigned long long* src8 = lane < 16 ? sb_pool_wp_ptr<PEERS>(pool, bsel) + (lane & 7) : Pt + bsel;
    unsigned long long* dst8 = lane < 16 ? &mt.wp[which][lane & 7] : &mt.Pt[which];
    const int* src4 = lane < 22 ? child_start + bsel + (lane >= 20 ? 1 : 0) : e + bsel;
    int* dst4 = lane < 20 ? &mt.cs[which] : (lane < 22 ? &mt.ce[which] : &mt.e[which]);
    if (lane < 18) sb_cp_async8(dst8, src8);
    else if (lane < 24) sb_cp_async4(dst4, src4);
    sb_cp_async_commit();
}

// Warp SB_META_WARP only, after the tile records of (rng, slot) were requested by this same warp (sb_pull_prefetch): wait for
// them and evaluate the CDF at the 2 x (8 warp boundaries + tile end) ONCE for the CTA (18 lanes, straight-line
// code).  A barrier must separate this from the sb_pull_ancestors call that consumes the table.
template <bool S32>
__device__ __forceinline__ void sb_pull_seg(const SbStepInfo& si, int c0, int2 rng, int slot, SbPullSmem& sm) {
    const int lane = threadIdx.x & 31;
    const int m = (int)si.m;
    const int c1 = min(c0 + SB_TILE, m);
    const int cw = max(c1 - c0, 0);
    const int b_lo = rng.x, b_hi = rng.y;
    const int bA = b_lo, bB = min(b_lo + 1, si.nt - 1);
    sb_cp_async_wait_all();
    __syncwarp();
    const SbPullMeta& mt = sm.meta[slot];
    const int which = lane >= SB_WARPS + 1 ? 1 : 0, bi = lane - which * (SB_WARPS + 1);
    const int cs = mt.cs[which], ce = mt.ce[which];
    bool use = !(ce <= c0 || cs >= c1 || ce == cs);
    if (which) use = use && bB != bA && bB <= b_hi;
    const int rb = (si.r0 + (si.E - mt.e[which])) & 63;           // < 64 when the tile owns children (else unused)
    const long long baseb = (long long)(mt.Pt[which] - si.U) + ((1ll << si.s) - 1ll) - ((long long)c0 << si.s);
    int kk = sb_pull_k<S32>(baseb, mt.wp[which][bi & (SB_WARPS - 1)] * (unsigned long long)si.R, rb, si.s, cw);
    if (bi == SB_WARPS) kk = min(max(ce - c0, 0), cw);             // the last warp ends at the tile's own child_start
    if (!use) kk = 0;
    if (lane < 2 * (SB_WARPS + 1)) sm.seg_k[which][bi] = kk;
    if (bi == 0 && lane < 2 * (SB_WARPS + 1)) { sm.seg_baseb[which] = baseb; sm.seg_rb[which] = rb; }
}

// Once per CTA, before its first children tile (the mbarrier is initialised and a barrier has passed): clear the
// marker array, request the first tile's parents and build its boundary table.
template <bool PEERS, bool CARRY = false>
__device__ __forceinline__ void sb_pull_first(const SbStepInfo& si, const SbPool& pool, const int* __restrict__ e,
                                              const unsigned long long* __restrict__ Pt, const int* __restrict__ child_start,
                                              int c0, int2 rng, SbPullSmem& sm) {
    const int tid = threadIdx.x;
    const int4 neg = make_int4(-1, -1, -1, -1);
    reinterpret_cast<int4*>(sm.marker)[tid * 2] = neg;
    reinterpret_cast<int4*>(sm.marker)[tid * 2 + 1] = neg;
    if ((tid >> 5) == 0 || (tid >> 5) == SB_META_WARP) sb_pull_prefetch<PEERS, CARRY>(si, pool, e, Pt, child_start, rng.x, 0, sm);
    if ((tid >> 5) == SB_META_WARP) {
        if (si.s >= 32) sb_pull_seg<true>(si, c0, rng, 0, sm); else sb_pull_seg<false>(si, c0, rng, 0, sm);
    }
    __syncthreads();
}

// Computes, for the children [c0, c0+SB_TILE) of this CTA, the pool index of each child's parent
// under systematic resampling on the weight grid; thread `tid` receives children c0+8*tid .. +7.
// The weights and records of the candidate parent tiles [rng.x, rng.x+1] were requested one tile ahead
// (sb_pull_prefetch, slot/parity of this tile) and the boundary table of this tile was built behind an
// earlier barrier (sb_pull_first / sb_pull_seg); further candidates -- rare: a children tile that spans
// more than two parent tiles -- are read directly.  When `more`, the copies for the next children
// tile (first candidate next_b_lo) are started as soon as this tile's staging buffers are free.
// ONE barrier.  The marker array is all -1 on entry and on exit (every thread resets the slots it reads);
// the caller runs sb_pull_seg for the next tile (warp SB_META_WARP) and at least one barrier before the next call.
// GEN: the record arrays (e, Pt, child_start, warp prefixes) may be shared memory or memory written earlier in the
// same launch (the resident sweep kernel): ordinary coherent loads instead of ld.global.nc.
template <bool GEN, typename T>
__device__ __forceinline__ T sb_ld_rec(const T* p) { return GEN ? *p : __ldg(p); }          // Pt, child_start: shared memory when GEN
template <bool GEN, typename T>
__device__ __forceinline__ T sb_ld_recg(const T* p) { return GEN ? __ldcg(p) : __ldg(p); }  // e, warp prefixes: always global
// CARRY: anc[] receives (parent index << xb) | parent state (see sb_pull_tile).
template <bool PEERS, bool S32, bool GEN = false, bool CARRY = false>
__device__ __forceinline__ void sb_pull_ancestors(const SbStepInfo& si, const SbPool& pool,
                                                  const int* __restrict__ e, const unsigned long long* __restrict__ Pt,
                                                  const int* __restrict__ child_start, int c0, int2 rng, int slot,
                                                  uint32_t parity, bool more, int next_b_lo, SbPullSmem& sm,
                                                  int anc[SB_ITEMS], int xb = 0) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = (int)si.m;
    const int c1 = min(c0 + SB_TILE, m);
    const int cw = max(c1 - c0, 0);
    const int b_lo = rng.x, b_hi = rng.y;
    const int bA = b_lo, bB = min(b_lo + 1, si.nt - 1);
    sb_mbar_wait(&sm.bar, parity);                               // the two staged weight tiles have landed
    {
        // (requesting the parents' states here, before the wait for the weights, was measured: 80.0 vs 76.4 us per launch)
        const SbPullMeta& mt = sm.meta[slot];
        if (sm.seg_k[0][warp + 1] > sm.seg_k[0][warp])
            sb_pull_tile<S32, CARRY, GEN>(si, bA, sm.w[0], sm.seg_baseb[0], sm.seg_rb[0], mt.wp[0][warp], cw, sm,
                                     CARRY ? sb_pool_xtile<PEERS>(pool, bA) : nullptr, xb);
#if defined(SB_ABLATE) && SB_ABLATE == 4                              /* diagnostics: at most one parent segment per warp (wrong results) */
        if (!(sm.seg_k[0][warp + 1] > sm.seg_k[0][warp]))
#endif
        if (sm.seg_k[1][warp + 1] > sm.seg_k[1][warp])
            sb_pull_tile<S32, CARRY, GEN>(si, bB, sm.w[1], sm.seg_baseb[1], sm.seg_rb[1], mt.wp[1][warp], cw, sm,
                                     CARRY ? sb_pool_xtile<PEERS>(pool, bB) : nullptr, xb);
    }
    for (int b = b_lo + 2; b <= b_hi; ++b) {
        const int cs = sb_ld_rec<GEN>(child_start + b), ce = sb_ld_rec<GEN>(child_start + b + 1);
        if (ce <= c0 || cs >= c1 || ce == cs) continue;          // no offspring in range (uniform branch)
        const unsigned long long* wpb = sb_pool_wp_ptr<PEERS>(pool, b);
        const int rb = si.r0 + (si.E - sb_ld_recg<GEN>(e + b));
        const long long baseb = (long long)(sb_ld_rec<GEN>(Pt + b) - si.U) + ((1ll << si.s) - 1ll) - ((long long)c0 << si.s);
        const unsigned long long wbase = sb_ld_recg<GEN>(wpb + warp);
        const int k0 = sb_pull_k<S32>(baseb, wbase * (unsigned long long)si.R, rb, si.s, cw);
        const int k1 = warp == SB_WARPS - 1 ? min(max(ce - c0, 0), cw)
                                            : sb_pull_k<S32>(baseb, sb_ld_recg<GEN>(wpb + warp + 1) * (unsigned long long)si.R, rb, si.s, cw);
        if (k1 > k0) sb_pull_tile<S32, CARRY, GEN>(si, b, sb_pool_tile<PEERS>(pool, b), baseb, rb, wbase, cw, sm,
                                              CARRY ? sb_pool_xtile<PEERS>(pool, b) : nullptr, xb);
    }
    // markers complete; sm.w and sm.meta[slot] are free.  (An mbarrier with a try_wait spin here instead of BAR.SYNC was
    // measured: 75.4 vs 73.9 us per launch; for the split barrier it is the other way round, see sb_split_arrive.)
    __syncthreads();
    if (more && (warp == 0 || warp == SB_META_WARP)) sb_pull_prefetch<PEERS, CARRY>(si, pool, e, Pt, child_start, next_b_lo, slot ^ 1, sm);
    // every child takes the last head marker at or before its slot.  Markers ascend with the slot (the CDF is
    // monotone) and the first slot of every warp has one, so the carry into a thread is the largest marker of
    // the nearest earlier thread OF ITS WARP that has one: a ballot and one shuffle
    int v[SB_ITEMS];
    {
        const int4 a = reinterpret_cast<const int4*>(sm.marker)[tid * 2];
        const int4 b = reinterpret_cast<const int4*>(sm.marker)[tid * 2 + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        const int4 neg = make_int4(-1, -1, -1, -1);
        reinterpret_cast<int4*>(sm.marker)[tid * 2] = neg;        // left clean for the next children tile
        reinterpret_cast<int4*>(sm.marker)[tid * 2 + 1] = neg;
    }
    const int inc = __vimax3_s32(__vimax3_s32(v[0], v[1], v[2]), __vimax3_s32(v[3], v[4], v[5]), max(v[6], v[7]));
    const unsigned has = __ballot_sync(0xffffffffu, inc >= 0);
    const unsigned below = has & ((1u << lane) - 1u);
    int prev = __shfl_sync(0xffffffffu, inc, (31 - __clz(below)) & 31);
    if (!below) prev = 0;
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) { prev = max(v[j], prev); anc[j] = prev; }   // running max: each child takes the last head at or before it
}

// Candidate parent tiles of this CTA's next SB_FT_CHUNK + 1 children tiles (tile, tile + grid, ...), looked
// up once per chunk; on the CTA's first tile also starts the copies for that tile.
template <bool PEERS, bool CARRY = false>
__device__ __forceinline__ void sb_pull_ranges(const SbStepInfo& si, const SbPool& pool, const int* __restrict__ e,
                                               const unsigned long long* __restrict__ Pt, const int* __restrict__ child_start,
                                               const int* __restrict__ first_tile, int tile, int ntiles, bool first, SbPullSmem& sm) {
    __syncthreads();                                             // previous chunk no longer read
    if (threadIdx.x <= SB_FT_CHUNK) {
        const long long tl = (long long)tile + (long long)threadIdx.x * gridDim.x;
        sm.ft[threadIdx.x] = tl < ntiles ? sb_pull_range(si, first_tile, pool.c_off + (int)tl * SB_TILE) : make_int2(0, -1);
    }
    __syncthreads();
    if (first) sb_pull_first<PEERS, CARRY>(si, pool, e, Pt, child_start, pool.c_off + tile * SB_TILE, sm.ft[0], sm);
}

// block maxima with ONE barrier (the caller's later barriers protect the scratch array)
__device__ __forceinline__ float sb_block_max_f32_1(float v, float* sm) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    const float4 a = reinterpret_cast<const float4*>(sm)[0], b = reinterpret_cast<const float4*>(sm)[1];
    return fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
}
__device__ __forceinline__ double sb_block_max_f64_1(double v, double* sm) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
#pragma unroll
    for (int w = 1; w < SB_WARPS; ++w) r = fmax(r, sm[w]);
    return r;
}

// Where the tile records (e, Q) of the pool being PRODUCED go (local tile index).  Single GPU: the
// running maximum of the exponents is kept with atomicMax for the tile scan.  Sharded: the scan of every
// GPU reads the records from the owner's arena (no remote stores inside the hot kernel).
struct SbRecOut {
    int* e;
    unsigned long long* Q;
    int* emax;                          // {exponent maximum, E_top, speculative total (u64)} of the share this GPU produces; nullptr: the
                                        // step has no factor statement (no scan follows, nothing may accumulate)
};
__device__ __forceinline__ void sb_put_e(const SbRecOut& r, int tile, int e) {
    r.e[tile] = e;
    if (r.emax && e != SB_E_EMPTY) atomicMax(r.emax, e);
}
__device__ __forceinline__ void sb_put_Q(const SbRecOut& r, int tile, unsigned long long q) { r.Q[tile] = q; }

// Sharded pool: the LAST CTA of a step kernel to finish tells every peer that this GPU's share of the pool
// (weights, states, warp prefixes, tile records) is complete: one 4-byte store over NVLink into flags[self]
// of every arena -- compute and notification in one kernel, no separate launch.  Peers wait for the flags
// in their tile scan.  counter == nullptr: not sharded, nothing to do.
// With the flag travels this GPU's SUMMARY of the pool share it produced (acc != nullptr; SB_REC_OFF bytes behind the
// flags of every arena, slot [rec_slot][self]): {largest tile exponent, the exponent the speculative total was taken
// at, the total} -- the words the single-GPU scan finds behind its exponent maximum (sb_spec_term).  Every GPU's scan
// then has the global exponent (a max over `world` records instead of a cluster-wide reduction over all tiles) and, when
// the exponents agree, the exact total without its first pass.  Two slots, alternating per record: a GPU can publish
// record r + 2 only after every peer has raised the flag of record r + 1, i.e. has finished the scan that read record r.
#define SB_REC_OFF 64
struct SbSignal {
    unsigned* counter;                  // CTAs of this launch that have finished
    char* base[SB_MAX_PEERS];
    unsigned long long off_flags;
    int world, self;
    unsigned epoch;
    int* acc;                           // this GPU's {emax, E_top, total (u64)} accumulators, or nullptr
    int rec_slot;
};
__device__ __forceinline__ void sb_signal_done(const SbSignal& sg) {
    if (!sg.counter) return;                                     // not sharded
    __syncthreads();                                             // the CTA's writes are ordered before thread 0's fence
    if (threadIdx.x == 0) {
        // device scope is enough here: peers read THIS GPU's memory through this GPU's L2.  The one
        // system-scope fence is the last CTA's, between everybody's writes (observed through the counter)
        // and the flag stores that leave the GPU.
        __threadfence();
        const unsigned done = atomicAdd(sg.counter, 1u) + 1u;
        if (done == gridDim.x) {
            *sg.counter = 0u;                                    // next launch
            if (sg.acc) {                                        // every CTA's atomics precede its counter increment
                __threadfence();
                const int em = atomicExch(sg.acc, SB_E_EMPTY), et = atomicExch(sg.acc + 1, SB_E_EMPTY);
                const unsigned long long tot = atomicExch(reinterpret_cast<unsigned long long*>(sg.acc + 2), 0ull);
                const unsigned long long w0 = ((unsigned long long)(unsigned)et << 32) | (unsigned)em;
                for (int p = 0; p < sg.world; ++p) {
                    volatile unsigned long long* r = reinterpret_cast<volatile unsigned long long*>(sg.base[p] + sg.off_flags + SB_REC_OFF) +
                                                     2 * (sg.rec_slot * SB_MAX_PEERS + sg.self);
                    r[0] = w0; r[1] = tot;
                }
            }
            __threadfence_system();
            for (int p = 0; p < sg.world; ++p)
                *(reinterpret_cast<volatile unsigned*>(sg.base[p] + sg.off_flags) + sg.self) = sg.epoch;
        }
    }
}

// Deferred form of sb_block_sum_wp for the persistent step kernels: sb_warp_total at the end of a tile,
// sb_flush_totals after ANY later barrier (the next tile's first one) -- no barrier of its own.
__device__ __forceinline__ void sb_warp_total(unsigned tsum, unsigned long long* sm) {
    // tsum <= 2^30: the two 16-bit halves sum to < 2^22 over a warp, so two REDUX.ADD give the exact total
    const unsigned lo = __reduce_add_sync(0xffffffffu, tsum & 0xFFFFu), hi = __reduce_add_sync(0xffffffffu, tsum >> 16);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ((unsigned long long)hi << 16) + lo;
}
// returns the tile sum to the one thread that wrote it (lane SB_WARPS of warp SB_FLUSH_WARP), 0 to everybody else
__device__ __forceinline__ unsigned long long sb_flush_totals(const unsigned long long* sm, unsigned long long* __restrict__ wp_tile,
                                                              const SbRecOut& rec, int tile) {
    const int t = (int)threadIdx.x - 32 * SB_FLUSH_WARP;          // lanes 0-8 of warp SB_FLUSH_WARP
    if (t >= 0 && t <= SB_WARPS) {                               // lane w < 8: prefix of warp w; lane 8: the tile sum
        unsigned long long pre = 0;
#pragma unroll
        for (int w = 0; w < SB_WARPS; ++w) if (w < t) pre += sm[w];
        if (t < SB_WARPS) wp_tile[t] = pre; else { sb_put_Q(rec, tile, pre); return pre; }
    }
    return 0ull;
}
// Speculative exact total (single GPU).  The tile scan's first pass sums (Q_b << LS) >> (E - e_b) over the tiles, E = the
// largest tile exponent -- which a step of the discrete model knows in advance whenever some tile contains a particle in
// the best-scoring state (E_top: virtually always).  The step kernel therefore accumulates that sum at E_top on the way
// (one thread per CTA, one 64-bit atomic per CTA at the end) into the words behind the exponent maximum:
// emax[1] = E_top, (u64*)(emax + 2) = the sum.  The scan takes it when its E equals E_top and skips its first pass (one
// cluster scan + barrier of its critical path, which is exposed between two step kernels); integer adds: the same bits.
__device__ __forceinline__ unsigned long long sb_spec_term(unsigned long long Q, int e, int E_top, int LS) {
    if (Q == 0ull || e == SB_E_EMPTY) return 0ull;
    const int sh = E_top - e;
    return sh >= 63 ? 0ull : ((Q << LS) >> sh);
}

// Tile sum of the weights a CTA just produced, and the exclusive prefix of its warp sums (what the
// pull of the NEXT step starts its warp scans from).  One barrier; the caller's own barriers keep
// `sm` (SB_WARPS u64) from being rewritten before warp 0 has read it.
__device__ __forceinline__ void sb_block_sum_wp(unsigned tsum, unsigned long long* sm, unsigned long long* __restrict__ wp_tile,
                                                unsigned long long* __restrict__ Q_tile) {
    unsigned long long v = tsum;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, d);
        uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), d);
        v += ((unsigned long long)hi << 32) | lo;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x <= SB_WARPS) {                               // thread w < 8: prefix of warp w; thread 8: the tile sum
        unsigned long long pre = 0;
#pragma unroll
        for (int w = 0; w < SB_WARPS; ++w) if (w < (int)threadIdx.x) pre += sm[w];
        if (threadIdx.x < SB_WARPS) wp_tile[threadIdx.x] = pre; else *Q_tile = pre;
    }
}
