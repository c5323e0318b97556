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
__device__ __forceinline__ unsigned long long sb_pool_wp(const SbPool& p, int b, int warp) {
    if (!PEERS) return __ldg(p.wp + (size_t)b * SB_WARPS + warp);
    const int ts = p.log2n - 11;
    return __ldg(reinterpret_cast<const unsigned long long*>(p.base[b >> ts] + p.off_wp) + (size_t)(b & ((1 << ts) - 1)) * SB_WARPS + warp);
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
    // p * 2^(31+k) truncated: exact integer arithmetic on the float's 24-bit significand
    int sh = SB_QBITS + (int)k - 23;                 // p = sig * 2^(ex-23)
    uint32_t bits = __float_as_uint(p);
    uint32_t sig = (bits & 0x7FFFFFu) | 0x800000u;
    sh += (int)(bits >> 23) - 127;
    if (sh >= 0) return sig << sh;                   // sh <= 4 because p * 2^(QBITS+k) <= 2^QBITS (+1 ulp)
    return sh > -24 ? (sig >> (-sh)) : 0u;
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

// ---------------------------------------------------------------- the pull: ancestors of one children tile
struct SbPullSmem {
    int marker[SB_TILE];
    unsigned long long scan[SB_WARPS + 1];
    int wmax[SB_WARPS];
};

// children (relative to c0, clamped to [0, cw]) whose threshold lies strictly below the CDF value
// Ptb + (acc >> rb);  D = baseb + (acc >> rb) with baseb = Ptb - U + 2^s - 1 - (c0 << s).
// S32: s >= 32, so D >> s is the arithmetic shift of D's high word alone.
template <bool S32>
__device__ __forceinline__ int sb_pull_k(long long baseb, unsigned long long acc, int rb, int s, int cw) {
    const long long D = baseb + (long long)(acc >> rb);
    if (S32) {
        const int k = (int)(D >> 32) >> (s - 32);
        return min(max(k, 0), cw);
    }
    const long long k = D >> s;
    return k <= 0 ? 0 : (k >= (long long)cw ? cw : (int)k);
}

// One parent tile's contribution to the children [c0, c0 + cw): warp scan of its weights on top of
// the warp prefix the producer left, then every parent that owns children in range writes its
// index at its first child's slot ("head marker").  No block-level synchronisation.
template <bool S32>
__device__ __forceinline__ void sb_pull_tile(const SbStepInfo& si, int b, const uint32_t w[SB_ITEMS], int eb,
                                             unsigned long long Ptb, unsigned long long wbase, int c0, int cw,
                                             int* __restrict__ marker) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int s = si.s;
    const unsigned R = si.R;
    unsigned tsum = 0;                                           // 8 * 2^27 fits 32 bits
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) tsum += w[j];
    unsigned long long inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = sb_shfl_up_u64(inc, d);
        if (lane >= d) inc += o;
    }
    if (tsum != 0) {
        const unsigned long long c = wbase + inc - tsum;        // tile-local exclusive prefix of this thread
        // tile-uniform constants; rb < 64 because the tile owns children
        const int rb = si.r0 + (si.E - eb);
        const long long baseb = (long long)(Ptb - si.U) + ((1ll << s) - 1ll) - ((long long)c0 << s);
        unsigned long long acc = c * (unsigned long long)R;
        int lo = sb_pull_k<S32>(baseb, acc, rb, s, cw);
        const int hi_t = sb_pull_k<S32>(baseb, acc + (unsigned long long)tsum * R, rb, s, cw);
        if (hi_t > lo) {                                         // this thread's parents own children in range
            const int base = b * SB_TILE + tid * SB_ITEMS;
#pragma unroll
            for (int j = 0; j < SB_ITEMS; ++j) {
                acc += (unsigned long long)w[j] * R;             // one IMAD.WIDE with 64-bit accumulate
                const int hi = (j == SB_ITEMS - 1) ? hi_t : sb_pull_k<S32>(baseb, acc, rb, s, cw);
                if (lo < hi) marker[lo] = base + j;
                lo = hi;
            }
        }
    }
}

// Computes, for the children [c0, c0+SB_TILE) of this CTA, the pool index of each child's parent
// under systematic resampling on the weight grid; thread `tid` receives children c0+8*tid .. +7.
// pool / e / Pt / child_start / first_tile describe the POOL (previous step).  Parent tiles are read
// with coalesced 128-bit loads.  The first two candidate tiles (almost always all there are) and
// their metadata are requested together, so the CTA pays one memory latency instead of a chain.
template <bool PEERS, bool S32>
__device__ __forceinline__ void sb_pull_ancestors(const SbStepInfo& si, const SbPool& pool,
                                                  const int* __restrict__ e, const unsigned long long* __restrict__ Pt,
                                                  const int* __restrict__ child_start, const int* __restrict__ first_tile,
                                                  int c0, SbPullSmem& sm, int anc[SB_ITEMS]) {
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m = (int)si.m;
    const int c1 = min(c0 + SB_TILE, m);
    const int cw = c1 - c0;
    const int cb = c0 / SB_TILE;                                 // global children tile
    const int b_lo = __ldg(first_tile + cb);
    const int b_hi = (c1 < m) ? __ldg(first_tile + cb + 1) : si.nt - 1;
    const int bA = b_lo, bB = min(b_lo + 1, si.nt - 1);
    uint32_t wA[SB_ITEMS], wB[SB_ITEMS];
    sb_load8_u32(sb_pool_tile<PEERS>(pool, bA) + tid * SB_ITEMS, wA);
    sb_load8_u32(sb_pool_tile<PEERS>(pool, bB) + tid * SB_ITEMS, wB);
    const unsigned long long wbA = sb_pool_wp<PEERS>(pool, bA, warp), wbB = sb_pool_wp<PEERS>(pool, bB, warp);
    const int csA = __ldg(child_start + bA), ceA = __ldg(child_start + bA + 1);
    const int csB = __ldg(child_start + bB), ceB = __ldg(child_start + bB + 1);
    const int eA = __ldg(e + bA), eB = __ldg(e + bB);
    const unsigned long long PtA = __ldg(Pt + bA), PtB = __ldg(Pt + bB);
    {
        const int4 neg = make_int4(-1, -1, -1, -1);
        reinterpret_cast<int4*>(sm.marker)[tid * 2] = neg;
        reinterpret_cast<int4*>(sm.marker)[tid * 2 + 1] = neg;
    }
    __syncthreads();
    if (!(ceA <= c0 || csA >= c1 || ceA == csA)) sb_pull_tile<S32>(si, bA, wA, eA, PtA, wbA, c0, cw, sm.marker);
    if (bB != bA && bB <= b_hi && !(ceB <= c0 || csB >= c1 || ceB == csB)) sb_pull_tile<S32>(si, bB, wB, eB, PtB, wbB, c0, cw, sm.marker);
    for (int b = b_lo + 2; b <= b_hi; ++b) {
        const int cs = __ldg(child_start + b), ce = __ldg(child_start + b + 1);
        if (ce <= c0 || cs >= c1 || ce == cs) continue;          // no offspring in range (uniform branch)
        uint32_t w[SB_ITEMS];
        sb_load8_u32(sb_pool_tile<PEERS>(pool, b) + tid * SB_ITEMS, w);
        sb_pull_tile<S32>(si, b, w, __ldg(e + b), __ldg(Pt + b), sb_pool_wp<PEERS>(pool, b, warp), c0, cw, sm.marker);
    }
    __syncthreads();
    // inclusive max-scan of the head markers fills every child slot with its parent
    int v[SB_ITEMS];
    {
        const int4 a = reinterpret_cast<const int4*>(sm.marker)[tid * 2];
        const int4 b = reinterpret_cast<const int4*>(sm.marker)[tid * 2 + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
#pragma unroll
    for (int j = 1; j < SB_ITEMS; ++j) v[j] = max(v[j], v[j - 1]);
    int inc = v[SB_ITEMS - 1];
    const int lane = tid & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc = max(inc, o);
    }
    if (lane == 31) sm.wmax[warp] = inc;
    int prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = -1;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < SB_WARPS; ++w) if (w < warp) prev = max(prev, sm.wmax[w]);
    prev = max(prev, 0);
#pragma unroll
    for (int j = 0; j < SB_ITEMS; ++j) anc[j] = max(v[j], prev);
}

// Tile sum of the weights a CTA just produced, and the exclusive prefix of its warp sums (what the
// pull of the NEXT step starts its warp scans from).  smem: SB_WARPS u64.  Returns the tile sum.
__device__ __forceinline__ unsigned long long sb_block_sum_wp(unsigned tsum, unsigned long long* sm,
                                                              unsigned long long* __restrict__ wp_tile) {
    unsigned long long v = tsum;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, d);
        uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), d);
        v += ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = 0, pre = 0;
#pragma unroll
    for (int w = 0; w < SB_WARPS; ++w) {
        const unsigned long long t = sm[w];
        if (w < (int)threadIdx.x) pre += t;
        r += t;
    }
    if (threadIdx.x < SB_WARPS) wp_tile[threadIdx.x] = pre;
    return r;
}
