// sb_dist_back.cuh -- the backward pass of one PMCMC iteration on a pool SHARDED over GPUs (sm_100a).
//
// Replaces, across GPUs: `retained = particles[1]` (src/pmcmc.jl:89) and `counts[p.value] += 1` for every particle
// (src/pmcmc.jl:90-92).  On one GPU this is k_count_final + T launches of k_count_back (sb_kernels.cuh): cnt_t[j] =
// how many of the N final lineages pass through pool entry j at step t, walked backwards over the ancestor rows.
// Sharded, a lineage hops between GPUs whenever a child and its parent live in different shards: ancestor rows hold
// GLOBAL pool indices, the count of entry j is added into the count array of the GPU that owns j's parent -- an
// atomic add on that GPU's arena (CUDA-IPC mapping, NVLink) -- and the steps of the pass are separated by the same
// flag barrier the forward sweep uses (every rank's last CTA raises flags[self] in every arena; the next launch of
// every rank waits for all flags).  Integer adds commute: the counts are the single-GPU counts bit for bit.
//
// Two count buffers per rank, ping-pong: step t reads AND CLEARS buf[t & 1] and accumulates into buf[(t - 1) & 1] of
// the parents' owners, so both are all-zero again when the pass ends (no memset between passes, no window in
// which a fast peer could add into a buffer its owner has yet to clear).
// The retained lineage is walked by whichever rank owns its current entry; that rank writes the entry's state into
// xstar[t] of EVERY arena (the retained particle is re-injected by the last GPU's step kernels) and the next index
// of the lineage into every arena's `lin` words.
#pragma once
#include "sb_device.cuh"

struct SbBackArgs {
    char* base[SB_MAX_PEERS];           // arenas (self included)
    unsigned long long off_cnt[2];      // int[NL] each
    unsigned long long off_lin;         // int[2]: global pool index of the retained lineage at step t in word t & 1 (a fast owner
                                        // writes step t-1's index while a slow rank still reads step t's)
    unsigned long long off_xstar;       // int[Tcap]: the xstar buffer this pass writes
    unsigned long long off_flags;       // unsigned[SB_MAX_PEERS] flags + the CTA counter behind them
    int world, self, log2n;
    long long N;                        // particles (global); pool entries >= N are the retained slot / padding
    long long n_pool;                   // N (+1 with a retained particle)
    const int* x_t;                     // local history row of step t
    const int* anc_prev;                // local row: GLOBAL ancestors drawn after step t-1 (nullptr at t == 0)
    int identity_prev;                  // step t-1 had no factor statement
    int K, t, T;
    long long pick;                     // k_count_final_dist: the child whose lineage becomes the retained particle
    unsigned long long* marg;           // [T*K] this rank's share of the counts
    unsigned wait_epoch;                // every peer's flag must have reached this before anything is read (0: no wait)
    unsigned sig_epoch;                 // raised when this rank's launch is complete
    unsigned* err;
};

__device__ __forceinline__ void sb_back_wait(const SbBackArgs& a) {
    if (a.wait_epoch != 0u && (int)threadIdx.x < a.world) {
        const volatile unsigned* f = reinterpret_cast<const volatile unsigned*>(a.base[a.self] + a.off_flags) + threadIdx.x;
        const long long t0 = clock64();
        while ((int)(*f - a.wait_epoch) < 0) {
            if (clock64() - t0 > 8000000000ll) { atomicOr(a.err, SB_ERRBIT_PEER); break; }   // ~4 s: a peer died
            __nanosleep(200);
        }
    }
    __syncthreads();
}
// every thread has fenced its own remote writes (system scope) before the barrier
__device__ __forceinline__ void sb_back_signal(const SbBackArgs& a) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* counter = reinterpret_cast<unsigned*>(a.base[a.self] + a.off_flags) + SB_MAX_PEERS;
        const unsigned done = atomicAdd(counter, 1u) + 1u;
        if (done == gridDim.x) {
            *counter = 0u;
            __threadfence_system();
            for (int p = 0; p < a.world; ++p)
                *(reinterpret_cast<volatile unsigned*>(a.base[p] + a.off_flags) + a.self) = a.sig_epoch;
        }
    }
}
__device__ __forceinline__ void sb_back_add(const SbBackArgs& a, int buf, long long j, int c) {
    const int g = (int)(j >> a.log2n);
    int* dst = reinterpret_cast<int*>(a.base[g] + a.off_cnt[buf]) + (j & ((1ll << a.log2n) - 1ll));
    if (g == a.self) atomicAdd(dst, c); else atomicAdd_system(dst, c);
}

// cnt_{T-1}[anc_last[c]] += 1 for this rank's children c < N of the final resample; the owner of child `pick`
// starts the retained lineage.  anc_prev = the final resample's local row.
__global__ void __launch_bounds__(256)
k_count_final_dist(const __grid_constant__ SbBackArgs a) {
    sb_back_wait(a);
    const long long NL = 1ll << a.log2n, c_off = (long long)a.self * NL;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NL && c_off + i < a.N) {
        const int p = a.anc_prev[i];
        sb_back_add(a, (a.T - 1) & 1, p, 1);
        if (c_off + i == a.pick)
            for (int g = 0; g < a.world; ++g) *(reinterpret_cast<volatile int*>(a.base[g] + a.off_lin) + ((a.T - 1) & 1)) = p;
    }
    sb_back_signal(a);
}

// one step of the pass (see k_count_back): eight consecutive local entries per thread
__global__ void __launch_bounds__(256)
k_count_back_dist(const __grid_constant__ SbBackArgs a) {
    extern __shared__ unsigned int s_hist[];
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) s_hist[k] = 0u;
    sb_back_wait(a);                                             // also orders s_hist's initialisation
    const long long NL = 1ll << a.log2n, c_off = (long long)a.self * NL;
    int* cnt_t = reinterpret_cast<int*>(a.base[a.self] + a.off_cnt[a.t & 1]);
    const long long j0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (j0 < NL && c_off + j0 < a.n_pool) {
        // written by the peers' atomics (they arrived in this GPU's L2): coherent loads, then cleared for step t-2
        const int4 u = __ldcg(reinterpret_cast<const int4*>(cnt_t + j0)), v = __ldcg(reinterpret_cast<const int4*>(cnt_t + j0) + 1);
        const int c[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
        reinterpret_cast<int4*>(cnt_t + j0)[0] = make_int4(0, 0, 0, 0);
        reinterpret_cast<int4*>(cnt_t + j0)[1] = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long j = c_off + j0 + i;                  // global pool index
            if (c[i] && j < a.n_pool) {
                atomicAdd(&s_hist[a.x_t[j0 + i]], (unsigned)c[i]);
                if (a.t > 0) {
                    const long long pj = (j >= a.N || a.identity_prev) ? j : (long long)a.anc_prev[j0 + i];
                    sb_back_add(a, (a.t - 1) & 1, pj, c[i]);
                }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long r = *(reinterpret_cast<const volatile int*>(a.base[a.self] + a.off_lin) + (a.t & 1));
        if ((int)(r >> a.log2n) == a.self) {                     // this rank owns the retained lineage's entry at step t
            const long long rl = r - c_off;
            const int xs = a.x_t[rl];
            const int nr = (a.t > 0 && !(r >= a.N || a.identity_prev)) ? a.anc_prev[rl] : (int)r;
            for (int g = 0; g < a.world; ++g) {
                *(reinterpret_cast<volatile int*>(a.base[g] + a.off_xstar) + a.t) = xs;
                if (a.t > 0) *(reinterpret_cast<volatile int*>(a.base[g] + a.off_lin) + ((a.t - 1) & 1)) = nr;
            }
        }
    }
    __syncthreads();
    if (a.marg)
        for (int k = threadIdx.x; k < a.K; k += blockDim.x)
            if (s_hist[k]) atomicAdd(a.marg + (size_t)a.t * a.K + k, (unsigned long long)s_hist[k]);
    sb_back_signal(a);
}

// the pass is over on EVERY rank (the last step's xstar entry has arrived, no peer will add into this rank's count
// buffers any more): the next sweep may start
__global__ void k_flag_wait(const __grid_constant__ SbBackArgs a) { sb_back_wait(a); }
