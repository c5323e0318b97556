// stochy_b200.cu -- C ABI of libstochy_b200.so (see include/stochy_b200.h).
// Host orchestration of the sm_100a kernels in sb_kernels.cuh.  No CPU fallback exists: every
// entry point that computes launches CUDA kernels and fails with SB_ECUDA when it cannot.
#include "../../include/stochy_b200.h"
#include "sb_kernels.cuh"
#include "sb_small.cuh"
#include "sb_resident.cuh"
#include "sb_dist_back.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t rc = cudaMalloc(&p, want);
        if (rc == cudaSuccess) cap = want;
        return rc;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum ModelKind { MODEL_NONE = 0, MODEL_HMM = 1, MODEL_LG = 2, MODEL_BNET = 3, MODEL_OPS = 4 };

}  // namespace

struct sb_handle {
    sb_options opt;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
    std::string err;
    sb_stats stats;
    int64_t launches = 0;

    // error word, logZ accumulators, lineage index, step infos
    unsigned* d_err = nullptr;
    DevBuf d_r53;         // per-step systematic offsets (53-bit uniforms drawn on the host)
    std::vector<unsigned long long> r53_host;
    std::vector<SbStepInfo> info_host;
    DevBuf d_info;        // SbStepInfo[max(T,1)]
    DevBuf d_misc;        // int lineage

    // grid scratch (one GPU: plain buffers; sharded: the same arrays live inside `arena`)
    DevBuf w32[2], tile_e[2], tile_Q[2], tile_wp[2], P, Pt, child_start, first_tile, cdf, lit_acc, lse_part;
    int num_sms = 148;
    std::map<std::pair<const void*, size_t>, int> occ;     // resident CTAs per kernel (persistent grids)

    // sharded pool (multi-GPU, one process per GPU): see DESIGN.md "Multi-GPU"
    int rank = 0, world = 1;
    bool dist_ready = false;
    int64_t NL = 0;                   // particles per GPU (power of two)
    void* arena = nullptr;
    size_t arena_bytes = 0;
    size_t off_w32[2] = {0, 0}, off_x[2] = {0, 0}, off_wp[2] = {0, 0}, off_e[2] = {0, 0}, off_Q[2] = {0, 0}, off_flags = 0;
    char* peer_base[SB_MAX_PEERS] = {nullptr};
    unsigned epoch = 0;
    unsigned rec_seq = 0;             // summary records published so far (slot = parity)
    int rec_slot_cur = -1;            // slot of the records that came with the latest step kernel's flags (-1: none)
    // sharded conditional SMC / PMCMC (sb_dist_reserve_history): state history rows, the backward pass's count
    // buffers, the retained lineage's index words and both retained trajectories live in the arena too
    int hist_rows = 0;                // rows reserved (0: plain SMC sweeps only)
    size_t off_xh = 0, off_cnt[2] = {0, 0}, off_lin = 0, off_xstar[2] = {0, 0};
    // Sharded pools: pool slot of step t is (slot_base + t) & 1 and slot_base advances by T per sweep, so the first
    // step kernel of a sweep never writes the slot whose tile records a slower peer's LAST scan of the previous sweep
    // may still be reading (every later overwrite is ordered behind the peers' flags; see DESIGN.md "Multi-GPU").
    unsigned slot_base = 0;
    // programmatic dependent launch between the kernels of a sweep: bit 0 = the scan may be scheduled while the step
    // kernel drains, bit 1 = the step kernel's prologue overlaps the scan.  Measured best: both on one GPU, only the
    // second on sharded pools (the scan spins on peer flags there).  Overrides: SB_PDL, SB_PDL_DIST (0..3).
    int pdl = 3;
    int pdl_dist = 2;
    int scan16 = -1;                  // 16-CTA scan cluster: -1 not probed, 0 unavailable, 1 available
    int* d_emax = nullptr;            // running max of the tile exponents of the pool in production (single GPU)
    DevBuf e_all;
    // hook staging
    DevBuf hk_in, hk_r, hk_anc, hk_out, hk_src;

    // model
    int model = MODEL_NONE;
    // HMM
    int K = 0, T = 0, x0 = -1;
    bool hmm_multi = false;           // some lookup bucket holds several thresholds (slow scan needed)
    std::vector<uint8_t> has_factor;
    DevBuf hmm_cumA, hmm_cumpi, hmm_tabA, hmm_tabpi, hmm_logF, hmm_expF, hmm_hasf, hmm_A, hmm_p0, hmm_logA, hmm_logp0, enum_buf;
    // LG
    double lg_a = 0, lg_q = 0, lg_c = 0, lg_r = 0, lg_m0 = 0, lg_P0 = 0;
    std::vector<double> lg_y;
    // OPS (op-list program, sb_model_ops)
    std::vector<sb_op_step> ops;
    int ops_D = 1;                    // latent dimension (Dirichlet programs: > 1)
    DevBuf ops_obs, ops_aux;
    // BNET
    int bn_D = 0, bn_F = 0, bn_stride = 0;
    unsigned bn_query = 0;
    DevBuf bn_sp, bn_cpt, bn_lp1, bn_lp0, bn_fp, bn_fac, bn_counts;

    // sweep state
    DevBuf xbuf, ancbuf, lwbuf, wlit, cnt[3], xstar[2], marg, joint;
    DevBuf res_trace;                 // SB_RES_TRACE diagnostics
    DevBuf res_rec;                   // resident sweep kernel: the tile records its CTAs exchange (two slots)
    DevBuf cb_live, cb_list;          // cooperative backward pass: live-entry counters + cursor, compacted (entry, count) list
    unsigned bar_count = 0;           // arrivals the grid barrier counter (d_err + 1) has seen: base of the next cooperative launch
    unsigned res_epoch = 0;           // epoch stamps handed out so far (fresh ones for every launch)
    DevBuf sm_x, sm_a, sm_xs, sm_lz;  // one-tile chains (sb_small.cuh)
    int small_cur = -1;               // >= 0: the retained particle of the last pmcmc call is in buffer small_cur of sm_xs
    int use_small = 1;                // SB_NO_SMALL=1 keeps pmcmc on the tiled kernels
    int use_carry = 1;                // SB_NO_CARRY=1: the step kernel gathers parent states instead of carrying them in the head markers
    int use_resident = 1;             // SB_NO_RESIDENT=1 keeps mid-size sweeps on the tiled kernels (one launch per step and scan)
    int coop = -1;                    // cooperative launch: -1 not probed, 0 unavailable, 1 available
    int last_route = 0;               // 0 tiled kernels, 1 resident sweep kernel, 2 one-tile chain kernel (last run call)
    int64_t sw_N = 0, sw_S = 0;       // particles and row stride of the last sweep
    int sw_T = 0;                     // steps of the last sweep (rows that exist in xbuf / ancbuf / lwbuf)
    int sw_slot = 0;                  // pool slot that holds the last sweep's final particle set
    int sw_hist = 0;                  // history rows kept
    int sw_cond = 0;
    int xstar_cur = 0;
    bool have_retained = false;

    // optional per-launch timing of the fused step kernel (sb_set_profile)
    int hook_exact = 0;     // the multinomial hook needs the exact CDF whatever the handle's mode
    int profile = 0;
    std::vector<cudaEvent_t> prof_ev;
    int prof_used = 0;
};

namespace {

int fail(sb_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}
#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(h, _e == cudaErrorMemoryAllocation ? SB_ENOMEM : SB_ECUDA,             \
                        std::string(#call) + ": " + cudaGetErrorString(_e));                   \
    } while (0)
#define LAUNCH_CHECK()                                                                         \
    do {                                                                                       \
        ++h->launches;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) return fail(h, SB_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

inline int64_t tiles_of(int64_t n) { return (n + SB_TILE - 1) / SB_TILE; }

// No sweep of the tiled kernels is current any more (a model was loaded, or the one-tile chain kernel ran, whose
// history lives in its own buffers): sb_get_history / sb_get_particles must not read stale rows.
inline void invalidate_sweep(sb_handle* h) { h->sw_N = 0; h->sw_S = 0; h->sw_T = 0; h->sw_hist = 0; h->sw_cond = 0; }

// ---- where the grid arrays of pool slot `slot` live (plain buffers, or the shared arena when sharded)
inline uint32_t* W32(sb_handle* h, int slot) {
    return h->dist_ready ? reinterpret_cast<uint32_t*>((char*)h->arena + h->off_w32[slot]) : h->w32[slot].as<uint32_t>();
}
// tile records by GLOBAL tile index (sharded: the replicated arrays every rank's step kernel pushes into)
inline int* TE(sb_handle* h, int slot) {
    return h->dist_ready ? reinterpret_cast<int*>((char*)h->arena + h->off_e[slot]) : h->tile_e[slot].as<int>();
}
inline unsigned long long* TQ(sb_handle* h, int slot) {
    return h->dist_ready ? reinterpret_cast<unsigned long long*>((char*)h->arena + h->off_Q[slot]) : h->tile_Q[slot].as<unsigned long long>();
}
// sharded: the step kernel that produces a pool raises this GPU's flag to a fresh epoch when it is done
inline SbSignal make_signal(sb_handle* h) {
    SbSignal sg;
    memset(&sg, 0, sizeof(sg));
    sg.world = 1;
    if (h->dist_ready) {
        sg.counter = reinterpret_cast<unsigned*>((char*)h->arena + h->off_flags) + SB_MAX_PEERS;
        for (int g = 0; g < h->world; ++g) sg.base[g] = h->peer_base[g];
        sg.off_flags = h->off_flags; sg.world = h->world; sg.self = h->rank; sg.epoch = ++h->epoch;
    }
    return sg;
}
// where the step kernels put the records of the tiles they produce (local tile index)
inline SbRecOut make_rec(sb_handle* h, int slot) {
    SbRecOut r;
    const int tile_off = h->dist_ready ? (int)(h->rank * (h->NL / SB_TILE)) : 0;
    r.e = TE(h, slot) + tile_off; r.Q = TQ(h, slot) + tile_off;
    r.emax = h->d_emax;               // sharded: this GPU's share; its last CTA hands the summary to every peer (sb_signal_done)
    return r;
}
// sharded: the summary record travels with the flag of the step kernel that accumulates into rec.emax (call after
// rec.emax is final); the scan that follows reads the records of all GPUs from slot h->rec_slot_cur
inline void attach_summary(sb_handle* h, SbSignal& sg, const SbRecOut& rec) {
    sg.acc = nullptr; sg.rec_slot = 0;
    h->rec_slot_cur = -1;
    if (h->dist_ready && rec.emax) { sg.acc = rec.emax; sg.rec_slot = (int)(h->rec_seq++ & 1u); h->rec_slot_cur = sg.rec_slot; }
}
inline unsigned long long* WP(sb_handle* h, int slot) {
    return h->dist_ready ? reinterpret_cast<unsigned long long*>((char*)h->arena + h->off_wp[slot]) : h->tile_wp[slot].as<unsigned long long>();
}
inline const int* TE_pull(sb_handle* h, int slot) { return TE(h, slot); }

// the pool held in slot `slot` whose particle states are at `x` (single GPU) / in the arena (sharded)
inline SbPool make_pool(sb_handle* h, int slot, const void* x, size_t off_x_override = 0) {
    SbPool p;
    memset(&p, 0, sizeof(p));
    p.w32 = W32(h, slot); p.x = x; p.wp = WP(h, slot);
    p.log2n = 31;
    if (h->dist_ready) {
        for (int g = 0; g < h->world; ++g) p.base[g] = h->peer_base[g];
        p.off_w32 = h->off_w32[slot]; p.off_x = h->off_x[slot]; p.off_wp = h->off_wp[slot];
        p.log2n = 0; while ((1ll << p.log2n) < h->NL) ++p.log2n;
        p.c_off = (int)(h->rank * h->NL);
        p.tile_off = (int)(h->rank * (h->NL / SB_TILE));
        p.x = (char*)h->arena + h->off_x[slot];
        if (off_x_override) { p.off_x = off_x_override; p.x = (char*)h->arena + off_x_override; }   // a history row
    }
    return p;
}

// host copy of the Philox4x32-10 stream (the per-step systematic offsets are drawn on the host so the
// single-CTA tile scan does not spend its serial section on ten dependent rounds)
void philox_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
uint64_t host_r53(uint64_t seed, uint64_t idx, uint32_t t, uint32_t iter, uint32_t purpose) {
    uint32_t o[4];
    philox_host((uint32_t)idx, (uint32_t)(idx >> 32), t, (iter << 4) | purpose, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return ((uint64_t)o[0] << 21) | (uint64_t)(o[1] >> 11);
}

SbRoundKeys round_keys(uint64_t seed) {
    SbRoundKeys rk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk.k[2 * r] = k0; rk.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return rk;
}

// log(mean weight) of a pool from its tile-scan record: the log-evidence increment (SURVEY D3 extension)
double log_mean_weight(const SbStepInfo& si) {
    return log((double)si.total) + (double)(si.E - SB_QBITS - si.LS) * SB_LN2 - log((double)si.n);
}

int check_err_word(sb_handle* h) {
    unsigned w = 0;
    CK(cudaMemcpyAsync(&w, h->d_err, sizeof(w), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (w) {
        cudaMemsetAsync(h->d_err, 0, sizeof(unsigned), h->stream);
        if (w & SB_ERRBIT_PEER) return fail(h, SB_EPEER, "multi-GPU barrier timed out: a peer GPU never signalled");
        if (w & SB_ERRBIT_NAN) return fail(h, SB_ENAN, "unexpected nan in rand");
        return fail(h, SB_EZEROMASS, "Distribution has zero probability mass.");
    }
    return SB_OK;
}

// scratch of the resampling grid for a pool of n_pool entries (GLOBAL count when sharded) and m children
int ensure_grid(sb_handle* h, int64_t n_pool, int64_t m) {
    const int64_t nt = tiles_of(n_pool);
    if (nt > (1 << 20) || m > 0x7fffffffLL - SB_TILE) return fail(h, SB_EINVAL, "pool too large (max 2^31 particles)");
    if (h->dist_ready) {
        if (n_pool != h->NL * h->world || (m != n_pool && m != n_pool - 1))
            return fail(h, SB_EINVAL, "sharded handle: the pool is fixed at world x N_local entries (particles, or particles + the retained one)");
    } else {
        for (int i = 0; i < 2; ++i) {
            CK(h->w32[i].ensure((size_t)nt * SB_TILE * 4));
            CK(h->tile_e[i].ensure((size_t)nt * 4));
            CK(h->tile_Q[i].ensure((size_t)nt * 8));
            CK(h->tile_wp[i].ensure((size_t)nt * SB_WARPS * 8));
        }
    }
    CK(h->P.ensure((size_t)(nt + 1) * 8));
    CK(h->Pt.ensure((size_t)(nt + 1) * 8));
    CK(h->child_start.ensure((size_t)(nt + 1) * 4));
    CK(h->first_tile.ensure((size_t)(tiles_of(m) + 1) * 4));
    return SB_OK;
}

// counter of sb_grid_barrier: the word behind the error word (allocated and zeroed together in sb_create)
inline unsigned* grid_bar(sb_handle* h) { return h->d_err + 1; }

// resident CTAs of a persistent kernel (cached per kernel and shared-memory size)
template <typename Kern>
int resident_ctas(sb_handle* h, Kern kern, size_t smem, int* out) {
    const std::pair<const void*, size_t> key((const void*)kern, smem);
    auto it = h->occ.find(key);
    if (it == h->occ.end()) {
        // static + dynamic shared memory above 48 KB needs the opt-in even when the dynamic part alone is small; the
        // attribute belongs to the kernel, not to this (kernel, size) entry: only ever raise it
        if (smem > 16 * 1024) {
            static std::mutex mu;
            static std::map<std::pair<int, const void*>, size_t> optin;   // (device, kernel) -> largest size opted into
            std::lock_guard<std::mutex> lk(mu);
            size_t& have = optin[std::make_pair(h->opt.device, (const void*)kern)];
            if (smem > have) {
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                have = smem;
            }
        }
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, SB_THREADS, smem));
        if (nb < 1) return fail(h, SB_ECUDA, "step kernel does not fit on an SM");
        if (const char* ov = getenv("SB_CTAS_PER_SM")) {             // tuning knob: fewer resident CTAs than fit
            const int want = atoi(ov);
            if (want >= 1 && want < nb) nb = want;
        }
        it = h->occ.emplace(key, nb * h->num_sms).first;
    }
    *out = it->second;
    return SB_OK;
}

// quantise (device data) into grid buffer `slot`
int launch_quantise(sb_handle* h, const void* d_data, int kind, int64_t n, int slot) {
    const int nt = (int)tiles_of(n);
    uint32_t* w32 = W32(h, slot);
    int* e = TE(h, slot);
    unsigned long long* Q = TQ(h, slot);
    unsigned long long* wp = WP(h, slot);
    if (kind == SB_W_LINEAR_F64) k_quantise<0><<<nt, SB_THREADS, 0, h->stream>>>(d_data, n, w32, e, Q, wp, h->d_emax, h->d_err);
    else if (kind == SB_W_LOG_F64) k_quantise<1><<<nt, SB_THREADS, 0, h->stream>>>(d_data, n, w32, e, Q, wp, h->d_emax, h->d_err);
    else if (kind == SB_W_LOG_F32) k_quantise<2><<<nt, SB_THREADS, 0, h->stream>>>(d_data, n, w32, e, Q, wp, h->d_emax, h->d_err);
    else return fail(h, SB_EINVAL, "unknown weight kind");
    LAUNCH_CHECK();
    return SB_OK;
}

template <int PER, bool PEERS>
int launch_tilescan_k(sb_handle* h, int ctas, const int* e, const unsigned long long* Q, int nt, int64_t n, int64_t m,
                      const unsigned long long* d_r53, SbStepInfo* info, int want_exact, const SbScanPeers& sp) {
    auto kern = k_tilescan<PER, PEERS>;
    if (ctas > SB_SCAN_CTAS) {
        const std::pair<const void*, size_t> key((const void*)kern, (size_t)1 << 40);     // "non-portable size allowed" set once
        if (h->occ.find(key) == h->occ.end()) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            h->occ.emplace(key, 1);
        }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(SB_SCAN_THREADS); cfg.stream = h->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // may be scheduled while the step kernel drains
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = ((h->dist_ready ? h->pdl_dist : h->pdl) & 1) ? 2 : 1;
    long long nn = n, mm = m;
    CK(cudaLaunchKernelEx(&cfg, kern, e, Q, nt, nn, mm, d_r53, info, h->P.as<unsigned long long>(), h->Pt.as<unsigned long long>(),
                          h->child_start.as<int>(), h->first_tile.as<int>(), want_exact, h->d_err, h->d_emax, sp));
    ++h->launches;
    return SB_OK;
}

template <bool PEERS>
int launch_tilescan_t(sb_handle* h, const int* e, const unsigned long long* Q, int nt, int64_t n, int64_t m,
                      const unsigned long long* d_r53, SbStepInfo* info, int want_exact, const SbScanPeers& sp) {
    // one cluster: as many CTAs as give one tile per thread (1, 2, 4 or 8: small pools skip the cluster barriers'
    // cross-SM latency), 8 CTAs (portable) up to 4 tiles per thread, 16 CTAs beyond that when the device takes them
    int ctas = 1;
    while (ctas < SB_SCAN_CTAS && nt > ctas * SB_SCAN_THREADS) ctas *= 2;
    if (nt > SB_SCAN_CTAS * SB_SCAN_THREADS * 4 && h->scan16 != 0) {
        if (h->scan16 < 0) {                                         // probe once
            auto kern = k_tilescan<4, PEERS>;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(SB_SCAN_CTAS_MAX); cfg.blockDim = dim3(SB_SCAN_THREADS);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = SB_SCAN_CTAS_MAX; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nc = 0;
            h->scan16 = (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                         cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess && nc >= 1) ? 1 : 0;
            cudaGetLastError();
        }
        if (h->scan16 == 1) ctas = SB_SCAN_CTAS_MAX;
    }
    const int per = (nt + SB_SCAN_THREADS * ctas - 1) / (SB_SCAN_THREADS * ctas);
    if (per <= 1) return launch_tilescan_k<1, PEERS>(h, ctas, e, Q, nt, n, m, d_r53, info, want_exact, sp);
    if (per <= 2) return launch_tilescan_k<2, PEERS>(h, ctas, e, Q, nt, n, m, d_r53, info, want_exact, sp);
    if (per <= 4) return launch_tilescan_k<4, PEERS>(h, ctas, e, Q, nt, n, m, d_r53, info, want_exact, sp);
    if (per <= 8) return launch_tilescan_k<8, PEERS>(h, ctas, e, Q, nt, n, m, d_r53, info, want_exact, sp);
    if (!PEERS) return launch_tilescan_k<0, PEERS>(h, ctas, e, Q, nt, n, m, d_r53, info, want_exact, sp);
    return fail(h, SB_EINVAL, "sharded pool too large (max 2^27 particles over all GPUs)");
}

// scan of pool `slot` (n entries, m children).  Sharded: first tell the peers that this GPU's share of
// the pool is complete, then scan all GPUs' tile records behind the flag barrier.
int launch_tilescan(sb_handle* h, int slot, int64_t n, int64_t m, const unsigned long long* d_r53, SbStepInfo* info) {
    const int nt = (int)tiles_of(n);
    const int want_exact = h->opt.resample_mode == SB_RESAMPLE_MULTINOMIAL || h->hook_exact;
    SbScanPeers sp;
    memset(&sp, 0, sizeof(sp));
    if (!h->dist_ready)
        return launch_tilescan_t<false>(h, TE(h, slot), TQ(h, slot), nt, n, m, d_r53, info, want_exact, sp);
    for (int g = 0; g < h->world; ++g) sp.base[g] = h->peer_base[g];
    sp.off_flags = h->off_flags; sp.off_e = h->off_e[slot]; sp.off_Q = h->off_Q[slot];
    sp.log2t = 0; while ((1ll << sp.log2t) < h->NL / SB_TILE) ++sp.log2t;
    sp.world = h->world; sp.self = h->rank; sp.epoch = h->epoch;      // the epoch the step kernel just signalled
    sp.rec_slot = h->rec_slot_cur;
    return launch_tilescan_t<true>(h, TE(h, slot), TQ(h, slot), nt, n, m, d_r53, info, want_exact, sp);
}

int launch_pull(sb_handle* h, int slot, const SbStepInfo* info, int64_t m_local, int* d_anc) {
    const SbPool pool = make_pool(h, slot, nullptr);
    const int ntiles = (int)tiles_of(m_local);
    // many tiles: the persistent grid with the step kernels' prefetch pipeline; few: one CTA per tile (no ramp to amortise)
    int cap = 0;
    if (ntiles >= 4 * h->num_sms) {
        int rc = h->dist_ready ? resident_ctas(h, k_pull_persistent<true>, 0, &cap) : resident_ctas(h, k_pull_persistent<false>, 0, &cap);
        if (rc != SB_OK) return rc;
    }
    if (cap > 0 && ntiles > cap) {
        if (h->dist_ready)
            k_pull_persistent<true><<<(unsigned)cap, SB_THREADS, 0, h->stream>>>(info, pool, TE_pull(h, slot), h->Pt.as<unsigned long long>(),
                                                                                 h->child_start.as<int>(), h->first_tile.as<int>(), d_anc, ntiles);
        else
            k_pull_persistent<false><<<(unsigned)cap, SB_THREADS, 0, h->stream>>>(info, pool, TE_pull(h, slot), h->Pt.as<unsigned long long>(),
                                                                                  h->child_start.as<int>(), h->first_tile.as<int>(), d_anc, ntiles);
        LAUNCH_CHECK();
        return SB_OK;
    }
    if (h->dist_ready)
        k_pull<true><<<(unsigned)ntiles, SB_THREADS, 0, h->stream>>>(info, pool, TE_pull(h, slot), h->Pt.as<unsigned long long>(),
                                                                    h->child_start.as<int>(), h->first_tile.as<int>(), d_anc);
    else
        k_pull<false><<<(unsigned)ntiles, SB_THREADS, 0, h->stream>>>(info, pool, TE_pull(h, slot), h->Pt.as<unsigned long long>(),
                                                                     h->child_start.as<int>(), h->first_tile.as<int>(), d_anc);
    LAUNCH_CHECK();
    return SB_OK;
}

// grid multinomial on pool `slot`: CDF + search.  d_r may be nullptr (Philox stream)
int launch_grid_multinomial(sb_handle* h, int slot, const SbStepInfo* info, int64_t n, int64_t m, const double* d_r,
                            unsigned t, unsigned iter, int* d_anc) {
    if (h->dist_ready) return fail(h, SB_EUNSUPPORTED, "sharded pools resample systematically");
    CK(h->cdf.ensure((size_t)n * 8));
    k_cdf<<<(unsigned)tiles_of(n), SB_THREADS, 0, h->stream>>>(info, W32(h, slot), TE(h, slot),
                                                               h->P.as<unsigned long long>(), h->cdf.as<unsigned long long>());
    LAUNCH_CHECK();
    k_multinomial_search<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(info, h->cdf.as<unsigned long long>(), d_r,
                                                                           h->opt.seed, t, iter, d_anc);
    LAUNCH_CHECK();
    return SB_OK;
}

int launch_literal(sb_handle* h, const double* d_w, int64_t n, int64_t m, const double* d_r, unsigned t, unsigned iter,
                   int* d_anc) {
    CK(h->lit_acc.ensure((size_t)n * 8));
    k_literal_cdf<<<1, 256, 0, h->stream>>>(d_w, n, h->lit_acc.as<double>(), h->d_err);
    LAUNCH_CHECK();
    k_literal_search<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(h->lit_acc.as<double>(), n, d_r, m, h->opt.seed,
                                                                       t, iter, d_anc);
    LAUNCH_CHECK();
    return SB_OK;
}

inline size_t kind_bytes(int kind) { return kind == SB_W_LOG_F32 ? 4 : 8; }

// bracket one launch of the step kernel with an event pair (at most prof_ev.size()/2 per run call)
struct ProfScope {
    sb_handle* h; bool on;
    ProfScope(sb_handle* h_, int t) : h(h_), on(false) {
        if (h->profile && (t % h->profile) == 0 && h->prof_used + 2 <= (int)h->prof_ev.size()) {
            on = true;
            cudaEventRecord(h->prof_ev[h->prof_used], h->stream);
        }
    }
    ~ProfScope() {
        if (on) { cudaEventRecord(h->prof_ev[h->prof_used + 1], h->stream); h->prof_used += 2; }
    }
};

int begin_timed(sb_handle* h) {
    h->launches = 0;
    h->prof_used = 0;
    static const int empty[4] = {SB_E_EMPTY, SB_E_EMPTY, 0, 0};   // an aborted call may have left a partial maximum / speculative total behind
    CK(cudaMemcpyAsync(h->d_emax, empty, sizeof(empty), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    return SB_OK;
}
int end_timed(sb_handle* h) {
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.device_ms = ms;
    h->stats.kernel_launches = h->launches;
    h->stats.step_kernel_ms = 0.0;
    if (h->prof_used) {
        double acc = 0.0;
        for (int i = 0; i + 1 < h->prof_used; i += 2) {
            float f = 0.f;
            CK(cudaEventElapsedTime(&f, h->prof_ev[i], h->prof_ev[i + 1]));
            acc += f;
        }
        h->stats.step_kernel_ms = acc / (h->prof_used / 2);      // average per launch
    }
    return SB_OK;
}

}  // namespace

// ======================================================================= lifecycle
extern "C" int32_t sb_version(void) { return SB_VERSION; }

extern "C" void sb_options_default(sb_options* o) {
    if (!o) return;
    o->device = 0;
    o->precision = SB_FP32;
    o->resample_mode = SB_RESAMPLE_SYSTEMATIC;
    o->retained_pick = -1;
    o->store_logw = 0;
    o->use_graph = 0;
    o->seed = 1;
}

static thread_local std::string g_create_err;

extern "C" int32_t sb_create(const sb_options* o, sb_handle** out) {
    if (!out) return SB_EINVAL;
    *out = nullptr;
    sb_options opt;
    if (o) opt = *o; else sb_options_default(&opt);
    if (opt.precision != SB_FP32 && opt.precision != SB_FP64) { g_create_err = "bad precision"; return SB_EINVAL; }
    if (opt.resample_mode < 0 || opt.resample_mode > 2) { g_create_err = "bad resample_mode"; return SB_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libstochy_b200 has no CPU fallback)";
        return SB_ECUDA;
    }
    if (opt.device < 0 || opt.device >= ndev) { g_create_err = "device ordinal out of range"; return SB_EINVAL; }
    if ((e = cudaSetDevice(opt.device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return SB_ECUDA; }
    sb_handle* h = new sb_handle();
    h->opt = opt;
    memset(&h->stats, 0, sizeof(h->stats));
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreate(&h->evt0) != cudaSuccess || cudaEventCreate(&h->evt1) != cudaSuccess ||
        cudaMalloc(&h->d_err, 16 * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(h->d_err, 0, 16 * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc(&h->d_emax, 4 * sizeof(int)) != cudaSuccess ||      // [0] exponent maximum, [1] E_top of the speculative total, [2..3] the total (u64)
        h->d_misc.ensure(64) != cudaSuccess) {
        g_create_err = std::string("CUDA init failed: ") + cudaGetErrorString(cudaGetLastError());
        sb_destroy(h);
        return SB_ECUDA;
    }
    {
        const int empty[4] = {SB_E_EMPTY, SB_E_EMPTY, 0, 0};
        cudaMemcpy(h->d_emax, empty, sizeof(empty), cudaMemcpyHostToDevice);
    }
    if (const char* np = getenv("SB_PDL")) h->pdl = atoi(np) & 3;
    if (const char* ns = getenv("SB_NO_SMALL")) h->use_small = atoi(ns) ? 0 : 1;
    if (const char* nr = getenv("SB_NO_RESIDENT")) h->use_resident = atoi(nr) ? 0 : 1;
    if (const char* nc = getenv("SB_NO_CARRY")) h->use_carry = atoi(nc) ? 0 : 1;
    if (const char* pd = getenv("SB_PDL_DIST")) h->pdl_dist = atoi(pd) & 3;
    if (cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, opt.device) != cudaSuccess || h->num_sms < 1)
        h->num_sms = 148;
    *out = h;
    return SB_OK;
}

extern "C" void sb_destroy(sb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->opt.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->d_r53, &h->d_info, &h->d_misc, &h->w32[0], &h->w32[1], &h->tile_e[0], &h->tile_e[1],
                      &h->tile_Q[0], &h->tile_Q[1], &h->tile_wp[0], &h->tile_wp[1], &h->e_all, &h->hmm_tabA, &h->hmm_tabpi, &h->hmm_A, &h->hmm_p0, &h->hmm_logA, &h->hmm_logp0, &h->enum_buf, &h->P, &h->Pt, &h->child_start, &h->first_tile, &h->cdf, &h->lit_acc, &h->lse_part,
                      &h->hk_in, &h->hk_r, &h->hk_anc, &h->hk_out, &h->hk_src, &h->hmm_cumA, &h->hmm_cumpi,
                      &h->hmm_logF, &h->hmm_expF, &h->hmm_hasf, &h->bn_sp, &h->bn_cpt, &h->bn_lp1, &h->bn_lp0,
                      &h->bn_fp, &h->bn_fac, &h->bn_counts, &h->xbuf, &h->ancbuf, &h->lwbuf, &h->wlit, &h->cnt[0],
                      &h->cnt[1], &h->cnt[2], &h->xstar[0], &h->xstar[1], &h->marg, &h->joint, &h->sm_x, &h->sm_a, &h->sm_xs, &h->sm_lz, &h->res_trace, &h->res_rec, &h->ops_obs, &h->ops_aux, &h->cb_live, &h->cb_list};
    for (DevBuf* b : bufs) b->release();
    for (int g = 0; g < SB_MAX_PEERS; ++g)
        if (h->peer_base[g] && g != h->rank) cudaIpcCloseMemHandle(h->peer_base[g]);
    if (h->arena) cudaFree(h->arena);
    if (h->d_err) cudaFree(h->d_err);
    if (h->d_emax) cudaFree(h->d_emax);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evt0) cudaEventDestroy(h->evt0);
    if (h->evt1) cudaEventDestroy(h->evt1);
    for (auto& e : h->prof_ev) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char* sb_last_error(const sb_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int32_t sb_get_stats(const sb_handle* h, sb_stats* out) {
    if (!h || !out) return SB_EINVAL;
    *out = h->stats;
    return SB_OK;
}

extern "C" int32_t sb_set_profile(sb_handle* h, int32_t every) {
    if (!h || every < 0) return SB_EINVAL;
    h->profile = every;
    if (every && h->prof_ev.empty()) {
        h->prof_ev.resize(256);
        for (auto& e : h->prof_ev) CK(cudaEventCreate(&e));
    }
    return SB_OK;
}

extern "C" int32_t sb_sync(sb_handle* h) {
    if (!h) return SB_EINVAL;
    CK(cudaSetDevice(h->opt.device));
    CK(cudaStreamSynchronize(h->stream));
    return check_err_word(h);
}
extern "C" int32_t sb_timer_start(sb_handle* h) {
    if (!h) return SB_EINVAL;
    CK(cudaEventRecord(h->evt0, h->stream));
    return SB_OK;
}
extern "C" int32_t sb_timer_stop(sb_handle* h, double* ms) {
    if (!h || !ms) return SB_EINVAL;
    CK(cudaEventRecord(h->evt1, h->stream));
    CK(cudaEventSynchronize(h->evt1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, h->evt0, h->evt1));
    *ms = f;
    return SB_OK;
}

// ======================================================================= models
extern "C" int32_t sb_model_discrete_hmm(sb_handle* h, int32_t K, int32_t T, int32_t x0, const double* pi,
                                         const double* A, const double* logF, const uint8_t* has_factor) {
    if (!h) return SB_EINVAL;
    if (!A || !logF || K < 1 || T < 1) return fail(h, SB_EINVAL, "discrete_hmm: bad arguments");
    if (K > 64) return fail(h, SB_EUNSUPPORTED, "discrete_hmm: K > 64 is not supported on the GPU path (no CPU fallback)");
    if (x0 >= K) return fail(h, SB_EINVAL, "discrete_hmm: x0 out of range");
    if (x0 < 0 && !pi) return fail(h, SB_EINVAL, "discrete_hmm: pi required when x0 < 0");
    CK(cudaSetDevice(h->opt.device));
    // sequential running sums, exactly the accumulation of src/rand.jl:3-7
    // u = R * 2^-32 (32-bit R): `u < acc` is exactly `R < ceil(acc * 2^32)`; saturated to 32 bits
    auto thr_of = [](double acc) -> uint32_t {
        const double t = ceil(acc * 4294967296.0);
        if (!(t > 0.0)) return 0u;
        return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    };
    // inverse-CDF lookup words, 256 buckets per row (bucket g = draws R with R >> 24 == g):
    //   bits 0..5   lo  = #{k < K-1 : thr[k] <= g << 24}        (where the scan of src/rand.jl:6-12 stands at the bucket start)
    //   bit  7      several thresholds fall inside the bucket (the kernel scans the thresholds from lo)
    //   bits 8..31  t24 - 1, t24 = low 24 bits of the first threshold inside the bucket, or 2^24 if there is none
    // so that  x = lo + ((R << 8) > word)  whenever bit 7 is clear: one shared-memory load and one compare.
    const int G = 1 << SB_LG_BUCKETS;
    std::vector<uint32_t> cumA((size_t)K * K, 0u), cumpi((size_t)K, 0u), tabA((size_t)K * G, 0u), tabpi((size_t)G, 0u);
    bool multi = false;
    auto fill_tab = [&](const uint32_t* thr, uint32_t* tab) {
        for (int g = 0; g < G; ++g) {
            const uint32_t r0 = (uint32_t)g << (32 - SB_LG_BUCKETS), r1 = r0 + ((1u << (32 - SB_LG_BUCKETS)) - 1u);
            int lo = 0;
            while (lo < K - 1 && thr[lo] <= r0) ++lo;
            int cnt = 0;
            while (lo + cnt < K - 1 && thr[lo + cnt] <= r1) ++cnt;
            const uint32_t t24 = cnt ? (thr[lo] & 0xFFFFFFu) : (1u << 24);
            tab[g] = (uint32_t)lo | ((t24 - 1u) << 8) | (cnt > 1 ? 128u : 0u);
            if (cnt > 1) multi = true;
        }
    };
    std::vector<double> expF((size_t)T * K);
    for (int r = 0; r < K; ++r) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) {
            const double p = A[(size_t)r * K + k];
            if (p != p) return fail(h, SB_ENAN, "unexpected nan in rand");
            if (!(p >= 0.0)) return fail(h, SB_EINVAL, "discrete_hmm: negative transition probability");
            acc += p; cumA[(size_t)r * K + k] = thr_of(acc);
        }
    }
    if (x0 < 0) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) {
            // rand(ps) on a NaN entry is the reference's error (src/rand.jl:8,15); a negative entry is a caller bug
            if (pi[k] != pi[k]) return fail(h, SB_ENAN, "unexpected nan in rand");
            if (!(pi[k] >= 0.0)) return fail(h, SB_EINVAL, "discrete_hmm: negative initial probability");
            acc += pi[k]; cumpi[k] = thr_of(acc);
        }
    }
    for (int r = 0; r < K; ++r) fill_tab(cumA.data() + (size_t)r * K, tabA.data() + (size_t)r * G);
    fill_tab(cumpi.data(), tabpi.data());
    h->hmm_multi = multi;
    for (size_t i = 0; i < (size_t)T * K; ++i) expF[i] = exp(logF[i]);       // src/pmcmc.jl:55 (literal mode)
    h->has_factor.assign((size_t)T, 1);
    if (has_factor) for (int t = 0; t < T; ++t) h->has_factor[t] = has_factor[t] ? 1 : 0;
    CK(h->hmm_cumA.ensure(cumA.size() * 4));
    CK(h->hmm_cumpi.ensure(cumpi.size() * 4));
    CK(h->hmm_tabA.ensure(tabA.size() * 4));
    CK(h->hmm_tabpi.ensure(tabpi.size() * 4));
    CK(cudaMemcpyAsync(h->hmm_tabA.p, tabA.data(), tabA.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->hmm_tabpi.p, tabpi.data(), tabpi.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(h->hmm_logF.ensure((size_t)T * K * 8));
    CK(h->hmm_expF.ensure((size_t)T * K * 8));
    CK(h->hmm_hasf.ensure((size_t)T));
    CK(cudaMemcpyAsync(h->hmm_cumA.p, cumA.data(), cumA.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->hmm_cumpi.p, cumpi.data(), cumpi.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->hmm_logF.p, logF, (size_t)T * K * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->hmm_expF.p, expF.data(), (size_t)T * K * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->hmm_hasf.p, h->has_factor.data(), (size_t)T, cudaMemcpyHostToDevice, h->stream));
    {   // probabilities and their logs for Enumerate (score(e, val) = logpdf(Categorical), src/Stochy.jl:8)
        std::vector<double> p0((size_t)K), lA((size_t)K * K), lp0((size_t)K);
        for (int k = 0; k < K; ++k) p0[k] = x0 >= 0 ? A[(size_t)x0 * K + k] : pi[k];
        for (size_t i = 0; i < (size_t)K * K; ++i) lA[i] = log(A[i]);
        for (int k = 0; k < K; ++k) lp0[k] = log(p0[k]);
        CK(h->hmm_A.ensure((size_t)K * K * 8)); CK(h->hmm_p0.ensure((size_t)K * 8));
        CK(h->hmm_logA.ensure((size_t)K * K * 8)); CK(h->hmm_logp0.ensure((size_t)K * 8));
        CK(cudaMemcpyAsync(h->hmm_A.p, A, (size_t)K * K * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->hmm_p0.p, p0.data(), (size_t)K * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->hmm_logA.p, lA.data(), (size_t)K * K * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->hmm_logp0.p, lp0.data(), (size_t)K * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));                      // the staging vectors go out of scope
    }
    CK(cudaStreamSynchronize(h->stream));
    h->model = MODEL_HMM; h->K = K; h->T = T; h->x0 = x0;
    h->have_retained = false;
    invalidate_sweep(h);
    return SB_OK;
}

extern "C" int32_t sb_model_linear_gaussian(sb_handle* h, int32_t T, double a, double q, double c, double r,
                                            double m0, double P0, const double* y) {
    if (!h) return SB_EINVAL;
    if (!y || T < 1 || !(q > 0) || !(r > 0) || !(P0 > 0)) return fail(h, SB_EINVAL, "linear_gaussian: bad arguments");
    h->model = MODEL_LG; h->T = T; h->K = 0;
    h->lg_a = a; h->lg_q = q; h->lg_c = c; h->lg_r = r; h->lg_m0 = m0; h->lg_P0 = P0;
    h->lg_y.assign(y, y + T);
    h->have_retained = false;
    invalidate_sweep(h);
    return SB_OK;
}

extern "C" int32_t sb_model_ops(sb_handle* h, int32_t T, const sb_op_step* steps, const double* obs, int64_t n_obs) {
    if (!h) return SB_EINVAL;
    if (!steps || T < 1 || n_obs < 0 || (n_obs > 0 && !obs)) return fail(h, SB_EINVAL, "ops: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    int D = 1;
    for (int t = 0; t < T; ++t)
        if (steps[t].draw == SB_DRAW_DIRICHLET) {
            const sb_op_step& s = steps[t];
            if (s.par_n < 2 || s.par_n > SB_OPS_MAXD) return fail(h, SB_EUNSUPPORTED, "ops: a Dirichlet latent has 2..8 components on the GPU path (no CPU fallback)");
            if (s.par_off < 0 || (int64_t)s.par_off + s.par_n > n_obs) return fail(h, SB_EINVAL, "ops: Dirichlet parameters outside obs[]");
            if (D > 1 && s.par_n != D) return fail(h, SB_EINVAL, "ops: every Dirichlet statement of a program must have the same dimension");
            for (int d = 0; d < s.par_n; ++d) if (!(obs[s.par_off + d] > 0)) return fail(h, SB_EINVAL, "ops: distribution parameter out of range");
            D = s.par_n;
        }
    if (D > 1 && h->dist_ready) return fail(h, SB_EUNSUPPORTED, "ops: vector latents are not supported on sharded pools");
    for (int t = 0; t < T; ++t) {
        const sb_op_step& s = steps[t];
        if (s.draw < SB_DRAW_KEEP || s.draw > SB_DRAW_DIRICHLET || s.score < SB_SCORE_ZERO || s.score > SB_SCORE_CATEGORICAL)
            return fail(h, SB_EUNSUPPORTED, "ops: unknown draw / score statement (no CPU fallback)");
        if (D > 1 && ((s.draw != SB_DRAW_DIRICHLET && s.draw != SB_DRAW_KEEP) || (s.score != SB_SCORE_CATEGORICAL && s.score != SB_SCORE_ZERO)))
            return fail(h, SB_EUNSUPPORTED, "ops: a Dirichlet program draws / keeps the vector latent and observes Categorical(theta) only");
        if (D == 1 && s.score == SB_SCORE_CATEGORICAL)
            return fail(h, SB_EUNSUPPORTED, "ops: observe(Categorical(theta), ...) needs a Dirichlet latent");
        if (t == 0 && s.draw == SB_DRAW_KEEP) return fail(h, SB_EINVAL, "ops: the first step must draw the latent");
        if (s.obs_n < 0 || s.obs_off < 0 || (int64_t)s.obs_off + s.obs_n > n_obs) return fail(h, SB_EINVAL, "ops: observation range outside obs[]");
        bool ok = true;
        switch (s.draw) {
            case SB_DRAW_NORMAL:  ok = s.dp[2] > 0; break;
            case SB_DRAW_BETA:    ok = s.dp[0] > 0 && s.dp[1] > 0; break;
            case SB_DRAW_GAMMA:   ok = s.dp[0] > 0 && s.dp[1] > 0; break;
            case SB_DRAW_UNIFORM: ok = s.dp[1] > s.dp[0]; break;
            default: break;
        }
        if (s.score == SB_SCORE_NORMAL) ok = ok && s.sp[2] > 0;
        if (!ok) return fail(h, SB_EINVAL, "ops: distribution parameter out of range");
        for (int i = 0; i < s.obs_n; ++i) {
            const double y = obs[s.obs_off + i];
            if (std::isnan(y)) return fail(h, SB_ENAN, "ops: NaN observation");
            if (s.score == SB_SCORE_BERNOULLI && y != 0.0 && y != 1.0) return fail(h, SB_EINVAL, "ops: Bernoulli observations must be 0 or 1");
            if (s.score == SB_SCORE_POISSON && (y < 0.0 || y != floor(y))) return fail(h, SB_EINVAL, "ops: Poisson observations must be non-negative integers");
            if (s.score == SB_SCORE_CATEGORICAL && (y < 1.0 || y > (double)D || y != floor(y))) return fail(h, SB_EINVAL, "ops: Categorical observations must be integers in 1..D");
        }
    }
    std::vector<double> aux((size_t)(n_obs > 0 ? n_obs : 1), 0.0);
    for (int64_t i = 0; i < n_obs; ++i) aux[(size_t)i] = obs[i] >= 0.0 ? lgamma(obs[i] + 1.0) : 0.0;
    CK(h->ops_obs.ensure((size_t)(n_obs > 0 ? n_obs : 1) * 8));
    CK(h->ops_aux.ensure((size_t)(n_obs > 0 ? n_obs : 1) * 8));
    if (n_obs > 0) {
        CK(cudaMemcpyAsync(h->ops_obs.p, obs, (size_t)n_obs * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->ops_aux.p, aux.data(), (size_t)n_obs * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    h->ops.assign(steps, steps + T);
    h->ops_D = D;
    h->model = MODEL_OPS; h->T = T; h->K = 0;
    h->have_retained = false;
    invalidate_sweep(h);
    return SB_OK;
}

extern "C" int32_t sb_model_bayesnet(sb_handle* h, int32_t D, int32_t F, const uint32_t* site_parents,
                                     const double* site_cpt, const uint32_t* fac_parents, const double* fac_logf,
                                     int32_t cpt_stride, uint32_t query_mask) {
    if (!h) return SB_EINVAL;
    if (D < 1 || D > SB_MH_MAXD || F < 0 || !site_parents || !site_cpt || cpt_stride < 1 || (F > 0 && (!fac_parents || !fac_logf)))
        return fail(h, SB_EINVAL, "bayesnet: bad arguments (1 <= D <= 16)");
    if (query_mask == 0 || (query_mask >> D) != 0 || __builtin_popcount(query_mask) > 10)
        return fail(h, SB_EINVAL, "bayesnet: query_mask must select 1..10 of the D sites");
    for (int d = 0; d < D; ++d) {
        if (site_parents[d] >> d) return fail(h, SB_EUNSUPPORTED, "bayesnet: sites must be in topological order (fixed structure)");
        if ((1 << __builtin_popcount(site_parents[d])) > cpt_stride) return fail(h, SB_EINVAL, "bayesnet: cpt_stride too small");
    }
    for (int f = 0; f < F; ++f) {
        if (fac_parents[f] >> D) return fail(h, SB_EINVAL, "bayesnet: factor parent outside the sites");
        if ((1 << __builtin_popcount(fac_parents[f])) > cpt_stride) return fail(h, SB_EINVAL, "bayesnet: cpt_stride too small");
    }
    if (D > SB_MH_TABD && (size_t)(3 * D + F) * cpt_stride * 8 > 190 * 1024)
        return fail(h, SB_EUNSUPPORTED, "bayesnet: conditional tables exceed the shared memory of one SM (no CPU fallback)");
    CK(cudaSetDevice(h->opt.device));
    const size_t ns = (size_t)D * cpt_stride, nf = (size_t)(F > 0 ? F : 1) * cpt_stride;
    std::vector<double> lp1(ns), lp0(ns), fac(nf, 0.0);
    for (size_t i = 0; i < ns; ++i) { lp1[i] = log(site_cpt[i]); lp0[i] = log(1.0 - site_cpt[i]); }   // logpdf(Bernoulli(p), x)
    for (size_t i = 0; i < (size_t)F * cpt_stride; ++i) fac[i] = fac_logf[i];
    std::vector<unsigned> fp((size_t)(F > 0 ? F : 1), 0u);
    for (int f = 0; f < F; ++f) fp[f] = fac_parents[f];
    CK(h->bn_sp.ensure((size_t)D * 4)); CK(h->bn_cpt.ensure(ns * 8)); CK(h->bn_lp1.ensure(ns * 8)); CK(h->bn_lp0.ensure(ns * 8));
    CK(h->bn_fp.ensure(fp.size() * 4)); CK(h->bn_fac.ensure(nf * 8));
    CK(cudaMemcpyAsync(h->bn_sp.p, site_parents, (size_t)D * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->bn_cpt.p, site_cpt, ns * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->bn_lp1.p, lp1.data(), ns * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->bn_lp0.p, lp0.data(), ns * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->bn_fp.p, fp.data(), fp.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->bn_fac.p, fac.data(), nf * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->model = MODEL_BNET; h->bn_D = D; h->bn_F = F; h->bn_stride = cpt_stride; h->bn_query = query_mask;
    invalidate_sweep(h);
    return SB_OK;
}

// ======================================================================= sweeps
namespace {

// A step kernel that pulls from the pool of the previous step follows that pool's tile scan in the stream:
// launched with programmatic stream serialisation its prologue overlaps the scan (see sb_grid_dep_wait).
template <typename Kern, typename Args>
int launch_step(sb_handle* h, Kern kern, unsigned grid, size_t smem, bool after_scan, const Args& a) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(SB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (after_scan && ((h->dist_ready ? h->pdl_dist : h->pdl) & 2)) ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, a));
    return SB_OK;
}

template <bool FP32, bool MULTI>
int launch_step_hmm_m(sb_handle* h, int anc_mode, const SbHmmStepArgs& a) {
    const size_t smem = ((size_t)a.rows << SB_LG_BUCKETS) * 4 + (size_t)a.K * 8 + (size_t)(a.rows * a.K + a.K + a.K * a.K) * 4;
    int cap = 0, rc;
#define SB_HMM_LAUNCH(ANC, PEERS, CARRY)                                                       \
    do {                                                                                       \
        auto kern = k_step_hmm<ANC, FP32, PEERS, MULTI, CARRY>;                                \
        if ((rc = resident_ctas(h, kern, smem, &cap)) != SB_OK) return rc;                     \
        if ((rc = launch_step(h, kern, (unsigned)(a.ntiles < cap ? a.ntiles : cap), smem, ANC == SB_ANC_PULL, a)) != SB_OK) return rc; \
    } while (0)
    const bool carry = a.xb > 0;
    if (h->dist_ready) {
        if (anc_mode == SB_ANC_NONE) SB_HMM_LAUNCH(SB_ANC_NONE, false, false);
        else if (anc_mode == SB_ANC_PULL) { if (carry) SB_HMM_LAUNCH(SB_ANC_PULL, true, true); else SB_HMM_LAUNCH(SB_ANC_PULL, true, false); }
        else return fail(h, SB_EUNSUPPORTED, "sharded pools resample systematically at every step");
    } else {
        switch (anc_mode) {
            case SB_ANC_NONE:     SB_HMM_LAUNCH(SB_ANC_NONE, false, false); break;
            case SB_ANC_PULL:     if (carry) SB_HMM_LAUNCH(SB_ANC_PULL, false, true); else SB_HMM_LAUNCH(SB_ANC_PULL, false, false); break;
            case SB_ANC_EXPLICIT: SB_HMM_LAUNCH(SB_ANC_EXPLICIT, false, false); break;
            default:              SB_HMM_LAUNCH(SB_ANC_IDENTITY, false, false); break;
        }
    }
#undef SB_HMM_LAUNCH
    LAUNCH_CHECK();
    return SB_OK;
}
int launch_step_hmm(sb_handle* h, int anc_mode, const SbHmmStepArgs& a) {
    const bool fp32 = h->opt.precision == SB_FP32;
    if (h->hmm_multi) return fp32 ? launch_step_hmm_m<true, true>(h, anc_mode, a) : launch_step_hmm_m<false, true>(h, anc_mode, a);
    return fp32 ? launch_step_hmm_m<true, false>(h, anc_mode, a) : launch_step_hmm_m<false, false>(h, anc_mode, a);
}

template <bool FP32, bool OPS = false>
int launch_step_lg(sb_handle* h, int anc_mode, const SbLgStepArgs& a) {
    int cap = 0, rc;
#define SB_LG_LAUNCH(ANC, PEERS)                                                               \
    do {                                                                                       \
        auto kern = k_step_lg<ANC, FP32, PEERS, OPS>;                                          \
        if ((rc = resident_ctas(h, kern, 0, &cap)) != SB_OK) return rc;                        \
        if ((rc = launch_step(h, kern, (unsigned)(a.ntiles < cap ? a.ntiles : cap), 0, ANC == SB_ANC_PULL, a)) != SB_OK) return rc; \
    } while (0)
    if (h->dist_ready) {
        if (anc_mode == SB_ANC_NONE) SB_LG_LAUNCH(SB_ANC_NONE, false);
        else if (anc_mode == SB_ANC_PULL) SB_LG_LAUNCH(SB_ANC_PULL, true);
        else return fail(h, SB_EUNSUPPORTED, "sharded pools resample systematically at every step");
    } else {
        switch (anc_mode) {
            case SB_ANC_NONE:     SB_LG_LAUNCH(SB_ANC_NONE, false); break;
            case SB_ANC_PULL:     SB_LG_LAUNCH(SB_ANC_PULL, false); break;
            default:              SB_LG_LAUNCH(SB_ANC_EXPLICIT, false); break;
        }
    }
#undef SB_LG_LAUNCH
    LAUNCH_CHECK();
    return SB_OK;
}

// resample pool `slot` of step t explicitly (multinomial / literal modes) into d_anc
int explicit_resample(sb_handle* h, int slot, int t, unsigned iter, int64_t n_pool, int64_t N, int* d_anc) {
    SbStepInfo* info = h->d_info.as<SbStepInfo>() + t;
    if (h->opt.resample_mode == SB_RESAMPLE_MULTINOMIAL)
        return launch_grid_multinomial(h, slot, info, n_pool, N, nullptr, (unsigned)t, iter, d_anc);
    return launch_literal(h, h->wlit.as<double>(), n_pool, N, nullptr, (unsigned)t, iter, d_anc);
}

// systematic offsets of every step of this sweep (Philox stream, purpose RESAMPLE, counter (0, t, iter));
// also clears the per-step records so steps without a factor statement read as invalid
int upload_offsets(sb_handle* h, unsigned iter) {
    const int T = h->T;
    CK(h->d_r53.ensure(sizeof(unsigned long long) * (size_t)T));
    CK(h->d_info.ensure(sizeof(SbStepInfo) * (size_t)T));
    h->r53_host.resize((size_t)T);
    for (int t = 0; t < T; ++t) h->r53_host[t] = host_r53(h->opt.seed, 0, (uint32_t)t, iter, SB_P_RESAMPLE);
    CK(cudaMemcpyAsync(h->d_r53.p, h->r53_host.data(), sizeof(unsigned long long) * (size_t)T, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_info.p, 0, sizeof(SbStepInfo) * (size_t)T, h->stream));
    return SB_OK;
}

// The whole sweep as ONE cooperative launch (sb_resident.cuh) when every tile of the pool fits on the GPU at once.
// *done = 0: not eligible here (too many tiles for the resident grid, no cooperative launch): the caller runs the
// tiled kernels.  Buffers, offsets and the per-step records were prepared by sweep_hmm.
template <bool FP32, bool MULTI, bool CARRY, bool TRACE = false>
int sweep_hmm_resident_k(sb_handle* h, const SbResArgs& a, int* done) {
    if (!TRACE && a.dbg) return sweep_hmm_resident_k<FP32, MULTI, CARRY, true>(h, a, done);   // SB_RES_TRACE=1: the instantiation with the stamps
    auto kern = k_sweep_hmm<FP32, MULTI, CARRY, TRACE>;
    const size_t smem = sb_res_smem_bytes(a.K, a.ntiles, MULTI);
    if (smem > 160 * 1024) return SB_OK;
    int cap = 0, rc;
    if ((rc = resident_ctas(h, kern, smem, &cap)) != SB_OK) { h->err.clear(); return SB_OK; }
    if (getenv("SB_DEBUG")) fprintf(stderr, "[sb] resident sweep: tiles %d, cap %d, dynamic smem %zu\n", a.ntiles, cap, smem);
    if (a.ntiles > cap) return SB_OK;
    void* args[] = {(void*)&a};
    // cooperative launch: the CTAs exchange records among themselves, so all of them must be resident at once
    CK(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)a.ntiles), dim3(SB_THREADS), args, smem, h->stream));
    ++h->launches;
    h->res_epoch += (unsigned)a.T + 1u;
    h->bar_count += (unsigned)a.T * (unsigned)a.ntiles;            // one barrier per step, one arrival per CTA
    *done = 1;
    if (a.dbg) {                                                    // diagnostics: phase stamps of CTA 0, steps 8..15 (ns)
        std::vector<unsigned long long> ts(1024 + 4096);
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(ts.data(), a.dbg, ts.size() * 8, cudaMemcpyDeviceToHost));
        if (a.T > 12 && a.ntiles <= 1024) {
            // when each CTA published step 11, relative to the first: distribution, and the mean per number of CTAs on its SM
            std::vector<long long> d; std::map<unsigned, int> per_sm;
            unsigned long long t0 = ~0ull;
            for (int i = 0; i < a.ntiles; ++i) { t0 = std::min(t0, ts[1024 + 2048 + 2 * i]); per_sm[(unsigned)ts[1024 + 2048 + 2 * i + 1]]++; }
            double sum[8] = {0}; int cnt[8] = {0};
            for (int i = 0; i < a.ntiles; ++i) {
                const long long v = (long long)(ts[1024 + 2048 + 2 * i] - t0);
                d.push_back(v);
                const int c = std::min(7, per_sm[(unsigned)ts[1024 + 2048 + 2 * i + 1]]);
                sum[c] += (double)v; cnt[c]++;
            }
            std::sort(d.begin(), d.end());
            fprintf(stderr, "[sb trace] publish times of step 11 over %d CTAs (ns after the first): p10 %lld p50 %lld p90 %lld p99 %lld max %lld; step-to-step of CTA0 %lld\n",
                    a.ntiles, d[d.size() / 10], d[d.size() / 2], d[d.size() * 9 / 10], d[d.size() * 99 / 100], d.back(),
                    (long long)(ts[1024 + 2048] - ts[1024]));
            for (int c = 1; c < 8; ++c) if (cnt[c]) fprintf(stderr, "[sb trace]   SMs holding %d CTAs: %d CTAs, mean publish time %+.0f ns\n", c, cnt[c], sum[c] / cnt[c]);
            {   // the latest publishers: (tile, SM, CTAs on that SM, +ns)
                std::vector<std::pair<long long, int>> late;
                for (int i = 0; i < a.ntiles; ++i) late.push_back(std::make_pair((long long)(ts[1024 + 2048 + 2 * i] - t0), i));
                std::sort(late.begin(), late.end());
                fprintf(stderr, "[sb trace]   latest:");
                for (int q = 0; q < 24 && q < (int)late.size(); ++q) {
                    const int i = late[late.size() - 1 - q].second;
                    fprintf(stderr, " t%d/sm%u/%d/+%lld", i, (unsigned)ts[1024 + 2048 + 2 * i + 1], per_sm[(unsigned)ts[1024 + 2048 + 2 * i + 1]], late[late.size() - 1 - q].first);
                }
                fprintf(stderr, "\n[sb trace]   earliest:");
                for (int q = 0; q < 12 && q < (int)late.size(); ++q) {
                    const int i = late[q].second;
                    fprintf(stderr, " t%d/sm%u/%d/+%lld", i, (unsigned)ts[1024 + 2048 + 2 * i + 1], per_sm[(unsigned)ts[1024 + 2048 + 2 * i + 1]], late[q].first);
                }
                fprintf(stderr, "\n");
            }
        }
        for (int t = 8; t < 16 && t + 1 < a.T; ++t) {
            fprintf(stderr, "[sb trace] t=%d:", t);
            for (int i = 1; i <= 7; ++i) fprintf(stderr, " p%d %+lld", i, (long long)(ts[t * 16 + i] - ts[t * 16 + i - 1]));
            fprintf(stderr, "  | step %lld ns\n", (long long)(ts[(t + 1) * 16] - ts[t * 16]));
        }
    }
    return SB_OK;
}
int sweep_hmm_resident(sb_handle* h, int64_t N, unsigned iter, bool conditional, bool history, bool final_anc, int64_t S, int* done) {
    *done = 0;
    if (h->coop < 0) {
        int v = 0;
        h->coop = (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->opt.device) == cudaSuccess && v) ? 1 : 0;
    }
    if (!h->coop) return SB_OK;
    const int n_out = (int)(N + (conditional ? 1 : 0));
    const int nt = (int)tiles_of(n_out);
    if (nt > SB_RES_MAXPER * SB_THREADS) return SB_OK;
    SbResArgs a;
    memset(&a, 0, sizeof(a));
    const int G = 1 << SB_LG_BUCKETS;
    a.tabA = h->hmm_tabA.as<uint32_t>(); a.thrA = h->hmm_cumA.as<uint32_t>();
    if (h->x0 >= 0) { a.tab0 = a.tabA + (size_t)h->x0 * G; a.thr0 = a.thrA + (size_t)h->x0 * h->K; }
    else { a.tab0 = h->hmm_tabpi.as<uint32_t>(); a.thr0 = h->hmm_cumpi.as<uint32_t>(); }
    a.logF = h->hmm_logF.as<double>();
    a.K = h->K; a.T = h->T;
    for (int sl = 0; sl < 2; ++sl) { a.w32[sl] = W32(h, sl); a.e[sl] = TE(h, sl); a.Q[sl] = TQ(h, sl); a.wp[sl] = WP(h, sl); }
    a.xbuf = h->xbuf.as<int>(); a.ancbuf = h->ancbuf.as<int>(); a.S = S;
    a.history = history ? 1 : 0; a.final_anc = final_anc ? 1 : 0;
    a.xstar = conditional ? h->xstar[h->xstar_cur].as<int>() : nullptr;
    a.N = N; a.n_out = n_out; a.ntiles = nt;
    a.r53 = h->d_r53.as<unsigned long long>(); a.info = h->d_info.as<SbStepInfo>();
    a.iter = iter; a.rk = round_keys(h->opt.seed); a.err = h->d_err;
    if (getenv("SB_RES_TRACE")) { CK(h->res_trace.ensure((1024 + 4096) * 8)); a.dbg = h->res_trace.as<unsigned long long>(); }
    // the records the CTAs exchange (zeroed once: epoch 0 is never waited for)
    const size_t rec_bytes = (size_t)2 * SB_RES_MAXPER * SB_THREADS * sizeof(ulonglong2);
    if (h->res_rec.cap < rec_bytes) {
        CK(h->res_rec.ensure(rec_bytes));
        CK(cudaMemsetAsync(h->res_rec.p, 0, h->res_rec.cap, h->stream));
    }
    a.rec[0] = h->res_rec.as<ulonglong2>(); a.rec[1] = a.rec[0] + SB_RES_MAXPER * SB_THREADS;
    a.epoch0 = h->res_epoch;
    a.bar = grid_bar(h); a.bar0 = h->bar_count;
    a.xb = 0;
    if (h->use_carry && h->K <= 64) a.xb = h->K <= 16 ? 4 : 6;      // n_out <= 2^21 here: (parent << xb) | state fits 31 bits
    const bool fp32 = h->opt.precision == SB_FP32;
#define SB_RES_GO(F, M) (a.xb ? sweep_hmm_resident_k<F, M, true>(h, a, done) : sweep_hmm_resident_k<F, M, false>(h, a, done))
    if (h->hmm_multi) return fp32 ? SB_RES_GO(true, true) : SB_RES_GO(false, true);
    return fp32 ? SB_RES_GO(true, false) : SB_RES_GO(false, false);
#undef SB_RES_GO
}

// One sweep of the discrete model.  conditional: pool of N+1 with the retained particle appended.
// Sharded handle: N is the GLOBAL particle count (world x N_local); this GPU produces its own share.
int sweep_hmm(sb_handle* h, int64_t N, unsigned iter, bool conditional, bool history, bool final_anc) {
    const int T = h->T, K = h->K;
    const bool dist = h->dist_ready;
    if (dist && (h->opt.store_logw || h->opt.resample_mode != SB_RESAMPLE_SYSTEMATIC))
        return fail(h, SB_EUNSUPPORTED, "sharded pools resample systematically and keep no log-weight history");
    // sharded conditional SMC: the history rows live in the arena (the parents of step t are read from row t-1 of the
    // owning GPU), reserved before the arena was exported
    const bool dhist = dist && (conditional || history || final_anc);
    if (dhist && h->hist_rows < T)
        return fail(h, SB_EUNSUPPORTED, "sharded conditional SMC / history: call sb_dist_reserve_history(h, T) before sb_dist_ipc_export");
    if (dhist) history = true;
    const int64_t S = dist ? h->NL : tiles_of(N + 1) * SB_TILE;        // row stride (slot N = retained)
    const int n_out = (int)(N + (conditional ? 1 : 0));
    const int mode = h->opt.resample_mode;
    const bool literal = mode == SB_RESAMPLE_LITERAL;
    int rc;
    if (dist && n_out > h->NL * h->world) return fail(h, SB_EINVAL, "sharded handle: the pool (particles + the retained one) exceeds world x N_local entries");
    if ((rc = ensure_grid(h, dist ? h->NL * h->world : N + 1, N)) != SB_OK) return rc;
    const int xrows = history ? T : 2;
    const int arows = history ? T : 1;
    if (!dist) {
        CK(h->xbuf.ensure((size_t)xrows * S * 4));
        CK(h->ancbuf.ensure((size_t)arows * S * 4));
    } else if (dhist) {
        CK(h->ancbuf.ensure((size_t)arows * S * 4));                   // local rows of GLOBAL parent indices
    }
    CK(h->d_info.ensure(sizeof(SbStepInfo) * (size_t)T));
    if (h->opt.store_logw) CK(h->lwbuf.ensure((size_t)T * S * 8));
    if (literal) CK(h->wlit.ensure((size_t)S * 8));
    if ((rc = upload_offsets(h, iter)) != SB_OK) return rc;
    h->sw_N = dist ? h->NL : N; h->sw_S = S; h->sw_hist = history ? 1 : 0; h->sw_cond = conditional ? 1 : 0;
    int* xb = h->xbuf.as<int>();
    int* ab = h->ancbuf.as<int>();
    const unsigned sb0 = dist ? h->slot_base : 0u;
    auto xoff = [&](int t) -> size_t { return dhist ? h->off_xh + (size_t)t * (size_t)S * 4 : h->off_x[(sb0 + (unsigned)t) & 1u]; };   // sharded: inside the arena
    auto xrow = [&](int t) -> int* {
        if (dist) return reinterpret_cast<int*>((char*)h->arena + xoff(t));
        return xb + (size_t)(history ? t : (t & 1)) * S;
    };
    const int ntiles = dist ? (int)(h->NL / SB_TILE) : (int)tiles_of(n_out);
    h->sw_T = T; h->sw_slot = (int)((sb0 + (unsigned)(T - 1)) & 1u);
    if (dist) h->slot_base += (unsigned)T;
    h->last_route = 0;
    if (!dist && h->use_resident && mode == SB_RESAMPLE_SYSTEMATIC && !h->opt.store_logw && !h->profile) {
        bool all_factors = true;
        for (int t = 0; t < T; ++t) all_factors = all_factors && h->has_factor[t];
        if (all_factors) {
            int done = 0;
            if ((rc = sweep_hmm_resident(h, N, iter, conditional, history, final_anc, S, &done)) != SB_OK) return rc;
            if (done) { h->last_route = 1; return SB_OK; }
        }
    }
    for (int t = 0; t < T; ++t) {
        const int cur = (int)((sb0 + (unsigned)t) & 1u), prev = cur ^ 1;
        int anc_mode = SB_ANC_NONE;
        int* anc_row_prev = (t && (!dist || dhist)) ? ab + (size_t)(history ? (t - 1) : 0) * S : nullptr;
        if (t > 0) {
            if (!h->has_factor[t - 1]) anc_mode = SB_ANC_IDENTITY;
            else if (mode == SB_RESAMPLE_SYSTEMATIC) anc_mode = SB_ANC_PULL;
            else {
                anc_mode = SB_ANC_EXPLICIT;
                const int64_t n_pool = N + (conditional ? 1 : 0);
                if ((rc = explicit_resample(h, prev, t - 1, iter, n_pool, N, anc_row_prev)) != SB_OK) return rc;
            }
        }
        if (dist && anc_mode == SB_ANC_IDENTITY)
            return fail(h, SB_EUNSUPPORTED, "sharded pools need a factor statement at every step");
        SbHmmStepArgs a;
        memset(&a, 0, sizeof(a));
        a.pool = make_pool(h, prev, t ? xrow(t - 1) : nullptr, (dhist && t) ? xoff(t - 1) : 0);
        a.e_prev = TE_pull(h, prev);
        a.Pt = h->Pt.as<unsigned long long>();
        a.child_start = h->child_start.as<int>();
        a.first_tile = h->first_tile.as<int>();
        a.info_prev = t ? h->d_info.as<SbStepInfo>() + (t - 1) : nullptr;
        a.anc_in = anc_row_prev;
        a.anc_out = (history && t) ? anc_row_prev : nullptr;
        a.x_out = xrow(t);
        a.w32_out = W32(h, cur);
        a.rec = make_rec(h, cur);
        // a step without a factor statement is followed by no tile scan: its tile exponents must not enter the running
        // maximum that the NEXT scan consumes (a stale larger exponent shifts that pool's sums right: beyond the scan's
        // head-room of LS bits the sums lose bits, and 2^-60 of them is "zero probability mass")
        if (!h->has_factor[t]) a.rec.emax = nullptr;
        a.sig = make_signal(h);
        attach_summary(h, a.sig, a.rec);
        a.wp_out = WP(h, cur);
        a.lw_out = h->opt.store_logw ? h->lwbuf.as<double>() + (size_t)t * S : nullptr;
        a.wlit_out = literal ? h->wlit.as<double>() : nullptr;
        const int G = 1 << SB_LG_BUCKETS;
        if (t == 0 && h->x0 >= 0) { a.thr = h->hmm_cumA.as<uint32_t>() + (size_t)h->x0 * K; a.tab = h->hmm_tabA.as<uint32_t>() + (size_t)h->x0 * G; a.rows = 1; }
        else if (t == 0) { a.thr = h->hmm_cumpi.as<uint32_t>(); a.tab = h->hmm_tabpi.as<uint32_t>(); a.rows = 1; }
        else { a.thr = h->hmm_cumA.as<uint32_t>(); a.tab = h->hmm_tabA.as<uint32_t>(); a.rows = K; }
        a.logF_t = h->hmm_logF.as<double>() + (size_t)t * K;
        a.expF_t = h->hmm_expF.as<double>() + (size_t)t * K;
        a.xstar = conditional ? (dist ? reinterpret_cast<const int*>((char*)h->arena + h->off_xstar[h->xstar_cur]) : h->xstar[h->xstar_cur].as<int>()) : nullptr;
        a.N = N; a.n_out = n_out; a.ntiles = ntiles; a.nt_pool = dist ? ntiles * h->world : ntiles; a.K = K; a.t = t; a.iter = iter; a.seed = h->opt.seed; a.rk = round_keys(h->opt.seed); a.err = h->d_err;
        // head markers that carry the parent's state ((parent << xb) | state must stay below 2^31): no gather
        a.xb = 0;
        if (anc_mode == SB_ANC_PULL && h->use_carry) {
            const int xb = K <= 16 ? 4 : 6;
            if (K <= 64 && ((tiles_of(n_out) * SB_TILE) << xb) <= (1ll << 31)) a.xb = xb;
        }
        {
            ProfScope ps(h, t ? t : -1);                               // t = 0 has no resample/gather: not sampled
            rc = launch_step_hmm(h, anc_mode, a);
        }
        if (rc != SB_OK) return rc;
        if (h->has_factor[t]) {
            // literal mode still runs the tile scan: it supplies the log-evidence term
            if ((rc = launch_tilescan(h, cur, n_out, N, h->d_r53.as<unsigned long long>() + t,
                                      h->d_info.as<SbStepInfo>() + t)) != SB_OK) return rc;
        }
    }
    if (final_anc) {
        const int t = T - 1, cur = (int)((sb0 + (unsigned)t) & 1u);
        int* row = ab + (size_t)(history ? t : 0) * S;
        if (!h->has_factor[t]) {
            if (dist) return fail(h, SB_EUNSUPPORTED, "sharded pools need a factor statement at every step");
            k_fill_identity<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(row, N);
            LAUNCH_CHECK();
        } else if (mode == SB_RESAMPLE_SYSTEMATIC) {
            if ((rc = launch_pull(h, cur, h->d_info.as<SbStepInfo>() + t, dist ? h->NL : N, row)) != SB_OK) return rc;
        } else {
            if ((rc = explicit_resample(h, cur, t, iter, n_out, N, row)) != SB_OK) return rc;
        }
    }
    return SB_OK;
}

// Sharded handle: N is the GLOBAL particle count.
int sweep_lg(sb_handle* h, int64_t N, unsigned iter) {
    const int T = h->T;
    const bool ops = h->model == MODEL_OPS;                         // op-list program: same sweep, fp64, its own propagate + weight
    const bool fp32 = h->opt.precision == SB_FP32 && !ops;
    const bool dist = h->dist_ready;
    const int XD = ops ? h->ops_D : 1;                              // planes of the latent (Dirichlet programs: D)
    const size_t sz = (fp32 ? 4 : 8) * (size_t)XD;                  // bytes per particle of one pool slot
    const int64_t S = dist ? h->NL : tiles_of(N) * SB_TILE;
    const int mode = h->opt.resample_mode;
    if (XD > 1 && dist) return fail(h, SB_EUNSUPPORTED, "ops: vector latents are not supported on sharded pools");
    if (mode == SB_RESAMPLE_LITERAL)
        return fail(h, SB_EUNSUPPORTED, "literal resampling is a parity mode of the discrete model only");
    if (dist && (mode != SB_RESAMPLE_SYSTEMATIC || h->opt.store_logw))
        return fail(h, SB_EUNSUPPORTED, "sharded pools run plain SMC sweeps with systematic resampling");
    int rc;
    if ((rc = ensure_grid(h, N, N)) != SB_OK) return rc;
    if (!dist) {
        CK(h->xbuf.ensure((size_t)2 * S * sz));
        CK(h->ancbuf.ensure((size_t)S * 4));
    }
    CK(h->d_info.ensure(sizeof(SbStepInfo) * (size_t)T));
    if (h->opt.store_logw) CK(h->lwbuf.ensure((size_t)T * S * 8));
    if ((rc = upload_offsets(h, iter)) != SB_OK) return rc;
    h->sw_N = dist ? h->NL : N; h->sw_S = S; h->sw_hist = 0; h->sw_cond = 0;
    char* xb = h->xbuf.as<char>();
    auto xrow = [&](int slot) -> char* { return dist ? (char*)h->arena + h->off_x[slot] : xb + (size_t)slot * S * sz; };
    const int ntiles = (int)(S / SB_TILE);
    const unsigned sb0 = dist ? h->slot_base : 0u;
    h->sw_T = T; h->sw_slot = (int)((sb0 + (unsigned)(T - 1)) & 1u);
    if (dist) h->slot_base += (unsigned)T;
    for (int t = 0; t < T; ++t) {
        const int cur = (int)((sb0 + (unsigned)t) & 1u), prev = cur ^ 1;
        int anc_mode = t ? (mode == SB_RESAMPLE_SYSTEMATIC ? SB_ANC_PULL : SB_ANC_EXPLICIT) : SB_ANC_NONE;
        if (anc_mode == SB_ANC_EXPLICIT)
            if ((rc = explicit_resample(h, prev, t - 1, iter, N, N, h->ancbuf.as<int>())) != SB_OK) return rc;
        SbLgStepArgs a;
        memset(&a, 0, sizeof(a));
        a.pool = make_pool(h, prev, xrow(prev));
        a.e_prev = TE_pull(h, prev);
        a.Pt = h->Pt.as<unsigned long long>();
        a.child_start = h->child_start.as<int>();
        a.first_tile = h->first_tile.as<int>();
        a.info_prev = t ? h->d_info.as<SbStepInfo>() + (t - 1) : nullptr;
        a.anc_in = h->ancbuf.as<int>();
        a.x_out = xrow(cur);
        a.w32_out = W32(h, cur);
        a.rec = make_rec(h, cur);
        a.sig = make_signal(h);
        attach_summary(h, a.sig, a.rec);
        a.wp_out = WP(h, cur);
        a.lw_out = h->opt.store_logw ? h->lwbuf.as<double>() + (size_t)t * S : nullptr;
        a.N = N; a.ntiles = ntiles; a.nt_pool = dist ? ntiles * h->world : ntiles; a.t = t; a.iter = iter; a.seed = h->opt.seed; a.rk = round_keys(h->opt.seed);
        a.xdim = XD; a.xplane = S;
        if (ops) {
            a.op = h->ops[(size_t)t]; a.obs = h->ops_obs.as<double>(); a.obs_aux = h->ops_aux.as<double>();
        } else {
            a.a = h->lg_a; a.sq = sqrt(h->lg_q); a.sp = sqrt(h->lg_P0); a.c = h->lg_c; a.r = h->lg_r; a.m0 = h->lg_m0;
            a.y = h->lg_y[t]; a.lc = -0.5 * log(SB_TWO_PI * h->lg_r);
            a.hr = -0.5 / a.r;
            a.fa = (float)a.a; a.fsq = (float)a.sq; a.fsp = (float)a.sp; a.fc = (float)a.c; a.fy = (float)a.y;
            a.fhr = -0.5f / (float)a.r; a.flc = (float)a.lc; a.fm0 = (float)a.m0;
        }
        a.err = h->d_err;
        {
            ProfScope ps(h, t ? t : -1);
            rc = ops ? launch_step_lg<false, true>(h, anc_mode, a) : (fp32 ? launch_step_lg<true>(h, anc_mode, a) : launch_step_lg<false>(h, anc_mode, a));
        }
        if (rc != SB_OK) return rc;
        if ((rc = launch_tilescan(h, cur, N, N, h->d_r53.as<unsigned long long>() + t,
                                  h->d_info.as<SbStepInfo>() + t)) != SB_OK) return rc;
    }
    return SB_OK;
}

// log-evidence of the last sweep: sum over its factor steps of log(mean weight), accumulated on the
// host in step order from the tile-scan records (integers), so it is reproducible bit for bit
int fetch_logZ(sb_handle* h, double* out) {
    const int T = h->T;
    h->info_host.resize((size_t)T);
    CK(cudaMemcpyAsync(h->info_host.data(), h->d_info.p, sizeof(SbStepInfo) * (size_t)T, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double v = 0.0;
    for (int t = 0; t < T; ++t)
        if (h->info_host[t].valid) v += log_mean_weight(h->info_host[t]);
    *out = v;
    return SB_OK;
}

}  // namespace

extern "C" int32_t sb_smc_sweep(sb_handle* h, int64_t N, uint32_t iter, int32_t store_history, double* out_logZ) {
    if (!h) return SB_EINVAL;
    if (N < 1) return fail(h, SB_EINVAL, "smc: N must be >= 1");
    if (h->model != MODEL_HMM && h->model != MODEL_LG && h->model != MODEL_OPS)
        return fail(h, SB_EUNSUPPORTED, "smc: no fixed-structure state-space model is loaded (no CPU fallback)");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    if (h->dist_ready && N != h->NL) return fail(h, SB_EINVAL, "smc: a sharded handle sweeps exactly N_local particles per GPU");
    const int64_t Ng = h->dist_ready ? N * h->world : N;            // the sweep is global over all GPUs
    if (h->model == MODEL_HMM) rc = sweep_hmm(h, Ng, iter, false, store_history != 0, store_history != 0);
    else rc = sweep_lg(h, Ng, iter);
    if (rc != SB_OK) return rc;
    if ((rc = end_timed(h)) != SB_OK) return rc;
    h->stats.particle_steps = N * (int64_t)h->T;
    double lz = 0.0;
    if ((rc = fetch_logZ(h, &lz)) != SB_OK) return rc;
    h->stats.logZ_first = h->stats.logZ_last = lz;
    if (out_logZ) *out_logZ = lz;
    return check_err_word(h);
}

// `chains` independent PMCMC chains of numparticles each, one CTA per chain, the whole iteration loop on the device
extern "C" int32_t sb_pmcmc_chains(sb_handle* h, int64_t chains, int64_t iters, int64_t N, int64_t chain0,
                                   int64_t* marg, int64_t* joint) {
    if (!h) return SB_EINVAL;
    if (h->model != MODEL_HMM)
        return fail(h, SB_EUNSUPPORTED, "pmcmc: only the fixed-structure discrete model is supported (no CPU fallback)");
    if (chains < 1 || iters < 1 || N < 1 || chain0 < 0) return fail(h, SB_EINVAL, "pmcmc_chains: bad arguments");
    if (N + 1 > SB_TILE) return fail(h, SB_EUNSUPPORTED, "pmcmc_chains: a chain's pool (numparticles + 1) must fit one tile of 2048; use sb_pmcmc_run");
    if (h->dist_ready) return fail(h, SB_EUNSUPPORTED, "pmcmc_chains: chain sets run one handle per GPU, not on a sharded handle");
    CK(cudaSetDevice(h->opt.device));
    const int T = h->T, K = h->K;
    int64_t njoint = 0;
    if (joint) {
        if (T * log2((double)K) > 26.0) return fail(h, SB_EUNSUPPORTED, "pmcmc: joint histogram needs K^T <= 2^26; pass joint = NULL");
        njoint = 1; for (int t = 0; t < T; ++t) njoint *= K;
    }
    const int S = (int)((N + 1 + 7) & ~7LL);
    invalidate_sweep(h);                                            // the chain kernel keeps its history in its own buffers
    CK(h->sm_x.ensure((size_t)chains * T * S * 4));
    CK(h->sm_a.ensure((size_t)chains * T * S * 4));
    CK(h->sm_xs.ensure((size_t)chains * 2 * T * 4));
    CK(h->sm_lz.ensure((size_t)chains * 2 * 8));
    CK(h->marg.ensure((size_t)T * K * 8));
    CK(cudaMemsetAsync(h->marg.p, 0, (size_t)T * K * 8, h->stream));
    if (joint) { CK(h->joint.ensure((size_t)njoint * 8)); CK(cudaMemsetAsync(h->joint.p, 0, (size_t)njoint * 8, h->stream)); }
    int pick = h->opt.retained_pick;
    if (pick < 0) pick = h->opt.resample_mode == SB_RESAMPLE_SYSTEMATIC ? 1 : 0;
    SbSmallArgs a;
    memset(&a, 0, sizeof(a));
    const int G = 1 << SB_LG_BUCKETS;
    a.K = K; a.T = T;
    a.tabA = h->hmm_tabA.as<uint32_t>(); a.thrA = h->hmm_cumA.as<uint32_t>();
    if (h->x0 >= 0) { a.tab0 = a.tabA + (size_t)h->x0 * G; a.thr0 = a.thrA + (size_t)h->x0 * K; }
    else { a.tab0 = h->hmm_tabpi.as<uint32_t>(); a.thr0 = h->hmm_cumpi.as<uint32_t>(); }
    a.logF = h->hmm_logF.as<double>(); a.expF = h->hmm_expF.as<double>(); a.has_factor = h->hmm_hasf.as<unsigned char>();
    a.N = (int)N; a.S = S; a.iters = iters; a.chain0 = chain0; a.mode = h->opt.resample_mode; a.pick = pick; a.seed = h->opt.seed;
    a.x_hist = h->sm_x.as<int>(); a.anc_hist = h->sm_a.as<int>(); a.xstar = h->sm_xs.as<int>();
    a.marg = h->marg.as<unsigned long long>(); a.joint = joint ? h->joint.as<unsigned long long>() : nullptr;
    a.logZ = h->sm_lz.as<double>(); a.err = h->d_err;
    // history and per-time counts stay on chip when they are small (they then cost no global round trips in the
    // lineage walks); counts: iters * N per cell must fit 32 bits
    a.marg_smem = (marg && (size_t)T * K <= 4096 && (double)iters * (double)N < 4.0e9) ? 1 : 0;
    a.hist_smem = ((size_t)2 * T * S * 4 <= 48 * 1024) ? 1 : 0;
    const size_t smem = sb_small_smem_bytes(K, T, S, a.hist_smem, a.marg_smem);
    if (smem > 200 * 1024) return fail(h, SB_EUNSUPPORTED, "pmcmc_chains: model tables do not fit in shared memory");
    int rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    const bool fp32 = h->opt.precision == SB_FP32;
    if (fp32) {
        CK(cudaFuncSetAttribute(k_pmcmc_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pmcmc_small<true><<<(unsigned)chains, SB_THREADS, smem, h->stream>>>(a);
    } else {
        CK(cudaFuncSetAttribute(k_pmcmc_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pmcmc_small<false><<<(unsigned)chains, SB_THREADS, smem, h->stream>>>(a);
    }
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    h->stats.particle_steps = chains * iters * N * (int64_t)T;
    double lz[2] = {0.0, 0.0};
    CK(cudaMemcpy(lz, h->sm_lz.p, sizeof(lz), cudaMemcpyDeviceToHost));      // chain chain0
    h->stats.logZ_first = lz[0]; h->stats.logZ_last = lz[1];
    h->small_cur = (int)(iters & 1);
    h->have_retained = true;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    if (marg) {
        std::vector<unsigned long long> tmp((size_t)T * K);
        CK(cudaMemcpy(tmp.data(), h->marg.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) marg[i] += (int64_t)tmp[i];
    }
    if (joint) {
        std::vector<unsigned long long> tmp((size_t)njoint);
        CK(cudaMemcpy(tmp.data(), h->joint.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) joint[i] += (int64_t)tmp[i];
    }
    return SB_OK;
}

namespace {
// PMCMC on a pool sharded over the GPUs (iterated conditional SMC, src/pmcmc.jl:78-99): every rank calls this with the
// same arguments.  N = world x N_local - 1 particles, so the pool of N + 1 entries fills the shards and the retained
// particle is the last entry of the last GPU.  marg receives this rank's share of the counts.
int pmcmc_run_dist(sb_handle* h, int64_t iters, int64_t N, int64_t* marg, int64_t* joint) {
    if (joint) return fail(h, SB_EUNSUPPORTED, "pmcmc on a sharded pool: the joint histogram is not kept (pass joint = NULL)");
    if (N + 1 != h->NL * h->world)
        return fail(h, SB_EINVAL, "pmcmc on a sharded pool: numparticles must be world x N_local - 1 (the retained particle takes the last entry)");
    if (h->opt.resample_mode != SB_RESAMPLE_SYSTEMATIC)
        return fail(h, SB_EUNSUPPORTED, "sharded pools resample systematically");
    CK(cudaSetDevice(h->opt.device));
    const int T = h->T, K = h->K;
    if (h->hist_rows < T)
        return fail(h, SB_EUNSUPPORTED, "pmcmc on a sharded pool: call sb_dist_reserve_history(h, T) before sb_dist_ipc_export");
    const int64_t S = h->NL;
    CK(h->marg.ensure((size_t)T * K * 8));
    CK(cudaMemsetAsync(h->marg.p, 0, (size_t)T * K * 8, h->stream));
    int pick = h->opt.retained_pick;
    if (pick < 0) pick = 1;                                           // systematic: uniform pick (as on one GPU)
    int rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    SbBackArgs b;
    memset(&b, 0, sizeof(b));
    for (int g = 0; g < h->world; ++g) b.base[g] = h->peer_base[g];
    b.off_cnt[0] = h->off_cnt[0]; b.off_cnt[1] = h->off_cnt[1]; b.off_lin = h->off_lin; b.off_flags = h->off_flags;
    b.world = h->world; b.self = h->rank; b.log2n = 0; while ((1ll << b.log2n) < h->NL) ++b.log2n;
    b.N = N; b.K = K; b.T = T; b.marg = h->marg.as<unsigned long long>(); b.err = h->d_err;
    const unsigned grid8 = (unsigned)((h->NL / 8 + 255) / 256), grid1 = (unsigned)((h->NL + 255) / 256);
    for (int64_t it = 0; it < iters; ++it) {
        if ((rc = sweep_hmm(h, N, (unsigned)it, it > 0, true, true)) != SB_OK) return rc;
        const int* ab = h->ancbuf.as<int>();
        b.n_pool = N + (it > 0 ? 1 : 0);
        b.off_xstar = h->off_xstar[h->xstar_cur ^ 1];
        // src/pmcmc.jl:89 -- the child whose lineage is retained (k_lineage_start's arithmetic, on the host)
        b.pick = 0;
        if (pick) {
            const double u = (double)host_r53(h->opt.seed, 0, (uint32_t)T, (uint32_t)it, SB_P_PICK) * 0x1p-53;
            long long c = (long long)(u * (double)N);
            b.pick = c >= N ? N - 1 : c;
        }
        // the forward sweep's last kernels (final pull) read the peers' pools, not their count buffers: no wait
        b.t = T - 1; b.x_t = nullptr; b.anc_prev = ab + (size_t)(T - 1) * S; b.identity_prev = 0;
        b.wait_epoch = 0u; b.sig_epoch = ++h->epoch;
        k_count_final_dist<<<grid1, 256, 0, h->stream>>>(b);
        LAUNCH_CHECK();
        for (int t = T - 1; t >= 0; --t) {
            b.t = t;
            b.x_t = reinterpret_cast<const int*>((char*)h->arena + h->off_xh + (size_t)t * (size_t)S * 4);
            b.anc_prev = t ? ab + (size_t)(t - 1) * S : nullptr;
            b.identity_prev = t ? !h->has_factor[t - 1] : 0;
            b.wait_epoch = h->epoch; b.sig_epoch = ++h->epoch;
            k_count_back_dist<<<grid8, 256, (size_t)K * 4, h->stream>>>(b);
            LAUNCH_CHECK();
        }
        b.wait_epoch = h->epoch;
        k_flag_wait<<<1, 32, 0, h->stream>>>(b);
        LAUNCH_CHECK();
        h->xstar_cur ^= 1;
        h->have_retained = true;
        if (it == 0) {
            double lz;
            if ((rc = fetch_logZ(h, &lz)) != SB_OK) return rc;
            h->stats.logZ_first = lz;
            if ((rc = check_err_word(h)) != SB_OK) return rc;
        }
    }
    if ((rc = end_timed(h)) != SB_OK) return rc;
    h->stats.particle_steps = iters * h->NL * (int64_t)T;
    double lz;
    if ((rc = fetch_logZ(h, &lz)) != SB_OK) return rc;
    h->stats.logZ_last = lz;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    if (marg) {
        std::vector<unsigned long long> tmp((size_t)T * K);
        CK(cudaMemcpy(tmp.data(), h->marg.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) marg[i] += (int64_t)tmp[i];
    }
    return SB_OK;
}
}  // namespace

extern "C" int32_t sb_pmcmc_run(sb_handle* h, int64_t iters, int64_t N, int64_t* marg, int64_t* joint) {
    if (!h) return SB_EINVAL;
    // a pool that fits one tile runs as a one-CTA chain: the whole iteration loop in one launch
    if (h->use_small && h->model == MODEL_HMM && !h->dist_ready && iters >= 1 && N >= 1 && N + 1 <= SB_TILE && !h->opt.store_logw)
        return sb_pmcmc_chains(h, 1, iters, N, 0, marg, joint);
    h->small_cur = -1;
    if (h->model != MODEL_HMM)
        return fail(h, SB_EUNSUPPORTED, "pmcmc: only the fixed-structure discrete model is supported (no CPU fallback)");
    if (iters < 1 || N < 1) return fail(h, SB_EINVAL, "pmcmc: numiterations and numparticles must be >= 1");
    if (h->dist_ready) return pmcmc_run_dist(h, iters, N, marg, joint);
    CK(cudaSetDevice(h->opt.device));
    const int T = h->T, K = h->K;
    int64_t njoint = 0;
    if (joint) {
        double lg = T * log2((double)K);
        if (lg > 26.0) return fail(h, SB_EUNSUPPORTED, "pmcmc: joint histogram needs K^T <= 2^26; pass joint = NULL");
        njoint = 1; for (int t = 0; t < T; ++t) njoint *= K;
    }
    const int64_t S = tiles_of(N + 1) * SB_TILE;
    CK(h->xstar[0].ensure((size_t)T * 4)); CK(h->xstar[1].ensure((size_t)T * 4));
    for (int i = 0; i < 3; ++i) CK(h->cnt[i].ensure((size_t)S * 4));
    CK(h->marg.ensure((size_t)T * K * 8));
    CK(cudaMemsetAsync(h->marg.p, 0, (size_t)T * K * 8, h->stream));
    if (joint) { CK(h->joint.ensure((size_t)njoint * 8)); CK(cudaMemsetAsync(h->joint.p, 0, (size_t)njoint * 8, h->stream)); }
    int pick = h->opt.retained_pick;
    if (pick < 0) pick = h->opt.resample_mode == SB_RESAMPLE_SYSTEMATIC ? 1 : 0;
    int rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    int* lineage = h->d_misc.as<int>();
    for (int64_t it = 0; it < iters; ++it) {
        if ((rc = sweep_hmm(h, N, (unsigned)it, it > 0, true, true)) != SB_OK) return rc;
        const int n_pool = (int)(N + (it > 0 ? 1 : 0));
        int* xb = h->xbuf.as<int>();
        int* ab = h->ancbuf.as<int>();
        // src/pmcmc.jl:89 -- the next retained particle
        k_lineage_start<<<1, 1, 0, h->stream>>>(ab + (size_t)(T - 1) * S, N, pick, h->opt.seed, (unsigned)T, (unsigned)it, lineage);
        LAUNCH_CHECK();
        // src/pmcmc.jl:90-92 -- count every particle's value
        // three count buffers in rotation: step t reads buf[t % 3], accumulates into buf[(t - 1) % 3] and clears
        // buf[(t + 1) % 3] (read by the previous launch) for the step after next -- no memset between the T launches
        int* cnt_t = h->cnt[(T - 1) % 3].as<int>();
        CK(cudaMemsetAsync(cnt_t, 0, (size_t)S * 4, h->stream));
        if (T > 1) CK(cudaMemsetAsync(h->cnt[(T - 2) % 3].p, 0, (size_t)S * 4, h->stream));
        k_count_final<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(ab + (size_t)(T - 1) * S, N, cnt_t);
        LAUNCH_CHECK();
        int* xs_next = h->xstar[h->xstar_cur ^ 1].as<int>();
        // the whole backward pass as one cooperative launch when its grid fits the device at once
        bool counted = false;
        if (h->use_resident && h->coop != 0) {
            if (h->coop < 0) {
                int v = 0;
                h->coop = (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->opt.device) == cudaSuccess && v) ? 1 : 0;
            }
            int cap = 0;
            const size_t smem_cb = (size_t)SB_WALK_CHUNK * K * 4;
            if (h->coop && resident_ctas(h, k_count_back_all, smem_cb, &cap) == SB_OK) {
                SbCountArgs ca;
                ca.x_hist = xb; ca.anc_hist = ab; ca.has_factor = h->hmm_hasf.as<unsigned char>();
                for (int i = 0; i < 3; ++i) ca.cnt[i] = h->cnt[i].as<int>();
                ca.S = S; ca.n_pool = n_pool; ca.N = N; ca.K = K; ca.T = T;
                ca.marg = h->marg.as<unsigned long long>(); ca.lineage = lineage; ca.xstar_out = xs_next;
                const int want = (int)((n_pool + 2047) / 2048);
                const unsigned ctas = (unsigned)(want < cap ? want : cap);
                // walk phase: live-entry counters per step + the list cursor (zeroed), the compacted list
                CK(h->cb_live.ensure(((size_t)T + 1) * 4));
                CK(h->cb_list.ensure((size_t)ctas * 256 * sizeof(int2)));
                CK(cudaMemsetAsync(h->cb_live.p, 0, ((size_t)T + 1) * 4, h->stream));
                ca.live = h->cb_live.as<unsigned>(); ca.cursor = ca.live + T; ca.list = h->cb_list.as<int2>();
                ca.bar = grid_bar(h); ca.bar0 = h->bar_count;
                void* args[] = {(void*)&ca};
                CK(cudaLaunchCooperativeKernel((const void*)k_count_back_all, dim3(ctas), dim3(256), args, smem_cb, h->stream));
                h->bar_count += (unsigned)T * ctas;                      // every CTA arrives T times whatever the switch point
                ++h->launches;
                counted = true;
            } else h->err.clear();
        }
        for (int t = T - 1; t >= 0 && !counted; --t) {
            int* c_t = h->cnt[t % 3].as<int>();
            int* c_p = t ? h->cnt[(t - 1) % 3].as<int>() : nullptr;
            int* c_clear = t >= 2 ? h->cnt[(t + 1) % 3].as<int>() : nullptr;
            k_count_back<<<(unsigned)((n_pool + 2047) / 2048), 256, (size_t)K * 4, h->stream>>>(
                xb + (size_t)t * S, c_t, t ? ab + (size_t)(t - 1) * S : nullptr, t ? !h->has_factor[t - 1] : 0,
                n_pool, N, K, t, c_p, c_clear, h->marg.as<unsigned long long>(), lineage, xs_next);
            LAUNCH_CHECK();
        }
        if (joint) {
            k_joint_hist<<<(unsigned)((N + 255) / 256), 256, 0, h->stream>>>(xb, ab, h->hmm_hasf.as<unsigned char>(), S, N, K, T,
                                                                            h->joint.as<unsigned long long>());
            LAUNCH_CHECK();
        }
        h->xstar_cur ^= 1;
        h->have_retained = true;
        if (it == 0) {
            double lz;
            if ((rc = fetch_logZ(h, &lz)) != SB_OK) return rc;
            h->stats.logZ_first = lz;
            if ((rc = check_err_word(h)) != SB_OK) return rc;
        }
    }
    if ((rc = end_timed(h)) != SB_OK) return rc;
    h->stats.particle_steps = iters * N * (int64_t)T;
    double lz;
    if ((rc = fetch_logZ(h, &lz)) != SB_OK) return rc;
    h->stats.logZ_last = lz;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    if (marg) {
        std::vector<unsigned long long> tmp((size_t)T * K);
        CK(cudaMemcpy(tmp.data(), h->marg.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) marg[i] += (int64_t)tmp[i];
    }
    if (joint) {
        std::vector<unsigned long long> tmp((size_t)njoint);
        CK(cudaMemcpy(tmp.data(), h->joint.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) joint[i] += (int64_t)tmp[i];
    }
    return SB_OK;
}

extern "C" int32_t sb_enum_hmm(sb_handle* h, double* joint, double* marg, double* out_logZ) {
    if (!h) return SB_EINVAL;
    if (h->model != MODEL_HMM)
        return fail(h, SB_EUNSUPPORTED, "enum: only the fixed-structure discrete model can be enumerated on the GPU (no CPU fallback)");
    CK(cudaSetDevice(h->opt.device));
    const int T = h->T, K = h->K;
    int rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    double lz = 0.0;
    // exact per-time marginals and log-evidence (forward-backward)
    CK(h->enum_buf.ensure(((size_t)2 * T * K + T + 1) * 8));
    double* alpha = h->enum_buf.as<double>();
    double* mg = alpha + (size_t)T * K;
    double* cn = mg + (size_t)T * K;
    double* d_lz = cn + T;
    k_forward_backward<<<1, 64, 0, h->stream>>>(h->hmm_A.as<double>(), h->hmm_p0.as<double>(), h->hmm_logF.as<double>(),
                                                h->hmm_hasf.as<unsigned char>(), K, T, alpha, cn, mg, d_lz);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(&lz, d_lz, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (!(lz > -INFINITY) || lz != lz) return fail(h, SB_EZEROMASS, "Distribution has zero probability mass.");
    if (marg) CK(cudaMemcpy(marg, mg, (size_t)T * K * 8, cudaMemcpyDeviceToHost));    // after the stream sync above
    if (joint) {
        if (T * log2((double)K) > 26.0) return fail(h, SB_EUNSUPPORTED, "enum: the joint needs K^T <= 2^26 paths; pass joint = NULL for per-time marginals");
        long long nc = 1; for (int t = 0; t < T; ++t) nc *= K;
        DevBuf tmp;
        CK(tmp.ensure((size_t)nc * 8));
        k_enum_joint<<<(unsigned)((nc + 255) / 256), 256, 0, h->stream>>>(h->hmm_logA.as<double>(), h->hmm_logp0.as<double>(),
                                                                       h->hmm_logF.as<double>(), h->hmm_hasf.as<unsigned char>(), K, T, nc, tmp.as<double>());
        ++h->launches;
        cudaError_t e1 = cudaGetLastError();
        // the handle's stream does not synchronise with the default stream: copy on it, then wait
        cudaError_t e2 = e1 == cudaSuccess ? cudaMemcpyAsync(joint, tmp.p, (size_t)nc * 8, cudaMemcpyDeviceToHost, h->stream) : e1;
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(h->stream);
        tmp.release();
        if (e2 != cudaSuccess) return fail(h, SB_ECUDA, std::string("enum joint: ") + cudaGetErrorString(e2));
    }
    if ((rc = end_timed(h)) != SB_OK) return rc;
    if (out_logZ) *out_logZ = lz;
    return SB_OK;
}

extern "C" int32_t sb_mh_run(sb_handle* h, int64_t chains, int64_t steps, int64_t chain0, int64_t* counts) {
    if (!h) return SB_EINVAL;
    if (h->model != MODEL_BNET)
        return fail(h, SB_EUNSUPPORTED, "mh: only fixed-structure binary Bayes nets are supported (no CPU fallback)");
    if (chains < 1 || steps < 0 || !counts) return fail(h, SB_EINVAL, "mh: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    const int nv = 1 << __builtin_popcount(h->bn_query);
    CK(h->bn_counts.ensure((size_t)(nv + 1) * 8));
    CK(cudaMemsetAsync(h->bn_counts.p, 0, (size_t)(nv + 1) * 8, h->stream));
    SbBnetArgs a;
    a.D = h->bn_D; a.F = h->bn_F; a.stride = h->bn_stride; a.query_mask = h->bn_query;
    a.site_parents = h->bn_sp.as<unsigned>(); a.site_cpt = h->bn_cpt.as<double>();
    a.site_lp1 = h->bn_lp1.as<double>(); a.site_lp0 = h->bn_lp0.as<double>();
    a.fac_parents = h->bn_fp.as<unsigned>(); a.fac_logf = h->bn_fac.as<double>();
    a.seed = h->opt.seed; a.chains = chains; a.steps = steps; a.chain0 = chain0;
    a.counts = h->bn_counts.as<unsigned long long>(); a.accepts = a.counts + nv;
    // CTA size: the one that leaves the fewest warps on the fullest SM (65536 chains on 148 SMs: 128 threads -> 4 CTAs = 16
    // warps on some SMs, 64 threads -> 7 CTAs = 14 warps; the kernel is bound by its fullest SMs); larger CTAs on a tie
    int threads = 128;
    {
        long long best = -1;
        for (int t = 128; t >= 32; t >>= 1) {
            const long long ctas = (chains + t - 1) / t, per_sm = (ctas + h->num_sms - 1) / h->num_sms, warps = per_sm * (t / 32);
            if (best < 0 || warps < best) { best = warps; threads = t; }
        }
        if (const char* ov = getenv("SB_MH_THREADS")) { const int v = atoi(ov); if (v == 32 || v == 64 || v == 128) threads = v; }
    }
    const size_t smem = (size_t)(3 * a.D + a.F) * a.stride * 8 + (size_t)(a.D + a.F) * 4 + (size_t)(threads / 32) * nv * 4;
    int rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    if (a.D <= SB_MH_TABD && !getenv("SB_MH_NOTAB")) {
        // small nets: per-(site, state) tables in shared memory (k_mh_bnet_tab); bit-identical chains
        const size_t ns = (size_t)1 << a.D;
        const size_t smem_t = (size_t)(4 * a.D + 1) * ns * 8 + (size_t)(threads / 32) * nv * 4 + ns * 2;   // + the value table (u16 per state)
        if (smem_t > 48 * 1024) CK(cudaFuncSetAttribute(k_mh_bnet_tab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
        k_mh_bnet_tab<<<(unsigned)((chains + threads - 1) / threads), threads, smem_t, h->stream>>>(a);
    } else {
        if (smem > 200 * 1024)
            return fail(h, SB_EUNSUPPORTED, "mh: the net's tables do not fit in shared memory (no CPU fallback)");
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_mh_bnet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_mh_bnet<<<(unsigned)((chains + threads - 1) / threads), threads, smem, h->stream>>>(a);
    }
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    std::vector<unsigned long long> tmp((size_t)nv + 1);
    CK(cudaMemcpy(tmp.data(), h->bn_counts.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
    for (int v = 0; v < nv; ++v) counts[v] += (int64_t)tmp[v];
    h->stats.accepts = (int64_t)tmp[nv];
    h->stats.particle_steps = chains * steps;
    return SB_OK;
}

// ======================================================================= hooks
namespace {
// one caller-supplied systematic offset u in [0,1) -> 53-bit integer in d_r53[0]
int stage_offset(sb_handle* h, double u) {
    const size_t need = sizeof(unsigned long long) * (size_t)(h->T > 1 ? h->T : 1);
    CK(h->d_r53.ensure(need));
    CK(h->d_info.ensure(sizeof(SbStepInfo) * (size_t)(h->T > 1 ? h->T : 1)));
    h->r53_host.resize(1);
    h->r53_host[0] = (unsigned long long)(u * 0x1p53);
    CK(cudaMemcpyAsync(h->d_r53.p, h->r53_host.data(), sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
    return SB_OK;
}
int stage_in(sb_handle* h, DevBuf& b, const void* src, size_t bytes) {
    CK(b.ensure(bytes));
    CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return SB_OK;
}
}  // namespace

extern "C" int32_t sb_dev_resample_systematic(sb_handle* h, const void* d_weights, int32_t kind, int64_t n,
                                              double u, int64_t m, int32_t* d_anc) {
    if (!h) return SB_EINVAL;
    if (!d_weights || !d_anc || n < 1 || m < 1 || !(u >= 0.0 && u < 1.0)) return fail(h, SB_EINVAL, "resample_systematic: bad arguments");
    int rc;
    if ((rc = ensure_grid(h, n, m)) != SB_OK) return rc;
    if ((rc = stage_offset(h, u)) != SB_OK) return rc;
    if ((rc = launch_quantise(h, d_weights, kind, n, 0)) != SB_OK) return rc;
    if ((rc = launch_tilescan(h, 0, n, m, h->d_r53.as<unsigned long long>(), h->d_info.as<SbStepInfo>())) != SB_OK) return rc;
    return launch_pull(h, 0, h->d_info.as<SbStepInfo>(), m, d_anc);
}

extern "C" int32_t sb_resample_systematic(sb_handle* h, const void* weights, int32_t kind, int64_t n,
                                          double u, int64_t m, int32_t* anc_out) {
    if (!h) return SB_EINVAL;
    if (!weights || !anc_out || n < 1 || m < 1) return fail(h, SB_EINVAL, "resample_systematic: bad arguments");
    if (kind < 0 || kind > 2) return fail(h, SB_EINVAL, "unknown weight kind");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_in, weights, (size_t)n * kind_bytes(kind))) != SB_OK) return rc;
    CK(h->hk_anc.ensure((size_t)tiles_of(m) * SB_TILE * 4));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    if ((rc = sb_dev_resample_systematic(h, h->hk_in.p, kind, n, u, m, h->hk_anc.as<int>())) != SB_OK) return rc;
    if ((rc = end_timed(h)) != SB_OK) return rc;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    CK(cudaMemcpy(anc_out, h->hk_anc.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int32_t sb_resample_multinomial(sb_handle* h, const void* weights, int32_t kind, int64_t n,
                                           const double* r, int64_t m, int32_t* anc_out) {
    if (!h) return SB_EINVAL;
    if (!weights || !anc_out || !r || n < 1 || m < 1) return fail(h, SB_EINVAL, "resample_multinomial: bad arguments");
    if (kind < 0 || kind > 2) return fail(h, SB_EINVAL, "unknown weight kind");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_in, weights, (size_t)n * kind_bytes(kind))) != SB_OK) return rc;
    if ((rc = stage_in(h, h->hk_r, r, (size_t)m * 8)) != SB_OK) return rc;
    CK(h->hk_anc.ensure((size_t)m * 4));
    if ((rc = ensure_grid(h, n, m)) != SB_OK) return rc;
    if ((rc = stage_offset(h, 0.0)) != SB_OK) return rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    if ((rc = launch_quantise(h, h->hk_in.p, kind, n, 0)) != SB_OK) return rc;
    h->hook_exact = 1;
    rc = launch_tilescan(h, 0, n, m, h->d_r53.as<unsigned long long>(), h->d_info.as<SbStepInfo>());
    h->hook_exact = 0;
    if (rc != SB_OK) return rc;
    if ((rc = launch_grid_multinomial(h, 0, h->d_info.as<SbStepInfo>(), n, m, h->hk_r.as<double>(), 0, 0, h->hk_anc.as<int>())) != SB_OK) return rc;
    if ((rc = end_timed(h)) != SB_OK) return rc;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    CK(cudaMemcpy(anc_out, h->hk_anc.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int32_t sb_resample_literal(sb_handle* h, const double* w, int64_t n, const double* r, int64_t m,
                                       int32_t* anc_out) {
    if (!h) return SB_EINVAL;
    if (!w || !anc_out || !r || n < 1 || m < 1) return fail(h, SB_EINVAL, "resample_literal: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_in, w, (size_t)n * 8)) != SB_OK) return rc;
    if ((rc = stage_in(h, h->hk_r, r, (size_t)m * 8)) != SB_OK) return rc;
    CK(h->hk_anc.ensure((size_t)m * 4));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    if ((rc = launch_literal(h, h->hk_in.as<double>(), n, m, h->hk_r.as<double>(), 0, 0, h->hk_anc.as<int>())) != SB_OK) return rc;
    if ((rc = end_timed(h)) != SB_OK) return rc;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    CK(cudaMemcpy(anc_out, h->hk_anc.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int32_t sb_logsumexp(sb_handle* h, const void* lw, int32_t kind, int64_t n, double* out) {
    if (!h) return SB_EINVAL;
    if (!lw || !out || n < 1 || (kind != SB_W_LOG_F64 && kind != SB_W_LOG_F32)) return fail(h, SB_EINVAL, "logsumexp: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_in, lw, (size_t)n * kind_bytes(kind))) != SB_OK) return rc;
    int np = (int)((n + 8191) / 8192);
    if (np > 1184) np = 1184;                       // 148 SMs x 8 CTAs
    CK(h->lse_part.ensure((size_t)np * sizeof(SbLse) + 8));
    SbLse* part = h->lse_part.as<SbLse>();
    double* d_out = reinterpret_cast<double*>(part + np);
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    if (kind == SB_W_LOG_F64) k_lse_partial<double><<<np, 256, 0, h->stream>>>(h->hk_in.as<double>(), n, part, h->d_err);
    else k_lse_partial<float><<<np, 256, 0, h->stream>>>(h->hk_in.as<float>(), n, part, h->d_err);
    LAUNCH_CHECK();
    k_lse_final<<<1, 256, 0, h->stream>>>(part, np, d_out);
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    CK(cudaMemcpy(out, d_out, 8, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int32_t sb_dev_gather(sb_handle* h, const void* d_src, int32_t elem_bytes, int64_t n,
                                 const int32_t* d_anc, int64_t m, void* d_dst) {
    if (!h) return SB_EINVAL;
    if (!d_src || !d_anc || !d_dst || n < 1 || m < 1) return fail(h, SB_EINVAL, "gather: bad arguments");
    const unsigned grid = (unsigned)((m + 1023) / 1024);
    if (elem_bytes == 4) k_gather<uint32_t><<<grid, 256, 0, h->stream>>>((const uint32_t*)d_src, d_anc, m, (uint32_t*)d_dst);
    else if (elem_bytes == 8) k_gather<unsigned long long><<<grid, 256, 0, h->stream>>>((const unsigned long long*)d_src, d_anc, m, (unsigned long long*)d_dst);
    else return fail(h, SB_EINVAL, "gather: elem_bytes must be 4 or 8");
    LAUNCH_CHECK();
    return SB_OK;
}

extern "C" int32_t sb_gather(sb_handle* h, const void* src, int32_t elem_bytes, int64_t n, const int32_t* anc,
                             int64_t m, void* dst) {
    if (!h) return SB_EINVAL;
    if (!src || !anc || !dst || n < 1 || m < 1 || (elem_bytes != 4 && elem_bytes != 8)) return fail(h, SB_EINVAL, "gather: bad arguments");
    for (int64_t i = 0; i < m; ++i) if (anc[i] < 0 || anc[i] >= n) return fail(h, SB_EINVAL, "gather: ancestor index out of range");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_src, src, (size_t)n * elem_bytes)) != SB_OK) return rc;
    if ((rc = stage_in(h, h->hk_anc, anc, (size_t)m * 4)) != SB_OK) return rc;
    CK(h->hk_out.ensure((size_t)m * elem_bytes));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    if ((rc = sb_dev_gather(h, h->hk_src.p, elem_bytes, n, h->hk_anc.as<int>(), m, h->hk_out.p)) != SB_OK) return rc;
    if ((rc = end_timed(h)) != SB_OK) return rc;
    CK(cudaMemcpy(dst, h->hk_out.p, (size_t)m * elem_bytes, cudaMemcpyDeviceToHost));
    return SB_OK;
}

// ======================================================================= ERP hooks
extern "C" int32_t sb_erp_sample(sb_handle* h, int32_t erp, const double* params, int64_t n, uint64_t seed, double* out) {
    if (!h) return SB_EINVAL;
    if (!params || !out || n < 1 || erp < 0 || erp > 5) return fail(h, SB_EINVAL, "erp_sample: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    CK(h->hk_out.ensure((size_t)n * 8));
    int rc;
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    k_erp_sample<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(erp, params[0], erp == SB_ERP_FLIP || erp == SB_ERP_RANDOMINTEGER ? 0.0 : params[1], n, seed, h->hk_out.as<double>());
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    CK(cudaMemcpy(out, h->hk_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return SB_OK;
}
extern "C" int32_t sb_erp_logpdf(sb_handle* h, int32_t erp, const double* params, const double* x, int64_t n, double* out) {
    if (!h) return SB_EINVAL;
    if (!params || !x || !out || n < 1 || erp < 0 || erp > 5) return fail(h, SB_EINVAL, "erp_logpdf: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_in, x, (size_t)n * 8)) != SB_OK) return rc;
    CK(h->hk_out.ensure((size_t)n * 8));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    k_erp_logpdf<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(erp, params[0], erp == SB_ERP_FLIP || erp == SB_ERP_RANDOMINTEGER ? 0.0 : params[1], h->hk_in.as<double>(), n, h->hk_out.as<double>());
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    CK(cudaMemcpy(out, h->hk_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return SB_OK;
}
// z = sum(values(dict)) (src/erp.jl:46), accumulated in order like the oracle; zero mass is the reference's error (:47)
static int discrete_z(sb_handle* h, const double* p, int32_t k, double* z) {
    double acc = 0.0;
    for (int32_t j = 0; j < k; ++j) acc += p[j];
    if (acc == 0.0) return fail(h, SB_EZEROMASS, "Distribution has zero probability mass.");
    *z = acc;
    return SB_OK;
}
extern "C" int32_t sb_erp_sample_discrete(sb_handle* h, const double* p, int32_t k, int64_t n, uint64_t seed, int32_t* out_index) {
    if (!h) return SB_EINVAL;
    if (!p || !out_index || n < 1 || k < 1) return fail(h, SB_EINVAL, "discrete: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    double z;
    if ((rc = discrete_z(h, p, k, &z)) != SB_OK) return rc;
    if ((rc = stage_in(h, h->hk_in, p, (size_t)k * 8)) != SB_OK) return rc;
    CK(h->hk_anc.ensure((size_t)n * 4));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    k_discrete_sample<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->hk_in.as<double>(), k, z, n, seed, h->hk_anc.as<int>(), h->d_err);
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    if ((rc = check_err_word(h)) != SB_OK) return rc;
    CK(cudaMemcpy(out_index, h->hk_anc.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}
extern "C" int32_t sb_erp_logpdf_discrete(sb_handle* h, const double* p, int32_t k, const int32_t* index, int64_t n, double* out) {
    if (!h) return SB_EINVAL;
    if (!p || !index || !out || n < 1 || k < 1) return fail(h, SB_EINVAL, "discrete: bad arguments");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    double z;
    if ((rc = discrete_z(h, p, k, &z)) != SB_OK) return rc;
    if ((rc = stage_in(h, h->hk_in, p, (size_t)k * 8)) != SB_OK) return rc;
    if ((rc = stage_in(h, h->hk_src, index, (size_t)n * 4)) != SB_OK) return rc;
    CK(h->hk_out.ensure((size_t)n * 8));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    k_discrete_logpdf<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->hk_in.as<double>(), k, z, h->hk_src.as<int>(), n, h->hk_out.as<double>());
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    CK(cudaMemcpy(out, h->hk_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return SB_OK;
}
extern "C" int32_t sb_erp_sample_dirichlet(sb_handle* h, const double* alpha, int32_t k, int64_t n, uint64_t seed, double* out) {
    if (!h) return SB_EINVAL;
    if (!alpha || !out || n < 1 || k < 1 || k > SB_DIR_MAXK) return fail(h, SB_EINVAL, "dirichlet: bad arguments (k <= 32)");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_in, alpha, (size_t)k * 8)) != SB_OK) return rc;
    CK(h->hk_out.ensure((size_t)n * k * 8));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    k_dirichlet_sample<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->hk_in.as<double>(), k, n, seed, h->hk_out.as<double>());
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    CK(cudaMemcpy(out, h->hk_out.p, (size_t)n * k * 8, cudaMemcpyDeviceToHost));
    return SB_OK;
}
extern "C" int32_t sb_erp_logpdf_dirichlet(sb_handle* h, const double* alpha, int32_t k, const double* x, int64_t n, double* out) {
    if (!h) return SB_EINVAL;
    if (!alpha || !x || !out || n < 1 || k < 1 || k > SB_DIR_MAXK) return fail(h, SB_EINVAL, "dirichlet: bad arguments (k <= 32)");
    CK(cudaSetDevice(h->opt.device));
    int rc;
    if ((rc = stage_in(h, h->hk_in, alpha, (size_t)k * 8)) != SB_OK) return rc;
    if ((rc = stage_in(h, h->hk_src, x, (size_t)n * k * 8)) != SB_OK) return rc;
    CK(h->hk_out.ensure((size_t)n * 8));
    if ((rc = begin_timed(h)) != SB_OK) return rc;
    k_dirichlet_logpdf<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->hk_in.as<double>(), k, h->hk_src.as<double>(), n, h->hk_out.as<double>());
    LAUNCH_CHECK();
    if ((rc = end_timed(h)) != SB_OK) return rc;
    CK(cudaMemcpy(out, h->hk_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return SB_OK;
}

// ======================================================================= state access
extern "C" int32_t sb_get_history(sb_handle* h, int32_t t, int64_t count, void* x_out, int32_t* anc_out, double* lw_out) {
    if (!h) return SB_EINVAL;
    if (h->sw_N == 0)
        return fail(h, SB_EINVAL, h->small_cur >= 0 ? "get_history: history not available (the last call ran as a one-tile chain set, whose history stays on chip)"
                                                    : "get_history: no sweep has run since the model was loaded");
    if (t < 0 || t >= h->sw_T || count < 1 || count > h->sw_S) return fail(h, SB_EINVAL, "get_history: bad index");
    CK(cudaSetDevice(h->opt.device));
    CK(cudaStreamSynchronize(h->stream));
    const int64_t S = h->sw_S;
    if (x_out) {
        if (h->model != MODEL_HMM || !h->sw_hist) return fail(h, SB_EINVAL, "get_history: state history was not kept");
        const int* src = h->dist_ready ? reinterpret_cast<const int*>((char*)h->arena + h->off_xh) : h->xbuf.as<int>();
        CK(cudaMemcpy(x_out, src + (size_t)t * S, (size_t)count * 4, cudaMemcpyDeviceToHost));
    }
    if (anc_out) {
        if (h->model != MODEL_HMM || !h->sw_hist) return fail(h, SB_EINVAL, "get_history: ancestor history was not kept");
        CK(cudaMemcpy(anc_out, h->ancbuf.as<int>() + (size_t)t * S, (size_t)count * 4, cudaMemcpyDeviceToHost));
    }
    if (lw_out) {
        if (!h->opt.store_logw) return fail(h, SB_EINVAL, "get_history: store_logw was off");
        CK(cudaMemcpy(lw_out, h->lwbuf.as<double>() + (size_t)t * S, (size_t)count * 8, cudaMemcpyDeviceToHost));
    }
    return SB_OK;
}

extern "C" int32_t sb_get_retained(sb_handle* h, int32_t* xstar_out) {
    if (!h) return SB_EINVAL;
    if (!xstar_out || !h->have_retained) return fail(h, SB_EINVAL, "get_retained: no retained particle");
    CK(cudaSetDevice(h->opt.device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->small_cur >= 0) CK(cudaMemcpy(xstar_out, h->sm_xs.as<int>() + (size_t)h->small_cur * h->T, (size_t)h->T * 4, cudaMemcpyDeviceToHost));
    else if (h->dist_ready) CK(cudaMemcpy(xstar_out, (char*)h->arena + h->off_xstar[h->xstar_cur], (size_t)h->T * 4, cudaMemcpyDeviceToHost));
    else CK(cudaMemcpy(xstar_out, h->xstar[h->xstar_cur].p, (size_t)h->T * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int32_t sb_get_particles(sb_handle* h, int64_t count, void* x_out) {
    if (!h) return SB_EINVAL;
    if (!x_out || h->sw_N == 0 || count < 1 || count > h->sw_N) return fail(h, SB_EINVAL, "get_particles: no sweep / bad count");
    CK(cudaSetDevice(h->opt.device));
    CK(cudaStreamSynchronize(h->stream));
    const int t = h->sw_T - 1;
    if (h->model == MODEL_HMM) {
        const int row = h->sw_hist ? t : h->sw_slot;
        const int* src = h->dist_ready ? reinterpret_cast<const int*>((char*)h->arena + (h->sw_hist ? h->off_xh + (size_t)t * (size_t)h->sw_S * 4 : h->off_x[h->sw_slot]))
                                       : h->xbuf.as<int>() + (size_t)row * h->sw_S;
        CK(cudaMemcpy(x_out, src, (size_t)count * 4, cudaMemcpyDeviceToHost));
    } else {
        const size_t sz = (h->opt.precision == SB_FP32 && h->model != MODEL_OPS) ? 4 : 8;
        const int XD = h->model == MODEL_OPS ? h->ops_D : 1;
        if (XD > 1) {                                                // Dirichlet program: D planes of `count` doubles
            const char* src = h->xbuf.as<char>() + (size_t)h->sw_slot * h->sw_S * 8 * XD;
            for (int d = 0; d < XD; ++d)
                CK(cudaMemcpy((double*)x_out + (size_t)d * count, src + (size_t)d * h->sw_S * 8, (size_t)count * 8, cudaMemcpyDeviceToHost));
            return SB_OK;
        }
        std::vector<char> tmp((size_t)count * sz);
        const char* src = h->dist_ready ? (char*)h->arena + h->off_x[h->sw_slot] : h->xbuf.as<char>() + (size_t)h->sw_slot * h->sw_S * sz;
        CK(cudaMemcpy(tmp.data(), src, tmp.size(), cudaMemcpyDeviceToHost));
        double* o = (double*)x_out;
        if (sz == 4) for (int64_t i = 0; i < count; ++i) o[i] = (double)((float*)tmp.data())[i];
        else memcpy(o, tmp.data(), tmp.size());
    }
    return SB_OK;
}

// ======================================================================= multi-GPU
// One process per GPU.  Every rank keeps its share of the particle pool (both double-buffer slots:
// weights, states, warp prefixes, tile records) and a flag word per peer in ONE device allocation,
// the arena, exported over CUDA IPC.  After the import every rank holds a mapping of every arena:
// the fused step kernel reads parent weights and states of any GPU with ordinary loads over
// NVLink, the tile scan reads every GPU's tile records behind a flag barrier (raised by the last CTA of the
// producing step kernel, sb_signal_done).  No NCCL
// call sits on the data path; the 64-byte handles travel over the host's own plumbing
// (torch.distributed.all_gather in the Python mirror).
extern "C" int32_t sb_dist_init(sb_handle* h, int32_t rank, int32_t world) {
    if (!h) return SB_EINVAL;
    if (world < 1 || world > SB_MAX_PEERS || (world & (world - 1)) || rank < 0 || rank >= world)
        return fail(h, SB_EINVAL, "dist_init: world must be 1, 2, 4 or 8 and 0 <= rank < world");
    if (h->arena) return fail(h, SB_EINVAL, "dist_init: handle is already sharded");
    h->rank = rank; h->world = world;
    return SB_OK;
}

extern "C" int32_t sb_dist_reserve_history(sb_handle* h, int32_t rows) {
    if (!h) return SB_EINVAL;
    if (h->arena) return fail(h, SB_EINVAL, "dist_reserve_history: call it before sb_dist_ipc_export (the rows live in the exported allocation)");
    if (rows < 0) return fail(h, SB_EINVAL, "dist_reserve_history: rows must be >= 0");
    h->hist_rows = rows;
    return SB_OK;
}

extern "C" int32_t sb_dist_ipc_export(sb_handle* h, int64_t N_local, uint8_t out_handle[64]) {
    if (!h || !out_handle) return SB_EINVAL;
    if (h->arena) return fail(h, SB_EINVAL, "dist_ipc_export: arena already allocated");
    if (N_local < SB_TILE || (N_local & (N_local - 1)) || N_local * h->world > (1ll << 27))
        return fail(h, SB_EINVAL, "dist_ipc_export: N_local must be a power of two >= 2048 with world x N_local <= 2^27");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CK(cudaSetDevice(h->opt.device));
    const size_t ntl = (size_t)(N_local / SB_TILE);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    for (int sl = 0; sl < 2; ++sl) {
        h->off_w32[sl] = take((size_t)N_local * 4);
        h->off_x[sl] = take((size_t)N_local * 8);              // room for fp64 states
        h->off_wp[sl] = take(ntl * SB_WARPS * 8);
        h->off_e[sl] = take(ntl * h->world * 4);               // replicated: every rank's records, global tile index
        h->off_Q[sl] = take(ntl * h->world * 8);
    }
    h->off_flags = take(SB_REC_OFF + 2 * SB_MAX_PEERS * 16);        // flags[8] + the CTA counter of the step kernel; the GPUs' summary records (2 slots)
    if (h->hist_rows > 0) {                                         // sharded conditional SMC / PMCMC (sb_dist_back.cuh)
        h->off_xh = take((size_t)h->hist_rows * (size_t)N_local * 4);
        h->off_cnt[0] = take((size_t)N_local * 4); h->off_cnt[1] = take((size_t)N_local * 4);
        h->off_lin = take(2 * sizeof(int));
        h->off_xstar[0] = take((size_t)h->hist_rows * 4); h->off_xstar[1] = take((size_t)h->hist_rows * 4);
    }
    CK(cudaMalloc(&h->arena, off));
    h->arena_bytes = off;
    CK(cudaMemset(h->arena, 0, off));
    CK(cudaDeviceSynchronize());
    h->NL = N_local;
    cudaIpcMemHandle_t ipc;
    CK(cudaIpcGetMemHandle(&ipc, h->arena));
    memcpy(out_handle, &ipc, 64);
    return SB_OK;
}

extern "C" int32_t sb_dist_ipc_import(sb_handle* h, const uint8_t* all_handles) {
    if (!h || !all_handles) return SB_EINVAL;
    if (!h->arena) return fail(h, SB_EINVAL, "dist_ipc_import: call sb_dist_ipc_export first");
    CK(cudaSetDevice(h->opt.device));
    for (int g = 0; g < h->world; ++g) {
        if (g == h->rank) { h->peer_base[g] = (char*)h->arena; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, all_handles + (size_t)g * 64, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
        h->peer_base[g] = (char*)p;
    }
    h->epoch = 0;
    h->dist_ready = true;
    return SB_OK;
}
