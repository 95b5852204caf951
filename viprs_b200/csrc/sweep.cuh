// viprs_b200 -- the one-pass Gauss-Seidel sweep (CAVI E-step) for sm_100a.
//
// Replaces e_step<T,U,I> / e_step_mixture<T,U,I> + update_q_factor<T,U,I> of the reference
// (/root/reference/viprs/model/vi/e_step.hpp:343-442, 447-551 and 307-338), threads=1 order.
//
// One CTA owns one LD block (rows r0..r1 that only reach columns inside [r0, r1)), keeps the strictly
// sequential per-SNP update order, and reads every LD entry ONCE from HBM:
//
//   q_j used at step j  =  dq * ( F_j + B_j ),
//       F_j = sum_{i<j} R_ij eta_i(new)     "forward"  -- axpy of finished rows
//       B_j = sum_{k>j} R_jk eta_k(old)     "backward" -- dot of row j with the not-yet-updated etas
//
// which is algebraically what the reference's incrementally maintained q holds at step j (its second
// pass, update_q_factor, is exactly the B_j term deferred to the end of the sweep).
//
// Warp roles (CTA = NBW + 2 warps, no __syncthreads after the prologue):
//   warps 0..NBW-1  bulk : A(u): B_j for the rows of panel u (full-row dots against eta_old in shared
//                      memory) + the rows' window coefficients; C(v): axpy of panel v's finished rows into
//                      f_s[] for the columns >= cut_j = ceil((j+33)/EPV)*EPV.
//   warp NBW     producer: one 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) per row panel into an NST-deep
//                      shared-memory ring, plus cp.async.bulk.prefetch.L2 a few panels further ahead;
//                      writes the per-row metadata ring.
//   warp NBW+1   chain   : the per-SNP scalar update, one row panel (1..16 rows) per hand-shake.  Lane l owns
//                      the block-local columns c == l (mod 32) of a sliding 64-column window (X0: the column
//                      in [j, j+32), X1: +32): the forward axpy of row j into the columns < cut_j is two FMAs
//                      per lane per step on register accumulators, so the serial dependence between
//                      consecutive SNPs never leaves the warp (one SHFL per step).  It is the highest-numbered
//                      warp of the CTA (issue priority).
// Hand-offs: mbarriers full/empty (TMA ring) and cdone (chain finished a panel), all waited on with the
// hardware-suspending try_wait, plus per-bulk-warp release/acquire progress counters a_prog / c_prog that
// only the chain warp polls.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "em_math.cuh"

namespace vb {

// shared-memory carve-up (byte offsets from the dynamic shared memory base, all 16-byte aligned)
struct SmemLayout {
    uint32_t stages, eta, f, rowmeta, panelmeta, partial, alpha, wwin, bars, counters, total;
};
__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }
__host__ __device__ inline uint32_t align128(uint32_t x) { return (x + 127u) & ~127u; }
inline int state_pad(int max_block) { return ((max_block + 63) / 64) * 64 + 64; }
inline SmemLayout make_layout(int bpad, int tsize, int stage_bytes, int nst) {
    SmemLayout L;
    uint32_t o = 0;
    L.stages = o;    o += (uint32_t)nst * (uint32_t)stage_bytes;       o = align128(o);
    L.eta = o;       o += align16((uint32_t)bpad * tsize);
    L.f = o;         o += align16((uint32_t)bpad * tsize);
    L.rowmeta = o;   o += RR * (uint32_t)sizeof(int4);
    L.panelmeta = o; o += NST_MAX * (uint32_t)sizeof(int4);
    L.partial = o;   o += align16((uint32_t)NBW * RR * tsize);
    L.alpha = o;     o += align16((uint32_t)RR * tsize);
    L.wwin = o;      o += align16((uint32_t)RR * WW * tsize);
    L.bars = o;      o += 3 * NST_MAX * (uint32_t)sizeof(uint64_t);
    L.counters = o;  o += 2 * NBW * (uint32_t)sizeof(uint32_t);
    L.total = o;
    return L;
}

struct SweepPlan {                 // what ld.cu prepared (device pointers) + the ring geometry
    const unsigned char* packed;
    const int64_t* prow;           // [M+1] element offset of each packed row
    const int32_t* pcs;            // [M]   first (aligned) column of each packed row, global index
    const int32_t* blk_row;        // [n_blocks+1]
    const int32_t* blk_panel;      // [n_blocks+1]
    const int32_t* panel_row;      // [n_panels+1]
    const int32_t* blk_order;      // [n_blocks] LPT order
    const int32_t* panel_need;     // [n_panels] panels of the block that must be C-complete before the chain
                                   //            warp starts this panel (rows <= last row of the panel - WIN)
    int n_blocks;
    int stage_bytes;
    int nst;
    int bpad;
    int l2_ahead;                  // panels of L2 prefetch distance beyond the ring
    int n_sm;                      // SMs of the device (the register-resident kernel: CTA b and CTA b + n_sm share an SM)
    int smsp_rot;                  // register-resident kernel: placement of the bulk warps' tile shares on the SM sub-partitions
    unsigned long long* trace;     // debug timeline of CTA 0 (VIPRS_B200_TRACE), else null
    SmemLayout L;
};

// debug timeline of CTA 0 (build with -DVB_TRACE, run with VIPRS_B200_TRACE=<file>):
// trace[(role * 8 + ev) * kTracePanels + panel] = clock64()   (plain stores, no atomics)
constexpr int kTracePanels = 4096;
constexpr int kTraceSlots = 12 * 8 * kTracePanels;
#ifdef VB_TRACE
__device__ __forceinline__ void trace_ev(const SweepPlan& p, int lane, int role, int ev, int panel) {
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && panel < kTracePanels)
        p.trace[(role * 8 + ev) * kTracePanels + panel] = (unsigned long long)clock64();
}
#else
__device__ __forceinline__ void trace_ev(const SweepPlan&, int, int, int, int) {}
#endif

// ---------------------------------------------------------------------------------------------
// per-SNP update models (the chain warp's scalar math)
// ---------------------------------------------------------------------------------------------
// Spike-and-slab, e_step.hpp:397-431.
template <typename T>
struct SlabModel {
    struct Args {
        const T* std_beta; const T* u_logs; const T* sqrt_half_var_tau; const T* mu_mult;
        T* var_gamma; T* var_mu; T dq;
    };
    static constexpr bool kHeavy = false;
    static constexpr bool kDualGroups = true;      // chain: separate branch-free code for full 4-step groups
    static constexpr bool kFusedSums = true;       // the output role can accumulate the M-step / ELBO sums (layout 0)
    struct Raw { T beta, mm, sv, ul; };
    struct Lane { T c0, c1, a0, a1, ul; };      // mu = c1 X + c0 ;  sqrt(tau/2) mu = a1 X + a0
    struct Out { T mu, g; };
    static __device__ __forceinline__ void load_raw(const Args& a, int row, bool ok, Raw& r) {
        r.beta = T(0); r.mm = T(0); r.sv = T(0); r.ul = T(0);
        if (ok) { r.beta = a.std_beta[row]; r.mm = a.mu_mult[row]; r.sv = a.sqrt_half_var_tau[row]; r.ul = a.u_logs[row]; }
    }
    static __device__ __forceinline__ void derive(const Args& a, const Raw& r, Lane& L) {
        L.c0 = mul_t(r.mm, r.beta);          // mu = fma(mu_mult, beta, -mu_mult*q)   (:401)  with q = dq * X
        L.c1 = -mul_t(r.mm, a.dq);
        L.a0 = mul_t(r.sv, L.c0);            // u = sqrt_half_var_tau * mu  (:404), one FMA off X instead of two ops
        L.a1 = mul_t(r.sv, L.c1);
        L.ul = r.ul;
        if constexpr (sizeof(T) == 4) {
            // float: the logit is carried in base-2 units (u^2 + u_logs) * log2(e), so that the sigmoid's exponential
            // is a bare MUFU.EX2 on the serial path (one multiply fewer per SNP)
            L.a0 = mul_t(L.a0, T(1.2011224087864498));       // sqrt(log2 e)
            L.a1 = mul_t(L.a1, T(1.2011224087864498));
            L.ul = mul_t(L.ul, T(1.4426950408889634));
        }
    }
    // X: F_j + B_j in LD-code units; eo: eta_j before the update
    static __device__ __forceinline__ void step(const Lane& L, T X, T eo, T eps, T& en, T& d, bool& skip, Out& o) {
        const T mu = fma_t(L.c1, X, L.c0);
        const T uu = fma_t(L.a1, X, L.a0);                        // :404
        T g;
        if constexpr (sizeof(T) == 4) g = sigmoid2_t(fma_t(uu, uu, L.ul));     // :405, logit in base-2 units
        else g = sigmoid_t(fma_t(uu, uu, L.ul));
        d = fma_t(g, mu, -eo);                                    // :408
        skip = abs_t(d) < eps;                                    // :410-413
        en = skip ? eo : add_t(eo, d);                            // :431
        o.mu = mu; o.g = g;
    }
    static __device__ __forceinline__ void store(const Args& a, int row, bool skip, const Out& o) {
        if (!skip) { a.var_mu[row] = o.mu; a.var_gamma[row] = o.g; }            // :416-418
    }
};

// Sparse mixture (K slabs + null), e_step.hpp:501-536.  (M,K) arrays are C-order; K <= KMAX.
template <typename T, int KMAX>
struct MixModel {
    struct Args {
        const T* std_beta; const T* u_logs; const T* sqrt_half_var_tau; const T* mu_mult; const T* log_null_pi;
        T* var_gamma; T* var_mu; T dq; int K;
    };
    static constexpr bool kFusedSums = false;        // (M,K) sums stay in the streaming kernel
    static constexpr bool kHeavy = (KMAX > 4);       // register-hungry: always one CTA per SM
    static constexpr bool kDualGroups = false;       // the step is large: a second inlined copy costs more (instruction
                                                     // cache) than its per-step branches (measured on the C4 workload)
    struct Raw { T beta, lnp; T mm[KMAX], sv[KMAX], ul[KMAX]; };
    struct Lane { T beta, lnp, dq; T mm[KMAX], sv[KMAX], ul[KMAX]; int K; };
    struct Out { T mu[KMAX], g[KMAX]; };
    static __device__ __forceinline__ void load_raw(const Args& a, int row, bool ok, Raw& r) {
        r.beta = ok ? a.std_beta[row] : T(0);
        r.lnp = ok ? a.log_null_pi[row] : T(0);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const bool v = ok && k < a.K;
            const size_t m = (size_t)row * a.K + k;
            r.mm[k] = v ? a.mu_mult[m] : T(0);
            r.sv[k] = v ? a.sqrt_half_var_tau[m] : T(0);
            r.ul[k] = v ? a.u_logs[m] : T(0);
        }
    }
    static __device__ __forceinline__ void derive(const Args& a, const Raw& r, Lane& L) {
        L.beta = r.beta; L.lnp = r.lnp; L.dq = a.dq; L.K = a.K;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) { L.mm[k] = r.mm[k]; L.sv[k] = r.sv[k]; L.ul[k] = r.ul[k]; }
    }
    static __device__ __forceinline__ void step(const Lane& L, T X, T eo, T eps, T& en, T& d, bool& skip, Out& o) {
        (void)eps;
        const T r = fma_t(-L.dq, X, L.beta);                      // :505
        T u[KMAX];
        T mx = L.lnp;                                             // :515, c_max :58-71
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            o.mu[k] = mul_t(L.mm[k], r);                          // :509
            const T t = mul_t(L.sv[k], o.mu[k]);                  // :510
            u[k] = fma_t(t, t, L.ul[k]);                          // :511
            if (k < L.K) mx = u[k] > mx ? u[k] : mx;
        }
        T sum = expneg_t(mx - L.lnp);                             // softmax :222-241 (max-shifted)
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            u[k] = (k < L.K) ? expneg_t(mx - u[k]) : T(0);
            sum = add_t(sum, u[k]);
        }
        const T inv = fast_rcp_t(sum);
        d = -eo;                                                  // :519
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            o.g[k] = mul_t(u[k], inv);
            d = fma_t(o.g[k], o.mu[k], d);                        // :523
        }
        skip = false;                                             // no skip branch in the mixture sweep
        en = add_t(eo, d);                                        // :536
    }
    static __device__ __forceinline__ void store(const Args& a, int row, bool skip, const Out& o) {
        (void)skip;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < a.K) { a.var_mu[(size_t)row * a.K + k] = o.mu[k]; a.var_gamma[(size_t)row * a.K + k] = o.g[k]; }
    }
};

// trips of the chain's 4-step group loop that are unrolled: a rolled loop keeps the chain warp's code inside the
// instruction cache (the three warp roles execute disjoint code)
#ifndef VB_CHAIN_UNROLL
#define VB_CHAIN_UNROLL 1
#endif
constexpr int kChainUnroll = VB_CHAIN_UNROLL;

template <typename T> __device__ __forceinline__ T eps_of();
template <> __device__ __forceinline__ float eps_of<float>() { return 1.1920928955078125e-7f; }   // max(FLT_EPSILON, 1e-8)
template <> __device__ __forceinline__ double eps_of<double>() { return 1e-8; }                    // max(DBL_EPSILON, 1e-8)

// fext / bext (nullable): per-row external contributions to F_j + B_j, in LD-code units after scaling fext by fscale.
//   fext: forward-type terms -- what earlier tiles of a tiled LD block contributed (sum_{i in earlier tiles} R_ij
//         eta_i(new)), and/or the caller's q offset (q_in - dq (R - I) eta_in, in q units, fscale = 1 / dq); they are
//         part of the q this sweep leaves behind.
//   bext: backward-type terms of a tiled block (sum_{k in later tiles} R_jk eta_k(old)); like the in-tile backward dots
//         they are NOT part of the forward q.
template <typename T>
struct StateArgs {
    T* eta; T* q; T* eta_diff; const T* fext = nullptr; const T* bext = nullptr; T fscale = T(1);
    // fused M-step / ELBO reductions (register-resident kernel, version 2, spike-and-slab): when unit_partial != nullptr
    // the output role also accumulates the VIPRS_B200_S_* sums of its sweep unit (float64, fixed order) and writes them
    // to unit_partial[unit][VIPRS_B200_NSUMS]; reduce_units_kernel folds them per chromosome afterwards
    const double* n_per_snp = nullptr; const Theta* theta = nullptr; double* unit_partial = nullptr;
};

template <typename T> __device__ __forceinline__ void load_state_vec(const T* src, T* dst, int n16) {
    const uint4* sp = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        if (e < n16) {
            const uint4 t = sp[e];
            memcpy(reinterpret_cast<unsigned char*>(dst) + 16 * e, &t, 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// shared-memory views and the two roles both kernels share (producer, chain)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SmemView {
    unsigned char* base;
    int4* rowmeta;      // [RR] {byte offset of column 0 of the row, vs, ve, C-complete panels the NEXT panel needs}
    int4* panelmeta;    // [NST] {P, vmin, vmax, first local row}
    T* partial;         // [n_A_warps][RR] backward-dot partials
    T* alpha;           // [RR] eta_new of finished rows
    T* wwin;            // [RR][WW] R[j][j+1+k] for k < cut_j-j-1, else 0
    uint64_t* full;     // [NST] TMA landed
    uint64_t* empty;    // [NST] every C warp is done with the stage
    uint64_t* cdone;    // [NST] chain finished the panel
    uint32_t* prog;     // a_prog[NA] | c_prog[NC]
    const T* fsrc;      // forward accumulator as the chain reads it: fsrc[col & fmask]
    int fmask;
    // register-resident kernel only: panelmeta / cdone are rings indexed by panel (u & (PMR - 1)) instead of by TMA
    // stage (the stage is recycled as soon as the A warps are done with it), the chain publishes the rows it has
    // finished, and never runs more than PMR - 2 panels ahead of the slowest C warp
    bool by_panel = false;
    uint32_t* chain_rows = nullptr;
};
constexpr int PMR = 16;       // panel-indexed rings (power of two)
constexpr int CR = 128;       // per-row ring read by the C warps of the register-resident kernel (power of two)

// window coefficient storage: the state type, or int16 (integer LD codes are exact in it) to save shared memory
template <typename W> __device__ __forceinline__ float lds_w(uint32_t addr);
template <> __device__ __forceinline__ float lds_w<float>(uint32_t addr) { return lds_t(addr, float()); }
template <> __device__ __forceinline__ float lds_w<int16_t>(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return (float)v;
}
template <typename W> __device__ __forceinline__ void sts_w(uint32_t addr, float v);
template <> __device__ __forceinline__ void sts_w<float>(uint32_t addr, float v) { sts_t(addr, v); }
template <> __device__ __forceinline__ void sts_w<int16_t>(uint32_t addr, float v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((short)__float2int_rn(v)) : "memory");
}

template <typename U>
__device__ __forceinline__ void producer_role(const SweepPlan& p, unsigned char* smem, int4* rowmeta, int4* panelmeta,
                                              uint64_t* full, uint64_t* empty, int r0, int pan0, int NP, int lane,
                                              int* rowbase = nullptr, int4* panelmeta2 = nullptr) {
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int ES = (int)sizeof(U);
    const int NST = p.nst;
    const unsigned char* gsrc = p.packed;
    if (lane == 0) {
        const int npf = min(NP, NST + p.l2_ahead);
        for (int v = NST; v < npf; ++v) {
            const int64_t o0 = p.prow[p.panel_row[pan0 + v]], o1 = p.prow[p.panel_row[pan0 + v + 1]];
            if (o1 > o0) tma_prefetch_l2(gsrc + o0 * ES, (uint32_t)((o1 - o0) * ES));
        }
    }
    int s = 0, k = 0;
    for (int v = 0; v < NP; ++v) {
        const int rs = p.panel_row[pan0 + v], re = p.panel_row[pan0 + v + 1];
        const int P = re - rs;
        const int64_t obase = p.prow[rs];
        const int64_t oend = p.prow[re];
        int64_t o0 = 0, o1 = 0;
        int c = 0;
        if (lane < P) { o0 = p.prow[rs + lane]; o1 = p.prow[rs + lane + 1]; c = p.pcs[rs + lane] - r0; }
        const int need_next = (v + 1 < NP) ? p.panel_need[pan0 + v + 1] : 0;   // what the chain needs before panel v+1
        // everything that does not touch shared memory comes before the wait: the TMA is issued a few
        // instructions after the stage is released
        int vs = 0x7fffffff, ve = 0;
        int lo = 0, hi = 0x7fffffff;                 // vectors [lo, hi) are stored by EVERY row of the panel
        int4 m = make_int4(0, 0, 0, 0);
        if (lane < P) {
            const int nv = (int)(o1 - o0) / EPV;
            const int vs_r = c / EPV;
            m.x = (int)p.L.stages + s * p.stage_bytes + (int)((o0 - obase) * ES) - vs_r * 16;
            m.y = vs_r; m.z = vs_r + nv; m.w = need_next;
            if (nv > 0) { vs = vs_r; ve = vs_r + nv; }
            lo = vs_r; hi = vs_r + nv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vs = min(vs, __shfl_xor_sync(0xffffffffu, vs, o));
            ve = max(ve, __shfl_xor_sync(0xffffffffu, ve, o));
            lo = max(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = min(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        const uint32_t bytes = (uint32_t)((oend - obase) * ES);
        trace_ev(p, lane, 9, 0, v);
        if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
        trace_ev(p, lane, 9, 1, v);
        if (lane < P) {
            rowmeta[(rs - r0 + lane) & (RR - 1)] = m;
            if (rowbase != nullptr) rowbase[(rs - r0 + lane) & (RR - 1)] = m.x;
        }
        if (lane == 0) {
            panelmeta[s] = make_int4(P, ve > 0 ? vs : 0, ve, rs - r0);
            if (panelmeta2 != nullptr) panelmeta2[s] = make_int4(lo, hi, 0, 0);
        }
        __syncwarp();
        if (lane == 0) {
            if (bytes > 0) {
                mbar_arrive_expect_tx(&full[s], bytes);
                tma_load_1d(smem + p.L.stages + (size_t)s * p.stage_bytes, gsrc + obase * ES, bytes, &full[s]);
            } else {
                mbar_arrive(&full[s]);
            }
            trace_ev(p, lane, 9, 2, v);
            const int vp = v + NST + p.l2_ahead;
            if (vp < NP) {
                const int64_t q0 = p.prow[p.panel_row[pan0 + vp]], q1 = p.prow[p.panel_row[pan0 + vp + 1]];
                if (q1 > q0) tma_prefetch_l2(gsrc + q0 * ES, (uint32_t)((q1 - q0) * ES));
            }
        }
        if (++s == NST) { s = 0; ++k; }
    }
}

// NA / NC: number of bulk warps that publish a_prog / c_prog (and dot partials)
// NW: warps that only publish a progress counter the chain waits on like an A counter (prog layout: A | W | C)
// INCR: the reference's own bookkeeping of q (e_step.hpp:421, 435-440) -- the forward axpys carry eta_diff instead of
// eta_new and the caller's q enters through fext, so that X = q_in / dq + sum_{i<j} R_ij eta_diff_i with no backward
// dots at all; the backward part is added to q by a second pass over the LD (update_q_factor) after the sweep.
template <typename T, typename Model, int NA, int NC, int NW = 0, typename W = T, bool INCR = false>
__device__ __forceinline__ void chain_role(const SweepPlan& p, const typename Model::Args& ma, const StateArgs<T>& sa,
                                           const SmemView<T>& sm, int r0, int B, int pan0, int NP, int lane) {
    const int NST = p.nst;
    const T eps = eps_of<T>();
    const T dq = ma.dq;
    const uint32_t a_partial = smem_u32(sm.partial), a_wwin = smem_u32(sm.wwin), a_alpha = smem_u32(sm.alpha),
                   a_f = smem_u32(sm.fsrc);
    typename Model::Lane L;
    typename Model::Raw pend;                       // parameters of the lane's next column, in flight
    bool has_pend = false;
    T eo_pend = T(0);
    {
        typename Model::Raw r;
        Model::load_raw(ma, r0 + lane, lane < B, r);
        Model::derive(ma, r, L);
        Model::load_raw(ma, r0 + lane, false, pend);
    }
    T eo = (lane < B) ? sa.eta[r0 + lane] : T(0);
    // external contributions of the lane's current / next column (tiled LD blocks, q offsets; see StateArgs)
    const bool has_f = sa.fext != nullptr, has_b = sa.bext != nullptr;
    auto load_fext = [&](int col) { return (has_f && col < B) ? mul_t(sa.fext[r0 + col], sa.fscale) : T(0); };
    auto load_bext = [&](int col) { return (has_b && col < B) ? sa.bext[r0 + col] : T(0); };
    T xf = load_fext(lane), xb = load_bext(lane), xf_pend = T(0), xb_pend = T(0);
    T X0 = T(0), X1 = T(0);
    const uint32_t a_rowmeta = smem_u32(sm.rowmeta), a_panelmeta = smem_u32(sm.panelmeta);
    int j0 = 0, s = 0;
    int need_c = 0;                                  // the first panel has no predecessors
    for (int u = 0; u < NP; ++u) {
        // one batch = one row panel (1..16 rows): rows [j0, j0 + nrows) of the block
        trace_ev(p, lane, 8, 0, u);
        const int pidx = sm.by_panel ? (u & (PMR - 1)) : s;
        if (sm.by_panel) need_c = max(need_c, u - (PMR - 2));       // keeps the panel-indexed cdone ring unambiguous
        wait_progress<NA + NW, NC>(sm.prog, (uint32_t)(u + 1), (uint32_t)need_c, lane);
        trace_ev(p, lane, 8, 1, u);
        const int nrows = (int)lds128(a_panelmeta + pidx * 16).x;
        need_c = (int)lds128(a_rowmeta + (j0 & (RR - 1)) * 16).w;
        const int base = j0 & 31;
        const int rel = (lane - base) & 31;

        // fold what the bulk warps prepared for this panel's columns
        T bsum = T(0);
        if (rel < nrows) {
            const int cl = j0 + rel;
#pragma unroll
            for (int w = 0; w < NA; ++w) bsum += lds_t(a_partial + (uint32_t)(w * RR + (cl & (RR - 1))) * sizeof(T), T());
            bsum += xb;                                          // backward part: in-tile dots + later tiles
            X0 += lds_t(a_f + (uint32_t)(cl & sm.fmask) * sizeof(T), T()) + xf + bsum;
        }
        T Xown = T(0);
#pragma unroll kChainUnroll
        for (int h = 0; h < PMAX; h += 4) {
            if (h < nrows) {
                T w0[4], w1[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k0 = (rel - (h + i) - 1) & 31;       // window slot of the lane's X0 column at this step
                    const uint32_t wr = a_wwin + (uint32_t)((((j0 + h + i) & (RR - 1)) * WW + k0) * sizeof(W));
                    w0[i] = T(0); w1[i] = T(0);
                    if constexpr (std::is_same<W, T>::value) {
                        if (h + i < nrows) w0[i] = lds_t(wr, T());
                        if (h + i < nrows && k0 + 32 < WW) w1[i] = lds_t(wr + 32 * sizeof(T), T());
                    } else {
                        if (h + i < nrows) w0[i] = lds_w<W>(wr);
                        if (h + i < nrows && k0 + 32 < WW) w1[i] = lds_w<W>(wr + 32 * sizeof(W));
                    }
                }
                // one SNP update; lanes other than the row's owner compute with a stale X0 and their result is unused
                auto one_step = [&](int i) {
                    T en, d;
                    bool skip;
                    typename Model::Out o;
                    Model::step(L, X0, eo, eps, en, d, skip, o);
                    const bool mine = (rel == h + i);
                    Xown = mine ? X0 : Xown;
                    const T x0n = mine ? X1 : X0;          // the window slides by one column: independent of the shuffle
                    const T x1n = mine ? T(0) : X1;
                    const T a = shfl_t(INCR ? (skip ? T(0) : d) : en, (base + h + i) & 31);
                    X0 = fma_t(w0[i], a, x0n);             // :421 restricted to the window
                    X1 = fma_t(w1[i], a, x1n);
                };
                if (Model::kDualGroups && h + 4 <= nrows) {   // full group: no per-step branches on the serial path
#pragma unroll
                    for (int i = 0; i < 4; ++i) one_step(i);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (h + i < nrows) one_step(i);
                }
            }
        }
        trace_ev(p, lane, 8, 2, u);
        // panel epilogue: outputs of the rows (re-evaluated from the saved X: same bits as on the chain),
        // eta_new for the bulk axpy, parameters of the lanes' next columns
        if (rel < nrows) {
            T en, d;
            bool skip;
            typename Model::Out o;
            Model::step(L, Xown, eo, eps, en, d, skip, o);
            const int cl = j0 + rel;
            const int row = r0 + cl;
            Model::store(ma, row, skip, o);
            if (!skip) sa.eta[row] = en;                                       // :431
            sa.eta_diff[row] = skip ? T(0) : d;                                // :413 / :418
            sa.q[row] = dq * (Xown - bsum);                                    // forward part of q (see header)
            sts_t(a_alpha + (uint32_t)(cl & (RR - 1)) * sizeof(T), INCR ? (skip ? T(0) : d) : en);
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&sm.cdone[pidx]);
            if (sm.chain_rows != nullptr) st_release(sm.chain_rows, (uint32_t)(j0 + nrows));
        }
        trace_ev(p, lane, 8, 3, u);
        if (has_pend) { Model::derive(ma, pend, L); eo = eo_pend; xf = xf_pend; xb = xb_pend; has_pend = false; }
        if (rel < nrows) {
            const int cn = j0 + rel + 32;
            Model::load_raw(ma, r0 + cn, cn < B, pend);
            eo_pend = (cn < B) ? sa.eta[r0 + cn] : T(0);                       // not yet rewritten: cn is >= 16 rows ahead
            xf_pend = load_fext(cn); xb_pend = load_bext(cn);
            has_pend = true;
        }
        j0 += nrows;
        if (++s == NST) s = 0;
    }
}

// window coefficients of row jl (block-local) of a landed panel -> wwin ring (one warp per row)
template <typename T, typename U>
__device__ __forceinline__ void window_row(unsigned char* smem, const int4* rowmeta, T* wwin, int jl, int lane) {
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int ES = (int)sizeof(U);
    const int4 m = rowmeta[jl & (RR - 1)];
    const int cut = ((jl + WIN + EPV - 1) / EPV) * EPV;
#pragma unroll
    for (int kk = lane; kk < WW; kk += WARP) {
        const int col = jl + 1 + kk;
        T v = T(0);
        if (col < cut && col / EPV >= m.y && col / EPV < m.z) v = ld_elem<T, U>(smem + m.x + col * ES);
        wwin[(jl & (RR - 1)) * WW + kk] = v;
    }
}

// Window coefficients of a whole landed panel, split over NP_ warps (part = 0..NP_-1).  Branch-free straight-line
// code with independent loads (an element outside the row's stored range or beyond the cut reads the "code 0"
// constant at zaddr): a warp-item is either (row r, kk = lane) -- P of those -- or (rows 2 i + (lane >> 4),
// kk = 32 + (lane & 15)) -- ceil(P / 2) of those; part `part` takes items part, part + NP_, ...
template <typename U> __device__ __forceinline__ float lds_code(uint32_t addr);
template <> __device__ __forceinline__ float lds_code<int8_t>(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return (float)((int)v - 128);
}
template <> __device__ __forceinline__ float lds_code<int16_t>(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return (float)((int)v - 32768);
}
template <> __device__ __forceinline__ float lds_code<float>(uint32_t addr) { return lds_t(addr, float()); }

template <typename U, int NP_, typename W = float>
__device__ __forceinline__ void window_panel(uint32_t sbase, uint32_t a_rowmeta, uint32_t a_wwin, uint32_t zaddr,
                                             int jl0, int P, int part, int lane) {
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int ES = (int)sizeof(U);
    constexpr int NIT = (PMAX + PMAX / 2 + NP_ - 1) / NP_;
    constexpr int NB = 2;                                  // items in flight per batch
    static_assert(WW == 48, "window_panel assumes 32 + 16 coefficients per row");
    const int nit = P + (P + 1) / 2;
#pragma unroll 1
    for (int q0 = 0; q0 < NIT; q0 += NB) {
        if (part + q0 * NP_ >= nit) break;                 // warp-uniform
        int jl[NB], kk[NB];
        bool ok[NB];
        uint4 m[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int i = part + (q0 + q) * NP_;
            const bool first = i < P;
            const int r = first ? i : 2 * (i - P) + (lane >> 4);
            kk[q] = first ? lane : 32 + (lane & 15);
            ok[q] = (i < nit) & (r < P);
            jl[q] = jl0 + (ok[q] ? r : 0);
            m[q] = lds128(a_rowmeta + (uint32_t)(jl[q] & (RR - 1)) * 16u);
        }
        float v[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            // block-local indices are non-negative: unsigned shifts, and plain & instead of && (no lazy-evaluation branches)
            const uint32_t cut = (((uint32_t)jl[q] + WIN + EPV - 1) / EPV) * EPV;
            const uint32_t col = (uint32_t)jl[q] + 1u + (uint32_t)kk[q];
            const uint32_t cv = col / EPV;
            const bool in = ok[q] & (col < cut) & (cv >= m[q].y) & (cv < m[q].z);
            v[q] = lds_code<U>(in ? sbase + m[q].x + col * ES : zaddr);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q)
            if (ok[q]) sts_w<W>(a_wwin + (uint32_t)((jl[q] & (RR - 1)) * WW + kk[q]) * (uint32_t)sizeof(W), v[q]);
    }
}

// int8 LD: the same job four coefficients at a time.  Lane l < 24 of part `part` takes word w = l % 12 (columns
// j + 1 + 4 w .. + 3) of row 2 part + l / 12 (+ 8 for the second half of a 16-row panel): two aligned 32-bit loads
// and a byte funnel fetch the four codes, one 128-bit store writes the four coefficients.
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int NP_, typename W = float>
__device__ __forceinline__ void window_panel_i8(uint32_t sbase, uint32_t a_rowmeta, uint32_t a_wwin, int jl0, int P,
                                                int part, int lane) {
    static_assert(WW == 48 && NP_ == 4 && PMAX == 16, "window_panel_i8: 12 words per row, 2 rows per part and pass");
    const uint32_t w = (uint32_t)lane % 12u, rsub = (uint32_t)lane / 12u;
#pragma unroll 1
    for (uint32_t it = 0; it < 2; ++it) {
        if (8u * it + 2u * (uint32_t)part >= (uint32_t)P) break;            // warp-uniform
        const uint32_t r = 2u * (uint32_t)part + rsub + 8u * it;
        const bool ok = (lane < 24) & (r < (uint32_t)P);
        const uint32_t jl = (uint32_t)jl0 + (ok ? r : 0u);
        const uint4 m = lds128(a_rowmeta + (jl & (RR - 1)) * 16u);
        const uint32_t a = sbase + m.x + jl + 1u + 4u * w;                  // byte address of the first of the four codes
        const uint32_t lo = lds_u32(a & ~3u);
        const uint32_t hi = lds_u32((a & ~3u) + 4u);
        const uint32_t word = __byte_perm(lo, hi, 0x3210u + 0x1111u * (a & 3u));
        const uint32_t col0 = jl + 1u + 4u * w;
        const uint32_t cut = ((jl + WIN + 15u) / 16u) * 16u;
        const uint32_t lim = min(cut, m.z * 16u), beg = m.y * 16u;          // stored and chain-owned: [beg, lim)
        if constexpr (std::is_same<W, float>::value) {
            float2 p0, p1;
            VecOps<float, int8_t>::pairs(word, p0, p1);
            uint4 o;
            o.x = (col0 >= beg) & (col0 < lim) ? __float_as_uint(p0.x) : 0u;
            o.y = (col0 + 1u >= beg) & (col0 + 1u < lim) ? __float_as_uint(p0.y) : 0u;
            o.z = (col0 + 2u >= beg) & (col0 + 2u < lim) ? __float_as_uint(p1.x) : 0u;
            o.w = (col0 + 3u >= beg) & (col0 + 3u < lim) ? __float_as_uint(p1.y) : 0u;
            if (ok) sts128(a_wwin + ((jl & (RR - 1)) * WW + 4u * w) * 4u, o);
        } else {
            // int16 storage: code = biased byte - 128, two codes per 32-bit word, one 64-bit store
            uint32_t c[4];
#pragma unroll
            for (uint32_t e = 0; e < 4; ++e) {
                const bool in = (col0 + e >= beg) & (col0 + e < lim);
                c[e] = in ? (((word >> (8u * e)) & 0xffu) - 128u) & 0xffffu : 0u;
            }
            const uint32_t o0 = c[0] | (c[1] << 16), o1 = c[2] | (c[3] << 16);
            if (ok) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a_wwin + ((jl & (RR - 1)) * WW + 4u * w) * 2u), "r"(o0), "r"(o1) : "memory");
        }
    }
}

// =============================================================================================
// Version 2 of the chain / helper roles of the register-resident kernel (sweep_fast.cuh, VER == 2)
// =============================================================================================
// Window coefficients, layout 2: wwin2[(row & 63) * 64 + (col & 63)] = R[row][col] for row < col < cut_row (and inside
// the row's stored run), 0 for every other slot.  Row jl only reaches chain-owned columns < cut_jl <= 32 (jl / 32) + 64,
// i.e. columns of its own 32-column block and of the next one, so the chain warp keeps exactly two accumulators per
// lane (X0: its column of the current block, X1: of the next block), switches them at 32-row boundaries (warp-uniform)
// and reads its two coefficients at addresses that advance by one constant stride per row -- no per-step address
// arithmetic, no per-lane window sliding.
constexpr int W2 = 64;            // slots per row in layout 2

// int8 LD: 16 lanes per row, one aligned 32-bit load + one 128-bit store each; two rows per warp and pass.
template <int NP_>
__device__ __forceinline__ void window_panel2_i8(uint32_t sbase, uint32_t a_rowmeta, uint32_t a_wwin, int jl0, int P,
                                                 int part, int lane) {
    static_assert(NP_ == 4 && PMAX == 16, "window_panel2_i8: 2 rows per part and pass");
    const uint32_t sw = (uint32_t)lane & 15u, rsub = (uint32_t)lane >> 4;
#pragma unroll 1
    for (uint32_t it = 0; it < 2; ++it) {
        if (8u * it + 2u * (uint32_t)part >= (uint32_t)P) break;            // warp-uniform
        const uint32_t r = 2u * (uint32_t)part + rsub + 8u * it;
        const bool ok = r < (uint32_t)P;
        const uint32_t jl = (uint32_t)jl0 + (ok ? r : 0u);
        const uint4 m = lds128(a_rowmeta + (jl & (RR - 1)) * 16u);
        const uint32_t jl4 = jl & ~3u;
        const uint32_t cb = jl4 + ((4u * sw - jl4) & 63u);                  // the column in [jl4, jl4 + 64) with slot 4 sw
        const uint32_t word = lds_u32(sbase + m.x + cb);
        const uint32_t cut = ((jl + WIN + 15u) / 16u) * 16u;
        const uint32_t lim = min(cut, m.z * 16u);                           // chain-owned and stored: (jl, lim)
        float2 p0, p1;
        VecOps<float, int8_t>::pairs(word, p0, p1);
        uint4 o;
        o.x = (cb > jl) & (cb < lim) ? __float_as_uint(p0.x) : 0u;
        o.y = (cb + 1u > jl) & (cb + 1u < lim) ? __float_as_uint(p0.y) : 0u;
        o.z = (cb + 2u > jl) & (cb + 2u < lim) ? __float_as_uint(p1.x) : 0u;
        o.w = (cb + 3u > jl) & (cb + 3u < lim) ? __float_as_uint(p1.y) : 0u;
        if (ok) sts128(a_wwin + ((jl & (RR - 1)) * W2 + 4u * sw) * 4u, o);
    }
}

// any LD type: lane l writes slots l and l + 32 of one row; rows part, part + NP_, ... of the panel
template <typename U, int NP_>
__device__ __forceinline__ void window_panel2(uint32_t sbase, uint32_t a_rowmeta, uint32_t a_wwin, uint32_t zaddr,
                                              int jl0, int P, int part, int lane) {
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int ES = (int)sizeof(U);
#pragma unroll 1
    for (int r = part; r < P; r += NP_) {
        const uint32_t jl = (uint32_t)(jl0 + r);
        const uint4 m = lds128(a_rowmeta + (jl & (RR - 1)) * 16u);
        const uint32_t cut = ((jl + WIN + EPV - 1) / EPV) * EPV;
        const uint32_t lim = min(cut, m.z * (uint32_t)EPV);
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t slot = (uint32_t)lane + 32u * h;
            const uint32_t col = jl + ((slot - jl) & 63u);                  // the column in [jl, jl + 64) with this slot
            const bool in = (col > jl) & (col < lim);
            v[h] = lds_code<U>(in ? sbase + m.x + col * ES : zaddr);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) sts_t(a_wwin + ((jl & (RR - 1)) * W2 + (uint32_t)lane + 32u * h) * 4u, v[h]);
    }
}

// shared-memory hand-off from the chain to the output role (rings indexed by block-local row & 63)
struct OutRings {
    uint32_t a_xown, a_bsum;      // X the row's update was computed from; backward part (in-tile dots + later tiles)
    uint32_t* chain_panels;       // panels the chain has finished (release / acquire)
    uint32_t* out_rows;           // rows whose outputs have been written (release / acquire)
};

template <typename T, typename Model, int NA, int NC>
__device__ __forceinline__ void chain_role2(const SweepPlan& p, const typename Model::Args& ma, const StateArgs<T>& sa,
                                            const SmemView<T>& sm, const OutRings& orr, int r0, int B, int pan0, int NP,
                                            int lane) {
    const int NST = p.nst;
    const T eps = eps_of<T>();
    const uint32_t a_partial = smem_u32(sm.partial), a_wwin = smem_u32(sm.wwin), a_alpha = smem_u32(sm.alpha),
                   a_f = smem_u32(sm.fsrc);
    typename Model::Lane L;
    typename Model::Raw pend;
    bool has_pend = false;
    T eo_pend = T(0);
    {
        typename Model::Raw r;
        Model::load_raw(ma, r0 + lane, lane < B, r);
        Model::derive(ma, r, L);
        Model::load_raw(ma, r0 + lane, false, pend);
    }
    T eo = (lane < B) ? sa.eta[r0 + lane] : T(0);
    const bool has_f = sa.fext != nullptr, has_b = sa.bext != nullptr;
    auto load_fext = [&](int col) { return (has_f && col < B) ? mul_t(sa.fext[r0 + col], sa.fscale) : T(0); };
    auto load_bext = [&](int col) { return (has_b && col < B) ? sa.bext[r0 + col] : T(0); };
    T xf = load_fext(lane), xb = load_bext(lane), xf_pend = T(0), xb_pend = T(0);
    T X0 = T(0), X1 = T(0);
    const uint32_t a_rowmeta = smem_u32(sm.rowmeta), a_panelmeta = smem_u32(sm.panelmeta);
    int j0 = 0, s = 0;
    int need_c = 0;
    for (int u = 0; u < NP; ++u) {
        trace_ev(p, lane, 8, 0, u);
        wait_progress<NA, NC>(sm.prog, (uint32_t)(u + 1), (uint32_t)need_c, lane);
        const int nrows = (int)lds128(a_panelmeta + s * 16).x;
        need_c = (int)lds128(a_rowmeta + (j0 & (RR - 1)) * 16).w;
        if (j0 + nrows > RR) wait_ge(orr.out_rows, (uint32_t)(j0 + nrows - RR));   // the output rings are RR rows deep
        trace_ev(p, lane, 8, 1, u);
        const int base = j0 & 31;
        const int rel = (lane - base) & 31;
        if (base == 0 && j0 > 0) { X0 = X1; X1 = T(0); }          // the panel opens the next 32-column block (warp-uniform)
        T bsum = T(0);
        if (rel < nrows) {
            const int cl = j0 + rel;
#pragma unroll
            for (int w = 0; w < NA; ++w) bsum += lds_t(a_partial + (uint32_t)(w * RR + (cl & (RR - 1))) * sizeof(T), T());
            bsum += xb;
            const T v = lds_t(a_f + (uint32_t)(cl & sm.fmask) * sizeof(T), T()) + xf + bsum;
            if ((base + rel) & 32) X1 += v; else X0 += v;        // a panel may straddle a 32-row boundary
        }
        T Xown = T(0);
        // one SNP update; lanes other than the row's owner compute with a stale X0 and their result is unused
        auto one_step = [&](int j, T w0, T w1) {
            T en, d;
            bool skip;
            typename Model::Out o;
            Model::step(L, X0, eo, eps, en, d, skip, o);
            const int src = j & 31;
            const bool mine = lane == src;
            Xown = mine ? X0 : Xown;
            const T a = shfl_t(en, src);
            if (mine) sts_t(a_alpha + (uint32_t)(j & (RR - 1)) * sizeof(T), en);      // eta_new for the C warps
            X0 = fma_t(w0, a, X0);                     // :421 restricted to the window
            X1 = fma_t(w1, a, X1);
        };
        auto waddr = [&](int j) {                      // address of the lane's current-block coefficient of row j
            return a_wwin + (uint32_t)(j & (RR - 1)) * (W2 * 4u) + (uint32_t)lane * 4u + ((j & 32) ? 128u : 0u);
        };
        if ((j0 & 3) == 0) {
            // panels are cut at multiples of 4 rows wherever the rows are short enough (ld.cu): groups of 4 steps never
            // straddle a 32-row boundary, the coefficients of a group are fetched up front at constant offsets
#pragma unroll 1
            for (int h = 0; h < nrows; h += 4) {
                const int jg = j0 + h;
                if ((jg & 31) == 0 && h > 0) { X0 = X1; X1 = T(0); }             // a 32-row boundary inside the panel
                const uint32_t g0 = waddr(jg);
                const uint32_t g1 = (jg & 32) ? g0 - 128u : g0 + 128u;
                const bool full = h + 4 <= nrows;
                T w0[4], w1[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w0[i] = T(0); w1[i] = T(0);
                    if (full || h + i < nrows) {
                        w0[i] = lds_t(g0 + (uint32_t)i * (W2 * 4u), T());
                        w1[i] = lds_t(g1 + (uint32_t)i * (W2 * 4u), T());
                    }
                }
                if (full) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) one_step(jg + i, w0[i], w1[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (h + i < nrows) one_step(jg + i, w0[i], w1[i]);
                }
            }
        } else {
            // long rows (a stage holds fewer than four of them): panels of 1..3 rows start anywhere
#pragma unroll 1
            for (int h = 0; h < nrows; ++h) {
                const int j = j0 + h;
                if ((j & 31) == 0 && h > 0) { X0 = X1; X1 = T(0); }
                const uint32_t g0 = waddr(j);
                const uint32_t g1 = (j & 32) ? g0 - 128u : g0 + 128u;
                one_step(j, lds_t(g0, T()), lds_t(g1, T()));
            }
        }
        trace_ev(p, lane, 8, 2, u);
        if (rel < nrows) {
            const uint32_t slot = (uint32_t)((j0 + rel) & (RR - 1)) * sizeof(T);
            sts_t(orr.a_xown + slot, Xown);
            sts_t(orr.a_bsum + slot, bsum);
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&sm.cdone[s]);
            st_release(orr.chain_panels, (uint32_t)(u + 1));
        }
        trace_ev(p, lane, 8, 3, u);
        if (has_pend) { Model::derive(ma, pend, L); eo = eo_pend; xf = xf_pend; xb = xb_pend; has_pend = false; }
        if (rel < nrows) {
            const int cn = j0 + rel + 32;
            Model::load_raw(ma, r0 + cn, cn < B, pend);
            eo_pend = (cn < B) ? sa.eta[r0 + cn] : T(0);      // written by the output role only after the chain is past cn
            xf_pend = load_fext(cn); xb_pend = load_bext(cn);
            has_pend = true;
        }
        j0 += nrows;
        if (++s == NST) s = 0;
    }
}

// Output role of version 2 (its own warp): writes the outputs of the panels the chain has finished.  The update of every
// row is re-evaluated from the X the chain saved (same function, same inputs: same bits) and var_mu / var_gamma / eta /
// eta_diff / q go to global memory from here, off the chain's serial path.  With sa.unit_partial set it also accumulates
// the M-step / ELBO sums of the sweep unit (the per-SNP terms of viprs_b200_sums_*, em.cu) from the values it writes.
template <typename T, typename Model>
__device__ __forceinline__ void output_role(const SweepPlan& p, const typename Model::Args& ma, const StateArgs<T>& sa,
                                            const OutRings& orr, int r0, int pan0, int NP, int lane, int unit) {
    const T eps = eps_of<T>();
    const T dq = ma.dq;
    [[maybe_unused]] const bool fuse = Model::kFusedSums && sa.unit_partial != nullptr;
    [[maybe_unused]] double acc[NS];
    [[maybe_unused]] double on = 0.0, on_next = 0.0, nscale = 0.0, tau_b = 0.0;
    if constexpr (Model::kFusedSums) {
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[i] = 0.0;
        if (fuse) { const Theta th = sa.theta[0]; nscale = (1.0 + th.lambda_min) / th.sigma_epsilon; tau_b = th.tau_beta; }
    }
    typename Model::Raw raw, raw_next;
    T eo = T(0), eo_next = T(0);
    int rs = 0, re = 0, rs_next = 0, re_next = 0;
    auto prefetch = [&](int u) {                       // inputs of panel u (not yet rewritten: only this warp writes them)
        if (u >= NP) return;
        rs_next = p.panel_row[pan0 + u]; re_next = p.panel_row[pan0 + u + 1];
        const bool ok = lane < re_next - rs_next;
        Model::load_raw(ma, rs_next + lane, ok, raw_next);
        eo_next = ok ? sa.eta[rs_next + lane] : T(0);
        if constexpr (Model::kFusedSums) on_next = (fuse && ok) ? sa.n_per_snp[rs_next + lane] : 0.0;
    };
    prefetch(0);
    for (int u = 0; u < NP; ++u) {
        raw = raw_next; eo = eo_next; rs = rs_next; re = re_next; on = on_next;
        prefetch(u + 1);
        uint32_t spins = 0;
        while (ld_acquire(orr.chain_panels) < (uint32_t)(u + 1)) {
            __nanosleep(256);
            if (++spins > kSpinLimit) __trap();
        }
        if (lane < re - rs) {
            typename Model::Lane L;
            Model::derive(ma, raw, L);
            const uint32_t slot = (uint32_t)((rs - r0 + lane) & (RR - 1)) * sizeof(T);
            const T Xown = lds_t(orr.a_xown + slot, T());
            const T bsum = lds_t(orr.a_bsum + slot, T());
            T en, d;
            bool skip;
            typename Model::Out o;
            Model::step(L, Xown, eo, eps, en, d, skip, o);
            const int row = rs + lane;
            Model::store(ma, row, skip, o);
            if (!skip) sa.eta[row] = en;                                       // :431
            sa.eta_diff[row] = skip ? T(0) : d;                                // :413 / :418
            const T qf = dq * (Xown - bsum);
            sa.q[row] = qf;                                                    // forward part of q
            if constexpr (Model::kFusedSums) {
                if (fuse) {
                    // a skipped update leaves var_gamma / var_mu as they were (e_step.hpp:410-413): read them back
                    const double g = skip ? (double)ma.var_gamma[row] : (double)o.g;
                    const double mu = skip ? (double)ma.var_mu[row] : (double)o.mu;
                    const double vt = on * nscale + tau_b;
                    const double gc = clip_res(g), ivt = rcp_em<T>(vt), et = (double)en;
                    const double ng = clip_res(1.0 - g);
                    acc[VIPRS_B200_S_GAMMA] += g;
                    acc[VIPRS_B200_S_GAMMA_MU2] += g * mu * mu;
                    acc[VIPRS_B200_S_G_INV_TAU] += g * ivt;
                    acc[VIPRS_B200_S_G_LOGG] += gc * log_unit<T>(gc);
                    acc[VIPRS_B200_S_GCLIP] += gc;
                    acc[VIPRS_B200_S_G_LOG_TAU] += gc * log_tau<T>(vt);
                    acc[VIPRS_B200_S_GC_ZETA] += gc * (mu * mu + ivt);
                    acc[VIPRS_B200_S_ETA_Q] += 2.0 * et * (double)qf;          // eta'(R - I)eta = 2 sum eta_j F_j
                    acc[VIPRS_B200_S_BETA_ETA] += (double)raw.beta * et;
                    acc[VIPRS_B200_S_NG_LOGNG] += ng * log_one_minus<T>(g, ng);
                    acc[VIPRS_B200_S_NGCLIP] += ng;
                    acc[VIPRS_B200_S_ETA2] += et * et;
                    acc[VIPRS_B200_S_MAX_DIFF] = fmax(acc[VIPRS_B200_S_MAX_DIFF], fabs(skip ? 0.0 : (double)d));
                }
            }
        }
        __syncwarp();
        if (lane == 0) st_release(orr.out_rows, (uint32_t)(re - r0));
    }
    if constexpr (Model::kFusedSums) {
        if (fuse) {
            // lanes by xor-shuffle (fixed order), one row of the unit table per sweep unit
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                double v = acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double w = __shfl_xor_sync(0xffffffffu, v, o);
                    v = (i == VIPRS_B200_S_MAX_DIFF) ? fmax(v, w) : v + w;
                }
                if (lane == 0) sa.unit_partial[(size_t)unit * NS + i] = v;
            }
        }
    }
}

// sums[seg][slot] = sum (max for MAX_DIFF) over the sweep units whose rows lie in chromosome segment seg, in unit order.
// grid = nseg, block = 32 * NS threads is more than needed: one warp per slot, lanes stride the units.
template <typename T>        // (a template only so that every translation unit may carry its own copy)
__global__ void reduce_units_kernel(int n_units, const int32_t* __restrict__ unit_row, const int32_t* __restrict__ seg_ptr,
                                    const double* __restrict__ unit_partial, double* __restrict__ sums) {
    const int seg = blockIdx.x, slot = threadIdx.x / WARP, lane = threadIdx.x % WARP;
    if (slot >= NS) return;
    const int row0 = seg_ptr[seg], row1 = seg_ptr[seg + 1];
    const bool is_max = slot == VIPRS_B200_S_MAX_DIFF;
    double v = 0.0;
    for (int u = lane; u < n_units; u += WARP) {
        const int r = unit_row[u];
        if (r >= row0 && r < row1) {
            const double w = unit_partial[(size_t)u * NS + slot];
            v = is_max ? fmax(v, w) : v + w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, w) : v + w;
    }
    if (lane == 0) sums[(size_t)seg * NS + slot] = v;
}

// ---------------------------------------------------------------------------------------------
// the generic sweep kernel: any state type, any block size that fits; block state in shared memory
// ---------------------------------------------------------------------------------------------
template <typename T, typename U, typename Model, int MINB>
__global__ void __launch_bounds__((NBW + 2) * WARP, MINB) sweep_kernel(const SweepPlan p, const typename Model::Args ma,
                                                                       const StateArgs<T> sa) {
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int TPV = EPV * (int)sizeof(T) / 16;       // 16-byte chunks of state per LD vector
    extern __shared__ __align__(128) unsigned char smem[];

    T* eta_s = reinterpret_cast<T*>(smem + p.L.eta);
    T* f_s = reinterpret_cast<T*>(smem + p.L.f);
    SmemView<T> sm;
    sm.base = smem;
    sm.rowmeta = reinterpret_cast<int4*>(smem + p.L.rowmeta);
    sm.panelmeta = reinterpret_cast<int4*>(smem + p.L.panelmeta);
    sm.partial = reinterpret_cast<T*>(smem + p.L.partial);
    sm.alpha = reinterpret_cast<T*>(smem + p.L.alpha);
    sm.wwin = reinterpret_cast<T*>(smem + p.L.wwin);
    sm.full = reinterpret_cast<uint64_t*>(smem + p.L.bars);
    sm.empty = sm.full + NST_MAX;
    sm.cdone = sm.full + 2 * NST_MAX;
    sm.prog = reinterpret_cast<uint32_t*>(smem + p.L.counters);
    sm.fsrc = f_s;
    sm.fmask = 0x7fffffff;
    int4* rowmeta = sm.rowmeta;
    int4* panelmeta = sm.panelmeta;
    T* partial = sm.partial;
    T* alpha = sm.alpha;
    uint64_t* full = sm.full;
    uint64_t* empty = sm.empty;
    uint64_t* cdone = sm.cdone;
    uint32_t* prog = sm.prog;

    const int tid = threadIdx.x, warp = tid / WARP, lane = tid % WARP;
    const int blk = p.blk_order[blockIdx.x];
    const int r0 = p.blk_row[blk], r1 = p.blk_row[blk + 1];
    const int B = r1 - r0;
    const int pan0 = p.blk_panel[blk];
    const int NP = p.blk_panel[blk + 1] - pan0;
    const int NST = p.nst;

    // ---- prologue: state into shared memory, barriers ---------------------------------------
    for (int i = tid; i < p.bpad; i += blockDim.x) {
        eta_s[i] = (i < B) ? sa.eta[r0 + i] : T(0);
        f_s[i] = T(0);
    }
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NBW); mbar_init(&cdone[s], 1); }
        for (int w = 0; w < 2 * NBW; ++w) prog[w] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NBW) {
        producer_role<U>(p, smem, rowmeta, panelmeta, full, empty, r0, pan0, NP, lane);
    } else if (warp == NBW + 1) {
        chain_role<T, Model, NBW, NBW>(p, ma, sa, sm, r0, B, pan0, NP, lane);
    } else {
        // =============================== bulk ===============================================
        const int wb = warp;
        const int tb = wb * WARP + lane;
        int cnext = 0, cs = 0, ck = 0;                     // next panel whose axpy (C) this warp has to do

        auto do_C = [&]() {
            trace_ev(p, lane, wb, 4, cnext);
            const int4 pm = panelmeta[cs];
            const int Pc = pm.x, vmaxc = pm.z, jl0 = pm.w;
            const int first = (jl0 + WIN + EPV - 1) / EPV;
            const int last_cut = (jl0 + Pc - 1 + WIN + EPV - 1) / EPV;     // cut vector of the panel's last row
            int vv = first + (((tb - first) % NBT) + NBT) % NBT;           // static ownership: vv == tb (mod NBT)
            for (; vv < vmaxc; vv += NBT) {
                T fs[EPV];
                T* fp = f_s + (size_t)vv * EPV;
                load_state_vec(fp, fs, TPV);
                for (int rg = 0; rg < Pc; rg += 4) {
                    const int nv = min(4, Pc - rg);
                    int mx[4], my[4], mz[4];
                    T al[4];
                    int lo_all = last_cut, hi_all = 0x7fffffff;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        mx[r] = 0; my[r] = 0; mz[r] = 0; al[r] = T(0);
                        if (r < nv) {
                            const int jl = jl0 + rg + r;
                            const int4 m = rowmeta[jl & (RR - 1)];
                            mx[r] = m.x; my[r] = max(m.y, (jl + WIN + EPV - 1) / EPV); mz[r] = m.z;
                            al[r] = alpha[jl & (RR - 1)];
                            lo_all = max(lo_all, my[r]); hi_all = min(hi_all, mz[r]);
                        }
                    }
                    if (vv >= lo_all && vv < hi_all) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < nv) {
                                const uint4 c = *reinterpret_cast<const uint4*>(smem + mx[r] + vv * 16);
                                VecOps<T, U>::axpy(c, al[r], fs);
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (vv >= my[r] && vv < mz[r]) {
                                const uint4 c = *reinterpret_cast<const uint4*>(smem + mx[r] + vv * 16);
                                VecOps<T, U>::axpy(c, al[r], fs);
                            }
                        }
                    }
                }
                uint4* fq = reinterpret_cast<uint4*>(fp);
#pragma unroll
                for (int e = 0; e < TPV; ++e) {
                    uint4 t;
                    memcpy(&t, reinterpret_cast<unsigned char*>(fs) + 16 * e, 16);
                    fq[e] = t;
                }
            }
            __syncwarp();
            if (lane == 0) {
                st_release(&prog[NBW + wb], (uint32_t)(cnext + 1));
                mbar_arrive(&empty[cs]);
            }
            trace_ev(p, lane, wb, 5, cnext);
            ++cnext;
            if (++cs == NST) { cs = 0; ++ck; }
        };

        int s = 0, k = 0;
        for (int u = 0; u < NP; ++u) {
            // the stage of panel u is only re-filled after C(u - NST): do the overdue axpys first (blocking)
            while (cnext <= u - NST) {
                mbar_wait(&cdone[cs], ck & 1);
                do_C();
            }
            trace_ev(p, lane, wb, 0, u);
            mbar_wait(&full[s], k & 1);
            trace_ev(p, lane, wb, 1, u);
            const int4 pm = panelmeta[s];
            const int P = pm.x, vmin = pm.y, vmax = pm.z, jl0 = pm.w;
            // ---- window coefficients of the panel's rows (row r of the panel -> bulk warp r % NBW) ----
            for (int r = wb; r < P; r += NBW) window_row<T, U>(smem, rowmeta, sm.wwin, jl0 + r, lane);
            // ---- A(u): backward dots of the rows of panel u ----------------------------------
            for (int rg = 0; rg < P; rg += 4) {
                const int nv = min(4, P - rg);
                typename Pk<T>::acc_t acc2[4];
                int mx[4], my[4], mz[4];
                int lo_all = 0, hi_all = 0x7fffffff;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc2[r] = Pk<T>::zero();
                    mx[r] = 0; my[r] = 0; mz[r] = 0;
                    if (r < nv) {
                        const int4 m = rowmeta[(jl0 + rg + r) & (RR - 1)];
                        mx[r] = m.x; my[r] = m.y; mz[r] = m.z;
                        lo_all = max(lo_all, m.y); hi_all = min(hi_all, m.z);
                    }
                }
                for (int v = vmin + tb; v < vmax; v += NBT) {
                    T es[EPV];
                    load_state_vec(eta_s + (size_t)v * EPV, es, TPV);
                    if (v >= lo_all && v < hi_all) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < nv) {
                                const uint4 c = *reinterpret_cast<const uint4*>(smem + mx[r] + v * 16);
                                VecOps<T, U>::dot(c, es, acc2[r]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (v >= my[r] && v < mz[r]) {
                                const uint4 c = *reinterpret_cast<const uint4*>(smem + mx[r] + v * 16);
                                VecOps<T, U>::dot(c, es, acc2[r]);
                            }
                        }
                    }
                }
                T acc[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = Pk<T>::sum(acc2[r]);
                const int rr = warp_reduce4(acc, lane);
                if ((lane & 7) == 0 && rr < nv) partial[wb * RR + ((jl0 + rg + rr) & (RR - 1))] = acc[0];
            }
            __syncwarp();
            if (lane == 0) st_release(&prog[wb], (uint32_t)(u + 1));
            trace_ev(p, lane, wb, 2, u);
            // opportunistic axpys: every panel the chain has already finished
            while (cnext <= u && mbar_test(&cdone[cs], ck & 1)) do_C();
            if (++s == NST) { s = 0; ++k; }
        }
        while (cnext < NP) {
            mbar_wait(&cdone[cs], ck & 1);
            do_C();
        }
    }
}

// q[j] += dq * sum_{k>j} R_jk x[k]   (update_q_factor, e_step.hpp:307-338); one warp per row.  Works on any row
// layout of ld.cu (packed in-unit rows, or the ext rows of a tiled block): `x` is indexed by global column.
template <typename T, typename U>
__global__ void backward_dot_kernel(int M, const unsigned char* __restrict__ packed, const int64_t* __restrict__ prow,
                                    const int32_t* __restrict__ pcs, const T* __restrict__ x, T* __restrict__ q, T dq) {
    constexpr int EPV = LdTraits<U>::EPV;
    const int row = blockIdx.x * (blockDim.x / WARP) + threadIdx.x / WARP;
    if (row >= M) return;
    const int lane = threadIdx.x % WARP;
    const int64_t o0 = prow[row];
    const int nv = (int)((prow[row + 1] - o0) / EPV);
    if (nv == 0) return;
    const int c0 = pcs[row];
    const uint4* src = reinterpret_cast<const uint4*>(packed + o0 * (int64_t)sizeof(U));
    typename Pk<T>::acc_t acc2 = Pk<T>::zero();
    for (int v = lane; v < nv; v += WARP) {
        const uint4 c = src[v];
        const int col = c0 + v * EPV;
        T xs[EPV];
#pragma unroll
        for (int e = 0; e < EPV; ++e) xs[e] = (col + e < M) ? x[col + e] : T(0);
        VecOps<T, U>::dot(c, xs, acc2);
    }
    const T acc = warp_sum(Pk<T>::sum(acc2));
    if (lane == 0) q[row] += dq * acc;
}

// The same product as backward_dot_kernel, organised for bandwidth: one CTA per item {row0, row1, col0, col1} = up to
// BWD_ROWS consecutive rows of one sweep unit (or of its ext rectangle) and the unit's column range.  x is staged
// through shared memory in BWD_COLS-column chunks aligned to col0 (every row layout of ld.cu is 16-byte aligned
// relative to it), so the dot reads it with 128-bit shared loads instead of per-element global gathers; one warp per
// row, four LD vectors in flight per lane.
constexpr int BWD_ROWS = 64;
constexpr int BWD_COLS = 4096;
constexpr int BWD_THREADS = 256;
#ifndef VB_BWD_RG
#define VB_BWD_RG 4
#endif
// rows a warp of row_dot_kernel handles together: 4 where a 16-byte LD vector needs more than 16 bytes of x from shared
// memory (integer LD codes, float LD with float64 state), 1 where it needs 16 (measured on the float64 C5 workload: the
// grouped form is slower there, 14.8 vs 13.5 ms per sweep)
template <typename T, typename U> constexpr int bwd_rows_per_warp() {
    return (int)sizeof(T) * LdTraits<U>::EPV > 16 ? VB_BWD_RG : 1;
}
template <typename T, typename U>
__global__ void __launch_bounds__(BWD_THREADS) row_dot_kernel(const int4* __restrict__ items,
                                                             const unsigned char* __restrict__ packed,
                                                             const int64_t* __restrict__ prow,
                                                             const int32_t* __restrict__ pcs, const T* __restrict__ x,
                                                             T* __restrict__ q, T dq) {
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int BWD_RG = bwd_rows_per_warp<T, U>();
    __shared__ __align__(16) T xs[BWD_COLS];
    __shared__ T racc[BWD_ROWS];
    const int4 it = items[blockIdx.x];
    const int row0 = it.x, row1 = it.y, col0 = it.z, col1 = it.w;
    const int warp = threadIdx.x / WARP, lane = threadIdx.x % WARP;
    if (threadIdx.x < BWD_ROWS) racc[threadIdx.x] = T(0);
    // columns below the first row's run are never touched: start at the chunk that holds it
    const int first_col = pcs[row0];
    for (int cbase = col0 + ((first_col - col0) / BWD_COLS) * BWD_COLS; cbase < col1; cbase += BWD_COLS) {
        __syncthreads();
        const int ncol = min(BWD_COLS, col1 - cbase);
        for (int i = threadIdx.x; i < BWD_COLS; i += BWD_THREADS) xs[i] = (i < ncol) ? x[cbase + i] : T(0);
        __syncthreads();
        if constexpr (BWD_RG == 1) {
            for (int r = row0 + warp; r < row1; r += BWD_THREADS / WARP) {
                const int64_t o0 = prow[r];
                const int nv = (int)((prow[r + 1] - o0) / EPV);
                const int c0 = pcs[r];
                const int v_lo = max(0, (cbase - c0) / EPV), v_hi = min(nv, (cbase + BWD_COLS - c0) / EPV);
                if (v_lo >= v_hi) continue;                                   // warp-uniform
                const uint4* src = reinterpret_cast<const uint4*>(packed + o0 * (int64_t)sizeof(U));
                const T* xr = xs + (c0 - cbase);
                typename Pk<T>::acc_t acc2 = Pk<T>::zero();
#pragma unroll 4
                for (int v = v_lo + lane; v < v_hi; v += WARP) {
                    const uint4 c = __ldg(src + v);
                    T xv[EPV];
                    load_state_vec(xr + (size_t)v * EPV, xv, EPV * (int)sizeof(T) / 16);
                    VecOps<T, U>::dot(c, xv, acc2);
                }
                const T acc = warp_sum(Pk<T>::sum(acc2));
                if (lane == 0) racc[r - row0] += acc;
            }
            continue;
        }
        // A warp takes BWD_RG consecutive rows at a time and walks the LD vectors of the column chunk ("block vectors":
        // 16-byte groups of columns counted from cbase -- every row layout of ld.cu starts at a multiple of EPV columns
        // from the unit start): the x values of a block vector are read from shared memory ONCE and used for all the
        // rows of the group (int8 LD: 64 bytes of x per 16 bytes of LD made the row-at-a-time form shared-memory
        // bound, 1.7 TB/s), and the group's loads are independent requests in flight.
        for (int r = row0 + BWD_RG * warp; r < row1; r += BWD_RG * (BWD_THREADS / WARP)) {
            const uint4* src[BWD_RG];
            int blo[BWD_RG], bhi[BWD_RG];
            int lo = BWD_COLS / EPV, hi = 0;
            bool aligned = true;
#pragma unroll
            for (int i = 0; i < BWD_RG; ++i) {
                blo[i] = 0; bhi[i] = 0; src[i] = nullptr;
                if (r + i < row1) {
                    const int64_t o0 = prow[r + i];
                    const int nv = (int)((prow[r + i + 1] - o0) / EPV);
                    const int d = pcs[r + i] - cbase;                     // first stored column relative to the chunk
                    aligned &= (d % EPV) == 0;
                    const int dv = d / EPV;                               // exact when aligned
                    blo[i] = max(0, dv); bhi[i] = min(BWD_COLS / EPV, dv + nv);
                    src[i] = reinterpret_cast<const uint4*>(packed + o0 * (int64_t)sizeof(U)) - dv;   // indexed by block vector
                    if (blo[i] < bhi[i]) { lo = min(lo, blo[i]); hi = max(hi, bhi[i]); }
                }
            }
            if (lo >= hi) continue;                                       // warp-uniform
            if (aligned) {
                typename Pk<T>::acc_t acc2[BWD_RG];
#pragma unroll
                for (int i = 0; i < BWD_RG; ++i) acc2[i] = Pk<T>::zero();
                for (int vb = lo + lane; vb < hi; vb += WARP) {
                    uint4 c[BWD_RG];
                    bool in[BWD_RG];
#pragma unroll
                    for (int i = 0; i < BWD_RG; ++i) {
                        in[i] = vb >= blo[i] && vb < bhi[i];
                        if (in[i]) c[i] = __ldg(src[i] + vb);
                    }
                    T xv[EPV];
                    load_state_vec(xs + (size_t)vb * EPV, xv, EPV * (int)sizeof(T) / 16);
#pragma unroll
                    for (int i = 0; i < BWD_RG; ++i)
                        if (in[i]) VecOps<T, U>::dot(c[i], xv, acc2[i]);
                }
#pragma unroll
                for (int i = 0; i < BWD_RG; ++i) {
                    const T acc = warp_sum(Pk<T>::sum(acc2[i]));
                    if (lane == 0 && r + i < row1) racc[r + i - row0] += acc;
                }
            } else {
                // a layout whose rows do not start on the chunk's vector grid: one row at a time
                for (int i = 0; i < BWD_RG && r + i < row1; ++i) {
                    const int rr = r + i;
                    const int64_t o0 = prow[rr];
                    const int nv = (int)((prow[rr + 1] - o0) / EPV);
                    const int c0 = pcs[rr];
                    const int v_lo = max(0, (cbase - c0) / EPV), v_hi = min(nv, (cbase + BWD_COLS - c0) / EPV);
                    if (v_lo >= v_hi) continue;
                    const uint4* s1 = reinterpret_cast<const uint4*>(packed + o0 * (int64_t)sizeof(U));
                    const T* xr = xs + (c0 - cbase);
                    typename Pk<T>::acc_t a2 = Pk<T>::zero();
                    for (int v = v_lo + lane; v < v_hi; v += WARP) {
                        const uint4 c = __ldg(s1 + v);
                        T xv[EPV];
                        load_state_vec(xr + (size_t)v * EPV, xv, EPV * (int)sizeof(T) / 16);
                        VecOps<T, U>::dot(c, xv, a2);
                    }
                    const T acc = warp_sum(Pk<T>::sum(a2));
                    if (lane == 0) racc[rr - row0] += acc;
                }
            }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < row1 - row0) q[row0 + threadIdx.x] += dq * racc[threadIdx.x];
}

// out[k] += scale * sum_{j in [row0, row1)} R_jk x[j]  for the columns k in [col0, col1) of every item
// {row0, row1, col0, col1}: the transposed product over a rectangle (or triangle) of stored rows.  Thread t of
// chunk c owns LD vector c * FWD_THREADS + t of the item's column range (the row layouts of ld.cu are aligned to
// 16 bytes relative to col0), walks the rows in order (fixed summation order: deterministic) and keeps its EPV
// sums in registers.  Used for the forward-external accumulator of tiled LD blocks and for q offsets.
#ifndef VB_FWD_THREADS
#define VB_FWD_THREADS 128
#endif
#ifndef VB_FWD_VPT
#define VB_FWD_VPT 1
#endif
constexpr int FWD_THREADS = VB_FWD_THREADS;
constexpr int FWD_VPT = VB_FWD_VPT;           // LD vectors per thread: a CTA covers FWD_THREADS * FWD_VPT * 16 contiguous bytes of every row
constexpr int FWD_MAX_ROWS = 4096;
constexpr int FWD_SLAB = 256;                 // rows whose offsets / first columns are staged at a time
// dynamic shared memory: the item's x values only (rows * sizeof(T)); a float64 tile of 2048 rows takes 16 KB, so a
// dozen CTAs share an SM and keep enough 16-byte loads in flight for HBM (a static 44 KB block held it to five:
// 1.06 ms for a 2.45 GB phase of the C5 workload)
inline size_t forward_smem_bytes(int max_rows, int tsize) { return (size_t)max_rows * (size_t)tsize; }
template <typename T, typename U>
__global__ void __launch_bounds__(FWD_THREADS) forward_axpy_kernel(const int4* __restrict__ items,
                                                                  const unsigned char* __restrict__ packed,
                                                                  const int64_t* __restrict__ prow,
                                                                  const int32_t* __restrict__ pcs,
                                                                  const T* __restrict__ x, T* __restrict__ out, T scale) {
    constexpr int EPV = LdTraits<U>::EPV;
    extern __shared__ __align__(16) unsigned char fwd_smem[];
    T* xs = reinterpret_cast<T*>(fwd_smem);
    __shared__ int64_t ro[FWD_SLAB + 1];               // row offsets / first columns of a slab of rows
    __shared__ int32_t rc[FWD_SLAB];
    const int4 it = items[blockIdx.y];
    const int row0 = it.x, row1 = it.y, col0 = it.z, col1 = it.w;
    const int cta_col0 = col0 + (int)blockIdx.x * FWD_THREADS * FWD_VPT * EPV;
    if (cta_col0 >= col1) return;                                            // whole CTA beyond the item (uniform)
    int col[FWD_VPT];
#pragma unroll
    for (int k = 0; k < FWD_VPT; ++k) col[k] = cta_col0 + (k * FWD_THREADS + (int)threadIdx.x) * EPV;
    for (int i = threadIdx.x; i < row1 - row0; i += FWD_THREADS) xs[i] = x[row0 + i];
    T acc[FWD_VPT][EPV];
#pragma unroll
    for (int k = 0; k < FWD_VPT; ++k)
#pragma unroll
        for (int e = 0; e < EPV; ++e) acc[k][e] = T(0);
    for (int s0 = row0; s0 < row1; s0 += FWD_SLAB) {
        const int ns = min(FWD_SLAB, row1 - s0);
        __syncthreads();
        for (int i = threadIdx.x; i <= ns; i += FWD_THREADS) ro[i] = prow[s0 + i];
        for (int i = threadIdx.x; i < ns; i += FWD_THREADS) rc[i] = pcs[s0 + i];
        __syncthreads();
        if (col[0] < col1) {
#pragma unroll 8 / FWD_VPT
            for (int i = 0; i < ns; ++i) {
                const int64_t o = ro[i];
                const int len = (int)(ro[i + 1] - o), c0 = rc[i];
                const T xv = xs[s0 - row0 + i];
#pragma unroll
                for (int k = 0; k < FWD_VPT; ++k) {
                    const int rel = col[k] - c0;
                    if (rel >= 0 && rel < len) {
                        const uint4 c = __ldg(reinterpret_cast<const uint4*>(packed + (o + rel) * (int64_t)sizeof(U)));
                        VecOps<T, U>::axpy(c, xv, acc[k]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < FWD_VPT; ++k) {
        if (col[k] < col1) {
#pragma unroll
            for (int e = 0; e < EPV; ++e)
                if (col[k] + e < col1) out[col[k] + e] = fma_t(scale, acc[k][e], out[col[k] + e]);
        }
    }
}

}  // namespace vb
