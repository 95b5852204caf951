// viprs_b200 -- the one-pass Gauss-Seidel sweep (spike-and-slab CAVI E-step) for sm_100a.
//
// Replaces e_step<T,U,I> + update_q_factor<T,U,I> of the reference
// (/root/reference/viprs/model/vi/e_step.hpp:343-442 and :307-338), threads=1 order.
//
// One CTA owns one LD block (rows r0..r1 that only reach columns inside [r0, r1)), keeps the
// strictly sequential per-SNP update order, and reads every LD entry ONCE:
//
//   q_j used at step j  =  dq * ( F_j + B_j ),
//       F_j = sum_{i<j} R_ij eta_i(new)     "forward"  -- axpy of finished rows into f_s[]
//       B_j = sum_{k>j} R_jk eta_k(old)     "backward" -- dot of row j with the not-yet-updated etas
//
// which is algebraically what the reference's incrementally maintained q holds at step j
// (its second pass, update_q_factor, is exactly the B_j term deferred to the end of the sweep).
//
// Warp roles (CTA = 2 + NBW warps):
//   warp 0  chain    : per-SNP scalar update, 32-lane register window covering the current panel's
//                      columns and the next panel's (in-window axpy via one FMA per step)
//   warp 1  producer : one 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) per row panel into a
//                      NSTAGE-deep shared-memory ring; writes the per-row metadata of the stage
//   warps 2.. bulk   : iteration u = { A(u): B_j for the rows of panel u ; C(u-2): axpy of panel
//                      u-2's finished rows into f_s[] for columns >= start of panel u }
// Hand-offs are mbarriers: full/empty (TMA ring), bulk_done (A(u),C(u-2) -> chain(u)),
// chain_done (eta_new of panel p -> C(p)).  No __syncthreads after the prologue.
#pragma once
#include "common.cuh"

namespace vb {

template <typename T>
struct SweepParams {
    const void* packed;
    const int64_t* prow;
    const int32_t* pcs;
    const int32_t* blk_row;
    const int32_t* blk_panel;
    const int32_t* panel_row;
    const int32_t* blk_order;
    int n_blocks;
    int stage_bytes;
    int bpad;                 // elements of T in each shared state array
    const T* std_beta;
    T* var_gamma;
    T* var_mu;
    T* eta;
    T* q;
    T* eta_diff;
    const T* u_logs;
    const T* sqrt_half_var_tau;
    const T* mu_mult;
    T dq_scale;
};

// shared-memory carve-up (all offsets 16-byte aligned)
struct SmemLayout {
    size_t stages, eta, f, rowmeta, panelmeta, partial, alpha, bars, total;
};
__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
__host__ __device__ inline SmemLayout make_layout(int bpad, int tsize, int stage_bytes, int nbw) {
    SmemLayout L;
    size_t o = 0;
    L.stages = o;    o += (size_t)NSTAGE * stage_bytes;
    L.eta = o;       o += align16((size_t)bpad * tsize);
    L.f = o;         o += align16((size_t)bpad * tsize);
    L.rowmeta = o;   o += (size_t)NSTAGE * PMAX * sizeof(int4);
    L.panelmeta = o; o += (size_t)NSTAGE * sizeof(int4);
    L.partial = o;   o += align16((size_t)NSLOT * nbw * PMAX * tsize);
    L.alpha = o;     o += align16((size_t)NSLOT * PMAX * tsize);
    L.bars = o;      o += (size_t)(2 * NSTAGE + 2 * NSLOT) * sizeof(uint64_t);
    L.total = o;
    return L;
}
inline int state_pad(int max_block) { return ((max_block + 15) / 16) * 16 + 16; }

template <typename T, typename U, int NBW>
__global__ void __launch_bounds__((NBW + 2) * WARP) sweep_kernel(const SweepParams<T> p) {
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int ES = (int)sizeof(U);
    constexpr int NBT = NBW * WARP;
    extern __shared__ __align__(128) unsigned char smem[];

    const SmemLayout L = make_layout(p.bpad, (int)sizeof(T), p.stage_bytes, NBW);
    unsigned char* stages = smem + L.stages;
    T* eta_s = reinterpret_cast<T*>(smem + L.eta);
    T* f_s = reinterpret_cast<T*>(smem + L.f);
    int4* rowmeta = reinterpret_cast<int4*>(smem + L.rowmeta);      // [NSTAGE][PMAX] {rowbase, vs, ve, -}
    int4* panelmeta = reinterpret_cast<int4*>(smem + L.panelmeta);  // [NSTAGE] {P, vmin, vmax, bytes}
    T* partial = reinterpret_cast<T*>(smem + L.partial);            // [NSLOT][NBW][PMAX]
    T* alpha = reinterpret_cast<T*>(smem + L.alpha);                // [NSLOT][PMAX]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTAGE;
    uint64_t* bulk_done = bars + 2 * NSTAGE;
    uint64_t* chain_done = bars + 2 * NSTAGE + NSLOT;

    const int tid = threadIdx.x, warp = tid / WARP, lane = tid % WARP;
    const int blk = p.blk_order[blockIdx.x];
    const int r0 = p.blk_row[blk], r1 = p.blk_row[blk + 1];
    const int B = r1 - r0;
    const int pan0 = p.blk_panel[blk];
    const int NP = p.blk_panel[blk + 1] - pan0;

    // ---- prologue: state into shared memory, barriers ---------------------------------------
    for (int i = tid; i < p.bpad; i += blockDim.x) {
        eta_s[i] = (i < B) ? p.eta[r0 + i] : T(0);
        f_s[i] = T(0);
    }
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NBW); }
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&bulk_done[s], NBW); mbar_init(&chain_done[s], 1); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 1) {
        // =============================== producer ===========================================
        const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(p.packed);
        for (int v = 0; v < NP; ++v) {
            const int s = v % NSTAGE, k = v / NSTAGE;
            if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
            const int rs = p.panel_row[pan0 + v], re = p.panel_row[pan0 + v + 1];
            const int P = re - rs;
            const int64_t obase = p.prow[rs];
            const int64_t oend = p.prow[re];
            int vs = 0x7fffffff, ve = 0;
            if (lane < P) {
                const int row = rs + lane;
                const int64_t o0 = p.prow[row], o1 = p.prow[row + 1];
                const int c = p.pcs[row] - r0;           // block-local first column, multiple of EPV
                const int nv = (int)(o1 - o0) / EPV;
                const int vs_r = c / EPV;
                int4 m;
                m.x = (int)((o0 - obase) * ES) - vs_r * 16;   // stage byte offset of local column 0
                m.y = vs_r; m.z = vs_r + nv; m.w = 0;
                rowmeta[s * PMAX + lane] = m;
                if (nv > 0) { vs = vs_r; ve = vs_r + nv; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vs = min(vs, __shfl_xor_sync(0xffffffffu, vs, o));
                ve = max(ve, __shfl_xor_sync(0xffffffffu, ve, o));
            }
            const uint32_t bytes = (uint32_t)((oend - obase) * ES);
            if (lane == 0) panelmeta[s] = make_int4(P, ve > 0 ? vs : 0, ve, (int)bytes);
            __syncwarp();
            if (lane == 0) {
                if (bytes > 0) {
                    mbar_arrive_expect_tx(&full[s], bytes);
                    tma_load_1d(stages + (size_t)s * p.stage_bytes, gsrc + obase * ES, bytes, &full[s]);
                } else {
                    mbar_arrive(&full[s]);
                }
            }
        }
    } else if (warp == 0) {
        // =============================== chain ==============================================
        const T eps = (sizeof(T) == 4) ? T(1.1920928955078125e-7) : T(1e-8);   // e_step.hpp:382
        const T dq = p.dq_scale;
        T carried = T(0);
        for (int pn = 0; pn < NP; ++pn) {
            const int s = pn % NSTAGE, slot = pn % NSLOT;
            const int rs = p.panel_row[pan0 + pn], re = p.panel_row[pan0 + pn + 1];
            const int P = re - rs;
            const int cut2 = ((pn + 1 < NP) ? p.panel_row[pan0 + pn + 2] : re) - r0;
            const int row = rs + lane;
            const bool valid = lane < P;
            T beta = T(0), mm = T(0), sv = T(0), ul = T(0), eo = T(0);
            if (valid) {
                beta = p.std_beta[row]; mm = p.mu_mult[row]; sv = p.sqrt_half_var_tau[row];
                ul = p.u_logs[row]; eo = p.eta[row];
            }
            mbar_wait(&bulk_done[slot], (pn / NSLOT) & 1);

            T bsum = T(0);
            if (valid) {
#pragma unroll
                for (int w = 0; w < NBW; ++w) bsum += partial[(slot * NBW + w) * PMAX + lane];
            }
            T X = valid ? (f_s[row - r0] + carried) : T(0);

            // window coefficients: lane <-> local column col; rows of this panel
            const unsigned char* st = stages + (size_t)s * p.stage_bytes;
            const int col = (rs - r0) + lane;
            const int vcol = col / EPV;
            T w[PMAX];
#pragma unroll
            for (int i = 0; i < PMAX; ++i) {
                w[i] = T(0);
                if (i < P) {
                    const int4 m = rowmeta[s * PMAX + i];
                    if (col > (rs - r0) + i && col < cut2 && vcol >= m.y && vcol < m.z)
                        w[i] = ld_elem<T, U>(st, m.x + col * ES);
                }
            }

            T o_mu = T(0), o_g = T(0), o_d = T(0), o_eta = T(0), o_F = T(0);
            bool o_skip = true;
#pragma unroll
            for (int i = 0; i < PMAX; ++i) {
                if (i < P) {
                    const T qv = dq * (X + bsum);
                    const T mu = fma_t(mm, beta, -mm * qv);                 // e_step.hpp:401
                    const T uu = sv * mu;                                   // :404
                    const T g = sigmoid_t(fma_t(uu, uu, ul));               // :405
                    const T d = fma_t(g, mu, -eo);                          // :408
                    const bool skip = abs_t(d) < eps;                       // :410
                    const T en = skip ? eo : (eo + d);                      // :431
                    if (lane == i) { o_mu = mu; o_g = g; o_d = skip ? T(0) : d; o_eta = en; o_F = X; o_skip = skip; }
                    const T a = shfl_t(en, i);
                    X = fma_t(w[i], a, X);                                  // :421 restricted to the window
                }
            }
            if (valid) {
                if (!o_skip) { p.var_mu[row] = o_mu; p.var_gamma[row] = o_g; p.eta[row] = o_eta; }   // :416-418,431
                p.eta_diff[row] = o_d;
                p.q[row] = dq * o_F;
                alpha[slot * PMAX + lane] = o_eta;
            }
            carried = shfl_down_t(X, P);
            __syncwarp();
            if (lane == 0) mbar_arrive(&chain_done[slot]);
        }
    } else {
        // =============================== bulk ===============================================
        const int wb = warp - 2;
        const int tb = wb * WARP + lane;
        for (int u = 0; u < NP; ++u) {
            const int s = u % NSTAGE, slot = u % NSLOT;
            mbar_wait(&full[s], (u / NSTAGE) & 1);
            {
                // ---- A(u): backward dots of the rows of panel u --------------------------------
                const int4 pm = panelmeta[s];
                const int P = pm.x, vmin = pm.y, vmax = pm.z;
                const unsigned char* st = stages + (size_t)s * p.stage_bytes;
                for (int rg = 0; rg < P; rg += 8) {
                    T acc[8];
                    int4 m[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        acc[r] = T(0);
                        m[r] = (rg + r < P) ? rowmeta[s * PMAX + rg + r] : make_int4(0, 0, 0, 0);
                    }
                    for (int v = vmin + tb; v < vmax; v += NBT) {
                        T es[EPV];
                        const uint4* ep = reinterpret_cast<const uint4*>(eta_s + (size_t)v * EPV);
#pragma unroll
                        for (int e = 0; e < (int)(EPV * sizeof(T) / 16); ++e) {
                            const uint4 t = ep[e];
                            memcpy(reinterpret_cast<unsigned char*>(es) + 16 * e, &t, 16);
                        }
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            if (v >= m[r].y && v < m[r].z) {
                                const uint4 c = *reinterpret_cast<const uint4*>(st + m[r].x + v * 16);
                                T vals[EPV];
                                Decode<T, U>::vec(c, vals);
#pragma unroll
                                for (int e = 0; e < EPV; ++e) acc[r] = fma_t(vals[e], es[e], acc[r]);
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        if (rg + r < P) {
                            const T sum = warp_sum(acc[r]);
                            if (lane == 0) partial[(slot * NBW + wb) * PMAX + rg + r] = sum;
                        }
                    }
                }
            }
            if (u >= 2) {
                // ---- C(u-2): axpy of panel u-2's finished rows into columns >= start(u) ---------
                const int pc = u - 2;
                const int sc = pc % NSTAGE, slotc = pc % NSLOT;
                mbar_wait(&chain_done[slotc], (pc / NSLOT) & 1);
                const int4 pm = panelmeta[sc];
                const int Pc = pm.x, vmaxc = pm.z;
                const unsigned char* st = stages + (size_t)sc * p.stage_bytes;
                const int cut = p.panel_row[pan0 + u] - r0;
                const int vlo = cut / EPV, rem = cut % EPV;
                int v = vlo + ((tb - vlo) % NBT + NBT) % NBT;      // static ownership: v == tb (mod NBT)
                for (; v < vmaxc; v += NBT) {
                    T fs[EPV];
                    uint4* fp = reinterpret_cast<uint4*>(f_s + (size_t)v * EPV);
#pragma unroll
                    for (int e = 0; e < (int)(EPV * sizeof(T) / 16); ++e) {
                        const uint4 t = fp[e];
                        memcpy(reinterpret_cast<unsigned char*>(fs) + 16 * e, &t, 16);
                    }
                    for (int r = 0; r < Pc; ++r) {
                        const int4 m = rowmeta[sc * PMAX + r];
                        if (v >= m.y && v < m.z) {
                            const uint4 c = *reinterpret_cast<const uint4*>(st + m.x + v * 16);
                            T vals[EPV];
                            Decode<T, U>::vec(c, vals);
                            if (v == vlo && rem) {
#pragma unroll
                                for (int e = 0; e < EPV; ++e) if (e < rem) vals[e] = T(0);
                            }
                            const T a = alpha[slotc * PMAX + r];
#pragma unroll
                            for (int e = 0; e < EPV; ++e) fs[e] = fma_t(vals[e], a, fs[e]);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < (int)(EPV * sizeof(T) / 16); ++e) {
                        uint4 t;
                        memcpy(&t, reinterpret_cast<unsigned char*>(fs) + 16 * e, 16);
                        fp[e] = t;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[sc]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bulk_done[slot]);
        }
    }
}

// q[j] += dq * sum_{k>j} R_jk x[k]   (update_q_factor, e_step.hpp:307-338); one warp per row.
template <typename T, typename U>
__global__ void backward_dot_kernel(int M, const U* __restrict__ packed, const int64_t* __restrict__ prow,
                                    const int32_t* __restrict__ pcs, const T* __restrict__ x,
                                    T* __restrict__ q, T dq) {
    constexpr int EPV = LdTraits<U>::EPV;
    const int row = blockIdx.x * (blockDim.x / WARP) + threadIdx.x / WARP;
    if (row >= M) return;
    const int lane = threadIdx.x % WARP;
    const int64_t o0 = prow[row];
    const int nv = (int)((prow[row + 1] - o0) / EPV);
    const int c0 = pcs[row];
    const uint4* src = reinterpret_cast<const uint4*>(packed + o0);
    T acc = T(0);
    for (int v = lane; v < nv; v += WARP) {
        const uint4 c = src[v];
        T vals[EPV];
        Decode<T, U>::vec(c, vals);
        const int col = c0 + v * EPV;
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
            const T xv = (col + e < M) ? x[col + e] : T(0);
            acc = fma_t(vals[e], xv, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0 && nv > 0) q[row] += dq * acc;
}

}  // namespace vb
