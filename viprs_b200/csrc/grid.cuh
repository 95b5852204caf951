// viprs_b200 -- the grid sweep: e_step_grid<T,U,I> + update_q_factor_matrix<T,U,I> of the reference
// (/root/reference/viprs/model/vi/e_step.hpp:555-647 and 266-303), threads=1 order, for sm_100a.
//
// One CTA owns one (LD block, tile of GT active grid columns) pair and walks the block's SNPs in order.  The LD
// row is read once per CTA and shared by the GT columns of the tile (and, through L2, by the other tiles of the
// same block, which are adjacent in blockIdx).
//
// q is kept the way the reference keeps it -- in/out and incremental -- but against the DENSE SYMMETRIC copy
// of the block (zero diagonal), so that ONE full-row axpy per updated SNP
//        q[k, g] = fma(R_jk, dq * eta_diff[j, g], q[k, g])      for every k of the block
// replaces both the reference's upper-row axpy (:623) and its second pass over the matrix (:291-302).
//
// Warp roles (no __syncthreads after the prologue):
//   warps 0..nbw-1  bulk : thread t owns GKPT = 16 columns x GT grid columns of q in REGISTERS for the whole
//                      sweep (128 registers).  Per LD row: one 128-bit shared load of its 16 codes (int8), the GT
//                      scaled deltas from a shared ring (two broadcast loads), decode, 64 FFMA2 (or 64 DFMA); the
//                      float tile pairs adjacent GRID columns so the decoded LD value is a 32-bit broadcast operand.
//                      After every panel the owner of the columns two panels ahead publishes them for the chain.
//   aux warp 0      producer : 1-D TMA bulk copies of (16 rows x <= 4 KB) stages into an nst-deep ring.
//   aux warp 1      chain : lane = (grid column g, slot w).  Per 16-row panel it holds X[c,g] = published q + the
//                      not-yet-published corrections of the previous and the current panel for the 16/NW columns it
//                      owns, runs the scalar update of GT columns at once and broadcasts dq*eta_diff with one SHFL per
//                      step.  Nothing but the dependent path: operands and window coefficients arrive ready-made.
//   aux warp 2      stager : stages the chain's operands (cp.async) in the form the steps consume and the decoded
//                      16 x 32 window of LD coefficients up to GPB panels ahead, and writes the outputs of every finished
//                      panel (measured on the C3 shape: the chain warp used to be busy 93 % of the time, 57 % of it in
//                      this per-panel work, and the bulk warps waited for it 13 % of theirs).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace vb {

constexpr int GP = 16;            // rows per panel
constexpr int GCW = 4096;         // bytes of one row inside a ring stage (column chunk)
constexpr int GKPT = 16;          // q columns per bulk thread
constexpr int GNST_MAX = 4;       // ring depth
constexpr int GAR = 4;            // depth (panels) of the delta ring
constexpr int GWW = 32;           // decoded window width: the panel's own 16 columns + the next 16
constexpr int GRID_MAX_BLOCK = 4096;   // 256 bulk threads x 16 columns
constexpr int GRID_MAX_BW = 8;
// Warps come in groups of four (one per SM sub-partition, each with its own 16K-register file).  The bulk warps
// fill whole warpgroups; the last warpgroup holds the producer, the chain warp and two idle warps.  With three
// warps per sub-partition ptxas may only assume 168 registers per thread, so the roles re-split the CTA's register
// pool with setmaxnreg: bulk warpgroups 200 (128 of them are the q tile), the auxiliary warpgroup 96.
constexpr int GRID_MAX_THREADS = (GRID_MAX_BW + 4) * WARP;
constexpr int GRID_BULK_REGS = 200;
constexpr int GRID_AUX_REGS = 96;
inline int grid_threads(int nbw) { return (((nbw + 3) / 4) * 4 + 4) * WARP; }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <typename T> struct GridT;
template <> struct GridT<float>  { static constexpr int GT = 8; };
template <> struct GridT<double> { static constexpr int GT = 4; };

constexpr int GPB = 3;            // depth (panels) of the parameter / window / output rings between the stager warp and the chain

struct GridLayout { uint32_t stages, wwin, wraw, pbuf, obuf, ybuf, alpha, qpub, bars, prog, total; };
inline GridLayout make_grid_layout(int tsize, int gt, int stage_bytes, int nst) {
    GridLayout L;
    uint32_t o = 0;
    L.stages = o; o += (uint32_t)nst * (uint32_t)stage_bytes; o = (o + 127u) & ~127u;
    L.wwin = o;   o += (uint32_t)GPB * GP * GWW * (uint32_t)tsize;
    L.wraw = o;   o += (uint32_t)GP * GWW * 8u;                       // raw window, sized for 8-byte LD elements
    L.pbuf = o;   o += (uint32_t)GPB * 6u * GP * (uint32_t)gt * (uint32_t)tsize;
    L.obuf = o;   o += (uint32_t)GPB * 2u * GP * (uint32_t)gt * (uint32_t)tsize;
    L.ybuf = o;   o += 2u * GP * (uint32_t)gt * (uint32_t)tsize;                // chain hand-over between its two warps
    L.alpha = o;  o += (uint32_t)GAR * GP * (uint32_t)gt * (uint32_t)tsize;
    L.qpub = o;   o += 2u * GP * (uint32_t)gt * (uint32_t)tsize;
    L.bars = o;   o += (2u * GNST_MAX + GAR) * 8u;
    L.prog = o;   o += (GRID_MAX_BW + 1) * 4u;                        // bulk warps' panel counters + the stager's
    L.total = o;
    return L;
}

struct GridPlan {
    const unsigned char* dense;    // dense symmetric blocks (biased integer codes), see ld.cu
    const int64_t* dblk_off;       // [n_blocks] byte offset of every block
    const int32_t* blk_row;        // [n_blocks+1]
    const int32_t* blk_order;      // [n_blocks] most expensive first
    const int32_t* active;         // [n_active] grid columns still iterating (e_step.hpp:606-609)
    int n_active, n_tiles, M, nst, stage_bytes, nbw;
    GridLayout L;
};

template <typename T>
struct GridArgs {
    const T* std_beta; T* var_gamma; T* var_mu; T* eta; T* q; T* eta_diff;
    const T* u_logs; const T* half_var_tau; const T* mu_mult; T dq;
};

// decode one 16-byte LD vector into EPV values of the state type
template <typename T, typename U> struct GridDecode;
template <> struct GridDecode<float, int8_t> {
    static __device__ __forceinline__ void vec(const uint4& c, float* o) {
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 p0, p1;
            VecOps<float, int8_t>::pairs(w[i], p0, p1);
            o[4 * i] = p0.x; o[4 * i + 1] = p0.y; o[4 * i + 2] = p1.x; o[4 * i + 3] = p1.y;
        }
    }
};
template <> struct GridDecode<float, int16_t> {
    static __device__ __forceinline__ void vec(const uint4& c, float* o) {
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 p = VecOps<float, int16_t>::pair(w[i]);
            o[2 * i] = p.x; o[2 * i + 1] = p.y;
        }
    }
};
template <> struct GridDecode<float, float> {
    static __device__ __forceinline__ void vec(const uint4& c, float* o) {
        o[0] = __uint_as_float(c.x); o[1] = __uint_as_float(c.y); o[2] = __uint_as_float(c.z); o[3] = __uint_as_float(c.w);
    }
};
template <typename U> struct GridDecode<double, U> {
    static __device__ __forceinline__ void vec(const uint4& c, double* o) { DecodeD<U>::vec(c, o); }
};

// cp.async (LDGSTS): global -> shared without a register round trip; N = 4, 8 or 16 bytes
template <int N>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
    if constexpr (N == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T, typename U>
__global__ void __launch_bounds__(GRID_MAX_THREADS, 1) grid_sweep_kernel(const GridPlan p, const GridArgs<T> a) {
    constexpr int GT = GridT<T>::GT;
    constexpr int NW = WARP / GT;                 // chain lanes per grid column
    constexpr int RPL = GP / NW;                  // panel rows (= columns) per chain lane
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int ES = (int)sizeof(U);
    constexpr int NVT = GKPT / EPV;               // LD vectors per bulk thread per row
    constexpr bool F32 = std::is_same<T, float>::value;
    extern __shared__ __align__(128) unsigned char smem[];

    const int tid = threadIdx.x, warp = tid / WARP, lane = tid % WARP;
    const int nbw = p.nbw, NT = nbw * WARP, NST = p.nst;
    const int blk = p.blk_order[blockIdx.x / p.n_tiles];
    const int tile = blockIdx.x % p.n_tiles;
    const int r0 = p.blk_row[blk];
    const int B = p.blk_row[blk + 1] - r0;
    const int Bp = (B + 15) & ~15;
    const int row_bytes = Bp * ES;
    const int NP = (B + GP - 1) / GP;
    const int nck = (row_bytes + GCW - 1) / GCW;
    const unsigned char* gblk = p.dense + p.dblk_off[blk];
    const size_t M = (size_t)p.M;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.L.bars);
    uint64_t* empty = full + GNST_MAX;
    uint64_t* cdone = full + 2 * GNST_MAX;
    uint32_t* prog = reinterpret_cast<uint32_t*>(smem + p.L.prog);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_alpha = sbase + p.L.alpha, a_qpub = sbase + p.L.qpub, a_wwin = sbase + p.L.wwin;
    const uint32_t a_full = sbase + p.L.bars, a_empty = a_full + 8u * GNST_MAX, a_cdone = a_full + 16u * GNST_MAX;
    const uint32_t a_prog = sbase + p.L.prog;

    if (tid == 0) {
        for (int s = 0; s < GNST_MAX; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)nbw); }
        for (int s = 0; s < GAR; ++s) mbar_init(&cdone[s], 1);
        for (int w = 0; w <= GRID_MAX_BW; ++w) prog[w] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    const int aux0 = ((nbw + 3) / 4) * 4;          // first warp of the auxiliary warpgroup
    // (each role's code must be dominated by its own setmaxnreg for ptxas to allocate against the new limit)
    if (warp >= aux0) {
      reg_dec<GRID_AUX_REGS>();
      // lane = (grid column g, slot w) in the chain and in the stager: RPL consecutive rows of one grid column
      const int g = lane % GT, w = lane / GT;
      const int gi = tile * GT + g;
      const bool gvalid = gi < p.n_active;
      const size_t colbase = (size_t)p.active[gvalid ? gi : tile * GT] * M + (size_t)r0;
      const uint32_t a_pbuf = sbase + p.L.pbuf, a_obuf = sbase + p.L.obuf;
      uint32_t* pready = prog + GRID_MAX_BW;       // panels whose parameters and window the stager has made ready
      // pbuf[GPB][6][GT][GP] of T: the chain's per-(SNP, column) operands, already in the form the steps consume
      //   0 mu_mult | 1 u_logs (base-2 units in float32) | 2 half_var_tau * mu_mult (ditto) | 3 std_beta | 4 -dq * eta_old | 5 eta_old
      // obuf[GPB][2][GT][GP] of T: var_mu, var_gamma of a finished panel (chain -> stager).
      // A lane's RPL rows of one array are RPL * sizeof(T) = 16 contiguous bytes.
      auto pslot = [&](int u, int arr) {
          return a_pbuf + (uint32_t)(((((u % GPB) * 6 + arr) * GT + g) * GP + RPL * w) * sizeof(T));
      };
      auto oslot = [&](int u, int arr) {
          return a_obuf + (uint32_t)(((((u % GPB) * 2 + arr) * GT + g) * GP + RPL * w) * sizeof(T));
      };
      if (warp == aux0) {
        // =============================== producer ============================================
        int s = 0, k = 0;
        for (int u = 0; u < NP; ++u) {
            const int nrows = min(GP, B - u * GP);
            for (int c = 0; c < nck; ++c) {
                const int cb = min(GCW, row_bytes - c * GCW);
                if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
                if (lane == 0) {
                    unsigned char* dst = smem + p.L.stages + (size_t)s * p.stage_bytes;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(nrows * cb));
                    if (nck == 1) {
                        tma_load_1d(dst, gblk + (size_t)u * GP * row_bytes, (uint32_t)(nrows * cb), &full[s]);
                    } else {
                        for (int r = 0; r < nrows; ++r)
                            tma_load_1d(dst + (size_t)r * cb, gblk + (size_t)(u * GP + r) * row_bytes + (size_t)c * GCW,
                                        (uint32_t)cb, &full[s]);
                    }
                }
                __syncwarp();
                if (++s == NST) { s = 0; ++k; }
            }
        }
      } else if (warp == aux0 + 2) {
        // =============================== stager ==============================================
        // Everything of the per-SNP update that is not on the dependent path: it stages the chain's operands and the
        // decoded LD window of a panel up to GPB panels ahead (cp.async, LDGSTS), and writes the outputs of every
        // finished panel (e_step.hpp:613-620, 633) from the (var_mu, var_gamma) pairs the chain leaves in shared memory.
        const T dq = a.dq;
        const uint32_t a_wraw = sbase + p.L.wraw;
        const bool vec_ok = ((colbase * sizeof(T)) % 16 == 0) && (((size_t)r0 * sizeof(T)) % 16 == 0) &&
                            ((reinterpret_cast<size_t>(a.mu_mult) | reinterpret_cast<size_t>(a.u_logs) |
                              reinterpret_cast<size_t>(a.half_var_tau) | reinterpret_cast<size_t>(a.eta) |
                              reinterpret_cast<size_t>(a.std_beta) | reinterpret_cast<size_t>(a.var_mu) |
                              reinterpret_cast<size_t>(a.var_gamma) | reinterpret_cast<size_t>(a.eta_diff)) % 16 == 0);
        const int wr = lane >> 1, wh = lane & 1;       // window: lane stages / decodes row wr, half wh (16 elements)
        auto stage = [&](int u) {
            // ---- raw copies: mu_mult -> 0, u_logs -> 1, half_var_tau -> 2, std_beta -> 3, eta -> 5
            const int j = u * GP + RPL * w;
            if (vec_ok && j + RPL <= B) {
                const size_t idx = colbase + (size_t)j;
                cp_async<16>(pslot(u, 0), a.mu_mult + idx);
                cp_async<16>(pslot(u, 1), a.u_logs + idx);
                cp_async<16>(pslot(u, 2), a.half_var_tau + idx);
                cp_async<16>(pslot(u, 3), a.std_beta + r0 + j);
                cp_async<16>(pslot(u, 5), a.eta + idx);
            } else {
#pragma unroll
                for (int m = 0; m < RPL; ++m) {
                    if (j + m < B) {
                        const size_t idx = colbase + (size_t)(j + m);
                        const uint32_t o = (uint32_t)(m * sizeof(T));
                        cp_async<sizeof(T)>(pslot(u, 0) + o, a.mu_mult + idx);
                        cp_async<sizeof(T)>(pslot(u, 1) + o, a.u_logs + idx);
                        cp_async<sizeof(T)>(pslot(u, 2) + o, a.half_var_tau + idx);
                        cp_async<sizeof(T)>(pslot(u, 3) + o, a.std_beta + r0 + j + m);
                        cp_async<sizeof(T)>(pslot(u, 5) + o, a.eta + idx);
                    }
                }
            }
            const int row = u * GP + wr, col = u * GP + 16 * wh;
            const bool wok = (row < B) && (col + 16 <= Bp);
            if (wok) {
                const unsigned char* src = gblk + (size_t)row * row_bytes + (size_t)col * ES;
#pragma unroll
                for (int c = 0; c < ES; ++c) cp_async<16>(a_wraw + (uint32_t)((wr * GWW + 16 * wh) * ES + 16 * c), src + 16 * c);
            }
            cp_async_wait_all();
            // ---- operands in the form the steps consume; rows past the block end and unused grid columns carry
            // all-zero operands (their delta is exactly 0)
            T v[6][RPL];
#pragma unroll
            for (int arr = 0; arr < 6; ++arr) {
                if (arr == 4) continue;
                const uint4 q4 = lds128(pslot(u, arr));
                memcpy(v[arr], &q4, 16);
            }
#pragma unroll
            for (int m = 0; m < RPL; ++m) {
                const bool ok = gvalid && (j + m < B);
                const T mmv = ok ? v[0][m] : T(0);
                T ulv = ok ? v[1][m] : T(0);
                T hmv = ok ? mul_t(v[2][m], mmv) : T(0);
                if constexpr (F32) {           // logit carried in base-2 units: the sigmoid's exponential is a bare MUFU.EX2
                    hmv = mul_t(hmv, T(1.4426950408889634));
                    ulv = mul_t(ulv, T(1.4426950408889634));
                }
                const T eov = ok ? v[5][m] : T(0);
                v[0][m] = mmv; v[1][m] = ulv; v[2][m] = hmv;
                v[3][m] = ok ? v[3][m] : T(0);
                v[4][m] = -mul_t(dq, eov);
                v[5][m] = eov;
            }
#pragma unroll
            for (int arr = 0; arr < 6; ++arr) {
                uint4 q4;
                memcpy(&q4, v[arr], 16);
                sts128(pslot(u, arr), q4);
            }
            // ---- decoded window
            const uint32_t dst = a_wwin + (uint32_t)((((u % GPB) * GP + wr) * GWW + 16 * wh) * sizeof(T));
#pragma unroll
            for (int c = 0; c < ES; ++c) {
                T x[EPV];
                if (wok) {
                    const uint4 rv = lds128(a_wraw + (uint32_t)((wr * GWW + 16 * wh) * ES + 16 * c));
                    GridDecode<T, U>::vec(rv, x);
                } else {
#pragma unroll
                    for (int e = 0; e < EPV; ++e) x[e] = T(0);
                }
#pragma unroll
                for (int e = 0; e < EPV; e += 16 / (int)sizeof(T)) {
                    uint4 q4;
                    memcpy(&q4, x + e, 16);
                    sts128(dst + (uint32_t)((c * EPV + e) * sizeof(T)), q4);
                }
            }
            __syncwarp();
            if (lane == 0) st_release(pready, (uint32_t)(u + 1));
        };
        for (int u = 0; u < min(GPB, NP); ++u) stage(u);
        for (int u = 0; u < NP; ++u) {
            mbar_wait_a(a_cdone + 8u * (uint32_t)(u % GAR), (u / GAR) & 1);          // the chain has finished panel u
            // panel outputs: every lane owns RPL consecutive rows of its grid column (16 contiguous bytes per array)
            T o_mu[RPL], o_g[RPL], eo[RPL], o_d[RPL], o_en[RPL];
            { uint4 q4 = lds128(oslot(u, 0)); memcpy(o_mu, &q4, 16);
              q4 = lds128(oslot(u, 1)); memcpy(o_g, &q4, 16);
              q4 = lds128(pslot(u, 5)); memcpy(eo, &q4, 16); }
#pragma unroll
            for (int m = 0; m < RPL; ++m) {
                o_d[m] = fma_t(o_g[m], o_mu[m], -eo[m]);                         // :620
                o_en[m] = add_t(eo[m], o_d[m]);                                   // :633
            }
            if (gvalid) {
                const int jr = u * GP + RPL * w;
                if (vec_ok && jr + RPL <= B) {
                    const size_t idx = colbase + (size_t)jr;
                    uint4 t4;
                    memcpy(&t4, o_mu, 16); *reinterpret_cast<uint4*>(a.var_mu + idx) = t4;
                    memcpy(&t4, o_g, 16); *reinterpret_cast<uint4*>(a.var_gamma + idx) = t4;
                    memcpy(&t4, o_d, 16); *reinterpret_cast<uint4*>(a.eta_diff + idx) = t4;
                    memcpy(&t4, o_en, 16); *reinterpret_cast<uint4*>(a.eta + idx) = t4;
                } else {
#pragma unroll
                    for (int m = 0; m < RPL; ++m) {
                        if (jr + m < B) {
                            const size_t idx = colbase + (size_t)(jr + m);
                            a.var_mu[idx] = o_mu[m]; a.var_gamma[idx] = o_g[m]; a.eta_diff[idx] = o_d[m];
                            a.eta[idx] = o_en[m];
                        }
                    }
                }
            }
            __syncwarp();                                       // every lane has read its slots of ring entry u % GPB
            if (u + GPB < NP) stage(u + GPB);
        }
      } else if (warp == aux0 + 1 || warp == aux0 + 3) {
        // =============================== chain ===============================================
        // Two warps on different SM sub-partitions take the panels in turns (even / odd): the two bulk warps that share a
        // sub-partition with the chain set the pace of the whole CTA through the 2-panel window, so the chain's issue
        // slots are spread over two sub-partitions.  The carry between panels (Y: corrections to the next panel's own
        // columns) goes through shared memory; the hand-over is the cdone barrier of the previous panel.
        const int which = (warp - aux0 - 1) >> 1;
        const uint32_t a_ybuf = sbase + p.L.ybuf;
        auto yslot = [&](int u) { return a_ybuf + (uint32_t)((((u & 1) * GT + g) * GP + RPL * w) * sizeof(T)); };
        const T dq = a.dq;
        T mm[RPL], ul[RPL], hm[RPL], bt[RPL], ndqeo[RPL];
        T X[RPL], Y[RPL];
        for (int u = which; u < NP; u += 2) {
            const int j0 = u * GP;
            if (u > 0) {
                mbar_wait_a(a_cdone + 8u * (uint32_t)((u - 1) % GAR), ((u - 1) / GAR) & 1);
                const uint4 y4 = lds128(yslot(u - 1));
                memcpy(Y, &y4, 16);
            } else {
#pragma unroll
                for (int m = 0; m < RPL; ++m) Y[m] = T(0);
            }
            // the stager has made panel u ready; every bulk warp has applied (and published past) panel u-2
            {
                uint32_t spins = 0;
                for (;;) {
                    bool ok = true;
                    if (lane < nbw) ok = (u < 2) || (ld_acquire_a(a_prog + 4u * (uint32_t)lane) >= (uint32_t)(u - 1));
                    else if (lane == GRID_MAX_BW) ok = ld_acquire_a(a_prog + 4u * GRID_MAX_BW) >= (uint32_t)(u + 1);
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (++spins > kSpinLimit) __trap();
                }
            }
            { uint4 q4 = lds128(pslot(u, 0)); memcpy(mm, &q4, 16);
              q4 = lds128(pslot(u, 1)); memcpy(ul, &q4, 16);
              q4 = lds128(pslot(u, 2)); memcpy(hm, &q4, 16);
              q4 = lds128(pslot(u, 3)); memcpy(bt, &q4, 16);
              q4 = lds128(pslot(u, 4)); memcpy(ndqeo, &q4, 16); }
#pragma unroll
            for (int m = 0; m < RPL; ++m) {
                const int cl = RPL * w + m;
                T base;
                if (u >= 2) base = lds_t(a_qpub + (uint32_t)((((u & 1) * GP + cl) * GT + g) * sizeof(T)), T());
                else base = (j0 + cl < B) ? a.q[colbase + (size_t)(j0 + cl)] : T(0);
                X[m] = add_t(base, Y[m]);
                Y[m] = T(0);
            }
            const uint32_t wrow = a_wwin + (uint32_t)(((u % GPB) * GP * GWW + RPL * w) * sizeof(T));
            // The steps below are branch-free and store-free: rows past the block end carry all-zero inputs (their
            // delta is exactly 0), outputs are kept by predicated moves and handed over once per panel.
            T o_mu[RPL], o_g[RPL], o_al[RPL];
#pragma unroll
            for (int m = 0; m < RPL; ++m) { o_mu[m] = T(0); o_g[m] = T(0); o_al[m] = T(0); }
#pragma unroll 1
            for (int ow = 0; ow < NW; ++ow) {            // the NW lanes of a grid column take turns: RPL rows each
                const bool own = (w == ow);
                const int src_lane = g + GT * ow;
                uint4 vx[RPL], vy[RPL];                  // window coefficients of the RPL steps, fetched up front
#pragma unroll
                for (int m = 0; m < RPL; ++m) {
                    vx[m] = lds128(wrow + (uint32_t)((ow * RPL + m) * GWW * sizeof(T)));
                    vy[m] = lds128(wrow + (uint32_t)(((ow * RPL + m) * GWW + 16) * sizeof(T)));
                }
#pragma unroll
                for (int m = 0; m < RPL; ++m) {
                    T cX[RPL], cY[RPL];                  // RPL * sizeof(T) == 16
                    memcpy(cX, &vx[m], 16);
                    memcpy(cY, &vy[m], 16);
                    // e_step.hpp:613-623 with the operations off the critical path hoisted: hm = half_var_tau*mu_mult,
                    // ndqeo = -dq*eta_old  =>  dq*eta_diff = fma(gamma, dq*mu, -dq*eta_old)
                    const T r = add_t(bt[m], -X[m]);
                    const T mu = mul_t(mm[m], r);                            // :613
                    T gam;                                                   // :616-617
                    if constexpr (F32) gam = sigmoid2_t(fma_t(mul_t(hm[m], r), mu, ul[m]));
                    else gam = sigmoid_t(fma_t(mul_t(hm[m], r), mu, ul[m]));
                    const T al = fma_t(gam, mul_t(dq, mu), ndqeo[m]);        // :620, :623
                    const T ab = shfl_t(al, src_lane);
#pragma unroll
                    for (int t = 0; t < RPL; ++t) X[t] = fma_t(cX[t], ab, X[t]);
#pragma unroll
                    for (int t = 0; t < RPL; ++t) Y[t] = fma_t(cY[t], ab, Y[t]);
                    o_mu[m] = own ? mu : o_mu[m];
                    o_g[m] = own ? gam : o_g[m];
                    o_al[m] = own ? al : o_al[m];
                }
            }
            // hand-over: scaled deltas to the bulk warps, (var_mu, var_gamma) to the stager
#pragma unroll
            for (int m = 0; m < RPL; ++m)
                sts_t(a_alpha + (uint32_t)((((u % GAR) * GP + RPL * w + m) * GT + g) * sizeof(T)), o_al[m]);
            { uint4 t4;
              memcpy(&t4, o_mu, 16); sts128(oslot(u, 0), t4);
              memcpy(&t4, o_g, 16); sts128(oslot(u, 1), t4);
              memcpy(&t4, Y, 16); sts128(yslot(u), t4); }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(a_cdone + 8u * (uint32_t)(u % GAR));
        }
      }
    } else {
        reg_inc<GRID_BULK_REGS>();
        if (warp >= nbw) return;                   // padding warps of the last bulk warpgroup
        // =============================== bulk ================================================
        const int t = tid;                                      // 0 .. NT-1
        int gcol[GT];
        bool gval[GT];
#pragma unroll
        for (int g = 0; g < GT; ++g) {
            const int gi = tile * GT + g;
            gval[g] = gi < p.n_active;
            gcol[g] = p.active[gval[g] ? gi : tile * GT];
        }
        // q registers.  float: pairs of adjacent GRID columns, q2[col][gp] = (q[col][2gp], q[col][2gp+1]), so that one
        // FFMA2 takes the decoded LD value as a 32-bit broadcast operand (R.F32), the two scaled deltas as a natural
        // 64-bit pair from the ring, and the accumulator pair -- measured 13% faster than pairing along the LD columns
        // (the loop body in isolation: scripts/microbench/grid_loop_bench.cu).  double: scalars.
        using QT = typename std::conditional<F32, float2, double>::type;
        constexpr int QG = F32 ? GT / 2 : GT;
        QT qr[NVT][EPV][QG];
        auto q_load = [&](int col, int g) -> T {
            return (col < B && gval[g]) ? a.q[(size_t)gcol[g] * M + (size_t)r0 + col] : T(0);
        };
        // float32: a thread's EPV columns of one grid column are EPV * 4 contiguous bytes (column-major state): 128-bit
        // loads / stores when the whole vector lies inside the block and the addresses are 16-byte aligned
        [[maybe_unused]] const bool q_vec_ok = F32 && (EPV % 4 == 0) && ((reinterpret_cast<size_t>(a.q) % 16) == 0) &&
                                               (M % 4 == 0) && (r0 % 4 == 0);
#pragma unroll
        for (int i = 0; i < NVT; ++i) {
            const int col0 = (t + NT * i) * EPV;
            if constexpr (F32 && (EPV % 4 == 0)) {
                if (q_vec_ok && col0 + EPV <= B) {
#pragma unroll
                    for (int g = 0; g < GT; ++g) {
                        const float4* src = reinterpret_cast<const float4*>(a.q + (size_t)gcol[g] * M + (size_t)r0 + col0);
#pragma unroll
                        for (int e4 = 0; e4 < EPV / 4; ++e4) {
                            const float4 v4 = gval[g] ? src[e4] : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (g & 1) qr[i][4 * e4 + k][g / 2].y = vv[k];
                                else qr[i][4 * e4 + k][g / 2].x = vv[k];
                            }
                        }
                    }
                    continue;
                }
            }
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
#pragma unroll
                for (int g = 0; g < QG; ++g) {
                    if constexpr (F32) {
                        qr[i][e][g].x = q_load(col0 + e, 2 * g);
                        qr[i][e][g].y = q_load(col0 + e, 2 * g + 1);
                    } else {
                        qr[i][e][g] = q_load(col0 + e, g);
                    }
                }
            }
        }
        uint32_t voff[NVT];      // byte offset of vector i inside its chunk's stage row
        int vchunk[NVT];         // chunk index, or -1 when the vector lies beyond the block's padded width
#pragma unroll
        for (int i = 0; i < NVT; ++i) {
            const int byte = 16 * (t + NT * i);
            const int ck = (nck == 1) ? 0 : ((NT * i) >> 8);       // chunk of vector i = floor(16 (t + NT i) / GCW)
            vchunk[i] = (byte < row_bytes) ? ck : -1;
            voff[i] = (uint32_t)(byte - ck * GCW);
        }

        int s = 0, k = 0;
        for (int u = 0; u < NP; ++u) {
            const int nrows = min(GP, B - u * GP);
            mbar_wait_a(a_cdone + 8u * (uint32_t)(u % GAR), (u / GAR) & 1);
            const uint32_t abase = a_alpha + (uint32_t)((u % GAR) * GP * GT * sizeof(T));
            for (int c = 0; c < nck; ++c) {
                const int cb = min(GCW, row_bytes - c * GCW);
                mbar_wait_a(a_full + 8u * (uint32_t)s, k & 1);
                const uint32_t st = sbase + p.L.stages + (uint32_t)s * (uint32_t)p.stage_bytes;
                bool any = false;
#pragma unroll
                for (int i = 0; i < NVT; ++i) any |= (vchunk[i] == c);
                if (any) {
                    // (software-pipelined forms of this loop -- loads only, or loads + decode one row ahead in a two-row body --
                    // measured slower on B200, as did decoding through I2F.S8 on the conversion unit: scripts/microbench/grid_loop_bench.cu)
#pragma unroll 1
                    for (int r = 0; r < nrows; ++r) {
                        // the GT scaled deltas of row r: GT * sizeof(T) = 32 bytes, two broadcast loads
                        const uint4 v0 = lds128(abase + (uint32_t)(r * GT * sizeof(T)));
                        const uint4 v1 = lds128(abase + (uint32_t)(r * GT * sizeof(T) + 16));
                        QT al[QG];
                        if constexpr (F32) {
                            al[0] = make_float2(__uint_as_float(v0.x), __uint_as_float(v0.y));
                            al[1] = make_float2(__uint_as_float(v0.z), __uint_as_float(v0.w));
                            al[2] = make_float2(__uint_as_float(v1.x), __uint_as_float(v1.y));
                            al[3] = make_float2(__uint_as_float(v1.z), __uint_as_float(v1.w));
                        } else {
                            al[0] = __hiloint2double((int)v0.y, (int)v0.x);
                            al[1] = __hiloint2double((int)v0.w, (int)v0.z);
                            al[2] = __hiloint2double((int)v1.y, (int)v1.x);
                            al[3] = __hiloint2double((int)v1.w, (int)v1.z);
                        }
#pragma unroll
                        for (int i = 0; i < NVT; ++i) {
                            if (vchunk[i] == c) {
                                const uint4 cv = lds128(st + (uint32_t)(r * cb) + voff[i]);
                                T v[EPV];
                                GridDecode<T, U>::vec(cv, v);
#pragma unroll
                                for (int e = 0; e < EPV; ++e) {
#pragma unroll
                                    for (int g = 0; g < QG; ++g) {
                                        if constexpr (F32) {
                                            qr[i][e][g] = fma2(make_float2(v[e], v[e]), al[g], qr[i][e][g]);
                                        } else {
                                            qr[i][e][g] = fma(v[e], al[g], qr[i][e][g]);
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_empty + 8u * (uint32_t)s);
                if (++s == NST) { s = 0; ++k; }
            }
            // publish the columns of panel u+2 (complete through panel u) for the chain
            const int pp = u + 2;
            if (pp < NP) {
#pragma unroll
                for (int i = 0; i < NVT; ++i) {
                    const int vv = t + NT * i;
                    if (vv / NVT == pp) {
                        const int kk = (vv % NVT) * EPV;
#pragma unroll
                        for (int e = 0; e < EPV; ++e) {
#pragma unroll
                            for (int g = 0; g < QG; ++g) {
                                const uint32_t ad = a_qpub + (uint32_t)((((pp & 1) * GP + kk + e) * GT) * sizeof(T));
                                if constexpr (F32) {
                                    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(ad + (uint32_t)(8 * g)), "f"(qr[i][e][g].x),
                                                 "f"(qr[i][e][g].y) : "memory");
                                } else {
                                    sts_t(ad + (uint32_t)(8 * g), qr[i][e][g]);
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) st_release_a(a_prog + 4u * (uint32_t)warp, (uint32_t)(u + 1));
        }
        // ---- epilogue: q back to global memory ------------------------------------------------
#pragma unroll
        for (int i = 0; i < NVT; ++i) {
            const int col0 = (t + NT * i) * EPV;
            if constexpr (F32 && (EPV % 4 == 0)) {
                if (q_vec_ok && col0 + EPV <= B) {
#pragma unroll
                    for (int g = 0; g < GT; ++g) {
                        if (!gval[g]) continue;
                        float4* dst = reinterpret_cast<float4*>(a.q + (size_t)gcol[g] * M + (size_t)r0 + col0);
#pragma unroll
                        for (int e4 = 0; e4 < EPV / 4; ++e4) {
                            float vv[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) vv[k] = (g & 1) ? qr[i][4 * e4 + k][g / 2].y : qr[i][4 * e4 + k][g / 2].x;
                            dst[e4] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                        }
                    }
                    continue;
                }
            }
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const int col = col0 + e;
                if (col >= B) continue;
#pragma unroll
                for (int g = 0; g < QG; ++g) {
                    if constexpr (F32) {
                        if (gval[2 * g]) a.q[(size_t)gcol[2 * g] * M + (size_t)r0 + col] = qr[i][e][g].x;
                        if (gval[2 * g + 1]) a.q[(size_t)gcol[2 * g + 1] * M + (size_t)r0 + col] = qr[i][e][g].y;
                    } else {
                        if (gval[g]) a.q[(size_t)gcol[g] * M + (size_t)r0 + col] = qr[i][e][g];
                    }
                }
            }
        }
    }
}

}  // namespace vb
