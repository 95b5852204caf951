// viprs_b200 -- the register-resident variant of the one-pass sweep (float32 state, LD blocks <= 4096 SNPs).
//
// Same algorithm, hand-offs, producer and chain warp as sweep.cuh; what changes is where the block state
// lives and how the bulk work is split:
//   warps 0..3  "A" : backward dots B_j = sum_{k>j} R_jk eta_k(old).  Thread (warp w, lane l) owns vector l of its NVT =
//                     32/EPV tiles of 32 LD vectors (fast_tile) of every row and keeps eta_old of its 32 columns in
//                     REGISTERS.  int8 LD: eta_old is held as NLIMB balanced base-128 digits (int8) of a
//                     block-scaled fixed-point value and the dot is IDP.4A (dp4a, u8 x s8 -> exact int32)
//                     on the raw biased bytes -- no dequantisation instruction at all on this side.
//   warps 4..7  "C" : forward axpy f_k += R_jk eta_j(new) with the same vector ownership and f in REGISTERS;
//                     columns whose accumulation is complete are published to a 128-entry ring for the chain.
//                     Both classify their tiles per panel (dead / interior / boundary, see the A role below).
// No eta / f arrays in shared memory => the whole 113 KB (two CTAs per SM) minus ~18 KB goes to the TMA ring.
// What bounds this kernel (measured): the latency of one panel through TMA -> A -> chain -> C with
// three stages in flight, not HBM and not issue slots.
#pragma once
#include <type_traits>

#include "sweep.cuh"

namespace vb {

constexpr int NAW = 4;                    // A warps
constexpr int NCW = 4;                    // C warps
constexpr int NWW = 0;                    // progress-only warps between the A and C counters (none)
// 10 warps per CTA: 0-3 A, 4-7 C, 8 chain, 9 producer (see VB_CHAIN_WARP).
constexpr int FAST_WARPS = 10;
#ifndef VB_CHAIN_WARP
#define VB_CHAIN_WARP 8         // 8: the chain shares its sub-partition with A0 / C0 (the lightest bulk warps), 9: with A1 / C1 (measured: C2 1.006 vs 1.020 ms, C4 1.702 vs 1.734 ms)
#endif
#ifndef VB_PRODUCER_WARP
#define VB_PRODUCER_WARP (17 - VB_CHAIN_WARP)
#endif
#ifndef VB_A_BASE
#define VB_A_BASE 0             // first of the four A warps
#endif
#ifndef VB_C_BASE
#define VB_C_BASE 4             // first of the four C warps
#endif
constexpr int FAST_CHAIN_WARP = VB_CHAIN_WARP, FAST_PRODUCER_WARP = VB_PRODUCER_WARP;
// Which share of the tiles (index 0..3, fast_tile) a bulk warp owns.  Index 3 owns the last tile of the block and is
// the only live one over the last eighth of the rows, index 2 joins it over the last quarter, ...  Warp w runs on SM
// sub-partition w % 4, so with index = warp for both roles the late, short rows of BOTH co-resident CTAs are served by
// the four warps of ONE sub-partition while the other three idle.  smsp_rot = 2 places the four roles (A, C of the
// CTA, A, C of its SM neighbour -- CTA b and CTA b + n_sm, by launch order) as a Latin square: every sub-partition
// hosts the indices 0, 1, 2, 3 once.  smsp_rot = 1: only the C warps are shifted by one; 0: index = warp.
__device__ __forceinline__ int fast_a_index(int warp, int rot) {
    return (unsigned)(warp - VB_A_BASE) < 4u ? ((warp - VB_A_BASE + rot) & 3) : -1;
}
__device__ __forceinline__ int fast_c_index(int warp, int rot) {
    return (unsigned)(warp - VB_C_BASE) < 4u ? ((warp - VB_C_BASE + rot) & 3) : -1;
}
constexpr int FAST_MAX_BLOCK = 4096;      // 128 threads x 32 columns
constexpr int FR = 128;                   // published-f ring (columns)
constexpr int HP = 258;                   // entries of the fixed-point prefix sum (>= 4096/16 + 1; HP * 8 a multiple of 16)
constexpr int NLIMB_MAX = 4;

struct FastLayout {
    uint32_t stages, rowmeta, panelmeta, panelmeta2, rowbase, partial, alpha, wwin, fring, hc, zero, red, xown, bsum, bars,
        counters, total;
};
inline FastLayout make_fast_layout(int stage_bytes, int nst) {
    FastLayout L;
    uint32_t o = 0;
    L.stages = o;    o += (uint32_t)nst * (uint32_t)stage_bytes;       o = align128(o);
    L.rowmeta = o;   o += RR * (uint32_t)sizeof(int4);
    L.panelmeta = o; o += NST_MAX * (uint32_t)sizeof(int4);
    L.panelmeta2 = o; o += NST_MAX * (uint32_t)sizeof(int4);       // {lo, hi}: vectors stored by every row of the panel
    L.rowbase = o;   o += RR * 4;                                  // rowmeta[].x alone: one LDS.128 = four rows
    L.partial = o;   o += NAW * RR * 4;
    L.alpha = o;     o += RR * 4;
    L.wwin = o;      o += RR * W2 * 4;               // version 2 layout (64 slots per row); version 1 uses 48 of them
    L.xown = o;      o += RR * 4;                    // version 2: chain -> output role
    L.bsum = o;      o += RR * 4;
    L.fring = o;     o += FR * 4;
    L.hc = o;        o += HP * 8;
    L.zero = o;      o += 32;
    L.red = L.partial;                   // prologue scratch, dead before the first partial is written
    L.bars = o;      o += 3 * NST_MAX * (uint32_t)sizeof(uint64_t);
    L.counters = o;  o += (NAW + NWW + NCW + 2) * (uint32_t)sizeof(uint32_t);     // + chain_panels, out_rows (version 2)
    L.total = o;
    return L;
}

// Static column ownership of the bulk warps: tile i = LD vectors [32 i, 32 i + 32); thread (warp w, lane l) owns vector
// l of its NVT tiles.  Row j only touches the vectors >= j / EPV, so a tile's work grows with its index.  Two deals:
// plain (warp w owns tiles w, 4 + w, 8 + w, ...) and serpentine (w, 7 - w, 8 + w, 15 - w, ...: every warp, i.e. every
// SM sub-partition, gets the same share of the triangle).  The sweep is bound by the latency of the panel hand-offs,
// not by issue slots, and the plain deal measures slightly faster; the serpentine one is kept as a build switch.
#ifndef VB_GROUP_UNROLL
#define VB_GROUP_UNROLL 1          // unroll factor of the A / C row-group loops (2: 1.02 -> 1.25 ms on C2: spills, instruction cache)
#endif
constexpr int kGroupUnroll = VB_GROUP_UNROLL;
#ifndef VB_FAST_BALANCE
#define VB_FAST_BALANCE 0       // measured on B200 (C2 workload): plain 1.007 ms, serpentine (1) 1.038 ms, load-aware (2) 1.014 ms
#endif
__device__ __forceinline__ int fast_tile(int w, int c) {
#if VB_FAST_BALANCE == 2
    // warp 0 keeps the plain (lightest) pair because its sub-partition also hosts the chain warp; warps 1..3 are
    // dealt the odd rounds in reverse: tiles {0,4}, {1,7}, {2,6}, {3,5} -> loads 5 : 9 : 9 : 9 (+ chain, + producer)
    return 4 * c + (((c & 1) && w > 0) ? 4 - w : w);
#elif VB_FAST_BALANCE
    return 4 * c + ((c & 1) ? 3 - w : w);
#else
    return 4 * c + w;
#endif
}

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {      // sum_b u8(a.b) * s8(b.b) + c
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// version 2 adds an output warp (warp FAST_WARPS): 11 warps per CTA, still two CTAs per SM (<= 88 registers)
constexpr int FAST_OUTPUT_WARP = FAST_WARPS;
template <int VER> constexpr int fast_threads() { return (VER == 2 ? FAST_WARPS + 1 : FAST_WARPS) * WARP; }

// INCR (version 1 only): q maintained incrementally like the reference (see chain_role) -- the A warps only cut the
// window coefficients, their backward dots are skipped.
template <typename U, typename Model, int NLIMB, int VER, bool INCR = false>
__global__ void __launch_bounds__(fast_threads<VER>(), 2) sweep_fast_kernel(const SweepPlan p, const FastLayout FL,
                                                                              const typename Model::Args ma,
                                                                              const StateArgs<float> sa) {
    using T = float;
    constexpr int EPV = LdTraits<U>::EPV;
    constexpr int NVT = 32 / EPV;                        // LD vectors per thread per row
    constexpr bool DP4A = std::is_same<U, int8_t>::value;
    // per-panel tile classification pays with two tiles per thread (int8); with four or eight (int16, float LD) the
    // flag arrays cost more registers and branches than they save (measured: C4 sweep 1.69 -> 1.94 ms), so those
    // types keep the general predicated form for every tile
    constexpr bool CLASSIFY = (NVT <= 2);
    extern __shared__ __align__(128) unsigned char smem[];

    SmemView<T> sm;
    sm.base = smem;
    sm.rowmeta = reinterpret_cast<int4*>(smem + FL.rowmeta);
    sm.panelmeta = reinterpret_cast<int4*>(smem + FL.panelmeta);
    sm.partial = reinterpret_cast<T*>(smem + FL.partial);
    sm.alpha = reinterpret_cast<T*>(smem + FL.alpha);
    sm.wwin = reinterpret_cast<T*>(smem + FL.wwin);
    sm.full = reinterpret_cast<uint64_t*>(smem + FL.bars);
    sm.empty = sm.full + NST_MAX;
    sm.cdone = sm.full + 2 * NST_MAX;
    sm.prog = reinterpret_cast<uint32_t*>(smem + FL.counters);
    T* fring = reinterpret_cast<T*>(smem + FL.fring);
    sm.fsrc = fring;
    sm.fmask = FR - 1;
    long long* hc = reinterpret_cast<long long*>(smem + FL.hc);    // [HP] sum of the fixed-point eta_old of the columns before vector v
    uint32_t* zvec = reinterpret_cast<uint32_t*>(smem + FL.zero);  // 16 B of zeros | 16 B of "code 0" in the biased encoding
    float* red = reinterpret_cast<float*>(smem + FL.red);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t zaddr_raw = sbase + FL.zero;          // bytes 0x00: contribute nothing to a dp4a
    const uint32_t zaddr_code = sbase + FL.zero + 16;    // code 0: contributes nothing to a decoded dot / axpy

    const int tid = threadIdx.x, warp = tid / WARP, lane = tid % WARP;
    const bool second = (int)blockIdx.x >= p.n_sm;
    const int rot_a = (p.smsp_rot == 2 && second) ? 2 : 0;
    const int rot_c = p.smsp_rot == 0 ? 0 : rot_a + 1;
    const int ai = fast_a_index(warp, rot_a), ci = fast_c_index(warp, rot_c);
    const int blk = p.blk_order[blockIdx.x];
    const int r0 = p.blk_row[blk], r1 = p.blk_row[blk + 1];
    const int B = r1 - r0;
    const int pan0 = p.blk_panel[blk];
    const int NP = p.blk_panel[blk + 1] - pan0;
    const int NST = p.nst;
    const int nvec = (B + EPV - 1) / EPV;

    // ---- prologue ----------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], NCW); mbar_init(&sm.cdone[s], 1); }
        for (int w = 0; w < NAW + NWW + NCW + 2; ++w) sm.prog[w] = 0;
        fence_mbar_init();
    }
    for (int i = tid; i < FR; i += blockDim.x) fring[i] = 0.f;
    if (tid < 4) zvec[tid] = 0u;
    else if (tid < 8) zvec[tid] = sizeof(U) == 1 ? 0x80808080u : (sizeof(U) == 2 && !std::is_floating_point<U>::value ? 0x80008000u : 0u);
    // eta_old of the thread's columns -> registers (A warps); int8: block-scaled balanced base-128 digits
    [[maybe_unused]] uint32_t hl[NVT][DP4A ? NLIMB : 1][4];
    [[maybe_unused]] float es[DP4A ? 1 : NVT][DP4A ? 1 : EPV];
    [[maybe_unused]] double wscale = 1.0;                // value of one fixed-point unit: bscale * 2^-(7 NLIMB - 1)
    if constexpr (INCR) {
        // no backward dots: nothing of eta_old is needed
    } else if constexpr (DP4A) {
        float mx = 0.f;
        for (int i = tid; i < B; i += blockDim.x) mx = fmaxf(mx, fabsf(sa.eta[r0 + i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = 0.f;
        for (int w = 0; w < (int)blockDim.x / WARP; ++w) mx = fmaxf(mx, red[w]);
        int ex = 0;
        if (mx > 0.f) frexpf(mx, &ex);                   // mx = m * 2^ex, m in [0.5, 1): block scale 2^ex > max |eta_old|
        const float qs = ldexpf(1.f, 7 * NLIMB - 1 - ex);        // eta -> fixed point Q, |Q| <= 2^(7 NLIMB - 1); exact scaling
        wscale = ldexp(1.0, ex - (7 * NLIMB - 1));
        if (ai >= 0) {
#pragma unroll
            for (int c = 0; c < NVT; ++c) {
                const int v = 32 * fast_tile(ai, c) + lane;
                long long qsum = 0;
#pragma unroll
                for (int l = 0; l < NLIMB; ++l)
#pragma unroll
                    for (int w = 0; w < 4; ++w) hl[c][l][w] = 0;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int col = v * 16 + e;
                    const float x = (col < B) ? sa.eta[r0 + col] : 0.f;
                    int Q = __float2int_rn(x * qs);
                    qsum += Q;
#pragma unroll
                    for (int l = NLIMB - 1; l >= 0; --l) {           // Q = sum_l d_l 128^(NLIMB-1-l), d_l in [-64, 64]
                        int d;
                        if (l > 0) { d = ((Q + 64) & 127) - 64; Q = (Q - d) >> 7; } else { d = Q; }
                        hl[c][l][e >> 2] |= (uint32_t)(d & 0xff) << (8 * (e & 3));
                    }
                }
                if (v < HP - 1) hc[v + 1] = qsum;
            }
        }
        __syncthreads();
        if (warp == 0) {
            // inclusive scan of hc[1..256] (8 entries per lane), hc[0] = 0
            long long loc[8];
            long long run = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { run += hc[1 + lane * 8 + i]; loc[i] = run; }
            long long tot = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long n = __shfl_up_sync(0xffffffffu, tot, o);
                if (lane >= o) tot += n;
            }
            const long long excl = tot - run;
#pragma unroll
            for (int i = 0; i < 8; ++i) hc[1 + lane * 8 + i] = loc[i] + excl;
            if (lane == 0) hc[0] = 0;
        }
    } else {
        if (ai >= 0) {
#pragma unroll
            for (int c = 0; c < NVT; ++c) {
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    const int col = (32 * fast_tile(ai, c) + lane) * EPV + e;
                    es[c][e] = (col < B) ? sa.eta[r0 + col] : 0.f;
                }
            }
        }
    }
    __syncthreads();

    [[maybe_unused]] OutRings orr;
    orr.a_xown = sbase + FL.xown; orr.a_bsum = sbase + FL.bsum;
    orr.chain_panels = sm.prog + NAW + NWW + NCW; orr.out_rows = sm.prog + NAW + NWW + NCW + 1;
#ifdef VB_CHAIN_SPLIT
    // timing experiment: the second CTA of an SM (launch order: block ids >= SM count) swaps its chain / producer warps
    const int swap = ((int)blockIdx.x >= VB_CHAIN_SPLIT) ? 1 : 0;
    const int chain_warp = swap ? FAST_PRODUCER_WARP : FAST_CHAIN_WARP, producer_warp = swap ? FAST_CHAIN_WARP : FAST_PRODUCER_WARP;
#else
    constexpr int chain_warp = FAST_CHAIN_WARP, producer_warp = FAST_PRODUCER_WARP;
#endif
    if (warp == producer_warp) {
        producer_role<U>(p, smem, sm.rowmeta, sm.panelmeta, sm.full, sm.empty, r0, pan0, NP, lane,
                         reinterpret_cast<int*>(smem + FL.rowbase), reinterpret_cast<int4*>(smem + FL.panelmeta2));
    } else if (VER == 2 && warp == FAST_OUTPUT_WARP) {
        if constexpr (VER == 2) output_role<T, Model>(p, ma, sa, orr, r0, pan0, NP, lane, blk);
    } else if (warp == chain_warp) {
        if constexpr (VER == 2) chain_role2<T, Model, NAW, NCW>(p, ma, sa, sm, orr, r0, B, pan0, NP, lane);
        else chain_role<T, Model, NAW, NCW, NWW, T, INCR>(p, ma, sa, sm, r0, B, pan0, NP, lane);
    } else if (ai >= 0) {
        // =============================== A: backward dots =====================================
        // Per panel every owned tile is classified (warp-uniform): dead (no row of the panel stores it), interior
        // (every row stores all of it: address = row base + own offset, no predicates) or boundary (the general,
        // branch-free form: a (row, vector) pair outside the row's range reads a 16-byte zero vector).
        const int wa = ai;
        uint32_t zaddr = DP4A ? zaddr_raw : zaddr_code;
        uint32_t a_rowbase = sbase + FL.rowbase, a_pm2 = sbase + FL.panelmeta2, a_rowmeta = sbase + FL.rowmeta;
#ifdef VB_PIN_BASES
        // experiment: force the ring bases to stay in registers (ptxas otherwise rebuilds some of them in front of their
        // uses: S2UR SR_CgaCtaId -> ULEA -> LDC layout field -> IADD3).  Measured: C2 sweep 1.005 ms pinned vs 0.985 ms
        // left alone -- the registers are worth more to the scheduler than the rebuilt addresses cost.
        asm volatile("" : "+r"(a_rowbase), "+r"(a_pm2), "+r"(a_rowmeta), "+r"(zaddr));
#endif
        int vown[NVT];
        uint32_t vaddr[NVT];
#pragma unroll
        for (int c = 0; c < NVT; ++c) { vown[c] = 32 * fast_tile(wa, c) + lane; vaddr[c] = sbase + (uint32_t)vown[c] * 16u; }
        int s = 0, k = 0;
        for (int u = 0; u < NP; ++u) {
            trace_ev(p, lane, wa, 0, u);
            mbar_wait(&sm.full[s], k & 1);
            trace_ev(p, lane, wa, 1, u);
            const int4 pm = sm.panelmeta[s];
            const uint4 pm2 = lds128(a_pm2 + s * 16);
            const int P = pm.x, vmin = pm.y, vmax = pm.z, jl0 = pm.w;
            const int lo_all = (int)pm2.x, hi_all = (int)pm2.y;
            if constexpr (VER == 2) {
                if constexpr (DP4A) window_panel2_i8<NAW>(sbase, sbase + FL.rowmeta, sbase + FL.wwin, jl0, P, wa, lane);
                else window_panel2<U, NAW>(sbase, sbase + FL.rowmeta, sbase + FL.wwin, zaddr_code, jl0, P, wa, lane);
            } else {
                if constexpr (DP4A) window_panel_i8<NAW>(sbase, sbase + FL.rowmeta, sbase + FL.wwin, jl0, P, wa, lane);
                else window_panel<U, NAW>(sbase, sbase + FL.rowmeta, sbase + FL.wwin, zaddr_code, jl0, P, wa, lane);
            }
            trace_ev(p, lane, wa, 3, u);
            const bool quad = ((jl0 | P) & 3) == 0;
            bool live[NVT], inter[NVT];
            bool anyl = false, anyb = false;
#pragma unroll
            for (int c = 0; c < NVT; ++c) {
                const int tb = 32 * fast_tile(wa, c);
                if constexpr (CLASSIFY) {
                    live[c] = (tb + 32 > vmin) && (tb < vmax);
                    inter[c] = quad && (tb >= lo_all) && (tb + 32 <= hi_all);
                } else {
#ifndef VB_NO_LIVE_ALL
                    live[c] = (tb + 32 > vmin) && (tb < vmax); inter[c] = false;     // dead tiles skipped, general form otherwise
#else
                    live[c] = true; inter[c] = false;        // general form only (see CLASSIFY)
#endif
                }
                anyl |= live[c];
                anyb |= live[c] && !inter[c];
            }
            [[maybe_unused]] int dig[DP4A ? NLIMB : 1];       // lane r: digit totals of row r of the panel
#pragma unroll
            for (int l = 0; l < (DP4A ? NLIMB : 1); ++l) dig[l] = 0;
#ifdef VB_WHATIF_SKIP_A
            anyl = false;                    // timing experiment: no backward-dot arithmetic (results are wrong)
#endif
            if constexpr (INCR) anyl = false;
            if (!anyl) {
                if constexpr (!DP4A || INCR) {
                    if (lane < P) sm.partial[wa * RR + ((jl0 + lane) & (RR - 1))] = 0.f;
                }
            } else {
#pragma unroll kGroupUnroll
                for (int rg = 0; rg < P; rg += 4) {
                    const int nv = min(4, P - rg);
                    int mx[4];
                    [[maybe_unused]] int my[4], mz[4];
#ifdef VB_WHATIF_PAD
                    // timing experiment: VB_WHATIF_PAD never-executed instructions in the middle of the row-group loop
                    // (how sensitive is the sweep to the size of a role's loop body?)
                    if (p.n_blocks < 0) {
#pragma unroll
                        for (int z = 0; z < VB_WHATIF_PAD; ++z) asm volatile("nanosleep.u32 1;");
                    }
#endif
#ifdef VB_TRACE2
                    const bool tr2 = (wa == 3 && rg == 4);
                    if (tr2) trace_ev(p, lane, 10, 0, u);
#endif
                    if (anyb) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            mx[r] = 0; my[r] = 0; mz[r] = 0;
                            if (r < nv) {
                                const uint4 m = lds128(a_rowmeta + (uint32_t)((jl0 + rg + r) & (RR - 1)) * 16u);
                                mx[r] = (int)m.x; my[r] = (int)m.y; mz[r] = (int)m.z;
                            }
                        }
                    } else {
                        const uint4 b = lds128(a_rowbase + (uint32_t)((jl0 + rg) & (RR - 1)) * 4u);
                        mx[0] = (int)b.x; mx[1] = (int)b.y; mx[2] = (int)b.z; mx[3] = (int)b.w;
#pragma unroll
                        for (int r = 0; r < 4; ++r) { my[r] = 0; mz[r] = 0; }
                    }
#ifdef VB_TRACE2
                    if (tr2) { asm volatile("" ::"r"(mx[0]), "r"(mx[3])); trace_ev(p, lane, 10, 1, u); }
#endif
                    [[maybe_unused]] int acc[4][DP4A ? NLIMB : 1];
                    [[maybe_unused]] typename Pk<T>::acc_t acc2[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        acc2[r] = Pk<T>::zero();
#pragma unroll
                        for (int l = 0; l < (DP4A ? NLIMB : 1); ++l) acc[r][l] = 0;
                    }
#pragma unroll
                    for (int c = 0; c < NVT; ++c) {
#ifdef VB_TRACE2
                        if (tr2 && c == NVT - 1) trace_ev(p, lane, 10, 6, u);
#endif
                        if (!live[c]) continue;
                        uint32_t ad[4];
                        if (inter[c]) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) ad[r] = vaddr[c] + (uint32_t)mx[r];
                        } else {
                            const int v = vown[c];
                            bool any = false;
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const bool in = (v >= my[r]) && (v < mz[r]);
                                any |= in;
                                ad[r] = in ? vaddr[c] + (uint32_t)mx[r] : zaddr;
                            }
                            if (!__any_sync(0xffffffffu, any)) continue;     // no lane of the warp owns a live vector here
                        }
#ifdef VB_TRACE2
                        if (tr2 && c == NVT - 1) { asm volatile("" ::"r"(ad[0]), "r"(ad[3])); trace_ev(p, lane, 10, 7, u); }
#endif
                        uint4 cv[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) cv[r] = lds128(ad[r]);
#ifdef VB_TRACE2
                        if (tr2 && c == NVT - 1) { trace_ev(p, lane, 10, 2, u); asm volatile("" ::"r"(cv[0].x), "r"(cv[3].w)); trace_ev(p, lane, 10, 3, u); }
#endif
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if constexpr (DP4A) {
#pragma unroll
                                for (int l = 0; l < NLIMB; ++l) {
                                    int a = acc[r][l];
                                    a = dp4a_us(cv[r].x, hl[c][l][0], a);
                                    a = dp4a_us(cv[r].y, hl[c][l][1], a);
                                    a = dp4a_us(cv[r].z, hl[c][l][2], a);
                                    a = dp4a_us(cv[r].w, hl[c][l][3], a);
                                    acc[r][l] = a;
                                }
                            } else {
                                VecOps<T, U>::dot(cv[r], es[c], acc2[r]);
                            }
                        }
                    }
                    if (rg == 0) trace_ev(p, lane, wa, 4, u);
#ifdef VB_TRACE2
                    if constexpr (DP4A) {
                        if (tr2) { asm volatile("" ::"r"(acc[0][0]), "r"(acc[3][NLIMB - 1]), "r"(acc[1][1]), "r"(acc[2][0])); trace_ev(p, lane, 10, 4, u); }
                    }
#endif
                    if constexpr (DP4A) {
                        // exact integer totals: one REDUX per (row, digit); lane (rg + r) keeps the digits of row rg + r
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
#pragma unroll
                            for (int l = 0; l < NLIMB; ++l) {
                                const int tot = __reduce_add_sync(0xffffffffu, acc[r][l]);
                                if (lane == rg + r) dig[l] = tot;
                            }
                        }
                    } else {
                        T acc1[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc1[r] = Pk<T>::sum(acc2[r]);
                        const int rr = warp_reduce4(acc1, lane);
                        if ((lane & 7) == 0 && rr < nv) sm.partial[wa * RR + ((jl0 + rg + rr) & (RR - 1))] = acc1[0];
                    }
                    if (rg == 0) trace_ev(p, lane, wa, 5, u);
#ifdef VB_TRACE2
                    if constexpr (DP4A) {
                        if (tr2) { asm volatile("" ::"r"(dig[0]), "r"(dig[NLIMB - 1])); trace_ev(p, lane, 10, 5, u); }
                    }
#endif
                }
            }
            trace_ev(p, lane, wa, 6, u);
            if constexpr (DP4A && !INCR) {
                // lane r < P: recombine the digits of row r in int64.  sum_b u8 * digit = sum code * digit +
                // 128 * sum digit over the row's packed range; the range term (128 * sum of Q over the range, from
                // the prefix array) is removed once per row, by A warp 0.
                if (lane < P) {
                    long long x = 0;
#pragma unroll
                    for (int l = 0; l < NLIMB; ++l) x = x * 128 + (long long)dig[l];
                    if (wa == 0) {
                        const int4 m = sm.rowmeta[(jl0 + lane) & (RR - 1)];
                        x -= 128 * (hc[min(m.z, HP - 1)] - hc[min(m.y, HP - 1)]);
                    }
                    sm.partial[wa * RR + ((jl0 + lane) & (RR - 1))] = (float)((double)x * wscale);
                }
            }
            __syncwarp();
            if (lane == 0) st_release(&sm.prog[wa], (uint32_t)(u + 1));
            trace_ev(p, lane, wa, 2, u);
            if (++s == NST) { s = 0; ++k; }
        }
    } else {
        // =============================== C: forward axpy ======================================
        const int wc = ci;
        uint32_t a_rowbase = sbase + FL.rowbase, a_pm2 = sbase + FL.panelmeta2, a_alpha = sbase + FL.alpha,
                 a_rowmeta = sbase + FL.rowmeta, zc = zaddr_code;
#ifdef VB_PIN_BASES
        asm volatile("" : "+r"(a_rowbase), "+r"(a_pm2), "+r"(a_alpha), "+r"(a_rowmeta), "+r"(zc));
#endif
        int vown[NVT];
        uint32_t vaddr[NVT];
#pragma unroll
        for (int c = 0; c < NVT; ++c) { vown[c] = 32 * fast_tile(wc, c) + lane; vaddr[c] = sbase + (uint32_t)vown[c] * 16u; }
        float f[NVT][EPV];
#pragma unroll
        for (int c = 0; c < NVT; ++c)
#pragma unroll
            for (int e = 0; e < EPV; ++e) f[c][e] = 0.f;
        int pubvec = 0;                                      // vectors below this index have been published
        int s = 0, k = 0;
        for (int v = 0; v < NP; ++v) {
            mbar_wait(&sm.cdone[s], k & 1);
            trace_ev(p, lane, 4 + wc, 4, v);
            const int4 pm = sm.panelmeta[s];
            const uint4 pm2 = lds128(a_pm2 + s * 16);
            const int Pc = pm.x, vmin = pm.y, vmax = pm.z, jl0 = pm.w;
            const int cut0 = (jl0 + WIN + EPV - 1) / EPV;                // first vector the bulk owns of the first row
            const int cutl = (jl0 + Pc - 1 + WIN + EPV - 1) / EPV;       // ... of the last row
            const int lo_c = max((int)pm2.x, cutl), hi_c = (int)pm2.y;
            const bool quad = ((jl0 | Pc) & 3) == 0;
            bool live[NVT], inter[NVT];
            bool anyl = false, anyb = false;
#pragma unroll
            for (int c = 0; c < NVT; ++c) {
                const int tb = 32 * fast_tile(wc, c);
                if constexpr (CLASSIFY) {
                    live[c] = (tb + 32 > max(vmin, cut0)) && (tb < vmax);
                    inter[c] = quad && (tb >= lo_c) && (tb + 32 <= hi_c);
                } else {
#ifndef VB_NO_LIVE_ALL
                    live[c] = (tb + 32 > max(vmin, cut0)) && (tb < vmax); inter[c] = false;
#else
                    live[c] = true; inter[c] = false;
#endif
                }
                anyl |= live[c];
                anyb |= live[c] && !inter[c];
            }
#ifdef VB_WHATIF_SKIP_C
            anyl = false;                    // timing experiment: no forward-axpy arithmetic (results are wrong)
#endif
            if (anyl) {
#pragma unroll kGroupUnroll
                for (int rg = 0; rg < Pc; rg += 4) {
                    const int nv = min(4, Pc - rg);
                    int mx[4];
                    [[maybe_unused]] int my[4], mz[4];
                    T al[4];
                    if (anyb) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            mx[r] = 0; my[r] = 0; mz[r] = 0; al[r] = T(0);
                            if (r < nv) {
                                const int jl = jl0 + rg + r;
                                const uint4 m = lds128(a_rowmeta + (uint32_t)(jl & (RR - 1)) * 16u);
                                mx[r] = (int)m.x; my[r] = max((int)m.y, (jl + WIN + EPV - 1) / EPV); mz[r] = (int)m.z;
                                al[r] = lds_t(a_alpha + (uint32_t)(jl & (RR - 1)) * 4u, T());
                            }
                        }
                    } else {
                        const uint32_t slot = (uint32_t)((jl0 + rg) & (RR - 1)) * 4u;
                        const uint4 b = lds128(a_rowbase + slot);
                        const uint4 a4 = lds128(a_alpha + slot);
                        mx[0] = (int)b.x; mx[1] = (int)b.y; mx[2] = (int)b.z; mx[3] = (int)b.w;
                        al[0] = __uint_as_float(a4.x); al[1] = __uint_as_float(a4.y);
                        al[2] = __uint_as_float(a4.z); al[3] = __uint_as_float(a4.w);
#pragma unroll
                        for (int r = 0; r < 4; ++r) { my[r] = 0; mz[r] = 0; }
                    }
#pragma unroll
                    for (int c = 0; c < NVT; ++c) {
                        if (!live[c]) continue;
                        uint32_t ad[4];
                        if (inter[c]) {
#pragma unroll
                            for (int r = 0; r < 4; ++r) ad[r] = vaddr[c] + (uint32_t)mx[r];
                        } else {
                            const int vv = vown[c];
                            bool any = false;
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const bool in = (vv >= my[r]) && (vv < mz[r]);
                                any |= in;
                                ad[r] = in ? vaddr[c] + (uint32_t)mx[r] : zc;
                            }
                            if (!__any_sync(0xffffffffu, any)) continue;
                        }
                        uint4 cv[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) cv[r] = lds128(ad[r]);
#pragma unroll
                        for (int r = 0; r < 4; ++r) VecOps<T, U>::axpy(cv[r], al[r], f[c]);
                    }
                }
            }
            trace_ev(p, lane, 4 + wc, 6, v);
            // columns below cut(first row after the panel) are complete: publish them for the chain
            const int newpub = min(nvec, (jl0 + Pc + WIN + EPV - 1) / EPV);
#pragma unroll
            for (int c = 0; c < NVT; ++c) {
                const int vv = vown[c];
                if (vv >= pubvec && vv < newpub) {
#pragma unroll
                    for (int e = 0; e < EPV; ++e) fring[(vv * EPV + e) & (FR - 1)] = f[c][e];
                }
            }
            pubvec = max(pubvec, newpub);
            __syncwarp();
            if (lane == 0) {
                st_release(&sm.prog[NAW + NWW + wc], (uint32_t)(v + 1));
                mbar_arrive(&sm.empty[s]);
            }
            trace_ev(p, lane, 4 + wc, 5, v);
            if (++s == NST) { s = 0; ++k; }
        }
    }
}

}  // namespace vb
