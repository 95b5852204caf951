// viprs_b200 -- launch helpers shared by the per-type translation units.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ld.h"
#include "sweep.cuh"
#include "sweep_fast.cuh"

namespace vb {

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int make_plan(const viprs_b200_ld* ld, int tsize, SweepPlan& p, RingGeometry& g) {
    g = ring_geometry(ld, tsize);
    p.packed = reinterpret_cast<const unsigned char*>(ld->d_packed);
    p.prow = ld->d_prow; p.pcs = ld->d_pcs; p.blk_row = ld->d_blk_row; p.blk_panel = ld->d_blk_panel;
    p.panel_row = ld->d_panel_row; p.blk_order = ld->d_blk_order; p.panel_need = ld->d_panel_need;
    p.n_blocks = ld->n_blocks; p.stage_bytes = ld->stage_bytes; p.nst = g.nst; p.bpad = state_pad(ld->max_block);
    p.l2_ahead = env_int("VIPRS_B200_L2_AHEAD", 0);
    p.n_sm = ld->n_sm > 0 ? ld->n_sm : 148;
    p.smsp_rot = env_int("VIPRS_B200_SMSP_ROT", 1);
    p.trace = nullptr;
    p.L = make_layout(p.bpad, tsize, ld->stage_bytes, g.nst);
    return g.nst == 0 ? VIPRS_B200_EBLOCK_TOO_LARGE : VIPRS_B200_OK;
}

// run `launch(plan)`; with VIPRS_B200_TRACE=<file> (debug only) also dump the timeline of CTA 0
template <typename F>
static int launch_traced(const SweepPlan& p, cudaStream_t st, F&& launch) {
    const char* trace_path = getenv("VIPRS_B200_TRACE");
    if (trace_path) {
        SweepPlan pt = p;
        const size_t nb = (size_t)kTraceSlots * sizeof(unsigned long long);
        if (cudaMalloc(&pt.trace, nb) != cudaSuccess) return VIPRS_B200_ENOMEM;
        cudaMemsetAsync(pt.trace, 0, nb, st);
        launch(pt);
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> h(kTraceSlots);
        cudaMemcpy(h.data(), pt.trace, nb, cudaMemcpyDeviceToHost);
        cudaFree(pt.trace);
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), 1, nb, f); fclose(f); }
    } else {
        launch(p);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

template <typename T, typename U, typename Model, int MINB>
static int launch_one(const viprs_b200_ld* ld, const SweepPlan& p, const RingGeometry& g,
                      const typename Model::Args& ma, const StateArgs<T>& sa, cudaStream_t st) {
    auto kern = sweep_kernel<T, U, Model, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    return launch_traced(p, st, [&](const SweepPlan& pp) {
        kern<<<p.n_blocks, (NBW + 2) * WARP, g.smem_bytes, st>>>(pp, ma, sa);
    });
}

// register-resident kernel (float32 state, LD blocks <= 4096 SNPs, two CTAs per SM)
template <typename U, typename Model, int NLIMB, int VER, bool INCR = false>
static int launch_fast_ver(const viprs_b200_ld* ld, SweepPlan p, const typename Model::Args& ma,
                           const StateArgs<float>& sa, cudaStream_t st) {
    const RingGeometry g = fast_ring_geometry(ld);
    const FastLayout FL = make_fast_layout(ld->stage_bytes, g.nst);
    p.nst = g.nst;
    p.L.stages = FL.stages;
    auto kern = sweep_fast_kernel<U, Model, NLIMB, VER, INCR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL.total);
    if (e != cudaSuccess) return (int)e;
    return launch_traced(p, st, [&](const SweepPlan& pp) {
        kern<<<p.n_blocks, fast_threads<VER>(), FL.total, st>>>(pp, FL, ma, sa);
    });
}

// VIPRS_B200_FAST = 1 (default): chain warp gathers its window coefficients per step and writes the outputs itself.
// 2: constant-stride window layout, two accumulators per chain lane switched at 32-row boundaries, an 11th warp writes
// the outputs and can accumulate the M-step / ELBO sums (viprs_b200_e_step_fused_f32).  Measured on the C2 workload
// (B200, profiles/r02_c2_fast_versions.txt): the chain's share of a panel drops from ~2,600 to ~1,400 cycles, but the
// sweep is bound by the A / C warps, not by the chain -- 1.021 ms (1) vs 1.053 ms (2), 1.19 ms with the fused sums.
template <typename U, typename Model, int NLIMB>
static int launch_fast_one(const viprs_b200_ld* ld, SweepPlan p, const typename Model::Args& ma,
                           const StateArgs<float>& sa, cudaStream_t st) {
    if (env_int("VIPRS_B200_FAST", 1) == 2) return launch_fast_ver<U, Model, NLIMB, 2>(ld, p, ma, sa, st);
    return launch_fast_ver<U, Model, NLIMB, 1>(ld, p, ma, sa, st);
}

template <typename T, typename U, typename Model>
static bool fast_path_ok(const viprs_b200_ld* ld) {
    if constexpr (sizeof(T) != 4 || Model::kHeavy || sizeof(U) == 8) return false;
    return ld->max_block <= FAST_MAX_BLOCK && fast_ring_geometry(ld).nst >= 3 && env_int("VIPRS_B200_FORCE_GENERIC", 0) == 0;
}

// A subset of the sweep units of a single-phase LD: row chunk `chunk` of the handle (ld.h), or everything (chunk < 0).
struct UnitSubset {
    const int32_t* order; int count; int item0, item1; int row0, row1;
};
static UnitSubset unit_subset(const viprs_b200_ld* ld, int chunk) {
    if (chunk < 0 || ld->n_chunks <= 0)
        return UnitSubset{ld->d_blk_order, ld->n_blocks, 0, ld->n_items_bwd, 0, ld->M};
    const int u0 = ld->h_chunk_unit[chunk], u1 = ld->h_chunk_unit[chunk + 1];
    return UnitSubset{ld->d_chunk_order + u0, u1 - u0, ld->h_chunk_item[chunk], ld->h_chunk_item[chunk + 1],
                      ld->h_chunk_row[chunk], ld->h_chunk_row[chunk + 1]};
}

// one launch over the sweep units of phase `ph` (all units when the LD has a single phase), or over `sub`
template <typename T, typename U, typename Model>
static int launch_phase(const viprs_b200_ld* ld, int ph, const typename Model::Args& ma, const StateArgs<T>& sa, cudaStream_t st,
                        const UnitSubset* sub = nullptr) {
    SweepPlan p;
    RingGeometry g;
    const int first = ld->h_phase_ptr[ph], count = sub ? sub->count : ld->h_phase_ptr[ph + 1] - first;
    if (count <= 0) return VIPRS_B200_OK;
    auto phase_of = [&](SweepPlan& q) { q.blk_order = sub ? sub->order : ld->d_blk_order + first; q.n_blocks = count; };
    if constexpr (sizeof(T) == 4 && !Model::kHeavy && sizeof(U) != 8) {
        if (fast_path_ok<T, U, Model>(ld)) {
            make_plan(ld, (int)sizeof(T), p, g);        // ring fields are overridden by the fast launcher
            phase_of(p);
            if (std::is_same<U, int8_t>::value && env_int("VIPRS_B200_LIMBS", 3) == 3)
                return launch_fast_one<U, Model, 3>(ld, p, ma, sa, st);
            return launch_fast_one<U, Model, 4>(ld, p, ma, sa, st);
        }
    }
    int rc = make_plan(ld, (int)sizeof(T), p, g);
    if (rc) return rc;
    phase_of(p);
    if constexpr (!Model::kHeavy) {
        if (g.ctas_per_sm >= 2) return launch_one<T, U, Model, 2>(ld, p, g, ma, sa, st);
    }
    return launch_one<T, U, Model, 1>(ld, p, g, ma, sa, st);
}

template <typename T, typename U>
static int launch_row_dots(const int4* items, int n_items, const void* rows, const int64_t* prow, const int32_t* pcs,
                           const T* x, T* q, T dq, cudaStream_t st) {
    if (n_items <= 0) return VIPRS_B200_OK;
    row_dot_kernel<T, U><<<n_items, BWD_THREADS, 0, st>>>(items, reinterpret_cast<const unsigned char*>(rows), prow, pcs, x, q, dq);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

// q[j] += dq * sum_{k>j} R_jk x[k] over the in-unit rows and (tiled LD) the ext rows
template <typename T, typename U>
static int launch_backward(const viprs_b200_ld* ld, const T* x, T* q, T dq, cudaStream_t st) {
    int rc = launch_row_dots<T, U>(ld->d_items_bwd, ld->n_items_bwd, ld->d_packed, ld->d_prow, ld->d_pcs, x, q, dq, st);
    if (rc == 0) rc = launch_row_dots<T, U>(ld->d_items_bwd_ext, ld->n_items_bwd_ext, ld->d_ext, ld->d_erow, ld->d_ecs, x, q, dq, st);
    return rc;
}

// out[k] += scale * sum_j R_jk x[j] over `n_items` items (see forward_axpy_kernel); max_cols: widest item
template <typename T, typename U>
static int launch_forward(const int4* items, int n_items, int max_cols, const void* rows, const int64_t* prow,
                          const int32_t* pcs, const T* x, T* out, T scale, cudaStream_t st, int max_rows = FWD_MAX_ROWS) {
    if (n_items <= 0 || max_cols <= 0) return VIPRS_B200_OK;
    constexpr int EPV = LdTraits<U>::EPV;
    dim3 grid((max_cols + FWD_THREADS * FWD_VPT * EPV - 1) / (FWD_THREADS * FWD_VPT * EPV), n_items);
    // an item's rows are those of one sweep unit: at most FWD_MAX_ROWS (= kTileLimit)
    forward_axpy_kernel<T, U><<<grid, FWD_THREADS, forward_smem_bytes(max_rows, (int)sizeof(T)), st>>>(
        items, reinterpret_cast<const unsigned char*>(rows), prow, pcs, x, out, scale);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

// side stream + events of a tiled LD handle, created on first use
static bool side_stream_ready(const viprs_b200_ld* ld) {
    if (ld->side_stream) return true;
    if (cudaStreamCreateWithFlags(&ld->side_stream, cudaStreamNonBlocking) != cudaSuccess) { ld->side_stream = nullptr; return false; }
    ld->side_events.assign(ld->n_phases + 1, nullptr);
    for (auto& ev : ld->side_events) {
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            for (cudaEvent_t x : ld->side_events) if (x) cudaEventDestroy(x);
            ld->side_events.clear();
            cudaStreamDestroy(ld->side_stream);
            ld->side_stream = nullptr;
            return false;
        }
    }
    return true;
}

template <typename T>
__global__ void scale_copy_kernel(int M, const T* __restrict__ src, T scale, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) dst[i] = src ? src[i] * scale : T(0);
}

// One Gauss-Seidel sweep over every LD block.  LD blocks of up to kTileLimit rows are one sweep unit and one launch
// covers them all.  Larger blocks are tiled (ld.cu): tile p of every block is swept by launch p with the strictly
// sequential one-pass kernel, and the rectangles between tiles enter as two streaming matrix-vector products,
//   bext[j]  = sum_{k in later tiles} R_jk eta_old[k]      (once, before anything is swept), and
//   fext[k] += sum_{j in tile p}      R_jk eta_new[j]      (after launch p, for the columns of the later tiles),
// so the order of the per-SNP updates -- and therefore the result -- is that of the reference's single chain.
// q_offset (nullable, q units): a caller-supplied constant part of q (see viprs_b200_q_offset_*).
template <typename T, typename U, typename Model>
static int launch_sweep(const viprs_b200_ld* ld, const typename Model::Args& ma, StateArgs<T> sa, const T* q_offset, T dq,
                        cudaStream_t st, const UnitSubset* sub = nullptr) {
    if (ld->n_phases == 1 && ld->ext_elems == 0) {
        if (q_offset) { sa.fext = q_offset; sa.fscale = T(1) / dq; }
        return launch_phase<T, U, Model>(ld, 0, ma, sa, st, sub);
    }
    if (sub) return VIPRS_B200_EUNSUPPORTED;
    T* fext = reinterpret_cast<T*>(ld->d_fext);
    T* bext = reinterpret_cast<T*>(ld->d_bext);
    const int M = ld->M;
    scale_copy_kernel<T><<<(M + 255) / 256, 256, 0, st>>>(M, q_offset, T(1) / dq, fext);
    cudaError_t e = cudaMemsetAsync(bext, 0, sizeof(T) * (size_t)M, st);
    if (e != cudaSuccess) return (int)e;
    // The backward-external dots of tile p only have to be complete before launch p, and the sequential launches keep
    // one CTA per LD block busy (73 of 148 SMs at the C5 shape): the dots run phase by phase on a side stream, next to
    // the sweeps and forward products of the earlier tiles (they read eta of later tiles only, which those do not
    // write).  The side stream is forked from and joined back into `st` with events, so the whole sweep stays
    // capturable in a CUDA graph.
    int rc = VIPRS_B200_OK;
    const bool side = ld->n_phases > 1 && env_int("VIPRS_B200_NO_SIDE_STREAM", 0) == 0 && side_stream_ready(ld);
    if (side) {
        e = cudaEventRecord(ld->side_events[0], st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ld->side_stream, ld->side_events[0], 0);
        for (int ph = 0; ph < ld->n_phases && rc == 0 && e == cudaSuccess; ++ph) {
            const int i0 = ld->h_bwd_ext_phase_ptr[ph], i1 = ld->h_bwd_ext_phase_ptr[ph + 1];
            rc = launch_row_dots<T, U>(ld->d_items_bwd_ext + i0, i1 - i0, ld->d_ext, ld->d_erow, ld->d_ecs, sa.eta, bext, T(1),
                                       ld->side_stream);
            if (rc == 0) e = cudaEventRecord(ld->side_events[ph + 1], ld->side_stream);
        }
        if (e != cudaSuccess) return (int)e;
    } else {
        rc = launch_row_dots<T, U>(ld->d_items_bwd_ext, ld->n_items_bwd_ext, ld->d_ext, ld->d_erow, ld->d_ecs, sa.eta, bext, T(1), st);
    }
    sa.fext = fext; sa.bext = bext; sa.fscale = T(1);
    for (int ph = 0; ph < ld->n_phases && rc == 0; ++ph) {
        if (side) {
            e = cudaStreamWaitEvent(st, ld->side_events[ph + 1], 0);
            if (e != cudaSuccess) return (int)e;
        }
        rc = launch_phase<T, U, Model>(ld, ph, ma, sa, st);
        const int i0 = ld->h_ext_phase_ptr[ph], i1 = ld->h_ext_phase_ptr[ph + 1];
        if (rc == 0 && i1 > i0)
            rc = launch_forward<T, U>(ld->d_items_ext + i0, i1 - i0, ld->h_items_cols[ph], ld->d_ext, ld->d_erow, ld->d_ecs,
                                      sa.eta, fext, T(1), st, ld->max_block);
    }
    return rc;
}

// out = q - dq * (R - I) eta : the part of a caller's q that is NOT explained by eta.  The reference maintains q
// incrementally (e_step.hpp:421,439), so whatever q holds beyond dq (R - I) eta on entry -- e.g. q = 0 next to
// eta != 0 after a `param_0` warm start, VIPRS.py:339-357 -- stays in q for ever.  Passing this vector as
// `q_offset` to the sweeps reproduces that.
template <typename T, typename U>
static int launch_q_offset(const viprs_b200_ld* ld, const T* eta, const T* q, T dq, T* out, cudaStream_t st) {
    cudaError_t e = cudaMemcpyAsync(out, q, sizeof(T) * (size_t)ld->M, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
    int rc = launch_backward<T, U>(ld, eta, out, -dq, st);
    if (rc == 0) rc = launch_forward<T, U>(ld->d_items_diag, ld->n_blocks, ld->h_items_cols[ld->n_phases], ld->d_packed,
                                          ld->d_prow, ld->d_pcs, eta, out, -dq, st);
    // the ext rectangles of different tiles of one LD block overlap in columns: one launch per phase (inside a phase
    // every item belongs to a different LD block, so no two CTAs update the same column)
    for (int ph = 0; ph < ld->n_phases && rc == 0; ++ph) {
        const int i0 = ld->h_ext_phase_ptr[ph], i1 = ld->h_ext_phase_ptr[ph + 1];
        if (i1 > i0)
            rc = launch_forward<T, U>(ld->d_items_ext + i0, i1 - i0, ld->h_items_cols[ph], ld->d_ext, ld->d_erow, ld->d_ecs,
                                      eta, out, -dq, st);
    }
    return rc;
}

// dispatch over the LD storage type; F(U tag) -> int
template <typename T, typename F>
static int for_ld_dtype(const viprs_b200_ld* ld, F&& f) {
    switch (ld->ld_dtype) {
        case VIPRS_B200_I8: return f(int8_t{});
        case VIPRS_B200_I16: return f(int16_t{});
        case VIPRS_B200_F32: return f(float{});
        case VIPRS_B200_F64:
            if constexpr (sizeof(T) == 8) return f(double{});
            return VIPRS_B200_EUNSUPPORTED;       // float64 LD with float32 state: the reference never builds it
    }
    return VIPRS_B200_EUNSUPPORTED;
}

template <typename T>
static int e_step_dispatch(const viprs_b200_ld* ld, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                           T* eta_diff, const T* u_logs, const T* shvt, const T* mu_mult, T dq,
                           int materialize_q, const T* q_offset, cudaStream_t st) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !shvt || !mu_mult)
        return VIPRS_B200_EINVAL;
    typename SlabModel<T>::Args ma{std_beta, u_logs, shvt, mu_mult, var_gamma, var_mu, dq};
    StateArgs<T> sa{eta, q, eta_diff};
    return for_ld_dtype<T>(ld, [&](auto tag) {
        using U = decltype(tag);
        int rc = launch_sweep<T, U, SlabModel<T>>(ld, ma, sa, q_offset, dq, st);
        if (rc == 0 && materialize_q) rc = launch_backward<T, U>(ld, eta, q, dq, st);
        return rc;
    });
}

// The reference's own q bookkeeping on device arrays (cpp_e_step / cpp_e_step_mixture semantics exactly, q in/out):
//   sweep:  q_j used at step j = q_in[j] + dq sum_{i<j} R_ij eta_diff_i           (e_step.hpp:421)
//   after:  q[j] += dq sum_{k>j} R_jk eta_diff_k                                  (update_q_factor, :435-440)
// Register-resident kernel only (float32 state, LD blocks <= 4096 SNPs, one phase): no backward dots inside the sweep,
// two reads of the LD per call like the reference.  `chunk` >= 0 restricts the call to one row chunk of the handle.
template <typename U, typename Model>
static int launch_incremental(const viprs_b200_ld* ld, const typename Model::Args& ma, StateArgs<float> sa, float dq, int chunk,
                              cudaStream_t st, cudaEvent_t swept = nullptr) {
    using T = float;
    if constexpr (Model::kHeavy || sizeof(U) == 8) {
        return VIPRS_B200_EUNSUPPORTED;
    } else {
        if (!fast_path_ok<T, U, Model>(ld) || ld->n_phases != 1 || ld->ext_elems != 0) return VIPRS_B200_EUNSUPPORTED;
        if (chunk >= ld->n_chunks && chunk >= 0) return VIPRS_B200_EINVAL;
        const UnitSubset sub = unit_subset(ld, chunk);
        if (sub.count <= 0) return VIPRS_B200_OK;
        SweepPlan p;
        RingGeometry g;
        make_plan(ld, (int)sizeof(T), p, g);
        p.blk_order = sub.order; p.n_blocks = sub.count;
        sa.fext = sa.q; sa.fscale = T(1) / dq;        // q_in: read by the chain 32 rows ahead of where it writes q
        int rc;
        if (std::is_same<U, int8_t>::value) rc = launch_fast_ver<U, Model, 3, 1, true>(ld, p, ma, sa, st);
        else rc = launch_fast_ver<U, Model, 4, 1, true>(ld, p, ma, sa, st);
        if (rc == 0 && swept != nullptr) {
            const cudaError_t e = cudaEventRecord(swept, st);
            if (e != cudaSuccess) return (int)e;
        }
        if (rc == 0)
            rc = launch_row_dots<T, U>(ld->d_items_bwd + sub.item0, sub.item1 - sub.item0, ld->d_packed, ld->d_prow, ld->d_pcs,
                                       sa.eta_diff, sa.q, dq, st);
        return rc;
    }
}

template <typename T>
static int e_step_incremental_dispatch(const viprs_b200_ld* ld, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                                       T* eta_diff, const T* u_logs, const T* shvt, const T* mu_mult, T dq, int chunk,
                                       cudaStream_t st, cudaEvent_t swept = nullptr) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !shvt || !mu_mult)
        return VIPRS_B200_EINVAL;
    if constexpr (sizeof(T) != 4) {
        return VIPRS_B200_EUNSUPPORTED;
    } else {
        typename SlabModel<T>::Args ma{std_beta, u_logs, shvt, mu_mult, var_gamma, var_mu, dq};
        StateArgs<T> sa{eta, q, eta_diff};
        return for_ld_dtype<T>(ld, [&](auto tag) { return launch_incremental<decltype(tag), SlabModel<T>>(ld, ma, sa, dq, chunk, st, swept); });
    }
}

template <typename T>
static int mixture_incremental_dispatch(const viprs_b200_ld* ld, int K, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                                        T* eta_diff, const T* log_null_pi, const T* u_logs, const T* shvt, const T* mu_mult,
                                        T dq, int chunk, cudaStream_t st, cudaEvent_t swept = nullptr) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !log_null_pi || !u_logs || !shvt || !mu_mult)
        return VIPRS_B200_EINVAL;
    if constexpr (sizeof(T) != 4) {
        return VIPRS_B200_EUNSUPPORTED;
    } else {
        if (K < 1 || K > 4) return VIPRS_B200_EUNSUPPORTED;
        typename MixModel<T, 4>::Args ma{std_beta, u_logs, shvt, mu_mult, log_null_pi, var_gamma, var_mu, dq, K};
        StateArgs<T> sa{eta, q, eta_diff};
        return for_ld_dtype<T>(ld, [&](auto tag) { return launch_incremental<decltype(tag), MixModel<T, 4>>(ld, ma, sa, dq, chunk, st, swept); });
    }
}

// Sweep with the M-step / ELBO reductions fused into its output role (register-resident kernel version 2, float32
// state, spike-and-slab, no q offset): one launch per phase writes every unit's partial sums, a one-warp-per-slot
// kernel folds them per chromosome segment into sums[nseg][VIPRS_B200_NSUMS].  VIPRS_B200_EUNSUPPORTED when the
// configuration is not covered (callers then run the sweep and viprs_b200_sums_* separately).
template <typename T>
static int e_step_fused_dispatch(const viprs_b200_ld* ld, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                                 T* eta_diff, const T* u_logs, const T* shvt, const T* mu_mult, T dq,
                                 const double* n_per_snp, const double* theta, int nseg, const int32_t* seg_ptr,
                                 double* sums, cudaStream_t st) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !shvt || !mu_mult ||
        !n_per_snp || !theta || !seg_ptr || !sums || nseg <= 0)
        return VIPRS_B200_EINVAL;
    if constexpr (sizeof(T) != 4) {
        return VIPRS_B200_EUNSUPPORTED;
    } else {
        if (env_int("VIPRS_B200_FAST", 1) != 2 || env_int("VIPRS_B200_FUSED_SUMS", 1) == 0) return VIPRS_B200_EUNSUPPORTED;
        typename SlabModel<T>::Args ma{std_beta, u_logs, shvt, mu_mult, var_gamma, var_mu, dq};
        StateArgs<T> sa{eta, q, eta_diff};
        sa.n_per_snp = n_per_snp; sa.theta = reinterpret_cast<const Theta*>(theta);
        sa.unit_partial = reinterpret_cast<double*>(ld->d_unit_partial);
        return for_ld_dtype<T>(ld, [&](auto tag) {
            using U = decltype(tag);
            if constexpr (sizeof(U) == 8) {
                return (int)VIPRS_B200_EUNSUPPORTED;
            } else {
                if (!fast_path_ok<T, U, SlabModel<T>>(ld)) return (int)VIPRS_B200_EUNSUPPORTED;
                int rc = launch_sweep<T, U, SlabModel<T>>(ld, ma, sa, nullptr, dq, st);
                if (rc) return rc;
                reduce_units_kernel<T><<<nseg, WARP * NS, 0, st>>>(ld->n_blocks, ld->d_blk_row, seg_ptr,
                                                                   reinterpret_cast<const double*>(ld->d_unit_partial), sums);
                const cudaError_t e = cudaGetLastError();
                return e == cudaSuccess ? (int)VIPRS_B200_OK : (int)e;
            }
        });
    }
}

template <typename T>
static int q_offset_dispatch(const viprs_b200_ld* ld, const T* eta, const T* q, T dq, T* out, cudaStream_t st) {
    if (!ld || !eta || !q || !out) return VIPRS_B200_EINVAL;
    return for_ld_dtype<T>(ld, [&](auto tag) { return launch_q_offset<T, decltype(tag)>(ld, eta, q, dq, out, st); });
}

template <typename T>
static int mixture_dispatch(const viprs_b200_ld* ld, int K, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                            T* eta_diff, const T* log_null_pi, const T* u_logs, const T* shvt, const T* mu_mult, T dq,
                            int materialize_q, const T* q_offset, cudaStream_t st) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !log_null_pi || !u_logs || !shvt ||
        !mu_mult)
        return VIPRS_B200_EINVAL;
    if (K < 1 || K > 16) return VIPRS_B200_EUNSUPPORTED;
    StateArgs<T> sa{eta, q, eta_diff};
    return for_ld_dtype<T>(ld, [&](auto tag) {
        using U = decltype(tag);
        int rc;
        if (K <= 4) {
            typename MixModel<T, 4>::Args ma{std_beta, u_logs, shvt, mu_mult, log_null_pi, var_gamma, var_mu, dq, K};
            rc = launch_sweep<T, U, MixModel<T, 4>>(ld, ma, sa, q_offset, dq, st);
        } else if (K <= 8) {
            typename MixModel<T, 8>::Args ma{std_beta, u_logs, shvt, mu_mult, log_null_pi, var_gamma, var_mu, dq, K};
            rc = launch_sweep<T, U, MixModel<T, 8>>(ld, ma, sa, q_offset, dq, st);
        } else {
            typename MixModel<T, 16>::Args ma{std_beta, u_logs, shvt, mu_mult, log_null_pi, var_gamma, var_mu, dq, K};
            rc = launch_sweep<T, U, MixModel<T, 16>>(ld, ma, sa, q_offset, dq, st);
        }
        if (rc == 0 && materialize_q) rc = launch_backward<T, U>(ld, eta, q, dq, st);
        return rc;
    });
}

template <typename T>
static int backward_dispatch(const viprs_b200_ld* ld, const T* x, T* q, T dq, cudaStream_t st) {
    if (!ld || !x || !q) return VIPRS_B200_EINVAL;
    return for_ld_dtype<T>(ld, [&](auto tag) { return launch_backward<T, decltype(tag)>(ld, x, q, dq, st); });
}

}  // namespace vb
