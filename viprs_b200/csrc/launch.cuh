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
        kern<<<ld->n_blocks, (NBW + 2) * WARP, g.smem_bytes, st>>>(pp, ma, sa);
    });
}

// register-resident kernel (float32 state, LD blocks <= 4096 SNPs, two CTAs per SM)
template <typename U, typename Model, int NLIMB>
static int launch_fast_one(const viprs_b200_ld* ld, SweepPlan p, const typename Model::Args& ma,
                           const StateArgs<float>& sa, cudaStream_t st) {
    const RingGeometry g = fast_ring_geometry(ld);
    const FastLayout FL = make_fast_layout(ld->stage_bytes, g.nst);
    p.nst = g.nst;
    p.L.stages = FL.stages;
    auto kern = sweep_fast_kernel<U, Model, NLIMB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL.total);
    if (e != cudaSuccess) return (int)e;
    return launch_traced(p, st, [&](const SweepPlan& pp) {
        kern<<<ld->n_blocks, FAST_WARPS * WARP, FL.total, st>>>(pp, FL, ma, sa);
    });
}

template <typename T, typename U, typename Model>
static bool fast_path_ok(const viprs_b200_ld* ld) {
    if constexpr (sizeof(T) != 4 || Model::kHeavy || sizeof(U) == 8) return false;
    return ld->max_block <= FAST_MAX_BLOCK && fast_ring_geometry(ld).nst >= 3 && env_int("VIPRS_B200_FORCE_GENERIC", 0) == 0;
}

template <typename T, typename U, typename Model>
static int launch_sweep(const viprs_b200_ld* ld, const typename Model::Args& ma, const StateArgs<T>& sa, cudaStream_t st) {
    SweepPlan p;
    RingGeometry g;
    if constexpr (sizeof(T) == 4 && !Model::kHeavy && sizeof(U) != 8) {
        if (fast_path_ok<T, U, Model>(ld)) {
            make_plan(ld, (int)sizeof(T), p, g);        // ring fields are overridden by the fast launcher
            if (std::is_same<U, int8_t>::value && env_int("VIPRS_B200_LIMBS", 3) == 3)
                return launch_fast_one<U, Model, 3>(ld, p, ma, sa, st);
            return launch_fast_one<U, Model, 4>(ld, p, ma, sa, st);
        }
    }
    int rc = make_plan(ld, (int)sizeof(T), p, g);
    if (rc) return rc;
    if constexpr (!Model::kHeavy) {
        if (g.ctas_per_sm >= 2) return launch_one<T, U, Model, 2>(ld, p, g, ma, sa, st);
    }
    return launch_one<T, U, Model, 1>(ld, p, g, ma, sa, st);
}

template <typename T, typename U>
static int launch_backward(const viprs_b200_ld* ld, const T* x, T* q, T dq, cudaStream_t st) {
    const int wpb = 8;
    backward_dot_kernel<T, U><<<(ld->M + wpb - 1) / wpb, wpb * WARP, 0, st>>>(
        ld->M, reinterpret_cast<const unsigned char*>(ld->d_packed), ld->d_prow, ld->d_pcs, x, q, dq);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
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
                           int materialize_q, cudaStream_t st) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !shvt || !mu_mult)
        return VIPRS_B200_EINVAL;
    typename SlabModel<T>::Args ma{std_beta, u_logs, shvt, mu_mult, var_gamma, var_mu, dq};
    StateArgs<T> sa{eta, q, eta_diff};
    return for_ld_dtype<T>(ld, [&](auto tag) {
        using U = decltype(tag);
        int rc = launch_sweep<T, U, SlabModel<T>>(ld, ma, sa, st);
        if (rc == 0 && materialize_q) rc = launch_backward<T, U>(ld, eta, q, dq, st);
        return rc;
    });
}

template <typename T>
static int mixture_dispatch(const viprs_b200_ld* ld, int K, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                            T* eta_diff, const T* log_null_pi, const T* u_logs, const T* shvt, const T* mu_mult, T dq,
                            int materialize_q, cudaStream_t st) {
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
            rc = launch_sweep<T, U, MixModel<T, 4>>(ld, ma, sa, st);
        } else if (K <= 8) {
            typename MixModel<T, 8>::Args ma{std_beta, u_logs, shvt, mu_mult, log_null_pi, var_gamma, var_mu, dq, K};
            rc = launch_sweep<T, U, MixModel<T, 8>>(ld, ma, sa, st);
        } else {
            typename MixModel<T, 16>::Args ma{std_beta, u_logs, shvt, mu_mult, log_null_pi, var_gamma, var_mu, dq, K};
            rc = launch_sweep<T, U, MixModel<T, 16>>(ld, ma, sa, st);
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
