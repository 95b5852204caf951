// viprs_b200 -- launch helpers shared by the per-type translation units.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ld.h"
#include "sweep.cuh"

namespace vb {

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int make_plan(const viprs_b200_ld* ld, int tsize, SweepPlan& p, RingGeometry& g) {
    g = ring_geometry(ld, tsize);
    if (g.nst == 0) return VIPRS_B200_EBLOCK_TOO_LARGE;
    p.packed = reinterpret_cast<const unsigned char*>(ld->d_packed);
    p.prow = ld->d_prow; p.pcs = ld->d_pcs; p.blk_row = ld->d_blk_row; p.blk_panel = ld->d_blk_panel;
    p.panel_row = ld->d_panel_row; p.blk_order = ld->d_blk_order; p.panel_need = ld->d_panel_need;
    p.n_blocks = ld->n_blocks; p.stage_bytes = ld->stage_bytes; p.nst = g.nst; p.bpad = state_pad(ld->max_block);
    p.l2_ahead = env_int("VIPRS_B200_L2_AHEAD", 8);
    p.trace = nullptr;
    p.L = make_layout(p.bpad, tsize, ld->stage_bytes, g.nst);
    return VIPRS_B200_OK;
}

template <typename T, typename U, typename Model, int MINB>
static int launch_one(const viprs_b200_ld* ld, const SweepPlan& p, const RingGeometry& g,
                      const typename Model::Args& ma, const StateArgs<T>& sa, cudaStream_t st) {
    auto kern = sweep_kernel<T, U, Model, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes);
    if (e != cudaSuccess) return (int)e;
    const char* trace_path = getenv("VIPRS_B200_TRACE");      // debug only: timeline of CTA 0 of every launch
    if (trace_path) {
        SweepPlan pt = p;
        const size_t nb = (size_t)kTraceSlots * sizeof(unsigned long long);
        if (cudaMalloc(&pt.trace, nb) != cudaSuccess) return VIPRS_B200_ENOMEM;
        cudaMemsetAsync(pt.trace, 0, nb, st);
        kern<<<ld->n_blocks, (NBW + 2) * WARP, g.smem_bytes, st>>>(pt, ma, sa);
        cudaStreamSynchronize(st);
        std::vector<unsigned long long> h(kTraceSlots);
        cudaMemcpy(h.data(), pt.trace, nb, cudaMemcpyDeviceToHost);
        cudaFree(pt.trace);
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), 1, nb, f); fclose(f); }
        e = cudaGetLastError();
        return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
    }
    kern<<<ld->n_blocks, (NBW + 2) * WARP, g.smem_bytes, st>>>(p, ma, sa);
    e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

template <typename T, typename U, typename Model>
static int launch_sweep(const viprs_b200_ld* ld, const typename Model::Args& ma, const StateArgs<T>& sa, cudaStream_t st) {
    SweepPlan p;
    RingGeometry g;
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
