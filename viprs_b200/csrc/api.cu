// viprs_b200 -- C ABI entry points (include/viprs_b200.h) for the sweep kernels.
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ld.h"
#include "sweep.cuh"

namespace vb {

constexpr int kNBW = 8;   // bulk warps per CTA (CTA = 320 threads)

size_t sweep_smem_bytes(int max_block, int epv, int tsize, int stage_bytes, int nbw) {
    (void)epv;
    return make_layout(state_pad(max_block), tsize, stage_bytes, nbw).total;
}

static int max_optin_smem(int device) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    return v;
}

template <typename T, typename U>
static int launch_sweep(const viprs_b200_ld* ld, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                        T* eta_diff, const T* u_logs, const T* shvt, const T* mu_mult, T dq, cudaStream_t st) {
    SweepParams<T> p;
    p.packed = ld->d_packed; p.prow = ld->d_prow; p.pcs = ld->d_pcs; p.blk_row = ld->d_blk_row;
    p.blk_panel = ld->d_blk_panel; p.panel_row = ld->d_panel_row; p.blk_order = ld->d_blk_order;
    p.n_blocks = ld->n_blocks; p.stage_bytes = ld->stage_bytes; p.bpad = state_pad(ld->max_block);
    p.std_beta = std_beta; p.var_gamma = var_gamma; p.var_mu = var_mu; p.eta = eta; p.q = q;
    p.eta_diff = eta_diff; p.u_logs = u_logs; p.sqrt_half_var_tau = shvt; p.mu_mult = mu_mult; p.dq_scale = dq;
    const size_t smem = make_layout(p.bpad, (int)sizeof(T), p.stage_bytes, kNBW).total;
    if ((int64_t)smem > (int64_t)max_optin_smem(ld->device)) return VIPRS_B200_EBLOCK_TOO_LARGE;
    auto kern = sweep_kernel<T, U, kNBW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<ld->n_blocks, (kNBW + 2) * WARP, smem, st>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

template <typename T, typename U>
static int launch_backward(const viprs_b200_ld* ld, const T* x, T* q, T dq, cudaStream_t st) {
    const int wpb = 8;
    backward_dot_kernel<T, U><<<(ld->M + wpb - 1) / wpb, wpb * WARP, 0, st>>>(
        ld->M, (const U*)ld->d_packed, ld->d_prow, ld->d_pcs, x, q, dq);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

template <typename T>
static int e_step_dispatch(const viprs_b200_ld* ld, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q,
                           T* eta_diff, const T* u_logs, const T* shvt, const T* mu_mult, T dq,
                           int materialize_q, cudaStream_t st) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !shvt || !mu_mult)
        return VIPRS_B200_EINVAL;
    int rc;
    switch (ld->ld_dtype) {
        case VIPRS_B200_I8:
            rc = launch_sweep<T, int8_t>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, shvt, mu_mult, dq, st);
            if (rc == 0 && materialize_q) rc = launch_backward<T, int8_t>(ld, eta, q, dq, st);
            return rc;
        case VIPRS_B200_I16:
            rc = launch_sweep<T, int16_t>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, shvt, mu_mult, dq, st);
            if (rc == 0 && materialize_q) rc = launch_backward<T, int16_t>(ld, eta, q, dq, st);
            return rc;
        case VIPRS_B200_F32:
            rc = launch_sweep<T, float>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, shvt, mu_mult, dq, st);
            if (rc == 0 && materialize_q) rc = launch_backward<T, float>(ld, eta, q, dq, st);
            return rc;
        case VIPRS_B200_F64:
            if (sizeof(T) == 4) return VIPRS_B200_EUNSUPPORTED;
            if constexpr (sizeof(T) == 8) {
                rc = launch_sweep<T, double>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, shvt, mu_mult, dq, st);
                if (rc == 0 && materialize_q) rc = launch_backward<T, double>(ld, eta, q, dq, st);
                return rc;
            }
    }
    return VIPRS_B200_EUNSUPPORTED;
}

template <typename T>
static int backward_dispatch(const viprs_b200_ld* ld, const T* x, T* q, T dq, cudaStream_t st) {
    if (!ld || !x || !q) return VIPRS_B200_EINVAL;
    switch (ld->ld_dtype) {
        case VIPRS_B200_I8: return launch_backward<T, int8_t>(ld, x, q, dq, st);
        case VIPRS_B200_I16: return launch_backward<T, int16_t>(ld, x, q, dq, st);
        case VIPRS_B200_F32: return launch_backward<T, float>(ld, x, q, dq, st);
        case VIPRS_B200_F64:
            if constexpr (sizeof(T) == 8) return launch_backward<T, double>(ld, x, q, dq, st);
    }
    return VIPRS_B200_EUNSUPPORTED;
}

}  // namespace vb

extern "C" int viprs_b200_ld_info(const viprs_b200_ld_t* h, viprs_b200_ld_info_t* info) {
    if (!h || !info) return VIPRS_B200_EINVAL;
    info->M = h->M; info->ld_dtype = h->ld_dtype; info->n_blocks = h->n_blocks; info->max_block = h->max_block;
    info->n_panels = h->n_panels; info->stage_bytes = h->stage_bytes; info->nnz = h->nnz;
    info->packed_elems = h->packed_elems;
    info->smem_bytes = (int64_t)vb::sweep_smem_bytes(h->max_block, h->epv, 4, h->stage_bytes, vb::kNBW);
    return VIPRS_B200_OK;
}

extern "C" int viprs_b200_e_step_f32(const viprs_b200_ld_t* ld, const float* std_beta, float* var_gamma,
                                     float* var_mu, float* eta, float* q, float* eta_diff, const float* u_logs,
                                     const float* sqrt_half_var_tau, const float* mu_mult, float dq_scale,
                                     int32_t materialize_q, void* stream) {
    return vb::e_step_dispatch<float>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs,
                                      sqrt_half_var_tau, mu_mult, dq_scale, materialize_q, (cudaStream_t)stream);
}

extern "C" int viprs_b200_e_step_f64(const viprs_b200_ld_t* ld, const double* std_beta, double* var_gamma,
                                     double* var_mu, double* eta, double* q, double* eta_diff,
                                     const double* u_logs, const double* sqrt_half_var_tau,
                                     const double* mu_mult, double dq_scale, int32_t materialize_q, void* stream) {
    return vb::e_step_dispatch<double>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs,
                                       sqrt_half_var_tau, mu_mult, dq_scale, materialize_q, (cudaStream_t)stream);
}

extern "C" int viprs_b200_backward_dot_f32(const viprs_b200_ld_t* ld, const float* x, float* q, float dq_scale,
                                           void* stream) {
    return vb::backward_dispatch<float>(ld, x, q, dq_scale, (cudaStream_t)stream);
}
extern "C" int viprs_b200_backward_dot_f64(const viprs_b200_ld_t* ld, const double* x, double* q,
                                           double dq_scale, void* stream) {
    return vb::backward_dispatch<double>(ld, x, q, dq_scale, (cudaStream_t)stream);
}

// One-shot host-pointer drop-in for cpp_e_step (e_step_cpp.pyx:91-122).
extern "C" int viprs_b200_cpp_e_step(int32_t M, const int32_t* ld_left_bound, const void* ld_indptr,
                                     int32_t indptr_is_i64, const void* ld_data, int32_t ld_dtype,
                                     int32_t float_dtype, const void* std_beta, void* var_gamma, void* var_mu,
                                     void* eta, void* q, void* eta_diff, const void* u_logs,
                                     const void* sqrt_half_var_tau, const void* mu_mult, double dq_scale,
                                     int32_t threads, int32_t low_memory) {
    (void)threads; (void)low_memory;   // always the threads=1 order; the layout is detected per row
    if (float_dtype != VIPRS_B200_F32 && float_dtype != VIPRS_B200_F64) return VIPRS_B200_EINVAL;
    if (!std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !sqrt_half_var_tau || !mu_mult)
        return VIPRS_B200_EINVAL;
    viprs_b200_ld_t* ld = nullptr;
    int rc = viprs_b200_ld_create(&ld, M, ld_left_bound, ld_indptr, indptr_is_i64, ld_data, ld_dtype,
                                  VIPRS_B200_MEM_HOST, 0, nullptr);
    if (rc) return rc;
    const size_t ts = float_dtype == VIPRS_B200_F32 ? 4 : 8;
    const size_t nb = (size_t)M * ts;
    unsigned char* d = nullptr;   // 9 arrays: beta, gamma, mu, eta, q, diff, ulogs, shvt, mm
    cudaError_t e = cudaMalloc(&d, 9 * nb);
    if (e != cudaSuccess) { viprs_b200_ld_destroy(ld); return (int)e; }
    const void* hin[9] = {std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult};
    for (int i = 0; i < 9 && e == cudaSuccess; ++i)
        e = cudaMemcpyAsync(d + i * nb, hin[i], nb, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) {
        if (ts == 4)
            rc = viprs_b200_e_step_f32(ld, (float*)(d), (float*)(d + nb), (float*)(d + 2 * nb), (float*)(d + 3 * nb),
                                       (float*)(d + 4 * nb), (float*)(d + 5 * nb), (float*)(d + 6 * nb),
                                       (float*)(d + 7 * nb), (float*)(d + 8 * nb), (float)dq_scale, 1, nullptr);
        else
            rc = viprs_b200_e_step_f64(ld, (double*)(d), (double*)(d + nb), (double*)(d + 2 * nb), (double*)(d + 3 * nb),
                                       (double*)(d + 4 * nb), (double*)(d + 5 * nb), (double*)(d + 6 * nb),
                                       (double*)(d + 7 * nb), (double*)(d + 8 * nb), dq_scale, 1, nullptr);
    }
    void* hout[5] = {var_gamma, var_mu, eta, q, eta_diff};
    for (int i = 0; i < 5 && e == cudaSuccess && rc == 0; ++i)
        e = cudaMemcpyAsync(hout[i], d + (i + 1) * nb, nb, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    cudaFree(d);
    viprs_b200_ld_destroy(ld);
    if (rc) return rc;
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}
