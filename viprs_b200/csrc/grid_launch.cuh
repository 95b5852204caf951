// viprs_b200 -- launch helper of the grid sweep (shared by grid_f32.cu / grid_f64.cu).
#pragma once
#include "grid.cuh"
#include "ld.h"

namespace vb {

template <typename T, typename U>
static int launch_grid_one(const viprs_b200_ld* ld, int n_active, const int32_t* active, const GridArgs<T>& ga,
                           cudaStream_t st) {
    constexpr int GT = GridT<T>::GT;
    constexpr int ES = (int)sizeof(U);
    GridPlan p;
    const int bp = (ld->max_block + 15) & ~15;
    int nbw = 1;
    while (nbw * WARP * GKPT < bp) nbw *= 2;                    // 1, 2, 4, 8 bulk warps
    const int row_bytes = bp * ES;
    p.stage_bytes = GP * (row_bytes < GCW ? row_bytes : GCW);
    const int fixed = (int)make_grid_layout((int)sizeof(T), GT, 0, 0).total + 256;
    int nst = (ld->smem_optin - fixed) / p.stage_bytes;
    if (nst > GNST_MAX) nst = GNST_MAX;
    if (nst < 2) return VIPRS_B200_EBLOCK_TOO_LARGE;
    p.nst = nst; p.nbw = nbw;
    p.L = make_grid_layout((int)sizeof(T), GT, p.stage_bytes, nst);
    p.dense = reinterpret_cast<const unsigned char*>(ld->d_dense);
    p.dblk_off = ld->d_dblk_off; p.blk_row = ld->d_blk_row; p.blk_order = ld->d_blk_order;
    p.active = active; p.n_active = n_active; p.n_tiles = (n_active + GT - 1) / GT; p.M = ld->M;
    auto kern = grid_sweep_kernel<T, U>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.L.total);
    if (e != cudaSuccess) return (int)e;
    kern<<<ld->n_blocks * p.n_tiles, grid_threads(nbw), p.L.total, st>>>(p, ga);
    e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

template <typename T>
static int grid_dispatch(const viprs_b200_ld* ld, int G, int n_active, const int32_t* active, const T* std_beta,
                         T* var_gamma, T* var_mu, T* eta, T* q, T* eta_diff, const T* u_logs, const T* half_var_tau,
                         const T* mu_mult, T dq, cudaStream_t st) {
    if (!ld || !std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !half_var_tau || !mu_mult)
        return VIPRS_B200_EINVAL;
    if (G < 1 || n_active < 0 || n_active > G || (n_active > 0 && !active)) return VIPRS_B200_EINVAL;
    if (n_active == 0) return VIPRS_B200_OK;
    if (ld->max_block > GRID_MAX_BLOCK) return VIPRS_B200_EBLOCK_TOO_LARGE;
    int rc = ensure_dense(ld, st);
    if (rc) return rc;
    GridArgs<T> ga{std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, half_var_tau, mu_mult, dq};
    switch (ld->ld_dtype) {
        case VIPRS_B200_I8: return launch_grid_one<T, int8_t>(ld, n_active, active, ga, st);
        case VIPRS_B200_I16: return launch_grid_one<T, int16_t>(ld, n_active, active, ga, st);
        case VIPRS_B200_F32: return launch_grid_one<T, float>(ld, n_active, active, ga, st);
        case VIPRS_B200_F64:
            if constexpr (sizeof(T) == 8) return launch_grid_one<T, double>(ld, n_active, active, ga, st);
            return VIPRS_B200_EUNSUPPORTED;
    }
    return VIPRS_B200_EUNSUPPORTED;
}

}  // namespace vb
