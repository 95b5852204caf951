// viprs_b200 -- C ABI entry points (include/viprs_b200.h): grid sweep, float64 state.
#include "grid_launch.cuh"

extern "C" int viprs_b200_e_step_grid_f64(const viprs_b200_ld_t* ld, int32_t G, int32_t n_active,
                                          const int32_t* active_model_idx, const double* std_beta, double* var_gamma,
                                          double* var_mu, double* eta, double* q, double* eta_diff, const double* u_logs,
                                          const double* half_var_tau, const double* mu_mult, double dq_scale,
                                          void* stream) {
    return vb::grid_dispatch<double>(ld, G, n_active, active_model_idx, std_beta, var_gamma, var_mu, eta, q, eta_diff,
                                     u_logs, half_var_tau, mu_mult, dq_scale, (cudaStream_t)stream);
}
