// viprs_b200 -- C ABI entry points (include/viprs_b200.h): sparse-mixture sweep, float64 state.
#include "launch.cuh"

extern "C" int viprs_b200_e_step_mixture_f64(const viprs_b200_ld_t* ld, int32_t K, const double* std_beta,
                                             double* var_gamma, double* var_mu, double* eta, double* q,
                                             double* eta_diff, const double* log_null_pi, const double* u_logs,
                                             const double* sqrt_half_var_tau, const double* mu_mult, double dq_scale,
                                             int32_t materialize_q, const double* q_offset, void* stream) {
    return vb::mixture_dispatch<double>(ld, K, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                                        sqrt_half_var_tau, mu_mult, dq_scale, materialize_q, q_offset, (cudaStream_t)stream);
}
