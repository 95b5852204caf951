// viprs_b200 -- C ABI entry points (include/viprs_b200.h): spike-and-slab sweep, float64 state.
#include "launch.cuh"

extern "C" int viprs_b200_e_step_f64(const viprs_b200_ld_t* ld, const double* std_beta, double* var_gamma,
                                     double* var_mu, double* eta, double* q, double* eta_diff,
                                     const double* u_logs, const double* sqrt_half_var_tau,
                                     const double* mu_mult, double dq_scale, int32_t materialize_q, const double* q_offset, void* stream) {
    return vb::e_step_dispatch<double>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs,
                                       sqrt_half_var_tau, mu_mult, dq_scale, materialize_q, q_offset, (cudaStream_t)stream);
}

extern "C" int viprs_b200_backward_dot_f64(const viprs_b200_ld_t* ld, const double* x, double* q,
                                           double dq_scale, void* stream) {
    return vb::backward_dispatch<double>(ld, x, q, dq_scale, (cudaStream_t)stream);
}

extern "C" int viprs_b200_q_offset_f64(const viprs_b200_ld_t* ld, const double* eta, const double* q, double dq_scale,
                                      double* q_offset_out, void* stream) {
    return vb::q_offset_dispatch<double>(ld, eta, q, dq_scale, q_offset_out, (cudaStream_t)stream);
}
