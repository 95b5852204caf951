// viprs_b200 -- C ABI entry points (include/viprs_b200.h): spike-and-slab sweep, float32 state.
#include "launch.cuh"

extern "C" int viprs_b200_e_step_f32(const viprs_b200_ld_t* ld, const float* std_beta, float* var_gamma,
                                     float* var_mu, float* eta, float* q, float* eta_diff, const float* u_logs,
                                     const float* sqrt_half_var_tau, const float* mu_mult, float dq_scale,
                                     int32_t materialize_q, const float* q_offset, void* stream) {
    return vb::e_step_dispatch<float>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs,
                                      sqrt_half_var_tau, mu_mult, dq_scale, materialize_q, q_offset, (cudaStream_t)stream);
}

extern "C" int viprs_b200_backward_dot_f32(const viprs_b200_ld_t* ld, const float* x, float* q, float dq_scale,
                                           void* stream) {
    return vb::backward_dispatch<float>(ld, x, q, dq_scale, (cudaStream_t)stream);
}

extern "C" int viprs_b200_q_offset_f32(const viprs_b200_ld_t* ld, const float* eta, const float* q, float dq_scale,
                                      float* q_offset_out, void* stream) {
    return vb::q_offset_dispatch<float>(ld, eta, q, dq_scale, q_offset_out, (cudaStream_t)stream);
}

extern "C" int viprs_b200_e_step_fused_f32(const viprs_b200_ld_t* ld, const float* std_beta, float* var_gamma, float* var_mu,
                                           float* eta, float* q, float* eta_diff, const float* u_logs,
                                           const float* sqrt_half_var_tau, const float* mu_mult, float dq_scale,
                                           const double* n_per_snp, const double* theta, int32_t nseg,
                                           const int32_t* seg_ptr, double* sums, void* stream) {
    return vb::e_step_fused_dispatch<float>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau,
                                            mu_mult, dq_scale, n_per_snp, theta, nseg, seg_ptr, sums, (cudaStream_t)stream);
}

int vb::incr_slab_f32(const viprs_b200_ld* ld, const float* std_beta, float* var_gamma, float* var_mu, float* eta, float* q,
                      float* eta_diff, const float* u_logs, const float* shvt, const float* mu_mult, float dq, int chunk,
                      cudaStream_t st, cudaEvent_t swept) {
    return vb::e_step_incremental_dispatch<float>(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, shvt, mu_mult, dq,
                                                  chunk, st, swept);
}

extern "C" int viprs_b200_e_step_incremental_f32(const viprs_b200_ld_t* ld, const float* std_beta, float* var_gamma,
                                                 float* var_mu, float* eta, float* q, float* eta_diff, const float* u_logs,
                                                 const float* sqrt_half_var_tau, const float* mu_mult, float dq_scale,
                                                 void* stream) {
    return vb::incr_slab_f32(ld, std_beta, var_gamma, var_mu, eta, q, eta_diff, u_logs, sqrt_half_var_tau, mu_mult, dq_scale, -1,
                             (cudaStream_t)stream);
}
