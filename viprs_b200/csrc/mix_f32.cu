// viprs_b200 -- C ABI entry points (include/viprs_b200.h): sparse-mixture sweep, float32 state.
#include "launch.cuh"

extern "C" int viprs_b200_e_step_mixture_f32(const viprs_b200_ld_t* ld, int32_t K, const float* std_beta,
                                             float* var_gamma, float* var_mu, float* eta, float* q, float* eta_diff,
                                             const float* log_null_pi, const float* u_logs,
                                             const float* sqrt_half_var_tau, const float* mu_mult, float dq_scale,
                                             int32_t materialize_q, const float* q_offset, void* stream) {
    return vb::mixture_dispatch<float>(ld, K, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                                       sqrt_half_var_tau, mu_mult, dq_scale, materialize_q, q_offset, (cudaStream_t)stream);
}

int vb::incr_mix_f32(const viprs_b200_ld* ld, int K, const float* std_beta, float* var_gamma, float* var_mu, float* eta, float* q,
                     float* eta_diff, const float* log_null_pi, const float* u_logs, const float* shvt, const float* mu_mult,
                     float dq, int chunk, cudaStream_t st, cudaEvent_t swept) {
    return vb::mixture_incremental_dispatch<float>(ld, K, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                                                   shvt, mu_mult, dq, chunk, st, swept);
}

extern "C" int viprs_b200_e_step_mixture_incremental_f32(const viprs_b200_ld_t* ld, int32_t K, const float* std_beta,
                                                         float* var_gamma, float* var_mu, float* eta, float* q,
                                                         float* eta_diff, const float* log_null_pi, const float* u_logs,
                                                         const float* sqrt_half_var_tau, const float* mu_mult,
                                                         float dq_scale, void* stream) {
    return vb::incr_mix_f32(ld, K, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs, sqrt_half_var_tau, mu_mult,
                            dq_scale, -1, (cudaStream_t)stream);
}
