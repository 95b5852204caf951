// viprs_b200 -- C ABI entry points (include/viprs_b200.h): grid sweep, float32 state.
#include "grid_launch.cuh"

extern "C" int viprs_b200_e_step_grid_f32(const viprs_b200_ld_t* ld, int32_t G, int32_t n_active,
                                          const int32_t* active_model_idx, const float* std_beta, float* var_gamma,
                                          float* var_mu, float* eta, float* q, float* eta_diff, const float* u_logs,
                                          const float* half_var_tau, const float* mu_mult, float dq_scale, void* stream) {
    return vb::grid_dispatch<float>(ld, G, n_active, active_model_idx, std_beta, var_gamma, var_mu, eta, q, eta_diff,
                                    u_logs, half_var_tau, mu_mult, dq_scale, (cudaStream_t)stream);
}
