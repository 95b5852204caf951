// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" instantiations of the *unmodified* reference header
// /root/reference/viprs/model/vi/e_step.hpp (e_step: 343-442, e_step_mixture:
// 447-551, e_step_grid: 555-647).  The header is included from where it lies
// (-I/root/reference/viprs/model/vi); nothing of it is copied into this repo.
// Built by oracle/Makefile into oracle/_ref/libviprs_ref.so with the
// reference's own flags (setup.py:211 "-O3 -std=c++17" + OpenMP, no BLAS).
#include <cstdint>
#include "e_step.hpp"

#define INST(TN, T, UN, U)                                                                 \
extern "C" void ref_e_step_##TN##_##UN(int c_size, int* lb, int64_t* indptr, U* ld,        \
        T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q, T* eta_diff, T* u_logs,        \
        T* sqrt_half_var_tau, T* mu_mult, T dq_scale, int threads, int low_memory) {       \
    e_step<T, U, int64_t>(c_size, lb, indptr, ld, std_beta, var_gamma, var_mu, eta, q,     \
        eta_diff, u_logs, sqrt_half_var_tau, mu_mult, dq_scale, threads, low_memory != 0); \
}                                                                                          \
extern "C" void ref_e_step_mixture_##TN##_##UN(int c_size, int K, int* lb, int64_t* indptr,\
        U* ld, T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q, T* eta_diff,            \
        T* log_null_pi, T* u_logs, T* sqrt_half_var_tau, T* mu_mult, T dq_scale,           \
        int threads, int low_memory) {                                                     \
    e_step_mixture<T, U, int64_t>(c_size, K, lb, indptr, ld, std_beta, var_gamma, var_mu,  \
        eta, q, eta_diff, log_null_pi, u_logs, sqrt_half_var_tau, mu_mult, dq_scale,       \
        threads, low_memory != 0);                                                         \
}                                                                                          \
extern "C" void ref_e_step_grid_##TN##_##UN(int c_size, int n_active, int* active,         \
        int* lb, int64_t* indptr, U* ld, T* std_beta, T* var_gamma, T* var_mu, T* eta,     \
        T* q, T* eta_diff, T* u_logs, T* half_var_tau, T* mu_mult, T dq_scale,             \
        int threads, int low_memory) {                                                     \
    e_step_grid<T, U, int64_t>(c_size, n_active, active, lb, indptr, ld, std_beta,         \
        var_gamma, var_mu, eta, q, eta_diff, u_logs, half_var_tau, mu_mult, dq_scale,      \
        threads, low_memory != 0);                                                         \
}

INST(f32, float, i8, int8_t)
INST(f32, float, i16, int16_t)
INST(f32, float, f32, float)
INST(f64, double, i8, int8_t)
INST(f64, double, i16, int16_t)
INST(f64, double, f32, float)
INST(f64, double, f64, double)

extern "C" int ref_omp_supported() { return omp_supported() ? 1 : 0; }
extern "C" int ref_blas_supported() { return blas_supported() ? 1 : 0; }
