/*
 * TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C restatement of the reference coordinate-ascent E-step, single thread
 * (the reference's `threads=1` order, which is the parity target).
 * Every function cites the lines of /root/reference/viprs/model/vi/e_step.hpp it follows.
 * The arithmetic (fma placement, operation order, two-branch sigmoid, max-shifted softmax,
 * dequantisation folded into the axpy scalar) is kept identical so that this port is
 * bit-comparable with oracle/_ref/libviprs_ref.so (tests/test_oracle.py pins it to the
 * compiled reference and to the committed golden vectors in tests/golden/).
 *
 * Generated per (T, U) pair by the PORT_DEFINE macro at the bottom:
 *   T in {float, double}; U in {int8_t, int16_t, float, double}; indptr is int64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <float.h>

#define FMA_float(a, b, c) fmaf((a), (b), (c))
#define FMA_double(a, b, c) fma((a), (b), (c))
#define EXP_float(x) expf(x)
#define EXP_double(x) exp(x)
#define ABS_float(x) fabsf(x)
#define ABS_double(x) fabs(x)
#define EPS_float FLT_EPSILON
#define EPS_double DBL_EPSILON

#define PORT_DEFINE(TN, T, UN, U)                                                              \
/* e_step.hpp:157-175  axpy: x[i] = fma(T(y[i]), alpha, x[i]) */                               \
static void axpy_##TN##_##UN(T* x, const U* y, T alpha, int64_t size) {                        \
    for (int64_t i = 0; i < size; ++i) x[i] = FMA_##T((T)y[i], alpha, x[i]);                   \
}                                                                                              \
/* e_step.hpp:82-104  dot: s = fma(T(y[i]), x[i], s), ascending i */                           \
static T dot_##TN##_##UN(const T* x, const U* y, int64_t size) {                               \
    T s = 0;                                                                                   \
    for (int64_t i = 0; i < size; ++i) s = FMA_##T((T)y[i], x[i], s);                          \
    return s;                                                                                  \
}                                                                                              \
/* e_step.hpp:307-338  update_q_factor (second pass for upper-triangular LD) */                \
static void update_q_##TN##_##UN(int c_size, const int32_t* lb, const int64_t* indptr,         \
        const U* ld, const T* eta_diff, T* q, T dq_scale) {                                    \
    for (int j = 0; j < c_size; ++j) {                                                         \
        int64_t s = indptr[j], e = indptr[j + 1];                                              \
        q[j] += dq_scale * dot_##TN##_##UN(eta_diff + lb[j], ld + s, e - s);                   \
    }                                                                                          \
}                                                                                              \
/* e_step.hpp:245-261  two-branch stable sigmoid (constants are double in the reference,       \
   so the division is carried out in double and rounded to T on return) */                     \
static T sigmoid_##TN##_##UN(T x) {                                                            \
    if (x < 0) { T ex = EXP_##T(x); return (T)(ex / (1. + ex)); }                              \
    return (T)(1. / (1. + EXP_##T(-x)));                                                       \
}                                                                                              \
/* e_step.hpp:343-442  e_step (spike-and-slab) */                                              \
void port_e_step_##TN##_##UN(int c_size, const int32_t* lb, const int64_t* indptr,             \
        const U* ld, const T* std_beta, T* var_gamma, T* var_mu, T* eta, T* q, T* eta_diff,    \
        const T* u_logs, const T* sqrt_half_var_tau, const T* mu_mult, T dq_scale,             \
        int low_memory) {                                                                      \
    T eps = EPS_##T > (T)1e-8 ? EPS_##T : (T)1e-8;               /* :382 */                    \
    for (int j = 0; j < c_size; ++j) {                                                         \
        int64_t s = indptr[j], e = indptr[j + 1];                                              \
        int start = lb[j];                                                                     \
        T mu = FMA_##T(mu_mult[j], std_beta[j], -mu_mult[j] * q[j]);          /* :401 */       \
        T u = sqrt_half_var_tau[j] * mu;                                      /* :404 */       \
        T g = sigmoid_##TN##_##UN(FMA_##T(u, u, u_logs[j]));                  /* :405 */       \
        T d = FMA_##T(g, mu, -eta[j]);                                        /* :408 */       \
        if (ABS_##T(d) < eps) { eta_diff[j] = 0; continue; }                  /* :410-413 */   \
        var_mu[j] = mu; var_gamma[j] = g; eta_diff[j] = d;                    /* :416-418 */   \
        axpy_##TN##_##UN(q + start, ld + s, dq_scale * d, e - s);             /* :421 */       \
        if (!low_memory) q[j] -= d;                                           /* :427 */       \
        eta[j] += d;                                                          /* :431 */       \
    }                                                                                          \
    if (low_memory) update_q_##TN##_##UN(c_size, lb, indptr, ld, eta_diff, q, dq_scale);       \
}                                                                                              \
/* e_step.hpp:222-241 softmax (writes size-1 outputs, overwrites logits) +                     \
   e_step.hpp:447-551 e_step_mixture; (M,K) arrays are C-order */                              \
void port_e_step_mixture_##TN##_##UN(int c_size, int K, const int32_t* lb,                     \
        const int64_t* indptr, const U* ld, const T* std_beta, T* var_gamma, T* var_mu,        \
        T* eta, T* q, T* eta_diff, const T* log_null_pi, const T* u_logs,                      \
        const T* sqrt_half_var_tau, const T* mu_mult, T dq_scale, int low_memory) {            \
    T* u = (T*)malloc(sizeof(T) * (size_t)(K + 1));                                            \
    for (int j = 0; j < c_size; ++j) {                                                         \
        int64_t s = indptr[j], e = indptr[j + 1];                                              \
        int start = lb[j];                                                                     \
        T r = std_beta[j] - q[j];                                             /* :505 */       \
        for (int k = 0; k < K; ++k) {                                                          \
            int64_t m = (int64_t)j * K + k;                                                    \
            var_mu[m] = mu_mult[m] * r;                                       /* :509 */       \
            T t = sqrt_half_var_tau[m] * var_mu[m];                           /* :510 */       \
            u[k] = FMA_##T(t, t, u_logs[m]);                                  /* :511 */       \
        }                                                                                      \
        u[K] = log_null_pi[j];                                                /* :515 */       \
        T mx = u[0];                                                          /* :58-71 */     \
        for (int k = 1; k <= K; ++k) if (mx < u[k]) mx = u[k];                                 \
        T sum = 0;                                                                             \
        for (int k = 0; k <= K; ++k) { u[k] = EXP_##T(u[k] - mx); sum += u[k]; } /* :233-236 */\
        for (int k = 0; k < K; ++k) var_gamma[(int64_t)j * K + k] = u[k] / sum;  /* :238-240 */\
        eta_diff[j] = -eta[j];                                                /* :519 */       \
        for (int k = 0; k < K; ++k) {                                                          \
            int64_t m = (int64_t)j * K + k;                                                    \
            eta_diff[j] = FMA_##T(var_gamma[m], var_mu[m], eta_diff[j]);      /* :523 */       \
        }                                                                                      \
        axpy_##TN##_##UN(q + start, ld + s, dq_scale * eta_diff[j], e - s);   /* :527 */       \
        if (!low_memory) q[j] -= eta_diff[j];                                 /* :533 */       \
        eta[j] += eta_diff[j];                                                /* :536 */       \
    }                                                                                          \
    free(u);                                                                                   \
    if (low_memory) update_q_##TN##_##UN(c_size, lb, indptr, ld, eta_diff, q, dq_scale);       \
}                                                                                              \
/* e_step.hpp:555-647 e_step_grid + 266-303 update_q_factor_matrix; (M,G) arrays are           \
   column-major, index model*c_size + j; SNP-outer / model-inner */                            \
void port_e_step_grid_##TN##_##UN(int c_size, int n_active, const int32_t* active,             \
        const int32_t* lb, const int64_t* indptr, const U* ld, const T* std_beta,              \
        T* var_gamma, T* var_mu, T* eta, T* q, T* eta_diff, const T* u_logs,                   \
        const T* half_var_tau, const T* mu_mult, T dq_scale, int low_memory) {                 \
    for (int j = 0; j < c_size; ++j) {                                                         \
        int64_t s = indptr[j], e = indptr[j + 1];                                              \
        int start = lb[j];                                                                     \
        for (int a = 0; a < n_active; ++a) {                                                   \
            int64_t col = (int64_t)active[a] * c_size;                                         \
            int64_t m = col + j;                                                               \
            var_mu[m] = mu_mult[m] * (std_beta[j] - q[m]);                    /* :613 */       \
            T uj = u_logs[m] + half_var_tau[m] * var_mu[m] * var_mu[m];       /* :616 */       \
            var_gamma[m] = sigmoid_##TN##_##UN(uj);                           /* :617 */       \
            eta_diff[m] = var_gamma[m] * var_mu[m] - eta[m];                  /* :620 */       \
            axpy_##TN##_##UN(q + col + start, ld + s, dq_scale * eta_diff[m], e - s); /* :623*/\
            if (!low_memory) q[m] -= eta_diff[m];                             /* :629 */       \
            eta[m] += eta_diff[m];                                            /* :633 */       \
        }                                                                                      \
    }                                                                                          \
    if (low_memory) {                                                         /* :291-302 */   \
        for (int j = 0; j < c_size; ++j) {                                                     \
            int64_t s = indptr[j], e = indptr[j + 1];                                          \
            for (int a = 0; a < n_active; ++a) {                                               \
                int64_t col = (int64_t)active[a] * c_size;                                     \
                q[col + j] += dq_scale *                                                       \
                    dot_##TN##_##UN(eta_diff + col + lb[j], ld + s, e - s);                    \
            }                                                                                  \
        }                                                                                      \
    }                                                                                          \
}

PORT_DEFINE(f32, float, i8, int8_t)
PORT_DEFINE(f32, float, i16, int16_t)
PORT_DEFINE(f32, float, f32, float)
PORT_DEFINE(f64, double, i8, int8_t)
PORT_DEFINE(f64, double, i16, int16_t)
PORT_DEFINE(f64, double, f32, float)
PORT_DEFINE(f64, double, f64, double)
