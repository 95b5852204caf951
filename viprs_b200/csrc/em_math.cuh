// viprs_b200 -- per-element math shared by the EM streaming kernels (em.cu) and the reductions fused into the sweep
// (sweep.cuh, producer_out_role): theta layout, float32-aware logarithms / reciprocals, the clipping of VIPRS.py:509-518.
#pragma once
#include "common.cuh"

namespace vb {

struct Theta { double sigma_epsilon, tau_beta, pi, lambda_min; };

// float32 state: the per-element transcendental work of the two streaming kernels is what bounds them at G = 256
// (profiles/r01b_c3_launches.csv), so where the INPUT is an exact float32 number the logarithm is taken in float32
// (<= 1 ulp of the result, i.e. as accurate as the float32 input deserves; accumulation stays float64), and double
// reciprocals start from the float32 reciprocal plus one Newton step (relative error ~4e-15).  float64 state: IEEE.
template <typename T> __device__ __forceinline__ double rcp_em(double d) {
    if constexpr (sizeof(T) == 4) {
        const double r = (double)__frcp_rn((float)d);
        return r * (2.0 - d * r);
    } else {
        return 1.0 / d;
    }
}
// log(var_tau) for a stream of nearby values (var_tau = n_j * nscale + tau_beta along one column).  float32 state:
// one float64 logarithm per thread for its first value v0, then log(v) = log(v0) + log1p((v - v0) / v0) with the
// log1p in float32 -- absolute error ~1e-8 on a value of ~13, i.e. two orders below float32 rounding of u_logs and
// (used by the prepare kernel; in the sums kernel the extra state cost more than the logarithm it saved).
// float64 state: the IEEE logarithm every time.
template <typename T>
struct LogNear {
    double v0 = 0.0, l0 = 0.0, r0 = 0.0;
    bool have = false;
    __device__ __forceinline__ double operator()(double v) {
        if constexpr (sizeof(T) == 4) {
            if (!have) { v0 = v; l0 = log(v); r0 = 1.0 / v; have = true; return l0; }
            return l0 + (double)log1pf((float)((v - v0) * r0));
        } else {
            return log(v);
        }
    }
};
// log(var_tau), var_tau ~ 1e4 .. 1e7 in float64.  float32 state: the float32 logarithm of the rounded value -- 1e-7
// relative on a value of ~13, far inside what the float32 gamma next to it carries; float64 state: IEEE.
template <typename T> __device__ __forceinline__ double log_tau(double vt) {
    if constexpr (sizeof(T) == 4) return (double)logf((float)vt);
    else return log(vt);
}
// log(x) for x = clip(g), g an exact value of type T in [0, 1]
template <typename T> __device__ __forceinline__ double log_unit(double xc) {
    if constexpr (sizeof(T) == 4) return (double)logf((float)xc);
    else return log(xc);
}
// log(clip(1 - p)) for an exact p of type T in [0, 1]: 1 - p is ill-conditioned near 1, log1p is not
template <typename T> __device__ __forceinline__ double log_one_minus(double p, double ngc) {
    if constexpr (sizeof(T) == 4) {
        const double res = 1e-15;
        if (ngc <= res || ngc >= 1.0 - res) return log(ngc);                  // clipped: rare, exact
        return p <= 0.5 ? (double)log1pf(-(float)p) : (double)logf((float)(1.0 - p));   // 1 - p is exact for p >= 0.5
    } else {
        return log(ngc);
    }
}

__device__ __forceinline__ double clip_res(double g) {                              // VIPRS.py:509-518
    const double res = 1e-15;                                                       // np.finfo(np.float64).resolution
    return fmin(fmax(g, res), 1.0 - res);
}

constexpr int NS = VIPRS_B200_NSUMS;

}  // namespace vb
