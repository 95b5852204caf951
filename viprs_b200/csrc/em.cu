// viprs_b200 -- the per-iteration work around the sweep, as two streaming kernels (HBM-bound, a few tens of bytes per
// SNP x model):
//   prepare : what VIPRS.e_step() / VIPRSMix.e_step() compute in numpy before calling cpp_e_step*
//             (/root/reference/viprs/model/VIPRS.py:400-406,418 ; VIPRSMix.py:187-204): var_tau, mu_mult, u_logs,
//             sqrt(var_tau/2) (or var_tau/2 for the grid kernel, e_step.hpp:616), log_null_pi -- float64 arithmetic,
//             cast to the state type like the reference.
//   sums    : the reductions m_step() / elbo() / mse() / the convergence test need (VIPRS.py:426-484, 497-581,
//             689-704, 997 ; VIPRSMix.py:227-260) -- float64 accumulation, fixed summation order (deterministic),
//             per chromosome segment and per model column / mixture component.
#include <cstdint>

#include "common.cuh"
#include "em_math.cuh"

namespace vb {

constexpr int EM_THREADS = 256;

// layout 0: (M, ncol) column-major (single model: ncol = 1; grid), 1: (M, ncol) row-major (mixture, ncol = K)
template <typename T>
__global__ void __launch_bounds__(EM_THREADS) prepare_kernel(int M, int ncol, int layout, int half_tau,
                                                             const double* __restrict__ n_per_snp,
                                                             const Theta* __restrict__ theta, T* __restrict__ u_logs,
                                                             T* __restrict__ tau_term, T* __restrict__ mu_mult,
                                                             T* __restrict__ log_null_pi) {
    const int c = blockIdx.y;
    const Theta th = theta[c];
    const double cst = log(th.pi) - log(1.0 - th.pi) + 0.5 * log(th.tau_beta);      // VIPRS.py:405
    const double nscale = (1.0 + th.lambda_min) / th.sigma_epsilon;                 // VIPRS.py:400
    double lnp = 0.0;
    if (log_null_pi != nullptr && c == 0) {                                         // VIPRSMix.py:191-194
        double s = 0.0;
        for (int k = 0; k < ncol; ++k) s += theta[k].pi;
        lnp = log(1.0 - s);
    }
    LogNear<T> lognear;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const double n = n_per_snp[j];
        const double vt = n * nscale + th.tau_beta;
        const size_t e = layout == 0 ? (size_t)c * M + j : (size_t)j * ncol + c;
        mu_mult[e] = (T)(n * rcp_em<T>(vt * th.sigma_epsilon));                              // VIPRS.py:404
        u_logs[e] = (T)(cst - 0.5 * lognear(vt));                                   // VIPRS.py:405-406
        tau_term[e] = (T)(half_tau ? 0.5 * vt : sqrt(0.5 * vt));                    // e_step.hpp:616 / VIPRS.py:418
        if (log_null_pi != nullptr && c == 0) log_null_pi[j] = (T)lnp;
    }
}

// grid = (chunks, nseg, ncol).  partial[((seg * ncol + c) * chunks + chunk) * NS + slot]; the last CTA of every
// (seg, c) adds the chunk partials in chunk order and writes sums[(seg * ncol + c) * NS + slot].
template <typename T>
__global__ void __launch_bounds__(EM_THREADS) sums_kernel(int M, int ncol, int layout, const int32_t* __restrict__ seg_ptr,
                                                          const T* __restrict__ var_gamma, const T* __restrict__ var_mu,
                                                          const T* __restrict__ eta, const T* __restrict__ q,
                                                          const T* __restrict__ eta_diff, const T* __restrict__ std_beta,
                                                          const double* __restrict__ n_per_snp,
                                                          const Theta* __restrict__ theta,
                                                          const Theta* __restrict__ theta_logtau, double q_scale,
                                                          double* __restrict__ partial, unsigned int* __restrict__ counters,
                                                          double* __restrict__ sums) {
    const int chunk = blockIdx.x, chunks = gridDim.x, seg = blockIdx.y, c = blockIdx.z;
    const int row0 = seg_ptr[seg], row1 = seg_ptr[seg + 1];
    const Theta th = theta[c], tl = theta_logtau[c];
    const double nscale = (1.0 + th.lambda_min) / th.sigma_epsilon;
    const double nscale_l = (1.0 + tl.lambda_min) / tl.sigma_epsilon;
    const bool per_snp = (layout == 0) || (c == 0);       // mixture: eta / q / pip terms are per SNP, kept in column 0
    const bool same_tau = (theta_logtau == theta);
    double acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = 0.0;
    // the operands of the next row are requested before the arithmetic of the current one (the loop is otherwise a chain of
    // load -> ~280 instructions -> load: 9 warps per issue slot waiting on memory at G = 256)
    struct Row { T g, mu, et, qv, df, sb; double n; };
    auto fetch = [&](int j, Row& r) {
        const size_t e = layout == 0 ? (size_t)c * M + j : (size_t)j * ncol + c;
        r.g = var_gamma[e]; r.mu = var_mu[e]; r.n = n_per_snp[j];
        if (per_snp) {
            const size_t ev = layout == 0 ? e : (size_t)j;
            r.et = eta[ev]; r.qv = q[ev]; r.df = eta_diff[ev]; r.sb = std_beta[j];
        }
    };
    const int jstep = chunks * blockDim.x;
    int j = row0 + chunk * blockDim.x + threadIdx.x;
    Row cur{}, nxt{};
    if (j < row1) fetch(j, cur);
    for (; j < row1; j += jstep, cur = nxt) {
        if (j + jstep < row1) fetch(j + jstep, nxt);
        const double g = (double)cur.g;
        const double mu = (double)cur.mu;
        const double n = cur.n;
        const double vt = n * nscale + th.tau_beta;
        const double gc = clip_res(g);
        acc[VIPRS_B200_S_GAMMA] += g;                                               // VIPRS.py:434
        acc[VIPRS_B200_S_GAMMA_MU2] += g * mu * mu;                                 // zeta, VIPRS.py:896
        const double ivt = rcp_em<T>(vt);
        acc[VIPRS_B200_S_G_INV_TAU] += g * ivt;
        acc[VIPRS_B200_S_G_LOGG] += gc * log_unit<T>(gc);                                   // VIPRS.py:562
        acc[VIPRS_B200_S_GCLIP] += gc;
        acc[VIPRS_B200_S_G_LOG_TAU] += gc * log_tau<T>(same_tau ? vt : n * nscale_l + tl.tau_beta);   // VIPRS.py:565 (log_var_tau cache)
        acc[VIPRS_B200_S_GC_ZETA] += gc * (mu * mu + ivt);                          // VIPRS.py:571-573
        if (per_snp) {
            const double et = (double)cur.et;
            double pip;
            if (layout == 0) {
                pip = g;
            } else {                                                               // VIPRSMix.py:297-301 (sum in T)
                T s = T(0);
                for (int k = 0; k < ncol; ++k) s += var_gamma[(size_t)j * ncol + k];
                pip = (double)s;
            }
            const double ng = clip_res(1.0 - pip);
            acc[VIPRS_B200_S_ETA_Q] += q_scale * et * (double)cur.qv;               // VIPRS.py:455
            acc[VIPRS_B200_S_BETA_ETA] += (double)cur.sb * et;                      // VIPRS.py:469
            acc[VIPRS_B200_S_NG_LOGNG] += ng * log_one_minus<T>(pip, ng);                             // VIPRS.py:563
            acc[VIPRS_B200_S_NGCLIP] += ng;
            acc[VIPRS_B200_S_ETA2] += et * et;                                      // VIPRS.py:703
            acc[VIPRS_B200_S_MAX_DIFF] = fmax(acc[VIPRS_B200_S_MAX_DIFF], fabs((double)cur.df));   // VIPRS.py:997
        }
    }
    // block reduction in a fixed order: lanes by xor-shuffle, then warps 0..7 sequentially
    __shared__ double red[EM_THREADS / WARP][NS];
    __shared__ bool is_last;
    const int lane = threadIdx.x % WARP, warp = threadIdx.x / WARP;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double v = acc[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_xor_sync(0xffffffffu, v, o);
            v = (s == VIPRS_B200_S_MAX_DIFF) ? fmax(v, w) : v + w;
        }
        if (lane == 0) red[warp][s] = v;
    }
    __syncthreads();
    const size_t slot = (size_t)seg * ncol + c;
    if (threadIdx.x < NS) {
        const int s = threadIdx.x;
        double v = red[0][s];
        for (int w = 1; w < EM_THREADS / WARP; ++w) v = (s == VIPRS_B200_S_MAX_DIFF) ? fmax(v, red[w][s]) : v + red[w][s];
        partial[(slot * chunks + chunk) * NS + s] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&counters[slot], 1u) == (unsigned int)(chunks - 1));
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed-order, 16-way parallel: thread (k0, s) folds chunks k0, k0+16, ... of slot s, then slot s folds its 16 partials
        constexpr int KW = EM_THREADS / NS;
        const int s = threadIdx.x % NS, k0 = threadIdx.x / NS;
        const double* pp = partial + slot * chunks * NS;
        double v = 0.0;
        for (int k = k0; k < chunks; k += KW) {
            const double w = __ldcg(pp + k * NS + s);
            v = (s == VIPRS_B200_S_MAX_DIFF) ? fmax(v, w) : v + w;
        }
        __shared__ double red2[KW][NS];
        red2[k0][s] = v;
        __syncthreads();
        if (threadIdx.x < NS) {
            double t = red2[0][s];
            for (int k = 1; k < KW; ++k) t = (s == VIPRS_B200_S_MAX_DIFF) ? fmax(t, red2[k][s]) : t + red2[k][s];
            sums[slot * NS + s] = t;
        }
        if (threadIdx.x == 0) counters[slot] = 0u;         // ready for the next call on this stream
    }
}

// CTAs per (segment, column): ~2 rows per thread (the float64 logarithms dominate, so parallelism matters more than
// bytes), capped so that the whole grid stays within 64K CTAs and the partial buffer within a few tens of MB.
static int sums_chunks(int M, int ncol, int nseg) {
    const int per_seg = (M + nseg - 1) / (nseg > 0 ? nseg : 1);
    int ch = (per_seg + 2 * EM_THREADS - 1) / (2 * EM_THREADS);
    const int64_t slots = (int64_t)nseg * ncol;
    const int cap = (int)(65536 / (slots > 0 ? slots : 1));
    if (ch > cap) ch = cap;
    if (ch > 1024) ch = 1024;
    if (ch < 1) ch = 1;
    return ch;
}

template <typename T>
static int prepare_dispatch(int M, int ncol, int layout, int half_tau, const double* n, const double* theta, T* u_logs,
                            T* tau_term, T* mu_mult, T* log_null_pi, cudaStream_t st) {
    if (M <= 0 || ncol <= 0 || !n || !theta || !u_logs || !tau_term || !mu_mult) return VIPRS_B200_EINVAL;
    if (layout != 0 && layout != 1) return VIPRS_B200_EINVAL;
    // many columns (grid): ~8 rows per thread amortise LogNear's first logarithm; few columns: one row per thread,
    // the launch is latency-sized and wants every SM busy
    const int rpt = ncol >= 8 ? 8 : 1;
    int bx = (M + rpt * EM_THREADS - 1) / (rpt * EM_THREADS);
    if (bx > 1184) bx = 1184;                       // 8 x 148: grid-stride beyond that
    if (bx < 1) bx = 1;
    dim3 grid(bx, ncol);
    prepare_kernel<T><<<grid, EM_THREADS, 0, st>>>(M, ncol, layout, half_tau, n, reinterpret_cast<const Theta*>(theta),
                                                   u_logs, tau_term, mu_mult, log_null_pi);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

template <typename T>
static int sums_dispatch(int M, int ncol, int layout, int nseg, const int32_t* seg_ptr, const T* var_gamma, const T* var_mu,
                         const T* eta, const T* q, const T* eta_diff, const T* std_beta, const double* n,
                         const double* theta, const double* theta_logtau, double q_scale, void* workspace,
                         int64_t workspace_bytes, double* sums, cudaStream_t st) {
    if (M <= 0 || ncol <= 0 || nseg <= 0 || !seg_ptr || !var_gamma || !var_mu || !eta || !q || !eta_diff || !std_beta ||
        !n || !theta || !workspace || !sums)
        return VIPRS_B200_EINVAL;
    if (layout != 0 && layout != 1) return VIPRS_B200_EINVAL;
    if (ncol > 65535 || nseg > 65535) return VIPRS_B200_EINVAL;
    const int chunks = sums_chunks(M, ncol, nseg);
    const int64_t need_c = (((int64_t)nseg * ncol * 4 + 255) / 256) * 256;
    const int64_t need = need_c + (int64_t)nseg * ncol * chunks * NS * 8;
    if (workspace_bytes < need) return VIPRS_B200_EINVAL;
    unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);      // zeroed by the caller once; self-resetting
    double* partial = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + need_c);
    dim3 grid(chunks, nseg, ncol);
    sums_kernel<T><<<grid, EM_THREADS, 0, st>>>(M, ncol, layout, seg_ptr, var_gamma, var_mu, eta, q, eta_diff, std_beta, n,
                                                reinterpret_cast<const Theta*>(theta),
                                                reinterpret_cast<const Theta*>(theta_logtau ? theta_logtau : theta), q_scale,
                                                partial, counters, sums);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

// ---------------------------------------------------------------------------------------------------------------
// The scalar side of one EM iteration on the device: M-step, ELBO, MSE, heritability and the convergence scalar from
// the reduced sums -- VIPRS.m_step (VIPRS.py:426-484), elbo (:497-581), mse (:689-704), get_heritability (:780-785);
// VIPRSMix.update_pi / update_tau_beta (VIPRSMix.py:227-260).  Float64 throughout, one thread per model column (the
// mixture is one model: thread 0).  With this kernel an EM iteration is prepare -> sweep -> sums -> update with no
// host round trip: theta is rewritten in place for the next prepare, and the per-iteration scalars go to a history
// ring the host reads whenever it wants to look (every iteration for the reference's exact stopping rules, or every
// k iterations).
// ---------------------------------------------------------------------------------------------------------------
struct EmUpdateArgs {
    const double* sums;        // [nseg][ncol][NS]; world > 1: the all-reduced table
    const double* max_onehot;  // nullable: [world][nseg][ncol] maxima of |eta_diff| per rank (the summed MAX_DIFF slot is void)
    const double* seg_sizes;   // [nseg] SNPs per chromosome (global)
    const int32_t* flags;      // [ncol] bit 0: pi fixed, 1: tau_beta fixed, 2: sigma_epsilon fixed; mixture: flags[0] bit 0:
                               //        'pis' fixed, 1: 'tau_betas' fixed, 2: sigma_epsilon fixed, 3: total 'pi' fixed
    const double* mix_d;       // mixture: [K] prior multipliers
    double* theta;             // [ncol][4] in/out: sigma_epsilon, tau_beta, pi, lambda_min
    double* theta_prev;        // [ncol][4] out: theta as the sweep of this iteration saw it
    double* sigma_g;           // [ncol] out (mixture: [0])
    double* scalars;           // [hist_len][ncol][8] ring: ELBO, mse, max |eta_diff|, h2, pi, tau_beta, sigma_epsilon, sigma_g
    int32_t* iter;             // device iteration counter (incremented by this kernel)
    double n_snps, n, mix_fix_pi;
    int nseg, ncol, world, layout, hist_len;
};

__global__ void em_update_kernel(const EmUpdateArgs a) {
    const int nseg = a.nseg, ncol = a.ncol;
    const int it = *a.iter;
    double* out = a.scalars + (size_t)(it % a.hist_len) * ncol * 8;
    auto S = [&](int sg, int c, int slot) { return a.sums[((size_t)sg * ncol + c) * NS + slot]; };
    auto maxdiff = [&](int c) {
        double m = 0.0;
        for (int sg = 0; sg < nseg; ++sg) {
            if (a.max_onehot != nullptr) {
                for (int r = 0; r < a.world; ++r) m = fmax(m, a.max_onehot[((size_t)r * nseg + sg) * ncol + c]);
            } else {
                m = fmax(m, S(sg, c, VIPRS_B200_S_MAX_DIFF));
            }
        }
        return m;
    };
    if (a.layout == 0) {
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x) {
            double se = a.theta[c * 4 + 0], tau = a.theta[c * 4 + 1], pi = a.theta[c * 4 + 2];
            const double lam = a.theta[c * 4 + 3];
            for (int k = 0; k < 4; ++k) a.theta_prev[c * 4 + k] = a.theta[c * 4 + k];
            const int fl = a.flags[c];
            double T[NS];
            for (int s = 0; s < NS; ++s) T[s] = 0.0;
            double pi_mean = 0.0, zeta_tot = 0.0, sg_sum = 0.0;
            for (int sg = 0; sg < nseg; ++sg) {
                for (int s = 0; s < NS; ++s) T[s] += S(sg, c, s);
                const double zeta = S(sg, c, VIPRS_B200_S_GAMMA_MU2) + S(sg, c, VIPRS_B200_S_G_INV_TAU);
                pi_mean += S(sg, c, VIPRS_B200_S_GAMMA) / a.seg_sizes[sg];      // dict_mean: mean of per-chromosome means
                zeta_tot += zeta;
                sg_sum += (1.0 + lam) * zeta + S(sg, c, VIPRS_B200_S_ETA_Q);      // VIPRS.py:454-457
            }
            pi_mean /= (double)nseg;
            if (!(fl & 1)) pi = pi_mean;                                          // :434
            if (!(fl & 2)) tau = pi * a.n_snps / zeta_tot;                        // :444
            const double sigma_g = sg_sum;
            if (!(fl & 4)) se = 1.0 - 2.0 * T[VIPRS_B200_S_BETA_ETA] + sigma_g;   // :466-471
            double e = -log(2.0 * 3.141592653589793 * se);                        // :545
            if (!(fl & 4)) e -= 1.0;                                              // :552
            else e -= (1.0 / se) * (1.0 - 2.0 * T[VIPRS_B200_S_BETA_ETA] + sigma_g);   // :558
            e *= 0.5 * a.n;                                                       // :560
            e -= T[VIPRS_B200_S_G_LOGG] - log(pi) * T[VIPRS_B200_S_GCLIP];        // :562
            e -= T[VIPRS_B200_S_NG_LOGNG] - log(1.0 - pi) * T[VIPRS_B200_S_NGCLIP];   // :563
            e += 0.5 * (T[VIPRS_B200_S_GCLIP] * (1.0 + log(tau)) - T[VIPRS_B200_S_G_LOG_TAU]);   // :565
            e -= 0.5 * tau * (T[VIPRS_B200_S_GAMMA_MU2] + T[VIPRS_B200_S_G_INV_TAU]);   // :568
            const double zt = T[VIPRS_B200_S_GAMMA_MU2] + T[VIPRS_B200_S_G_INV_TAU];
            const double mse = 1.0 - 2.0 * T[VIPRS_B200_S_BETA_ETA] + (sigma_g - zt + T[VIPRS_B200_S_ETA2]);   // :689-704
            a.theta[c * 4 + 0] = se; a.theta[c * 4 + 1] = tau; a.theta[c * 4 + 2] = pi;
            a.sigma_g[c] = sigma_g;
            double* o = out + (size_t)c * 8;
            o[0] = e; o[1] = mse; o[2] = maxdiff(c); o[3] = sigma_g / (sigma_g + se);
            o[4] = pi; o[5] = tau; o[6] = se; o[7] = sigma_g;
        }
    } else if (blockIdx.x == 0 && threadIdx.x == 0) {
        // sparse mixture: one model with K = ncol components (VIPRSMix.py:227-260 + VIPRS.py:454-471, 497-581)
        const int K = ncol;
        const int fl = a.flags[0];
        double se = a.theta[0];
        const double lam = a.theta[3];
        for (int k = 0; k < 4 * K; ++k) a.theta_prev[k] = a.theta[k];
        double pis[16], taus[16], zetas[16], gam[16];
        double pi_sum_new = 0.0;
        for (int k = 0; k < K; ++k) {
            pis[k] = a.theta[k * 4 + 2]; taus[k] = a.theta[k * 4 + 1];
            double g = 0.0, z = 0.0;
            for (int sg = 0; sg < nseg; ++sg) {
                g += S(sg, k, VIPRS_B200_S_GAMMA);
                z += S(sg, k, VIPRS_B200_S_GAMMA_MU2) + S(sg, k, VIPRS_B200_S_G_INV_TAU);
            }
            gam[k] = g; zetas[k] = z; pi_sum_new += g;
        }
        if (!(fl & 1)) {
            for (int k = 0; k < K; ++k) pis[k] = (fl & 8) ? a.mix_fix_pi * gam[k] / pi_sum_new : gam[k] / a.n_snps;   // :237-239
        }
        double pi_tot = 0.0, dz = 0.0, zsum = 0.0;
        for (int k = 0; k < K; ++k) { pi_tot += pis[k]; dz += a.mix_d[k] * zetas[k]; zsum += zetas[k]; }
        if (!(fl & 2)) {
            const double t = pi_tot * a.n_snps / dz;                               // :257
            for (int k = 0; k < K; ++k) taus[k] = fmax(a.mix_d[k] * t, 1.0);       // :258-260
        }
        double eq = 0.0, be = 0.0, e2 = 0.0, nglog = 0.0, ngc = 0.0;
        for (int sg = 0; sg < nseg; ++sg) {
            eq += S(sg, 0, VIPRS_B200_S_ETA_Q); be += S(sg, 0, VIPRS_B200_S_BETA_ETA); e2 += S(sg, 0, VIPRS_B200_S_ETA2);
            nglog += S(sg, 0, VIPRS_B200_S_NG_LOGNG); ngc += S(sg, 0, VIPRS_B200_S_NGCLIP);
        }
        const double sigma_g = (1.0 + lam) * zsum + eq;
        if (!(fl & 4)) se = 1.0 - 2.0 * be + sigma_g;
        double e = -log(2.0 * 3.141592653589793 * se);
        if (!(fl & 4)) e -= 1.0;
        else e -= (1.0 / se) * (1.0 - 2.0 * be + sigma_g);
        e *= 0.5 * a.n;
        double t1 = 0.0, t3 = 0.0, t4 = 0.0;
        for (int k = 0; k < K; ++k) {
            double glogg = 0.0, gclip = 0.0, glogtau = 0.0, gczeta = 0.0;
            for (int sg = 0; sg < nseg; ++sg) {
                glogg += S(sg, k, VIPRS_B200_S_G_LOGG); gclip += S(sg, k, VIPRS_B200_S_GCLIP);
                glogtau += S(sg, k, VIPRS_B200_S_G_LOG_TAU); gczeta += S(sg, k, VIPRS_B200_S_GC_ZETA);
            }
            t1 += glogg - log(pis[k]) * gclip;
            t3 += gclip * (1.0 + log(taus[k])) - glogtau;
            t4 += taus[k] * gczeta;
        }
        e -= t1;
        e -= nglog - log(1.0 - pi_tot) * ngc;
        e += 0.5 * t3;
        e -= 0.5 * t4;
        const double mse = 1.0 - 2.0 * be + (sigma_g - zsum + e2);
        for (int k = 0; k < K; ++k) { a.theta[k * 4 + 0] = se; a.theta[k * 4 + 1] = taus[k]; a.theta[k * 4 + 2] = pis[k]; }
        a.sigma_g[0] = sigma_g;
        double* o = out;
        o[0] = e; o[1] = mse; o[2] = maxdiff(0); o[3] = sigma_g / (sigma_g + se);
        o[4] = pi_tot; o[5] = taus[K - 1]; o[6] = se; o[7] = sigma_g;
    }
    // every thread has read *a.iter before any increment: the kernel is launched with ONE CTA
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.iter = it + 1;
}

}  // namespace vb

extern "C" int viprs_b200_em_update(int32_t nseg, int32_t ncol, int32_t layout, int32_t world, const double* sums,
                                    const double* max_onehot, const double* seg_sizes, const int32_t* flags,
                                    const double* mix_d, double n_snps, double n, double mix_fix_pi, double* theta,
                                    double* theta_prev, double* sigma_g, double* scalars, int32_t hist_len, int32_t* iter,
                                    void* stream) {
    if (nseg <= 0 || ncol <= 0 || !sums || !seg_sizes || !flags || !theta || !theta_prev || !sigma_g || !scalars || !iter ||
        hist_len <= 0)
        return VIPRS_B200_EINVAL;
    if (layout != 0 && layout != 1) return VIPRS_B200_EINVAL;
    if (layout == 1 && (ncol > 16 || !mix_d)) return VIPRS_B200_EINVAL;
    vb::EmUpdateArgs a{sums, max_onehot, seg_sizes, flags, mix_d, theta, theta_prev, sigma_g, scalars, iter,
                       n_snps, n, mix_fix_pi, nseg, ncol, world > 0 ? world : 1, layout, hist_len};
    vb::em_update_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

extern "C" int64_t viprs_b200_sums_workspace_bytes(int32_t M, int32_t ncol, int32_t nseg) {
    if (M <= 0 || ncol <= 0 || nseg <= 0) return 0;
    const int chunks = vb::sums_chunks(M, ncol, nseg);
    return (((int64_t)nseg * ncol * 4 + 255) / 256) * 256 + (int64_t)nseg * ncol * chunks * vb::NS * 8;
}

extern "C" int viprs_b200_prepare_f32(int32_t M, int32_t ncol, int32_t layout, int32_t half_tau, const double* n_per_snp,
                                      const double* theta, float* u_logs, float* tau_term, float* mu_mult,
                                      float* log_null_pi, void* stream) {
    return vb::prepare_dispatch<float>(M, ncol, layout, half_tau, n_per_snp, theta, u_logs, tau_term, mu_mult, log_null_pi,
                                       (cudaStream_t)stream);
}
extern "C" int viprs_b200_prepare_f64(int32_t M, int32_t ncol, int32_t layout, int32_t half_tau, const double* n_per_snp,
                                      const double* theta, double* u_logs, double* tau_term, double* mu_mult,
                                      double* log_null_pi, void* stream) {
    return vb::prepare_dispatch<double>(M, ncol, layout, half_tau, n_per_snp, theta, u_logs, tau_term, mu_mult, log_null_pi,
                                        (cudaStream_t)stream);
}
extern "C" int viprs_b200_sums_f32(int32_t M, int32_t ncol, int32_t layout, int32_t nseg, const int32_t* seg_ptr,
                                   const float* var_gamma, const float* var_mu, const float* eta, const float* q,
                                   const float* eta_diff, const float* std_beta, const double* n_per_snp,
                                   const double* theta, const double* theta_logtau, double q_scale, void* workspace,
                                   int64_t workspace_bytes, double* sums, void* stream) {
    return vb::sums_dispatch<float>(M, ncol, layout, nseg, seg_ptr, var_gamma, var_mu, eta, q, eta_diff, std_beta, n_per_snp,
                                    theta, theta_logtau, q_scale, workspace, workspace_bytes, sums, (cudaStream_t)stream);
}
extern "C" int viprs_b200_sums_f64(int32_t M, int32_t ncol, int32_t layout, int32_t nseg, const int32_t* seg_ptr,
                                   const double* var_gamma, const double* var_mu, const double* eta, const double* q,
                                   const double* eta_diff, const double* std_beta, const double* n_per_snp,
                                   const double* theta, const double* theta_logtau, double q_scale, void* workspace,
                                   int64_t workspace_bytes, double* sums, void* stream) {
    return vb::sums_dispatch<double>(M, ncol, layout, nseg, seg_ptr, var_gamma, var_mu, eta, q, eta_diff, std_beta, n_per_snp,
                                     theta, theta_logtau, q_scale, workspace, workspace_bytes, sums, (cudaStream_t)stream);
}
