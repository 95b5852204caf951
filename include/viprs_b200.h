/*
 * viprs_b200 -- C ABI of the B200-native coordinate-ascent E-step.
 *
 * This is the drop-in boundary for the reference's Cython layer
 * /root/reference/viprs/model/vi/e_step_cpp.pyx (cpp_e_step :91-122, cpp_e_step_mixture :125-159,
 * cpp_e_step_grid :161-195, check_omp_support/check_blas_support :71-76), i.e. what a ctypes / cffi
 * binding on the reference side would load instead of the compiled `e_step_cpp` module.
 * Plain pointers and sizes only; no torch / numpy types.  Every function returns 0 on success,
 * a negative VIPRS_B200_E* code for argument errors, or a positive cudaError_t.  Nothing throws.
 *
 * Conventions
 *   T  : floating type of the variational state (f32 | f64)         -- `floating` in e_step_cpp.pxd
 *   U  : LD storage type (i8 | i16 | f32 | f64)                     -- `noncomplex_numeric` (.pxd:11-17)
 *   I  : CSR index type (int32 | int64)                             -- `indptr_type` (.pxd:7-9)
 *   LD : magenpy's "CSR without column indices": row j stores the contiguous column run
 *        [left_bound[j], left_bound[j] + indptr[j+1]-indptr[j])      (e_step.hpp:389-392).
 *        Both the upper-triangular (`low_memory=True`, diagonal excluded) and the symmetric layout
 *        (`low_memory=False`, run includes the unit diagonal) are accepted; only the strictly
 *        upper part (column > row) is kept on the device.
 *   "dev" pointers are CUDA device pointers; "host" pointers are ordinary host memory.
 *   `stream` is a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef VIPRS_B200_H
#define VIPRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- dtype / memory-kind enums ------------------------------------------------------------ */
#define VIPRS_B200_I8   0
#define VIPRS_B200_I16  1
#define VIPRS_B200_F32  2
#define VIPRS_B200_F64  3

#define VIPRS_B200_MEM_HOST   0
#define VIPRS_B200_MEM_DEVICE 1

/* ---- error codes (negative; positive values are cudaError_t) ------------------------------ */
#define VIPRS_B200_OK               0
#define VIPRS_B200_EINVAL          -1   /* bad argument (null pointer, bad dtype, M <= 0 ...)        */
#define VIPRS_B200_ELAYOUT         -2   /* LD arrays inconsistent (run leaves [0,M), negative length) */
#define VIPRS_B200_EBLOCK_TOO_LARGE -3  /* grid sweep only: an LD block (connected run of overlapping
                                           rows) is larger than 4096 SNPs.  The single-model and mixture
                                           sweeps tile larger blocks (and banded / windowed LD)       */
#define VIPRS_B200_ENOMEM          -4
#define VIPRS_B200_ENODEVICE       -5   /* no CUDA device: there is NO CPU fallback                  */
#define VIPRS_B200_EUNSUPPORTED    -6   /* (T,U) combination or K not built                          */

/* ---- library / device queries ------------------------------------------------------------- */
const char* viprs_b200_version(void);
const char* viprs_b200_strerror(int code);
/* replaces check_omp_support / check_blas_support (e_step_cpp.pyx:71-76): number of visible CUDA
 * devices (0 => every compute entry point returns VIPRS_B200_ENODEVICE). */
int viprs_b200_device_count(void);

/* ---- device-resident LD matrix ------------------------------------------------------------ */
typedef struct viprs_b200_ld viprs_b200_ld_t;

typedef struct {
    int32_t M;               /* rows (SNPs)                                                   */
    int32_t ld_dtype;        /* VIPRS_B200_I8 ...                                             */
    int32_t n_blocks;        /* independent LD blocks found                                   */
    int32_t max_block;       /* rows in the largest block                                     */
    int32_t n_panels;        /* row panels (units of the TMA ring)                            */
    int32_t stage_bytes;     /* shared-memory ring stage size the panels were cut for         */
    int64_t nnz;             /* strictly-upper stored entries (algorithmic LD elements)       */
    int64_t packed_elems;    /* elements in the 16B-aligned device layout (nnz + padding)     */
    int64_t smem_bytes;      /* dynamic shared memory per CTA the float32 sweep will request  */
    int32_t ring_stages;     /* depth of the TMA ring (float32 state)                         */
    int32_t ctas_per_sm;     /* CTAs (LD blocks) that share one SM (float32 state)            */
    int32_t n_units;         /* sweep units: n_blocks, plus the extra tiles of LD blocks larger
                                than 4096 rows (tiled sweep, see viprs_b200_e_step_*)         */
    int32_t n_phases;        /* sweep launches per E-step: 1, or the tile count of the largest
                                LD block                                                      */
    int64_t ext_elems;       /* elements stored in the between-tile rectangles (0 if untiled) */
} viprs_b200_ld_info_t;

/* Replaces the reference's LD load (VIPRS.py:153-172: ld_mat.load(...) -> ld_data/ld_indptr/
 * leftmost_idx kept in host RAM).  Copies (if mem_kind == HOST) and re-lays the matrix into the
 * 16-byte-aligned row layout described in DESIGN.md, finds the independent LD blocks and cuts the
 * row panels.  indptr_is_i64: 1 for int64 indptr, 0 for int32.  stage_bytes = 0 picks the default. */
int viprs_b200_ld_create(viprs_b200_ld_t** out, int32_t M, const int32_t* left_bound,
                         const void* indptr, int32_t indptr_is_i64, const void* ld_data,
                         int32_t ld_dtype, int32_t mem_kind, int32_t stage_bytes, void* stream);
int viprs_b200_ld_info(const viprs_b200_ld_t* ld, viprs_b200_ld_info_t* info);
/* first row of every LD block, n_blocks+1 entries (host pointer out) */
int viprs_b200_ld_block_rows(const viprs_b200_ld_t* ld, int32_t* out_host);
int viprs_b200_ld_destroy(viprs_b200_ld_t* ld);

/* ---- E-step sweeps on a device-resident LD matrix (all array arguments are DEVICE pointers) --
 *
 * viprs_b200_e_step_{f32,f64}: one Gauss-Seidel sweep with the semantics of
 * e_step<T,U,I>(..., threads=1, low_memory=true) (e_step.hpp:343-442), same argument meaning and
 * order as cpp_e_step (e_step_cpp.pyx:91-105).  `threads` / `low_memory` have no equivalent: the
 * sweep is always the strictly sequential order and the LD is always upper-triangular on device.
 *
 * q handling: the sweep reads the LD once.  It leaves in `q` the forward part
 * dq_scale * sum_{i<j} R_ij eta_i(new) (+ q_offset).  If materialize_q != 0 a second streaming pass adds
 * the backward part so that `q` equals the reference's q after its second pass
 * (update_q_factor, e_step.hpp:307-338).  The array `q` itself is not read on entry: the sweep recomputes
 * dq (R - I) eta from `eta`.  The reference instead maintains q incrementally (e_step.hpp:421,439), so
 * whatever its q holds beyond dq (R - I) eta on entry stays there for ever -- e.g. q = 0 next to eta != 0
 * after a `param_0` warm start (VIPRS.py:339-357).  To reproduce that, compute that constant part ONCE with
 * viprs_b200_q_offset_*(ld, eta_in, q_in, dq, out) and pass it as `q_offset` (device array, M entries, q
 * units; NULL = none) to every sweep.  The host drop-ins viprs_b200_cpp_e_step* always do this.
 *
 * LD blocks larger than 4096 SNPs (e.g. 10,240-SNP float64 blocks, or banded / windowed LD where a whole
 * chromosome is one block) are swept in 1024-row tiles: same per-SNP order, one launch per tile index plus
 * streaming products over the rectangles between tiles (viprs_b200_ld_info_t.n_phases launches).
 *
 * The reductions the M-step / ELBO need are produced by viprs_b200_sums_* (VIPRS_B200_S_* slots below).
 */
#define VIPRS_B200_NSUMS 16
#define VIPRS_B200_S_GAMMA        0  /* sum gamma                             (VIPRS.py:434, VIPRSMix.py:233)  */
#define VIPRS_B200_S_GAMMA_MU2    1  /* sum gamma mu^2                        (zeta, VIPRS.py:896)             */
#define VIPRS_B200_S_ETA_Q        2  /* q_scale * sum eta q                   (VIPRS.py:455)                   */
#define VIPRS_B200_S_BETA_ETA     3  /* sum std_beta eta                      (VIPRS.py:469)                   */
#define VIPRS_B200_S_G_LOGG       4  /* sum gc log gc, gc = clip(gamma, 1e-15, 1-1e-15)   (VIPRS.py:509,562)   */
#define VIPRS_B200_S_NG_LOGNG     5  /* sum ngc log ngc, ngc = clip(1 - pip)  (VIPRS.py:514-518,563)           */
#define VIPRS_B200_S_ETA2         6  /* sum eta^2                             (VIPRS.py:703)                   */
#define VIPRS_B200_S_MAX_DIFF     7  /* max |eta_diff|                        (VIPRS.py:997)                   */
#define VIPRS_B200_S_G_INV_TAU    8  /* sum gamma / var_tau                   (zeta, VIPRS.py:896)             */
#define VIPRS_B200_S_G_LOG_TAU    9  /* sum gc log var_tau(theta_logtau)      (VIPRS.py:519,565)               */
#define VIPRS_B200_S_GCLIP        10 /* sum gc                                (VIPRS.py:562,565)               */
#define VIPRS_B200_S_GC_ZETA      11 /* sum gc (mu^2 + 1/var_tau)             (VIPRS.py:571-573, mixture)      */
#define VIPRS_B200_S_NGCLIP       12 /* sum ngc                               (VIPRS.py:563)                   */

int viprs_b200_e_step_f32(const viprs_b200_ld_t* ld, const float* std_beta, float* var_gamma,
                          float* var_mu, float* eta, float* q, float* eta_diff,
                          const float* u_logs, const float* sqrt_half_var_tau, const float* mu_mult,
                          float dq_scale, int32_t materialize_q, const float* q_offset, void* stream);
int viprs_b200_e_step_f64(const viprs_b200_ld_t* ld, const double* std_beta, double* var_gamma,
                          double* var_mu, double* eta, double* q, double* eta_diff,
                          const double* u_logs, const double* sqrt_half_var_tau,
                          const double* mu_mult, double dq_scale, int32_t materialize_q,
                          const double* q_offset, void* stream);
/* q_offset_out[j] = q[j] - dq_scale * sum_{k != j} R_jk eta[k]   (two streaming passes over the LD) */
int viprs_b200_q_offset_f32(const viprs_b200_ld_t* ld, const float* eta, const float* q, float dq_scale,
                            float* q_offset_out, void* stream);
int viprs_b200_q_offset_f64(const viprs_b200_ld_t* ld, const double* eta, const double* q, double dq_scale,
                            double* q_offset_out, void* stream);

/* viprs_b200_e_step_mixture_{f32,f64}: one sweep with the semantics of e_step_mixture<T,U,I>(..., threads=1,
 * low_memory=true) (e_step.hpp:447-551); argument meaning and order of cpp_e_step_mixture
 * (e_step_cpp.pyx:125-141).  var_gamma, var_mu, u_logs, sqrt_half_var_tau, mu_mult are (M,K) C-order;
 * eta, q, eta_diff, std_beta, log_null_pi have M entries.  1 <= K <= 16. */
int viprs_b200_e_step_mixture_f32(const viprs_b200_ld_t* ld, int32_t K, const float* std_beta, float* var_gamma,
                                  float* var_mu, float* eta, float* q, float* eta_diff, const float* log_null_pi,
                                  const float* u_logs, const float* sqrt_half_var_tau, const float* mu_mult,
                                  float dq_scale, int32_t materialize_q, const float* q_offset, void* stream);
int viprs_b200_e_step_mixture_f64(const viprs_b200_ld_t* ld, int32_t K, const double* std_beta, double* var_gamma,
                                  double* var_mu, double* eta, double* q, double* eta_diff,
                                  const double* log_null_pi, const double* u_logs, const double* sqrt_half_var_tau,
                                  const double* mu_mult, double dq_scale, int32_t materialize_q,
                                  const double* q_offset, void* stream);

/* viprs_b200_e_step_grid_{f32,f64}: one sweep with the semantics of e_step_grid<T,U,I>(..., threads=1)
 * followed by update_q_factor_matrix (e_step.hpp:555-647, 266-303); argument meaning of cpp_e_step_grid
 * (e_step_cpp.pyx:161-177).  var_gamma, var_mu, eta, q, eta_diff, u_logs, half_var_tau, mu_mult are (M,G)
 * COLUMN-MAJOR (index model*M + j); std_beta has M entries; active_model_idx is a DEVICE array of n_active
 * column indices -- the other columns are not touched.  q is in/out exactly as in the reference (maintained
 * incrementally across sweeps).  LD blocks of up to 4096 SNPs. */
int viprs_b200_e_step_grid_f32(const viprs_b200_ld_t* ld, int32_t G, int32_t n_active,
                               const int32_t* active_model_idx, const float* std_beta, float* var_gamma,
                               float* var_mu, float* eta, float* q, float* eta_diff, const float* u_logs,
                               const float* half_var_tau, const float* mu_mult, float dq_scale, void* stream);
int viprs_b200_e_step_grid_f64(const viprs_b200_ld_t* ld, int32_t G, int32_t n_active,
                               const int32_t* active_model_idx, const double* std_beta, double* var_gamma,
                               double* var_mu, double* eta, double* q, double* eta_diff, const double* u_logs,
                               const double* half_var_tau, const double* mu_mult, double dq_scale, void* stream);

/* q[j] += dq_scale * sum_{k>j} R_jk x[k]  -- the reference's update_q_factor (e_step.hpp:307-338)
 * as a stand-alone streaming kernel (x = eta for materialisation, or any vector). */
int viprs_b200_backward_dot_f32(const viprs_b200_ld_t* ld, const float* x, float* q,
                                float dq_scale, void* stream);
int viprs_b200_backward_dot_f64(const viprs_b200_ld_t* ld, const double* x, double* q,
                                double dq_scale, void* stream);

/* viprs_b200_e_step_incremental_f32 / viprs_b200_e_step_mixture_incremental_f32: cpp_e_step / cpp_e_step_mixture
 * (e_step_cpp.pyx:91-105, 125-141) on DEVICE arrays with the reference's own bookkeeping of q -- q is in/out and
 * maintained incrementally: inside the sweep q_j = q_in[j] + dq_scale * sum_{i<j} R_ij eta_diff_i (e_step.hpp:421),
 * after it q[j] += dq_scale * sum_{k>j} R_jk eta_diff_k (update_q_factor, e_step.hpp:435-440).  Whatever the caller's q
 * holds is honoured exactly as the reference does; the LD is read twice per call (like the reference), with no backward
 * dots inside the sequential sweep.  float32 state, LD blocks <= 4096 SNPs, K <= 4; VIPRS_B200_EUNSUPPORTED otherwise
 * (use viprs_b200_e_step_* with viprs_b200_q_offset_*). */
int viprs_b200_e_step_incremental_f32(const viprs_b200_ld_t* ld, const float* std_beta, float* var_gamma, float* var_mu,
                                      float* eta, float* q, float* eta_diff, const float* u_logs,
                                      const float* sqrt_half_var_tau, const float* mu_mult, float dq_scale, void* stream);
int viprs_b200_e_step_mixture_incremental_f32(const viprs_b200_ld_t* ld, int32_t K, const float* std_beta, float* var_gamma,
                                              float* var_mu, float* eta, float* q, float* eta_diff,
                                              const float* log_null_pi, const float* u_logs,
                                              const float* sqrt_half_var_tau, const float* mu_mult, float dq_scale,
                                              void* stream);

/* ---- one-shot host-pointer drop-ins: exactly the reference's cpdef signatures -----------------
 * Same arrays, same in-place outputs as cpp_e_step(...) (e_step_cpp.pyx:91-122) with host (numpy)
 * buffers: upload, pack, sweep, materialise q, download.  `threads` is accepted and ignored
 * (the result is the threads=1 result); `low_memory` selects how the LD run is interpreted
 * (1: upper-triangular, 0: symmetric with diagonal).  q is in/out like the reference's: the part of
 * the incoming q that eta does not explain is carried through the sweep (viprs_b200_q_offset_*). */
int viprs_b200_cpp_e_step(int32_t M, const int32_t* ld_left_bound, const void* ld_indptr,
                          int32_t indptr_is_i64, const void* ld_data, int32_t ld_dtype,
                          int32_t float_dtype, const void* std_beta, void* var_gamma, void* var_mu,
                          void* eta, void* q, void* eta_diff, const void* u_logs,
                          const void* sqrt_half_var_tau, const void* mu_mult, double dq_scale,
                          int32_t threads, int32_t low_memory);

/* cpp_e_step_mixture(...) (e_step_cpp.pyx:125-159) with host buffers; K = var_mu.shape[1]. */
int viprs_b200_cpp_e_step_mixture(int32_t M, int32_t K, const int32_t* ld_left_bound, const void* ld_indptr,
                                  int32_t indptr_is_i64, const void* ld_data, int32_t ld_dtype,
                                  int32_t float_dtype, const void* std_beta, void* var_gamma, void* var_mu,
                                  void* eta, void* q, void* eta_diff, const void* log_null_pi, const void* u_logs,
                                  const void* sqrt_half_var_tau, const void* mu_mult, double dq_scale,
                                  int32_t threads, int32_t low_memory);

/* The same two calls for a caller that keeps the LD resident on the device (viprs_b200_ld_create once) but its state in
 * host memory, as the reference does: per call the arrays cpp_e_step reads go host->device (std_beta, var_gamma, var_mu,
 * eta, q, u_logs, sqrt_half_var_tau, mu_mult), one sweep runs, q is materialised, and what cpp_e_step writes
 * (var_gamma, var_mu, eta, q, eta_diff) comes back; returns after the stream has drained.  Pinned host buffers make the
 * copies asynchronous (and let the uploads of a chunk go out as one batched copy).  Where the incremental sweep applies (float32 state, LD blocks <= 4096 SNPs) the call runs
 * viprs_b200_e_step_*incremental_f32 in row chunks on internal streams, so that the copies of one chunk overlap the sweep
 * of another (everything but q goes back to the host while the update_q_factor pass of its chunk still runs), and
 * q_is_consistent is irrelevant.  VIPRS_B200_E2E_TIMING=1 in the environment prints the device timeline of a call, per
 * chunk, to stderr.  Otherwise: q_is_consistent != 0: the caller vouches that
 * q = dq (R - I) eta on entry (true on every iteration of VIPRS.fit unless `param_0` was given) and the two extra LD
 * passes of viprs_b200_q_offset_* are skipped. */
int viprs_b200_cpp_e_step_resident(const viprs_b200_ld_t* ld, int32_t float_dtype, const void* std_beta, void* var_gamma,
                                   void* var_mu, void* eta, void* q, void* eta_diff, const void* u_logs,
                                   const void* sqrt_half_var_tau, const void* mu_mult, double dq_scale,
                                   int32_t q_is_consistent, void* stream);
int viprs_b200_cpp_e_step_mixture_resident(const viprs_b200_ld_t* ld, int32_t K, int32_t float_dtype, const void* std_beta,
                                           void* var_gamma, void* var_mu, void* eta, void* q, void* eta_diff,
                                           const void* log_null_pi, const void* u_logs, const void* sqrt_half_var_tau,
                                           const void* mu_mult, double dq_scale, int32_t q_is_consistent, void* stream);

/* ---- the per-iteration work around the sweep ------------------------------------------------------------
 *
 * theta: DEVICE double[ncol][4] = {sigma_epsilon, tau_beta, pi, lambda_min} per model column (grid) or per mixture
 * component (sigma_epsilon and lambda_min repeated).  layout 0: (M, ncol) arrays are column-major (single model:
 * ncol = 1; grid), layout 1: row-major (mixture, ncol = K).
 *
 * viprs_b200_prepare_*: the numpy pre-compute of VIPRS.e_step() / VIPRSMix.e_step() (VIPRS.py:400-406,418;
 * VIPRSMix.py:187-204) in float64, cast to the state type: u_logs, mu_mult and tau_term = sqrt(var_tau/2)
 * (half_tau = 0; cpp_e_step / cpp_e_step_mixture) or var_tau/2 (half_tau = 1; cpp_e_step_grid).  log_null_pi
 * (M entries, mixture only) may be NULL. */
int viprs_b200_prepare_f32(int32_t M, int32_t ncol, int32_t layout, int32_t half_tau, const double* n_per_snp,
                           const double* theta, float* u_logs, float* tau_term, float* mu_mult, float* log_null_pi,
                           void* stream);
int viprs_b200_prepare_f64(int32_t M, int32_t ncol, int32_t layout, int32_t half_tau, const double* n_per_snp,
                           const double* theta, double* u_logs, double* tau_term, double* mu_mult,
                           double* log_null_pi, void* stream);

/* viprs_b200_sums_*: sums[nseg][ncol][VIPRS_B200_NSUMS] (device doubles) over the row segments
 * [seg_ptr[s], seg_ptr[s+1]) (chromosomes; device int32[nseg+1]) -- everything m_step() / elbo() / mse() and the
 * convergence test read from the per-SNP arrays (VIPRS.py:426-484, 497-581, 689-704, 997; VIPRSMix.py:227-260).
 * float64 accumulation in a fixed order (bit-reproducible).  eta / q / eta_diff / std_beta: M entries (layout 1)
 * or (M, ncol) column-major (layout 0); mixture per-SNP slots (ETA_Q, BETA_ETA, NG_*, ETA2, MAX_DIFF) are
 * reported in column 0.  q_scale multiplies sum eta*q: 2 when q holds only the forward part of the one-pass
 * sweep (materialize_q = 0), else 1.  theta_logtau (may be NULL = theta) parameterises the cached log var_tau
 * of the ELBO (VIPRS.py:401,519 -- VIPRSMix never refreshes it).  workspace: viprs_b200_sums_workspace_bytes()
 * bytes of device memory, zero-filled once by the caller. */
int64_t viprs_b200_sums_workspace_bytes(int32_t M, int32_t ncol, int32_t nseg);
int viprs_b200_sums_f32(int32_t M, int32_t ncol, int32_t layout, int32_t nseg, const int32_t* seg_ptr,
                        const float* var_gamma, const float* var_mu, const float* eta, const float* q,
                        const float* eta_diff, const float* std_beta, const double* n_per_snp, const double* theta,
                        const double* theta_logtau, double q_scale, void* workspace, int64_t workspace_bytes,
                        double* sums, void* stream);
int viprs_b200_sums_f64(int32_t M, int32_t ncol, int32_t layout, int32_t nseg, const int32_t* seg_ptr,
                        const double* var_gamma, const double* var_mu, const double* eta, const double* q,
                        const double* eta_diff, const double* std_beta, const double* n_per_snp, const double* theta,
                        const double* theta_logtau, double q_scale, void* workspace, int64_t workspace_bytes,
                        double* sums, void* stream);

/* viprs_b200_e_step_fused_f32: viprs_b200_e_step_f32 (materialize_q = 0, no q_offset) with the reductions of
 * viprs_b200_sums_f32 (ncol = 1, layout 0, q_scale = 2, theta_logtau = theta) FUSED into the sweep: the warp that writes a
 * row's outputs also accumulates its VIPRS_B200_S_* terms (float64, fixed order), one table row per LD block, folded per
 * chromosome segment by a one-CTA-per-segment kernel -- the per-SNP arrays are not read a second time.  theta / n_per_snp /
 * seg_ptr / sums as in viprs_b200_sums_f32.  Returns VIPRS_B200_EUNSUPPORTED for configurations the fused kernel does not
 * cover (float64 LD, LD blocks the register-resident kernel cannot take): run the sweep and the sums separately then. */
int viprs_b200_e_step_fused_f32(const viprs_b200_ld_t* ld, const float* std_beta, float* var_gamma, float* var_mu, float* eta,
                                float* q, float* eta_diff, const float* u_logs, const float* sqrt_half_var_tau,
                                const float* mu_mult, float dq_scale, const double* n_per_snp, const double* theta,
                                int32_t nseg, const int32_t* seg_ptr, double* sums, void* stream);

/* viprs_b200_em_update: the scalar side of one EM iteration ON THE DEVICE -- VIPRS.m_step (VIPRS.py:426-484), elbo
 * (:497-581), mse (:689-704), get_heritability (:780-785), VIPRSMix.update_pi / update_tau_beta (VIPRSMix.py:227-260) --
 * from the reduced sums, in float64.  theta (see above) is rewritten in place for the next viprs_b200_prepare_*;
 * theta_prev receives theta as this iteration's sweep saw it (var_tau of the outputs, VIPRS.py:888-897); per model
 * column scalars[(iter % hist_len)][col][8] = {ELBO, mse, max |eta_diff|, h2, pi, tau_beta, sigma_epsilon, sigma_g}
 * and the device counter *iter is incremented, so prepare -> sweep -> sums -> em_update can be replayed (e.g. as a CUDA
 * graph) without a host round trip.  flags[col]: bit 0 pi fixed, bit 1 tau_beta fixed, bit 2 sigma_epsilon fixed
 * (layout 1 = mixture, one model: flags[0] bit 0 'pis', bit 1 'tau_betas', bit 2 sigma_epsilon, bit 3 total 'pi' fixed at
 * mix_fix_pi; mix_d = prior multipliers).  seg_sizes: SNPs per chromosome; n_snps their sum; n = max n_per_snp.
 * world > 1: `sums` is the all-reduced table and max_onehot[world][nseg][ncol] carries the per-rank maxima. */
int viprs_b200_em_update(int32_t nseg, int32_t ncol, int32_t layout, int32_t world, const double* sums,
                         const double* max_onehot, const double* seg_sizes, const int32_t* flags, const double* mix_d,
                         double n_snps, double n, double mix_fix_pi, double* theta, double* theta_prev, double* sigma_g,
                         double* scalars, int32_t hist_len, int32_t* iter, void* stream);

/* cpp_e_step_grid(...) (e_step_cpp.pyx:161-195) with host buffers; (M,G) arrays Fortran-order, active_model_idx
 * a host array of n_active column indices.  q is in/out. */
int viprs_b200_cpp_e_step_grid(int32_t M, int32_t G, int32_t n_active, const int32_t* active_model_idx,
                               const int32_t* ld_left_bound, const void* ld_indptr, int32_t indptr_is_i64,
                               const void* ld_data, int32_t ld_dtype, int32_t float_dtype, const void* std_beta,
                               void* var_gamma, void* var_mu, void* eta, void* q, void* eta_diff, const void* u_logs,
                               const void* half_var_tau, const void* mu_mult, double dq_scale, int32_t threads,
                               int32_t low_memory);

#ifdef __cplusplus
}
#endif
#endif /* VIPRS_B200_H */
