// viprs_b200 -- C ABI entry points (include/viprs_b200.h): LD info and the one-shot host-pointer drop-ins.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ld.h"

extern "C" int viprs_b200_ld_info(const viprs_b200_ld_t* h, viprs_b200_ld_info_t* info) {
    if (!h || !info) return VIPRS_B200_EINVAL;
    info->M = h->M; info->ld_dtype = h->ld_dtype; info->n_blocks = h->n_blocks; info->max_block = h->max_block;
    info->n_panels = h->n_panels; info->stage_bytes = h->stage_bytes; info->nnz = h->nnz;
    info->packed_elems = h->packed_elems;
    info->n_blocks = h->n_ld_blocks; info->max_block = h->max_ld_block;
    info->n_units = h->n_blocks; info->n_phases = h->n_phases; info->ext_elems = h->ext_elems;
    vb::RingGeometry g = vb::fast_ring_geometry(h);
    if (g.nst == 0) g = vb::ring_geometry(h, 4);
    info->smem_bytes = g.smem_bytes;
    info->ring_stages = g.nst;
    info->ctas_per_sm = g.ctas_per_sm;
    return VIPRS_B200_OK;
}

// ---- host-pointer drop-ins ------------------------------------------------------------------------------
namespace {

struct DevBuf {            // scoped device allocation
    unsigned char* d = nullptr;
    ~DevBuf() { cudaFree(d); }
};

// internal streams / events of the chunked host-state round trip, created once per device on first use
struct HostStreams {
    static constexpr int kMax = 8;
    cudaStream_t s[kMax] = {};
    cudaStream_t out[kMax] = {};           // downloads of what is final before the update_q_factor pass
    cudaEvent_t join[kMax] = {};
    cudaEvent_t swept[kMax] = {};
    cudaEvent_t copied[kMax] = {};
    cudaEvent_t fork = nullptr;
    int n = 0;
};
HostStreams* host_streams(int n) {
    static HostStreams per_dev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || n > HostStreams::kMax) return nullptr;
    HostStreams& h = per_dev[dev];
    if (!h.fork && cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (; h.n < n; ++h.n) {
        if (cudaStreamCreateWithFlags(&h.s[h.n], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&h.out[h.n], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&h.join[h.n], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&h.swept[h.n], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&h.copied[h.n], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &h;
}

// One sweep with HOST state arrays on a device-resident LD matrix: upload what the reference's cpp_e_step* reads,
// sweep, materialise q, download what it writes.  K == 0: cpp_e_step; K > 0: cpp_e_step_mixture.  The staging buffer
// lives in the LD handle (grown on demand), so a per-iteration caller allocates nothing.
int state_roundtrip(const viprs_b200_ld_t* ld, int32_t K, int32_t float_dtype, const void* std_beta, void* var_gamma,
                    void* var_mu, void* eta, void* q, void* eta_diff, const void* log_null_pi, const void* u_logs,
                    const void* shvt, const void* mu_mult, double dq_scale, int32_t q_is_consistent, cudaStream_t st) {
    if (!ld) return VIPRS_B200_EINVAL;
    if (float_dtype != VIPRS_B200_F32 && float_dtype != VIPRS_B200_F64) return VIPRS_B200_EINVAL;
    if (!std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !shvt || !mu_mult)
        return VIPRS_B200_EINVAL;
    if (K > 0 && !log_null_pi) return VIPRS_B200_EINVAL;
    const int32_t M = ld->M;
    const size_t ts = float_dtype == VIPRS_B200_F32 ? 4 : 8;
    const size_t kk = K > 0 ? (size_t)K : 1;
    const size_t n1 = (((size_t)M * ts) + 255) & ~(size_t)255, nk = (((size_t)M * ts * kk) + 255) & ~(size_t)255;
    const size_t c1 = (size_t)M * ts, ck = c1 * kk;
    // layout: [beta][eta][q][diff][lnp][q offset] (n1 each) [gamma][mu][ulogs][shvt][mm] (nk each)
    const size_t need = 6 * n1 + 5 * nk;
    if (ld->host_ws_bytes < (int64_t)need) {
        cudaFree(ld->d_host_ws);
        ld->d_host_ws = nullptr; ld->host_ws_bytes = 0;
        cudaError_t ea = cudaMalloc(&ld->d_host_ws, need);
        if (ea != cudaSuccess) { cudaGetLastError(); return ea == cudaErrorMemoryAllocation ? VIPRS_B200_ENOMEM : (int)ea; }
        ld->host_ws_bytes = (int64_t)need;
    }
    unsigned char* d = reinterpret_cast<unsigned char*>(ld->d_host_ws);
    unsigned char *d_beta = d, *d_eta = d + n1, *d_q = d + 2 * n1, *d_diff = d + 3 * n1, *d_lnp = d + 4 * n1,
                  *d_off = d + 5 * n1;
    unsigned char *d_g = d + 6 * n1, *d_mu = d_g + nk, *d_ul = d_mu + nk, *d_sv = d_ul + nk, *d_mm = d_sv + nk;
    cudaError_t e = cudaSuccess;
    int rc = VIPRS_B200_OK;
    // ---- float32, register-resident sweep: the reference's incremental q (no q offset, no vouching), in row chunks
    // on internal streams -- chunk c's sweep overlaps chunk c+1's uploads and chunk c-1's downloads
    if (ts == 4 && (K == 0 || K <= 4) && getenv("VIPRS_B200_NO_INCREMENTAL") == nullptr) {
        using F = float;
        const int nch = ld->n_chunks > 0 ? ld->n_chunks : 1;
        HostStreams* hs = host_streams(nch);
        if (hs) {
            // VIPRS_B200_E2E_TIMING=1 (debug): device timeline of the call, per chunk, printed to stderr
            const bool timing = getenv("VIPRS_B200_E2E_TIMING") != nullptr;
            cudaEvent_t tev[1 + 3 * HostStreams::kMax] = {};
            if (timing) {
                for (int i = 0; i < 1 + 3 * nch; ++i) cudaEventCreate(&tev[i]);
                cudaEventRecord(tev[0], st);
            }
            // batched uploads need page-locked sources (a pageable source would have to be staged inside the call)
            auto pinned = [](const void* p_) {
                cudaPointerAttributes a_;
                if (cudaPointerGetAttributes(&a_, p_) != cudaSuccess) { cudaGetLastError(); return false; }
                return a_.type == cudaMemoryTypeHost;
            };
            bool batch_uploads = getenv("VIPRS_B200_NO_BATCH_COPY") == nullptr && pinned(std_beta) && pinned(eta) && pinned(q) &&
                                 pinned(var_gamma) && pinned(var_mu) && pinned(u_logs) && pinned(shvt) && pinned(mu_mult) &&
                                 (K == 0 || pinned(log_null_pi));
            // all chunks wait for whatever the caller queued on `st`
            if (e == cudaSuccess) e = cudaEventRecord(hs->fork, st);
            bool unsupported = false;
            for (int c = 0; c < nch && e == cudaSuccess && rc == 0; ++c) {
                cudaStream_t sc = hs->s[c];
                e = cudaStreamWaitEvent(sc, hs->fork, 0);
                const size_t r0 = ld->n_chunks > 0 ? (size_t)ld->h_chunk_row[c] : 0;
                const size_t r1 = ld->n_chunks > 0 ? (size_t)ld->h_chunk_row[c + 1] : (size_t)M;
                const size_t o1 = r0 * ts, b1 = (r1 - r0) * ts, ok_ = o1 * kk, bk = b1 * kk;
                // (moving the arrays with kernels over mapped page-locked memory instead of DMA copies was measured: same
                // PCIe rate, more contention with the sweeps -- 2.9 vs 2.5 ms per call; scripts/microbench/h2d_bench.cu)
                // The eight (nine) uploads of a chunk go out as ONE batched copy when the runtime has cudaMemcpyBatchAsync and
                // every source is page-locked: 53 instead of 46 GB/s for ~1 MB pieces (same micro-benchmark).
                void* bd[10]; void* bs[10]; size_t bn[10]; int nb_ = 0;
                auto upc = [&](unsigned char* dst, const void* src, size_t off, size_t n) {
                    if (src && n && nb_ < 10) {
                        bd[nb_] = dst + off; bs[nb_] = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(src)) + off;
                        bn[nb_] = n; ++nb_;
                    }
                };
                upc(d_beta, std_beta, o1, b1); upc(d_eta, eta, o1, b1); upc(d_q, q, o1, b1);
                if (K > 0) upc(d_lnp, log_null_pi, o1, b1);
                upc(d_g, var_gamma, ok_, bk); upc(d_mu, var_mu, ok_, bk); upc(d_ul, u_logs, ok_, bk); upc(d_sv, shvt, ok_, bk);
                upc(d_mm, mu_mult, ok_, bk);
                bool batched = false;
#if CUDART_VERSION >= 12080
                if (batch_uploads && e == cudaSuccess && nb_ > 1) {
                    cudaMemcpyAttributes at = {};
                    at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                    size_t idx0 = 0, fail = 0;
                    const cudaError_t eb = cudaMemcpyBatchAsync(bd, bs, bn, (size_t)nb_, &at, &idx0, 1, &fail, sc);
                    if (eb == cudaSuccess) batched = true;
                    else { cudaGetLastError(); batch_uploads = false; }      // not supported here: plain copies from now on
                }
#endif
                for (int i = 0; i < nb_ && !batched && e == cudaSuccess; ++i)
                    e = cudaMemcpyAsync(bd[i], bs[i], bn[i], cudaMemcpyHostToDevice, sc);
                if (e != cudaSuccess) break;
                if (timing) cudaEventRecord(tev[1 + 3 * c], sc);
                const int chunk = ld->n_chunks > 0 ? c : -1;
                rc = K > 0 ? vb::incr_mix_f32(ld, K, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta, (F*)d_q, (F*)d_diff, (F*)d_lnp,
                                              (F*)d_ul, (F*)d_sv, (F*)d_mm, (F)dq_scale, chunk, sc, hs->swept[c])
                           : vb::incr_slab_f32(ld, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta, (F*)d_q, (F*)d_diff, (F*)d_ul,
                                               (F*)d_sv, (F*)d_mm, (F)dq_scale, chunk, sc, hs->swept[c]);
                if (rc == VIPRS_B200_EUNSUPPORTED && c == 0) { unsupported = true; rc = 0; break; }
                if (timing) cudaEventRecord(tev[2 + 3 * c], sc);
                // everything but q is final once the sweep is done: it goes back on a second stream while the
                // update_q_factor pass runs; q follows the pass
                cudaStream_t so = hs->out[c];
                auto downc = [&](void* dst, unsigned char* src, size_t off, size_t n, cudaStream_t s_) {
                    if (e == cudaSuccess && rc == 0 && n)
                        e = cudaMemcpyAsync(reinterpret_cast<unsigned char*>(dst) + off, src + off, n, cudaMemcpyDeviceToHost, s_);
                };
                if (e == cudaSuccess && rc == 0) e = cudaStreamWaitEvent(so, hs->swept[c], 0);
                downc(var_gamma, d_g, ok_, bk, so); downc(var_mu, d_mu, ok_, bk, so); downc(eta, d_eta, o1, b1, so);
                downc(eta_diff, d_diff, o1, b1, so);
                if (e == cudaSuccess && rc == 0) e = cudaEventRecord(hs->copied[c], so);
                downc(q, d_q, o1, b1, sc);
                if (e == cudaSuccess && rc == 0) e = cudaStreamWaitEvent(sc, hs->copied[c], 0);
                if (timing) cudaEventRecord(tev[3 + 3 * c], sc);
                if (e == cudaSuccess) e = cudaEventRecord(hs->join[c], sc);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(st, hs->join[c], 0);
            }
            if (!unsupported) {
                cudaError_t e2 = cudaStreamSynchronize(st);
                for (int c = 0; c < nch; ++c) { cudaStreamSynchronize(hs->s[c]); cudaStreamSynchronize(hs->out[c]); }
                if (timing) {
                    for (int c = 0; c < nch; ++c) {
                        float a = 0, b = 0, d2 = 0;
                        cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * c]); cudaEventElapsedTime(&b, tev[0], tev[2 + 3 * c]);
                        cudaEventElapsedTime(&d2, tev[0], tev[3 + 3 * c]);
                        fprintf(stderr, "[viprs_b200 e2e] chunk %d: inputs on device %.3f ms, sweep + q pass done %.3f ms, outputs on host %.3f ms\n",
                                c, a, b, d2);
                    }
                    for (int i = 0; i < 1 + 3 * nch; ++i) cudaEventDestroy(tev[i]);
                }
                if (e == cudaSuccess) e = e2;
                if (rc) return rc;
                return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
            }
            // nothing was launched: chunk 0 reported that the incremental sweep does not cover this LD / model
            cudaStreamSynchronize(hs->s[0]);
            e = cudaSuccess;
        }
    }
    auto up = [&](void* dst, const void* src, size_t n) {
        if (e == cudaSuccess && src) e = cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
    };
    up(d_beta, std_beta, c1); up(d_eta, eta, c1); up(d_q, q, c1);
    if (K > 0) up(d_lnp, log_null_pi, c1);
    up(d_g, var_gamma, ck); up(d_mu, var_mu, ck); up(d_ul, u_logs, ck); up(d_sv, shvt, ck); up(d_mm, mu_mult, ck);
    // q is in/out in the reference (maintained incrementally): unless the caller vouches that q_in = dq (R - I) eta_in
    // (true on every iteration of VIPRS.fit without param_0), carry the part of q_in that eta_in does not explain
    if (e == cudaSuccess) {
        if (ts == 4) {
            using F = float;
            F* off = q_is_consistent ? nullptr : (F*)d_off;
            if (off) rc = viprs_b200_q_offset_f32(ld, (F*)d_eta, (F*)d_q, (F)dq_scale, off, st);
            if (rc == 0)
                rc = K > 0 ? viprs_b200_e_step_mixture_f32(ld, K, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta, (F*)d_q, (F*)d_diff,
                                                           (F*)d_lnp, (F*)d_ul, (F*)d_sv, (F*)d_mm, (F)dq_scale, 1, off, st)
                           : viprs_b200_e_step_f32(ld, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta, (F*)d_q, (F*)d_diff, (F*)d_ul,
                                                   (F*)d_sv, (F*)d_mm, (F)dq_scale, 1, off, st);
        } else {
            using F = double;
            F* off = q_is_consistent ? nullptr : (F*)d_off;
            if (off) rc = viprs_b200_q_offset_f64(ld, (F*)d_eta, (F*)d_q, (F)dq_scale, off, st);
            if (rc == 0)
                rc = K > 0 ? viprs_b200_e_step_mixture_f64(ld, K, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta, (F*)d_q, (F*)d_diff,
                                                           (F*)d_lnp, (F*)d_ul, (F*)d_sv, (F*)d_mm, (F)dq_scale, 1, off, st)
                           : viprs_b200_e_step_f64(ld, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta, (F*)d_q, (F*)d_diff, (F*)d_ul,
                                                   (F*)d_sv, (F*)d_mm, (F)dq_scale, 1, off, st);
        }
    }
    auto down = [&](void* dst, const void* src, size_t n) {
        if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st);
    };
    down(var_gamma, d_g, ck); down(var_mu, d_mu, ck); down(eta, d_eta, c1); down(q, d_q, c1); down(eta_diff, d_diff, c1);
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = e2;
    if (rc) return rc;
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}

// one-shot: upload + pack the LD, one sweep, destroy
int host_dropin(int32_t M, int32_t K, const int32_t* lb, const void* indptr, int32_t is64, const void* ld_data,
                int32_t ld_dtype, int32_t float_dtype, const void* std_beta, void* var_gamma, void* var_mu, void* eta,
                void* q, void* eta_diff, const void* log_null_pi, const void* u_logs, const void* shvt,
                const void* mu_mult, double dq_scale) {
    if (float_dtype != VIPRS_B200_F32 && float_dtype != VIPRS_B200_F64) return VIPRS_B200_EINVAL;
    if (!std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !shvt || !mu_mult)
        return VIPRS_B200_EINVAL;
    if (K > 0 && !log_null_pi) return VIPRS_B200_EINVAL;
    viprs_b200_ld_t* ld = nullptr;
    int rc = viprs_b200_ld_create(&ld, M, lb, indptr, is64, ld_data, ld_dtype, VIPRS_B200_MEM_HOST, 0, nullptr);
    if (rc) return rc;
    rc = state_roundtrip(ld, K, float_dtype, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs, shvt,
                         mu_mult, dq_scale, 0, nullptr);
    viprs_b200_ld_destroy(ld);
    return rc;
}

}  // namespace

// cpp_e_step / cpp_e_step_mixture argument lists (e_step_cpp.pyx:91-105, 125-141) with HOST state arrays on a
// device-resident LD matrix: what a per-iteration caller that keeps its state in numpy uses.
extern "C" int viprs_b200_cpp_e_step_resident(const viprs_b200_ld_t* ld, int32_t float_dtype, const void* std_beta,
                                              void* var_gamma, void* var_mu, void* eta, void* q, void* eta_diff,
                                              const void* u_logs, const void* sqrt_half_var_tau, const void* mu_mult,
                                              double dq_scale, int32_t q_is_consistent, void* stream) {
    return state_roundtrip(ld, 0, float_dtype, std_beta, var_gamma, var_mu, eta, q, eta_diff, nullptr, u_logs,
                           sqrt_half_var_tau, mu_mult, dq_scale, q_is_consistent, (cudaStream_t)stream);
}
extern "C" int viprs_b200_cpp_e_step_mixture_resident(const viprs_b200_ld_t* ld, int32_t K, int32_t float_dtype,
                                                      const void* std_beta, void* var_gamma, void* var_mu, void* eta,
                                                      void* q, void* eta_diff, const void* log_null_pi, const void* u_logs,
                                                      const void* sqrt_half_var_tau, const void* mu_mult, double dq_scale,
                                                      int32_t q_is_consistent, void* stream) {
    if (K < 1) return VIPRS_B200_EINVAL;
    return state_roundtrip(ld, K, float_dtype, std_beta, var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs,
                           sqrt_half_var_tau, mu_mult, dq_scale, q_is_consistent, (cudaStream_t)stream);
}

// cpp_e_step (e_step_cpp.pyx:91-122)
extern "C" int viprs_b200_cpp_e_step(int32_t M, const int32_t* ld_left_bound, const void* ld_indptr,
                                     int32_t indptr_is_i64, const void* ld_data, int32_t ld_dtype,
                                     int32_t float_dtype, const void* std_beta, void* var_gamma, void* var_mu,
                                     void* eta, void* q, void* eta_diff, const void* u_logs,
                                     const void* sqrt_half_var_tau, const void* mu_mult, double dq_scale,
                                     int32_t threads, int32_t low_memory) {
    (void)threads; (void)low_memory;   // always the threads=1 order; the layout is detected per row
    return host_dropin(M, 0, ld_left_bound, ld_indptr, indptr_is_i64, ld_data, ld_dtype, float_dtype, std_beta,
                       var_gamma, var_mu, eta, q, eta_diff, nullptr, u_logs, sqrt_half_var_tau, mu_mult, dq_scale);
}

// cpp_e_step_mixture (e_step_cpp.pyx:125-159)
extern "C" int viprs_b200_cpp_e_step_mixture(int32_t M, int32_t K, const int32_t* ld_left_bound, const void* ld_indptr,
                                             int32_t indptr_is_i64, const void* ld_data, int32_t ld_dtype,
                                             int32_t float_dtype, const void* std_beta, void* var_gamma, void* var_mu,
                                             void* eta, void* q, void* eta_diff, const void* log_null_pi,
                                             const void* u_logs, const void* sqrt_half_var_tau, const void* mu_mult,
                                             double dq_scale, int32_t threads, int32_t low_memory) {
    (void)threads; (void)low_memory;
    if (K < 1) return VIPRS_B200_EINVAL;
    return host_dropin(M, K, ld_left_bound, ld_indptr, indptr_is_i64, ld_data, ld_dtype, float_dtype, std_beta,
                       var_gamma, var_mu, eta, q, eta_diff, log_null_pi, u_logs, sqrt_half_var_tau, mu_mult, dq_scale);
}

// cpp_e_step_grid (e_step_cpp.pyx:161-195)
extern "C" int viprs_b200_cpp_e_step_grid(int32_t M, int32_t G, int32_t n_active, const int32_t* active_model_idx,
                                          const int32_t* ld_left_bound, const void* ld_indptr, int32_t indptr_is_i64,
                                          const void* ld_data, int32_t ld_dtype, int32_t float_dtype,
                                          const void* std_beta, void* var_gamma, void* var_mu, void* eta, void* q,
                                          void* eta_diff, const void* u_logs, const void* half_var_tau,
                                          const void* mu_mult, double dq_scale, int32_t threads, int32_t low_memory) {
    (void)threads; (void)low_memory;
    if (float_dtype != VIPRS_B200_F32 && float_dtype != VIPRS_B200_F64) return VIPRS_B200_EINVAL;
    if (G < 1 || n_active < 0 || n_active > G || (n_active > 0 && !active_model_idx)) return VIPRS_B200_EINVAL;
    if (!std_beta || !var_gamma || !var_mu || !eta || !q || !eta_diff || !u_logs || !half_var_tau || !mu_mult)
        return VIPRS_B200_EINVAL;
    viprs_b200_ld_t* ld = nullptr;
    int rc = viprs_b200_ld_create(&ld, M, ld_left_bound, ld_indptr, indptr_is_i64, ld_data, ld_dtype,
                                  VIPRS_B200_MEM_HOST, 0, nullptr);
    if (rc) return rc;
    const size_t ts = float_dtype == VIPRS_B200_F32 ? 4 : 8;
    const size_t n1 = (size_t)M * ts, ng = n1 * (size_t)G, na = ((size_t)(n_active > 0 ? n_active : 1) * 4 + 15) & ~(size_t)15;
    // layout: [gamma][mu][eta][q][diff][ulogs][hvt][mm] (ng each) [beta n1] [active]
    DevBuf buf;
    cudaError_t e = cudaMalloc(&buf.d, 8 * ng + n1 + 16 + na);
    if (e != cudaSuccess) { viprs_b200_ld_destroy(ld); return (int)e; }
    unsigned char* d = buf.d;
    unsigned char *d_g = d, *d_mu = d + ng, *d_eta = d + 2 * ng, *d_q = d + 3 * ng, *d_diff = d + 4 * ng,
                  *d_ul = d + 5 * ng, *d_hv = d + 6 * ng, *d_mm = d + 7 * ng, *d_beta = d + 8 * ng,
                  *d_act = d + 8 * ng + ((n1 + 15) & ~(size_t)15);
    auto up = [&](void* dst, const void* src, size_t n) {
        if (e == cudaSuccess && n > 0) e = cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, 0);
    };
    up(d_g, var_gamma, ng); up(d_mu, var_mu, ng); up(d_eta, eta, ng); up(d_q, q, ng); up(d_diff, eta_diff, ng);
    up(d_ul, u_logs, ng); up(d_hv, half_var_tau, ng); up(d_mm, mu_mult, ng); up(d_beta, std_beta, n1);
    up(d_act, active_model_idx, (size_t)n_active * 4);
    if (e == cudaSuccess) {
        if (ts == 4) {
            using F = float;
            rc = viprs_b200_e_step_grid_f32(ld, G, n_active, (const int32_t*)d_act, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta,
                                            (F*)d_q, (F*)d_diff, (F*)d_ul, (F*)d_hv, (F*)d_mm, (F)dq_scale, nullptr);
        } else {
            using F = double;
            rc = viprs_b200_e_step_grid_f64(ld, G, n_active, (const int32_t*)d_act, (F*)d_beta, (F*)d_g, (F*)d_mu, (F*)d_eta,
                                            (F*)d_q, (F*)d_diff, (F*)d_ul, (F*)d_hv, (F*)d_mm, (F)dq_scale, nullptr);
        }
    }
    auto down = [&](void* dst, const void* src, size_t n) {
        if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, 0);
    };
    down(var_gamma, d_g, ng); down(var_mu, d_mu, ng); down(eta, d_eta, ng); down(q, d_q, ng); down(eta_diff, d_diff, ng);
    cudaError_t e2 = cudaStreamSynchronize(0);
    if (e == cudaSuccess) e = e2;
    viprs_b200_ld_destroy(ld);
    if (rc) return rc;
    return e == cudaSuccess ? VIPRS_B200_OK : (int)e;
}
