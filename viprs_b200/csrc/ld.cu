// viprs_b200 -- device-resident LD matrix: host-side planning (LD blocks, row panels) and the
// re-layout kernel that turns magenpy's CSR-without-column-indices (consumer: e_step.hpp:389-392,
// producer call site: VIPRS.py:167-172) into the 16-byte-aligned row layout the sweep streams
// with 1-D TMA bulk copies.  See DESIGN.md "Data layout in HBM".
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <numeric>
#include <vector>

#include <cstdlib>

#include "common.cuh"
#include "ld.h"
#include "sweep.cuh"
#include "sweep_fast.cuh"

namespace vb {

// integer codes are stored biased (see common.cuh): int8 ^ 0x80, int16 ^ 0x8000
__device__ __forceinline__ int8_t bias(int8_t v) { return (int8_t)(v ^ (int8_t)0x80); }
__device__ __forceinline__ int16_t bias(int16_t v) { return (int16_t)(v ^ (int16_t)0x8000); }
__device__ __forceinline__ float bias(float v) { return v; }
__device__ __forceinline__ double bias(double v) { return v; }

template <typename U>
__global__ void pack_rows_kernel(int M, const U* __restrict__ src, const int64_t* __restrict__ src_off,
                                 const int32_t* __restrict__ cs, const int32_t* __restrict__ ce,
                                 const int64_t* __restrict__ prow, const int32_t* __restrict__ pcs,
                                 U* __restrict__ dst) {
    // one warp per row; 8 rows per CTA
    const int row = blockIdx.x * (blockDim.x / WARP) + (threadIdx.x / WARP);
    if (row >= M) return;
    const int lane = threadIdx.x % WARP;
    const int64_t p0 = prow[row];
    const int plen = (int)(prow[row + 1] - p0);
    const int c0 = pcs[row], a = cs[row], b = ce[row];
    const int64_t so = src_off[row];
    for (int e = lane; e < plen; e += WARP) {
        const int col = c0 + e;
        U v = U(0);
        if (col >= a && col < b) v = src[so + (col - a)];
        dst[p0 + e] = bias(v);
    }
}

// dense symmetric copy of one LD block row: out[j][k] = R[min(j,k)][max(j,k)] (0 on the diagonal / outside the
// stored run / in the padding), read from the packed upper-triangular rows.  One CTA per row.
template <typename U>
__global__ void mirror_dense_kernel(int n_blocks, const int32_t* __restrict__ blk_row, const int64_t* __restrict__ dblk_off,
                                    const U* __restrict__ packed, const int64_t* __restrict__ prow,
                                    const int32_t* __restrict__ pcs, unsigned char* __restrict__ dense) {
    const int row = blockIdx.x;
    int lo = 0, hi = n_blocks;                     // block of this row: blk_row[lo] <= row < blk_row[lo+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blk_row[mid] <= row) lo = mid; else hi = mid;
    }
    const int r0 = blk_row[lo], B = blk_row[lo + 1] - r0, Bp = (B + 15) & ~15;
    const int j = row - r0;
    U* out = reinterpret_cast<U*>(dense + dblk_off[lo]) + (size_t)j * Bp;
    const U zero = bias(U(0));
    for (int k = threadIdx.x; k < Bp; k += blockDim.x) {
        U v = zero;
        if (k < B && k != j) {
            const int a = min(j, k), b = max(j, k);
            const int64_t p0 = prow[r0 + a];
            const int64_t idx = (int64_t)(r0 + b) - pcs[r0 + a];
            if (idx >= 0 && idx < prow[r0 + a + 1] - p0) v = packed[p0 + idx];
        }
        out[k] = v;
    }
}

static int elem_size(int dt) {
    switch (dt) {
        case VIPRS_B200_I8: return 1;
        case VIPRS_B200_I16: return 2;
        case VIPRS_B200_F32: return 4;
        case VIPRS_B200_F64: return 8;
    }
    return 0;
}

}  // namespace vb

#define CUDA_TRY(x)                          \
    do {                                     \
        cudaError_t _e = (x);                \
        if (_e != cudaSuccess) {             \
            rc = (int)_e;                    \
            goto fail;                       \
        }                                    \
    } while (0)

extern "C" int viprs_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" const char* viprs_b200_version(void) { return "viprs_b200 0.1.0 (sm_100a)"; }

extern "C" const char* viprs_b200_strerror(int code) {
    switch (code) {
        case VIPRS_B200_OK: return "ok";
        case VIPRS_B200_EINVAL: return "invalid argument";
        case VIPRS_B200_ELAYOUT: return "inconsistent LD layout (row run leaves [0, M) or negative length)";
        case VIPRS_B200_EBLOCK_TOO_LARGE:
            return "grid sweep: an LD block is larger than 4096 SNPs (only the single-model and mixture sweeps tile larger blocks)";
        case VIPRS_B200_ENOMEM: return "out of memory";
        case VIPRS_B200_ENODEVICE: return "no CUDA device (there is no CPU fallback)";
        case VIPRS_B200_EUNSUPPORTED: return "unsupported dtype combination";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

namespace vb {

// bytes of dynamic shared memory the sweep needs besides the TMA ring
static int state_bytes(int max_block, int tsize) {
    return (int)make_layout(state_pad(max_block), tsize, 0, 0).total;
}

RingGeometry ring_geometry(const viprs_b200_ld* ld, int tsize) {
    RingGeometry g{0, 0, 0};
    const int st = state_bytes(ld->max_block, tsize);
    const int cap1 = ld->smem_optin;
    int nst2 = (kSmemTwoPerSM - st) / ld->stage_bytes;
    int nst1 = (cap1 - st) / ld->stage_bytes;
    if (nst2 > NST_MAX) nst2 = NST_MAX;
    if (nst1 > NST_MAX) nst1 = NST_MAX;
    if (nst2 >= 4) { g.nst = nst2; g.ctas_per_sm = 2; }
    else if (nst1 >= 3) { g.nst = nst1; g.ctas_per_sm = 1; }
    else return g;
    g.smem_bytes = (int)make_layout(state_pad(ld->max_block), tsize, ld->stage_bytes, g.nst).total;
    return g;
}

RingGeometry fast_ring_geometry(const viprs_b200_ld* ld) {
    RingGeometry g{0, 0, 0};
    if (ld->max_block > FAST_MAX_BLOCK) return g;
    const int st = (int)make_fast_layout(0, 0).total;
    int nst = (kSmemTwoPerSM - st) / ld->stage_bytes;
    if (nst > NST_MAX) nst = NST_MAX;
    if (nst < 3) return g;
    g.nst = nst; g.ctas_per_sm = 2;
    g.smem_bytes = (int)make_fast_layout(ld->stage_bytes, nst).total;
    return g;
}

// pick the TMA stage size: four stages next to the float32 block state if two CTAs can share an SM,
// otherwise four stages of a single CTA per SM; never smaller than the longest packed row.
static int choose_stage_bytes(int max_block, int max_row_bytes, int smem_optin, int requested) {
    const char* env = getenv("VIPRS_B200_STAGE_BYTES");
    if (requested <= 0 && env) requested = atoi(env);
    int sb;
    if (requested > 0) {
        sb = requested;
    } else {
        const int st = max_block <= FAST_MAX_BLOCK ? (int)make_fast_layout(0, 0).total : state_bytes(max_block, 4);
        sb = (kSmemTwoPerSM - st) / 4;
        // long rows: 8-row panels in a 3-deep ring beat 4-row panels in a 4-deep one (per-panel hand-off costs
        // dominate the first quarter of a 4096-SNP block; measured 1.102 -> 1.069 ms on the C2 workload)
        if (max_block <= FAST_MAX_BLOCK && (int64_t)max_row_bytes * 8 > sb) sb = ((kSmemTwoPerSM - st) / 3) & ~127;
        if (sb < max_row_bytes || sb < 2048) sb = (smem_optin - st) / 4;
        if (sb > 48 * 1024) sb = 48 * 1024;
    }
    sb &= ~127;
    if (sb < max_row_bytes) sb = (max_row_bytes + 127) & ~127;
    return sb;
}

}  // namespace vb

extern "C" int viprs_b200_ld_create(viprs_b200_ld_t** out, int32_t M, const int32_t* left_bound,
                                    const void* indptr, int32_t indptr_is_i64, const void* ld_data,
                                    int32_t ld_dtype, int32_t mem_kind, int32_t stage_bytes, void* stream_) {
    if (!out || M <= 0 || !left_bound || !indptr) return VIPRS_B200_EINVAL;
    const int esize = vb::elem_size(ld_dtype);
    if (esize == 0) return VIPRS_B200_EINVAL;
    if (viprs_b200_device_count() <= 0) return VIPRS_B200_ENODEVICE;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int epv = 16 / esize;

    int rc = VIPRS_B200_OK;
    viprs_b200_ld* h = new (std::nothrow) viprs_b200_ld();
    if (!h) return VIPRS_B200_ENOMEM;
    void* d_raw = nullptr;        // staging copy of the caller's ld_data when it lives on the host
    int64_t* d_src_off = nullptr;
    int32_t *d_cs = nullptr, *d_ce = nullptr;

    // ---- host copies of the index arrays -------------------------------------------------
    std::vector<int32_t> lb(M);
    std::vector<int64_t> ip((size_t)M + 1);
    {
        const size_t ipb = (size_t)(M + 1) * (indptr_is_i64 ? 8 : 4);
        std::vector<unsigned char> tmp(ipb);
        if (mem_kind == VIPRS_B200_MEM_DEVICE) {
            CUDA_TRY(cudaMemcpyAsync(lb.data(), left_bound, (size_t)M * 4, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(tmp.data(), indptr, ipb, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
        } else {
            std::memcpy(lb.data(), left_bound, (size_t)M * 4);
            std::memcpy(tmp.data(), indptr, ipb);
        }
        if (indptr_is_i64) {
            std::memcpy(ip.data(), tmp.data(), ipb);
        } else {
            const int32_t* p32 = reinterpret_cast<const int32_t*>(tmp.data());
            for (int j = 0; j <= M; ++j) ip[j] = p32[j];
        }
    }
    {
        // ---- per-row runs (strictly upper part), LD blocks ----------------------------------
        std::vector<int32_t> cs(M), ce(M), pcs(M);
        std::vector<int64_t> src_off(M), prow((size_t)M + 1);
        std::vector<int32_t> ldblk_row;
        int64_t nnz = 0;
        int32_t runmax = 0;
        for (int j = 0; j < M; ++j) {
            const int64_t len = ip[j + 1] - ip[j];
            if (len < 0 || lb[j] < 0 || (int64_t)lb[j] + len > M) { rc = VIPRS_B200_ELAYOUT; goto fail; }
            int64_t a = lb[j], b = (int64_t)lb[j] + len;
            int64_t skip = std::max<int64_t>(0, (int64_t)j + 1 - a);   // symmetric layout: drop columns <= j
            if (skip > len) skip = len;
            a += skip;
            if (a >= b) { a = b = j + 1; }
            cs[j] = (int32_t)a; ce[j] = (int32_t)b; src_off[j] = ip[j] + skip;
            nnz += b - a;
            if (j == 0 || runmax <= j) ldblk_row.push_back(j);   // no earlier row reaches column j
            runmax = std::max<int32_t>(runmax, (int32_t)b);
        }
        ldblk_row.push_back(M);
        const int nlb = (int)ldblk_row.size() - 1;

        // ---- sweep units: an LD block, or kTileRows-row tiles of a block larger than kTileLimit -----------
        // (VIPRS_B200_TILE_LIMIT / VIPRS_B200_TILE_ROWS: timing experiments with the blocked decomposition on blocks that
        // would fit one unit)
        int tile_limit = vb::kTileLimit, tile_rows = vb::kTileRows;
        if (const char* e = getenv("VIPRS_B200_TILE_LIMIT")) { const int v = atoi(e); if (v >= 64 && v <= vb::kTileLimit) tile_limit = v; }
        if (const char* e = getenv("VIPRS_B200_TILE_ROWS")) { const int v = atoi(e); if (v >= 64 && v <= 2048 && v % 16 == 0) tile_rows = v; }
        std::vector<int32_t> blk_row, unit_phase;
        int32_t max_ld_block = 0, n_phases = 1;
        for (int b = 0; b < nlb; ++b) {
            const int r0 = ldblk_row[b], r1 = ldblk_row[b + 1], B = r1 - r0;
            max_ld_block = std::max(max_ld_block, B);
            if (B <= tile_limit) {
                blk_row.push_back(r0); unit_phase.push_back(0);
            } else {
                int ph = 0;
                for (int t = r0; t < r1; t += tile_rows) { blk_row.push_back(t); unit_phase.push_back(ph++); }
                n_phases = std::max(n_phases, ph);
            }
        }
        blk_row.push_back(M);
        const int nb = (int)blk_row.size() - 1;

        // ---- aligned packed rows (columns inside the unit) and ext rows (columns beyond it) ----------------
        std::vector<int32_t> ce_in(M), ecs(M);
        std::vector<int64_t> erow((size_t)M + 1);
        std::vector<int64_t> blk_cost(nb);
        std::vector<int4> items_diag(nb);
        std::vector<int32_t> unit_ext_end(nb, 0);
        int64_t off = 0, eoff = 0;
        int32_t max_block = 0, max_row_bytes = 0;
        for (int b = 0; b < nb; ++b) {
            const int r0 = blk_row[b], r1 = blk_row[b + 1];
            max_block = std::max(max_block, r1 - r0);
            int64_t cost = 0;
            for (int j = r0; j < r1; ++j) {
                const int32_t cin = std::min(ce[j], r1);             // end of the in-unit part
                ce_in[j] = cin;
                int32_t a = r0 + ((cs[j] - r0) / epv) * epv;
                int32_t e = r0 + ((cin - r0 + epv - 1) / epv) * epv;
                if (cin <= cs[j]) { a = r0 + ((std::min(cs[j], r1) - r0) / epv) * epv; e = a; }
                pcs[j] = a; prow[j] = off;
                max_row_bytes = std::max<int32_t>(max_row_bytes, (e - a) * esize);
                off += (e - a); cost += (int64_t)(e - a) * esize + 256;
                // ext part: columns [max(cs, r1), ce), aligned to epv relative to r1
                const int32_t xs = std::max(cs[j], r1);
                int32_t xa = r1, xe = r1;
                if (ce[j] > xs) {
                    xa = r1 + ((xs - r1) / epv) * epv;
                    xe = r1 + ((ce[j] - r1 + epv - 1) / epv) * epv;
                    unit_ext_end[b] = std::max(unit_ext_end[b], ce[j]);
                }
                ecs[j] = xa; erow[j] = eoff;
                eoff += (xe - xa);
            }
            blk_cost[b] = cost;
            items_diag[b] = make_int4(r0, r1, r0, r1);
        }
        prow[M] = off;
        erow[M] = eoff;

        CUDA_TRY(cudaGetDevice(&h->device));
        CUDA_TRY(cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
        CUDA_TRY(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device));
        h->M = M; h->ld_dtype = ld_dtype; h->esize = esize; h->epv = epv; h->nnz = nnz;
        h->packed_elems = off; h->ext_elems = eoff; h->n_blocks = nb; h->max_block = max_block; h->max_row_bytes = max_row_bytes;
        h->n_ld_blocks = nlb; h->max_ld_block = max_ld_block; h->n_phases = n_phases;
        h->h_blk_row = blk_row;
        h->h_ldblk_row = ldblk_row;
        stage_bytes = vb::choose_stage_bytes(max_block, max_row_bytes, h->smem_optin, stage_bytes);
        h->stage_bytes = stage_bytes;
        if (vb::ring_geometry(h, 4).nst == 0 && vb::fast_ring_geometry(h).nst == 0) { rc = VIPRS_B200_EBLOCK_TOO_LARGE; goto fail; }

        // ---- row panels (one TMA bulk copy each) and the chain warp's axpy prerequisites --------
        std::vector<int32_t> blk_panel(nb + 1), panel_row, panel_need;
        for (int b = 0; b < nb; ++b) {
            const int r0 = blk_row[b], r1 = blk_row[b + 1];
            const int first_panel = (int)panel_row.size();
            blk_panel[b] = first_panel;
            // greedy cut: as many rows as fit one stage (<= PMAX), rounded down to a multiple of 4 because the
            // bulk warps work on groups of 4 rows (a 5-row panel would cost them as much as an 8-row one)
            int j = r0;
            while (j < r1) {
                int64_t pbytes = 0; int prows = 0;
                while (j + prows < r1 && prows < vb::PMAX) {
                    const int64_t rb = (prow[j + prows + 1] - prow[j + prows]) * esize;
                    if (prows > 0 && pbytes + rb > stage_bytes) break;
                    pbytes += rb; ++prows;
                }
                if (prows > 4 && j + prows < r1) prows &= ~3;
                panel_row.push_back(j);
                j += prows;
            }
            // panel u = rows [ps, pe): its columns need the bulk axpy of every row <= pe-1-WIN, i.e. of
            // every panel up to the one holding that row (which always ends before ps because PMAX <= WIN/2)
            const int np_b = (int)panel_row.size() - first_panel;
            int pc = 0;
            for (int u = 0; u < np_b; ++u) {
                const int pe = (u + 1 < np_b) ? panel_row[first_panel + u + 1] : r1;
                const int need_row = pe - 1 - vb::WIN;          // global row index, may be < r0
                int need = 0;
                if (need_row >= r0) {
                    while (pc + 1 < np_b && panel_row[first_panel + pc + 1] <= need_row) ++pc;
                    need = pc + 1;
                }
                panel_need.push_back(need);
            }
        }
        blk_panel[nb] = (int32_t)panel_row.size();
        const int np = (int)panel_row.size();
        panel_row.push_back(M);
        h->n_panels = np;

        // launch p sweeps the units of phase p, most expensive first (LPT schedule for ragged sizes)
        std::vector<int32_t> order(nb);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (unit_phase[a] != unit_phase[b]) return unit_phase[a] < unit_phase[b];
            return blk_cost[a] > blk_cost[b];
        });
        h->h_phase_ptr.assign(n_phases + 1, 0);
        for (int b = 0; b < nb; ++b) h->h_phase_ptr[unit_phase[b] + 1]++;
        for (int p = 0; p < n_phases; ++p) h->h_phase_ptr[p + 1] += h->h_phase_ptr[p];
        // ext rectangles grouped by the phase of the unit that owns their rows
        std::vector<int4> items_ext;
        h->h_ext_phase_ptr.assign(n_phases + 1, 0);
        h->h_items_cols.assign(n_phases + 1, 0);
        for (int p = 0; p < n_phases; ++p) {
            for (int b = 0; b < nb; ++b) {
                if (unit_phase[b] != p || unit_ext_end[b] <= blk_row[b + 1]) continue;
                items_ext.push_back(make_int4(blk_row[b], blk_row[b + 1], blk_row[b + 1], unit_ext_end[b]));
                h->h_items_cols[p] = std::max(h->h_items_cols[p], unit_ext_end[b] - blk_row[b + 1]);
            }
            h->h_ext_phase_ptr[p + 1] = (int32_t)items_ext.size();
        }
        h->n_items_ext = (int32_t)items_ext.size();
        h->h_items_cols[n_phases] = max_block;
        // row-dot items (update_q_factor / backward-external products): <= 64-row chunks of every unit
        std::vector<int4> items_bwd, items_bwd_ext;
        for (int b = 0; b < nb; ++b) {
            const int r0 = blk_row[b], r1 = blk_row[b + 1];
            for (int j = r0; j < r1; j += 64) {
                const int je = std::min(j + 64, r1);
                if (prow[je] > prow[j]) items_bwd.push_back(make_int4(j, je, r0, r1));
                if (erow[je] > erow[j]) items_bwd_ext.push_back(make_int4(j, je, r1, unit_ext_end[b]));
            }
        }
        // the ext row-dot items grouped by the phase of the unit that owns their rows: the dots of tile p + 1 only have to
        // be there before launch p + 1, so they can run next to the sweep of tile p (launch.cuh: launch_sweep)
        {
            std::vector<int32_t> item_phase(items_bwd_ext.size());
            size_t k = 0;
            for (int b = 0; b < nb; ++b) {
                const int r0 = blk_row[b], r1 = blk_row[b + 1];
                for (int j = r0; j < r1; j += 64) {
                    const int je = std::min(j + 64, r1);
                    if (erow[je] > erow[j]) item_phase[k++] = unit_phase[b];
                }
            }
            std::vector<int32_t> perm(items_bwd_ext.size());
            std::iota(perm.begin(), perm.end(), 0);
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int c) { return item_phase[a] < item_phase[c]; });
            std::vector<int4> sorted(items_bwd_ext.size());
            h->h_bwd_ext_phase_ptr.assign(n_phases + 1, 0);
            for (size_t i = 0; i < perm.size(); ++i) { sorted[i] = items_bwd_ext[perm[i]]; h->h_bwd_ext_phase_ptr[item_phase[perm[i]] + 1]++; }
            for (int p = 0; p < n_phases; ++p) h->h_bwd_ext_phase_ptr[p + 1] += h->h_bwd_ext_phase_ptr[p];
            items_bwd_ext.swap(sorted);
        }
        h->n_items_bwd = (int32_t)items_bwd.size();
        h->n_items_bwd_ext = (int32_t)items_bwd_ext.size();
        // row chunks of about equal sweep cost (single-phase LD with enough units)
        std::vector<int32_t> chunk_order;
        int n_chunks_want = vb::kChunks;
        if (const char* e = getenv("VIPRS_B200_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= 8) n_chunks_want = v; }
        if (n_phases == 1 && nb >= 2 * n_chunks_want && n_chunks_want > 1) {
            double total = 0.0;
            for (int b = 0; b < nb; ++b) total += (double)blk_cost[b];
            h->h_chunk_unit.assign(1, 0);
            double run = 0.0;
            for (int b = 0; b < nb; ++b) {
                run += (double)blk_cost[b];
                const int c = (int)h->h_chunk_unit.size();
                if (c < n_chunks_want && run >= total * c / n_chunks_want && b + 1 < nb) h->h_chunk_unit.push_back(b + 1);
            }
            h->h_chunk_unit.push_back(nb);
            h->n_chunks = (int32_t)h->h_chunk_unit.size() - 1;
            chunk_order.resize(nb);
            std::iota(chunk_order.begin(), chunk_order.end(), 0);
            size_t it = 0;
            for (int c = 0; c < h->n_chunks; ++c) {
                const int u0 = h->h_chunk_unit[c], u1 = h->h_chunk_unit[c + 1];
                std::stable_sort(chunk_order.begin() + u0, chunk_order.begin() + u1, [&](int a, int b) { return blk_cost[a] > blk_cost[b]; });
                h->h_chunk_row.push_back(blk_row[u0]);
                while (it < items_bwd.size() && items_bwd[it].x < blk_row[u0]) ++it;
                h->h_chunk_item.push_back((int32_t)it);
            }
            h->h_chunk_row.push_back(M);
            h->h_chunk_item.push_back((int32_t)items_bwd.size());
        }

        // ---- device arrays ------------------------------------------------------------------
        CUDA_TRY(cudaMalloc(&h->d_packed, (size_t)std::max<int64_t>(off, 16) * esize + 64));
        CUDA_TRY(cudaMalloc(&h->d_prow, sizeof(int64_t) * ((size_t)M + 1)));
        CUDA_TRY(cudaMalloc(&h->d_pcs, sizeof(int32_t) * (size_t)M));
        CUDA_TRY(cudaMalloc(&h->d_blk_row, sizeof(int32_t) * (nb + 1)));
        CUDA_TRY(cudaMalloc(&h->d_blk_panel, sizeof(int32_t) * (nb + 1)));
        CUDA_TRY(cudaMalloc(&h->d_panel_row, sizeof(int32_t) * (np + 1)));
        CUDA_TRY(cudaMalloc(&h->d_panel_need, sizeof(int32_t) * std::max(np, 1)));
        CUDA_TRY(cudaMalloc(&h->d_blk_order, sizeof(int32_t) * nb));
        CUDA_TRY(cudaMalloc(&h->d_items_diag, sizeof(int4) * nb));
        CUDA_TRY(cudaMalloc(&h->d_unit_partial, sizeof(double) * VIPRS_B200_NSUMS * (size_t)nb));
        CUDA_TRY(cudaMemsetAsync(h->d_unit_partial, 0, sizeof(double) * VIPRS_B200_NSUMS * (size_t)nb, stream));
        CUDA_TRY(cudaMalloc(&h->d_items_bwd, sizeof(int4) * std::max<size_t>(items_bwd.size(), 1)));
        CUDA_TRY(cudaMalloc(&h->d_items_bwd_ext, sizeof(int4) * std::max<size_t>(items_bwd_ext.size(), 1)));
        CUDA_TRY(cudaMemcpyAsync(h->d_items_bwd, items_bwd.data(), sizeof(int4) * items_bwd.size(), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_items_bwd_ext, items_bwd_ext.data(), sizeof(int4) * items_bwd_ext.size(), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMalloc(&d_src_off, sizeof(int64_t) * (size_t)M));
        CUDA_TRY(cudaMalloc(&d_cs, sizeof(int32_t) * (size_t)M));
        CUDA_TRY(cudaMalloc(&d_ce, sizeof(int32_t) * (size_t)M));
        CUDA_TRY(cudaMemcpyAsync(h->d_prow, prow.data(), sizeof(int64_t) * ((size_t)M + 1), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_pcs, pcs.data(), sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_blk_row, blk_row.data(), sizeof(int32_t) * (nb + 1), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_blk_panel, blk_panel.data(), sizeof(int32_t) * (nb + 1), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_panel_row, panel_row.data(), sizeof(int32_t) * (np + 1), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_panel_need, panel_need.data(), sizeof(int32_t) * np, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_blk_order, order.data(), sizeof(int32_t) * nb, cudaMemcpyHostToDevice, stream));
        if (h->n_chunks > 0) {
            CUDA_TRY(cudaMalloc(&h->d_chunk_order, sizeof(int32_t) * nb));
            CUDA_TRY(cudaMemcpyAsync(h->d_chunk_order, chunk_order.data(), sizeof(int32_t) * nb, cudaMemcpyHostToDevice, stream));
        }
        CUDA_TRY(cudaMemcpyAsync(h->d_items_diag, items_diag.data(), sizeof(int4) * nb, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_src_off, src_off.data(), sizeof(int64_t) * (size_t)M, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_cs, cs.data(), sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(d_ce, ce_in.data(), sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, stream));

        const void* d_src = ld_data;
        const int64_t total_in = ip[M];
        if (mem_kind != VIPRS_B200_MEM_DEVICE) {
            CUDA_TRY(cudaMalloc(&d_raw, (size_t)std::max<int64_t>(total_in, 1) * esize));
            if (total_in > 0) {
                if (!ld_data) { rc = VIPRS_B200_EINVAL; goto fail; }
                CUDA_TRY(cudaMemcpyAsync(d_raw, ld_data, (size_t)total_in * esize, cudaMemcpyHostToDevice, stream));
            }
            d_src = d_raw;
        }
        auto pack = [&](const int32_t* a_cs, const int32_t* a_ce, const int64_t* a_prow, const int32_t* a_pcs, void* dst) {
            const int wpb = 8;
            dim3 grid((M + wpb - 1) / wpb), block(wpb * vb::WARP);
            switch (ld_dtype) {
                case VIPRS_B200_I8:
                    vb::pack_rows_kernel<int8_t><<<grid, block, 0, stream>>>(M, (const int8_t*)d_src, d_src_off, a_cs, a_ce, a_prow, a_pcs, (int8_t*)dst);
                    break;
                case VIPRS_B200_I16:
                    vb::pack_rows_kernel<int16_t><<<grid, block, 0, stream>>>(M, (const int16_t*)d_src, d_src_off, a_cs, a_ce, a_prow, a_pcs, (int16_t*)dst);
                    break;
                case VIPRS_B200_F32:
                    vb::pack_rows_kernel<float><<<grid, block, 0, stream>>>(M, (const float*)d_src, d_src_off, a_cs, a_ce, a_prow, a_pcs, (float*)dst);
                    break;
                default:
                    vb::pack_rows_kernel<double><<<grid, block, 0, stream>>>(M, (const double*)d_src, d_src_off, a_cs, a_ce, a_prow, a_pcs, (double*)dst);
                    break;
            }
            return cudaGetLastError();
        };
        CUDA_TRY(pack(d_cs, d_ce, h->d_prow, h->d_pcs, h->d_packed));
        if (eoff > 0) {
            // the ext rows: same packing kernel, source columns [cs, ce) clipped to [unit end, ce)
            CUDA_TRY(cudaMalloc(&h->d_ext, (size_t)eoff * esize + 64));
            CUDA_TRY(cudaMalloc(&h->d_erow, sizeof(int64_t) * ((size_t)M + 1)));
            CUDA_TRY(cudaMalloc(&h->d_ecs, sizeof(int32_t) * (size_t)M));
            CUDA_TRY(cudaMalloc(&h->d_items_ext, sizeof(int4) * std::max<size_t>(items_ext.size(), 1)));
            CUDA_TRY(cudaMalloc(&h->d_fext, 8 * (size_t)M));
            CUDA_TRY(cudaMalloc(&h->d_bext, 8 * (size_t)M));
            CUDA_TRY(cudaMemcpyAsync(h->d_erow, erow.data(), sizeof(int64_t) * ((size_t)M + 1), cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(h->d_ecs, ecs.data(), sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(h->d_items_ext, items_ext.data(), sizeof(int4) * items_ext.size(), cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));      // d_cs / d_ce are re-used below
            // pack_rows reads source element (col - cs[row]) + src_off[row] for cs <= col < ce: keep cs, raise nothing;
            // the clip to [unit end, ce) comes from the destination range [ecs, ecs + len) and the test col >= xs
            std::vector<int32_t> xs(M);
            for (int b = 0; b < nb; ++b)
                for (int j = blk_row[b]; j < blk_row[b + 1]; ++j) xs[j] = std::max(cs[j], blk_row[b + 1]);
            // source offset of column xs[j]
            std::vector<int64_t> xoff(M);
            for (int j = 0; j < M; ++j) xoff[j] = src_off[j] + (xs[j] - cs[j]);
            CUDA_TRY(cudaMemcpyAsync(d_src_off, xoff.data(), sizeof(int64_t) * (size_t)M, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(d_cs, xs.data(), sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(d_ce, ce.data(), sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(pack(d_cs, d_ce, h->d_erow, h->d_ecs, h->d_ext));
            CUDA_TRY(cudaStreamSynchronize(stream));
        }
        CUDA_TRY(cudaStreamSynchronize(stream));   // host vectors above are about to go out of scope
    }
    cudaFree(d_raw); cudaFree(d_src_off); cudaFree(d_cs); cudaFree(d_ce);
    *out = h;
    return VIPRS_B200_OK;

fail:
    cudaGetLastError();
    cudaFree(d_raw); cudaFree(d_src_off); cudaFree(d_cs); cudaFree(d_ce);
    viprs_b200_ld_destroy(h);
    return rc;
}

int vb::ensure_dense(const viprs_b200_ld* h, cudaStream_t stream) {
    if (h->d_dense) return VIPRS_B200_OK;
    if (h->n_phases > 1 || h->ext_elems > 0) return VIPRS_B200_EBLOCK_TOO_LARGE;    // the grid sweep does not tile
    const int nb = h->n_blocks;
    std::vector<int64_t> off(nb);
    int64_t o = 0;
    for (int b = 0; b < nb; ++b) {
        const int64_t B = h->h_blk_row[b + 1] - h->h_blk_row[b], Bp = (B + 15) & ~15LL;
        off[b] = o;
        o += (B * Bp * h->esize + 127) & ~127LL;
    }
    void* dd = nullptr;
    int64_t* doff = nullptr;
    cudaError_t e = cudaMalloc(&dd, (size_t)o + 1024);     // slack: the chain warp's window loads may run past a row
    if (e == cudaSuccess) e = cudaMalloc(&doff, sizeof(int64_t) * nb);
    if (e == cudaSuccess) e = cudaMemcpyAsync(doff, off.data(), sizeof(int64_t) * nb, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dd, 0, (size_t)o + 1024, stream);
    if (e == cudaSuccess) {
        dim3 grid(h->M), block(128);
        unsigned char* d8 = reinterpret_cast<unsigned char*>(dd);
        switch (h->ld_dtype) {
            case VIPRS_B200_I8:
                vb::mirror_dense_kernel<int8_t><<<grid, block, 0, stream>>>(nb, h->d_blk_row, doff, (const int8_t*)h->d_packed, h->d_prow, h->d_pcs, d8);
                break;
            case VIPRS_B200_I16:
                vb::mirror_dense_kernel<int16_t><<<grid, block, 0, stream>>>(nb, h->d_blk_row, doff, (const int16_t*)h->d_packed, h->d_prow, h->d_pcs, d8);
                break;
            case VIPRS_B200_F32:
                vb::mirror_dense_kernel<float><<<grid, block, 0, stream>>>(nb, h->d_blk_row, doff, (const float*)h->d_packed, h->d_prow, h->d_pcs, d8);
                break;
            default:
                vb::mirror_dense_kernel<double><<<grid, block, 0, stream>>>(nb, h->d_blk_row, doff, (const double*)h->d_packed, h->d_prow, h->d_pcs, d8);
                break;
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);       // `off` goes out of scope
    if (e != cudaSuccess) {
        cudaFree(dd); cudaFree(doff);
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? VIPRS_B200_ENOMEM : (int)e;
    }
    h->d_dense = dd; h->d_dblk_off = doff; h->dense_bytes = o;
    return VIPRS_B200_OK;
}

extern "C" int viprs_b200_ld_destroy(viprs_b200_ld_t* h) {
    if (!h) return VIPRS_B200_OK;
    cudaFree(h->d_dense); cudaFree(h->d_dblk_off);
    cudaFree(h->d_ext); cudaFree(h->d_erow); cudaFree(h->d_ecs); cudaFree(h->d_items_diag); cudaFree(h->d_items_ext); cudaFree(h->d_items_bwd); cudaFree(h->d_items_bwd_ext);
    cudaFree(h->d_fext); cudaFree(h->d_bext); cudaFree(h->d_host_ws); cudaFree(h->d_unit_partial);
    cudaFree(h->d_packed); cudaFree(h->d_prow); cudaFree(h->d_pcs); cudaFree(h->d_blk_row);
    if (h->side_stream) { cudaStreamDestroy(h->side_stream); for (cudaEvent_t e : h->side_events) cudaEventDestroy(e); }
    cudaFree(h->d_blk_panel); cudaFree(h->d_panel_row); cudaFree(h->d_panel_need); cudaFree(h->d_blk_order); cudaFree(h->d_chunk_order);
    delete h;
    return VIPRS_B200_OK;
}

extern "C" int viprs_b200_ld_block_rows(const viprs_b200_ld_t* h, int32_t* out_host) {
    if (!h || !out_host) return VIPRS_B200_EINVAL;
    std::memcpy(out_host, h->h_ldblk_row.data(), sizeof(int32_t) * h->h_ldblk_row.size());
    return VIPRS_B200_OK;
}
