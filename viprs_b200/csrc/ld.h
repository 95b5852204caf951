// Internal definition of the opaque viprs_b200_ld handle (device-resident LD matrix).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

struct viprs_b200_ld {
    int32_t M = 0;
    int32_t ld_dtype = 0;
    int32_t esize = 0;          // bytes per LD element
    int32_t epv = 0;            // elements per 16-byte vector
    int64_t nnz = 0;            // strictly-upper stored entries kept (algorithmic elements)
    int64_t packed_elems = 0;   // elements in the aligned device layout
    int32_t n_blocks = 0;       // sweep units ("tiles"): one per LD block, or several for an LD block > kTileLimit rows
    int32_t n_ld_blocks = 0;    // independent LD blocks found (maximal row ranges no earlier row reaches into)
    int32_t n_phases = 1;       // max tiles per LD block: tile p of every block is swept in launch p (see below)
    int32_t max_block = 0;      // rows of the largest sweep unit
    int32_t max_ld_block = 0;   // rows of the largest LD block
    int32_t max_row_bytes = 0;  // longest packed row
    int32_t n_panels = 0;
    int32_t stage_bytes = 0;
    int device = 0;
    int smem_optin = 0;         // cudaDevAttrMaxSharedMemoryPerBlockOptin
    int n_sm = 0;               // cudaDevAttrMultiProcessorCount

    // device arrays
    void* d_packed = nullptr;      // [packed_elems] LD entries (biased integer codes), rows 16B-aligned
    int64_t* d_prow = nullptr;     // [M+1] element offset of each packed row (multiple of epv)
    int32_t* d_pcs = nullptr;      // [M]   first column (global index) of the packed row, aligned
                                   //       to epv relative to the block start
    int32_t* d_blk_row = nullptr;  // [n_blocks+1] first row of every LD block
    int32_t* d_blk_panel = nullptr;// [n_blocks+1] first panel of every LD block
    int32_t* d_panel_row = nullptr;// [n_panels+1] first row of every panel
    int32_t* d_panel_need = nullptr;// [n_panels] panels of the block whose forward axpy must be complete
                                   //       before the chain warp may start this panel
    int32_t* d_blk_order = nullptr;// [n_blocks] unit ids grouped by phase, most expensive first inside a phase (LPT)

    std::vector<int32_t> h_blk_row;     // host copy of the sweep-unit boundaries
    std::vector<int32_t> h_ldblk_row;   // [n_ld_blocks+1] LD-block boundaries, for callers (sharding across GPUs)
    std::vector<int32_t> h_phase_ptr;   // [n_phases+1] slice of d_blk_order swept by launch p
    // Row chunks (single-phase LD only): contiguous runs of sweep units of about equal cost, so that a caller with HOST
    // state can overlap the copies of one chunk with the sweep of another (viprs_b200_cpp_e_step_resident).
    int32_t n_chunks = 0;
    int32_t* d_chunk_order = nullptr;      // [n_blocks] unit ids grouped by chunk, most expensive first inside a chunk
    std::vector<int32_t> h_chunk_unit;     // [n_chunks+1] slice of d_chunk_order (= unit range, units are in row order)
    std::vector<int32_t> h_chunk_row;      // [n_chunks+1] first row of every chunk
    std::vector<int32_t> h_chunk_item;     // [n_chunks+1] slice of d_items_bwd

    // ---- tiled LD blocks (blocks larger than kTileLimit rows, e.g. 10,240-SNP float64 blocks or banded LD) ----
    // Row j of tile [t0, t1) keeps its columns (j, t1) in the packed layout above (the sequential part, swept by the
    // one-pass kernels) and its columns [t1, row end) in the "ext" layout below: these rectangles enter the sweep as
    // two streaming matrix-vector products, bext[j] = sum_k R_jk eta_old[k] before anything is swept and
    // fext[k] += sum_j R_jk eta_new[j] after tile j's launch (DESIGN.md "tiled sweep").
    void* d_ext = nullptr;         // [ext_elems] ext entries, rows 16B-aligned, biased integer codes
    int64_t* d_erow = nullptr;     // [M+1] element offset of each ext row
    int32_t* d_ecs = nullptr;      // [M]   first (aligned) column of each ext row, global index
    int64_t ext_elems = 0;
    int4* d_items_diag = nullptr;  // [n_blocks] {row0, row1, col0, col1} of every unit's packed triangle
    int4* d_items_ext = nullptr;   // [n_items_ext] the same for every unit's ext rectangle, grouped by phase
    int32_t n_items_ext = 0;
    int4* d_items_bwd = nullptr;   // row_dot_kernel items: <= 64-row chunks of every unit's packed rows ...
    int4* d_items_bwd_ext = nullptr;   // ... and of every unit's ext rows
    int32_t n_items_bwd = 0, n_items_bwd_ext = 0;
    std::vector<int32_t> h_ext_phase_ptr;   // [n_phases+1] slice of d_items_ext belonging to phase p
    std::vector<int32_t> h_bwd_ext_phase_ptr;   // [n_phases+1] slice of d_items_bwd_ext (sorted by phase)
    mutable cudaStream_t side_stream = nullptr;     // the backward-external dots of later tiles run here, next to the sweeps
    mutable std::vector<cudaEvent_t> side_events;   // [n_phases + 1]: fork + one per phase
    std::vector<int32_t> h_items_cols;      // max columns of an item per phase / overall (grid sizing)
    void* d_unit_partial = nullptr;   // [n_blocks][VIPRS_B200_NSUMS] doubles: per-unit sums of the fused sweep
    mutable void* d_host_ws = nullptr;   // staging of HOST state arrays (viprs_b200_cpp_e_step_resident), grown on demand
    mutable int64_t host_ws_bytes = 0;
    mutable void* d_fext = nullptr;   // [M] scratch of the state type (8 bytes per row): forward-external accumulator
    mutable void* d_bext = nullptr;   // [M] scratch: backward-external dots

    // dense symmetric block layout for the grid sweep (built lazily by vb::ensure_dense): block b is a
    // B_b x Bp_b row-major matrix (Bp_b = B_b rounded up to 16 elements), zero diagonal, biased integer codes
    mutable void* d_dense = nullptr;
    mutable int64_t* d_dblk_off = nullptr;   // [n_blocks] byte offset of every block (128-byte aligned)
    mutable int64_t dense_bytes = 0;
};

namespace vb {
constexpr int kTileLimit = 4096;   // LD blocks up to this many rows are one sweep unit
constexpr int kTileRows = 1024;    // tile size of larger blocks (measured on the C5 workload: 2048 -> 19.6 ms, 1024 -> 15.3 ms, 512 -> 15.0 ms per sweep)
constexpr int kChunks = 4;         // row chunks for copy / sweep overlap of host-state callers
// shared memory per CTA that lets two CTAs share one SM (228 KB per SM, 1 KB reserved per CTA)
constexpr int kSmemTwoPerSM = 113 * 1024;
struct RingGeometry { int nst; int smem_bytes; int ctas_per_sm; };
// ring depth / dynamic shared memory of the sweep for a state type of `tsize` bytes; nst == 0: does not fit
RingGeometry ring_geometry(const viprs_b200_ld* ld, int tsize);
// same for the register-resident kernel (float32 state, blocks <= 4096 SNPs); nst == 0: not applicable
RingGeometry fast_ring_geometry(const viprs_b200_ld* ld);
// build the dense symmetric block layout if it does not exist yet; 0 or an error code
int ensure_dense(const viprs_b200_ld* ld, cudaStream_t stream);
// incremental-q sweeps of one row chunk (chunk < 0: everything), see launch.cuh launch_incremental; defined in
// slab_f32.cu / mix_f32.cu, used by the host-state drop-ins in api.cu.  `swept` (nullable) is recorded on `st` between
// the sweep and the update_q_factor pass: everything but q is final from there on.
int incr_slab_f32(const viprs_b200_ld* ld, const float* std_beta, float* var_gamma, float* var_mu, float* eta, float* q,
                  float* eta_diff, const float* u_logs, const float* shvt, const float* mu_mult, float dq, int chunk,
                  cudaStream_t st, cudaEvent_t swept = nullptr);
int incr_mix_f32(const viprs_b200_ld* ld, int K, const float* std_beta, float* var_gamma, float* var_mu, float* eta, float* q,
                 float* eta_diff, const float* log_null_pi, const float* u_logs, const float* shvt, const float* mu_mult,
                 float dq, int chunk, cudaStream_t st, cudaEvent_t swept = nullptr);
}  // namespace vb
