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
    int32_t n_blocks = 0;
    int32_t max_block = 0;      // rows of the largest LD block
    int32_t max_row_bytes = 0;  // longest packed row
    int32_t n_panels = 0;
    int32_t stage_bytes = 0;
    int device = 0;
    int smem_optin = 0;         // cudaDevAttrMaxSharedMemoryPerBlockOptin

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
    int32_t* d_blk_order = nullptr;// [n_blocks] block ids, most expensive first (LPT schedule)

    std::vector<int32_t> h_blk_row;  // host copy for callers (sharding across GPUs)

    // dense symmetric block layout for the grid sweep (built lazily by vb::ensure_dense): block b is a
    // B_b x Bp_b row-major matrix (Bp_b = B_b rounded up to 16 elements), zero diagonal, biased integer codes
    mutable void* d_dense = nullptr;
    mutable int64_t* d_dblk_off = nullptr;   // [n_blocks] byte offset of every block (128-byte aligned)
    mutable int64_t dense_bytes = 0;
};

namespace vb {
// shared memory per CTA that lets two CTAs share one SM (228 KB per SM, 1 KB reserved per CTA)
constexpr int kSmemTwoPerSM = 113 * 1024;
struct RingGeometry { int nst; int smem_bytes; int ctas_per_sm; };
// ring depth / dynamic shared memory of the sweep for a state type of `tsize` bytes; nst == 0: does not fit
RingGeometry ring_geometry(const viprs_b200_ld* ld, int tsize);
// same for the register-resident kernel (float32 state, blocks <= 4096 SNPs); nst == 0: not applicable
RingGeometry fast_ring_geometry(const viprs_b200_ld* ld);
// build the dense symmetric block layout if it does not exist yet; 0 or an error code
int ensure_dense(const viprs_b200_ld* ld, cudaStream_t stream);
}  // namespace vb
