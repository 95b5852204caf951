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
    int32_t n_panels = 0;
    int32_t stage_bytes = 0;
    int device = 0;

    // device arrays
    void* d_packed = nullptr;      // [packed_elems] LD entries, row-major, rows 16B-aligned
    int64_t* d_prow = nullptr;     // [M+1] element offset of each packed row (multiple of epv)
    int32_t* d_pcs = nullptr;      // [M]   first column (global index) of the packed row, aligned
                                   //       to epv relative to the block start
    int32_t* d_blk_row = nullptr;  // [n_blocks+1] first row of every LD block
    int32_t* d_blk_panel = nullptr;// [n_blocks+1] first panel of every LD block
    int32_t* d_panel_row = nullptr;// [n_panels+1] first row of every panel
    int32_t* d_blk_order = nullptr;// [n_blocks] block ids, most expensive first (LPT schedule)

    std::vector<int32_t> h_blk_row;  // host copy for callers (sharding across GPUs)
};

namespace vb {
constexpr int kDefaultStageBytes = 20 * 1024;
// dynamic shared memory the sweep kernel needs for a matrix whose largest block has `max_block`
// rows, with state type of `tsize` bytes (see sweep.cuh for the carve-up)
size_t sweep_smem_bytes(int max_block, int epv, int tsize, int stage_bytes, int n_bulk_warps);
}  // namespace vb
