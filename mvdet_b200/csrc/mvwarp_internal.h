// mvwarp_internal.h -- plan layout and kernel launch interfaces shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mvwarp.h"

namespace mvw {

// One sparse-map entry: column index (a grid cell y*Wg + x, or a source texel ys*Ws + xs) and its weight.
struct CsrEntry {
  uint32_t col;
  float w;
};

// Work decomposition of a CSR gather (built with the CSR): a regular task is a run of consecutive rows whose
// weight (1 per row + 1 per entry) fits TASK_BUDGET (32), processed by one warp; a row longer than HEAVY_LEN
// is "heavy" and gets a whole CTA (its entries split over the warps, partials combined in warp order).
constexpr uint32_t TASK_BUDGET = 32, TASK_MAX_ROWS = 31, HEAVY_LEN = 31;  // swept on B200: profiles/README.md
constexpr int CSR_TILE = 8;  // rows are ordered tile-major over CSR_TILE x CSR_TILE tiles of the row image

struct ZeroSeg {
  uint32_t id;     // linear index of the segment's first texel in the destination image
  uint32_t view;
};

struct Task {
  uint32_t row0;   // first row (tile-major position, view-relative)
  uint16_t nrows;  // <= TASK_MAX_ROWS
  uint16_t view;
};

// Host-side CSR for all views (built once per rig): dst[row_id[r]] = sum over the row of w * src[col].
// Rows follow a tile-major order of the destination image so that rows close in the array are close in 2-D.
struct HostCsr {
  size_t rows_per_view = 0;
  std::vector<uint32_t> row_ptr;     // [V * (rows_per_view + 1)], offsets relative to view_base[v]
  std::vector<uint32_t> row_id;      // [V * rows_per_view]  row -> linear index y*W + x in the destination image
  std::vector<CsrEntry> entries;     // [nnz], per view, rows in order, row entries in a fixed order
  std::vector<uint64_t> view_base;   // [V + 1]
  std::vector<Task> tasks;           // regular tasks, grouped by view
  std::vector<uint32_t> task_base;   // [V + 1]
  std::vector<uint32_t> heavy;       // heavy rows (view-relative row index), grouped by view
  std::vector<uint32_t> heavy_base;  // [V + 1]
  uint64_t nonempty_rows = 0;
  uint32_t max_row = 0;
  // Zero segments (backward maps): tile-row segments -- the <= CSR_TILE consecutive rows that are one image row of one tile, i.e.
  // consecutive texels in memory -- in which EVERY row is empty (60 % of a Wildtrack feature map lies outside the cameras' ground
  // footprints).  They belong to no task: dedicated CTAs of the same launch stream zeros over them with no per-row bookkeeping,
  // and the gather warps only walk rows inside the footprints.  {first texel id (y * W + x, x a multiple of CSR_TILE), view}.
  std::vector<ZeroSeg> zseg;
  std::vector<uint32_t> zseg_base;   // [V + 1]
  // Staging blocks (forward map of fused up-sample plans only): rows are cut into blocks of STAGE_ROWS consecutive rows (one 8x8
  // cell tile); per block the DISTINCT columns (low-res texels) it references, ascending, and per entry its index in that list.
  // The forward kernel brings a block's distinct texel rows into shared memory once (bulk async copies) and blends from there.
  std::vector<uint32_t> blk_ptr;     // [V * (nblk + 1)] offsets into blk_tex (absolute)
  std::vector<uint32_t> blk_tex;     // distinct column ids per block
  std::vector<uint16_t> slot;        // [nnz] slot of every entry within its block's list
};
constexpr uint32_t STAGE_ROWS = 64;

struct PlanMaps {
  HostCsr bwd;            // rows = source texels, columns = grid cells (always built)
  HostCsr fwd;            // rows = grid cells, columns = source texels (only for fused up-sample plans)
  uint64_t valid_cells = 0;
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

// Host: normalised grid [V,Hg,Wg,2] from the matrices (bit-identical to the oracle's torch ops).
void host_grid(const float* T, int V, int Hg, int Wg, float* grid_out);
// Host: inverse map (texel -> cells) of a normalised grid.
void host_build_maps(const float* grid, int V, int Hs, int Ws, int Hg, int Wg, PlanMaps& maps);
// Host: maps of  warp(bilinear_upsample(feats[hl,wl] -> [Hs,Ws]))  composed into one sparse operator on the low-res image.
void host_build_maps_upsampled(const float* grid, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg, PlanMaps& maps);

// Device-side CSR (pointers into the plan's allocation) + host copies of what launches need.
struct DevCsr {
  uint32_t rows_per_view = 0;
  uint32_t* row_ptr = nullptr;
  uint32_t* row_id = nullptr;
  CsrEntry* entries = nullptr;
  uint64_t* view_base = nullptr;
  Task* tasks = nullptr;
  uint32_t* heavy = nullptr;
  uint32_t* heavy_base_dev = nullptr;
  ZeroSeg* zseg = nullptr;                      // zero segments (see HostCsr), or nullptr
  std::vector<uint32_t> task_base, heavy_base, zseg_base;  // host copies, [V+1]
  uint32_t max_row = 0;                         // longest row (host copy)
  uint32_t* blk_ptr = nullptr;                  // staging blocks (see HostCsr), or nullptr
  uint32_t* blk_tex = nullptr;
  uint16_t* slot = nullptr;
  uint32_t max_blk_tex = 0;                     // most distinct columns in one block
};

// Byte offsets of one CSR's arrays inside the plan's single device allocation (also the layout of a saved plan).
struct CsrLayout {
  uint64_t rows_per_view = 0, max_row = 0;
  uint64_t row_ptr = 0, row_id = 0, entries = 0, view_base = 0, tasks = 0, heavy = 0, heavy_base = 0;
  uint64_t blk_ptr = 0, blk_tex = 0, slot = 0, n_blk_tex = 0, max_blk_tex = 0;   // staging blocks (n_blk_tex == 0: none)
  uint64_t zseg = 0, n_zseg = 0;                                                 // zero segments (n_zseg == 0: none)
};
struct PlanLayout {
  uint64_t T = 0, grid = 0, coord_x = 0, coord_y = 0;
  CsrLayout bwd, fwd;
  uint64_t touched = 0, touched_base = 0;   // feature texels with >= 1 tap, ascending per view; [V+1] offsets (mvwarp_copy_touched_texels)
};

}  // namespace mvw

// The plan: immutable after creation.
struct mvwarp_plan {
  int device = 0;
  int V = 0, Hs = 0, Ws = 0, Hg = 0, Wg = 0;
  int hl = 0, wl = 0;      // fused up-sample plans: size of the low-res feature map the kernels actually read
  int upsampled = 0;
  int inline_coords = 0;
  int num_sms = 148;
  // device memory (one allocation, carved)
  void* d_base = nullptr;
  size_t d_bytes = 0;
  float* d_T = nullptr;            // [V,9]          (inline mode)
  float* d_grid = nullptr;         // [V,Hg,Wg,2]    (table mode)
  float* d_coord_x = nullptr;      // [Wg]  coord_map channel 0 along x
  float* d_coord_y = nullptr;      // [Hg]  coord_map channel 1 along y
  mvw::DevCsr bwd, fwd;
  uint32_t* d_touched = nullptr;           // [U] linear texel ids (ys*W + xs of the feature map the kernels read), ascending per view
  std::vector<uint32_t> touched_base;      // host copy, [V+1]
  mvw::PlanLayout layout;          // where everything lives inside d_base (mvwarp_plan_save / _load)
  std::vector<float> h_T;          // host copy of the matrices (inline mode), for mvwarp_plan_matrices
  mvwarp_stats_t stats{};
};

namespace mvw {

struct LaunchArgs {
  const mvwarp_plan* plan;
  int view0, nviews, B, C, dtype;
  const void* feats_or_grad_out;  // forward: feats; backward: grad_out
  void* out_or_grad_feats;        // forward: out;   backward: grad_feats
  int64_t fs[5];                  // feats / grad_feats strides (B,V,C,H,W)
  int64_t os[4];                  // out / grad_out strides (B,C,H,W)
  int out_chan0, coord_chan;
  cudaStream_t stream;
  // "list" calls (mvwarp_warp_views_fwd_list / _bwd_list): one device pointer per local view for the feature side (feats in the
  // forward, grad_feats in the backward); fs[1] (the V stride) is then unused.  nullptr: one dense [B,V,C,H,W] tensor.
  const void* const* view_ptrs = nullptr;
};

// CUtensorMap (passed as void*: 128 bytes, 64-byte aligned) over packed weight blocks [nblocks][512][64] bf16 seen as a 2-D
// tensor of 128-byte rows, box = 128 rows, no swizzle (the rows are pre-swizzled).  Implemented in mvwarp_conv2.cu.
int make_weight_tensor_map(void* tmap_out, const void* wpack, uint64_t nblocks);

int launch_fwd(const LaunchArgs& a);
int launch_bwd(const LaunchArgs& a);
int launch_device_grid(const mvwarp_plan* plan, float* d_grid_out, cudaStream_t stream);

// RAII: make `dev` current for the scope of a call (the autograd worker thread's current device is not ours to assume)
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

}  // namespace mvw
