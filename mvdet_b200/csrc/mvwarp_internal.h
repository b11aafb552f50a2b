// mvwarp_internal.h -- plan layout and kernel launch interfaces shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mvwarp.h"

namespace mvw {

// One inverse-map entry: grid cell (y*Wg + x) and the bilinear weight with which it samples the texel.
struct CsrEntry {
  uint32_t cell;
  float w;
};

// Backward work decomposition (built with the CSR): a regular task is a run of consecutive rows whose
// weight (1 per texel + 1 per entry) fits TASK_BUDGET (32), processed by one warp; a row longer than HEAVY_LEN
// is "heavy" and gets a whole CTA (its entries split over the warps, partials combined in warp order).
constexpr uint32_t TASK_BUDGET = 32, TASK_MAX_TEX = 31, HEAVY_LEN = 31;  // swept on B200: profiles/README.md
constexpr int CSR_TILE = 8;  // rows are ordered tile-major over CSR_TILE x CSR_TILE texel tiles

struct Task {
  uint32_t row0;   // first row (tile-major texel position, view-relative)
  uint16_t nrows;  // <= TASK_MAX_TEX
  uint16_t view;
};

// Host-side CSR for all views (built once per rig).  Row r of view v describes the texel tex_of[v][r]:
// rows follow a tile-major order of the source image so that rows close in the array are close in 2-D.
struct HostCsr {
  std::vector<uint32_t> row_ptr;     // [V * (Hs*Ws + 1)], offsets relative to view_base[v]
  std::vector<uint32_t> tex_of;      // [V * Hs*Ws]  row -> linear texel index ys*Ws + xs
  std::vector<CsrEntry> entries;     // [nnz], per view, rows in order, row entries in (cell, tap) order
  std::vector<uint64_t> view_base;   // [V + 1]
  std::vector<Task> tasks;           // regular tasks, grouped by view
  std::vector<uint32_t> task_base;   // [V + 1]
  std::vector<uint32_t> heavy;       // heavy rows (view-relative row index), grouped by view
  std::vector<uint32_t> heavy_base;  // [V + 1]
  uint64_t unique_texels = 0, valid_cells = 0;
  uint32_t max_row = 0;
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

// Host: normalised grid [V,Hg,Wg,2] from the matrices (bit-identical to the oracle's torch ops).
void host_grid(const float* T, int V, int Hg, int Wg, float* grid_out);
// Host: inverse CSR map from a normalised grid.
void host_build_csr(const float* grid, int V, int Hs, int Ws, int Hg, int Wg, HostCsr& csr);

}  // namespace mvw

// The plan: immutable after creation.
struct mvwarp_plan {
  int device = 0;
  int V = 0, Hs = 0, Ws = 0, Hg = 0, Wg = 0;
  int inline_coords = 0;
  int num_sms = 148;
  // device memory (one allocation, carved)
  void* d_base = nullptr;
  size_t d_bytes = 0;
  float* d_T = nullptr;            // [V,9]          (inline mode)
  float* d_grid = nullptr;         // [V,Hg,Wg,2]    (table mode)
  float* d_coord_x = nullptr;      // [Wg]  coord_map channel 0 along x
  float* d_coord_y = nullptr;      // [Hg]  coord_map channel 1 along y
  uint32_t* d_row_ptr = nullptr;   // [V*(Hs*Ws+1)]
  uint32_t* d_tex_of = nullptr;    // [V*Hs*Ws]
  mvw::CsrEntry* d_entries = nullptr;  // [nnz]
  uint64_t* d_view_base = nullptr; // [V+1]
  mvw::Task* d_tasks = nullptr;    // regular backward tasks
  uint32_t* d_heavy = nullptr;     // heavy rows
  uint32_t* d_heavy_base = nullptr;  // [V+1]
  std::vector<uint32_t> task_base, heavy_base;  // host copies, [V+1]
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
};

int launch_fwd(const LaunchArgs& a);
int launch_bwd(const LaunchArgs& a);
int launch_device_grid(const mvwarp_plan* plan, float* d_grid_out, cudaStream_t stream);

}  // namespace mvw
