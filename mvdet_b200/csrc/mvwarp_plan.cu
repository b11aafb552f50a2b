// mvwarp_plan.cu -- rig-constant plan: host coordinate chain, inverse CSR map, device upload, C ABI.
//
// The reference redoes all of this on every forward call (persp_trans_detector.py:68-69: H2D of
// the matrix, ~25 tiny kornia launches, two torch.inverse syncs per camera); here it is done once
// per camera rig.  Compile with -Xcompiler -ffp-contract=off (see mvwarp_coords.h).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <thread>

#include "mvwarp_coords.h"
#include "mvwarp_internal.h"

namespace mvw {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  set_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
  return MVWARP_ECUDA;
}

void host_grid(const float* T, int V, int Hg, int Wg, float* grid_out) {
  for (int v = 0; v < V; ++v)
    for (int y = 0; y < Hg; ++y)
      for (int x = 0; x < Wg; ++x) {
        float sx, sy;
        norm_src_coord(T + 9 * v, x, y, Hg, Wg, sx, sy);
        float* g = grid_out + (((size_t)v * Hg + y) * Wg + x) * 2;
        g[0] = sx;
        g[1] = sy;
      }
}

// Tile-major order of the source image: CSR_TILE x CSR_TILE tiles row-major, texels row-major inside a tile.
static void tile_major_order(int Hs, int Ws, std::vector<uint32_t>& tex_of, std::vector<uint32_t>& row_of) {
  const size_t ntex = (size_t)Hs * Ws;
  tex_of.resize(ntex);
  row_of.resize(ntex);
  size_t r = 0;
  for (int ty = 0; ty < Hs; ty += CSR_TILE)
    for (int tx = 0; tx < Ws; tx += CSR_TILE)
      for (int y = ty; y < std::min(Hs, ty + CSR_TILE); ++y)
        for (int x = tx; x < std::min(Ws, tx + CSR_TILE); ++x) {
          const uint32_t tex = (uint32_t)((size_t)y * Ws + x);
          tex_of[r] = tex;
          row_of[tex] = (uint32_t)r++;
        }
}

struct ViewCsr {
  std::vector<CsrEntry> entries;
  std::vector<Task> tasks;
  std::vector<uint32_t> heavy;
  uint64_t unique = 0, valid = 0;
  uint32_t max_row = 0;
};

// Counting sort of the in-bounds taps of one view by row; entry order inside a row is (cell, tap).
// Then the backward tasks: greedy runs of consecutive light rows under TASK_BUDGET, heavy rows apart.
static void build_view_csr(const float* grid, int view, int Hs, int Ws, int Hg, int Wg, const std::vector<uint32_t>& row_of,
                           uint32_t* row_ptr, ViewCsr& out) {
  const size_t ntex = (size_t)Hs * Ws, ncell = (size_t)Hg * Wg;
  std::vector<uint32_t> count(ntex + 1, 0);
  static const int dx[4] = {0, 1, 0, 1}, dy[4] = {0, 0, 1, 1};
  out.valid = 0;
  for (size_t cell = 0; cell < ncell; ++cell) {
    const Taps t = make_taps(grid[2 * cell], grid[2 * cell + 1], Hs, Ws);
    if (t.mask) ++out.valid;
    for (int k = 0; k < 4; ++k)
      if (t.mask >> k & 1) ++count[row_of[(size_t)(t.y0 + dy[k]) * Ws + (t.x0 + dx[k])]];
  }
  uint32_t run = 0;
  out.unique = 0;
  out.max_row = 0;
  for (size_t i = 0; i < ntex; ++i) {
    const uint32_t c = count[i];
    row_ptr[i] = run;
    count[i] = run;  // becomes the fill cursor
    run += c;
    if (c) ++out.unique;
    out.max_row = std::max(out.max_row, c);
  }
  row_ptr[ntex] = run;
  out.entries.resize(run);
  for (size_t cell = 0; cell < ncell; ++cell) {
    const Taps t = make_taps(grid[2 * cell], grid[2 * cell + 1], Hs, Ws);
    for (int k = 0; k < 4; ++k)
      if (t.mask >> k & 1) {
        const size_t row = row_of[(size_t)(t.y0 + dy[k]) * Ws + (t.x0 + dx[k])];
        out.entries[count[row]++] = CsrEntry{(uint32_t)cell, t.w[k]};
      }
  }
  // backward tasks (budget / heavy threshold overridable for tuning runs only)
  static const uint32_t budget = getenv("MVWARP_TASK_BUDGET") ? (uint32_t)atoi(getenv("MVWARP_TASK_BUDGET")) : TASK_BUDGET;
  static const uint32_t heavy_len = getenv("MVWARP_HEAVY_LEN") ? (uint32_t)atoi(getenv("MVWARP_HEAVY_LEN")) : HEAVY_LEN;
  Task cur{0, 0, (uint16_t)view};
  uint32_t weight = 0;
  auto flush = [&]() {
    if (cur.nrows) out.tasks.push_back(cur);
    cur.nrows = 0;
    weight = 0;
  };
  for (size_t r = 0; r < ntex; ++r) {
    const uint32_t len = row_ptr[r + 1] - row_ptr[r];
    if (len > heavy_len) {
      flush();
      out.heavy.push_back((uint32_t)r);
      continue;
    }
    if (cur.nrows && (weight + 1 + len > budget || cur.nrows >= TASK_MAX_TEX)) flush();
    if (!cur.nrows) cur.row0 = (uint32_t)r;
    ++cur.nrows;
    weight += 1 + len;
  }
  flush();
}

void host_build_csr(const float* grid, int V, int Hs, int Ws, int Hg, int Wg, HostCsr& csr) {
  const size_t ntex = (size_t)Hs * Ws, ncell = (size_t)Hg * Wg;
  std::vector<uint32_t> tex_of, row_of;
  tile_major_order(Hs, Ws, tex_of, row_of);
  csr.row_ptr.assign((size_t)V * (ntex + 1), 0);
  csr.tex_of.resize((size_t)V * ntex);
  for (int v = 0; v < V; ++v) memcpy(csr.tex_of.data() + (size_t)v * ntex, tex_of.data(), ntex * sizeof(uint32_t));
  csr.view_base.assign(V + 1, 0);
  csr.task_base.assign(V + 1, 0);
  csr.heavy_base.assign(V + 1, 0);
  std::vector<ViewCsr> per_view(V);
  const int nthreads = std::max(1, std::min(V, (int)std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&, t]() {
      for (int v = t; v < V; v += nthreads)
        build_view_csr(grid + (size_t)v * ncell * 2, v, Hs, Ws, Hg, Wg, row_of, csr.row_ptr.data() + (size_t)v * (ntex + 1),
                       per_view[v]);
    });
  for (auto& th : pool) th.join();
  uint64_t total = 0;
  for (int v = 0; v < V; ++v) {
    csr.view_base[v] = total;
    total += per_view[v].entries.size();
    csr.task_base[v + 1] = csr.task_base[v] + (uint32_t)per_view[v].tasks.size();
    csr.heavy_base[v + 1] = csr.heavy_base[v] + (uint32_t)per_view[v].heavy.size();
  }
  csr.view_base[V] = total;
  csr.entries.resize(total);
  csr.tasks.reserve(csr.task_base[V]);
  csr.heavy.reserve(csr.heavy_base[V]);
  csr.unique_texels = csr.valid_cells = 0;
  csr.max_row = 0;
  for (int v = 0; v < V; ++v) {
    ViewCsr& pv = per_view[v];
    if (!pv.entries.empty())
      memcpy(csr.entries.data() + csr.view_base[v], pv.entries.data(), pv.entries.size() * sizeof(CsrEntry));
    csr.tasks.insert(csr.tasks.end(), pv.tasks.begin(), pv.tasks.end());
    csr.heavy.insert(csr.heavy.end(), pv.heavy.begin(), pv.heavy.end());
    std::vector<CsrEntry>().swap(pv.entries);
    csr.unique_texels += pv.unique;
    csr.valid_cells += pv.valid;
    csr.max_row = std::max(csr.max_row, pv.max_row);
  }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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

static int check_shapes(int V, int Hs, int Ws, int Hg, int Wg) {
  if (V <= 0 || Hs <= 0 || Ws <= 0 || Hg <= 0 || Wg <= 0) {
    set_error("plan: V, Hs, Ws, Hg, Wg must all be positive");
    return MVWARP_EINVAL;
  }
  if ((uint64_t)Hs * Ws >= (1ull << 31) || (uint64_t)Hg * Wg >= (1ull << 31)) {
    set_error("plan: image or grid too large for 32-bit texel/cell indices");
    return MVWARP_EUNSUPPORTED;
  }
  return MVWARP_OK;
}

static int plan_create_impl(const float* T, const float* grid_in, int V, int Hs, int Ws, int Hg, int Wg, int device,
                            mvwarp_plan_t** plan_out) {
  if (!plan_out || (!T && !grid_in)) {
    set_error("plan_create: null argument");
    return MVWARP_EINVAL;
  }
  *plan_out = nullptr;
  int rc = check_shapes(V, Hs, Ws, Hg, Wg);
  if (rc) return rc;
  try {
    const size_t ncell = (size_t)Hg * Wg, ntex = (size_t)Hs * Ws;
    std::vector<float> grid_local;
    const float* grid = grid_in;
    if (!grid) {
      grid_local.resize((size_t)V * ncell * 2);
      host_grid(T, V, Hg, Wg, grid_local.data());
      grid = grid_local.data();
    }
    HostCsr csr;
    host_build_csr(grid, V, Hs, Ws, Hg, Wg, csr);
    if (csr.view_base[V] >= (1ull << 32)) {
      set_error("plan: more than 2^32 taps");
      return MVWARP_EUNSUPPORTED;
    }
    std::vector<float> cx(Wg), cy(Hg);
    for (int x = 0; x < Wg; ++x) cx[x] = coord_map_value(x, Wg);
    for (int y = 0; y < Hg; ++y) cy[y] = coord_map_value(y, Hg);

    mvwarp_plan* p = new mvwarp_plan();
    p->device = device;
    p->V = V; p->Hs = Hs; p->Ws = Ws; p->Hg = Hg; p->Wg = Wg;
    p->inline_coords = T ? 1 : 0;

    DeviceGuard guard(device);
    if (!guard.ok) { delete p; return cuda_fail(cudaGetLastError(), "cudaSetDevice"); }
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaGetDeviceProperties"); }
    p->num_sms = prop.multiProcessorCount;

    // carve one allocation
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_T = carve((size_t)V * 9 * sizeof(float));
    const size_t o_grid = carve(T ? 0 : (size_t)V * ncell * 2 * sizeof(float));
    const size_t o_cx = carve(Wg * sizeof(float)), o_cy = carve(Hg * sizeof(float));
    const size_t o_rp = carve((size_t)V * (ntex + 1) * sizeof(uint32_t));
    const size_t o_to = carve((size_t)V * ntex * sizeof(uint32_t));
    const size_t o_en = carve(csr.entries.size() * sizeof(CsrEntry));
    const size_t o_vb = carve((V + 1) * sizeof(uint64_t));
    const size_t o_tk = carve(csr.tasks.size() * sizeof(Task));
    const size_t o_hv = carve(csr.heavy.size() * sizeof(uint32_t));
    const size_t o_hb = carve((V + 1) * sizeof(uint32_t));
    p->d_bytes = off;
    e = cudaMalloc(&p->d_base, p->d_bytes);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(plan)"); }
    char* base = (char*)p->d_base;
    p->d_T = (float*)(base + o_T);
    p->d_grid = T ? nullptr : (float*)(base + o_grid);
    p->d_coord_x = (float*)(base + o_cx);
    p->d_coord_y = (float*)(base + o_cy);
    p->d_row_ptr = (uint32_t*)(base + o_rp);
    p->d_tex_of = (uint32_t*)(base + o_to);
    p->d_entries = (CsrEntry*)(base + o_en);
    p->d_view_base = (uint64_t*)(base + o_vb);
    p->d_tasks = (Task*)(base + o_tk);
    p->d_heavy = (uint32_t*)(base + o_hv);
    p->d_heavy_base = (uint32_t*)(base + o_hb);
    p->task_base = csr.task_base;
    p->heavy_base = csr.heavy_base;

    auto up = [&](void* dst, const void* src, size_t bytes) {
      if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    };
    if (T) up(p->d_T, T, (size_t)V * 9 * sizeof(float));
    else { e = cudaMemset(p->d_T, 0, (size_t)V * 9 * sizeof(float)); up(p->d_grid, grid, (size_t)V * ncell * 2 * sizeof(float)); }
    up(p->d_coord_x, cx.data(), Wg * sizeof(float));
    up(p->d_coord_y, cy.data(), Hg * sizeof(float));
    up(p->d_row_ptr, csr.row_ptr.data(), csr.row_ptr.size() * sizeof(uint32_t));
    up(p->d_tex_of, csr.tex_of.data(), csr.tex_of.size() * sizeof(uint32_t));
    up(p->d_entries, csr.entries.data(), csr.entries.size() * sizeof(CsrEntry));
    up(p->d_tasks, csr.tasks.data(), csr.tasks.size() * sizeof(Task));
    up(p->d_heavy, csr.heavy.data(), csr.heavy.size() * sizeof(uint32_t));
    up(p->d_heavy_base, csr.heavy_base.data(), csr.heavy_base.size() * sizeof(uint32_t));
    up(p->d_view_base, csr.view_base.data(), csr.view_base.size() * sizeof(uint64_t));
    if (e != cudaSuccess) { cudaFree(p->d_base); delete p; return cuda_fail(e, "plan upload"); }

    p->stats.V = V; p->stats.Hs = Hs; p->stats.Ws = Ws; p->stats.Hg = Hg; p->stats.Wg = Wg;
    p->stats.inline_coords = p->inline_coords;
    p->stats.unique_texels = csr.unique_texels;
    p->stats.nnz = csr.view_base[V];
    p->stats.valid_cells = csr.valid_cells;
    p->stats.max_row = csr.max_row;
    p->stats.bwd_tasks = csr.task_base[V];
    p->stats.heavy_rows = csr.heavy_base[V];
    p->stats.device_bytes = p->d_bytes;
    *plan_out = p;
    return MVWARP_OK;
  } catch (const std::bad_alloc&) {
    set_error("plan_create: host allocation failed");
    return MVWARP_ENOMEM;
  } catch (const std::exception& ex) {
    set_error(std::string("plan_create: ") + ex.what());
    return MVWARP_EINVAL;
  }
}

}  // namespace mvw

using namespace mvw;

extern "C" {

int mvwarp_version(void) { return MVWARP_VERSION; }

const char* mvwarp_last_error(void) { return g_last_error.c_str(); }

int mvwarp_host_grid(const float* T, int V, int Hg, int Wg, float* grid_out) {
  if (!T || !grid_out || V <= 0 || Hg <= 0 || Wg <= 0) {
    set_error("host_grid: bad argument");
    return MVWARP_EINVAL;
  }
  host_grid(T, V, Hg, Wg, grid_out);
  return MVWARP_OK;
}

int mvwarp_host_csr(const float* grid_host, int V, int Hs, int Ws, int Hg, int Wg, mvwarp_stats_t* stats_out,
                    uint32_t* tex_of_out, uint32_t* row_ptr_out, uint32_t* cells_out, float* weights_out,
                    uint64_t capacity) {
  if (!grid_host || !stats_out) { set_error("host_csr: null argument"); return MVWARP_EINVAL; }
  int rc = check_shapes(V, Hs, Ws, Hg, Wg);
  if (rc) return rc;
  try {
    HostCsr csr;
    host_build_csr(grid_host, V, Hs, Ws, Hg, Wg, csr);
    memset(stats_out, 0, sizeof(*stats_out));
    stats_out->V = V; stats_out->Hs = Hs; stats_out->Ws = Ws; stats_out->Hg = Hg; stats_out->Wg = Wg;
    stats_out->unique_texels = csr.unique_texels;
    stats_out->nnz = csr.view_base[V];
    stats_out->valid_cells = csr.valid_cells;
    stats_out->max_row = csr.max_row;
    stats_out->bwd_tasks = csr.task_base[V];
    stats_out->heavy_rows = csr.heavy_base[V];
    if (tex_of_out) memcpy(tex_of_out, csr.tex_of.data(), csr.tex_of.size() * sizeof(uint32_t));
    if (row_ptr_out) memcpy(row_ptr_out, csr.row_ptr.data(), csr.row_ptr.size() * sizeof(uint32_t));
    if (cells_out || weights_out) {
      if (capacity < csr.entries.size()) { set_error("host_csr: capacity smaller than nnz"); return MVWARP_EINVAL; }
      for (size_t i = 0; i < csr.entries.size(); ++i) {
        if (cells_out) cells_out[i] = csr.entries[i].cell;
        if (weights_out) weights_out[i] = csr.entries[i].w;
      }
    }
    return MVWARP_OK;
  } catch (const std::bad_alloc&) {
    set_error("host_csr: host allocation failed");
    return MVWARP_ENOMEM;
  }
}

int mvwarp_plan_create(const float* T, int V, int Hs, int Ws, int Hg, int Wg, int device, mvwarp_plan_t** plan_out) {
  if (!T) { set_error("plan_create: null matrices"); return MVWARP_EINVAL; }
  return plan_create_impl(T, nullptr, V, Hs, Ws, Hg, Wg, device, plan_out);
}

int mvwarp_plan_create_from_grid(const float* grid_host, int V, int Hs, int Ws, int Hg, int Wg, int device,
                                 mvwarp_plan_t** plan_out) {
  if (!grid_host) { set_error("plan_create_from_grid: null grid"); return MVWARP_EINVAL; }
  return plan_create_impl(nullptr, grid_host, V, Hs, Ws, Hg, Wg, device, plan_out);
}

int mvwarp_plan_destroy(mvwarp_plan_t* plan) {
  if (!plan) return MVWARP_OK;
  DeviceGuard guard(plan->device);
  cudaError_t e = plan->d_base ? cudaFree(plan->d_base) : cudaSuccess;
  delete plan;
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "cudaFree(plan)");
}

int mvwarp_plan_stats(const mvwarp_plan_t* plan, mvwarp_stats_t* out) {
  if (!plan || !out) { set_error("plan_stats: null argument"); return MVWARP_EINVAL; }
  *out = plan->stats;
  return MVWARP_OK;
}

int mvwarp_plan_device_grid(const mvwarp_plan_t* plan, float* grid_host_out, void* stream) {
  if (!plan || !grid_host_out) { set_error("plan_device_grid: null argument"); return MVWARP_EINVAL; }
  DeviceGuard guard(plan->device);
  const size_t bytes = (size_t)plan->V * plan->Hg * plan->Wg * 2 * sizeof(float);
  float* d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(grid)");
  int rc = launch_device_grid(plan, d, (cudaStream_t)stream);
  if (rc == MVWARP_OK) {
    e = cudaMemcpyAsync(grid_host_out, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "grid copy");
  }
  cudaFree(d);
  return rc;
}

static int check_call(const mvwarp_plan_t* plan, int view0, int nviews, const void* a, const void* b,
                      const int64_t* fs, const int64_t* os, int B, int C, int dtype, const char* who) {
  if (!plan || !a || !b || !fs || !os) { set_error(std::string(who) + ": null argument"); return MVWARP_EINVAL; }
  if (B <= 0 || C <= 0 || nviews <= 0 || view0 < 0 || view0 + nviews > plan->V) {
    set_error(std::string(who) + ": bad B/C/view range");
    return MVWARP_EINVAL;
  }
  if (dtype != MVWARP_F32 && dtype != MVWARP_BF16) { set_error(std::string(who) + ": unknown dtype"); return MVWARP_EINVAL; }
  return MVWARP_OK;
}

int mvwarp_warp_views_fwd(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats,
                          const int64_t feat_strides[5], void* out, const int64_t out_strides[4], int out_chan0,
                          int coord_chan, int B, int C, int dtype, void* stream) {
  int rc = check_call(plan, view0, nviews, feats, out, feat_strides, out_strides, B, C, dtype, "warp_views_fwd");
  if (rc) return rc;
  if (out_chan0 < 0) { set_error("warp_views_fwd: negative out_chan0"); return MVWARP_EINVAL; }
  DeviceGuard guard(plan->device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  LaunchArgs a{};
  a.plan = plan; a.view0 = view0; a.nviews = nviews; a.B = B; a.C = C; a.dtype = dtype;
  a.feats_or_grad_out = feats; a.out_or_grad_feats = out;
  memcpy(a.fs, feat_strides, sizeof(a.fs));
  memcpy(a.os, out_strides, sizeof(a.os));
  a.out_chan0 = out_chan0; a.coord_chan = coord_chan; a.stream = (cudaStream_t)stream;
  return launch_fwd(a);
}

int mvwarp_warp_views_bwd(const mvwarp_plan_t* plan, int view0, int nviews, const void* grad_out,
                          const int64_t out_strides[4], int out_chan0, void* grad_feats,
                          const int64_t feat_strides[5], int B, int C, int dtype, void* stream) {
  int rc = check_call(plan, view0, nviews, grad_out, grad_feats, feat_strides, out_strides, B, C, dtype, "warp_views_bwd");
  if (rc) return rc;
  if (out_chan0 < 0) { set_error("warp_views_bwd: negative out_chan0"); return MVWARP_EINVAL; }
  DeviceGuard guard(plan->device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  LaunchArgs a{};
  a.plan = plan; a.view0 = view0; a.nviews = nviews; a.B = B; a.C = C; a.dtype = dtype;
  a.feats_or_grad_out = grad_out; a.out_or_grad_feats = grad_feats;
  memcpy(a.fs, feat_strides, sizeof(a.fs));
  memcpy(a.os, out_strides, sizeof(a.os));
  a.out_chan0 = out_chan0; a.coord_chan = -1; a.stream = (cudaStream_t)stream;
  return launch_bwd(a);
}

}  // extern "C"
