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

// Tile-major order of an H x W image: CSR_TILE x CSR_TILE tiles row-major, pixels row-major inside a tile.
static void tile_major_order(int H, int W, std::vector<uint32_t>& id_of_row, std::vector<uint32_t>& row_of_id) {
  const size_t n = (size_t)H * W;
  id_of_row.resize(n);
  row_of_id.resize(n);
  size_t r = 0;
  for (int ty = 0; ty < H; ty += CSR_TILE)
    for (int tx = 0; tx < W; tx += CSR_TILE)
      for (int y = ty; y < std::min(H, ty + CSR_TILE); ++y)
        for (int x = tx; x < std::min(W, tx + CSR_TILE); ++x) {
          const uint32_t id = (uint32_t)((size_t)y * W + x);
          id_of_row[r] = id;
          row_of_id[id] = (uint32_t)r++;
        }
}

struct Triplet {
  uint32_t row;
  CsrEntry e;
};

struct ViewCsr {
  std::vector<CsrEntry> entries;
  std::vector<Task> tasks;
  std::vector<uint32_t> heavy;
  uint64_t nonempty = 0;
  uint32_t max_row = 0;
};

// Stable counting sort of one view's triplets by row (entry order inside a row = insertion order), then the
// gather tasks: greedy runs of consecutive light rows under TASK_BUDGET, heavy rows apart.
static void csr_from_triplets(const std::vector<Triplet>& trip, size_t nrows, int view, uint32_t* row_ptr, ViewCsr& out) {
  std::vector<uint32_t> cursor(nrows + 1, 0);
  for (const Triplet& t : trip) ++cursor[t.row];
  uint32_t run = 0;
  out.nonempty = 0;
  out.max_row = 0;
  for (size_t i = 0; i < nrows; ++i) {
    const uint32_t c = cursor[i];
    row_ptr[i] = run;
    cursor[i] = run;
    run += c;
    if (c) ++out.nonempty;
    out.max_row = std::max(out.max_row, c);
  }
  row_ptr[nrows] = run;
  out.entries.resize(run);
  for (const Triplet& t : trip) out.entries[cursor[t.row]++] = t.e;

  // tasks (budget / heavy threshold overridable for tuning runs only)
  static const uint32_t budget = getenv("MVWARP_TASK_BUDGET") ? (uint32_t)atoi(getenv("MVWARP_TASK_BUDGET")) : TASK_BUDGET;
  static const uint32_t heavy_len = getenv("MVWARP_HEAVY_LEN") ? (uint32_t)atoi(getenv("MVWARP_HEAVY_LEN")) : HEAVY_LEN;
  Task cur{0, 0, (uint16_t)view};
  uint32_t weight = 0;
  auto flush = [&]() {
    if (cur.nrows) out.tasks.push_back(cur);
    cur.nrows = 0;
    weight = 0;
  };
  for (size_t r = 0; r < nrows; ++r) {
    const uint32_t len = row_ptr[r + 1] - row_ptr[r];
    if (len > heavy_len) {
      flush();
      out.heavy.push_back((uint32_t)r);
      continue;
    }
    if (cur.nrows && (weight + 1 + len > budget || cur.nrows >= TASK_MAX_ROWS)) flush();
    if (!cur.nrows) cur.row0 = (uint32_t)r;
    ++cur.nrows;
    weight += 1 + len;
  }
  flush();
}

static void csr_init(HostCsr& csr, int V, size_t rows_per_view, const std::vector<uint32_t>& id_of_row) {
  csr.rows_per_view = rows_per_view;
  csr.row_ptr.assign((size_t)V * (rows_per_view + 1), 0);
  csr.row_id.resize((size_t)V * rows_per_view);
  for (int v = 0; v < V; ++v)
    memcpy(csr.row_id.data() + (size_t)v * rows_per_view, id_of_row.data(), rows_per_view * sizeof(uint32_t));
  csr.view_base.assign(V + 1, 0);
  csr.task_base.assign(V + 1, 0);
  csr.heavy_base.assign(V + 1, 0);
}

static void csr_assemble(HostCsr& csr, int V, std::vector<ViewCsr>& per_view) {
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
  csr.nonempty_rows = 0;
  csr.max_row = 0;
  for (int v = 0; v < V; ++v) {
    ViewCsr& pv = per_view[v];
    if (!pv.entries.empty())
      memcpy(csr.entries.data() + csr.view_base[v], pv.entries.data(), pv.entries.size() * sizeof(CsrEntry));
    csr.tasks.insert(csr.tasks.end(), pv.tasks.begin(), pv.tasks.end());
    csr.heavy.insert(csr.heavy.end(), pv.heavy.begin(), pv.heavy.end());
    std::vector<CsrEntry>().swap(pv.entries);
    csr.nonempty_rows += pv.nonempty;
    csr.max_row = std::max(csr.max_row, pv.max_row);
  }
}

template <typename F>
static void for_each_view_parallel(int V, F&& fn) {
  const int nthreads = std::max(1, std::min(V, (int)std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&, t]() {
      for (int v = t; v < V; v += nthreads) fn(v);
    });
  for (auto& th : pool) th.join();
}

static const int kDx[4] = {0, 1, 0, 1}, kDy[4] = {0, 0, 1, 1};

void host_build_maps(const float* grid, int V, int Hs, int Ws, int Hg, int Wg, PlanMaps& maps) {
  const size_t ntex = (size_t)Hs * Ws, ncell = (size_t)Hg * Wg;
  std::vector<uint32_t> tex_of_row, row_of_tex;
  tile_major_order(Hs, Ws, tex_of_row, row_of_tex);
  csr_init(maps.bwd, V, ntex, tex_of_row);
  std::vector<ViewCsr> per_view(V);
  std::vector<uint64_t> valid(V, 0);
  for_each_view_parallel(V, [&](int v) {
    const float* g = grid + (size_t)v * ncell * 2;
    std::vector<Triplet> trip;
    trip.reserve(4 * ncell);
    for (size_t cell = 0; cell < ncell; ++cell) {  // ascending cell, then tap: the backward's summation order
      const Taps t = make_taps(g[2 * cell], g[2 * cell + 1], Hs, Ws);
      if (t.mask) ++valid[v];
      for (int k = 0; k < 4; ++k)
        if (t.mask >> k & 1)
          trip.push_back({row_of_tex[(size_t)(t.y0 + kDy[k]) * Ws + (t.x0 + kDx[k])], CsrEntry{(uint32_t)cell, t.w[k]}});
    }
    csr_from_triplets(trip, ntex, v, maps.bwd.row_ptr.data() + (size_t)v * (ntex + 1), per_view[v]);
  });
  csr_assemble(maps.bwd, V, per_view);
  maps.valid_cells = 0;
  for (int v = 0; v < V; ++v) maps.valid_cells += valid[v];
}

// One axis of torch.nn.functional.interpolate(mode='bilinear', align_corners=False) (ATen upsample_bilinear2d,
// persp_trans_detector.py:65): source index and weights of every output index, computed in fp32 as ATen does.
struct UpsampleAxis {
  std::vector<int> i0, i1;
  std::vector<float> l0, l1;
};
static UpsampleAxis upsample_axis(int in, int out) {
  UpsampleAxis a;
  a.i0.resize(out); a.i1.resize(out); a.l0.resize(out); a.l1.resize(out);
  const float scale = (float)in / (float)out;
  for (int d = 0; d < out; ++d) {
    float real = scale * ((float)d + 0.5f) - 0.5f;
    if (real < 0.f) real = 0.f;
    const int i0 = std::min((int)real, in - 1);
    const float l1 = std::min(std::max(real - (float)i0, 0.f), 1.f);
    a.i0[d] = i0;
    a.i1[d] = i0 + (i0 < in - 1 ? 1 : 0);
    a.l1[d] = l1;
    a.l0[d] = 1.f - l1;
  }
  return a;
}

void host_build_maps_upsampled(const float* grid, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg, PlanMaps& maps) {
  const size_t nlow = (size_t)hl * wl, ncell = (size_t)Hg * Wg;
  std::vector<uint32_t> low_of_row, row_of_low, cell_of_row, row_of_cell;
  tile_major_order(hl, wl, low_of_row, row_of_low);
  tile_major_order(Hg, Wg, cell_of_row, row_of_cell);
  csr_init(maps.bwd, V, nlow, low_of_row);
  csr_init(maps.fwd, V, ncell, cell_of_row);
  const UpsampleAxis ax = upsample_axis(wl, Ws), ay = upsample_axis(hl, Hs);
  std::vector<ViewCsr> pv_bwd(V), pv_fwd(V);
  std::vector<uint64_t> valid(V, 0);
  for_each_view_parallel(V, [&](int v) {
    const float* g = grid + (size_t)v * ncell * 2;
    std::vector<Triplet> tb, tf;
    tb.reserve(9 * ncell);
    tf.reserve(9 * ncell);
    for (size_t cell = 0; cell < ncell; ++cell) {
      const Taps t = make_taps(g[2 * cell], g[2 * cell + 1], Hs, Ws);
      if (!t.mask) continue;
      ++valid[v];
      // compose: each in-range high-res tap is itself 4 low-res taps; merge duplicates, weights summed in double
      uint32_t col[16];
      double wt[16];
      int n = 0;
      for (int k = 0; k < 4; ++k) {
        if (!(t.mask >> k & 1)) continue;
        const int X = t.x0 + kDx[k], Y = t.y0 + kDy[k];
        const int xs[2] = {ax.i0[X], ax.i1[X]}, ys[2] = {ay.i0[Y], ay.i1[Y]};
        const float lx[2] = {ax.l0[X], ax.l1[X]}, ly[2] = {ay.l0[Y], ay.l1[Y]};
        for (int j = 0; j < 2; ++j)
          for (int i = 0; i < 2; ++i) {
            const double w = (double)t.w[k] * (double)ly[j] * (double)lx[i];
            const uint32_t c = (uint32_t)((size_t)ys[j] * wl + xs[i]);
            int m = 0;
            while (m < n && col[m] != c) ++m;
            if (m == n) { col[n] = c; wt[n] = 0.0; ++n; }
            wt[m] += w;
          }
      }
      // ascending low-res texel: the forward's summation order
      for (int a = 1; a < n; ++a)
        for (int b = a; b > 0 && col[b - 1] > col[b]; --b) { std::swap(col[b - 1], col[b]); std::swap(wt[b - 1], wt[b]); }
      for (int m = 0; m < n; ++m) {
        const float w = (float)wt[m];
        tf.push_back({row_of_cell[cell], CsrEntry{col[m], w}});
        tb.push_back({row_of_low[col[m]], CsrEntry{(uint32_t)cell, w}});   // ascending cell within a row
      }
    }
    csr_from_triplets(tf, ncell, v, maps.fwd.row_ptr.data() + (size_t)v * (ncell + 1), pv_fwd[v]);
    csr_from_triplets(tb, nlow, v, maps.bwd.row_ptr.data() + (size_t)v * (nlow + 1), pv_bwd[v]);
  });
  csr_assemble(maps.fwd, V, pv_fwd);
  csr_assemble(maps.bwd, V, pv_bwd);
  maps.valid_cells = 0;
  for (int v = 0; v < V; ++v) maps.valid_cells += valid[v];
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
  if (V > 65535) {
    set_error("plan: more than 65535 views");
    return MVWARP_EUNSUPPORTED;
  }
  if ((uint64_t)Hs * Ws >= (1ull << 31) || (uint64_t)Hg * Wg >= (1ull << 31)) {
    set_error("plan: image or grid too large for 32-bit texel/cell indices");
    return MVWARP_EUNSUPPORTED;
  }
  return MVWARP_OK;
}

static void fill_stats(mvwarp_stats_t& st, const PlanMaps& maps, int V, int Hs, int Ws, int Hg, int Wg, int hl, int wl) {
  memset(&st, 0, sizeof(st));
  st.V = V; st.Hs = Hs; st.Ws = Ws; st.Hg = Hg; st.Wg = Wg;
  st.unique_texels = maps.bwd.nonempty_rows;
  st.nnz = maps.bwd.view_base[V];
  st.valid_cells = maps.valid_cells;
  st.max_row = maps.bwd.max_row;
  st.bwd_tasks = maps.bwd.task_base[V];
  st.heavy_rows = maps.bwd.heavy_base[V];
  st.lowres_h = hl; st.lowres_w = wl;
  st.fwd_tasks = maps.fwd.task_base.empty() ? 0 : maps.fwd.task_base[V];
}

static int plan_create_impl(const float* T, const float* grid_in, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg,
                            int device, mvwarp_plan_t** plan_out) {
  if (!plan_out || (!T && !grid_in)) {
    set_error("plan_create: null argument");
    return MVWARP_EINVAL;
  }
  *plan_out = nullptr;
  int rc = check_shapes(V, Hs, Ws, Hg, Wg);
  if (rc) return rc;
  const bool upsampled = hl > 0 && wl > 0;
  try {
    const size_t ncell = (size_t)Hg * Wg;
    std::vector<float> grid_local;
    const float* grid = grid_in;
    if (!grid) {
      grid_local.resize((size_t)V * ncell * 2);
      host_grid(T, V, Hg, Wg, grid_local.data());
      grid = grid_local.data();
    }
    PlanMaps maps;
    if (upsampled) host_build_maps_upsampled(grid, V, hl, wl, Hs, Ws, Hg, Wg, maps);
    else host_build_maps(grid, V, Hs, Ws, Hg, Wg, maps);
    if (maps.bwd.view_base[V] >= (1ull << 32)) {
      set_error("plan: more than 2^32 taps");
      return MVWARP_EUNSUPPORTED;
    }
    std::vector<float> cx(Wg), cy(Hg);
    for (int x = 0; x < Wg; ++x) cx[x] = coord_map_value(x, Wg);
    for (int y = 0; y < Hg; ++y) cy[y] = coord_map_value(y, Hg);

    mvwarp_plan* p = new mvwarp_plan();
    p->device = device;
    p->V = V; p->Hs = Hs; p->Ws = Ws; p->Hg = Hg; p->Wg = Wg;
    p->hl = upsampled ? hl : 0; p->wl = upsampled ? wl : 0;
    p->upsampled = upsampled ? 1 : 0;
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
    struct CsrOff { size_t rp, id, en, vb, tk, hv, hb; };
    auto carve_csr = [&](const HostCsr& c) {
      CsrOff o{};
      if (c.rows_per_view == 0) return o;
      o.rp = carve(c.row_ptr.size() * sizeof(uint32_t));
      o.id = carve(c.row_id.size() * sizeof(uint32_t));
      o.en = carve(c.entries.size() * sizeof(CsrEntry));
      o.vb = carve(c.view_base.size() * sizeof(uint64_t));
      o.tk = carve(c.tasks.size() * sizeof(Task));
      o.hv = carve(c.heavy.size() * sizeof(uint32_t));
      o.hb = carve(c.heavy_base.size() * sizeof(uint32_t));
      return o;
    };
    const size_t o_T = carve((size_t)V * 9 * sizeof(float));
    const size_t o_grid = carve(T ? 0 : (size_t)V * ncell * 2 * sizeof(float));
    const size_t o_cx = carve(Wg * sizeof(float)), o_cy = carve(Hg * sizeof(float));
    const CsrOff ob = carve_csr(maps.bwd), of = carve_csr(maps.fwd);
    p->d_bytes = off;
    e = cudaMalloc(&p->d_base, p->d_bytes);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(plan)"); }
    char* base = (char*)p->d_base;
    p->d_T = (float*)(base + o_T);
    p->d_grid = T ? nullptr : (float*)(base + o_grid);
    p->d_coord_x = (float*)(base + o_cx);
    p->d_coord_y = (float*)(base + o_cy);

    auto up = [&](void* dst, const void* src, size_t bytes) {
      if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    };
    auto up_csr = [&](DevCsr& d, const HostCsr& c, const CsrOff& o) {
      if (c.rows_per_view == 0) return;
      d.rows_per_view = (uint32_t)c.rows_per_view;
      d.row_ptr = (uint32_t*)(base + o.rp); d.row_id = (uint32_t*)(base + o.id);
      d.entries = (CsrEntry*)(base + o.en); d.view_base = (uint64_t*)(base + o.vb);
      d.tasks = (Task*)(base + o.tk); d.heavy = (uint32_t*)(base + o.hv); d.heavy_base_dev = (uint32_t*)(base + o.hb);
      d.task_base = c.task_base; d.heavy_base = c.heavy_base;
      up(d.row_ptr, c.row_ptr.data(), c.row_ptr.size() * sizeof(uint32_t));
      up(d.row_id, c.row_id.data(), c.row_id.size() * sizeof(uint32_t));
      up(d.entries, c.entries.data(), c.entries.size() * sizeof(CsrEntry));
      up(d.view_base, c.view_base.data(), c.view_base.size() * sizeof(uint64_t));
      up(d.tasks, c.tasks.data(), c.tasks.size() * sizeof(Task));
      up(d.heavy, c.heavy.data(), c.heavy.size() * sizeof(uint32_t));
      up(d.heavy_base_dev, c.heavy_base.data(), c.heavy_base.size() * sizeof(uint32_t));
    };
    if (T) up(p->d_T, T, (size_t)V * 9 * sizeof(float));
    else { e = cudaMemset(p->d_T, 0, (size_t)V * 9 * sizeof(float)); up(p->d_grid, grid, (size_t)V * ncell * 2 * sizeof(float)); }
    up(p->d_coord_x, cx.data(), Wg * sizeof(float));
    up(p->d_coord_y, cy.data(), Hg * sizeof(float));
    up_csr(p->bwd, maps.bwd, ob);
    up_csr(p->fwd, maps.fwd, of);
    if (e != cudaSuccess) { cudaFree(p->d_base); delete p; return cuda_fail(e, "plan upload"); }

    fill_stats(p->stats, maps, V, Hs, Ws, Hg, Wg, p->hl, p->wl);
    p->stats.inline_coords = p->inline_coords;
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
  return mvwarp_host_csr_upsampled(grid_host, V, 0, 0, Hs, Ws, Hg, Wg, stats_out, tex_of_out, row_ptr_out, cells_out,
                                   weights_out, capacity);
}

int mvwarp_host_csr_upsampled(const float* grid_host, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg,
                              mvwarp_stats_t* stats_out, uint32_t* tex_of_out, uint32_t* row_ptr_out, uint32_t* cells_out,
                              float* weights_out, uint64_t capacity) {
  if (!grid_host || !stats_out) { set_error("host_csr: null argument"); return MVWARP_EINVAL; }
  int rc = check_shapes(V, Hs, Ws, Hg, Wg);
  if (rc) return rc;
  if ((hl > 0) != (wl > 0) || hl < 0) { set_error("host_csr: bad low-res size"); return MVWARP_EINVAL; }
  try {
    PlanMaps maps;
    if (hl > 0) host_build_maps_upsampled(grid_host, V, hl, wl, Hs, Ws, Hg, Wg, maps);
    else host_build_maps(grid_host, V, Hs, Ws, Hg, Wg, maps);
    fill_stats(*stats_out, maps, V, Hs, Ws, Hg, Wg, hl, wl);
    const HostCsr& csr = maps.bwd;
    if (tex_of_out) memcpy(tex_of_out, csr.row_id.data(), csr.row_id.size() * sizeof(uint32_t));
    if (row_ptr_out) memcpy(row_ptr_out, csr.row_ptr.data(), csr.row_ptr.size() * sizeof(uint32_t));
    if (cells_out || weights_out) {
      if (capacity < csr.entries.size()) { set_error("host_csr: capacity smaller than nnz"); return MVWARP_EINVAL; }
      for (size_t i = 0; i < csr.entries.size(); ++i) {
        if (cells_out) cells_out[i] = csr.entries[i].col;
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
  return plan_create_impl(T, nullptr, V, 0, 0, Hs, Ws, Hg, Wg, device, plan_out);
}

int mvwarp_plan_create_from_grid(const float* grid_host, int V, int Hs, int Ws, int Hg, int Wg, int device,
                                 mvwarp_plan_t** plan_out) {
  if (!grid_host) { set_error("plan_create_from_grid: null grid"); return MVWARP_EINVAL; }
  return plan_create_impl(nullptr, grid_host, V, 0, 0, Hs, Ws, Hg, Wg, device, plan_out);
}

int mvwarp_plan_create_upsampled(const float* T, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg, int device,
                                 mvwarp_plan_t** plan_out) {
  if (!T) { set_error("plan_create_upsampled: null matrices"); return MVWARP_EINVAL; }
  if (hl <= 0 || wl <= 0) { set_error("plan_create_upsampled: low-res size must be positive"); return MVWARP_EINVAL; }
  if ((uint64_t)hl * wl >= (1ull << 31)) { set_error("plan_create_upsampled: low-res image too large"); return MVWARP_EUNSUPPORTED; }
  return plan_create_impl(T, nullptr, V, hl, wl, Hs, Ws, Hg, Wg, device, plan_out);
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
