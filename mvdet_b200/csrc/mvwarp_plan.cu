// mvwarp_plan.cu -- rig-constant plan: host coordinate chain, inverse CSR map, device upload, C ABI.
//
// The reference redoes all of this on every forward call (persp_trans_detector.py:68-69: H2D of
// the matrix, ~25 tiny kornia launches, two torch.inverse syncs per camera); here it is done once
// per camera rig.  Compile with -Xcompiler -ffp-contract=off (see mvwarp_coords.h).
#include <stdlib.h>
#include <string.h>

#include <stdio.h>

#include <algorithm>
#include <memory>
#include <new>
#include <stdexcept>
#include <thread>

#include "mvwarp_coords.h"
#include "mvwarp_uptaps.h"
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
  std::vector<ZeroSeg> zseg;
  uint64_t nonempty = 0;
  uint32_t max_row = 0;
};

// Stable counting sort of one view's triplets by row (entry order inside a row = insertion order), then the
// gather tasks: greedy runs of consecutive light rows under TASK_BUDGET, heavy rows apart.
// zero_width > 0 (backward maps; the width of the destination image whose tile-major order the rows follow): tile-row segments
// whose rows are all empty become zero segments instead of task rows (HostCsr::zseg).
static void csr_from_triplets(const std::vector<Triplet>& trip, size_t nrows, int view, uint32_t* row_ptr, ViewCsr& out,
                              const uint32_t* id_of_row = nullptr, int zero_width = 0) {
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
#ifdef MVWARP_DEV
  static const uint32_t budget = getenv("MVWARP_TASK_BUDGET") ? (uint32_t)atoi(getenv("MVWARP_TASK_BUDGET")) : TASK_BUDGET;
  static const uint32_t heavy_len = getenv("MVWARP_HEAVY_LEN") ? (uint32_t)atoi(getenv("MVWARP_HEAVY_LEN")) : HEAVY_LEN;
#else
  constexpr uint32_t budget = TASK_BUDGET, heavy_len = HEAVY_LEN;
#endif
  Task cur{0, 0, (uint16_t)view};
  uint32_t weight = 0;
  auto flush = [&]() {
    if (cur.nrows) out.tasks.push_back(cur);
    cur.nrows = 0;
    weight = 0;
  };
  for (size_t r = 0; r < nrows; ++r) {
    const uint32_t len = row_ptr[r + 1] - row_ptr[r];
    if (zero_width > 0 && len == 0) {
      // start of a tile-row segment (x a multiple of CSR_TILE; its rows are consecutive in tile-major order and in memory)?
      const uint32_t id = id_of_row[r];
      const int x = (int)(id % (uint32_t)zero_width);
      if (x % CSR_TILE == 0) {
        const size_t tw = (size_t)std::min(CSR_TILE, zero_width - x);
        bool all_empty = r + tw <= nrows;
        for (size_t i = 0; all_empty && i < tw; ++i) all_empty = row_ptr[r + i + 1] == row_ptr[r + i] && id_of_row[r + i] == id + i;
        if (all_empty) {
          flush();   // tasks are runs of CONSECUTIVE rows
          out.zseg.push_back(ZeroSeg{id, (uint32_t)view});
          r += tw - 1;
          continue;
        }
      }
    }
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
  csr.zseg_base.assign(V + 1, 0);
}

static void csr_assemble(HostCsr& csr, int V, std::vector<ViewCsr>& per_view) {
  uint64_t total = 0;
  for (int v = 0; v < V; ++v) {
    csr.view_base[v] = total;
    total += per_view[v].entries.size();
    csr.task_base[v + 1] = csr.task_base[v] + (uint32_t)per_view[v].tasks.size();
    csr.heavy_base[v + 1] = csr.heavy_base[v] + (uint32_t)per_view[v].heavy.size();
    csr.zseg_base[v + 1] = csr.zseg_base[v] + (uint32_t)per_view[v].zseg.size();
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
    csr.zseg.insert(csr.zseg.end(), pv.zseg.begin(), pv.zseg.end());
    std::vector<CsrEntry>().swap(pv.entries);
    csr.nonempty_rows += pv.nonempty;
    csr.max_row = std::max(csr.max_row, pv.max_row);
  }
}

// Every row of every view must belong to exactly one task, be a heavy row, or lie in a zero segment: a row left out would keep
// whatever the output buffer held.  Checked once per plan on the host (cheap: one pass over the work lists).
static void csr_check_coverage(const HostCsr& csr, int V, int zero_width) {
  const size_t R = csr.rows_per_view;
  std::vector<uint8_t> seen(R);
  std::vector<uint32_t> row_of_id;
  if (!csr.zseg.empty()) {
    row_of_id.resize(R);
    for (size_t r = 0; r < R; ++r) row_of_id[csr.row_id[r]] = (uint32_t)r;   // the same order for every view
  }
  for (int v = 0; v < V; ++v) {
    std::fill(seen.begin(), seen.end(), 0);
    auto mark = [&](size_t r) {
      if (r >= R || seen[r]) throw std::logic_error("plan: a CSR row is covered twice or lies outside the map");
      seen[r] = 1;
    };
    for (uint32_t t = csr.task_base[v]; t < csr.task_base[v + 1]; ++t)
      for (uint32_t i = 0; i < csr.tasks[t].nrows; ++i) mark((size_t)csr.tasks[t].row0 + i);
    for (uint32_t i = csr.heavy_base[v]; i < csr.heavy_base[v + 1]; ++i) mark(csr.heavy[i]);
    for (uint32_t i = csr.zseg_base[v]; i < csr.zseg_base[v + 1]; ++i) {
      const uint32_t id = csr.zseg[i].id;
      const int tw = std::min<int>(CSR_TILE, zero_width - (int)(id % (uint32_t)zero_width));
      for (int k = 0; k < tw; ++k) {
        if ((size_t)id + k >= R) throw std::logic_error("plan: a zero segment leaves the map");
        const size_t r = row_of_id[id + k];
        const uint32_t* rp = csr.row_ptr.data() + (size_t)v * (R + 1);
        if (rp[r + 1] != rp[r]) throw std::logic_error("plan: a zero segment covers a non-empty row");
        mark(r);
      }
    }
    for (size_t r = 0; r < R; ++r)
      if (!seen[r]) throw std::logic_error("plan: a CSR row belongs to no task, heavy row or zero segment");
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
    csr_from_triplets(trip, ntex, v, maps.bwd.row_ptr.data() + (size_t)v * (ntex + 1), per_view[v], tex_of_row.data(), Ws);
  });
  csr_assemble(maps.bwd, V, per_view);
  csr_check_coverage(maps.bwd, V, Ws);
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

// Staging blocks of a forward map: distinct columns per block of STAGE_ROWS rows + the slot of every entry (HostCsr).
static void build_stage_blocks(HostCsr& c, int V) {
  const size_t R = c.rows_per_view, nblk = (R + STAGE_ROWS - 1) / STAGE_ROWS;
  c.blk_ptr.assign((size_t)V * (nblk + 1), 0);
  c.blk_tex.clear();
  c.slot.assign(c.entries.size(), 0);
  std::vector<uint32_t> cols;
  for (int v = 0; v < V; ++v) {
    const uint32_t* rp = c.row_ptr.data() + (size_t)v * (R + 1);
    const CsrEntry* en = c.entries.data() + c.view_base[v];
    uint16_t* sl = c.slot.data() + c.view_base[v];
    for (size_t b = 0; b < nblk; ++b) {
      const size_t r0 = b * STAGE_ROWS, r1 = std::min(R, r0 + STAGE_ROWS);
      const uint32_t k0 = rp[r0], k1 = rp[r1];
      cols.clear();
      for (uint32_t k = k0; k < k1; ++k) cols.push_back(en[k].col);
      std::sort(cols.begin(), cols.end());
      cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
      c.blk_ptr[(size_t)v * (nblk + 1) + b] = (uint32_t)c.blk_tex.size();
      for (uint32_t k = k0; k < k1; ++k)
        sl[k] = (uint16_t)(std::lower_bound(cols.begin(), cols.end(), en[k].col) - cols.begin());
      c.blk_tex.insert(c.blk_tex.end(), cols.begin(), cols.end());
    }
    c.blk_ptr[(size_t)v * (nblk + 1) + nblk] = (uint32_t)c.blk_tex.size();
  }
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
    csr_from_triplets(tb, nlow, v, maps.bwd.row_ptr.data() + (size_t)v * (nlow + 1), pv_bwd[v], low_of_row.data(), wl);
  });
  csr_assemble(maps.fwd, V, pv_fwd);
  csr_assemble(maps.bwd, V, pv_bwd);
  csr_check_coverage(maps.fwd, V, Wg);
  csr_check_coverage(maps.bwd, V, wl);
  build_stage_blocks(maps.fwd, V);
  maps.valid_cells = 0;
  for (int v = 0; v < V; ++v) maps.valid_cells += valid[v];
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }


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
  const int zw = hl > 0 ? wl : Ws;   // width of the image the backward's rows are texels of
  st.zero_segments = (uint32_t)maps.bwd.zseg.size();
  uint64_t zr = 0;
  for (const ZeroSeg& z : maps.bwd.zseg) zr += (uint64_t)std::min<int>(CSR_TILE, zw - (int)(z.id % (uint32_t)zw));
  st.zero_rows = (uint32_t)zr;
}

// Device pointers of a plan from its layout (shared by plan creation and mvwarp_plan_load).
static void bind_pointers(mvwarp_plan* p) {
  char* base = (char*)p->d_base;
  const PlanLayout& L = p->layout;
  p->d_T = (float*)(base + L.T);
  p->d_grid = p->inline_coords ? nullptr : (float*)(base + L.grid);
  p->d_coord_x = (float*)(base + L.coord_x);
  p->d_coord_y = (float*)(base + L.coord_y);
  auto bind = [&](DevCsr& d, const CsrLayout& o) {
    if (o.rows_per_view == 0) return;
    d.rows_per_view = (uint32_t)o.rows_per_view;
    d.max_row = (uint32_t)o.max_row;
    d.row_ptr = (uint32_t*)(base + o.row_ptr); d.row_id = (uint32_t*)(base + o.row_id);
    d.entries = (CsrEntry*)(base + o.entries); d.view_base = (uint64_t*)(base + o.view_base);
    d.tasks = (Task*)(base + o.tasks); d.heavy = (uint32_t*)(base + o.heavy);
    d.heavy_base_dev = (uint32_t*)(base + o.heavy_base);
    if (o.n_blk_tex) {
      d.blk_ptr = (uint32_t*)(base + o.blk_ptr); d.blk_tex = (uint32_t*)(base + o.blk_tex); d.slot = (uint16_t*)(base + o.slot);
      d.max_blk_tex = (uint32_t)o.max_blk_tex;
    }
    d.zseg = o.n_zseg ? (ZeroSeg*)(base + o.zseg) : nullptr;
  };
  bind(p->bwd, L.bwd);
  bind(p->fwd, L.fwd);
  p->d_touched = (uint32_t*)(base + L.touched);
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
    auto carve_csr = [&](const HostCsr& c) {
      CsrLayout o{};
      if (c.rows_per_view == 0) return o;
      o.rows_per_view = c.rows_per_view;
      o.max_row = c.max_row;
      o.row_ptr = carve(c.row_ptr.size() * sizeof(uint32_t));
      o.row_id = carve(c.row_id.size() * sizeof(uint32_t));
      o.entries = carve(c.entries.size() * sizeof(CsrEntry));
      o.view_base = carve(c.view_base.size() * sizeof(uint64_t));
      o.tasks = carve(c.tasks.size() * sizeof(Task));
      o.heavy = carve(c.heavy.size() * sizeof(uint32_t));
      o.heavy_base = carve(c.heavy_base.size() * sizeof(uint32_t));
      if (!c.zseg.empty()) {
        o.n_zseg = c.zseg.size();
        o.zseg = carve(c.zseg.size() * sizeof(ZeroSeg));
      }
      if (!c.blk_tex.empty()) {
        o.n_blk_tex = c.blk_tex.size();
        o.blk_ptr = carve(c.blk_ptr.size() * sizeof(uint32_t));
        o.blk_tex = carve(c.blk_tex.size() * sizeof(uint32_t));
        o.slot = carve(c.slot.size() * sizeof(uint16_t));
        const size_t nb1 = c.blk_ptr.size() / (size_t)V;
        for (int v = 0; v < V; ++v)
          for (size_t b = 0; b + 1 < nb1; ++b)
            o.max_blk_tex = std::max<uint64_t>(o.max_blk_tex, c.blk_ptr[v * nb1 + b + 1] - c.blk_ptr[v * nb1 + b]);
      }
      return o;
    };
    PlanLayout& L = p->layout;
    L.T = carve((size_t)V * 9 * sizeof(float));
    L.grid = carve(T ? 0 : (size_t)V * ncell * 2 * sizeof(float));
    L.coord_x = carve(Wg * sizeof(float));
    L.coord_y = carve(Hg * sizeof(float));
    L.bwd = carve_csr(maps.bwd);
    L.fwd = carve_csr(maps.fwd);
    // touched feature texels (rows of the inverse map with >= 1 entry), ascending per view: what a host-resident feature tensor
    // has to send to the device, and the only rows of the feature gradient that are ever non-zero
    std::vector<uint32_t> touched, touched_base((size_t)V + 1, 0);
    {
      const HostCsr& c = maps.bwd;
      touched.reserve(c.nonempty_rows);
      for (int v = 0; v < V; ++v) {
        const uint32_t* rp = c.row_ptr.data() + (size_t)v * (c.rows_per_view + 1);
        const uint32_t* id = c.row_id.data() + (size_t)v * c.rows_per_view;
        const size_t first = touched.size();
        for (size_t r = 0; r < c.rows_per_view; ++r)
          if (rp[r + 1] > rp[r]) touched.push_back(id[r]);
        std::sort(touched.begin() + first, touched.end());
        touched_base[v + 1] = (uint32_t)touched.size();
      }
    }
    L.touched = carve(touched.size() * sizeof(uint32_t));
    L.touched_base = carve(touched_base.size() * sizeof(uint32_t));
    p->touched_base = touched_base;
    p->d_bytes = off;
    e = cudaMalloc(&p->d_base, p->d_bytes);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(plan)"); }
    p->bwd.task_base = maps.bwd.task_base; p->bwd.heavy_base = maps.bwd.heavy_base; p->bwd.zseg_base = maps.bwd.zseg_base;
    p->fwd.task_base = maps.fwd.task_base; p->fwd.heavy_base = maps.fwd.heavy_base; p->fwd.zseg_base = maps.fwd.zseg_base;
    if (T) p->h_T.assign(T, T + (size_t)V * 9);
    bind_pointers(p);

    auto up = [&](void* dst, const void* src, size_t bytes) {
      if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    };
    auto up_csr = [&](DevCsr& d, const HostCsr& c) {
      if (c.rows_per_view == 0) return;
      up(d.row_ptr, c.row_ptr.data(), c.row_ptr.size() * sizeof(uint32_t));
      up(d.row_id, c.row_id.data(), c.row_id.size() * sizeof(uint32_t));
      up(d.entries, c.entries.data(), c.entries.size() * sizeof(CsrEntry));
      up(d.view_base, c.view_base.data(), c.view_base.size() * sizeof(uint64_t));
      up(d.tasks, c.tasks.data(), c.tasks.size() * sizeof(Task));
      up(d.heavy, c.heavy.data(), c.heavy.size() * sizeof(uint32_t));
      up(d.heavy_base_dev, c.heavy_base.data(), c.heavy_base.size() * sizeof(uint32_t));
      if (!c.zseg.empty()) up(d.zseg, c.zseg.data(), c.zseg.size() * sizeof(ZeroSeg));
      if (!c.blk_tex.empty()) {
        up(d.blk_ptr, c.blk_ptr.data(), c.blk_ptr.size() * sizeof(uint32_t));
        up(d.blk_tex, c.blk_tex.data(), c.blk_tex.size() * sizeof(uint32_t));
        up(d.slot, c.slot.data(), c.slot.size() * sizeof(uint16_t));
      }
    };
    e = cudaMemset(p->d_base, 0, p->d_bytes);  // alignment gaps are part of a saved plan's checksum
    if (T) up(p->d_T, T, (size_t)V * 9 * sizeof(float));
    else up(p->d_grid, grid, (size_t)V * ncell * 2 * sizeof(float));
    up(p->d_coord_x, cx.data(), Wg * sizeof(float));
    up(p->d_coord_y, cy.data(), Hg * sizeof(float));
    up_csr(p->bwd, maps.bwd);
    up_csr(p->fwd, maps.fwd);
    up(p->d_touched, touched.data(), touched.size() * sizeof(uint32_t));
    up((char*)p->d_base + L.touched_base, touched_base.data(), touched_base.size() * sizeof(uint32_t));
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
  } catch (...) {
    set_error("plan_create: unknown C++ exception");
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
  } catch (const std::exception& ex) {  // the header's promise: nothing is ever thrown across the boundary
    set_error(std::string("host_csr: ") + ex.what());
    return MVWARP_EINVAL;
  } catch (...) {
    set_error("host_csr: unknown C++ exception");
    return MVWARP_EINVAL;
  }
}

int mvwarp_host_upsample_windows(const float* T, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg, int32_t* origin_out,
                                 uint32_t* mask_out, float* wx_out, float* wy_out) {
  if (!T || !origin_out || !mask_out || !wx_out || !wy_out) { set_error("host_upsample_windows: null argument"); return MVWARP_EINVAL; }
  int rc = check_shapes(V, Hs, Ws, Hg, Wg);
  if (rc) return rc;
  if (hl <= 0 || wl <= 0 || hl > Hs || wl > Ws || hl > 32767 || wl > 32767) {
    set_error("host_upsample_windows: the low-res size must be positive, at most the up-sampled size and below 32768");
    return MVWARP_EINVAL;
  }
  // exactly what the fused warp -> conv1 producers compute per halo cell (mvwarp_uptaps.h, shared source)
  const float scx = upsample_scale(wl, Ws), scy = upsample_scale(hl, Hs);
  for (int v = 0; v < V; ++v)
    for (int y = 0; y < Hg; ++y)
      for (int x = 0; x < Wg; ++x) {
        float sx, sy;
        norm_src_coord(T + 9 * v, x, y, Hg, Wg, sx, sy);
        const ConvTapUp t = compose_cell(sx, sy, Hs, Ws, hl, wl, scx, scy);
        const size_t i = ((size_t)v * Hg + y) * Wg + x;
        origin_out[2 * i] = t.xy & 0xffff; origin_out[2 * i + 1] = t.xy >> 16;
        mask_out[i] = t.mask;
        for (int k = 0; k < 3; ++k) { wx_out[3 * i + k] = t.wx[k]; wy_out[3 * i + k] = t.wy[k]; }
      }
  return MVWARP_OK;
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

// ---- plan persistence ----------------------------------------------------------------------
// File = PlanFileHeader | bwd task_base[V+1] | bwd heavy_base[V+1] | bwd zseg_base[V+1] | fwd task_base, heavy_base, zseg_base (fused
// up-sample plans only) | matrices [V*9] (inline plans only) | the device allocation, byte for byte.
namespace {
struct PlanFileHeader {
  char magic[8];             // "MVWPLAN2"
  uint32_t header_bytes;     // sizeof(PlanFileHeader): a struct change invalidates old files
  uint32_t schema;           // plan_schema(): bumped by hand when the MEANING of an array changes, mixed with the record sizes
  uint32_t abi_version;      // MVWARP_VERSION
  uint32_t task_budget, task_max_rows, heavy_len, csr_tile;  // work decomposition baked into the tasks
  int32_t V, Hs, Ws, Hg, Wg, hl, wl, upsampled, inline_coords, has_fwd;
  uint64_t d_bytes;
  uint64_t blob_fnv1a;       // FNV-1a 64 of the device blob
  mvw::PlanLayout layout;
  mvwarp_stats_t stats;
};
const char kMagic[8] = {'M', 'V', 'W', 'P', 'L', 'A', 'N', '2'};

// Layout / schema hash: any builder change that keeps the constants but changes what the arrays mean must bump PLAN_SCHEMA, so
// that files cached under $MVWARP_PLAN_DIR by an older build are refused instead of being dereferenced by the gather kernels.
constexpr uint32_t PLAN_SCHEMA = 4;   // 4: zero segments (rows of all-empty tile-row segments belong to no task)
constexpr uint32_t plan_schema() {
  return PLAN_SCHEMA * 0x9E3779B1u ^ (uint32_t)(sizeof(mvw::Task) << 24 | sizeof(mvw::CsrEntry) << 16 | sizeof(mvw::PlanLayout));
}

uint64_t fnv1a64(const unsigned char* p, size_t n) {
  // 8 bytes per step (word-wise FNV variant): a checksum against truncation / bit rot, not a cryptographic hash
  uint64_t h = 1469598103934665603ull;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 1099511628211ull;
  }
  for (; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}

struct FileCloser {
  FILE* f;
  ~FileCloser() { if (f) fclose(f); }
};
}  // namespace

int mvwarp_plan_save(const mvwarp_plan_t* plan, const char* path) {
  if (!plan || !path) { set_error("plan_save: null argument"); return MVWARP_EINVAL; }
  try {
    DeviceGuard guard(plan->device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    std::vector<unsigned char> blob(plan->d_bytes);
    cudaError_t e = cudaMemcpy(blob.data(), plan->d_base, plan->d_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "plan_save: device -> host copy");
    PlanFileHeader h;
    memset((void*)&h, 0, sizeof(h));  // padding bytes are written to the file too
    memcpy(h.magic, kMagic, 8);
    h.header_bytes = sizeof(PlanFileHeader);
    h.schema = plan_schema();
    h.abi_version = MVWARP_VERSION;
    h.task_budget = TASK_BUDGET; h.task_max_rows = TASK_MAX_ROWS; h.heavy_len = HEAVY_LEN; h.csr_tile = CSR_TILE;
    h.V = plan->V; h.Hs = plan->Hs; h.Ws = plan->Ws; h.Hg = plan->Hg; h.Wg = plan->Wg; h.hl = plan->hl; h.wl = plan->wl;
    h.upsampled = plan->upsampled; h.inline_coords = plan->inline_coords;
    h.has_fwd = plan->fwd.rows_per_view ? 1 : 0;
    h.d_bytes = plan->d_bytes;
    h.blob_fnv1a = fnv1a64(blob.data(), blob.size());
    h.layout = plan->layout;
    h.stats = plan->stats;
    const std::string tmp = std::string(path) + ".tmp";
    {
      FileCloser fc{fopen(tmp.c_str(), "wb")};
      if (!fc.f) { set_error(std::string("plan_save: cannot open ") + tmp); return MVWARP_EINVAL; }
      bool ok = fwrite(&h, sizeof(h), 1, fc.f) == 1;
      auto put = [&](const std::vector<uint32_t>& v) {
        ok = ok && v.size() == (size_t)plan->V + 1 && fwrite(v.data(), sizeof(uint32_t), v.size(), fc.f) == v.size();
      };
      put(plan->bwd.task_base); put(plan->bwd.heavy_base); put(plan->bwd.zseg_base);
      if (h.has_fwd) { put(plan->fwd.task_base); put(plan->fwd.heavy_base); put(plan->fwd.zseg_base); }
      if (h.inline_coords) ok = ok && plan->h_T.size() == (size_t)plan->V * 9 &&
                                fwrite(plan->h_T.data(), sizeof(float), plan->h_T.size(), fc.f) == plan->h_T.size();
      ok = ok && fwrite(blob.data(), 1, blob.size(), fc.f) == blob.size();
      ok = ok && fflush(fc.f) == 0;
      if (!ok) { set_error(std::string("plan_save: write to ") + tmp + " failed"); remove(tmp.c_str()); return MVWARP_EINVAL; }
    }
    if (rename(tmp.c_str(), path) != 0) {  // atomic replace: concurrent ranks may save the same rig
      set_error(std::string("plan_save: cannot rename to ") + path);
      remove(tmp.c_str());
      return MVWARP_EINVAL;
    }
    return MVWARP_OK;
  } catch (const std::bad_alloc&) {
    set_error("plan_save: host allocation failed");
    return MVWARP_ENOMEM;
  } catch (const std::exception& ex) {  // the header's promise: nothing is ever thrown across the boundary
    set_error(std::string("plan_save: ") + ex.what());
    return MVWARP_EINVAL;
  } catch (...) {
    set_error("plan_save: unknown C++ exception");
    return MVWARP_EINVAL;
  }
}

int mvwarp_plan_load(const char* path, int device, mvwarp_plan_t** plan_out) {
  if (!path || !plan_out) { set_error("plan_load: null argument"); return MVWARP_EINVAL; }
  *plan_out = nullptr;
  try {
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) { set_error(std::string("plan_load: cannot open ") + path); return MVWARP_EINVAL; }
    PlanFileHeader h;
    if (fread(&h, sizeof(h), 1, fc.f) != 1 || memcmp(h.magic, kMagic, 8) != 0) {
      set_error(std::string("plan_load: ") + path + " is not a plan file");
      return MVWARP_EINVAL;
    }
    if (h.header_bytes != sizeof(PlanFileHeader) || h.schema != plan_schema() || h.abi_version != MVWARP_VERSION || h.task_budget != TASK_BUDGET ||
        h.task_max_rows != TASK_MAX_ROWS || h.heavy_len != HEAVY_LEN || h.csr_tile != CSR_TILE) {
      set_error(std::string("plan_load: ") + path + " was written by an incompatible library build; rebuild the plan");
      return MVWARP_EUNSUPPORTED;
    }
    int rc = check_shapes(h.V, h.Hs, h.Ws, h.Hg, h.Wg);
    if (rc) return rc;
    if (h.d_bytes == 0 || h.d_bytes > (1ull << 40)) { set_error("plan_load: implausible plan size"); return MVWARP_EINVAL; }
    std::unique_ptr<mvwarp_plan> p(new mvwarp_plan());
    p->device = device;
    p->V = h.V; p->Hs = h.Hs; p->Ws = h.Ws; p->Hg = h.Hg; p->Wg = h.Wg; p->hl = h.hl; p->wl = h.wl;
    p->upsampled = h.upsampled; p->inline_coords = h.inline_coords;
    p->layout = h.layout;
    p->stats = h.stats;
    p->d_bytes = h.d_bytes;
    bool ok = true;
    auto get = [&](std::vector<uint32_t>& v) {
      v.resize((size_t)h.V + 1);
      ok = ok && fread(v.data(), sizeof(uint32_t), v.size(), fc.f) == v.size();
    };
    get(p->bwd.task_base); get(p->bwd.heavy_base); get(p->bwd.zseg_base);
    if (h.has_fwd) { get(p->fwd.task_base); get(p->fwd.heavy_base); get(p->fwd.zseg_base); }
    if (h.inline_coords) {
      p->h_T.resize((size_t)h.V * 9);
      ok = ok && fread(p->h_T.data(), sizeof(float), p->h_T.size(), fc.f) == p->h_T.size();
    }
    std::vector<unsigned char> blob(h.d_bytes);
    ok = ok && fread(blob.data(), 1, blob.size(), fc.f) == blob.size();
    if (!ok) { set_error(std::string("plan_load: ") + path + " is truncated"); return MVWARP_EINVAL; }
    if (fnv1a64(blob.data(), blob.size()) != h.blob_fnv1a) {
      set_error(std::string("plan_load: checksum mismatch in ") + path);
      return MVWARP_EINVAL;
    }
    // every array must lie inside the blob (a corrupt header must not become a wild device pointer)
    const uint64_t nrows_b = h.layout.bwd.rows_per_view, nrows_f = h.layout.fwd.rows_per_view;
    auto inside = [&](uint64_t off, uint64_t bytes) { return off <= h.d_bytes && bytes <= h.d_bytes - off; };
    auto csr_ok = [&](const CsrLayout& o, const std::vector<uint32_t>& tb, const std::vector<uint32_t>& hb, const std::vector<uint32_t>& zb) {
      if (o.rows_per_view == 0) return true;
      if (zb.back() != o.n_zseg || (o.n_zseg && !inside(o.zseg, o.n_zseg * sizeof(ZeroSeg)))) return false;
      return inside(o.row_ptr, (uint64_t)h.V * (o.rows_per_view + 1) * 4) && inside(o.row_id, (uint64_t)h.V * o.rows_per_view * 4) &&
             inside(o.view_base, ((uint64_t)h.V + 1) * 8) && inside(o.tasks, (uint64_t)tb.back() * sizeof(Task)) &&
             inside(o.heavy, (uint64_t)hb.back() * 4) && inside(o.heavy_base, ((uint64_t)h.V + 1) * 4) && inside(o.entries, 0);
    };
    if (!inside(h.layout.T, (uint64_t)h.V * 36) || !inside(h.layout.coord_x, (uint64_t)h.Wg * 4) ||
        !inside(h.layout.coord_y, (uint64_t)h.Hg * 4) || nrows_b == 0 || (h.has_fwd != 0) != (nrows_f != 0) ||
        !csr_ok(h.layout.bwd, p->bwd.task_base, p->bwd.heavy_base, p->bwd.zseg_base) ||
        (h.has_fwd && !csr_ok(h.layout.fwd, p->fwd.task_base, p->fwd.heavy_base, p->fwd.zseg_base)) ||
        (!h.inline_coords && !inside(h.layout.grid, (uint64_t)h.V * h.Hg * h.Wg * 8))) {
      set_error(std::string("plan_load: inconsistent layout in ") + path);
      return MVWARP_EINVAL;
    }
    // ... and every INDEX the gather kernels will dereference must be in range: a well-formed file with a valid checksum but
    // wrong contents (a foreign tool, a shared cache directory) must be refused here, not found out by an out-of-bounds read.
    auto contents_ok = [&](const CsrLayout& o, const std::vector<uint32_t>& tb, const std::vector<uint32_t>& hb, const std::vector<uint32_t>& zb,
                           uint64_t n_dst, uint64_t dst_width, uint64_t n_cols) {
      if (o.rows_per_view == 0) return true;
      const uint64_t R = o.rows_per_view;
      if (R > n_dst) return false;
      const uint64_t* vb = reinterpret_cast<const uint64_t*>(blob.data() + o.view_base);
      if (vb[0] != 0) return false;
      for (int v = 0; v < h.V; ++v) if (vb[v + 1] < vb[v]) return false;
      const uint64_t nnz = vb[h.V];
      if (!inside(o.entries, nnz * sizeof(CsrEntry))) return false;
      const uint32_t* rp = reinterpret_cast<const uint32_t*>(blob.data() + o.row_ptr);
      const uint32_t* rid = reinterpret_cast<const uint32_t*>(blob.data() + o.row_id);
      const CsrEntry* en = reinterpret_cast<const CsrEntry*>(blob.data() + o.entries);
      for (int v = 0; v < h.V; ++v) {
        const uint32_t* r = rp + (uint64_t)v * (R + 1);
        if (r[0] != 0 || r[R] != vb[v + 1] - vb[v]) return false;
        for (uint64_t i = 0; i < R; ++i) if (r[i + 1] < r[i]) return false;
        for (uint64_t i = 0; i < R; ++i) if (rid[(uint64_t)v * R + i] >= n_dst) return false;
      }
      for (uint64_t k = 0; k < nnz; ++k) if (en[k].col >= n_cols) return false;
      if (tb[0] != 0 || hb[0] != 0) return false;
      const Task* tasks = reinterpret_cast<const Task*>(blob.data() + o.tasks);
      const uint32_t* heavy = reinterpret_cast<const uint32_t*>(blob.data() + o.heavy);
      const uint32_t* hbd = reinterpret_cast<const uint32_t*>(blob.data() + o.heavy_base);
      for (int v = 0; v < h.V; ++v) {
        if (tb[v + 1] < tb[v] || hb[v + 1] < hb[v] || hbd[v] != hb[v]) return false;
        for (uint32_t t = tb[v]; t < tb[v + 1]; ++t)
          if (tasks[t].view != v || tasks[t].nrows == 0 || tasks[t].nrows > TASK_MAX_ROWS || (uint64_t)tasks[t].row0 + tasks[t].nrows > R) return false;
        for (uint32_t i = hb[v]; i < hb[v + 1]; ++i) if (heavy[i] >= R) return false;
      }
      if (hbd[h.V] != hb[h.V]) return false;
      if (zb[0] != 0) return false;
      for (int v = 0; v < h.V; ++v) if (zb[v + 1] < zb[v]) return false;
      if (o.n_zseg) {   // zero segments: whole tile-row segments inside the destination image, filed under their own view
        const ZeroSeg* zs = reinterpret_cast<const ZeroSeg*>(blob.data() + o.zseg);
        for (int v = 0; v < h.V; ++v)
          for (uint32_t i = zb[v]; i < zb[v + 1]; ++i) {
            const uint64_t x = zs[i].id % dst_width;
            if (zs[i].view != (uint32_t)v || zs[i].id >= n_dst || x % CSR_TILE != 0) return false;
          }
      }
      if (o.n_blk_tex) {   // staging blocks: block lists inside the blob, columns in range, every entry's slot inside its block
        const uint64_t nblk = (R + STAGE_ROWS - 1) / STAGE_ROWS;
        if (!inside(o.blk_ptr, (uint64_t)h.V * (nblk + 1) * 4) || !inside(o.blk_tex, o.n_blk_tex * 4) || !inside(o.slot, nnz * 2)) return false;
        const uint32_t* bp = reinterpret_cast<const uint32_t*>(blob.data() + o.blk_ptr);
        const uint32_t* bt = reinterpret_cast<const uint32_t*>(blob.data() + o.blk_tex);
        const uint16_t* sl = reinterpret_cast<const uint16_t*>(blob.data() + o.slot);
        for (uint64_t i = 0; i < o.n_blk_tex; ++i) if (bt[i] >= n_cols) return false;
        uint32_t prev = 0;
        for (int v = 0; v < h.V; ++v) {
          const uint32_t* r = rp + (uint64_t)v * (R + 1);
          for (uint64_t b = 0; b < nblk; ++b) {
            const uint32_t t0 = bp[(uint64_t)v * (nblk + 1) + b], t1 = bp[(uint64_t)v * (nblk + 1) + b + 1];
            if (t0 < prev || t1 < t0 || t1 > o.n_blk_tex || t1 - t0 > o.max_blk_tex) return false;
            prev = t0;
            const uint64_t r0 = b * STAGE_ROWS, r1 = std::min<uint64_t>(R, r0 + STAGE_ROWS);
            for (uint64_t k = vb[v] + r[r0]; k < vb[v] + r[r1]; ++k) if (sl[k] >= t1 - t0) return false;
          }
        }
      }
      return true;
    };
    const uint64_t n_cells = (uint64_t)h.Hg * h.Wg, n_src = h.upsampled ? (uint64_t)h.hl * h.wl : (uint64_t)h.Hs * h.Ws;
    {
      bool tok = inside(h.layout.touched_base, ((uint64_t)h.V + 1) * 4);
      if (tok) {
        const uint32_t* tb = reinterpret_cast<const uint32_t*>(blob.data() + h.layout.touched_base);
        tok = tb[0] == 0;
        for (int v = 0; v < h.V && tok; ++v) tok = tb[v + 1] >= tb[v];
        tok = tok && inside(h.layout.touched, (uint64_t)tb[h.V] * 4);
        if (tok) {
          const uint32_t* t = reinterpret_cast<const uint32_t*>(blob.data() + h.layout.touched);
          for (uint32_t i = 0; i < tb[h.V] && tok; ++i) tok = t[i] < n_src;
          p->touched_base.assign(tb, tb + h.V + 1);
        }
      }
      if (!tok) { set_error(std::string("plan_load: bad touched-texel list in ") + path); return MVWARP_EINVAL; }
    }
    const uint64_t w_src = h.upsampled ? (uint64_t)h.wl : (uint64_t)h.Ws;
    if ((h.upsampled && (h.hl <= 0 || h.wl <= 0)) ||
        !contents_ok(h.layout.bwd, p->bwd.task_base, p->bwd.heavy_base, p->bwd.zseg_base, n_src, w_src, n_cells) ||
        (h.has_fwd && !contents_ok(h.layout.fwd, p->fwd.task_base, p->fwd.heavy_base, p->fwd.zseg_base, n_cells, (uint64_t)h.Wg, n_src))) {
      set_error(std::string("plan_load: index out of range in ") + path + " (corrupt or foreign contents)");
      return MVWARP_EINVAL;
    }
    DeviceGuard guard(device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    p->num_sms = prop.multiProcessorCount;
    e = cudaMalloc(&p->d_base, p->d_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(plan)");
    e = cudaMemcpy(p->d_base, blob.data(), p->d_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->d_base); return cuda_fail(e, "plan_load: upload"); }
    bind_pointers(p.get());
    p->stats.device_bytes = p->d_bytes;
    *plan_out = p.release();
    return MVWARP_OK;
  } catch (const std::bad_alloc&) {
    set_error("plan_load: host allocation failed");
    return MVWARP_ENOMEM;
  } catch (const std::exception& ex) {  // the header's promise: nothing is ever thrown across the boundary
    set_error(std::string("plan_load: ") + ex.what());
    return MVWARP_EINVAL;
  } catch (...) {
    set_error("plan_load: unknown C++ exception");
    return MVWARP_EINVAL;
  }
}

int mvwarp_plan_matrices(const mvwarp_plan_t* plan, float* T_out) {
  if (!plan || !T_out) { set_error("plan_matrices: null argument"); return MVWARP_EINVAL; }
  if (!plan->inline_coords || plan->h_T.size() != (size_t)plan->V * 9) {
    set_error("plan_matrices: the plan was built from an explicit grid, it has no matrices");
    return MVWARP_EUNSUPPORTED;
  }
  memcpy(T_out, plan->h_T.data(), plan->h_T.size() * sizeof(float));
  return MVWARP_OK;
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
  if (B <= 0 || C <= 0 || nviews <= 0 || view0 < 0 || view0 > plan->V || nviews > plan->V - view0) {
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

static int list_call(const mvwarp_plan_t* plan, int view0, int nviews, const void* const* views, const void* concat,
                     const int64_t feat_strides[4], const int64_t out_strides[4], int out_chan0, int coord_chan, int B, int C,
                     int dtype, void* stream, bool forward) {
  const char* who = forward ? "warp_views_fwd_list" : "warp_views_bwd_list";
  if (!plan || !views || !concat || !feat_strides || !out_strides) { set_error(std::string(who) + ": null argument"); return MVWARP_EINVAL; }
  if (nviews <= 0) { set_error(std::string(who) + ": bad B/C/view range"); return MVWARP_EINVAL; }
  if (nviews > MVWARP_MAX_LIST_VIEWS) {
    set_error(std::string(who) + ": more than MVWARP_MAX_LIST_VIEWS views in one call (split the list)");
    return MVWARP_EUNSUPPORTED;
  }
  for (int v = 0; v < nviews; ++v)
    if (!views[v]) { set_error(std::string(who) + ": null view pointer"); return MVWARP_EINVAL; }
  const int64_t fs5[5] = {feat_strides[0], 0, feat_strides[1], feat_strides[2], feat_strides[3]};
  int rc = check_call(plan, view0, nviews, views[0], concat, fs5, out_strides, B, C, dtype, who);
  if (rc) return rc;
  if (out_chan0 < 0) { set_error(std::string(who) + ": negative out_chan0"); return MVWARP_EINVAL; }
  DeviceGuard guard(plan->device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  LaunchArgs a{};
  a.plan = plan; a.view0 = view0; a.nviews = nviews; a.B = B; a.C = C; a.dtype = dtype;
  a.feats_or_grad_out = forward ? views[0] : concat;
  a.out_or_grad_feats = forward ? const_cast<void*>(concat) : const_cast<void*>(views[0]);
  memcpy(a.fs, fs5, sizeof(a.fs));
  memcpy(a.os, out_strides, sizeof(a.os));
  a.out_chan0 = out_chan0; a.coord_chan = forward ? coord_chan : -1; a.stream = (cudaStream_t)stream;
  a.view_ptrs = views;
  return forward ? launch_fwd(a) : launch_bwd(a);
}

int mvwarp_warp_views_fwd_list(const mvwarp_plan_t* plan, int view0, int nviews, const void* const* feats,
                               const int64_t feat_strides[4], void* out, const int64_t out_strides[4], int out_chan0,
                               int coord_chan, int B, int C, int dtype, void* stream) {
  return list_call(plan, view0, nviews, feats, out, feat_strides, out_strides, out_chan0, coord_chan, B, C, dtype, stream, true);
}

int mvwarp_warp_views_bwd_list(const mvwarp_plan_t* plan, int view0, int nviews, const void* grad_out,
                               const int64_t out_strides[4], int out_chan0, void* const* grad_feats,
                               const int64_t feat_strides[4], int B, int C, int dtype, void* stream) {
  return list_call(plan, view0, nviews, grad_feats, grad_out, feat_strides, out_strides, out_chan0, -1, B, C, dtype, stream, false);
}

}  // extern "C"
