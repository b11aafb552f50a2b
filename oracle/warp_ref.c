/* warp_ref.c -- plain-C restatement of the perspective-warp hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independent implementation of the CPU oracle (oracle/kornia_warp.py), without torch:
 * scalar loops that follow, step by step, what the reference call site
 *   /root/reference/multiview_detector/models/persp_trans_detector.py:68-69,77
 * computes through kornia==0.6.11 (requirements.txt:1; not vendored) and ATen's CPU
 * grid_sampler_2d (bilinear, zeros padding, align_corners=True).  PARITY UNPINNED: the reference
 * holds no golden vectors for this path; tests/test_oracle_c.py pins this file against the torch
 * oracle bit-for-bit instead.  Only tests/ may load the library built from this file.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: every rounding below is explicit)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* kornia create_meshgrid(normalized_coordinates=True): (i / (n-1) - 0.5) * 2 in fp32 */
static float mesh(int i, int n) { return ((float)i / (float)(n - 1) - 0.5f) * 2.0f; }

/* kornia transform_points + convert_points_from_homogeneous for one camera.
 * T: 3x3 row-major fp32 `src_norm_trans_dst_norm`; grid: [Hg,Wg,2].
 * torch.bmm (K=3) on CPU: MKL sgemm accumulates sequentially with FMA from a zero accumulator; below
 * 3*Wg*3 < 400 at::native uses its scalar small-matrix kernel instead (plain multiply-adds). */
void warp_ref_grid(const float* T, int Hg, int Wg, float* grid) {
  for (int y = 0; y < Hg; ++y)
    for (int x = 0; x < Wg; ++x) {
      const float p[3] = {mesh(x, Wg), mesh(y, Hg), 1.0f};
      float q[3];
      for (int j = 0; j < 3; ++j) {
        float acc = p[0] * T[3 * j];
        if (9 * Wg >= 400) {
          acc = fmaf(p[1], T[3 * j + 1], acc);
          acc = fmaf(p[2], T[3 * j + 2], acc);
        } else {
          acc = acc + p[1] * T[3 * j + 1];
          acc = acc + p[2] * T[3 * j + 2];
        }
        q[j] = acc;
      }
      const float eps = 1e-8f;
      const float scale = fabsf(q[2]) > eps ? 1.0f / (q[2] + eps) : 1.0f;
      grid[(y * Wg + x) * 2 + 0] = scale * q[0];
      grid[(y * Wg + x) * 2 + 1] = scale * q[1];
    }
}

typedef struct { float w[4]; long off[4]; int ok[4]; } taps_t;

/* ATen grid_sampler_2d CPU kernel, bilinear / zeros / align_corners=True */
static taps_t taps_of(float gx, float gy, int Hs, int Ws) {
  taps_t t;
  const float ix = (gx + 1.0f) * ((float)(Ws - 1) / 2.0f);
  const float iy = (gy + 1.0f) * ((float)(Hs - 1) / 2.0f);
  const float x_w = floorf(ix), y_n = floorf(iy);
  const float w = ix - x_w, e = 1.0f - w, n = iy - y_n, s = 1.0f - n;
  t.w[0] = s * e; t.w[1] = s * w; t.w[2] = n * e; t.w[3] = n * w;
  const float xs[4] = {x_w, x_w + 1.0f, x_w, x_w + 1.0f}, ys[4] = {y_n, y_n, y_n + 1.0f, y_n + 1.0f};
  for (int k = 0; k < 4; ++k) {
    t.ok[k] = xs[k] > -1.0f && xs[k] < (float)Ws && ys[k] > -1.0f && ys[k] < (float)Hs;
    t.off[k] = t.ok[k] ? (long)ys[k] * Ws + (long)xs[k] : 0;
  }
  return t;
}

/* src [N,C,Hs,Ws] NCHW, grid [Hg,Wg,2] shared by the batch, out [N,C,Hg,Wg] (one camera).
 * Blend order as compiled in ATen: ((nw*a + ne*b) + sw*c) + se*d with fused adds. */
void warp_ref_forward(const float* src, int N, int C, int Hs, int Ws, const float* grid, int Hg, int Wg, float* out) {
  for (int cell = 0; cell < Hg * Wg; ++cell) {
    const taps_t t = taps_of(grid[2 * cell], grid[2 * cell + 1], Hs, Ws);
    for (int nc = 0; nc < N * C; ++nc) {
      const float* plane = src + (long)nc * Hs * Ws;
      float v[4];
      for (int k = 0; k < 4; ++k) v[k] = t.ok[k] ? plane[t.off[k]] : 0.0f;
      float acc = v[0] * t.w[0];
      acc = fmaf(v[1], t.w[1], acc);
      acc = fmaf(v[2], t.w[2], acc);
      acc = fmaf(v[3], t.w[3], acc);
      out[(long)nc * Hg * Wg + cell] = acc;
    }
  }
}

/* grid_sampler_2d_backward w.r.t. the input: scatter-add g*w to the four taps, cells in row-major order. */
void warp_ref_backward(const float* grad_out, int N, int C, int Hs, int Ws, const float* grid, int Hg, int Wg,
                       float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)N * C * Hs * Ws);
  for (int cell = 0; cell < Hg * Wg; ++cell) {
    const taps_t t = taps_of(grid[2 * cell], grid[2 * cell + 1], Hs, Ws);
    for (int nc = 0; nc < N * C; ++nc) {
      const float g = grad_out[(long)nc * Hg * Wg + cell];
      float* plane = grad_in + (long)nc * Hs * Ws;
      for (int k = 0; k < 4; ++k)
        if (t.ok[k]) plane[t.off[k]] += g * t.w[k];
    }
  }
}

/* persp_trans_detector.py:103-108 create_coord_map: float64 i/(n-1)*2-1, then .float() */
void warp_ref_coord_map(int Hg, int Wg, float* out /* [2,Hg,Wg] */) {
  for (int y = 0; y < Hg; ++y)
    for (int x = 0; x < Wg; ++x) {
      out[y * Wg + x] = (float)((double)x / (double)(Wg - 1) * 2.0 - 1.0);
      out[Hg * Wg + y * Wg + x] = (float)((double)y / (double)(Hg - 1) * 2.0 - 1.0);
    }
}
