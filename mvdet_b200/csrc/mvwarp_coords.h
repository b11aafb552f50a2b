// mvwarp_coords.h -- the per-cell coordinate chain, shared by host (plan builder) and device (kernels).
//
// Restates, operation for operation in IEEE fp32 with explicit rounding points, what the CPU
// oracle's torch ops compute between the 3x3 matrix `src_norm_trans_dst_norm` and the four bilinear
// taps:
//   create_meshgrid            xs = (i / (Wg-1) - 0.5) * 2                       (kornia utils/grid.py)
//   transform_points / bmm     q_j = fma(1, T[j][2], fma(gy, T[j][1], gx*T[j][0]))   sequential-K FMA sgemm (MKL), or, when
//                              torch.bmm takes its small-matrix path (9*Wg < 400), the unfused (gx*T0 + gy*T1) + T2
//   convert_points_from_homogeneous   scale = |z| > 1e-8 ? 1/(z + 1e-8) : 1 ; (sx, sy) = scale * (q0, q1)
//   grid_sampler unnormalize (align_corners=True, ATen CPU kernel)  ix = (sx + 1) * ((Ws-1)/2)
//   bilinear weights (ATen CPU kernel)  w = ix - floor(ix); e = 1 - w; n = iy - floor(iy); s = 1 - n
//                                        nw = s*e, ne = s*w, sw = n*e, se = n*w
// Every operation is a single correctly-rounded fp32 op on both sides (device: __f*_rn intrinsics,
// never contracted; host: compiled with -ffp-contract=off, fmaf where the oracle fuses), so host
// CSR weights, device forward weights and the oracle's weights are bit-identical.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define MVW_HD __host__ __device__ __forceinline__
#define MVW_MUL(a, b) __fmul_rn((a), (b))
#define MVW_ADD(a, b) __fadd_rn((a), (b))
#define MVW_SUB(a, b) __fsub_rn((a), (b))
#define MVW_DIV(a, b) __fdiv_rn((a), (b))
#define MVW_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#if defined(__CUDACC__)
#define MVW_HD __host__ __device__ inline
#else
#define MVW_HD inline
#endif
#define MVW_MUL(a, b) ((a) * (b))
#define MVW_ADD(a, b) ((a) + (b))
#define MVW_SUB(a, b) ((a) - (b))
#define MVW_DIV(a, b) ((a) / (b))
#define MVW_FMA(a, b, c) fmaf((a), (b), (c))
#endif

namespace mvw {

struct Taps {
  float w[4];     // nw, ne, sw, se  (weights; not masked)
  int x0, y0;     // north-west texel (may be out of range)
  unsigned mask;  // bit k set: tap k is inside the source image
};

// Normalised destination coordinate of pixel index i on an axis of `size` pixels.
MVW_HD float mesh_coord(int i, int size) {
  return MVW_MUL(MVW_SUB(MVW_DIV((float)i, (float)(size - 1)), 0.5f), 2.0f);
}

// torch.bmm([Hg, Wg, 3] x [Hg, 3, 3]) on CPU: at::native's bmm takes a scalar small-matrix kernel (plain
// multiply-adds, no FMA) when contraction * rows * cols = 3 * Wg * 3 < 400, and MKL's sgemm (FMA chain) otherwise.
// Measured: each form is 100 % bit-equal to torch on its side of the threshold (Wg = 44 / 45) and 77-99 % on the other.
MVW_HD bool bmm_is_fused(int Wg) { return 9 * Wg >= 400; }

MVW_HD float dot3(float gx, float gy, const float* __restrict__ r, bool fused) {
  return fused ? MVW_ADD(MVW_FMA(gy, r[1], MVW_MUL(gx, r[0])), r[2])
               : MVW_ADD(MVW_ADD(MVW_MUL(gx, r[0]), MVW_MUL(gy, r[1])), r[2]);
}

// Normalised source coordinates (what kornia hands to grid_sample) of grid cell (x, y).
MVW_HD void norm_src_coord(const float* __restrict__ T, int x, int y, int Hg, int Wg, float& sx, float& sy) {
  const float gx = mesh_coord(x, Wg), gy = mesh_coord(y, Hg);
  const bool fused = bmm_is_fused(Wg);
  const float q0 = dot3(gx, gy, T, fused);
  const float q1 = dot3(gx, gy, T + 3, fused);
  const float z = dot3(gx, gy, T + 6, fused);
  const float eps = 1e-8f;
  const float scale = (fabsf(z) > eps) ? MVW_DIV(1.0f, MVW_ADD(z, eps)) : 1.0f;
  sx = MVW_MUL(scale, q0);
  sy = MVW_MUL(scale, q1);
}

// Bilinear taps of normalised source coordinate (sx, sy) in an Hs x Ws image, align_corners=True, zeros padding.
MVW_HD Taps make_taps(float sx, float sy, int Hs, int Ws) {
  Taps t;
  const float ix = MVW_MUL(MVW_ADD(sx, 1.0f), MVW_DIV((float)(Ws - 1), 2.0f));
  const float iy = MVW_MUL(MVW_ADD(sy, 1.0f), MVW_DIV((float)(Hs - 1), 2.0f));
  const float xw = floorf(ix), yn = floorf(iy);
  const float w = MVW_SUB(ix, xw), e = MVW_SUB(1.0f, w);
  const float n = MVW_SUB(iy, yn), s = MVW_SUB(1.0f, n);
  t.w[0] = MVW_MUL(s, e);
  t.w[1] = MVW_MUL(s, w);
  t.w[2] = MVW_MUL(n, e);
  t.w[3] = MVW_MUL(n, w);
  // Bounds are tested in float (as ATen does) so NaN / huge coordinates simply fail every test.
  const bool wx = (xw >= 0.0f) && (xw <= (float)(Ws - 1));
  const bool ex = (xw >= -1.0f) && (xw <= (float)(Ws - 2));
  const bool ny = (yn >= 0.0f) && (yn <= (float)(Hs - 1));
  const bool sy_ = (yn >= -1.0f) && (yn <= (float)(Hs - 2));
  t.mask = (unsigned)(wx && ny) | ((unsigned)(ex && ny) << 1) | ((unsigned)(wx && sy_) << 2) | ((unsigned)(ex && sy_) << 3);
  // Only meaningful when some tap is in range; clamp so the int conversion is defined.
  const float cx = fminf(fmaxf(xw, -2.0f), (float)Ws), cy = fminf(fmaxf(yn, -2.0f), (float)Hs);
  t.x0 = (xw == xw) ? (int)cx : -2;
  t.y0 = (yn == yn) ? (int)cy : -2;
  return t;
}

// coord_map channel value (persp_trans_detector.py:103-108): float64 i/(size-1)*2-1 rounded once to fp32.
MVW_HD float coord_map_value(int i, int size) { return (float)((double)i / (double)(size - 1) * 2.0 - 1.0); }

}  // namespace mvw
