// mvwarp_uptaps.h -- warp(bilinear up-sample(low-res map)) composed per grid cell into ONE 3 x 3 low-res window.
//
// The reference warps F.interpolate(feat, upsample_shape, mode='bilinear') (persp_trans_detector.py:65, 69).  Both steps are
// bilinear and separable:
//   * grid_sample's four high-res taps of a cell are (x0, x0+1) x (y0, y0+1) with weights s*e, s*w, n*e, n*w (mvwarp_coords.h
//     make_taps) and its zero-padding mask is an x flag times a y flag;
//   * a high-res column X is the two low-res columns (i0, i1) with weights (l0, l1) -- ATen upsample_bilinear2d with
//     align_corners=False, the fp32 formulas of mvwarp_plan.cu `upsample_axis`:  real = max(scale * (X + 0.5) - 0.5, 0),
//     i0 = min(int(real), in - 1), l1 = real - i0, i1 = i0 + (i0 < in - 1), l0 = 1 - l1,  scale = in / out.
// Two neighbouring high-res columns fall into a window of at most THREE low-res columns when scale <= 1 (an up-sample), so a
// cell is  sum_{j,i < 3} wy[j] * wx[i] * low[oy + j][ox + i]:  six weights, an origin and a 9-bit mask of the live slots.
// Shared by the fused warp -> conv1 kernel (device) and the CPU test harness (host, tests/uptaps_harness.cpp), which pins it
// against the plan builder's composed CSR (mvwarp_host_csr_upsampled).  Every operation is a single rounded fp32 op on both sides.
#pragma once
#include "mvwarp_coords.h"

namespace mvw {

struct ConvTapUp {
  float wx[3], wy[3];
  int xy;          // window origin: ox | oy << 16 (low-res texel; in range whenever a slot is live)
  unsigned mask;   // bit j*3 + i: slot (row j, column i) has a non-zero weight
};

struct UpAxis {
  int i0, i1;
  float l0, l1;
};

MVW_HD float upsample_scale(int in, int out) { return MVW_DIV((float)in, (float)out); }

MVW_HD UpAxis up_axis(int d, int in, float scale) {
  float real = MVW_SUB(MVW_MUL(scale, MVW_ADD((float)d, 0.5f)), 0.5f);
  if (real < 0.f) real = 0.f;
  UpAxis a;
  const int i = (int)real;
  a.i0 = i < in - 1 ? i : in - 1;
  const float l1 = MVW_SUB(real, (float)a.i0);
  a.l1 = l1 < 0.f ? 0.f : (l1 > 1.f ? 1.f : l1);
  a.i1 = a.i0 + (a.i0 < in - 1 ? 1 : 0);
  a.l0 = MVW_SUB(1.f, a.l1);
  return a;
}

// One axis of a cell: high-res taps x0 (weight e, live if in0) and x0 + 1 (weight w, live if in1) -> low-res window.
MVW_HD void compose_axis(int x0, float e, float w, bool in0, bool in1, int in, float scale, int& origin, float out[3]) {
  out[0] = out[1] = out[2] = 0.f;
  origin = 0;
  UpAxis a{0, 0, 0.f, 0.f}, b{0, 0, 0.f, 0.f};
  if (in0) a = up_axis(x0, in, scale);
  if (in1) b = up_axis(x0 + 1, in, scale);
  if (in0) origin = a.i0;
  else if (in1) origin = b.i0;
  // i - origin is 0, 1 or 2 for an up-sampling scale (callers check scale <= 1); anything else cannot happen and is dropped
#define MVW_UP_ADD(i, v)                                   \
  do {                                                     \
    const int o_ = (i) - origin;                           \
    const float v_ = (v);                                  \
    if (o_ == 0) out[0] = MVW_ADD(out[0], v_);             \
    else if (o_ == 1) out[1] = MVW_ADD(out[1], v_);        \
    else if (o_ == 2) out[2] = MVW_ADD(out[2], v_);        \
  } while (0)
  if (in0) { MVW_UP_ADD(a.i0, MVW_MUL(e, a.l0)); MVW_UP_ADD(a.i1, MVW_MUL(e, a.l1)); }
  if (in1) { MVW_UP_ADD(b.i0, MVW_MUL(w, b.l0)); MVW_UP_ADD(b.i1, MVW_MUL(w, b.l1)); }
#undef MVW_UP_ADD
}

// The composed window of the cell whose normalised source coordinate (what kornia hands to grid_sample) is (sx, sy).
// (Hs, Ws): the up-sampled size the warp addresses; (hl, wl): the low-res map actually read; scx = wl / Ws, scy = hl / Hs.
MVW_HD ConvTapUp compose_cell(float sx, float sy, int Hs, int Ws, int hl, int wl, float scx, float scy) {
  ConvTapUp t;
  t.mask = 0; t.xy = 0;
  t.wx[0] = t.wx[1] = t.wx[2] = t.wy[0] = t.wy[1] = t.wy[2] = 0.f;
  const Taps tp = make_taps(sx, sy, Hs, Ws);          // the four HIGH-RES taps, exactly as the un-fused warp sees them
  if (!tp.mask) return t;
  // their separable factors (make_taps: nw = s*e, ne = s*w, sw = n*e, se = n*w), the same fp32 operations
  const float ix = MVW_MUL(MVW_ADD(sx, 1.0f), MVW_DIV((float)(Ws - 1), 2.0f));
  const float iy = MVW_MUL(MVW_ADD(sy, 1.0f), MVW_DIV((float)(Hs - 1), 2.0f));
  const float w = MVW_SUB(ix, floorf(ix)), e = MVW_SUB(1.0f, w);
  const float n = MVW_SUB(iy, floorf(iy)), s = MVW_SUB(1.0f, n);
  int ox, oy;
  compose_axis(tp.x0, e, w, (tp.mask & 5u) != 0, (tp.mask & 10u) != 0, wl, scx, ox, t.wx);
  compose_axis(tp.y0, s, n, (tp.mask & 3u) != 0, (tp.mask & 12u) != 0, hl, scy, oy, t.wy);
  t.xy = ox | (oy << 16);
  for (int k = 0; k < 9; ++k)
    if (t.wy[k / 3] != 0.f && t.wx[k % 3] != 0.f) t.mask |= 1u << k;
  return t;
}

}  // namespace mvw
