// mvwarp_conv1_bwd.cu -- backward of the fused warp -> map_classifier[0] (+ReLU): the element-wise front end.
//
// The training step (BASELINE.json configs[2]; reached from loss.backward(), /root/reference/multiview_detector/trainer.py:48)
// back-propagates through relu(conv1(warp_views(feats)) + bias).  Before the two convolution gradients can run, the incoming
// gradient has to be masked by the ReLU, rounded to bf16 (the tensor cores' operand type) and summed per output channel for the
// bias gradient.  As separate torch ops that is three passes over [B,512,Hg,Wg] fp32 tensors (4.5 GB of traffic at B=8); this
// kernel does all three in one pass and writes the bf16 gradient straight into the zero-padded buffer the weight-gradient GEMMs
// read (mvdet_b200.conv1.conv1_wgrad_gemm), so no separate copy into it is needed either.
//
//   g16[b, n, y, x]  = bf16( grad_out[b, n, y, x] * (out[b, n, y, x] > 0) )          (out == NULL: no mask)
//   bias_grad[n]     = sum_{b,y,x} of the masked fp32 gradient, in a fixed order (per-CTA partials, then one ordered pass)
#include <cuda_bf16.h>

#include <algorithm>

#include "mvwarp_internal.h"

namespace mvw {

constexpr int RB_THREADS = 256;

// C channels = 8 per thread; the CTA's 256 threads cover (256 * 8 / C) cells per iteration.
__global__ void __launch_bounds__(RB_THREADS) relu_bwd_cast_kernel(const float* __restrict__ go, int64_t goB, int64_t goH, int64_t goW,
                                                                   const float* __restrict__ out, int64_t oB, int64_t oH, int64_t oW,
                                                                   __nv_bfloat16* __restrict__ g, int64_t gB, int64_t gH, int64_t gW,
                                                                   float* __restrict__ partial, int B, int H, int W, int C) {
  const int lanes_per_cell = C / 8;                       // threads that share one cell
  const int cells_per_iter = RB_THREADS / lanes_per_cell;
  const int sub = threadIdx.x / lanes_per_cell, c0 = (threadIdx.x % lanes_per_cell) * 8;
  const int64_t ncell = (int64_t)B * H * W;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int64_t cell = (int64_t)blockIdx.x * cells_per_iter + sub; cell < ncell; cell += (int64_t)gridDim.x * cells_per_iter) {
    const int x = (int)(cell % W), y = (int)((cell / W) % H), b = (int)(cell / ((int64_t)W * H));
    const float4* pg = reinterpret_cast<const float4*>(go + b * goB + y * goH + x * goW + c0);
    float4 a0 = __ldcs(pg), a1 = __ldcs(pg + 1);
    if (out) {
      const float4* po = reinterpret_cast<const float4*>(out + b * oB + y * oH + x * oW + c0);
      const float4 m0 = __ldcs(po), m1 = __ldcs(po + 1);
      a0.x = m0.x > 0.f ? a0.x : 0.f; a0.y = m0.y > 0.f ? a0.y : 0.f; a0.z = m0.z > 0.f ? a0.z : 0.f; a0.w = m0.w > 0.f ? a0.w : 0.f;
      a1.x = m1.x > 0.f ? a1.x : 0.f; a1.y = m1.y > 0.f ? a1.y : 0.f; a1.z = m1.z > 0.f ? a1.z : 0.f; a1.w = m1.w > 0.f ? a1.w : 0.f;
    }
    acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
    acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(a0.x, a0.y), h1 = __floats2bfloat162_rn(a0.z, a0.w);
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a1.x, a1.y), h3 = __floats2bfloat162_rn(a1.z, a1.w);
    uint4 v;
    v.x = *reinterpret_cast<const unsigned*>(&h0); v.y = *reinterpret_cast<const unsigned*>(&h1);
    v.z = *reinterpret_cast<const unsigned*>(&h2); v.w = *reinterpret_cast<const unsigned*>(&h3);
    *reinterpret_cast<uint4*>(g + b * gB + y * gH + x * gW + c0) = v;
  }
  if (!partial) return;
  // per-CTA partial sums: the cells_per_iter sub-rows are added in sub order (fixed), one row of C floats per CTA
  __shared__ float red[RB_THREADS * 8];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[(sub * C) + c0 + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += RB_THREADS) {
    float s = red[c];
    for (int k = 1; k < cells_per_iter; ++k) s += red[k * C + c];
    partial[(int64_t)blockIdx.x * C + c] = s;
  }
}

// bias_grad[c] = sum over the per-CTA partial rows.  Block = 32 channels x 8 part groups: thread (cx, g) adds the rows g, g+8, ... with
// four independent accumulators (k mod 32 classes, combined in a fixed order), the 8 groups meet in shared memory and are added
// in group order -- a fixed summation tree, hence bitwise reproducible, and ~40 dependent loads deep instead of ~1200.
__global__ void __launch_bounds__(256) bias_partials_reduce_kernel(const float* __restrict__ partial, int nparts, int C, float* __restrict__ bias_grad) {
  __shared__ float red[8][32];
  const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int k = g;
    for (; k + 24 < nparts; k += 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] += partial[(int64_t)(k + 8 * i) * C + c];
    }
    for (int i = 0; k < nparts; k += 8, ++i) s[i] += partial[(int64_t)k * C + c];
  }
  red[g][cx] = (s[0] + s[1]) + (s[2] + s[3]);
  __syncthreads();
  if (g == 0 && c < C) {
    float t = red[0][cx];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][cx];
    bias_grad[c] = t;
  }
}

// SM count of a device (cudaDeviceGetAttribute: microseconds; cudaGetDeviceProperties costs milliseconds per call)
static int sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) return 0;
  return n;
}

}  // namespace mvw

using namespace mvw;

extern "C" int64_t mvwarp_relu_bwd_scratch_floats(int C, int device) {
  const int sms = sm_count(device);
  if (C <= 0 || !sms) return 0;
  return (int64_t)sms * 8 * C;
}

extern "C" int mvwarp_relu_bwd_cast(const float* grad_out, const int64_t go_strides[4], const float* out, const int64_t out_strides[4],
                                    void* g16, const int64_t g_strides[4], float* bias_grad, float* scratch, int64_t scratch_floats,
                                    int B, int H, int W, int C, int device, void* stream) {
  if (!grad_out || !go_strides || !g16 || !g_strides || (out && !out_strides)) { set_error("relu_bwd_cast: null argument"); return MVWARP_EINVAL; }
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0) { set_error("relu_bwd_cast: bad B / H / W / C"); return MVWARP_EINVAL; }
  if (C % 8 || (RB_THREADS * 8) % C || C > RB_THREADS * 8) {
    set_error("relu_bwd_cast: C must divide 2048 and be a multiple of 8 (map_classifier[0] has 512 output channels)");
    return MVWARP_EUNSUPPORTED;
  }
  auto cl4 = [](const int64_t* s, int al) { return s[1] == 1 && s[0] % al == 0 && s[2] % al == 0 && s[3] % al == 0; };
  if (!cl4(go_strides, 4) || (out && !cl4(out_strides, 4)) || !cl4(g_strides, 8) || (reinterpret_cast<uintptr_t>(grad_out) & 15u) ||
      (out && (reinterpret_cast<uintptr_t>(out) & 15u)) || (reinterpret_cast<uintptr_t>(g16) & 15u)) {
    set_error("relu_bwd_cast: tensors must be channels-last with 16-byte aligned cell rows");
    return MVWARP_EUNSUPPORTED;
  }
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  const int sms = sm_count(device);
  if (!sms) return cuda_fail(cudaGetLastError(), "cudaDeviceGetAttribute");
  cudaError_t e;
  const int cells_per_iter = RB_THREADS / (C / 8);
  const int64_t ncell = (int64_t)B * H * W;
  int blocks = (int)std::min<int64_t>((ncell + cells_per_iter - 1) / cells_per_iter, (int64_t)sms * 8);
  float* partial = nullptr;
  if (bias_grad) {
    if (!scratch || scratch_floats < (int64_t)blocks * C) { set_error("relu_bwd_cast: scratch smaller than mvwarp_relu_bwd_scratch_floats()"); return MVWARP_EINVAL; }
    partial = scratch;
  }
  relu_bwd_cast_kernel<<<blocks, RB_THREADS, 0, (cudaStream_t)stream>>>(grad_out, go_strides[0], go_strides[2], go_strides[3], out,
                                                                         out ? out_strides[0] : 0, out ? out_strides[2] : 0, out ? out_strides[3] : 0,
                                                                         (__nv_bfloat16*)g16, g_strides[0], g_strides[2], g_strides[3], partial, B, H, W, C);
  e = cudaGetLastError();
  if (e == cudaSuccess && bias_grad) {
    bias_partials_reduce_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partial, blocks, C, bias_grad);
    e = cudaGetLastError();
  }
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "relu_bwd_cast launch");
}
