// mvwarp_layout.cu -- planar (NCHW) <-> channel-contiguous (NHWC) feature maps, for trunks that emit NCHW.
//
// The reference's trunk hands over NCHW-contiguous tensors (persp_trans_detector.py:63-65).  The warp kernels gather a texel's
// channels as one contiguous run, so an NCHW tensor costs them one 4-byte gather per channel and tap (a quarter of the HBM
// roofline, profiles/README.md).  For the LOW-RES trunk output that fused up-sample plans read (90x160: 29.5 MB per camera) the
// cheap way out is to transpose it once -- 2 x 206 MB of traffic per Wildtrack frame against the 1.86 GB tensor the reference
// materialises with F.interpolate -- and run the channel-contiguous kernels; the backward transposes the 9x smaller low-res
// gradient back.  Classic shared-memory tile transpose: 32 x 32 element tiles, +1 padding, coalesced on both sides.
#include <cuda_bf16.h>

#include "mvwarp_internal.h"

namespace mvw {

// src: [N][R][S] (S contiguous)  ->  dst: [N][S][R] (R contiguous).  NCHW -> NHWC: R = C, S = H*W.  NHWC -> NCHW: R = H*W, S = C.
template <typename T>
__global__ void __launch_bounds__(256) transpose_tiles_kernel(const T* __restrict__ src, T* __restrict__ dst, int R, int S,
                                                              int64_t src_n, int64_t dst_n) {
  __shared__ T tile[32][33];
  const int n = blockIdx.z;
  const T* s = src + (int64_t)n * src_n;
  T* d = dst + (int64_t)n * dst_n;
  const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = s0 + tx;
    if (r < R && c < S) tile[ty + 8 * k][tx] = __ldcs(s + (int64_t)r * S + c);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = s0 + ty + 8 * k, r = r0 + tx;
    if (r < R && c < S) __stcs(d + (int64_t)c * R + r, tile[tx][ty + 8 * k]);
  }
}

}  // namespace mvw

using namespace mvw;

extern "C" int mvwarp_layout_transpose(const void* src, int64_t src_image_stride, void* dst, int64_t dst_image_stride, int N, int C,
                                       int HW, int dtype, int to_channels_last, int device, void* stream) {
  if (!src || !dst) { set_error("layout_transpose: null argument"); return MVWARP_EINVAL; }
  if (N <= 0 || C <= 0 || HW <= 0) { set_error("layout_transpose: bad N / C / HW"); return MVWARP_EINVAL; }
  if (dtype != MVWARP_F32 && dtype != MVWARP_BF16) { set_error("layout_transpose: unknown dtype"); return MVWARP_EINVAL; }
  if (N > 65535) { set_error("layout_transpose: N exceeds the CUDA grid limit"); return MVWARP_EUNSUPPORTED; }
  const int R = to_channels_last ? C : HW, S = to_channels_last ? HW : C;
  if ((R + 31) / 32 > 65535) { set_error("layout_transpose: too many row tiles"); return MVWARP_EUNSUPPORTED; }
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  dim3 grid((S + 31) / 32, (R + 31) / 32, N), block(256);
  const int64_t per = (int64_t)C * HW;
  const int64_t sn = src_image_stride > 0 ? src_image_stride : per, dn = dst_image_stride > 0 ? dst_image_stride : per;
  if (sn < per || dn < per) { set_error("layout_transpose: image stride smaller than an image"); return MVWARP_EINVAL; }
  if (dtype == MVWARP_F32)
    transpose_tiles_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)src, (float*)dst, R, S, sn, dn);
  else
    transpose_tiles_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, R, S, sn, dn);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "layout_transpose launch");
}
