// mvwarp_conv1.cu -- tensor-core side of the path (north_star (c)): tcgen05 / TMEM GEMM building blocks for the
// fused warp -> map_classifier[0] implicit GEMM (/root/reference/multiview_detector/models/persp_trans_detector.py:51,81).
#include <cuda_bf16.h>

#include "mvwarp_internal.h"
#include "mvwarp_umma.h"

namespace mvw {
using namespace umma;

// -----------------------------------------------------------------------------------------------------
// Stage-1 validation kernel: D[M, 256] = A[M, K] * B[256, K]^T, bf16 in, fp32 out, K % 64 == 0, M % 128 == 0.
// One CTA per 128-row tile, operands staged by all threads into 128-byte-swizzled K-major shared memory,
// one thread issues tcgen05.mma, accumulators live in TMEM, epilogue tcgen05.ld -> global.
// No pipelining: this kernel exists to pin the descriptor / swizzle / TMEM conventions, not for speed.
// -----------------------------------------------------------------------------------------------------
constexpr int GT_M = 128, GT_N = 256, GT_K = 64;

__global__ void __launch_bounds__(128) gemm_tcgen05_test_kernel(const __nv_bfloat16* __restrict__ A,
                                                               const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // 128 x 128 B = 16 KB
  uint8_t* sB = smem + GT_M * 128;          // 256 x 128 B = 32 KB
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * GT_M;

  if (warp == 0) tmem_alloc(&tmem_holder, GT_N);
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t idesc = idesc_bf16_f32(GT_M, GT_N);
  uint32_t phase = 0;

  for (int kb = 0; kb < K / GT_K; ++kb) {
    // stage A (row tid) and B (rows tid, tid+128): 8 chunks of 16 bytes per row, swizzled
    const uint4* ga = reinterpret_cast<const uint4*>(A + (size_t)(m0 + tid) * K + kb * GT_K);
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sA + sw128_offset(tid, c)) = __ldg(ga + c);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = tid + 128 * h;
      const uint4* gb = reinterpret_cast<const uint4*>(B + (size_t)r * K + kb * GT_K);
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sB + sw128_offset(r, c)) = __ldg(gb + c);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t da = smem_desc_sw128(smem_u32(sA)), db = smem_desc_sw128(smem_u32(sB));
#pragma unroll
      for (int k = 0; k < GT_K / 16; ++k)  // 16 bf16 = 32 bytes per step: advance the start address by 2 (x16 B)
        mma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, kb > 0 || k > 0);
      mma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, phase);   // the MMAs have consumed this stage: shared memory may be overwritten
    phase ^= 1;
    tc_fence_after();
  }

  // epilogue: warp w owns TMEM lanes 32w..32w+31 = tile rows
  float* drow = D + (size_t)(m0 + warp * 32 + lane) * GT_N;
#pragma unroll 1
  for (int c0 = 0; c0 < GT_N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(drow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, GT_N);
}

}  // namespace mvw

using namespace mvw;

extern "C" int mvwarp_debug_gemm(const void* A, const void* B, void* D, int M, int N, int K, void* stream) {
  if (!A || !B || !D || N != GT_N || M <= 0 || M % GT_M || K <= 0 || K % GT_K) {
    set_error("debug_gemm: needs N == 256, M % 128 == 0, K % 64 == 0");
    return MVWARP_EINVAL;
  }
  const int smem = (GT_M + GT_N) * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm test)");
  gemm_tcgen05_test_kernel<<<M / GT_M, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (float*)D, K);
  e = cudaGetLastError();
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "debug_gemm launch");
}
