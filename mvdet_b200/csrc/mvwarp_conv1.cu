// mvwarp_conv1.cu -- tensor-core side of the path (north_star (c)): the fused warp -> map_classifier[0] implicit GEMM on
// tcgen05 / TMEM (/root/reference/multiview_detector/models/persp_trans_detector.py:51,81).  Kernels, in the order they were
// developed (each kept because the next one is measured against it; profiles/README.md has the numbers):
//   gemm_tcgen05_test_kernel   validation GEMM that pins the descriptor / swizzle / TMEM conventions (and the layout probe)
//   warp_conv1_kernel          correctness-first fused kernel, no pipelining            (MVWARP_CONV1_VARIANT=s)
//   warp_conv1_pipe_kernel     warp-specialised mbarrier pipeline, tap stages copied out of the halo patch; optionally a
//                              cluster with multicast weights (<CL>, MVWARP_CONV1_VARIANT=2|4) or CTA pairs with
//                              cta_group::2 MMAs (<1, true>, MVWARP_CONV1_PAIR=1); single CTAs: MVWARP_CONV1_PAIR=0
//   warp_conv1_direct_kernel   THE DEFAULT: CTA pairs, taps read in place from the halo patch (no tap copies)
//   conv1_tail_reduce_kernel   adds the K-split partial sums of the last partial wave of tiles
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "mvwarp_internal.h"
#include "mvwarp_umma.h"

namespace mvw {
using namespace umma;

#ifdef MVWARP_DEV  // development build only: validation GEMM + layout probe
// -----------------------------------------------------------------------------------------------------
// Stage-1 validation kernel: D[M, 256] = A[M, K] * B[256, K]^T, bf16 in, fp32 out, K % 64 == 0, M % 128 == 0.
// One CTA per 128-row tile, operands staged by all threads into 128-byte-swizzled K-major shared memory,
// one thread issues tcgen05.mma, accumulators live in TMEM, epilogue tcgen05.ld -> global.
// No pipelining: this kernel exists to pin the descriptor / swizzle / TMEM conventions, not for speed.
// -----------------------------------------------------------------------------------------------------
constexpr int GT_M = 128, GT_N = 256, GT_K = 64;

// Layout probe (mvwarp_debug_gemm_layout): the A tile's 8-row groups are placed `sbo` bytes apart (1024 = the canonical layout),
// the whole tile starts `start_cells` * 128 bytes into its buffer, every 16-byte chunk is swizzled by the ABSOLUTE address bits
// (chunk ^ ((row address >> 7) & 7)), and the descriptor gets stride-byte-offset = sbo and, if `use_base_offset`, the
// base-offset field (start >> 7) & 7.  Answers: can the nine conv taps be read straight out of a halo patch whose rows are not
// 1024 bytes apart (tile 16 rows x 8 cells, patch pitch 10 cells = 1280 bytes) instead of being copied into canonical stages?
__global__ void __launch_bounds__(128) gemm_tcgen05_test_kernel(const __nv_bfloat16* __restrict__ A,
                                                               const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int K,
                                                               int sbo, int start_cells, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // up to 32 KB for the A probe layouts
  uint8_t* sB = smem + 32768;               // 256 x 128 B = 32 KB
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
    const uint32_t a_row = (uint32_t)(start_cells * 128 + (tid >> 3) * sbo + (tid & 7) * 128);   // byte offset of row tid
    const uint32_t a_phase = ((smem_u32(sA) + a_row) >> 7) & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sA + a_row + ((c ^ a_phase) << 4)) = __ldg(ga + c);
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
      const uint32_t a_start = smem_u32(sA) + start_cells * 128;
      uint64_t da = smem_desc_sw128(a_start);
      da = (da & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32);
      if (use_base_offset) da |= (uint64_t)((a_start >> 7) & 7) << 49;
      const uint64_t db = smem_desc_sw128(smem_u32(sB));
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

#endif  // MVWARP_DEV
}  // namespace mvw

using namespace mvw;

#ifdef MVWARP_DEV  // validation hooks: exported by the development build only (mvdet_b200/libmvwarp_dev.so)
extern "C" int mvwarp_debug_gemm(const void* A, const void* B, void* D, int M, int N, int K, void* stream) {
  return mvwarp_debug_gemm_layout(A, B, D, M, N, K, 1024, 0, 0, stream);
}

extern "C" int mvwarp_debug_gemm_layout(const void* A, const void* B, void* D, int M, int N, int K, int sbo_bytes, int start_cells,
                                        int use_base_offset, void* stream) {
  if (!A || !B || !D || N != GT_N || M <= 0 || M % GT_M || K <= 0 || K % GT_K) {
    set_error("debug_gemm: needs N == 256, M % 128 == 0, K % 64 == 0");
    return MVWARP_EINVAL;
  }
  if (sbo_bytes < 1024 || sbo_bytes % 128 || start_cells < 0 || start_cells * 128 + 15 * sbo_bytes + 1024 > 32768) {
    set_error("debug_gemm_layout: the probe layout must fit 32 KB (sbo >= 1024, multiple of 128)");
    return MVWARP_EINVAL;
  }
  const int smem = 32768 + GT_N * 128 + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm test)");
  gemm_tcgen05_test_kernel<<<M / GT_M, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (float*)D, K, sbo_bytes,
                                                                           start_cells, use_base_offset);
  e = cudaGetLastError();
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "debug_gemm launch");
}

#endif  // MVWARP_DEV
// =====================================================================================================
// Fused warp -> conv1 (north_star (c)):  out1[b, n, y, x] = bias[n] + sum_{dy,dx,ch} W[n, ch, dy, dx] * X[b, ch, y+dy-1, x+dx-1]
// with X = warp_views(feats) (never materialised), zero padding, optional ReLU.
// Implicit GEMM on tcgen05: M = grid cells (128 per CTA: an 8 x 16 tile), N = 512 output channels (all of TMEM),
// K = 9 taps x (V*C + 2) channels in blocks of 64.  The A operand is PRODUCED: for each (view, 64-channel block)
// the CTA warps the tile's 10 x 18 halo patch once (bilinear blend in fp32, rounded to bf16) and the nine conv
// taps are shifted copies of that patch written into the 128-byte-swizzled K-major layout the tensor core reads.
// The B operand (weights) is pre-packed on the host in exactly that shared-memory image, one 64 KB block per
// (view, channel block, tap), and arrives by one bulk async copy (TMA engine) per block.
// =====================================================================================================
#include "mvwarp_coords.h"
#include "mvwarp_uptaps.h"

namespace mvw {

constexpr int CT_H = 8, CT_W = 16, CP_H = CT_H + 2, CP_W = CT_W + 2, CP_CELLS = CP_H * CP_W;  // tile and halo patch
constexpr int CV_N = 512, CV_KB = 64;
constexpr uint32_t CV_B_BYTES = CV_N * CV_KB * 2;  // one packed weight block (64 KB)

struct ConvParams {
  const float* T;
  const float* coord_x;
  const float* coord_y;
  int Hs, Ws, Hg, Wg, view0, nviews, C, B;
  int hl, wl;                  // fused up-sample plans (direct kernel, UPS): feats is the LOW-RES map [hl][wl]; (Hs, Ws) is the size it is
                               // up-sampled to before the warp (persp_trans_detector.py:65).  0, 0: feats is the map the warp reads
  const __nv_bfloat16* feats;
  int64_t fB, fV, fH, fW;
  const __nv_bfloat16* wpack;  // [nviews * C/64 * 9 (+1 coord block)][512][64] bf16, rows pre-swizzled
  const float* bias;           // [512] or nullptr
  void* out;                   // [B, Hg, Wg, 512] fp32 or bf16 (logical [B,512,Hg,Wg] channels-last)
  int64_t oB, oH, oW;          // in elements of the output dtype
  int out_bf16;
  int relu, with_coord, tiles_x, tiles_y;
  int tile0, ntiles;           // this launch covers tiles [tile0, tile0 + ntiles) (a band of tile rows) of every frame
  // work items = (frame, tile) pairs, one CTA each -- except the last partial wave: when nitems is not a multiple of the SM
  // count, the n_tail = nitems - n_full leftover items would occupy a few SMs for a whole tile time while the rest idle
  // (345 tiles on 148 SMs = 2.33 waves cost 3).  Those items are split along K (by views) into `split` CTAs each, which
  // write raw fp32 partial sums to `scratch`; conv1_tail_reduce_kernel adds them in a fixed order (+ bias, ReLU).
  int nitems, n_full, split, pair_mode;
  float* scratch;              // [n_tail * split][128][512] fp32, or nullptr (split == 1)
  // CTA-pair mode (PAIR kernels): CTAs 2i, 2i+1 of the launch form a cluster; whole items first, pair i takes items 2i, 2i+1
  // (n_full even), then for tail pair j and K part q cluster n_full/2 + j*split + q takes items n_full + 2j (+1), part q:
  // the two CTAs of a pair always walk the same weight blocks.  Scratch slot of (tail item n_full + 2j + r, part q) = (j*split + q)*2 + r.
  unsigned long long* stats;   // development: per CTA {cycles waiting for A, for B, total} of the MMA thread, or nullptr
  // Fused reduce-scatter over NVLink peer memory (view-sharded conv1, direct kernel only).  Rank r of p2p_n owns the grid rows
  // [p2p_row0[r], p2p_row0[r+1]); instead of storing its partial sums locally, the epilogue of every tile writes them -- raw
  // fp32, no bias / ReLU -- straight into the OWNER's inbox (a symmetric-memory buffer mapped into this process):
  //   inbox[owner][p2p_src][b][y - row0[owner]][x][512].  The stores ride along with the GEMM tile by tile, so the exchange costs
  // no extra kernel and no SMs; afterwards every rank sums the p2p_n slots of its own inbox (mvdet_b200.sharding).
  int p2p_n, p2p_src;
  int p2p_tile_shift;          // tiles are visited starting at the band AFTER this rank's own, so that at any moment the ranks store
                               // into different owners' inboxes (all ranks walking the map in the same order made every rank hit the
                               // same owner at once: 32 ms instead of 8 on 8 GPUs)
  int p2p_row0[9];
  unsigned long long p2p_inbox[8];
  long long p2p_src_pitch, p2p_b_pitch;   // elements between source slots / batch elements of an inbox
};

struct ConvTap {
  float w[4];
  int x0, y0;
  unsigned mask;
  int pad;
};

__device__ __forceinline__ float bf_lo(unsigned v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(unsigned v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ unsigned pack_bf2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const unsigned*>(&h);
}

#ifdef MVWARP_DEV  // superseded kernel generation, kept in the development build for comparison runs
// Step A: correctness-first version -- one stage, every thread takes part in every phase.
__global__ void __launch_bounds__(128) warp_conv1_kernel(const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                               // 128 x 128 B
  uint8_t* sB = smem + 16384;                       // 512 x 128 B
  uint8_t* sP = smem + 16384 + CV_B_BYTES;          // patch: 180 cells x 128 B (plain layout)
  ConvTap* sT = reinterpret_cast<ConvTap*>(sP + CP_CELLS * 128);
  __shared__ __align__(8) uint64_t b_full, mma_done;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = p.tile0 + blockIdx.x, tx0 = (tile % p.tiles_x) * CT_W, ty0 = (tile / p.tiles_x) * CT_H, b = blockIdx.z;

  if (warp == 0) tmem_alloc(&tmem_holder, CV_N);
  if (tid == 0) {
    mbar_init(&b_full, 1);
    mbar_init(&mma_done, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t idesc = idesc_bf16_f32(128, 256);
  uint32_t ph_b = 0, ph_m = 0;
  int kcount = 0;
  const int ncb = p.C / CV_KB;
  const int my = tid / CT_W, mx = tid % CT_W;  // this thread's A row = tile cell (my, mx)

  auto run_block = [&](const __nv_bfloat16* wblock) {  // A stage is in sA; fetch B, run the 128 x 512 x 64 MMA block
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&b_full, CV_B_BYTES);
      bulk_g2s(sB, wblock, CV_B_BYTES, &b_full);
      mbar_wait(&b_full, ph_b);
      tc_fence_after();
      const uint64_t da = smem_desc_sw128(smem_u32(sA)), db0 = smem_desc_sw128(smem_u32(sB)),
                     db1 = smem_desc_sw128(smem_u32(sB) + 256 * 128);
#pragma unroll
      for (int k = 0; k < CV_KB / 16; ++k) {
        mma_bf16(tmem, da + 2 * k, db0 + 2 * k, idesc, kcount > 0 || k > 0);
        mma_bf16(tmem + 256, da + 2 * k, db1 + 2 * k, idesc, kcount > 0 || k > 0);
      }
      mma_commit(&mma_done);
    }
    ph_b ^= 1;
    ++kcount;
    mbar_wait(&mma_done, ph_m);  // sA / sB free again
    ph_m ^= 1;
    tc_fence_after();
  };

  for (int vl = 0; vl < p.nviews; ++vl) {
    const int v = p.view0 + vl;
    // taps of the halo patch for this view (cells outside the grid are the conv's zero padding)
    for (int i = tid; i < CP_CELLS; i += 128) {
      const int gy = ty0 + i / CP_W - 1, gx = tx0 + i % CP_W - 1;
      ConvTap t;
      t.mask = 0; t.x0 = t.y0 = 0; t.pad = 0;
      t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
      if (gy >= 0 && gy < p.Hg && gx >= 0 && gx < p.Wg) {
        float Tm[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) Tm[q] = __ldg(p.T + 9 * v + q);
        float sx, sy;
        norm_src_coord(Tm, gx, gy, p.Hg, p.Wg, sx, sy);
        const Taps tp = make_taps(sx, sy, p.Hs, p.Ws);
#pragma unroll
        for (int k = 0; k < 4; ++k) t.w[k] = tp.w[k];
        t.x0 = tp.x0; t.y0 = tp.y0; t.mask = tp.mask;
      }
      sT[i] = t;
    }
    __syncthreads();
    const __nv_bfloat16* fv = p.feats + (int64_t)b * p.fB + (int64_t)vl * p.fV;
    for (int cb = 0; cb < ncb; ++cb) {
      // warp the patch: 180 cells x 8 chunks of 8 channels
      for (int task = tid; task < CP_CELLS * 8; task += 128) {
        const int cell = task >> 3, ch = task & 7;
        const ConvTap t = sT[cell];
        uint4 r = make_uint4(0, 0, 0, 0);
        if (t.mask) {
          uint4 tv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
            tv[k] = (t.mask >> k & 1) ? __ldg(reinterpret_cast<const uint4*>(fv + (int64_t)yy * p.fH + (int64_t)xx * p.fW + cb * CV_KB + ch * 8))
                                      : make_uint4(0, 0, 0, 0);
          }
          const unsigned* a = &tv[0].x; const unsigned* bq = &tv[1].x; const unsigned* c = &tv[2].x; const unsigned* d = &tv[3].x;
          unsigned o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float lo = __fmaf_rn(bf_lo(d[q]), t.w[3], __fmaf_rn(bf_lo(c[q]), t.w[2], __fmaf_rn(bf_lo(bq[q]), t.w[1], __fmul_rn(bf_lo(a[q]), t.w[0]))));
            const float hi = __fmaf_rn(bf_hi(d[q]), t.w[3], __fmaf_rn(bf_hi(c[q]), t.w[2], __fmaf_rn(bf_hi(bq[q]), t.w[1], __fmul_rn(bf_hi(a[q]), t.w[0]))));
            o[q] = pack_bf2(lo, hi);
          }
          r = make_uint4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<uint4*>(sP + cell * 128 + ch * 16) = r;
      }
      __syncthreads();
      for (int tap = 0; tap < 9; ++tap) {
        const int srcc = (my + tap / 3) * CP_W + (mx + tap % 3);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(sA + sw128_offset(tid, c)) = *reinterpret_cast<const uint4*>(sP + srcc * 128 + c * 16);
        run_block(p.wpack + (size_t)((vl * ncb + cb) * 9 + tap) * (CV_N * CV_KB));
      }
    }
    __syncthreads();
  }
  if (p.with_coord) {  // the two coord_map channels: all 9 taps x 2 channels in one K block (k = tap*2 + ch), rest zero
    unsigned short vals[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) vals[k] = 0;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int gy = ty0 + my + tap / 3 - 1, gx = tx0 + mx + tap % 3 - 1;
      if (gy >= 0 && gy < p.Hg && gx >= 0 && gx < p.Wg) {
        vals[tap * 2] = __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p.coord_x + gx)));
        vals[tap * 2 + 1] = __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p.coord_y + gy)));
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 r;
      r.x = vals[c * 8 + 0] | ((unsigned)vals[c * 8 + 1] << 16); r.y = vals[c * 8 + 2] | ((unsigned)vals[c * 8 + 3] << 16);
      r.z = vals[c * 8 + 4] | ((unsigned)vals[c * 8 + 5] << 16); r.w = vals[c * 8 + 6] | ((unsigned)vals[c * 8 + 7] << 16);
      *reinterpret_cast<uint4*>(sA + sw128_offset(tid, c)) = r;
    }
    run_block(p.wpack + (size_t)(p.nviews * ncb * 9) * (CV_N * CV_KB));
  }

  // epilogue: thread = tile cell = TMEM lane; 512 fp32 columns in chunks of 32
  const int gy = ty0 + my, gx = tx0 + mx;
  const bool inside = gy < p.Hg && gx < p.Wg;
  float* orow = reinterpret_cast<float*>(p.out) + (int64_t)b * p.oB + (int64_t)gy * p.oH + (int64_t)gx * p.oW;   // fp32 output only
#pragma unroll 1
  for (int c0 = 0; c0 < CV_N; c0 += 32) {
    float vv[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, vv);
    if (inside) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 o = make_float4(vv[i], vv[i + 1], vv[i + 2], vv[i + 3]);
        if (p.bias) { o.x += __ldg(p.bias + c0 + i); o.y += __ldg(p.bias + c0 + i + 1); o.z += __ldg(p.bias + c0 + i + 2); o.w += __ldg(p.bias + c0 + i + 3); }
        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(orow + c0 + i) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, CV_N);
}
#endif  // MVWARP_DEV


// -----------------------------------------------------------------------------------------------------
// Pipelined version: warp-specialised, mbarrier rings.
//   warps 0-7  producers (256 threads): warp the halo patch (once per view x 64-channel block) and write the nine
//              shifted A stages; the gathers of the NEXT patch are issued before the nine copies of the current
//              one and stay in flight in registers meanwhile, so their latency hides behind ~9 K blocks of MMA work.
//              At the end the same warps drain TMEM (epilogue): warp w reads lane quarter w%4, column half w/4.
//   warp 8     one thread issues tcgen05.mma (8 per K block: 4 k-steps x 2 halves of N) and commits the stage barriers;
//   warp 9     one thread streams the packed weight blocks with bulk async copies (64 KB each).
// Rings: A = PA stages x 16 KB, B = PB stages x 64 KB.  full barriers: producers / TMA -> MMA; empty barriers:
// tcgen05.commit -> producers / TMA.
// CL = thread-block cluster size: the CL CTAs of a cluster work on CL different cell tiles but stream the SAME weight
// blocks, so each CTA fetches 1/CL of every block and multicasts it into all CL shared memories.
// -----------------------------------------------------------------------------------------------------
constexpr int PA = 3, PB = 2;
constexpr int NPROD = 256, PIPE_THREADS = NPROD + 64;
constexpr int PATCH_TASKS = CP_CELLS * 8, PATCH_ROUNDS = (PATCH_TASKS + NPROD - 1) / NPROD;  // 1440 tasks, 6 per thread

__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// PAIR: CTA-pair mode (cta_group::2).  The two CTAs of a cluster work on two cell tiles and run ONE MMA of M = 256 per step,
// issued by the even CTA; each CTA keeps only HALF of every weight block (N rows [128r, 128r+128) and [256+128r, ...)) in
// shared memory.  The single-CTA kernel moves 16 (A stage write) + 16 (patch read) + 16 (A read) + 64 (weights in) + 64 (weights
// read) = 176 KB through shared memory per K block against 128 B per clock = 0.72 us, more than the 0.55 us the tensor core
// needs: shared-memory bandwidth, not the tensor pipe, bounds it.  Halving the weight traffic (112 KB, 0.46 us) lifts that.
// Weight halves arrive as TMA tensor boxes (.cta_group::2) that complete on the leader's barrier; the peer's producer warps
// publish their A stages with one remote arrive per warp; the leader's commits reach both CTAs' empty barriers by multicast.
#ifndef MVW_PA2
#define MVW_PA2 5
#define MVW_PB2 3
#endif
constexpr int PA2 = MVW_PA2, PB2 = MVW_PB2;           // ring depths in pair mode (16 KB A stages, 32 KB weight stages)
constexpr uint32_t CV_B2_BYTES = CV_B_BYTES / 2;

__device__ __forceinline__ void tma_wbox_2cta(uint32_t smem_dst, const CUtensorMap* tmap, int row, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(0), "r"(row), "r"(bar_cluster)
               : "memory");
}

#ifdef MVWARP_DEV  // superseded kernel generations (single CTA / cluster multicast / copying CTA pairs)
template <int CL, bool PAIR>
__global__ void __launch_bounds__(PIPE_THREADS, 1) warp_conv1_pipe_kernel(const __grid_constant__ CUtensorMap tmap_w, const ConvParams p) {
  static_assert(!PAIR || CL == 1, "pair mode does not multicast the weights");
  // rings: 3 x 16 KB of A + 2 x 64 KB of weights, or in pair mode 5 x 16 KB + 3 x 32 KB (the same 176 KB): the A ring is
  // deeper there because a stage's round trip now crosses SMs (remote arrive -> leader's MMA -> multicast commit)
  constexpr int NPA = PAIR ? PA2 : PA, NPB = PAIR ? PB2 : PB;
  constexpr uint32_t BSTAGE = PAIR ? CV_B2_BYTES : CV_B_BYTES;
  static_assert(NPA * 16384 + NPB * BSTAGE == PA * 16384 + PB * CV_B_BYTES, "both modes use the same shared-memory budget");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                          // NPA x 16 KB
  uint8_t* sB = smem + NPA * 16384;                            // NPB x BSTAGE
  uint8_t* sP = sB + NPB * BSTAGE;                             // patch: 180 cells x 128 B (plain layout)
  ConvTap* sT = reinterpret_cast<ConvTap*>(sP + CP_CELLS * 128);  // taps of the halo cells: 2 buffers (current / next view)
  __shared__ __align__(8) uint64_t a_full[PA2], a_empty[PA2], b_full[PB2], b_empty[PB2], acc_full;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t prank = PAIR ? cluster_ctarank() : 0;      // rank in the CTA pair (0 = leader)
  // work item of this CTA (see ConvParams): whole items first, then the K-split parts of the tail items
  int item = blockIdx.x, part = 0, nparts = 1, slot = 0;
  if (PAIR) {
    const int cpair = blockIdx.x >> 1, nfp = p.n_full >> 1;
    if (cpair >= nfp) {
      const int q = cpair - nfp, j = q / p.split;
      part = q % p.split; nparts = p.split;
      item = p.n_full + 2 * j + (int)prank;
      slot = q * 2 + (int)prank;
    }
  } else if ((int)blockIdx.x >= p.n_full && p.split > 1) {
    const int q = blockIdx.x - p.n_full;
    item = p.n_full + q / p.split; part = q % p.split; nparts = p.split;
    slot = q;
  }
  const bool real_tile = item < p.nitems;                   // with clusters the grid is padded to a multiple of CL: a padding CTA
  if (!real_tile) item = p.nitems - 1;                      // recomputes the last item (it must feed the multicast), stores nothing
  const int b = item / p.ntiles, tile = p.tile0 + item % p.ntiles;
  const int tx0 = (tile % p.tiles_x) * CT_W, ty0 = (tile / p.tiles_x) * CT_H;
  const int ncb = p.C / CV_KB;
  const int vlo = (part * p.nviews) / nparts, vhi = ((part + 1) * p.nviews) / nparts;   // this CTA's views (local indices)
  const bool coord_here = p.with_coord && part == nparts - 1;
  const bool partial = nparts > 1;
  const int NK = (vhi - vlo) * ncb * 9 + (coord_here ? 1 : 0);
  const int kb_main = (vhi - vlo) * ncb * 9;                // K blocks before the coord block
  const int wblk0 = vlo * ncb * 9, wblk_coord = p.nviews * ncb * 9;   // indices into the packed weight
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

  if (warp == 0) { if (PAIR) tmem_alloc_2cta(&tmem_holder, CV_N); else tmem_alloc(&tmem_holder, CV_N); }
  if (tid == 0) {
    // pair mode: one arrival per producer warp on the CTA's own barrier; the peer's relay thread forwards each completed
    // stage to the leader's barrier with ONE remote arrive (a cluster-scope release per producer warp and stage made the
    // peer's producers, and with them the pair, 35 % slower)
    for (int i = 0; i < NPA; ++i) { mbar_init(&a_full[i], PAIR ? NPROD / 32 + (prank == 0 ? 1 : 0) : NPROD); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NPB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], PAIR ? 1 : CL); }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  if (CL > 1 || PAIR) cluster_sync(); else __syncthreads();   // every CTA's barriers exist before any remote arrival
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 9) {
    // ===== weight streamer =====
    if (lane == 0) {
      for (int kb = 0; kb < NK; ++kb) {
        const int s = kb % NPB;
        const int wblk = kb < kb_main ? wblk0 + kb : wblk_coord;
        if (kb >= NPB) mbar_wait(&b_empty[s], ((kb / NPB) - 1) & 1);
        if (PAIR) {   // my half of the block: two boxes of 128 rows, completing on the leader's barrier
          if (prank == 0) mbar_expect_tx(&b_full[s], CV_B_BYTES);
          const uint32_t dst = smem_u32(sB + (size_t)s * BSTAGE), bar = map_to_cta(smem_u32(&b_full[s]), 0);
          tma_wbox_2cta(dst, &tmap_w, wblk * CV_N + 128 * (int)prank, bar);
          tma_wbox_2cta(dst + BSTAGE / 2, &tmap_w, wblk * CV_N + 256 + 128 * (int)prank, bar);
          continue;
        }
        mbar_expect_tx(&b_full[s], CV_B_BYTES);   // the whole block lands here: my slice and the peers' multicasts
        if (CL == 1) {
          bulk_g2s(sB + (size_t)s * CV_B_BYTES, p.wpack + (size_t)wblk * (CV_N * CV_KB), CV_B_BYTES, &b_full[s]);
        } else {
          constexpr uint32_t SLICE = CV_B_BYTES / CL;
          bulk_g2s_multicast(sB + (size_t)s * CV_B_BYTES + crank * SLICE,
                             reinterpret_cast<const uint8_t*>(p.wpack + (size_t)wblk * (CV_N * CV_KB)) + crank * SLICE, SLICE, &b_full[s], CMASK);
        }
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0 && prank == 0) {   // pair mode: the leader issues for both CTAs
      const uint32_t idesc = PAIR ? idesc_bf16_f32(256, 256) : idesc_bf16_f32(128, 256);
      long long wa = 0, wb = 0;
      const long long t_start = clock64();
      for (int kb = 0; kb < NK; ++kb) {
        const int sa = kb % NPA, sb = kb % NPB;
        const long long t0 = clock64();
        if (PAIR) mbar_wait_cluster(&a_full[sa], (kb / NPA) & 1);   // arrivals from the peer's producer warps
        else mbar_wait(&a_full[sa], (kb / NPA) & 1);
        const long long t1 = clock64();
        mbar_wait(&b_full[sb], (kb / NPB) & 1);
        const long long t2 = clock64();
        wa += t1 - t0; wb += t2 - t1;
        tc_fence_after();
        const uint32_t bbase = smem_u32(sB + (size_t)sb * BSTAGE);
        const uint64_t da = smem_desc_sw128(smem_u32(sA + sa * 16384)), db0 = smem_desc_sw128(bbase),
                       db1 = smem_desc_sw128(bbase + BSTAGE / 2);
#pragma unroll
        for (int k = 0; k < CV_KB / 16; ++k) {
          if (PAIR) {
            mma_bf16_2cta(tmem, da + 2 * k, db0 + 2 * k, idesc, kb > 0 || k > 0);
            mma_bf16_2cta(tmem + 256, da + 2 * k, db1 + 2 * k, idesc, kb > 0 || k > 0);
          } else {
            mma_bf16(tmem, da + 2 * k, db0 + 2 * k, idesc, kb > 0 || k > 0);
            mma_bf16(tmem + 256, da + 2 * k, db1 + 2 * k, idesc, kb > 0 || k > 0);
          }
        }
        if (PAIR) {
          mma_commit_2cta(&a_empty[sa], 3);
          mma_commit_2cta(&b_empty[sb], 3);
        } else {
          mma_commit(&a_empty[sa]);
          if (CL == 1) mma_commit(&b_empty[sb]);
          else mma_commit_multicast(&b_empty[sb], CMASK);   // every CTA of the cluster writes into this stage of mine
        }
      }
      if (PAIR) mma_commit_2cta(&acc_full, 3);
      else mma_commit(&acc_full);
      if (p.stats) {
        unsigned long long* st = p.stats + 3 * (size_t)blockIdx.x;
        st[0] = (unsigned long long)wa; st[1] = (unsigned long long)wb; st[2] = (unsigned long long)(clock64() - t_start);
      }
    } else if (PAIR && lane == 0) {
      // ===== relay (peer CTA): forward "my A stage is complete" to the leader's barrier =====
      for (int kb = 0; kb < NK; ++kb) {
        const int sa = kb % NPA;
        mbar_wait(&a_full[sa], (kb / NPA) & 1);
        mbar_arrive_cluster(map_to_cta(smem_u32(&a_full[sa]), 0));
      }
    }
  } else {
    // ===== producers (256 threads), then epilogue =====
    const int m = tid & 127, half = tid >> 7;    // A row = tile cell (my, mx); this thread writes chunks [4*half, 4*half+4)
    const int my = m / CT_W, mx = m % CT_W;
    int kb = 0;
    auto acquire_stage = [&]() -> uint8_t* {
      const int sa = kb % NPA;
      if (kb >= NPA) mbar_wait(&a_empty[sa], ((kb / NPA) - 1) & 1);
      return sA + sa * 16384;
    };
    auto publish_stage = [&]() {
      fence_proxy_async();
      if (PAIR) {   // one arrival per warp
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[kb % NPA]);
      } else {
        mbar_arrive(&a_full[kb % NPA]);
      }
      ++kb;
    };
    auto compute_taps = [&](int vl, ConvTap* dst) {  // taps of the halo patch for one view (outside the grid = conv zero padding)
      const int v = p.view0 + vl;
      for (int i = tid; i < CP_CELLS; i += NPROD) {
        const int gy = ty0 + i / CP_W - 1, gx = tx0 + i % CP_W - 1;
        ConvTap t;
        t.mask = 0; t.x0 = t.y0 = 0; t.pad = 0;
        t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
        if (gy >= 0 && gy < p.Hg && gx >= 0 && gx < p.Wg) {
          float Tm[9];
#pragma unroll
          for (int q = 0; q < 9; ++q) Tm[q] = __ldg(p.T + 9 * v + q);
          float sx, sy;
          norm_src_coord(Tm, gx, gy, p.Hg, p.Wg, sx, sy);
          const Taps tp = make_taps(sx, sy, p.Hs, p.Ws);
#pragma unroll
          for (int k = 0; k < 4; ++k) t.w[k] = tp.w[k];
          t.x0 = tp.x0; t.y0 = tp.y0; t.mask = tp.mask;
        }
        dst[i] = t;
      }
    };
    uint4 tv[PATCH_ROUNDS][4];   // gathers in flight for the next patch
    auto issue_gathers = [&](int vl, int cb, const ConvTap* taps) {
      const __nv_bfloat16* fv = p.feats + (int64_t)b * p.fB + (int64_t)vl * p.fV + cb * CV_KB;
#pragma unroll
      for (int j = 0; j < PATCH_ROUNDS; ++j) {
        const int task = tid + j * NPROD;
        const int cell = task >> 3, ch = task & 7;
        if (task < PATCH_TASKS) {
          const ConvTap t = taps[cell];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
            tv[j][k] = (t.mask >> k & 1) ? __ldg(reinterpret_cast<const uint4*>(fv + (int64_t)yy * p.fH + (int64_t)xx * p.fW + ch * 8))
                                         : make_uint4(0, 0, 0, 0);
          }
        }
      }
    };
    auto blend_patch = [&](const ConvTap* taps) {
#pragma unroll
      for (int j = 0; j < PATCH_ROUNDS; ++j) {
        const int task = tid + j * NPROD;
        const int cell = task >> 3, ch = task & 7;
        if (task < PATCH_TASKS) {
          const ConvTap t = taps[cell];
          const unsigned* a = &tv[j][0].x; const unsigned* bq = &tv[j][1].x; const unsigned* c = &tv[j][2].x; const unsigned* d = &tv[j][3].x;
          unsigned o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float lo = __fmaf_rn(bf_lo(d[q]), t.w[3], __fmaf_rn(bf_lo(c[q]), t.w[2], __fmaf_rn(bf_lo(bq[q]), t.w[1], __fmul_rn(bf_lo(a[q]), t.w[0]))));
            const float hi = __fmaf_rn(bf_hi(d[q]), t.w[3], __fmaf_rn(bf_hi(c[q]), t.w[2], __fmaf_rn(bf_hi(bq[q]), t.w[1], __fmul_rn(bf_hi(a[q]), t.w[0]))));
            o[q] = pack_bf2(lo, hi);
          }
          // the patch is XOR-swizzled like the stages (chunk position = ch ^ (cell & 7)): the tap copies below read one chunk
          // of 8 consecutive cells per quarter-warp, which would otherwise be an 8-way bank conflict on every load
          *reinterpret_cast<uint4*>(sP + cell * 128 + ((ch ^ (cell & 7)) << 4)) = t.mask ? make_uint4(o[0], o[1], o[2], o[3]) : make_uint4(0, 0, 0, 0);
        }
      }
    };

    const int nblk = vhi * ncb;        // (view, channel-block) pairs [vlo * ncb, vhi * ncb)
    compute_taps(vlo, sT + (vlo & 1) * CP_CELLS);
    producers_sync();
    issue_gathers(vlo, 0, sT + (vlo & 1) * CP_CELLS);
    for (int blk = vlo * ncb; blk < nblk; ++blk) {
      const int vl = blk / ncb;
      const ConvTap* taps = sT + (vl & 1) * CP_CELLS;
      blend_patch(taps);                 // consumes the gathers issued one block earlier
      producers_sync();                  // patch complete
      if (blk + 1 < nblk) {              // start the next block's gathers; they fly during the nine copies below
        const int nvl = (blk + 1) / ncb, ncbi = (blk + 1) % ncb;
        if (ncbi == 0) {                 // next block opens a new view: its taps go to the other buffer
          compute_taps(nvl, sT + (nvl & 1) * CP_CELLS);
          producers_sync();
        }
        issue_gathers(nvl, ncbi, sT + (nvl & 1) * CP_CELLS);
      }
      for (int tap = 0; tap < 9; ++tap) {
        uint8_t* stage = acquire_stage();
        const int srcc = (my + tap / 3) * CP_W + (mx + tap % 3);
#pragma unroll
        for (int c = 4 * half; c < 4 * half + 4; ++c)
          *reinterpret_cast<uint4*>(stage + sw128_offset(m, c)) = *reinterpret_cast<const uint4*>(sP + srcc * 128 + ((c ^ (srcc & 7)) << 4));
        publish_stage();
      }
      producers_sync();                  // every thread has finished reading the patch: it may be overwritten
    }
    if (coord_here) {  // the two coord_map channels: all 9 taps x 2 channels in one K block (k = tap*2 + ch), rest zero
      unsigned short vals[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) vals[k] = 0;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int gy = ty0 + my + tap / 3 - 1, gx = tx0 + mx + tap % 3 - 1;
        if (gy >= 0 && gy < p.Hg && gx >= 0 && gx < p.Wg) {
          vals[tap * 2] = __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p.coord_x + gx)));
          vals[tap * 2 + 1] = __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p.coord_y + gy)));
        }
      }
      uint8_t* stage = acquire_stage();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = 4 * half + cc;     // chunks 0..3 carry k = 0..31 (18 used), chunks 4..7 are zero
        uint4 r = make_uint4(0, 0, 0, 0);
        if (half == 0) {
          r.x = vals[cc * 8 + 0] | ((unsigned)vals[cc * 8 + 1] << 16); r.y = vals[cc * 8 + 2] | ((unsigned)vals[cc * 8 + 3] << 16);
          r.z = vals[cc * 8 + 4] | ((unsigned)vals[cc * 8 + 5] << 16); r.w = vals[cc * 8 + 6] | ((unsigned)vals[cc * 8 + 7] << 16);
        }
        *reinterpret_cast<uint4*>(stage + sw128_offset(m, c)) = r;
      }
      publish_stage();
    }
    // ----- epilogue: thread = tile cell = TMEM lane (quarter warp%4); warps 0-3 take columns 0-255, warps 4-7 256-511 -----
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int er = (warp & 3) * 32 + lane, ey = er / CT_W, ex = er % CT_W;
    const int gy = ty0 + ey, gx = tx0 + ex;
    const bool inside = real_tile && gy < p.Hg && gx < p.Wg;
    const int64_t ooff = (int64_t)b * p.oB + (int64_t)gy * p.oH + (int64_t)gx * p.oW;
    const int cbeg = (warp >> 2) * 256;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + 256; c0 += 32) {
      float vv[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, vv);
      if (partial) {   // K-split tail item: raw partial sums, combined by conv1_tail_reduce_kernel
        float4* o = reinterpret_cast<float4*>(p.scratch + ((size_t)slot * 128 + er) * CV_N + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
      } else if (inside) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (p.bias) vv[i] += __ldg(p.bias + c0 + i);
          if (p.relu) vv[i] = fmaxf(vv[i], 0.f);
        }
        if (p.out_bf16) {   // feeds the tcgen05 conv of map_classifier[2] (mvwarp_conv2.cu), which reads bf16 boxes
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = make_uint4(pack_bf2(vv[8 * i], vv[8 * i + 1]), pack_bf2(vv[8 * i + 2], vv[8 * i + 3]),
                              pack_bf2(vv[8 * i + 4], vv[8 * i + 5]), pack_bf2(vv[8 * i + 6], vv[8 * i + 7]));
        } else {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + ooff + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  if (CL > 1 || PAIR) cluster_sync(); else __syncthreads();   // nobody leaves while a peer may still write into it / read from it
  if (warp == 0) { if (PAIR) tmem_dealloc_2cta(tmem, CV_N); else tmem_dealloc(tmem, CV_N); }
}
#endif  // MVWARP_DEV

// -----------------------------------------------------------------------------------------------------
// Direct version (the default): CTA pairs as above, and the nine conv taps are no longer COPIED out of the halo patch --
// the tensor core reads them in place.  The swizzle of a K-major operand is a function of the absolute shared-memory
// address (probed: scripts/probe_umma_layout.py), so a descriptor may start at any 128-byte row of a buffer whose rows were
// written with that rule, and the stride between its 8-row groups is free.  With the GEMM rows ordered x-major (row m = cell
// (mx = m / 8, my = m % 8) of a tile 16 cells wide and 8 high) and the patch stored x-major too ([18][10] cells of 128 bytes),
// the 8-row group g of tap (dy, dx) is the 8 consecutive patch cells (g + dx, dy .. dy+7): start = patch + ((dx * 10 + dy) * 128,
// group stride = 10 * 128 = 1280 bytes.  So a patch (one view x 64 channels) is produced once and consumed by nine K blocks
// straight away; the A stages and their 2 x 16 KB of shared-memory traffic per K block are gone (82 KB per K block in total
// against the 112 KB of the copying pair kernel), and the ring is a ring of PATCHES.
// -----------------------------------------------------------------------------------------------------
constexpr int DT_X = 16, DT_Y = 8;                                             // tile: 16 cells in x, 8 in y
constexpr int DP_X = DT_X + 2, DP_Y = DT_Y + 2, DP_CELLS = DP_X * DP_Y;        // halo patch 18 x 10 = 180 cells, x-major
constexpr int DPATCH_BYTES = (DP_CELLS * 128 + 1023) / 1024 * 1024;           // 23 KB: every patch buffer starts 1024-aligned
constexpr int NPATCH = 3, DPB = 4;                                             // patch ring, weight ring (32 KB stages)
constexpr int D_SBO = DP_Y * 128;                                              // 1280 bytes between the 8-row groups of a tap
constexpr int DP_TASKS = DP_CELLS * 8, DP_ROUNDS = (DP_TASKS + NPROD - 1) / NPROD;

// UPS (fused up-sample plans): the warp reads F.interpolate(low-res, (Hs, Ws), bilinear, align_corners=False) -- which is never
// materialised.  Both steps are bilinear and separable: a high-res tap column X is the two low-res columns (i0, i1) with weights
// (l0, l1) (ATen upsample_bilinear2d: mvwarp_plan.cu upsample_axis, the same fp32 formulas), two neighbouring high-res columns
// fall into a window of at most THREE low-res columns when the map is up-sampled (scale <= 1), and the out-of-range mask of
// grid_sample's zero padding is a product of an x flag and a y flag.  So a halo cell is a 3 x 3 low-res window with weights
// wy[j] * wx[i]: 32 bytes per cell like ConvTap, nine (masked) 16-byte gathers per task instead of four.
static_assert(sizeof(ConvTapUp) == sizeof(ConvTap), "both tap records share the shared-memory budget");

template <bool UPS>
__global__ void __launch_bounds__(PIPE_THREADS, 1) warp_conv1_direct_kernel(const __grid_constant__ CUtensorMap tmap_w, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sPatch = smem;                                        // NPATCH x 23 KB
  uint8_t* sC = smem + NPATCH * DPATCH_BYTES;                    // 16 KB: canonical A stage of the coord_map K block
  uint8_t* sB = sC + 16384;                                      // DPB x 32 KB
  ConvTap* sT = reinterpret_cast<ConvTap*>(sB + DPB * CV_B2_BYTES);   // taps of the halo cells: 2 buffers (current / next view)
  __shared__ __align__(8) uint64_t p_full[NPATCH], p_empty[NPATCH], c_full, b_full[DPB], b_empty[DPB], acc_full;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t prank = cluster_ctarank();
  int item = blockIdx.x, part = 0, nparts = 1, slot = 0;
  {
    const int cpair = blockIdx.x >> 1, nfp = p.n_full >> 1;
    if (cpair >= nfp) {
      const int q = cpair - nfp, j = q / p.split;
      part = q % p.split; nparts = p.split;
      item = p.n_full + 2 * j + (int)prank;
      slot = q * 2 + (int)prank;
    }
  }
  const int item_peer = min(item ^ 1, p.nitems - 1);          // the other CTA of the pair (n_full is even: items pair as 2k, 2k + 1)
  const bool real_tile = item < p.nitems;
  if (!real_tile) item = p.nitems - 1;
  const int b = item / p.ntiles, tile = p.tile0 + (item % p.ntiles + p.p2p_tile_shift) % p.ntiles;
  const int tx0 = (tile % p.tiles_x) * DT_X, ty0 = (tile / p.tiles_x) * DT_Y;
  const int ncb = p.C / CV_KB;
  const int vlo = (part * p.nviews) / nparts, vhi = ((part + 1) * p.nviews) / nparts;
  const bool coord_here = p.with_coord && part == nparts - 1;
  const bool partial = nparts > 1;

  // ----- views that reach neither tile of the pair are skipped: their K blocks would multiply zeros -----
  // A camera sees only part of the ground plane (81 % of the cells on Wildtrack, 72 % on MultiviewX), and the area it misses is
  // contiguous, so for 12.5 % (Wildtrack) to 22 % (MultiviewX, the 32-camera rig) of the (tile pair, view) combinations every halo
  // cell of BOTH tiles falls outside the view's image: the whole patch is zero and its 9 * C/64 K blocks add +0 to the accumulator.
  // Both CTAs of a pair walk the same weight blocks, so a view is dropped only when it misses both tiles; every thread of both CTAs
  // evaluates the same predicate (the coordinate chain of the 2 x 180 halo cells per view, ~1 us per tile) and so arrives at the
  // same list without any exchange.  Same bits as the dense loop.
  constexpr int SKIP_MAX_VIEWS = 128;   // 276 bytes of static shared memory: the kernel is within 1 KB of the 227 KB limit
  __shared__ uint32_t s_vmask[SKIP_MAX_VIEWS / 32];
  __shared__ uint16_t s_av[SKIP_MAX_VIEWS];
  __shared__ int s_nav;
  const int nv_part = vhi - vlo;
  const bool sparse = nv_part <= SKIP_MAX_VIEWS;
  if (tid < SKIP_MAX_VIEWS / 32) s_vmask[tid] = 0u;
  __syncthreads();
  if (sparse) {
    const int tile_p = p.tile0 + (item_peer % p.ntiles + p.p2p_tile_shift) % p.ntiles;
    const int tx0_p = (tile_p % p.tiles_x) * DT_X, ty0_p = (tile_p / p.tiles_x) * DT_Y;
    for (int w = tid; w < nv_part * 2 * DP_CELLS; w += PIPE_THREADS) {
      const int vi = w / (2 * DP_CELLS), r = w - vi * (2 * DP_CELLS), which = r / DP_CELLS, i = r - which * DP_CELLS;
      if ((*reinterpret_cast<volatile uint32_t*>(&s_vmask[vi >> 5]) >> (vi & 31)) & 1u) continue;   // already known to be seen
      const int gx = (which ? tx0_p : tx0) + i / DP_Y - 1, gy = (which ? ty0_p : ty0) + i % DP_Y - 1;
      if (gy < 0 || gy >= p.Hg || gx < 0 || gx >= p.Wg) continue;
      float Tm[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) Tm[q] = __ldg(p.T + 9 * (p.view0 + vlo + vi) + q);
      float sx, sy;
      norm_src_coord(Tm, gx, gy, p.Hg, p.Wg, sx, sy);
      if (make_taps(sx, sy, p.Hs, p.Ws).mask) atomicOr(&s_vmask[vi >> 5], 1u << (vi & 31));
    }
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    if (sparse)
      for (int vi = 0; vi < nv_part; ++vi)
        if ((s_vmask[vi >> 5] >> (vi & 31)) & 1u) s_av[n++] = (uint16_t)(vlo + vi);
    s_nav = sparse ? n : nv_part;
  }
  __syncthreads();
  const int nav = s_nav;                                       // active views of this CTA pair (and K part)
  auto view_of = [&](int idx) -> int { return sparse ? (int)s_av[idx] : vlo + idx; };   // idx-th active view (local index)
  const int npatch = nav * ncb;                               // patches = (active view, channel block) pairs of this CTA
  const int NK = npatch * 9 + (coord_here ? 1 : 0);
  const int wblk_coord = p.nviews * ncb * 9;

  if (warp == 0) tmem_alloc_2cta(&tmem_holder, CV_N);
  if (tid == 0) {
    const uint32_t nprod = NPROD / 32 + (prank == 0 ? 1 : 0);   // producer warps (+ the peer's relay on the leader)
    for (int i = 0; i < NPATCH; ++i) { mbar_init(&p_full[i], nprod); mbar_init(&p_empty[i], 1); }
    mbar_init(&c_full, nprod);
    for (int i = 0; i < DPB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 9) {
    // ===== my half of every weight block: two TMA boxes of 128 rows, completing on the leader's barrier =====
    if (lane == 0) {
      for (int kb = 0; kb < NK; ++kb) {
        const int s = kb % DPB;
        const int pkb = kb / 9;                                   // patch of this K block -> (active view, channel block), tap
        const int wblk = kb < npatch * 9 ? (view_of(pkb / ncb) * ncb + pkb % ncb) * 9 + (kb - 9 * pkb) : wblk_coord;
        if (kb >= DPB) mbar_wait(&b_empty[s], ((kb / DPB) - 1) & 1);
        if (prank == 0) mbar_expect_tx(&b_full[s], CV_B_BYTES);
        const uint32_t dst = smem_u32(sB + (size_t)s * CV_B2_BYTES), bar = map_to_cta(smem_u32(&b_full[s]), 0);
        tma_wbox_2cta(dst, &tmap_w, wblk * CV_N + 128 * (int)prank, bar);
        tma_wbox_2cta(dst + CV_B2_BYTES / 2, &tmap_w, wblk * CV_N + 256 + 128 * (int)prank, bar);
      }
    }
  } else if (warp == 8) {
    if (lane == 0 && prank == 0) {
      // ===== MMA issuer (leader, for both CTAs) =====
      const uint32_t idesc = idesc_bf16_f32(256, 256);
      long long wa = 0, wb = 0;
      const long long t_start = clock64();
      int kb = 0;
      auto mma_block = [&](uint64_t da) {   // one K block against weight stage kb % DPB
        const int sb = kb % DPB;
        const long long t1 = clock64();
        mbar_wait(&b_full[sb], (kb / DPB) & 1);
        wb += clock64() - t1;
        tc_fence_after();
        const uint32_t bbase = smem_u32(sB + (size_t)sb * CV_B2_BYTES);
        const uint64_t db0 = smem_desc_sw128(bbase), db1 = smem_desc_sw128(bbase + CV_B2_BYTES / 2);
#pragma unroll
        for (int k = 0; k < CV_KB / 16; ++k) {
          mma_bf16_2cta(tmem, da + 2 * k, db0 + 2 * k, idesc, kb > 0 || k > 0);
          mma_bf16_2cta(tmem + 256, da + 2 * k, db1 + 2 * k, idesc, kb > 0 || k > 0);
        }
        mma_commit_2cta(&b_empty[sb], 3);
        ++kb;
      };
      for (int pk = 0; pk < npatch; ++pk) {
        const int i = pk % NPATCH;
        const long long t0 = clock64();
        mbar_wait_cluster(&p_full[i], (pk / NPATCH) & 1);
        wa += clock64() - t0;
        const uint32_t pbase = smem_u32(sPatch + (size_t)i * DPATCH_BYTES);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap)   // tap = ky * 3 + kx: the patch cell of tile cell (0, 0) is (x = kx, y = ky)
          mma_block(smem_desc_sw128_sbo(pbase + (uint32_t)(((tap % 3) * DP_Y + tap / 3) * 128), D_SBO));
        mma_commit_2cta(&p_empty[i], 3);
      }
      if (coord_here) {
        const long long t0 = clock64();
        mbar_wait_cluster(&c_full, 0);
        wa += clock64() - t0;
        mma_block(smem_desc_sw128(smem_u32(sC)));
      }
      mma_commit_2cta(&acc_full, 3);
      if (p.stats) {
        unsigned long long* st = p.stats + 3 * (size_t)blockIdx.x;
        st[0] = (unsigned long long)wa; st[1] = (unsigned long long)wb; st[2] = (unsigned long long)(clock64() - t_start);
      }
    } else if (lane == 0) {
      // ===== relay (peer): forward "my patch is complete" to the leader's barrier with one remote arrive =====
      for (int pk = 0; pk < npatch; ++pk) {
        const int i = pk % NPATCH;
        mbar_wait(&p_full[i], (pk / NPATCH) & 1);
        mbar_arrive_cluster(map_to_cta(smem_u32(&p_full[i]), 0));
      }
      if (coord_here) {
        mbar_wait(&c_full, 0);
        mbar_arrive_cluster(map_to_cta(smem_u32(&c_full), 0));
      }
    }
  } else {
    // ===== producers (256 threads), then epilogue =====
    auto publish = [&](uint64_t* bar) {   // all of this warp's writes -> async proxy, then one arrival per warp
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    if constexpr (!UPS) {
    auto compute_taps = [&](int vl, ConvTap* dst) {   // patch cell i = (px = i / 10, py = i % 10) <-> grid cell (tx0 + px - 1, ty0 + py - 1)
      const int v = p.view0 + vl;
      for (int i = tid; i < DP_CELLS; i += NPROD) {
        const int gx = tx0 + i / DP_Y - 1, gy = ty0 + i % DP_Y - 1;
        ConvTap t;
        t.mask = 0; t.x0 = t.y0 = 0; t.pad = 0;
        t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
        if (gy >= 0 && gy < p.Hg && gx >= 0 && gx < p.Wg) {
          float Tm[9];
#pragma unroll
          for (int q = 0; q < 9; ++q) Tm[q] = __ldg(p.T + 9 * v + q);
          float sx, sy;
          norm_src_coord(Tm, gx, gy, p.Hg, p.Wg, sx, sy);
          const Taps tp = make_taps(sx, sy, p.Hs, p.Ws);
#pragma unroll
          for (int k = 0; k < 4; ++k) t.w[k] = tp.w[k];
          t.x0 = tp.x0; t.y0 = tp.y0; t.mask = tp.mask;
        }
        dst[i] = t;
      }
    };
    uint4 tv[DP_ROUNDS][4];   // gathers in flight for the next patch
    auto issue_gathers = [&](int vl, int cb, const ConvTap* taps) {
      const __nv_bfloat16* fv = p.feats + (int64_t)b * p.fB + (int64_t)vl * p.fV + cb * CV_KB;
#pragma unroll
      for (int j = 0; j < DP_ROUNDS; ++j) {
        const int task = tid + j * NPROD;
        const int cell = task >> 3, ch = task & 7;
        if (task < DP_TASKS) {
          const ConvTap t = taps[cell];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
            tv[j][k] = (t.mask >> k & 1) ? __ldg(reinterpret_cast<const uint4*>(fv + (int64_t)yy * p.fH + (int64_t)xx * p.fW + ch * 8))
                                         : make_uint4(0, 0, 0, 0);
          }
        }
      }
    };
    auto blend_patch = [&](const ConvTap* taps, uint8_t* patch) {   // patch buffers are 1024-aligned: swizzle phase = cell & 7
#pragma unroll
      for (int j = 0; j < DP_ROUNDS; ++j) {
        const int task = tid + j * NPROD;
        const int cell = task >> 3, ch = task & 7;
        if (task < DP_TASKS) {
          const ConvTap t = taps[cell];
          const unsigned* a = &tv[j][0].x; const unsigned* bq = &tv[j][1].x; const unsigned* c = &tv[j][2].x; const unsigned* d = &tv[j][3].x;
          unsigned o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float lo = __fmaf_rn(bf_lo(d[q]), t.w[3], __fmaf_rn(bf_lo(c[q]), t.w[2], __fmaf_rn(bf_lo(bq[q]), t.w[1], __fmul_rn(bf_lo(a[q]), t.w[0]))));
            const float hi = __fmaf_rn(bf_hi(d[q]), t.w[3], __fmaf_rn(bf_hi(c[q]), t.w[2], __fmaf_rn(bf_hi(bq[q]), t.w[1], __fmul_rn(bf_hi(a[q]), t.w[0]))));
            o[q] = pack_bf2(lo, hi);
          }
          *reinterpret_cast<uint4*>(patch + cell * 128 + ((ch ^ (cell & 7)) << 4)) = t.mask ? make_uint4(o[0], o[1], o[2], o[3]) : make_uint4(0, 0, 0, 0);
        }
      }
    };

    // tap buffers alternate with the INDEX of the active view (two consecutive active views may both be even)
    if (npatch > 0) {
      compute_taps(view_of(0), sT);
      producers_sync();
      issue_gathers(view_of(0), 0, sT);
    }
    for (int pk = 0; pk < npatch; ++pk) {
      const int vidx = pk / ncb;
      const int i = pk % NPATCH;
      if (pk >= NPATCH) mbar_wait(&p_empty[i], ((pk / NPATCH) - 1) & 1);   // the nine K blocks that read this buffer are done
      blend_patch(sT + (vidx & 1) * DP_CELLS, sPatch + (size_t)i * DPATCH_BYTES);    // consumes the gathers issued one patch earlier
      publish(&p_full[i]);
      if (pk + 1 < npatch) {             // the next patch's gathers fly while the tensor core works on this one
        const int nidx = (pk + 1) / ncb, ncbi = (pk + 1) % ncb, nvl = view_of(nidx);
        if (ncbi == 0) {                 // next patch opens a new view: its taps go to the other buffer
          producers_sync();              // everybody is done reading the buffer of the view before this one
          compute_taps(nvl, sT + (nidx & 1) * DP_CELLS);
          producers_sync();
        }
        issue_gathers(nvl, ncbi, sT + (nidx & 1) * DP_CELLS);
      }
    }
    } else {
    // ----- UPS: the patch is warp(up-sample(low-res)) composed per halo cell into a 3 x 3 low-res window (ConvTapUp) -----
    static_assert(DP_ROUNDS == 6, "the round schedule below is written out for six rounds");
    ConvTapUp* sU = reinterpret_cast<ConvTapUp*>(sT);
    const float scx = upsample_scale(p.wl, p.Ws), scy = upsample_scale(p.hl, p.Hs);
    auto compute_taps_up = [&](int vl, ConvTapUp* dst) {
      const int v = p.view0 + vl;
      for (int i = tid; i < DP_CELLS; i += NPROD) {
        const int gx = tx0 + i / DP_Y - 1, gy = ty0 + i % DP_Y - 1;
        ConvTapUp t;
        t.mask = 0; t.xy = 0;
        t.wx[0] = t.wx[1] = t.wx[2] = t.wy[0] = t.wy[1] = t.wy[2] = 0.f;
        if (gy >= 0 && gy < p.Hg && gx >= 0 && gx < p.Wg) {
          float Tm[9];
#pragma unroll
          for (int q = 0; q < 9; ++q) Tm[q] = __ldg(p.T + 9 * v + q);
          float sx, sy;
          norm_src_coord(Tm, gx, gy, p.Hg, p.Wg, sx, sy);
          t = compose_cell(sx, sy, p.Hs, p.Ws, p.hl, p.wl, scx, scy);
        }
        dst[i] = t;
      }
    };
    // round j of a patch: task = tid + j * 256 -> (halo cell, 16-byte channel chunk); nine masked gathers, then the blend
    auto issue = [&](uint4 (&g)[9], int vl, int cb, int j, const ConvTapUp* taps) {
      const int task = tid + j * NPROD;
      if (task < DP_TASKS) {
        const int cell = task >> 3, ch = task & 7;
        const ConvTapUp t = taps[cell];
        const __nv_bfloat16* f0 = p.feats + (int64_t)b * p.fB + (int64_t)vl * p.fV + cb * CV_KB + ch * 8 +
                                  (int64_t)(t.xy >> 16) * p.fH + (int64_t)(t.xy & 0xffff) * p.fW;
#pragma unroll
        for (int k = 0; k < 9; ++k)
          g[k] = (t.mask >> k & 1) ? __ldg(reinterpret_cast<const uint4*>(f0 + (int64_t)(k / 3) * p.fH + (int64_t)(k % 3) * p.fW))
                                   : make_uint4(0, 0, 0, 0);
      }
    };
    auto blend = [&](const uint4 (&g)[9], int j, const ConvTapUp* taps, uint8_t* patch) {
      const int task = tid + j * NPROD;
      if (task < DP_TASKS) {
        const int cell = task >> 3, ch = task & 7;
        const ConvTapUp t = taps[cell];
        float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 9; ++k) {     // ascending low-res texel; a dead slot holds zeros and a zero weight
          const float wk = __fmul_rn(t.wy[k / 3], t.wx[k % 3]);
          const unsigned* q = &g[k].x;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            lo[c] = __fmaf_rn(bf_lo(q[c]), wk, lo[c]);
            hi[c] = __fmaf_rn(bf_hi(q[c]), wk, hi[c]);
          }
        }
        *reinterpret_cast<uint4*>(patch + cell * 128 + ((ch ^ (cell & 7)) << 4)) =
            t.mask ? make_uint4(pack_bf2(lo[0], hi[0]), pack_bf2(lo[1], hi[1]), pack_bf2(lo[2], hi[2]), pack_bf2(lo[3], hi[3]))
                   : make_uint4(0, 0, 0, 0);
      }
    };
    uint4 g0[9], g1[9];   // two rounds of gathers in flight: the low-res map is L2-resident (103 MB bf16 on Wildtrack) and a patch's
                          // 180 cells share ~60 texel rows, so two rounds cover the latency; nine K blocks of MMA hide the rest
    if (npatch > 0) {
      compute_taps_up(view_of(0), sU);
      producers_sync();
      issue(g0, view_of(0), 0, 0, sU);
    }
    for (int pk = 0; pk < npatch; ++pk) {
      const int vidx = pk / ncb, cb = pk % ncb, vl = view_of(vidx);
      const int i = pk % NPATCH;
      const ConvTapUp* taps = sU + (vidx & 1) * DP_CELLS;          // tap buffers alternate with the INDEX of the active view
      uint8_t* patch = sPatch + (size_t)i * DPATCH_BYTES;
      if (pk >= NPATCH) mbar_wait(&p_empty[i], ((pk / NPATCH) - 1) & 1);   // the nine K blocks that read this buffer are done
      issue(g1, vl, cb, 1, taps); blend(g0, 0, taps, patch);
      issue(g0, vl, cb, 2, taps); blend(g1, 1, taps, patch);
      issue(g1, vl, cb, 3, taps); blend(g0, 2, taps, patch);
      issue(g0, vl, cb, 4, taps); blend(g1, 3, taps, patch);
      issue(g1, vl, cb, 5, taps); blend(g0, 4, taps, patch);
      const bool more = pk + 1 < npatch;
      const int nidx = (pk + 1) / ncb, ncbi = (pk + 1) % ncb;
      if (more && ncbi != 0) issue(g0, vl, ncbi, 0, taps);                // same view: the next patch's first round flies already
      blend(g1, 5, taps, patch);
      publish(&p_full[i]);
      if (more && ncbi == 0) {            // next patch opens a new view: its taps go to the other buffer (last read one view ago,
        const int nvl = view_of(nidx);    // and every thread has passed this view's barrier since)
        compute_taps_up(nvl, sU + (nidx & 1) * DP_CELLS);
        producers_sync();
        issue(g0, nvl, 0, 0, sU + (nidx & 1) * DP_CELLS);
      }
    }
    }
    if (coord_here) {  // the two coord_map channels: all 9 taps x 2 channels in one K block (k = tap*2 + ch), canonical layout
      const int m = tid & 127, half = tid >> 7, mx = m / DT_Y, my = m % DT_Y;
      unsigned short vals[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) vals[k] = 0;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int gy = ty0 + my + tap / 3 - 1, gx = tx0 + mx + tap % 3 - 1;
        if (gy >= 0 && gy < p.Hg && gx >= 0 && gx < p.Wg) {
          vals[tap * 2] = __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p.coord_x + gx)));
          vals[tap * 2 + 1] = __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p.coord_y + gy)));
        }
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = 4 * half + cc;     // chunks 0..3 carry k = 0..31 (18 used), chunks 4..7 are zero
        uint4 r = make_uint4(0, 0, 0, 0);
        if (half == 0) {
          r.x = vals[cc * 8 + 0] | ((unsigned)vals[cc * 8 + 1] << 16); r.y = vals[cc * 8 + 2] | ((unsigned)vals[cc * 8 + 3] << 16);
          r.z = vals[cc * 8 + 4] | ((unsigned)vals[cc * 8 + 5] << 16); r.w = vals[cc * 8 + 6] | ((unsigned)vals[cc * 8 + 7] << 16);
        }
        *reinterpret_cast<uint4*>(sC + sw128_offset(m, c)) = r;
      }
      publish(&c_full);
    }
    // ----- epilogue: thread = GEMM row = TMEM lane (quarter warp % 4); warps 0-3 take columns 0-255, warps 4-7 256-511 -----
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int er = (warp & 3) * 32 + lane, ex = er / DT_Y, ey = er % DT_Y;
    const int gy = ty0 + ey, gx = tx0 + ex;
    const bool inside = real_tile && gy < p.Hg && gx < p.Wg;
    const int64_t ooff = (int64_t)b * p.oB + (int64_t)gy * p.oH + (int64_t)gx * p.oW;
    float* p2p_dst = nullptr;
    if (p.p2p_n > 0 && inside) {
      int o = 0;
      while (o + 1 < p.p2p_n && gy >= p.p2p_row0[o + 1]) ++o;
      p2p_dst = reinterpret_cast<float*>(p.p2p_inbox[o]) + (int64_t)p.p2p_src * p.p2p_src_pitch + (int64_t)b * p.p2p_b_pitch +
                ((int64_t)(gy - p.p2p_row0[o]) * p.Wg + gx) * CV_N;
    }
    const int cbeg = (warp >> 2) * 256;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + 256; c0 += 32) {
      float vv[32];
      if (NK > 0) {
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, vv);
      } else {   // no K block at all (every view skipped and no coord block here): the sum is zero, TMEM was never written
#pragma unroll
        for (int i = 0; i < 32; ++i) vv[i] = 0.f;
      }
      if (partial) {
        float4* o = reinterpret_cast<float4*>(p.scratch + ((size_t)slot * 128 + er) * CV_N + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
      } else if (p.p2p_n > 0) {
        if (inside) {   // raw partial sums into the inbox of the rank that owns this grid row (peer memory over NVLink)
          float4* o = reinterpret_cast<float4*>(p2p_dst + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
        }
      } else if (inside) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (p.bias) vv[i] += __ldg(p.bias + c0 + i);
          if (p.relu) vv[i] = fmaxf(vv[i], 0.f);
        }
        if (p.out_bf16) {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = make_uint4(pack_bf2(vv[8 * i], vv[8 * i + 1]), pack_bf2(vv[8 * i + 2], vv[8 * i + 3]),
                              pack_bf2(vv[8 * i + 4], vv[8 * i + 5]), pack_bf2(vv[8 * i + 6], vv[8 * i + 7]));
        } else {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + ooff + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 0) tmem_dealloc_2cta(tmem, CV_N);
}

// out[item's cells] = [relu](bias + sum over the item's K-split parts, in part order) for the tail items of a launch.
__global__ void __launch_bounds__(256) conv1_tail_reduce_kernel(const ConvParams p) {
  const int item = p.n_full + blockIdx.x;
  const int b = item / p.ntiles, tile = p.tile0 + item % p.ntiles;
  const int tx0 = (tile % p.tiles_x) * CT_W, ty0 = (tile / p.tiles_x) * CT_H;
  // slot of (tail item t, part s): single-CTA launches t * split + s; pair launches ((t/2) * split + s) * 2 + t % 2
  const bool pair = p.pair_mode != 0;
  const size_t slot0 = pair ? (size_t)(blockIdx.x >> 1) * p.split * 2 + (blockIdx.x & 1) : (size_t)blockIdx.x * p.split;
  const size_t slot_step = pair ? 2 : 1;
  const float* base = p.scratch + slot0 * 128 * CV_N;
  // grid.y = 8 slices of 16 cells: 49 tail items alone would leave two thirds of the SMs idle (55 us instead of ~10)
  for (int i = threadIdx.x; i < 16 * CV_N / 4; i += blockDim.x) {
    const int row = blockIdx.y * 16 + i / (CV_N / 4), c = (i % (CV_N / 4)) * 4;
    // GEMM row -> tile cell: y-major 8 x 16 in the copying kernels, x-major (16 wide, 8 high) in the direct kernel; the tile is 16 x 8 cells either way
    const int gy = ty0 + (p.pair_mode == 2 ? row % DT_Y : row / CT_W), gx = tx0 + (p.pair_mode == 2 ? row / DT_Y : row % CT_W);
    if (gy >= p.Hg || gx >= p.Wg) continue;
    float4 a = *reinterpret_cast<const float4*>(base + (size_t)row * CV_N + c);
    for (int s = 1; s < p.split; ++s) {
      const float4 q = *reinterpret_cast<const float4*>(base + ((size_t)s * slot_step * 128 + row) * CV_N + c);
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    if (p.bias) { a.x += __ldg(p.bias + c); a.y += __ldg(p.bias + c + 1); a.z += __ldg(p.bias + c + 2); a.w += __ldg(p.bias + c + 3); }
    if (p.relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    const int64_t ooff = (int64_t)b * p.oB + (int64_t)gy * p.oH + (int64_t)gx * p.oW + c;
    if (p.out_bf16) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff) = make_uint2(pack_bf2(a.x, a.y), pack_bf2(a.z, a.w));
    else *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + ooff) = a;
  }
}

// Tail decomposition of a launch with `nitems` work items of `nviews` views on `num_sms` SMs: how many leftover items and
// into how many K parts each is split (1 = no split).
static void tail_split(int nitems, int nviews, int num_sms, int& n_tail, int& split) {
  n_tail = nitems % num_sms;
  split = 1;
  if (!n_tail) return;
  // cost of the last wave in tile times: the parts run in ceil(n_tail * s / SMs) rounds of 1/s tile each, plus a small
  // fixed cost per part (scratch epilogue + reduce); s = 1 costs one whole round
  double best = 1.0;
  for (int s = 2; s <= std::min(nviews, 8); ++s) {
    const double cost = (double)(((int64_t)n_tail * s + num_sms - 1) / num_sms) / s + 0.02 * s;
    if (cost < best - 1e-9) { best = cost; split = s; }
  }
  if (split == 1) n_tail = 0;
}

// Work decomposition of a launch (ConvParams): n_full whole items, `split` K parts for each of the rest, scratch slots needed.
// pair: CTA-pair kernels -- the unit is a pair of items on a pair of SMs, n_full is even.
struct Decomp {
  int n_full, split;
  int64_t slots;
};
static Decomp decompose(int nitems, int nviews, int num_sms, bool pair) {
  Decomp d{nitems, 1, 0};
  if (!pair) {
    int n_tail, split;
    tail_split(nitems, nviews, num_sms, n_tail, split);
    d.split = split; d.n_full = nitems - n_tail; d.slots = (int64_t)n_tail * split;
    return d;
  }
  const int npairs = (nitems + 1) / 2;
  int tail_pairs, split;
  tail_split(npairs, nviews, std::max(1, num_sms / 2), tail_pairs, split);
  if (split > 1) { d.split = split; d.n_full = 2 * (npairs - tail_pairs); d.slots = (int64_t)tail_pairs * split * 2; }
  else { d.n_full = nitems / 2 * 2; }   // an odd last item runs as a "tail pair" of one part with a padding partner
  return d;
}

// Development knobs, read ONCE (first launch) and only by the development build; the product library has no knobs.
struct Conv1Knobs {
  int pair_mode = 2;      // 0 = single-CTA kernel, 1 = CTA pairs with copied tap stages, 2 = CTA pairs reading the taps in place (default)
  int cluster = 1;        // MVWARP_CONV1_VARIANT=2|4: cluster multicast of the weight blocks (single-CTA kernel)
  bool simple = false;    // MVWARP_CONV1_VARIANT=s: the unpipelined validation kernel
  bool nosplit = false;   // MVWARP_CONV1_NOSPLIT: no K split of the last partial wave
  bool stats = false;     // MVWARP_CONV1_STATS: synchronous, prints where the MMA thread waits
};
static const Conv1Knobs& conv1_knobs() {
  static const Conv1Knobs k = [] {
    Conv1Knobs r;
#ifdef MVWARP_DEV
    const char* e = getenv("MVWARP_CONV1_PAIR");
    const char* v = getenv("MVWARP_CONV1_VARIANT");
    if (e && e[0] >= '0' && e[0] <= '2') r.pair_mode = e[0] - '0';
    if (v) {
      r.pair_mode = 0;
      if (v[0] == 's') r.simple = true;
      else if (v[0] >= '1' && v[0] <= '4') r.cluster = v[0] - '0';
    }
    r.nosplit = getenv("MVWARP_CONV1_NOSPLIT") != nullptr;
    r.stats = getenv("MVWARP_CONV1_STATS") != nullptr;
#endif
    return r;
  }();
  return k;
}
static int pair_mode_wanted() { return conv1_knobs().pair_mode; }

}  // namespace mvw

extern "C" int64_t mvwarp_warp_conv1_scratch_bytes(const mvwarp_plan_t* plan, int nviews, int B, int row0, int nrows) {
  using namespace mvw;
  if (!plan || nviews <= 0 || B <= 0) return 0;
  if (nrows <= 0) { row0 = 0; nrows = plan->Hg; }
  const int tiles_x = (plan->Wg + CT_W - 1) / CT_W;
  const int64_t nitems = (int64_t)((nrows + CT_H - 1) / CT_H) * tiles_x * B;
  if (nitems > 0x7fffffff) return 0;
  const Decomp d = decompose((int)nitems, nviews, plan->num_sms, pair_mode_wanted() != 0);
  return d.slots * 128 * CV_N * (int64_t)sizeof(float);
}

extern "C" int mvwarp_warp_conv1_fwd(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats,
                                     const int64_t feat_strides[5], const void* wpack, const float* bias, float* out,
                                     const int64_t out_strides[4], int B, int C, int relu, int with_coord, int row0, int nrows,
                                     void* stream) {
  return mvwarp_warp_conv1_fwd_ex(plan, view0, nviews, feats, feat_strides, wpack, bias, out, MVWARP_F32, out_strides, B, C, relu,
                                  with_coord, row0, nrows, nullptr, 0, stream);
}

namespace mvw {
// out[b][i] = [relu](bias[i % 512] + sum over slots s (in slot order) of inbox[s][b][i]),  i < n_per_b = rows * Wg * 512:
// the owner's side of the fused reduce-scatter.  One pass: nslots reads and one write per element, 128-bit accesses.
template <typename TO>
__global__ void __launch_bounds__(256) inbox_reduce_kernel(const float* __restrict__ inbox, int nslots, int64_t slot_pitch, int64_t b_pitch,
                                                           int64_t n_per_b, const float* __restrict__ bias, int relu, TO* __restrict__ out) {
  const int b = blockIdx.y;
  const float* src = inbox + (int64_t)b * b_pitch;
  TO* dst = out + (int64_t)b * n_per_b;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n_per_b; i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(src + i));
    for (int s = 1; s < nslots; ++s) {
      const float4 q = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)s * slot_pitch + i));
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    if (bias) {
      const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + (i & 511)));
      a.x += bq.x; a.y += bq.y; a.z += bq.z; a.w += bq.w;
    }
    if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    if (sizeof(TO) == 4) *reinterpret_cast<float4*>(dst + i) = a;
    else *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf2(a.x, a.y), pack_bf2(a.z, a.w));
  }
}
}  // namespace mvw

extern "C" int mvwarp_inbox_reduce(const float* inbox, int nslots, int64_t slot_pitch, int64_t b_pitch, int B, int64_t n_per_b,
                                   const float* bias, int relu, void* out, int out_dtype, int device, void* stream) {
  using namespace mvw;
  if (!inbox || !out || nslots < 1 || B < 1 || B > 65535 || n_per_b < 0 || n_per_b % 512 || slot_pitch % 4 || b_pitch % 4 ||
      ((reinterpret_cast<uintptr_t>(inbox) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15u) ||
      (out_dtype != MVWARP_F32 && out_dtype != MVWARP_BF16)) {
    set_error("inbox_reduce: bad argument (elements per batch item must be a multiple of 512, pointers 16-byte aligned)");
    return MVWARP_EINVAL;
  }
  if (n_per_b == 0) return MVWARP_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device && cudaSetDevice(device) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  const int64_t want = (n_per_b / 4 + 255) / 256;
  dim3 grid((unsigned)std::min<int64_t>(want, 148 * 16), B);
  if (out_dtype == MVWARP_F32)
    inbox_reduce_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(inbox, nslots, slot_pitch, b_pitch, n_per_b, bias, relu, (float*)out);
  else
    inbox_reduce_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(inbox, nslots, slot_pitch, b_pitch, n_per_b, bias, relu,
                                                                                 (__nv_bfloat16*)out);
  const cudaError_t e = cudaGetLastError();
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "inbox_reduce launch");
}

namespace {
struct P2PArgs {
  int n = 0, src = 0, band_rows_max = 0;
  const uint64_t* inbox = nullptr;
  const int32_t* row0 = nullptr;
};
int conv1_launch(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats, const int64_t feat_strides[5], const void* wpack,
                 const float* bias, void* out, int out_dtype, const int64_t out_strides[4], int B, int C, int relu, int with_coord, int row0,
                 int nrows, void* scratch, int64_t scratch_bytes, const P2PArgs& p2p, void* stream);
}  // namespace

extern "C" int mvwarp_warp_conv1_fwd_ex(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats,
                                        const int64_t feat_strides[5], const void* wpack, const float* bias, void* out, int out_dtype,
                                        const int64_t out_strides[4], int B, int C, int relu, int with_coord, int row0, int nrows,
                                        void* scratch, int64_t scratch_bytes, void* stream) {
  return conv1_launch(plan, view0, nviews, feats, feat_strides, wpack, bias, out, out_dtype, out_strides, B, C, relu, with_coord, row0, nrows,
                      scratch, scratch_bytes, P2PArgs{}, stream);
}

extern "C" int mvwarp_warp_conv1_fwd_p2p(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats,
                                         const int64_t feat_strides[5], const void* wpack, int B, int C, int with_coord, int nranks,
                                         int my_rank, const uint64_t* peer_inbox, const int32_t* band_row0, int band_rows_max, void* stream) {
  using namespace mvw;
  if (!plan || !peer_inbox || !band_row0 || nranks < 1 || nranks > 8 || my_rank < 0 || my_rank >= nranks || band_rows_max <= 0) {
    set_error("warp_conv1_fwd_p2p: bad rank / inbox arguments (1..8 ranks)");
    return MVWARP_EINVAL;
  }
  if (band_row0[0] != 0 || band_row0[nranks] != plan->Hg) { set_error("warp_conv1_fwd_p2p: the row bands must cover [0, Hg)"); return MVWARP_EINVAL; }
  for (int r = 0; r < nranks; ++r) {
    if (band_row0[r + 1] < band_row0[r] || band_row0[r + 1] - band_row0[r] > band_rows_max || (peer_inbox[r] & 15u) || !peer_inbox[r]) {
      set_error("warp_conv1_fwd_p2p: bad row band or inbox pointer");
      return MVWARP_EINVAL;
    }
  }
  P2PArgs a;
  a.n = nranks; a.src = my_rank; a.band_rows_max = band_rows_max; a.inbox = peer_inbox; a.row0 = band_row0;
  // `out` is unused in this mode; the shared argument checks want a non-null, 16-byte aligned pointer and channels-last strides
  const int64_t os[4] = {(int64_t)plan->Hg * plan->Wg * 512, 1, (int64_t)plan->Wg * 512, 512};
  return conv1_launch(plan, view0, nviews, feats, feat_strides, wpack, nullptr, reinterpret_cast<void*>(uintptr_t(16)), MVWARP_F32, os, B, C, 0,
                      with_coord, 0, 0, nullptr, 0, a, stream);
}

namespace {
int conv1_launch(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats, const int64_t feat_strides[5], const void* wpack,
                 const float* bias, void* out, int out_dtype, const int64_t out_strides[4], int B, int C, int relu, int with_coord, int row0,
                 int nrows, void* scratch, int64_t scratch_bytes, const P2PArgs& p2p, void* stream) {
  using namespace mvw;
  if (out_dtype != MVWARP_F32 && out_dtype != MVWARP_BF16) { set_error("warp_conv1_fwd: unknown output dtype"); return MVWARP_EINVAL; }
  if (!plan || !feats || !feat_strides || !wpack || !out || !out_strides) { set_error("warp_conv1_fwd: null argument"); return MVWARP_EINVAL; }
  if (B <= 0 || nviews <= 0 || view0 < 0 || view0 > plan->V || nviews > plan->V - view0) { set_error("warp_conv1_fwd: bad B / view range"); return MVWARP_EINVAL; }
  if (!plan->inline_coords) { set_error("warp_conv1_fwd: needs a plan made from matrices (not from a sampling grid)"); return MVWARP_EUNSUPPORTED; }
  // fused up-sample plans (mvwarp_plan_create_upsampled): `feats` is the low-res map [B,V,C,hl,wl]; the producers compose the two
  // bilinear steps per halo cell into a 3 x 3 low-res window, which holds for an up-sampling ratio only
  if (plan->upsampled && (plan->hl > plan->Hs || plan->wl > plan->Ws || plan->hl > 32767 || plan->wl > 32767)) {
    set_error("warp_conv1_fwd: a fused up-sample plan must up-sample (low-res size <= up-sampled size, below 32768)");
    return MVWARP_EUNSUPPORTED;
  }
  if (C <= 0 || C % CV_KB || feat_strides[2] != 1 || (reinterpret_cast<uintptr_t>(feats) & 15u) || feat_strides[0] % 8 ||
      feat_strides[1] % 8 || feat_strides[3] % 8 || feat_strides[4] % 8) {
    set_error("warp_conv1_fwd: feats must be bf16 channels-last with C % 64 == 0 and 16-byte aligned texel rows");
    return MVWARP_EUNSUPPORTED;
  }
  const int oal = out_dtype == MVWARP_BF16 ? 8 : 4;
  if (out_strides[1] != 1 || out_strides[0] % oal || out_strides[2] % oal || out_strides[3] % oal || (reinterpret_cast<uintptr_t>(out) & 15u)) {
    set_error("warp_conv1_fwd: out must be channels-last [B,512,Hg,Wg] with 16-byte aligned cell rows");
    return MVWARP_EUNSUPPORTED;
  }
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != plan->device && cudaSetDevice(plan->device) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  ConvParams p{};
  p.T = plan->d_T; p.coord_x = plan->d_coord_x; p.coord_y = plan->d_coord_y;
  p.Hs = plan->Hs; p.Ws = plan->Ws; p.Hg = plan->Hg; p.Wg = plan->Wg;
  p.hl = plan->upsampled ? plan->hl : 0; p.wl = plan->upsampled ? plan->wl : 0;
  p.view0 = view0; p.nviews = nviews; p.C = C; p.B = B;
  p.feats = (const __nv_bfloat16*)feats;
  p.fB = feat_strides[0]; p.fV = feat_strides[1]; p.fH = feat_strides[3]; p.fW = feat_strides[4];
  p.wpack = (const __nv_bfloat16*)wpack; p.bias = bias; p.out = out; p.out_bf16 = out_dtype == MVWARP_BF16;
  p.oB = out_strides[0]; p.oH = out_strides[2]; p.oW = out_strides[3];
  p.relu = relu; p.with_coord = with_coord;
  p.tiles_x = (plan->Wg + CT_W - 1) / CT_W; p.tiles_y = (plan->Hg + CT_H - 1) / CT_H;
  if (nrows <= 0) { row0 = 0; nrows = plan->Hg; }
  if (row0 < 0 || row0 % CT_H || row0 + nrows > plan->Hg || (nrows % CT_H && row0 + nrows != plan->Hg)) {
    if (prev >= 0 && prev != plan->device) cudaSetDevice(prev);
    set_error("warp_conv1_fwd: a row band must start on a multiple of 8 rows and end on one (or at the last row)");
    return MVWARP_EINVAL;
  }
  p.tile0 = (row0 / CT_H) * p.tiles_x;
  p.ntiles = ((nrows + CT_H - 1) / CT_H) * p.tiles_x;
  if ((int64_t)p.ntiles * B > 0x7fffffff) {
    if (prev >= 0 && prev != plan->device) cudaSetDevice(prev);
    set_error("warp_conv1_fwd: too many tiles");
    return MVWARP_EUNSUPPORTED;
  }
  p.nitems = p.ntiles * B; p.n_full = p.nitems; p.split = 1; p.scratch = nullptr;
  cudaError_t e;
  const Conv1Knobs& knobs = conv1_knobs();
#ifdef MVWARP_DEV
  if (knobs.simple && out_dtype == MVWARP_F32) {
    const int smem = 16384 + (int)CV_B_BYTES + CP_CELLS * 128 + CP_CELLS * (int)sizeof(ConvTap) + 1024;
    e = cudaFuncSetAttribute(warp_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) {
      warp_conv1_kernel<<<dim3(p.ntiles, 1, B), 128, smem, (cudaStream_t)stream>>>(p);
      e = cudaGetLastError();
    }
  } else
#endif
  {
    const int cl = knobs.cluster;
    const int pmode = cl == 1 ? knobs.pair_mode : 0;
    const bool pair = pmode != 0;
    CUtensorMap tmap_w;
    memset(&tmap_w, 0, sizeof(tmap_w));
    if (pair) {
      const uint64_t nblocks = (uint64_t)nviews * (C / CV_KB) * 9 + (with_coord ? 1 : 0);
      const int wrc = make_weight_tensor_map(&tmap_w, wpack, nblocks);
      if (wrc) { if (prev >= 0 && prev != plan->device) cudaSetDevice(prev); return wrc; }
    }
    p.pair_mode = pmode;
    if (plan->upsampled && pmode != 2) {
      if (prev >= 0 && prev != plan->device) cudaSetDevice(prev);
      set_error("warp_conv1_fwd: fused up-sample plans run on the direct kernel only (unset MVWARP_CONV1_PAIR / _VARIANT)");
      return MVWARP_EUNSUPPORTED;
    }
    if (p2p.n > 0) {
      if (pmode != 2) {
        if (prev >= 0 && prev != plan->device) cudaSetDevice(prev);
        set_error("warp_conv1_fwd_p2p: only the direct kernel writes to peer inboxes (unset MVWARP_CONV1_PAIR / _VARIANT)");
        return MVWARP_EUNSUPPORTED;
      }
      p.p2p_n = p2p.n; p.p2p_src = p2p.src;
      for (int r = 0; r <= p2p.n; ++r) p.p2p_row0[r] = p2p.row0[r];
      for (int r = 0; r < p2p.n; ++r) p.p2p_inbox[r] = p2p.inbox[r];
      p.p2p_tile_shift = (p2p.row0[(p2p.src + 1) % p2p.n] / CT_H) * p.tiles_x % p.ntiles;   // first tile of the next rank's band
      p.p2p_b_pitch = (long long)p2p.band_rows_max * plan->Wg * CV_N;
      p.p2p_src_pitch = p.p2p_b_pitch * B;
    }
    if (pair) p.n_full = p.nitems / 2 * 2;
    if (cl == 1 && scratch && !knobs.nosplit) {   // K-split of the last partial wave (see ConvParams)
      const Decomp d = decompose(p.nitems, nviews, plan->num_sms, pair);
      if (d.split > 1 && scratch_bytes >= d.slots * 128 * CV_N * (int64_t)sizeof(float) && !(reinterpret_cast<uintptr_t>(scratch) & 15u)) {
        p.split = d.split; p.n_full = d.n_full; p.scratch = reinterpret_cast<float*>(scratch);
      }
    }
#ifdef MVWARP_DEV
    const int smem = PA * 16384 + PB * (int)CV_B_BYTES + CP_CELLS * 128 + 2 * CP_CELLS * (int)sizeof(ConvTap) + 1024;
    auto launch = [&](auto kernel, int cluster) -> cudaError_t {
      cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (err != cudaSuccess) return err;
      cudaLaunchConfig_t cfg{};
      const int n_tail = p.nitems - p.n_full;
      if (p.pair_mode) cfg.gridDim = dim3(p.n_full + (n_tail + 1) / 2 * 2 * p.split, 1, 1);   // clusters: n_full / 2 whole pairs + tail pairs x parts
      else cfg.gridDim = dim3(p.split > 1 ? p.n_full + n_tail * p.split : (p.nitems + cluster - 1) / cluster * cluster, 1, 1);
      cfg.blockDim = dim3(PIPE_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, kernel, tmap_w, p);
    };
#endif
    unsigned long long* d_stats = nullptr;
    const size_t nst = (size_t)3 * ((size_t)p.nitems * 8 + 8);
    if (knobs.stats) {   // development only: synchronous, prints where the MMA thread waits
      cudaMalloc(&d_stats, nst * sizeof(unsigned long long));
      cudaMemset(d_stats, 0, nst * sizeof(unsigned long long));
      p.stats = d_stats;
    }
    if (pmode == 2) {
      const int smem_d = NPATCH * DPATCH_BYTES + 16384 + DPB * (int)CV_B2_BYTES + 2 * DP_CELLS * (int)sizeof(ConvTap) + 1024;
      auto kern = plan->upsampled ? warp_conv1_direct_kernel<true> : warp_conv1_direct_kernel<false>;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_d);
      if (e == cudaSuccess) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(p.n_full + (p.nitems - p.n_full + 1) / 2 * 2 * p.split, 1, 1);
        cfg.blockDim = dim3(PIPE_THREADS);
        cfg.dynamicSmemBytes = smem_d;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, tmap_w, p);
      }
    }
#ifdef MVWARP_DEV
    else if (pair) e = launch(warp_conv1_pipe_kernel<1, true>, 2);
    else if (cl == 1) e = launch(warp_conv1_pipe_kernel<1, false>, 1);
    else if (cl == 4) e = launch(warp_conv1_pipe_kernel<4, false>, 4);
    else e = launch(warp_conv1_pipe_kernel<2, false>, 2);
#else
    else { set_error("warp_conv1_fwd: this build carries the direct kernel only"); e = cudaErrorNotSupported; }
#endif
    if (e == cudaSuccess && p.split > 1) {
      conv1_tail_reduce_kernel<<<dim3(p.nitems - p.n_full, 8), 256, 0, (cudaStream_t)stream>>>(p);
      e = cudaGetLastError();
    }
    if (d_stats) {
      std::vector<unsigned long long> h(nst);
      cudaMemcpy(h.data(), d_stats, nst * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      double a = 0, bq = 0, t = 0; int n = 0;
      for (size_t i = 0; i + 2 < nst; i += 3) if (h[i + 2]) { a += h[i]; bq += h[i + 1]; t += h[i + 2]; ++n; }
      if (n) fprintf(stderr, "[conv1 stats] CTAs %d: MMA thread waits A %.1f%%, B %.1f%% of %.0f cycles per tile\n", n, 100 * a / t, 100 * bq / t, t / n);
      cudaFree(d_stats);
    }
  }
  if (prev >= 0 && prev != plan->device) cudaSetDevice(prev);
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "warp_conv1_fwd launch");
}
}  // namespace
