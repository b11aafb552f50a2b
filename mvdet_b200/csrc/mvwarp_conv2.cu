// mvwarp_conv2.cu -- the rest of map_classifier behind the fused warp -> conv1:
//   map_classifier[2] = Conv2d(512, 512, 3, padding=2, dilation=2) + ReLU     conv3x3_tc_kernel  (tcgen05 / TMEM / TMA)
//   map_classifier[4] = Conv2d(512,   1, 3, padding=4, dilation=4, bias=False) conv3x3_head_kernel (memory-bound, no tensor cores)
// (/root/reference/multiview_detector/models/persp_trans_detector.py:53-54, applied at :81).
//
// conv3x3_tc_kernel is an implicit GEMM like warp_conv1_pipe_kernel (M = an 8 x 16 tile of cells, N = 512 = all of
// TMEM, K = 9 taps x Cin in blocks of 64), but here the A operand already exists in memory as a channels-last bf16
// activation tensor, so nothing has to be produced: the K block of tap (dy, dx) and channel block cb is the BOX
// [8 rows][16 cells][64 channels] of that tensor displaced by ((dy-1)*d, (dx-1)*d), which one TMA tensor copy
// (cp.async.bulk.tensor.4d, SWIZZLE_128B) drops into shared memory in exactly the K-major layout the tensor core
// reads; cells outside the image are the convolution's zero padding and come from the TMA unit's out-of-bounds fill.
// Three single-thread roles (A boxes, packed weight blocks, MMA issue) and eight epilogue warps.
// Two kernels: conv3x3_tc2_kernel (default) runs CTA PAIRS with cta_group::2 MMAs of M = 256, each CTA streaming half of
// every weight block; conv3x3_tc_kernel is the single-CTA version it grew out of (kept behind MVWARP_CONV2_1CTA=1).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "mvwarp_internal.h"
#include "mvwarp_umma.h"

namespace mvw {
using namespace umma;

constexpr int C2_TH = 8, C2_TW = 16;              // cell tile = 128 GEMM rows
constexpr int C2_N = 512, C2_KB = 64;
constexpr uint32_t C2_A_BYTES = 128 * C2_KB * 2;  // 16 KB per A stage
constexpr uint32_t C2_B_BYTES = C2_N * C2_KB * 2; // 64 KB per B stage
constexpr int C2_PA = 4, C2_PB = 2;
constexpr int C2_EPI = 256, C2_THREADS = C2_EPI + 96;   // 8 epilogue warps + MMA warp + weight warp + box warp

struct Conv2Params {
  const __nv_bfloat16* wpack;   // [Cin/64 * 9][512][64] bf16, rows pre-swizzled (pack_conv1_weight with one "view", no coord block)
  const float* bias;            // [512] or nullptr
  void* out;                    // [B, H, W, 512] bf16 or fp32 (logical [B,512,H,W] channels-last)
  int64_t oB, oH, oW;           // in elements of the output dtype
  int H, W, Cin, dilation, relu, out_bf16;
  int tiles_x, tiles_y, ntiles;
};

// TMA tensor copy global -> shared, 4-D tile mode, completion on an mbarrier.  Coordinates innermost first and signed:
// parts of the box outside the tensor are filled with zeros.
__device__ __forceinline__ void tma_box_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const unsigned*>(&h);
}

#ifdef MVWARP_DEV  // single-CTA generation of the TMA convolution (development build only)
template <int CL>
__global__ void __launch_bounds__(C2_THREADS, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const Conv2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                // C2_PA x 16 KB
  uint8_t* sB = smem + C2_PA * C2_A_BYTES;           // C2_PB x 64 KB
  __shared__ __align__(8) uint64_t a_full[C2_PA], a_empty[C2_PA], b_full[C2_PB], b_empty[C2_PB], acc_full;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool real_tile = (int)blockIdx.x < p.ntiles;   // grid padded to a multiple of CL: a padding CTA redoes the last tile, stores nothing
  const int tile = real_tile ? blockIdx.x : p.ntiles - 1;
  const int tx0 = (tile % p.tiles_x) * C2_TW, ty0 = (tile / p.tiles_x) * C2_TH, b = blockIdx.z;
  const int ncb = p.Cin / C2_KB;
  const int NK = ncb * 9;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

  if (warp == 0) tmem_alloc(&tmem_holder, C2_N);
  if (tid == 0) {
    for (int i = 0; i < C2_PA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C2_PB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], CL); }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 10) {
    // ===== A boxes: one TMA tensor copy per K block =====
    if (lane == 0) {
      tma_prefetch_desc(&tmap_x);
      for (int kb = 0; kb < NK; ++kb) {
        const int s = kb % C2_PA, cb = kb / 9, tap = kb % 9;
        if (kb >= C2_PA) mbar_wait(&a_empty[s], ((kb / C2_PA) - 1) & 1);
        mbar_expect_tx(&a_full[s], C2_A_BYTES);
        tma_box_4d(sA + (size_t)s * C2_A_BYTES, &tmap_x, cb * C2_KB, tx0 + (tap % 3 - 1) * p.dilation, ty0 + (tap / 3 - 1) * p.dilation, b,
                   &a_full[s]);
      }
    }
  } else if (warp == 9) {
    // ===== weight blocks: one bulk copy per K block (a 1/CL slice multicast to the cluster) =====
    if (lane == 0) {
      for (int kb = 0; kb < NK; ++kb) {
        const int s = kb % C2_PB;
        if (kb >= C2_PB) mbar_wait(&b_empty[s], ((kb / C2_PB) - 1) & 1);
        mbar_expect_tx(&b_full[s], C2_B_BYTES);
        if (CL == 1) {
          bulk_g2s(sB + (size_t)s * C2_B_BYTES, p.wpack + (size_t)kb * (C2_N * C2_KB), C2_B_BYTES, &b_full[s]);
        } else {
          constexpr uint32_t SLICE = C2_B_BYTES / CL;
          bulk_g2s_multicast(sB + (size_t)s * C2_B_BYTES + crank * SLICE,
                             reinterpret_cast<const uint8_t*>(p.wpack + (size_t)kb * (C2_N * C2_KB)) + crank * SLICE, SLICE, &b_full[s], CMASK);
        }
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(128, 256);
      for (int kb = 0; kb < NK; ++kb) {
        const int sa = kb % C2_PA, sb = kb % C2_PB;
        mbar_wait(&a_full[sa], (kb / C2_PA) & 1);
        mbar_wait(&b_full[sb], (kb / C2_PB) & 1);
        tc_fence_after();
        const uint32_t bbase = smem_u32(sB + (size_t)sb * C2_B_BYTES);
        const uint64_t da = smem_desc_sw128(smem_u32(sA + (size_t)sa * C2_A_BYTES)), db0 = smem_desc_sw128(bbase),
                       db1 = smem_desc_sw128(bbase + 256 * 128);
#pragma unroll
        for (int k = 0; k < C2_KB / 16; ++k) {
          mma_bf16(tmem, da + 2 * k, db0 + 2 * k, idesc, kb > 0 || k > 0);
          mma_bf16(tmem + 256, da + 2 * k, db1 + 2 * k, idesc, kb > 0 || k > 0);
        }
        mma_commit(&a_empty[sa]);
        if (CL == 1) mma_commit(&b_empty[sb]);
        else mma_commit_multicast(&b_empty[sb], CMASK);
      }
      mma_commit(&acc_full);
    }
  } else {
    // ===== epilogue: thread = tile cell = TMEM lane (quarter warp % 4); warps 0-3 take channels 0-255, warps 4-7 256-511 =====
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int er = (warp & 3) * 32 + lane, ey = er / C2_TW, ex = er % C2_TW;
    const int gy = ty0 + ey, gx = tx0 + ex;
    const bool inside = real_tile && gy < p.H && gx < p.W;
    const int64_t ooff = (int64_t)b * p.oB + (int64_t)gy * p.oH + (int64_t)gx * p.oW;
    const int cbeg = (warp >> 2) * 256;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + 256; c0 += 32) {
      float vv[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, vv);
      if (inside) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (p.bias) vv[i] += __ldg(p.bias + c0 + i);
          if (p.relu) vv[i] = fmaxf(vv[i], 0.f);
        }
        if (p.out_bf16) {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = make_uint4(pack_bf16x2(vv[8 * i], vv[8 * i + 1]), pack_bf16x2(vv[8 * i + 2], vv[8 * i + 3]),
                              pack_bf16x2(vv[8 * i + 4], vv[8 * i + 5]), pack_bf16x2(vv[8 * i + 6], vv[8 * i + 7]));
        } else {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + ooff + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync(); else __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C2_N);
}
#endif  // MVWARP_DEV

// -----------------------------------------------------------------------------------------------------
// CTA-pair version (cta_group::2): the two CTAs of a cluster take two neighbouring cell tiles and run ONE MMA of
// M = 256 per step.  Each CTA fetches its own A boxes and only HALF of every weight block (the N rows [128r, 128r+128) and
// [256+128r, 256+128r+128) of the 512, r = rank in the pair) -- the weight stream from L2, which bounds the single-CTA kernel
// (64 KB per 0.55 us of tensor work per SM), is halved.  All TMA copies of the pair (tensor maps for A and for the packed
// weights, .cta_group::2) complete on the LEADER's full barriers, the leader's MMA thread issues for both CTAs, and its
// commits arrive on the empty barriers of both (multicast).  Each CTA drains its own 128 x 512 accumulator from its own TMEM.
// Tried and dropped: fetching the tile's halo patch ONCE per channel block (one x-major TMA box through a tensor map with the
// dimensions listed as (C, y, x, batch)) and reading the nine taps in place like warp_conv1_direct_kernel does -- correct, but
// 0.200 vs 0.193 ms at Wildtrack B=1 and 0.716 vs 0.685 ms at B=4: with only 72 K blocks per tile this kernel is bound by the
// per-tile prologue and the un-overlapped epilogue (all 512 TMEM columns hold the accumulator), not by its A side.  Staging the
// bf16 tile in shared memory for coalesced 512-byte store runs (instead of 16 bytes per lane, 1 KB apart) changed nothing either
// (0.183 / 0.694 ms): what a short tile loses is the pipeline fill and the two cluster barriers, not the stores.
// -----------------------------------------------------------------------------------------------------
constexpr int T2_PA = 4, T2_PB = 4;
constexpr uint32_t T2_B_BYTES = C2_B_BYTES / 2;   // 32 KB of weights per CTA per K block

__device__ __forceinline__ void tma_box_4d_2cta(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_cluster)
      : "memory");
}
__device__ __forceinline__ void tma_box_2d_2cta(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar_cluster)
               : "memory");
}

__global__ void __launch_bounds__(C2_THREADS, 1) conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                                                                   const Conv2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                                // T2_PA x 16 KB
  uint8_t* sB = smem + T2_PA * C2_A_BYTES;           // T2_PB x 32 KB
  __shared__ __align__(8) uint64_t a_full[T2_PA], a_empty[T2_PA], b_full[T2_PB], b_empty[T2_PB], acc_full;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader
  const int nitems = p.ntiles * (int)gridDim.z;
  (void)nitems;
  const bool real_tile = (int)blockIdx.x < p.ntiles;   // grid.x padded to an even count: the padding CTA redoes the last tile, stores nothing
  const int tile = real_tile ? blockIdx.x : p.ntiles - 1;
  const int tx0 = (tile % p.tiles_x) * C2_TW, ty0 = (tile / p.tiles_x) * C2_TH, b = blockIdx.z;
  const int NK = p.Cin / C2_KB * 9;

  if (warp == 0) tmem_alloc_2cta(&tmem_holder, C2_N);
  if (tid == 0) {
    for (int i = 0; i < T2_PA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < T2_PB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 10) {
    // ===== A boxes (own tile), completing on the leader's barrier =====
    if (lane == 0) {
      tma_prefetch_desc(&tmap_x);
      for (int kb = 0; kb < NK; ++kb) {
        const int s = kb % T2_PA, cb = kb / 9, tap = kb % 9;
        if (kb >= T2_PA) mbar_wait(&a_empty[s], ((kb / T2_PA) - 1) & 1);
        if (rank == 0) mbar_expect_tx(&a_full[s], 2 * C2_A_BYTES);   // both CTAs' boxes
        tma_box_4d_2cta(smem_u32(sA + (size_t)s * C2_A_BYTES), &tmap_x, cb * C2_KB, tx0 + (tap % 3 - 1) * p.dilation,
                        ty0 + (tap / 3 - 1) * p.dilation, b, map_to_cta(smem_u32(&a_full[s]), 0));
      }
    }
  } else if (warp == 9) {
    // ===== this CTA's half of every weight block: two boxes of 128 rows =====
    if (lane == 0) {
      tma_prefetch_desc(&tmap_w);
      for (int kb = 0; kb < NK; ++kb) {
        const int s = kb % T2_PB;
        if (kb >= T2_PB) mbar_wait(&b_empty[s], ((kb / T2_PB) - 1) & 1);
        if (rank == 0) mbar_expect_tx(&b_full[s], 2 * T2_B_BYTES);
        const uint32_t dst = smem_u32(sB + (size_t)s * T2_B_BYTES), bar = map_to_cta(smem_u32(&b_full[s]), 0);
        tma_box_2d_2cta(dst, &tmap_w, 0, kb * C2_N + 128 * (int)rank, bar);
        tma_box_2d_2cta(dst + T2_B_BYTES / 2, &tmap_w, 0, kb * C2_N + 256 + 128 * (int)rank, bar);
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer: the leader, for both CTAs =====
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = idesc_bf16_f32(256, 256);
      for (int kb = 0; kb < NK; ++kb) {
        const int sa = kb % T2_PA, sb = kb % T2_PB;
        mbar_wait(&a_full[sa], (kb / T2_PA) & 1);
        mbar_wait(&b_full[sb], (kb / T2_PB) & 1);
        tc_fence_after();
        const uint32_t bbase = smem_u32(sB + (size_t)sb * T2_B_BYTES);
        const uint64_t da = smem_desc_sw128(smem_u32(sA + (size_t)sa * C2_A_BYTES)), db0 = smem_desc_sw128(bbase),
                       db1 = smem_desc_sw128(bbase + T2_B_BYTES / 2);
#pragma unroll
        for (int k = 0; k < C2_KB / 16; ++k) {
          mma_bf16_2cta(tmem, da + 2 * k, db0 + 2 * k, idesc, kb > 0 || k > 0);
          mma_bf16_2cta(tmem + 256, da + 2 * k, db1 + 2 * k, idesc, kb > 0 || k > 0);
        }
        mma_commit_2cta(&a_empty[sa], 3);
        mma_commit_2cta(&b_empty[sb], 3);
      }
      mma_commit_2cta(&acc_full, 3);
    }
  } else {
    // ===== epilogue (each CTA its own tile) =====
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int er = (warp & 3) * 32 + lane, ey = er / C2_TW, ex = er % C2_TW;
    const int gy = ty0 + ey, gx = tx0 + ex;
    const bool inside = real_tile && gy < p.H && gx < p.W;
    const int64_t ooff = (int64_t)b * p.oB + (int64_t)gy * p.oH + (int64_t)gx * p.oW;
    const int cbeg = (warp >> 2) * 256;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + 256; c0 += 32) {
      float vv[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, vv);
      if (inside) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (p.bias) vv[i] += __ldg(p.bias + c0 + i);
          if (p.relu) vv[i] = fmaxf(vv[i], 0.f);
        }
        if (p.out_bf16) {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = make_uint4(pack_bf16x2(vv[8 * i], vv[8 * i + 1]), pack_bf16x2(vv[8 * i + 2], vv[8 * i + 3]),
                              pack_bf16x2(vv[8 * i + 4], vv[8 * i + 5]), pack_bf16x2(vv[8 * i + 6], vv[8 * i + 7]));
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
  if (warp == 0) tmem_dealloc_2cta(tmem, C2_N);
}

// -----------------------------------------------------------------------------------------------------
// map_classifier[4]: 512 -> 1 channel, 3x3, dilation d, no bias.  0.4 GFLOP per Wildtrack frame against 44 MB of
// activations: memory-bound, so no tensor cores -- one warp per output cell, lanes across channels (128-bit loads of
// the nine tap rows, the 9 x Cin weights staged in shared memory as fp32), warp-shuffle reduction, fp32 out.
// -----------------------------------------------------------------------------------------------------
constexpr int HD_WARPS = 8, HD_CELLS = 16;   // cells per warp: a column run (x fixed, rows `dilation` apart)

// A warp walks DOWN a column in steps of the dilation: cells (x, ph + d*j), j = j0 .. j0+15.  The three tap rows of cell j
// (y-d, y, y+d) are then the rows of cells j-1, j, j+1, so a four-slot register window slides down the column: each step
// loads ONE new tap row (3 taps x NTRIP 128-bit loads per lane, issued one step ahead) instead of three, and the other
// two come from registers.  Lane l owns channels [8l, 8l+8) + 256 t, t < NTRIP = ceil(Cin / 256); its 9 x 8 x NTRIP
// weights stay in registers as packed bf16 pairs.  Neighbouring warps take neighbouring columns, whose x +- d taps
// overlap in L1.  (First version: warp per cell, nine rows per cell, weights read from shared memory: 5x slower.)
template <int NTRIP>
__global__ void __launch_bounds__(HD_WARPS * 32, 1) conv3x3_head_kernel(const __nv_bfloat16* __restrict__ x, int64_t xB, int64_t xH, int64_t xW,
                                                                        const __nv_bfloat16* __restrict__ w /*[9][Cin]*/, float* __restrict__ out,
                                                                        int64_t oB, int64_t oH, int64_t oW, int B, int H, int W, int Cin,
                                                                        int dilation, int nblk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warp -> (batch, phase, run of the phase's rows, column); column fastest
  int64_t wid = (int64_t)blockIdx.x * HD_WARPS + warp;
  const int xq = (int)(wid % W); wid /= W;
  const int jb = (int)(wid % nblk); wid /= nblk;
  const int ph = (int)(wid % dilation);
  const int64_t b = wid / dilation;
  if (b >= B || ph >= H) return;
  const int rows = (H - ph + dilation - 1) / dilation;   // cells of this phase in a column
  const int j0 = jb * HD_CELLS, nj = min(HD_CELLS, rows - j0);
  if (nj <= 0) return;

  uint4 wt[NTRIP][9];   // packed bf16 pairs
  bool act[NTRIP];
#pragma unroll
  for (int t = 0; t < NTRIP; ++t) {
    const int c = lane * 8 + 256 * t;
    act[t] = c < Cin;   // Cin % 8 == 0: a lane's eight channels are all inside or all outside
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      wt[t][tap] = act[t] ? __ldg(reinterpret_cast<const uint4*>(w + (size_t)tap * Cin + c)) : make_uint4(0, 0, 0, 0);
  }
  const __nv_bfloat16* xb = x + b * xB + lane * 8;
  uint4 win[4][3][NTRIP];   // [slot][dx][trip]
  auto load_row = [&](uint4 (&dst)[3][NTRIP], int yy) {
#pragma unroll
    for (int dxi = 0; dxi < 3; ++dxi) {
      const int xx = xq + (dxi - 1) * dilation;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;   // warp-uniform; outside = the conv's zero padding
#pragma unroll
      for (int t = 0; t < NTRIP; ++t)
        dst[dxi][t] = (in && act[t]) ? __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)yy * xH + (int64_t)xx * xW + 256 * t)) : make_uint4(0, 0, 0, 0);
    }
  };
  const int y0 = ph + dilation * j0;
  load_row(win[0], y0 - dilation);
  load_row(win[1], y0);
  load_row(win[2], y0 + dilation);
  for (int jj = 0; jj < nj; jj += 4) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {   // jj % 4 == 0: the slot numbers below are compile-time constants
      const int j = jj + r;
      if (j < nj) {
        const int y = y0 + dilation * j;
        if (j + 1 < nj) load_row(win[(r + 3) & 3], y + 2 * dilation);   // bottom row of the NEXT cell, one step ahead
        float a4[4] = {0.f, 0.f, 0.f, 0.f};   // four independent FMA chains (one chain of 144 would cost 4 cycles per FMA)
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
          for (int dxi = 0; dxi < 3; ++dxi)
#pragma unroll
            for (int t = 0; t < NTRIP; ++t) {
              const uint4 v = win[(r + dyi) & 3][dxi][t], q = wt[t][dyi * 3 + dxi];
              const unsigned vu[4] = {v.x, v.y, v.z, v.w}, qu[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                a4[i] = __fmaf_rn(__uint_as_float(vu[i] << 16), __uint_as_float(qu[i] << 16), a4[i]);
                a4[i] = __fmaf_rn(__uint_as_float(vu[i] & 0xffff0000u), __uint_as_float(qu[i] & 0xffff0000u), a4[i]);
              }
            }
        float acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[b * oB + (int64_t)y * oH + (int64_t)xq * oW] = acc;
      }
    }
  }
}

// -----------------------------------------------------------------------------------------------------
// map_classifier[4] in two passes (the default when the caller provides scratch): every activation is read ONCE.
//   pass 1  P[cell][tap] = sum_c x[cell][c] * w[tap][c]      nine partial dot products per cell (head_partial_kernel)
//   pass 2  out[y][x]    = sum_tap P[(y + dy*d, x + dx*d)][tap]   a 9-term stencil over the 36-byte rows of P (head_stencil_kernel)
// The one-pass kernel above reads each cell's 1 KB row nine times (once per tap; three from registers) through L1 / L2 and lost to
// cuDNN at B >= 4.  Pass 1 is a skinny GEMM -- M = cells, N = 9, K = Cin -- and runs on the tensor cores with mma.sync
// (m16n8k16, bf16 x bf16 -> fp32; two n-tiles of eight taps, the second holds tap 8 only): with N = 9 the contraction is far below
// the HBM ridge whatever issues it, so this is a bandwidth kernel, and mma.sync takes its A fragments straight from registers --
// no shared-memory staging of the activations, no TMEM round trip.  The K order inside a k-step is a free choice as long as both
// operands agree, so lane (g, t) of a warp reads 16 contiguous bytes (channels c0 + 8t .. + 8 of rows g and g + 8: LDG.128) and uses
// register pairs (0, 1) as the first k-step and (2, 3) as the second; the weights follow the same rule from shared memory.
// -----------------------------------------------------------------------------------------------------
constexpr int HP_WARPS = 8, HP_MT = 2, HP_CELLS = HP_WARPS * HP_MT * 16;   // 256 cells per CTA
constexpr int HP_TAPS = 9;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(HP_WARPS * 32, 2) head_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t xB, int64_t xH, int64_t xW,
                                                                        const __nv_bfloat16* __restrict__ w /*[9][Cin]*/, float* __restrict__ P,
                                                                        int64_t ncells, int H, int W, int Cin) {
  extern __shared__ __align__(16) uint8_t hp_smem[];
  const int pitch = Cin * 2 + 16;                        // bytes; +16: the eight tap rows of a quarter-warp's LDS.128 fall into different banks
  for (int i = threadIdx.x; i < HP_TAPS * (Cin / 8); i += blockDim.x) {
    const int tap = i / (Cin / 8), c8 = i % (Cin / 8);
    *reinterpret_cast<uint4*>(hp_smem + (size_t)tap * pitch + c8 * 16) = __ldg(reinterpret_cast<const uint4*>(w + (size_t)tap * Cin + c8 * 8));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t base = (int64_t)blockIdx.x * HP_CELLS + (int64_t)warp * (HP_MT * 16);
  if (base >= ncells) return;
  const __nv_bfloat16* pa[HP_MT][2];
#pragma unroll
  for (int m = 0; m < HP_MT; ++m)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int64_t cell = base + m * 16 + h * 8 + g;
      if (cell >= ncells) cell = ncells - 1;             // a clamped row is computed and dropped at the store
      const int64_t hw = (int64_t)H * W, b = cell / hw, r = cell - b * hw;
      const int y = (int)(r / W), xx = (int)(r - (int64_t)y * W);
      pa[m][h] = x + b * xB + (int64_t)y * xH + (int64_t)xx * xW + 8 * t;
    }
  const uint8_t* wb0 = hp_smem + (size_t)g * pitch + 16 * t;            // taps 0..7 (n-tile 0, column g)
  const uint8_t* wb1 = hp_smem + (size_t)8 * pitch + 16 * t;            // tap 8 (n-tile 1, column 0; the other columns are zero)
  float acc[HP_MT][2][4];
#pragma unroll
  for (int m = 0; m < HP_MT; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
  constexpr int UB = 4;                                  // channel blocks (32 channels each) per trip: 2 * HP_MT * UB loads in flight per lane
  for (int c0 = 0; c0 < Cin; c0 += 32 * UB) {
    uint4 a[UB][HP_MT][2];
#pragma unroll
    for (int u = 0; u < UB; ++u)
#pragma unroll
      for (int m = 0; m < HP_MT; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          a[u][m][h] = (c0 + 32 * u < Cin) ? __ldg(reinterpret_cast<const uint4*>(pa[m][h] + c0 + 32 * u)) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (c0 + 32 * u < Cin) {
        const uint4 b0 = *reinterpret_cast<const uint4*>(wb0 + (c0 + 32 * u) * 2);
        const uint4 b1 = g == 0 ? *reinterpret_cast<const uint4*>(wb1 + (c0 + 32 * u) * 2) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int m = 0; m < HP_MT; ++m) {
          const uint4 lo = a[u][m][0], hi = a[u][m][1];
          mma_bf16_16816(acc[m][0], lo.x, hi.x, lo.y, hi.y, b0.x, b0.y);
          mma_bf16_16816(acc[m][0], lo.z, hi.z, lo.w, hi.w, b0.z, b0.w);
          mma_bf16_16816(acc[m][1], lo.x, hi.x, lo.y, hi.y, b1.x, b1.y);
          mma_bf16_16816(acc[m][1], lo.z, hi.z, lo.w, hi.w, b1.z, b1.w);
        }
      }
    }
  }
  // accumulator (row g / g + 8, columns 2t, 2t + 1) -> P[cell][tap]: n-tile 0 = taps 2t, 2t + 1; n-tile 1 = tap 8 in column 0 (t == 0)
#pragma unroll
  for (int m = 0; m < HP_MT; ++m)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t cell = base + m * 16 + h * 8 + g;
      if (cell < ncells) {
        float* o = P + cell * HP_TAPS;
        o[2 * t] = acc[m][0][2 * h];
        o[2 * t + 1] = acc[m][0][2 * h + 1];
        if (t == 0) o[8] = acc[m][1][2 * h];
      }
    }
}

__global__ void __launch_bounds__(256) head_stencil_kernel(const float* __restrict__ P, float* __restrict__ out, int64_t oB, int64_t oH, int64_t oW,
                                                           int64_t ncells, int H, int W, int dilation) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncells) return;
  const int64_t hw = (int64_t)H * W, b = cell / hw, r = cell - b * hw;
  const int y = (int)(r / W), x = (int)(r - (int64_t)y * W);
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < HP_TAPS; ++tap) {              // tap order = the one-pass kernel's: (dy, dx) row-major
    const int yy = y + (tap / 3 - 1) * dilation, xx = x + (tap % 3 - 1) * dilation;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc += __ldg(P + ((b * H + yy) * (int64_t)W + xx) * HP_TAPS + tap);
  }
  out[b * oB + (int64_t)y * oH + (int64_t)x * oW] = acc;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_weight_tensor_map(void* tmap_out, const void* wpack, uint64_t nblocks) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("the driver does not export cuTensorMapEncodeTiled"); return MVWARP_EUNSUPPORTED; }
  if (reinterpret_cast<uintptr_t>(wpack) & 15u) { set_error("packed weights must be 16-byte aligned"); return MVWARP_EUNSUPPORTED; }
  const cuuint64_t wdim[2] = {(cuuint64_t)C2_KB, (cuuint64_t)nblocks * C2_N};
  const cuuint64_t wstr[1] = {(cuuint64_t)C2_KB * 2};
  const cuuint32_t wbox[2] = {C2_KB, 128};
  const cuuint32_t west[2] = {1, 1};
  const CUresult wr = enc(reinterpret_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wpack), wdim, wstr, wbox,
                          west, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (wr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed (code " + std::to_string((int)wr) + ")"); return MVWARP_ECUDA; }
  return MVWARP_OK;
}

}  // namespace mvw

using namespace mvw;

extern "C" int mvwarp_conv3x3_fwd(const void* x, const int64_t x_strides[4], const void* wpack, const float* bias, void* out,
                                  const int64_t out_strides[4], int B, int H, int W, int Cin, int dilation, int relu, int out_dtype,
                                  int device, void* stream) {
  if (!x || !x_strides || !wpack || !out || !out_strides) { set_error("conv3x3_fwd: null argument"); return MVWARP_EINVAL; }
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || dilation <= 0 || B > 65535) { set_error("conv3x3_fwd: bad B / H / W / Cin / dilation"); return MVWARP_EINVAL; }
  if (out_dtype != MVWARP_F32 && out_dtype != MVWARP_BF16) { set_error("conv3x3_fwd: unknown output dtype"); return MVWARP_EINVAL; }
  if (Cin % C2_KB || x_strides[1] != 1 || (reinterpret_cast<uintptr_t>(x) & 15u) || x_strides[0] % 8 || x_strides[2] % 8 || x_strides[3] % 8) {
    set_error("conv3x3_fwd: x must be bf16 channels-last [B,Cin,H,W] with Cin % 64 == 0 and 16-byte aligned cell rows");
    return MVWARP_EUNSUPPORTED;
  }
  const int oal = out_dtype == MVWARP_BF16 ? 8 : 4;
  if (out_strides[1] != 1 || out_strides[0] % oal || out_strides[2] % oal || out_strides[3] % oal || (reinterpret_cast<uintptr_t>(out) & 15u)) {
    set_error("conv3x3_fwd: out must be channels-last [B,512,H,W] with 16-byte aligned cell rows");
    return MVWARP_EUNSUPPORTED;
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) { set_error("conv3x3_fwd: the driver does not export cuTensorMapEncodeTiled"); return MVWARP_EUNSUPPORTED; }
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device && cudaSetDevice(device) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  auto restore = [&]() { if (prev >= 0 && prev != device) cudaSetDevice(prev); };

  // x as a 4-D tensor, innermost first: (channel, x, y, batch); box = 64 channels x 16 cells x 8 rows x 1 frame
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)x_strides[3] * 2, (cuuint64_t)x_strides[2] * 2, (cuuint64_t)x_strides[0] * 2};
  const cuuint32_t box[4] = {C2_KB, C2_TW, C2_TH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    restore();
    set_error("conv3x3_fwd: cuTensorMapEncodeTiled failed (code " + std::to_string((int)cr) + ")");
    return MVWARP_ECUDA;
  }
  CUtensorMap tmap_w;   // the packed weight as a 2-D tensor of 128-byte rows, boxes of 128 rows
  {
    const int wrc = make_weight_tensor_map(&tmap_w, wpack, (uint64_t)(Cin / C2_KB) * 9);
    if (wrc) { restore(); return wrc; }
  }
  Conv2Params p{};
  p.wpack = (const __nv_bfloat16*)wpack; p.bias = bias; p.out = out;
  p.oB = out_strides[0]; p.oH = out_strides[2]; p.oW = out_strides[3];
  p.H = H; p.W = W; p.Cin = Cin; p.dilation = dilation; p.relu = relu; p.out_bf16 = out_dtype == MVWARP_BF16;
  p.tiles_x = (W + C2_TW - 1) / C2_TW; p.tiles_y = (H + C2_TH - 1) / C2_TH; p.ntiles = p.tiles_x * p.tiles_y;
  const int smem = C2_PA * (int)C2_A_BYTES + C2_PB * (int)C2_B_BYTES + 1024;
  // Cluster of 2 with multicast weight blocks: measured equal to no cluster (0.190 vs 0.185 ms at the Wildtrack shape);
  // a K-split of the last partial wave (as in the fused conv1) does not pay here either (0.200 ms): the CTAs of a thin
  // last wave already run faster because the weight stream from L2 is shared by fewer SMs.
#ifdef MVWARP_DEV
  // development knobs, read once: MVWARP_CONV2_CLUSTER=2 (cluster of two, single-CTA kernel), MVWARP_CONV2_1CTA=1 (the single-CTA
  // kernel: 0.203 vs 0.182 ms at Wildtrack B=1, 0.701 vs 0.672 at B=4)
  static const int cl = [] { const char* v = getenv("MVWARP_CONV2_CLUSTER"); return (v && v[0] == '2') ? 2 : 1; }();
  static const bool one_cta = [] { const char* v = getenv("MVWARP_CONV2_1CTA"); return v && v[0] == '1'; }();
  auto launch = [&](auto kernel, int cluster) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((p.ntiles + cluster - 1) / cluster * cluster, 1, B);
    cfg.blockDim = dim3(C2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, tmap, p);
  };
#else
  constexpr bool one_cta = false;
  (void)smem;
#endif
  cudaError_t e;
  if (!one_cta) {                                  // default: CTA pairs (cta_group::2)
    const int smem2 = T2_PA * (int)C2_A_BYTES + T2_PB * (int)T2_B_BYTES + 1024;
    e = cudaFuncSetAttribute(conv3x3_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    if (e == cudaSuccess) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((p.ntiles + 1) / 2 * 2, 1, B);
      cfg.blockDim = dim3(C2_THREADS);
      cfg.dynamicSmemBytes = smem2;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, conv3x3_tc2_kernel, tmap, tmap_w, p);
    }
  }
#ifdef MVWARP_DEV
  else {
    e = cl == 1 ? launch(conv3x3_tc_kernel<1>, 1) : launch(conv3x3_tc_kernel<2>, 2);
  }
#endif
  restore();
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "conv3x3_fwd launch");
}

extern "C" int64_t mvwarp_conv3x3_head_scratch_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)B * H * W * HP_TAPS * (int64_t)sizeof(float);
}

extern "C" int mvwarp_conv3x3_head_fwd(const void* x, const int64_t x_strides[4], const void* w, float* out, const int64_t out_strides[4],
                                       int B, int H, int W, int Cin, int dilation, int device, void* stream) {
  return mvwarp_conv3x3_head_fwd_ex(x, x_strides, w, out, out_strides, B, H, W, Cin, dilation, nullptr, 0, device, stream);
}

extern "C" int mvwarp_conv3x3_head_fwd_ex(const void* x, const int64_t x_strides[4], const void* w, float* out, const int64_t out_strides[4],
                                          int B, int H, int W, int Cin, int dilation, void* scratch, int64_t scratch_bytes, int device,
                                          void* stream) {
  if (!x || !x_strides || !w || !out || !out_strides) { set_error("conv3x3_head_fwd: null argument"); return MVWARP_EINVAL; }
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || dilation <= 0) { set_error("conv3x3_head_fwd: bad B / H / W / Cin / dilation"); return MVWARP_EINVAL; }
  if (Cin % 8 || Cin > 512 || x_strides[1] != 1 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15u) || x_strides[0] % 8 ||
      x_strides[2] % 8 || x_strides[3] % 8) {
    set_error("conv3x3_head_fwd: x must be bf16 channels-last [B,Cin,H,W] with Cin % 8 == 0, Cin <= 512 and 16-byte aligned cell rows");
    return MVWARP_EUNSUPPORTED;
  }
  const int nblk = ((H + dilation - 1) / dilation + HD_CELLS - 1) / HD_CELLS;   // column runs per phase
  const int64_t nwarps = (int64_t)B * dilation * nblk * W;
  if ((nwarps + HD_WARPS - 1) / HD_WARPS > 0x7fffffff) { set_error("conv3x3_head_fwd: too many cells"); return MVWARP_EUNSUPPORTED; }
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device && cudaSetDevice(device) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  const int64_t ncells = (int64_t)B * H * W;
  if (scratch && Cin % 32 == 0 && !(reinterpret_cast<uintptr_t>(scratch) & 15u) && scratch_bytes >= ncells * HP_TAPS * (int64_t)sizeof(float) &&
      (ncells + HP_CELLS - 1) / HP_CELLS <= 0x7fffffff) {
    // two passes: nine partial dot products per cell on the tensor cores (every activation read once), then the 9-term stencil
    float* P = reinterpret_cast<float*>(scratch);
    const int smem = HP_TAPS * (Cin * 2 + 16);
    head_partial_kernel<<<(unsigned)((ncells + HP_CELLS - 1) / HP_CELLS), HP_WARPS * 32, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, x_strides[0], x_strides[2], x_strides[3], (const __nv_bfloat16*)w, P, ncells, H, W, Cin);
    cudaError_t e2 = cudaGetLastError();
    if (e2 == cudaSuccess) {
      head_stencil_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, out, out_strides[0], out_strides[2], out_strides[3],
                                                                                           ncells, H, W, dilation);
      e2 = cudaGetLastError();
    }
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return e2 == cudaSuccess ? MVWARP_OK : cuda_fail(e2, "conv3x3_head_fwd (two-pass) launch");
  }
  const int grid = (int)((nwarps + HD_WARPS - 1) / HD_WARPS);
  if (Cin <= 256)
    conv3x3_head_kernel<1><<<grid, HD_WARPS * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, x_strides[0], x_strides[2], x_strides[3],
                                                                             (const __nv_bfloat16*)w, out, out_strides[0], out_strides[2],
                                                                             out_strides[3], B, H, W, Cin, dilation, nblk);
  else
    conv3x3_head_kernel<2><<<grid, HD_WARPS * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, x_strides[0], x_strides[2], x_strides[3],
                                                                             (const __nv_bfloat16*)w, out, out_strides[0], out_strides[2],
                                                                             out_strides[3], B, H, W, Cin, dilation, nblk);
  const cudaError_t e = cudaGetLastError();
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "conv3x3_head_fwd launch");
}
