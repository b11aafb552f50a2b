// mvwarp_host_io.cu -- sparse staging of the feature side between host and device.
//
// When the features live in (pinned) host memory, the end-to-end cost of the warp path is the PCIe transfer, not the
// kernels: 2.48 GB each way per Wildtrack frame against 0.65 ms of GPU work.  But the warp only ever READS the feature
// texels that some grid cell samples, and the feature gradient is non-zero only on those same texels -- 26 % of the
// texels on the Wildtrack rig, a set that depends on the calibration alone (it is the set of non-empty rows of the
// plan's inverse CSR map).  mvwarp_copy_touched_texels moves exactly those rows:
//   host -> device : a gather kernel reads the touched C-channel rows straight out of mapped pinned host memory
//                    (unified addressing: the host pointer is valid on the device) into the dense device tensor;
//   device -> host : the same kernel with the roles swapped scatters the non-zero gradient rows into a host tensor
//                    that was zeroed once and stays zero everywhere else.
// One warp per texel row, 16-byte accesses, rows in ascending address order per view, so the bus sees long runs.
// Replaces: the dense .to('cuda:0') hops of the reference forward (persp_trans_detector.py:63-68) for callers whose
// features are host-resident; nothing on the device-resident path uses it.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "mvwarp_internal.h"
#include "mvwarp_umma.h"

namespace mvw {

struct CopyParams {
  const uint32_t* touched;   // [U] linear texel ids, ascending per view
  uint32_t t0, nrows;        // range of the touched list covered by this launch (views [view0, view0 + nviews))
  uint32_t view_first[MVWARP_MAX_LIST_VIEWS + 1];   // touched_base of the launch's views, relative to t0
  int nviews, width, B;
  int row_bytes;             // C * element size
  const char* src;
  char* dst;
  int64_t sB, sV, sH, sW;    // byte strides
  int64_t dB, dV, dH, dW;
};

constexpr int COPY_WARPS = 8;

template <int VEC>
__global__ void __launch_bounds__(COPY_WARPS * 32) copy_touched_kernel(const CopyParams p) {
  const int lane = threadIdx.x & 31;
  const uint32_t nwarps = gridDim.x * COPY_WARPS;
  const int b = blockIdx.y;
  for (uint32_t r = blockIdx.x * COPY_WARPS + (threadIdx.x >> 5); r < p.nrows; r += nwarps) {
    int vl = 0;
    while (r >= p.view_first[vl + 1]) ++vl;   // few views: linear search in the parameter space
    const uint32_t id = __ldg(p.touched + p.t0 + r);
    const int y = id / p.width, x = id - y * p.width;
    const char* s = p.src + (int64_t)b * p.sB + (int64_t)vl * p.sV + (int64_t)y * p.sH + (int64_t)x * p.sW;
    char* d = p.dst + (int64_t)b * p.dB + (int64_t)vl * p.dV + (int64_t)y * p.dH + (int64_t)x * p.dW;
    if (VEC == 16) {
      // four independent 16-byte loads per lane in flight (a 2 KB fp32 texel row in one trip): reads from host memory are
      // latency-bound (~2 us per round trip over PCIe), so the bytes in flight, not the instruction count, set the bandwidth
      for (int o0 = lane * 16; o0 < p.row_bytes; o0 += 4 * 32 * 16) {
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int o = o0 + i * 32 * 16;
          if (o < p.row_bytes) v[i] = *reinterpret_cast<const uint4*>(s + o);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int o = o0 + i * 32 * 16;
          if (o < p.row_bytes) *reinterpret_cast<uint4*>(d + o) = v[i];
        }
      }
    } else {
      for (int o = lane * VEC; o < p.row_bytes; o += 32 * VEC) {
        if (VEC == 4) *reinterpret_cast<uint32_t*>(d + o) = *reinterpret_cast<const uint32_t*>(s + o);
        else *reinterpret_cast<uint16_t*>(d + o) = *reinterpret_cast<const uint16_t*>(s + o);
      }
    }
  }
}

#ifdef MVWARP_DEV
// The same copy on the TMA engine: every THREAD is an independent copier with a ring of NSLOT shared-memory slots -- one bulk
// async copy global/host -> shared per texel row (completion on the slot's mbarrier), one bulk async copy shared -> global/host
// back out (bulk group).  No LSU instructions touch the data, the SM only issues descriptors, and the copies are as long as the
// rows (2 KB), which is what a PCIe link likes.  32 threads x NSLOT rows are in flight per CTA.
template <int NSLOT>
__global__ void __launch_bounds__(32) copy_touched_tma_kernel(const CopyParams p) {
  using namespace umma;
  extern __shared__ __align__(128) unsigned char tbuf[];
  __shared__ __align__(8) uint64_t bars[32 * NSLOT];
  const int t = threadIdx.x, b = blockIdx.y;
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) mbar_init(&bars[t * NSLOT + s], 1);
  fence_barrier_init();
  __syncwarp();
  const uint32_t nthreads = gridDim.x * 32, r0 = blockIdx.x * 32 + t;
  const uint32_t count = p.nrows > r0 ? (p.nrows - r0 - 1) / nthreads + 1 : 0;
  const uint32_t bytes = (uint32_t)p.row_bytes;
  unsigned char* mine = tbuf + (size_t)t * NSLOT * bytes;
  auto locate = [&](uint32_t i, const char*& s, char*& d) {
    const uint32_t r = r0 + i * nthreads;
    int vl = 0;
    while (r >= p.view_first[vl + 1]) ++vl;
    const uint32_t id = __ldg(p.touched + p.t0 + r);
    const int y = id / p.width, x = id - y * p.width;
    s = p.src + (int64_t)b * p.sB + (int64_t)vl * p.sV + (int64_t)y * p.sH + (int64_t)x * p.sW;
    d = p.dst + (int64_t)b * p.dB + (int64_t)vl * p.dV + (int64_t)y * p.dH + (int64_t)x * p.dW;
  };
  auto load = [&](uint32_t i) {
    const char* s; char* d;
    locate(i, s, d);
    const int slot = i % NSLOT;
    mbar_expect_tx(&bars[t * NSLOT + slot], bytes);
    bulk_g2s(mine + (size_t)slot * bytes, s, bytes, &bars[t * NSLOT + slot]);
  };
  for (uint32_t i = 0; i < count && i < NSLOT; ++i) load(i);
  for (uint32_t i = 0; i < count; ++i) {
    const int slot = i % NSLOT;
    mbar_wait(&bars[t * NSLOT + slot], (i / NSLOT) & 1);
    const char* s; char* d;
    locate(i, s, d);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(smem_u32(mine + (size_t)slot * bytes)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (i + NSLOT < count) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read the slot: it may be refilled
      load(i + NSLOT);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

#endif  // MVWARP_DEV

}  // namespace mvw

using namespace mvw;

extern "C" int mvwarp_copy_touched_texels(const mvwarp_plan_t* plan, int view0, int nviews, const void* src,
                                          const int64_t src_strides[5], void* dst, const int64_t dst_strides[5], int B, int C,
                                          int dtype, void* stream) {
  if (!plan || !src || !dst || !src_strides || !dst_strides) { set_error("copy_touched_texels: null argument"); return MVWARP_EINVAL; }
  if (B <= 0 || C <= 0 || nviews <= 0 || view0 < 0 || view0 > plan->V || nviews > plan->V - view0) {
    set_error("copy_touched_texels: bad B/C/view range");
    return MVWARP_EINVAL;
  }
  if (dtype != MVWARP_F32 && dtype != MVWARP_BF16) { set_error("copy_touched_texels: unknown dtype"); return MVWARP_EINVAL; }
  if (nviews > MVWARP_MAX_LIST_VIEWS) { set_error("copy_touched_texels: more than MVWARP_MAX_LIST_VIEWS views per call"); return MVWARP_EUNSUPPORTED; }
  if (src_strides[2] != 1 || dst_strides[2] != 1) {
    set_error("copy_touched_texels: both tensors must be channel-contiguous (a texel's channels are one run of memory)");
    return MVWARP_EUNSUPPORTED;
  }
  if (B > 65535) { set_error("copy_touched_texels: B exceeds the CUDA grid limit"); return MVWARP_EUNSUPPORTED; }
  if ((int)plan->touched_base.size() != plan->V + 1) { set_error("copy_touched_texels: the plan carries no touched-texel list"); return MVWARP_EINVAL; }
  const int es = dtype == MVWARP_F32 ? 4 : 2;
  CopyParams p{};
  p.touched = plan->d_touched;
  p.t0 = plan->touched_base[view0];
  p.nrows = plan->touched_base[view0 + nviews] - p.t0;
  for (int v = 0; v <= nviews; ++v) p.view_first[v] = plan->touched_base[view0 + v] - p.t0;
  p.nviews = nviews; p.B = B;
  p.width = plan->upsampled ? plan->wl : plan->Ws;
  p.row_bytes = C * es;
  p.src = (const char*)src; p.dst = (char*)dst;
  p.sB = src_strides[0] * es; p.sV = src_strides[1] * es; p.sH = src_strides[3] * es; p.sW = src_strides[4] * es;
  p.dB = dst_strides[0] * es; p.dV = dst_strides[1] * es; p.dH = dst_strides[3] * es; p.dW = dst_strides[4] * es;
  if (p.nrows == 0) return MVWARP_OK;
  DeviceGuard guard(plan->device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
  // vector width: every row start and the row length must be multiples of it
  auto aligned = [&](int n) {
    const int64_t all[] = {p.sB, p.sV, p.sH, p.sW, p.dB, p.dV, p.dH, p.dW, (int64_t)p.row_bytes,
                           (int64_t)(reinterpret_cast<uintptr_t>(src) & 15), (int64_t)(reinterpret_cast<uintptr_t>(dst) & 15)};
    for (int64_t a : all) if (a % n) return false;
    return true;
  };
  // Enough warps to keep the bus full (4 CTAs x 8 warps x 2 KB per SM = 9 MB in flight against a bandwidth-delay product of
  // ~0.1 MB), few enough that the warp kernels of neighbouring frames keep most of every SM while a copy is in progress.
  const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)p.nrows + COPY_WARPS - 1) / COPY_WARPS, (uint64_t)plan->num_sms * 4);
  dim3 grid(blocks, B), block(COPY_WARPS * 32);
  // The TMA-engine variant is a development knob: on two boxes of the pool it measured 30.1 vs 26.9 and 30.6 vs 34.3 frames/s end
  // to end against the LSU kernel below -- the PCIe path of the box, not the copy method, decides (profiles/README.md).
#ifdef MVWARP_DEV
  int use_tma = 0, nslot = 2;
  static const int knob_tma = [] { const char* e = getenv("MVWARP_COPY_TMA"); return e ? atoi(e) : -1; }();   // 0 = LSU kernel, n = CTAs per SM
  static const int knob_slot = [] { const char* e = getenv("MVWARP_COPY_SLOTS"); return e ? atoi(e) : 0; }();
  if (knob_tma >= 0) use_tma = knob_tma;
  if (knob_slot == 4) nslot = 4;
  if (use_tma && aligned(16) && p.row_bytes <= 2048) {
    const int smem = 32 * nslot * p.row_bytes;
    auto kernel = nslot == 4 ? copy_touched_tma_kernel<4> : copy_touched_tma_kernel<2>;
    cudaError_t ea = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ea != cudaSuccess) return cuda_fail(ea, "cudaFuncSetAttribute(copy_touched_tma_kernel)");
    dim3 tgrid((unsigned)std::min<uint64_t>(((uint64_t)p.nrows + 31) / 32, (uint64_t)plan->num_sms * use_tma), B);
    kernel<<<tgrid, 32, smem, (cudaStream_t)stream>>>(p);
    const cudaError_t e2 = cudaGetLastError();
    return e2 == cudaSuccess ? MVWARP_OK : cuda_fail(e2, "copy_touched_texels (TMA) launch");
  }
#endif
  if (aligned(16)) copy_touched_kernel<16><<<grid, block, 0, (cudaStream_t)stream>>>(p);
  else if (aligned(4)) copy_touched_kernel<4><<<grid, block, 0, (cudaStream_t)stream>>>(p);
  else copy_touched_kernel<2><<<grid, block, 0, (cudaStream_t)stream>>>(p);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MVWARP_OK : cuda_fail(e, "copy_touched_texels launch");
}

extern "C" int mvwarp_plan_touched(const mvwarp_plan_t* plan, uint32_t* touched_base_out, uint32_t* touched_out, uint64_t capacity) {
  if (!plan) { set_error("plan_touched: null plan"); return MVWARP_EINVAL; }
  if ((int)plan->touched_base.size() != plan->V + 1) { set_error("plan_touched: the plan carries no touched-texel list"); return MVWARP_EINVAL; }
  if (touched_base_out) memcpy(touched_base_out, plan->touched_base.data(), plan->touched_base.size() * sizeof(uint32_t));
  if (touched_out) {
    const uint64_t n = plan->touched_base[plan->V];
    if (capacity < n) { set_error("plan_touched: capacity smaller than the list"); return MVWARP_EINVAL; }
    DeviceGuard guard(plan->device);
    if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    const cudaError_t e = cudaMemcpy(touched_out, plan->d_touched, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "plan_touched: device -> host copy");
  }
  return MVWARP_OK;
}
