/* mvwarp.h -- C ABI of libmvwarp.so, the B200 (sm_100a) perspective-warp library.
 *
 * This is the drop-in boundary for MVDet's perspective-transformation hot path.  In the
 * reference that path is a Python call sequence, not an FFI:
 *
 *   proj_mat = self.proj_mats[cam].repeat([B,1,1]).float().to('cuda:0')       persp_trans_detector.py:68
 *   world_feature = kornia.geometry.transform.warp_perspective(                persp_trans_detector.py:69
 *                       img_feature, proj_mat, self.reducedgrid_shape)
 *   world_features = torch.cat(world_features + [coord_map.repeat(B)], dim=1)  persp_trans_detector.py:77
 *   (autograd: CatBackward + grid_sampler_2d_backward, driven from trainer.py:48)
 *
 * (paths relative to /root/reference/multiview_detector/models/ resp. /root/reference/multiview_detector/).
 * The entry points below are what a ctypes binding for that sequence binds; INTEGRATION.md
 * shows the reference-side stub.  Conventions:
 *   - every function returns 0 on success, a negative MVWARP_E* code on failure, and never
 *     throws or aborts across the boundary; mvwarp_last_error() returns a thread-local message;
 *   - the caller (PyTorch) owns every tensor; the library owns only the opaque plan;
 *   - all device work is enqueued on the cudaStream_t passed as `stream` (void*), on the
 *     plan's device (the library sets the device itself: backward is called from autograd's
 *     worker thread, whose current device is not ours to assume);
 *   - strides are in ELEMENTS of the tensor's dtype, logical index order given per argument;
 *   - functions are re-entrant; a plan is immutable after creation and may be shared by threads.
 */
#ifndef MVWARP_H_
#define MVWARP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVWARP_VERSION 1

/* error codes */
#define MVWARP_OK            0
#define MVWARP_EINVAL       -1   /* bad argument (null pointer, non-positive size, bad dtype ...) */
#define MVWARP_ECUDA        -2   /* a CUDA runtime call failed; see mvwarp_last_error() */
#define MVWARP_ENOMEM       -3   /* host allocation failed */
#define MVWARP_EUNSUPPORTED -4   /* valid request the library has no kernel for */

/* element types of feature / output tensors (coordinates, weights and accumulation are always fp32) */
#define MVWARP_F32  0
#define MVWARP_BF16 1

typedef struct mvwarp_plan mvwarp_plan_t;

typedef struct mvwarp_stats {
  int32_t  V, Hs, Ws, Hg, Wg;
  int32_t  inline_coords;     /* 1: kernels recompute coordinates from the matrices; 0: read the grid table */
  uint64_t unique_texels;     /* U: source texels with >= 1 in-bounds tap, summed over views */
  uint64_t nnz;               /* in-bounds taps = entries of the inverse (texel -> cells) CSR map */
  uint64_t valid_cells;       /* grid cells with >= 1 in-bounds tap, summed over views */
  uint32_t max_row;           /* longest CSR row (taps landing on one texel) */
  uint32_t bwd_tasks;         /* backward: warp-sized tasks (runs of light rows) */
  uint32_t heavy_rows;        /* backward: rows long enough to get a whole CTA */
  uint32_t fwd_tasks;         /* fused up-sample plans: warp-sized tasks of the forward table gather (else 0) */
  int32_t  lowres_h, lowres_w;/* fused up-sample plans: the low-res feature map the kernels read (else 0); for such plans
                                 unique_texels / nnz / max_row describe the composed map on the low-res image */
  uint32_t zero_segments;     /* backward: runs of <= 8 consecutive texels (one image row of one 8x8 tile) whose rows are all empty;
                                 dedicated CTAs stream zeros over them, they belong to no task */
  uint32_t zero_rows;         /* texels covered by those segments, summed over views */
  uint64_t device_bytes;      /* HBM held by the plan */
} mvwarp_stats_t;

/* Library identity / diagnostics. */
int         mvwarp_version(void);
const char* mvwarp_last_error(void);

/* ---- host-only helpers (no GPU needed; used by the plan builder and by CPU tests) -------------
 * mvwarp_host_grid: the normalised sampling grid [V,Hg,Wg,2] (x,y in [-1,1] source coordinates) that
 * kornia's transform_points produces from src_norm_trans_dst_norm (fp32 [V,3,3], row-major, the
 * matrix kornia calls `src_norm_trans_dst_norm`: inverse of the normalised homography).  Bit-identical
 * to the CPU oracle's torch op sequence (create_meshgrid -> bmm -> convert_points_from_homogeneous).
 * Replaces: kornia create_meshgrid + transform_points inside warp_perspective (persp_trans_detector.py:69). */
int mvwarp_host_grid(const float* src_norm_trans_dst_norm, int V, int Hg, int Wg, float* grid_out);
/* mvwarp_host_csr: build the inverse (texel -> cells) CSR map of a normalised grid [V,Hg,Wg,2] on the host
 * and report its statistics (device_bytes and inline_coords are not set).
 * Rows are ordered tile-major over 8x8 texel tiles (rows near in the array are near in the image):
 * row r of view v is the texel tex_of[v*Hs*Ws + r] (= ys*Ws + xs).
 * If tex_of_out / row_ptr_out / cells_out / weights_out are non-null they receive the map itself:
 * tex_of_out [V*Hs*Ws], row_ptr_out [V*(Hs*Ws+1)] offsets relative to each view's first entry,
 * cells_out / weights_out [>= nnz] (pass the capacity in `capacity`; MVWARP_EINVAL if too small).
 * Entry order in a row: ascending cell, then tap (nw, ne, sw, se) -- the backward's summation order. */
int mvwarp_host_csr(const float* grid_host, int V, int Hs, int Ws, int Hg, int Wg, mvwarp_stats_t* stats_out,
                    uint32_t* tex_of_out, uint32_t* row_ptr_out, uint32_t* cells_out, float* weights_out,
                    uint64_t capacity);

/* Same for the composed map of a fused up-sample plan (see mvwarp_plan_create_upsampled); hl = wl = 0 gives the plain map.
 * tex_of_out then has V*hl*wl entries and row_ptr_out V*(hl*wl+1). */
int mvwarp_host_csr_upsampled(const float* grid_host, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg,
                              mvwarp_stats_t* stats_out, uint32_t* tex_of_out, uint32_t* row_ptr_out, uint32_t* cells_out,
                              float* weights_out, uint64_t capacity);

/* The composed 3 x 3 low-res windows the fused warp -> conv1 kernel uses on fused up-sample plans (mvwarp_warp_conv1_fwd_ex with a
 * plan from mvwarp_plan_create_upsampled), computed on the host by the kernel's own source (csrc/mvwarp_uptaps.h):
 *   warp(up-sample(low))[v, cell] = sum_{j,i<3, bit j*3+i of mask set} wy[j] * wx[i] * low[v][oy + j][ox + i].
 * T [V,9] as for mvwarp_host_grid; origin_out [V*Hg*Wg][2] = (ox, oy); mask_out [V*Hg*Wg]; wx_out, wy_out [V*Hg*Wg][3].
 * Needs hl <= Hs and wl <= Ws (an up-sample).  Test / introspection entry point: no GPU needed. */
int mvwarp_host_upsample_windows(const float* T, int V, int hl, int wl, int Hs, int Ws, int Hg, int Wg, int32_t* origin_out,
                                 uint32_t* mask_out, float* wx_out, float* wy_out);

/* ---- plan: everything that depends only on the camera rig ---------------------------------------
 * Built once per (proj_mats, shapes, device); holds the per-view matrices (or grid table) and the
 * inverse CSR map used by the deterministic backward.
 * mvwarp_plan_create:            kernels compute coordinates in-kernel from the matrices.
 * mvwarp_plan_create_from_grid:  kernels read an explicit normalised grid [V,Hg,Wg,2] (host fp32),
 *                                e.g. one produced by other torch code.  Same semantics as
 *                                F.grid_sample(src, grid, 'bilinear', 'zeros', align_corners=True).
 * Replaces: the per-forward matrix staging + kornia normalise/invert/meshgrid/transform chain
 * (persp_trans_detector.py:68-69), which the reference redoes on every call. */
int mvwarp_plan_create(const float* src_norm_trans_dst_norm, int V, int Hs, int Ws, int Hg, int Wg,
                       int device, mvwarp_plan_t** plan_out);
int mvwarp_plan_create_from_grid(const float* grid_host, int V, int Hs, int Ws, int Hg, int Wg,
                                 int device, mvwarp_plan_t** plan_out);
/* mvwarp_plan_create_upsampled: SURVEY.md 8(f) rank 1 -- fuse the trunk's bilinear up-sample into the warp.
 * The reference warps  F.interpolate(feat[B,C,hl,wl], [Hs,Ws], mode='bilinear')  (persp_trans_detector.py:65, ATen
 * upsample_bilinear2d, align_corners=False).  Both maps are linear and rig-constant, so their composition is one sparse
 * operator from the LOW-RES feature map to the grid: every grid cell becomes <= 16 (typically 4-9) weighted low-res taps.
 * With such a plan mvwarp_warp_views_fwd / _bwd take feats / grad_feats of shape [B, nviews, C, hl, wl]: the
 * up-sampled tensor is never read (forward) nor written (backward).  Weights are composed in double and rounded once;
 * results agree with the two-step reference to rounding (1e-6 relative), not bit for bit.
 * Replaces: F.interpolate at persp_trans_detector.py:65 (for the warp's consumer) + :68-69,75,77. */
int mvwarp_plan_create_upsampled(const float* src_norm_trans_dst_norm, int V, int hl, int wl, int Hs, int Ws,
                                 int Hg, int Wg, int device, mvwarp_plan_t** plan_out);
int mvwarp_plan_destroy(mvwarp_plan_t* plan);
/* Plan persistence: a plan depends only on the calibration, so it can be stored next to it
 * (the reference rebuilds its matrices from the calibration files on every start, persp_trans_detector.py:17-30, and the
 * sampling grid on every forward).  mvwarp_plan_save writes the device-resident plan to `path` (atomic replace);
 * mvwarp_plan_load re-creates it on `device` without running the host builder.  The file carries a checksum and the
 * library's work-decomposition constants: a truncated / corrupt file or one written by an incompatible build is
 * refused (MVWARP_EINVAL / MVWARP_EUNSUPPORTED), never half-loaded.  mvwarp_plan_matrices returns the [V,3,3]
 * src_norm_trans_dst_norm matrices a plan was built from, so a caller can check a loaded plan against its rig. */
int mvwarp_plan_save(const mvwarp_plan_t* plan, const char* path);
int mvwarp_plan_load(const char* path, int device, mvwarp_plan_t** plan_out);
int mvwarp_plan_matrices(const mvwarp_plan_t* plan, float* T_out);
int mvwarp_plan_stats(const mvwarp_plan_t* plan, mvwarp_stats_t* stats_out);
/* Debug/test: run the device coordinate kernel and copy the normalised grid [V,Hg,Wg,2] to host. */
int mvwarp_plan_device_grid(const mvwarp_plan_t* plan, float* grid_host_out, void* stream);

/* ---- forward: warp + view concat + coord_map in one pass ----------------------------------------
 * For views v in [view0, view0+nviews), batch b, channel c, grid cell (y,x):
 *   out[b, out_chan0 + (v-view0)*C + c, y, x] = bilinear(feats[b, v-view0, c], src_xy(v, y, x)), zeros outside
 * and, if coord_chan >= 0,
 *   out[b, coord_chan, y, x] = x/(Wg-1)*2-1 ;  out[b, coord_chan+1, y, x] = y/(Hg-1)*2-1
 * feats:  logical [B, nviews, C, Hs, Ws], strides feat_strides[5] = (B, V, C, H, W);  `feats` points at view0.
 * out:    logical [B, Ctot, Hg, Wg],      strides out_strides[4]  = (B, C, H, W).
 * Fast path: channel-contiguous (stride C == 1, "channels_last") feats and out; any other strides
 * take the generic kernel.  Replaces persp_trans_detector.py:68-69,75,77 (warp_perspective x V, cat, coord_map). */
int mvwarp_warp_views_fwd(const mvwarp_plan_t* plan, int view0, int nviews,
                          const void* feats, const int64_t feat_strides[5],
                          void* out, const int64_t out_strides[4], int out_chan0, int coord_chan,
                          int B, int C, int dtype, void* stream);

/* ---- backward: deterministic gather through the inverse CSR map ---------------------------------
 *   grad_feats[b, v-view0, c, ys, xs] = sum over CSR row (v, ys, xs) of  w * grad_out[b, out_chan0+(v-view0)*C+c, cell]
 * in a fixed order (ascending cell, then tap nw,ne,sw,se): bitwise run-to-run reproducible, no atomics.
 * Every texel of grad_feats is written (zeros where no cell samples it): no memset needed.
 * The coord_map channels and the matrices receive no gradient (they never require grad in the reference).
 * Replaces CatBackward + grid_sampler_2d_backward (atomicAdd scatter) reached from trainer.py:48. */
int mvwarp_warp_views_bwd(const mvwarp_plan_t* plan, int view0, int nviews,
                          const void* grad_out, const int64_t out_strides[4], int out_chan0,
                          void* grad_feats, const int64_t feat_strides[5],
                          int B, int C, int dtype, void* stream);

/* ---- "list" variants: one tensor per camera ------------------------------------------------------
 * The reference's trunk produces ONE feature tensor per camera inside the camera loop (persp_trans_detector.py:62-66) and appends
 * the warped result to a Python list (:75).  These entry points take the per-camera tensors as they are: feats[v] (resp.
 * grad_feats[v]) is the DEVICE pointer of local view v, logical [B, C, Hs, Ws] with strides feat_strides[4] = (B, C, H, W) shared
 * by all views (same shape, same memory format); `feats` itself is a HOST array of nviews pointers (nviews <= MVWARP_MAX_LIST_VIEWS).
 * One launch covers all views (the pointers travel in the kernel's parameter space), so no [B,V,C,H,W] stack copy -- a 1.86 GB read
 * plus a 1.86 GB write per Wildtrack frame, more than the warp itself moves -- is ever made, and the backward writes each camera's
 * gradient straight into its own tensor.  Everything else as mvwarp_warp_views_fwd / _bwd.
 * Replaces persp_trans_detector.py:68-69,75,77 without torch.stack. */
#define MVWARP_MAX_LIST_VIEWS 64
int mvwarp_warp_views_fwd_list(const mvwarp_plan_t* plan, int view0, int nviews, const void* const* feats,
                               const int64_t feat_strides[4], void* out, const int64_t out_strides[4], int out_chan0,
                               int coord_chan, int B, int C, int dtype, void* stream);
int mvwarp_warp_views_bwd_list(const mvwarp_plan_t* plan, int view0, int nviews, const void* grad_out,
                               const int64_t out_strides[4], int out_chan0, void* const* grad_feats,
                               const int64_t feat_strides[4], int B, int C, int dtype, void* stream);

/* ---- NCHW trunks: planar <-> channel-contiguous feature maps ------------------------------------------------------------
 * mvwarp_layout_transpose: N images of C planes x HW texels.  to_channels_last = 1: src [N][C][HW] (the NCHW tensors the
 * reference's trunk emits, persp_trans_detector.py:63-65) -> dst [N][HW][C]; 0: the reverse (the feature gradient on its way back
 * to an NCHW trunk).  Each image dense; consecutive images *_image_stride elements apart (0 = back to back), so a per-camera tensor
 * [B,C,H,W] can be written into / read from its slice of a stacked [B,V,H,W,C] buffer.  Device pointers, MVWARP_F32 / MVWARP_BF16.  Used by warp_views_list / warp_views_upsampled on
 * the LOW-RES trunk output of fused up-sample plans, where transposing once (2 x 206 MB per Wildtrack frame) and running the
 * channel-contiguous kernels beats gathering 4 bytes per channel and tap from planes. */
int mvwarp_layout_transpose(const void* src, int64_t src_image_stride, void* dst, int64_t dst_image_stride, int N, int C, int HW,
                            int dtype, int to_channels_last, int device, void* stream);

/* ---- host-resident features: move only what the warp can touch ---------------------------------------------------
 * mvwarp_copy_touched_texels: for every batch element, every view in [view0, view0+nviews) and every feature texel (ys, xs) that
 * at least one grid cell samples (= the non-empty rows of the plan's inverse CSR map; 26 % of the texels on the Wildtrack rig,
 * a function of the calibration alone), copy the texel's C channel values from `src` to `dst`.  Both are logical
 * [B, nviews, C, H, W] tensors, channel-contiguous, strides in elements as (B, V, C, H, W); H x W is the feature map the plan's
 * kernels read (the low-res map for fused up-sample plans).  Either pointer may be device memory or PINNED HOST memory
 * (unified addressing: the gather / scatter kernel reads or writes it across PCIe directly), so
 *   host feats   -> device feats        sends the warp everything it will read            (0.49 GB instead of 1.86 GB per frame), and
 *   device grads -> host grads          returns every non-zero row of the feature gradient (the host tensor is zeroed once).
 * mvwarp_plan_touched: the list itself -- touched_base_out [V+1] offsets, touched_out [>= U] linear texel ids ys*W + xs,
 * ascending per view (either may be NULL).
 * Replaces the dense host -> device hops of persp_trans_detector.py:63-68 for callers whose features are host-resident. */
int mvwarp_copy_touched_texels(const mvwarp_plan_t* plan, int view0, int nviews, const void* src, const int64_t src_strides[5],
                               void* dst, const int64_t dst_strides[5], int B, int C, int dtype, void* stream);
int mvwarp_plan_touched(const mvwarp_plan_t* plan, uint32_t* touched_base_out, uint32_t* touched_out, uint64_t capacity);

#ifdef MVWARP_DEV
/* ---- tensor-core building blocks (north_star (c)): validation hooks, DEVELOPMENT BUILD ONLY ------------------
 * (mvdet_b200/libmvwarp_dev.so, compiled with -DMVWARP_DEV; the product library does not export them)
 * mvwarp_debug_gemm: D[M,256] (fp32) = A[M,K] (bf16, row-major) x B[256,K]^T (bf16, row-major) on tcgen05 / TMEM.
 * Validation hook for the descriptor / swizzle / TMEM conventions the fused warp -> conv1 kernel relies on
 * (M % 128 == 0, K % 64 == 0, device pointers). */
int mvwarp_debug_gemm(const void* A, const void* B, void* D, int M, int N, int K, void* stream);
/* The same with a probe layout of the A tile in shared memory: 8-row groups sbo_bytes apart (1024 = canonical), the tile shifted by
 * start_cells * 128 bytes, chunks swizzled by absolute address bits, descriptor stride offset = sbo_bytes (+ the base-offset
 * field if use_base_offset).  Development hook: which non-canonical operand layouts does the tensor core read correctly? */
int mvwarp_debug_gemm_layout(const void* A, const void* B, void* D, int M, int N, int K, int sbo_bytes, int start_cells,
                             int use_base_offset, void* stream);
#endif /* MVWARP_DEV */

/* mvwarp_warp_conv1_fwd (north_star (c)): the warp fused into map_classifier[0] = Conv2d(V*C+2, 512, 3, padding=1)
 * (persp_trans_detector.py:51,81) as a tcgen05 implicit GEMM; the warped tensor is never materialised.  Runs in CTA pairs
 * (cta_group::2, M = 256 per MMA, half of every weight block per SM) whose nine conv taps the tensor core reads in place from
 * the warped halo patch; needs a device with thread-block clusters and TMA (sm_100a).
 *   out[b, n, y, x] = (bias ? bias[n] : 0) + sum_{dy,dx} sum_{ch} W[n, ch, dy, dx] * X[b, ch, y+dy-1, x+dx-1]   (+ ReLU if relu)
 * with X = the warp of views [view0, view0+nviews) (+ the two coord_map channels if with_coord), zero padded.
 * feats: bf16, logical [B, nviews, C, Hs, Ws], channel-contiguous, C % 64 == 0.  out: fp32 [B, 512, Hg, Wg] channels-last.
 * wpack: the weight pre-packed by mvdet_b200.conv1.pack_conv1_weight: per (view, 64-channel block, tap) one
 *        [512][64] bf16 block whose rows carry the 128-byte shared-memory swizzle, then one block for the coord channels.
 * X is rounded to bf16 before the contraction (fp32 accumulate in TMEM).  View shards pass their own weight slices,
 * bias / with_coord only on the coord-map owner, relu = 0, and all-reduce the partial sums (mvdet_b200.sharding).
 * row0 / nrows restrict the launch to the grid rows [row0, row0+nrows) (row0 % 8 == 0; nrows <= 0 = the whole grid): the
 * sharded path computes the map band by band so that the all-reduce of one band overlaps the GEMM of the next. */
int mvwarp_warp_conv1_fwd(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats,
                          const int64_t feat_strides[5], const void* wpack, const float* bias, float* out,
                          const int64_t out_strides[4], int B, int C, int relu, int with_coord, int row0, int nrows,
                          void* stream);

/* The same with a choice of output dtype (MVWARP_F32 or MVWARP_BF16; bf16 feeds mvwarp_conv3x3_fwd below) and an optional
 * scratch buffer.  The launch has one CTA per (frame, 8x16 cell tile); when that count is not a multiple of the SM count the
 * last partial wave would keep a few SMs busy for a whole tile time (Wildtrack B=1: 345 tiles on 148 SMs = 2.33 waves cost
 * 3).  With scratch (device memory, 16-byte aligned, >= mvwarp_warp_conv1_scratch_bytes(...) bytes, owned by the caller)
 * the leftover tiles are split along K (by views) over the idle SMs, write fp32 partial sums to scratch, and a second
 * small kernel adds them in a fixed order (deterministic) with bias / ReLU.  scratch == NULL: no split. */
int mvwarp_warp_conv1_fwd_ex(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats,
                             const int64_t feat_strides[5], const void* wpack, const float* bias, void* out, int out_dtype,
                             const int64_t out_strides[4], int B, int C, int relu, int with_coord, int row0, int nrows,
                             void* scratch, int64_t scratch_bytes, void* stream);
int64_t mvwarp_warp_conv1_scratch_bytes(const mvwarp_plan_t* plan, int nviews, int B, int row0, int nrows);

/* mvwarp_warp_conv1_fwd_p2p: the view-sharded fused conv1 with the reduce-scatter of its partial sums fused into the GEMM's
 * epilogue ("each rank computes its per-view conv1 partial sum, partials are combined with a reduce over NVLink").
 * Rank r of nranks (<= 8, one process per GPU) owns the grid rows [band_row0[r], band_row0[r+1]).  peer_inbox[r] is rank r's
 * inbox as mapped into THIS process (symmetric / IPC memory; fp32 [nranks][B][band_rows_max][Wg][512]): the epilogue of every
 * tile stores this rank's raw partial sums (no bias, no ReLU) for the tile's cells into slot my_rank of the owner's inbox --
 * peer-to-peer stores over NVLink / NVSwitch that ride along with the tensor-core work, no extra kernel, no collective call.
 * After a cross-rank barrier every rank adds the nranks slots of its own inbox (+ bias, ReLU): mvdet_b200.sharding.
 * Replaces: conv1 partial + torch.distributed.all_reduce in the view-sharded map_classifier[0] (persp_trans_detector.py:51,81). */
/* mvwarp_inbox_reduce: the owner's side of that exchange.  out[b][i] = [relu](bias[i % 512] + sum_s inbox[s][b][i]) for the nslots
 * slots (added in slot order: deterministic), i < n_per_b (a multiple of 512: whole cells); slot_pitch / b_pitch in elements;
 * out contiguous [B][n_per_b], MVWARP_F32 or MVWARP_BF16.  One pass over the data. */
int mvwarp_inbox_reduce(const float* inbox, int nslots, int64_t slot_pitch, int64_t b_pitch, int B, int64_t n_per_b,
                        const float* bias, int relu, void* out, int out_dtype, int device, void* stream);
int mvwarp_warp_conv1_fwd_p2p(const mvwarp_plan_t* plan, int view0, int nviews, const void* feats, const int64_t feat_strides[5],
                              const void* wpack, int B, int C, int with_coord, int nranks, int my_rank, const uint64_t* peer_inbox,
                              const int32_t* band_row0, int band_rows_max, void* stream);

/* ---- backward of relu(conv1 + bias): the element-wise front end (training step, trainer.py:48) ---------------------------
 * mvwarp_relu_bwd_cast:  g16[b,n,y,x] = bf16(grad_out[b,n,y,x] * (out[b,n,y,x] > 0))  (out == NULL: no mask) and, if bias_grad != NULL,
 * bias_grad[n] = sum_{b,y,x} of the masked fp32 gradient in a fixed order (deterministic) -- one pass instead of three torch ops.
 * grad_out / out: fp32 [B,C,H,W] channels-last; g16: bf16 channels-last with any (8-element aligned) strides, e.g. the interior of
 * the zero-padded buffer the weight-gradient GEMMs read.  Strides as (B, C, H, W) in elements.  scratch: device floats,
 * >= mvwarp_relu_bwd_scratch_floats(C, device) (only needed with bias_grad).  C must divide 2048 and be a multiple of 8. */
int64_t mvwarp_relu_bwd_scratch_floats(int C, int device);
int mvwarp_relu_bwd_cast(const float* grad_out, const int64_t go_strides[4], const float* out, const int64_t out_strides[4],
                         void* g16, const int64_t g_strides[4], float* bias_grad, float* scratch, int64_t scratch_floats,
                         int B, int H, int W, int C, int device, void* stream);

/* ---- the rest of map_classifier (persp_trans_detector.py:53-54, applied at :81) -------------
 * mvwarp_conv3x3_fwd: out = [relu](conv2d(x, W, bias, padding = dilation, dilation = dilation)) for a 3x3 kernel with
 * 512 output channels -- map_classifier[2] = Conv2d(512, 512, 3, padding=2, dilation=2) + ReLU -- as a tcgen05 implicit
 * GEMM whose A operand is fetched by TMA tensor copies (one [8 rows][16 cells][64 channels] box per tap and channel
 * block, zero padding = the TMA unit's out-of-bounds fill).
 *   x:     bf16, logical [B, Cin, H, W] channels-last (stride C == 1), Cin % 64 == 0; x_strides = (B, C, H, W) in elements.
 *   wpack: the reference weight [512, Cin, 3, 3] packed by mvdet_b200.map_head.pack_conv3x3_weight (= pack_conv1_weight
 *          for one "view" without the coord block): [Cin/64 * 9][512][64] bf16, rows carrying the 128-byte swizzle.
 *   out:   [B, 512, H, W] channels-last, out_dtype MVWARP_F32 or MVWARP_BF16.  fp32 accumulation in TMEM.
 * mvwarp_conv3x3_head_fwd: out[b, y, x] = sum_{tap, c} w[tap][c] * x[b, c, y + (tap/3-1)*d, x + (tap%3-1)*d], zero padded
 * -- map_classifier[4] = Conv2d(512, 1, 3, padding=4, dilation=4, bias=False); w = bf16 [9][Cin] (tap-major), out fp32
 * [B, 1, H, W] (strides given as (B, C, H, W)).  Memory-bound: plain CUDA cores, fp32 accumulation. */
int mvwarp_conv3x3_fwd(const void* x, const int64_t x_strides[4], const void* wpack, const float* bias, void* out,
                       const int64_t out_strides[4], int B, int H, int W, int Cin, int dilation, int relu, int out_dtype,
                       int device, void* stream);
int mvwarp_conv3x3_head_fwd(const void* x, const int64_t x_strides[4], const void* w, float* out, const int64_t out_strides[4],
                            int B, int H, int W, int Cin, int dilation, int device, void* stream);
/* The same with caller-provided scratch (mvwarp_conv3x3_head_scratch_bytes(B, H, W) bytes, 16-byte aligned; any allocator): with it
 * and Cin % 32 == 0 the convolution runs in two passes that read every activation ONCE -- nine partial dot products per cell
 * (P[cell][tap], a skinny GEMM on the tensor cores), then a 9-term stencil over P.  Without scratch: the one-pass kernel. */
int64_t mvwarp_conv3x3_head_scratch_bytes(int B, int H, int W);
int mvwarp_conv3x3_head_fwd_ex(const void* x, const int64_t x_strides[4], const void* w, float* out, const int64_t out_strides[4],
                               int B, int H, int W, int Cin, int dilation, void* scratch, int64_t scratch_bytes, int device,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVWARP_H_ */
