"""Fused warp -> map_classifier[0] (north_star (c)): host side.

``map_classifier[0]`` is ``Conv2d(V*C+2, 512, 3, padding=1)`` (/root/reference/multiview_detector/models/
persp_trans_detector.py:51) applied to the concat that ``warp_views`` produces (:81).  ``warp_conv1`` computes
``conv(warp_views(feats))`` as one tcgen05 implicit GEMM without materialising the warped tensor; this module
packs the reference's weight ``[512, V*C+2, 3, 3]`` (e.g. straight from a reference checkpoint, ``main.py:121``)
into the shared-memory image the kernel's B operand expects.  CUDA only, bf16 inputs, fp32 accumulate / output.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from .plan import get_plan, get_plan_upsampled

KB = 64          # channels per K block
N_OUT = 512


def pack_conv1_weight(weight: torch.Tensor, num_views: int, channels: int, with_coord: bool = True) -> torch.Tensor:
    """[512, V*C(+2), 3, 3] -> bf16 [(V * C/64 * 9 (+1)), 512, 64]: one block per (view, 64-channel block, tap), K-major,
    every 128-byte row with its 16-byte chunks XOR-swizzled by (row & 7) -- the tensor core's SWIZZLE_128B layout.
    The optional last block holds the two coord_map channels of all nine taps at k = tap*2 + ch (rest zero)."""
    V, Cc = int(num_views), int(channels)
    if weight.shape[0] != N_OUT or tuple(weight.shape[2:]) != (3, 3) or weight.shape[1] != V * Cc + (2 if with_coord else 0):
        raise ValueError(f"weight must be [512, {V * Cc + (2 if with_coord else 0)}, 3, 3]. Got {tuple(weight.shape)}")
    if Cc % KB:
        raise ValueError("channels must be a multiple of 64")
    w = weight.detach().to(torch.bfloat16)
    ncb = Cc // KB
    main = w[:, :V * Cc].reshape(N_OUT, V, ncb, KB, 9)                  # n, v, cb, k, tap
    blocks = main.permute(1, 2, 4, 0, 3).reshape(V * ncb * 9, N_OUT, KB)  # (v, cb, tap), n, k
    if with_coord:
        cw = w[:, V * Cc:].reshape(N_OUT, 2, 9).permute(0, 2, 1).reshape(N_OUT, 18)   # n, tap*2 + ch
        cblock = torch.zeros(1, N_OUT, KB, dtype=torch.bfloat16, device=w.device)
        cblock[0, :, :18] = cw
        blocks = torch.cat([blocks, cblock], dim=0)
    # swizzle: chunk c of row n lands at chunk position c ^ (n & 7)
    nb = blocks.shape[0]
    chunks = blocks.reshape(nb, N_OUT, 8, 8)
    n_idx = torch.arange(N_OUT, device=w.device)[:, None]
    pos = torch.arange(8, device=w.device)[None, :]
    src = pos ^ (n_idx & 7)                                              # chunk stored at position pos comes from chunk src
    packed = torch.gather(chunks, 2, src[None, :, :, None].expand(nb, N_OUT, 8, 8))
    return packed.reshape(nb, N_OUT, KB).contiguous()


def unpack_conv1_weight(wpack: torch.Tensor, num_views: int, channels: int, with_coord: bool = True) -> torch.Tensor:
    """Inverse of :func:`pack_conv1_weight`: the packed blocks back to the reference's ``map_classifier.0.weight`` layout
    ``[512, V*C(+2), 3, 3]`` (bf16 values, exact), e.g. to write a checkpoint the reference's ``main.py:121`` can load."""
    V, Cc = int(num_views), int(channels)
    ncb = Cc // KB
    nb = V * ncb * 9 + (1 if with_coord else 0)
    if Cc % KB or tuple(wpack.shape) != (nb, N_OUT, KB):
        raise ValueError(f"wpack must be [{nb}, 512, 64] for V={V}, C={Cc}. Got {tuple(wpack.shape)}")
    chunks = wpack.reshape(nb, N_OUT, 8, 8)
    n_idx = torch.arange(N_OUT, device=wpack.device)[:, None]
    pos = torch.arange(8, device=wpack.device)[None, :]
    src = pos ^ (n_idx & 7)                                              # the XOR swizzle is its own inverse
    blocks = torch.gather(chunks, 2, src[None, :, :, None].expand(nb, N_OUT, 8, 8)).reshape(nb, N_OUT, KB)
    main = blocks[:V * ncb * 9].reshape(V, ncb, 9, N_OUT, KB).permute(3, 0, 1, 4, 2).reshape(N_OUT, V * Cc, 3, 3)
    if not with_coord:
        return main.contiguous()
    cw = blocks[-1][:, :18].reshape(N_OUT, 9, 2).permute(0, 2, 1).reshape(N_OUT, 2, 3, 3)
    return torch.cat([main, cw], dim=1).contiguous()


def conv1_from_state_dict(state_dict, num_views: int, channels: int = 512, prefix: str = "map_classifier.0.",
                          device=None):
    """``(wpack, bias)`` for :func:`warp_conv1` from a reference checkpoint's state dict (``torch.load`` of the file
    ``main.py:121`` saves / ``main.py:125`` loads: keys ``map_classifier.0.weight`` [512, V*C+2, 3, 3] and ``.bias`` [512]).
    Keys with a DataParallel ``module.`` prefix are accepted too."""
    def find(name):
        for k in (prefix + name, "module." + prefix + name):
            if k in state_dict:
                return state_dict[k]
        raise KeyError(f"{prefix + name} not in the state dict")
    w, b = find("weight"), find("bias")
    if device is not None:
        w, b = w.to(device), b.to(device)
    return pack_conv1_weight(w, num_views, channels, with_coord=True), b.detach().float().contiguous()


def conv1_to_state_dict(wpack: torch.Tensor, bias: torch.Tensor, num_views: int, channels: int = 512,
                        prefix: str = "map_classifier.0.", dtype=torch.float32):
    """The two reference state-dict entries back from the packed weight (values are the bf16-rounded ones)."""
    return {prefix + "weight": unpack_conv1_weight(wpack, num_views, channels).to(dtype),
            prefix + "bias": bias.detach().to(dtype).clone()}


def pack_conv1_dgrad_weight(weight: torch.Tensor, num_views: int, channels: int):
    """Per-view packed weights of the DATA gradient of ``map_classifier[0]``: the gradient w.r.t. view v's warped channels is itself a
    3x3 convolution, ``gX_v[c] = sum_{n,ky,kx} g[n](y+1-ky, x+1-kx) * W[n, v*C+c, ky, kx]``, i.e. ``conv3x3(g, W'_v, padding=1)`` with
    ``W'_v[c, n, ky, kx] = W[n, v*C + c, 2-ky, 2-kx]`` (in/out transposed, taps flipped).  Returns a list of V packed blocks for
    :func:`mvdet_b200.map_head.conv3x3_cl` (the TMA-fed tcgen05 convolution; needs C == 512 output channels per view)."""
    V, Cc = int(num_views), int(channels)
    if Cc != N_OUT:
        raise ValueError("the tcgen05 data-gradient path needs 512 channels per view (the convolution kernel has 512 outputs)")
    w = weight.detach()[:, :V * Cc].reshape(N_OUT, V, Cc, 3, 3)
    out = []
    for v in range(V):
        wt = w[:, v].flip(-1, -2).permute(1, 0, 2, 3).contiguous()            # [c, n, ky, kx]
        out.append(pack_conv1_weight(wt, 1, N_OUT, with_coord=False))
    return out


_pad_cache = {}


def _padded_buffers(device, B: int, Hg: int, Wg: int, P: int):
    """``(Xp [B,Hg+2,Wg+2,P], Gp [B,Hg+2,Wg+2,512])`` bf16, zeroed ONCE per shape: every step overwrites the interiors completely (the
    warp forward writes every cell, zeros included), so the zero borders / pad channels survive and no per-step memset is needed."""
    key = (str(device), B, Hg, Wg, P)
    hit = _pad_cache.get("k")
    if hit is None or hit[0] != key:
        _pad_cache.pop("k", None)
        Xp = torch.zeros((B, Hg + 2, Wg + 2, P), dtype=torch.bfloat16, device=device)
        Gp = torch.zeros((B, Hg + 2, Wg + 2, N_OUT), dtype=torch.bfloat16, device=device)
        _pad_cache["k"] = (key, Xp, Gp)
    return _pad_cache["k"][1], _pad_cache["k"][2]


def relu_bwd_cast(grad_out: torch.Tensor, out: Optional[torch.Tensor], g16: torch.Tensor, want_bias_grad: bool) -> Optional[torch.Tensor]:
    """``g16 <- bf16(grad_out * (out > 0))`` and the bias gradient ``sum_{b,y,x}`` of the masked gradient in ONE pass
    (``mvwarp_relu_bwd_cast``).  grad_out / out fp32 channels-last [B,512,H,W]; g16 a bf16 channels-last view of any strides."""
    B, Cn, H, W = grad_out.shape
    lib = _lib.load()
    dev = grad_out.device
    bias_grad = scratch = None
    nscr = 0
    if want_bias_grad:
        nscr = int(lib.mvwarp_relu_bwd_scratch_floats(Cn, dev.index))
        scratch = torch.empty(nscr, dtype=torch.float32, device=dev)
        bias_grad = torch.empty(Cn, dtype=torch.float32, device=dev)
    rc = lib.mvwarp_relu_bwd_cast(C.c_void_p(grad_out.data_ptr()), _lib.i64x(grad_out.stride()),
                                  C.c_void_p(out.data_ptr()) if out is not None else None,
                                  _lib.i64x(out.stride()) if out is not None else None,
                                  C.c_void_p(g16.data_ptr()), _lib.i64x(g16.stride()),
                                  C.c_void_p(bias_grad.data_ptr()) if want_bias_grad else None,
                                  C.c_void_p(scratch.data_ptr()) if want_bias_grad else None, nscr, B, H, W, Cn, dev.index,
                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "mvwarp_relu_bwd_cast")
    return bias_grad


def conv1_wgrad_gemm(plan, feats: torch.Tensor, Xp: torch.Tensor, Gp: torch.Tensor, with_coord: bool = True) -> torch.Tensor:
    """Weight gradient of ``map_classifier[0]`` as NINE plain GEMMs on shifted views -- no im2col, no cuDNN.

    ``gW[n, ch, ky, kx] = sum_{b,y,x} g[b,n,y,x] * X[b,ch,y+ky-1,x+kx-1]`` with X = warp_views(feats) zero-padded.  X goes into the
    interior of the zero-bordered buffer ``Xp[B, Hg+2, Wg+2, P]`` (P = V*C+2 rounded up to 8 channels, so every row is 16-byte
    aligned) -- the warp forward kernel stores straight into that view -- and ``g`` sits in the top-left corner of the zero-bordered
    ``Gp[B, Hg+2, Wg+2, 512]`` (written there by :func:`relu_bwd_cast`).  In the flattened position index
    ``p = (b*(Hg+2) + y)*(Wg+2) + x`` the tap (ky, kx) is then a CONSTANT shift ``ky*(Wg+2) + kx``:
    ``gW[:, :, ky, kx] = Gp[:K].T @ Xp[shift : shift+K]``  (K = all positions minus the largest shift; the positions cut off hold
    zeros of Gp).  Each is a cuBLAS bf16 GEMM (M = 512, N = P, K = B*(Hg+2)*(Wg+2)) with fp32 accumulation and fp32 output: a
    plain library GEMM is all this contraction needs once the shifts are views (DESIGN.md section 6 has the traffic analysis of a
    hand-written tcgen05 version and why it would not beat this)."""
    from . import ops
    B, V, Cc = feats.shape[:3]
    Hg, Wg = plan.grid_hw
    ctot = V * Cc + (2 if with_coord else 0)
    P = Xp.shape[-1]
    ops.warp_views_fwd_into(plan, feats, Xp[:, 1:Hg + 1, 1:Wg + 1, :ctot].permute(0, 3, 1, 2), 0, 0, V * Cc if with_coord else -1)
    Xf, Gf = Xp.view(-1, P), Gp.view(-1, N_OUT)
    K = Xf.shape[0] - (2 * (Wg + 2) + 2)
    Gt = Gf[:K].t()                                                     # [512, K] view (column-major operand for cuBLAS)
    # the three kx taps of one ky are three GEMMs whose B operands start one row apart: ONE batched GEMM over an overlapping
    # strided view (batch stride = one row of Xp).  Measured at B=8: 3 x bmm(batch 3) 26.9 ms per training step against 30.3 ms
    # for 9 separate GEMMs -- each has only 56 output tiles of 128 x 256, a batch of three fills the 148 SMs
    Gb = Gt.unsqueeze(0).expand(3, N_OUT, K)
    rows = []
    for ky in range(3):
        Xk = torch.as_strided(Xf, (3, K, P), (P, P, 1), Xf.storage_offset() + ky * (Wg + 2) * P)
        rows.append(torch.bmm(Gb, Xk, out_dtype=torch.float32))       # bf16 x bf16 -> fp32 accumulate, fp32 out
    gw = torch.stack(rows).view(3, 3, N_OUT, P)
    return gw[..., :ctot].permute(2, 3, 0, 1)                           # [512, V*C+2, 3, 3] (a view)


_dgrad_cache = {}
_wpack_cache = {}


def _dgrad_packs(weight: torch.Tensor, V: int, Cc: int):
    """Packed data-gradient weights, re-packed only when the weight changes (keyed on storage address + version counter)."""
    key = (weight.data_ptr(), weight._version, weight.device, V, Cc)
    hit = _dgrad_cache.get("k")
    if hit is None or hit[0] != key:
        _dgrad_cache["k"] = (key, pack_conv1_dgrad_weight(weight, V, Cc))
    return _dgrad_cache["k"][1]


def clear_caches() -> None:
    """Drop the training path's cached buffers: the zero-bordered ``Xp`` / ``Gp`` (2.9 GB at the Wildtrack shape with B=8), the packed
    forward weight and the packed data-gradient weights.  They are rebuilt on the next call."""
    _pad_cache.clear()
    _dgrad_cache.clear()
    _wpack_cache.clear()


class _WarpConv1Fn(torch.autograd.Function):
    """Training wrapper: fused tcgen05 forward.  Backward, no cuDNN anywhere:
    * ``mvwarp_relu_bwd_cast``: ReLU mask + bf16 rounding + bias gradient in one pass, written into the zero-bordered buffer ``Gp``;
    * DATA gradient: the library's own TMA-fed tcgen05 convolution, one launch per view with the transposed, tap-flipped weight
      block, into a 16-byte aligned ``[B,Hg,Wg,V*C]`` buffer which the deterministic CSR warp backward reads on its aligned path;
    * WEIGHT gradient: the warped tensor is re-materialised into the zero-bordered buffer ``Xp`` (0.12 ms per Wildtrack frame) and the
      nine taps are nine cuBLAS GEMMs on shifted views of ``Gp`` / ``Xp`` (:func:`conv1_wgrad_gemm`).
    ``MVWARP_CUDNN_DGRAD=1`` / ``MVWARP_CUDNN_WGRAD=1`` switch either gradient back to ``aten.convolution_backward`` (comparison runs)."""

    @staticmethod
    def forward(ctx, img_feats, weight, bias, proj_mats, grid_shape, wpack, relu, upsample_shape=None):
        out = warp_conv1(img_feats, proj_mats, grid_shape, wpack, bias, relu=relu, upsample_shape=upsample_shape)
        ctx.save_for_backward(img_feats, weight, bias if bias is not None else torch.empty(0), out if relu else torch.empty(0))
        ctx.proj_mats, ctx.grid_shape, ctx.relu, ctx.has_bias = proj_mats, tuple(int(g) for g in grid_shape), relu, bias is not None
        ctx.upsample_shape = None if upsample_shape is None else tuple(int(u) for u in upsample_shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from . import ops
        from .map_head import conv3x3_cl
        img_feats, weight, bias, out = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        feats = img_feats.detach()
        B, V, Cc, Hs, Ws = feats.shape
        # fused up-sample: the SAME backward over the composed plan -- the re-materialised warp reads the low-res map through the
        # composed forward CSR, and the CSR warp backward writes the low-res gradient directly (no F.interpolate backward)
        if ctx.upsample_shape is not None:
            plan = get_plan_upsampled(ctx.proj_mats, (Hs, Ws), ctx.upsample_shape, ctx.grid_shape, feats.device)
        else:
            plan = get_plan(ctx.proj_mats, (Hs, Ws), ctx.grid_shape, feats.device)
        Hg, Wg = plan.grid_hw
        ctot = V * Cc + 2
        Xp, Gp = _padded_buffers(feats.device, B, Hg, Wg, (ctot + 7) // 8 * 8)
        # ReLU mask + bf16 rounding + bias gradient in one pass, written straight into the padded buffer of the weight-gradient GEMMs;
        # g16 is the [B,512,Hg,Wg] channels-last VIEW of that buffer's interior, which the data-gradient convolution reads through TMA
        g16 = Gp[:, :Hg, :Wg].permute(0, 3, 1, 2)
        go = grad_out
        if go.dtype != torch.float32 or go.stride(1) != 1:
            go = go.float().contiguous(memory_format=torch.channels_last)
        g_b = relu_bwd_cast(go, out if ctx.relu else None, g16, need_b)
        own_dgrad = need_x and Cc == N_OUT and os.environ.get("MVWARP_CUDNN_DGRAD") != "1"
        gemm_wgrad = need_w and os.environ.get("MVWARP_CUDNN_WGRAD") != "1"
        gx = gw = None
        if gemm_wgrad:
            gw = conv1_wgrad_gemm(plan, feats, Xp, Gp)
        if (need_w and not gemm_wgrad) or (need_x and not own_dgrad):      # comparison paths: cuDNN's convolution backward
            x = ops.warp_views_with_plan(feats, plan)
            w = weight.detach().to(x.dtype).contiguous(memory_format=torch.channels_last)
            gx, gw2, _ = torch.ops.aten.convolution_backward(g16.contiguous(memory_format=torch.channels_last), x, w, None, [1, 1], [1, 1],
                                                             [1, 1], False, [0, 0], 1, [need_x and not own_dgrad, need_w and not gemm_wgrad, False])
            gw = gw if gemm_wgrad else gw2
        if own_dgrad:
            gx_buf = torch.empty((B, Hg, Wg, V * Cc), dtype=feats.dtype, device=feats.device)    # pitch V*C: every cell row 16-byte aligned
            gx = gx_buf.permute(0, 3, 1, 2)
            for v, wp in enumerate(_dgrad_packs(weight, V, Cc)):
                conv3x3_cl(g16, wp, None, dilation=1, relu=False, out_dtype=feats.dtype, out=gx[:, v * Cc:(v + 1) * Cc])
        g_feats = None
        if need_x:
            g_feats = torch.empty((B, V, Hs, Ws, Cc), dtype=feats.dtype, device=feats.device).permute(0, 1, 4, 2, 3)
            ops.warp_views_bwd_into(plan, gx, g_feats)        # deterministic CSR gather; the two coord channels of gx get no gradient
        g_w = gw.to(weight.dtype) if need_w else None
        return g_feats, g_w, (g_b.to(bias.dtype) if need_b else None), None, None, None, None, None


def warp_conv1_trainable(img_feats: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], proj_mats: torch.Tensor,
                         grid_shape: Sequence[int], wpack: Optional[torch.Tensor] = None, relu: bool = True,
                         upsample_shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    """``relu(map_classifier[0](warp_views(img_feats)))`` with gradients for ``img_feats``, ``weight`` [512, V*C+2, 3, 3]
    and ``bias``: fused tcgen05 forward; backward = ReLU-mask / bias-gradient kernel, data gradient on the TMA tcgen05 convolution,
    weight gradient as GEMMs on shifted views, deterministic CSR warp backward (see :class:`_WarpConv1Fn`).
    Pass ``wpack`` (from :func:`pack_conv1_weight`) to skip re-packing when the weight has not changed.  With ``upsample_shape``
    ``img_feats`` is the trunk's low-res output and ``F.interpolate`` (persp_trans_detector.py:65) is fused into forward and
    backward alike (see :func:`warp_conv1`): the gradient comes back at low resolution."""
    if wpack is None:
        # re-packed only when the weight changed (storage address + version counter): once per optimizer step in training,
        # never in an evaluation loop
        key = (weight.data_ptr(), weight._version, weight.device, img_feats.shape[1], img_feats.shape[2])
        hit = _wpack_cache.get("k")
        if hit is None or hit[0] != key:
            _wpack_cache["k"] = (key, pack_conv1_weight(weight, img_feats.shape[1], img_feats.shape[2]))
        wpack = _wpack_cache["k"][1]
    return _WarpConv1Fn.apply(img_feats, weight, bias, proj_mats, grid_shape, wpack, relu, upsample_shape)


def warp_conv1(img_feats: torch.Tensor, proj_mats: torch.Tensor, grid_shape: Sequence[int], wpack: torch.Tensor,
               bias: Optional[torch.Tensor] = None, relu: bool = True, with_coord: bool = True,
               out: Optional[torch.Tensor] = None, rows: Optional[Sequence[int]] = None,
               out_dtype: torch.dtype = torch.float32, upsample_shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    """relu(conv3x3(warp_views(img_feats, proj_mats, grid_shape), W) + bias) as one fused kernel (forward only).

    ``upsample_shape`` (the reference's ``upsample_shape``, persp_trans_detector.py:41): ``img_feats`` is then the trunk's LOW-RES
    output and the kernel computes ``conv(warp_views(F.interpolate(img_feats, upsample_shape, mode='bilinear')))`` -- line 65 of
    the reference fused in as well: the producers compose up-sample and warp per halo cell into one 3 x 3 low-res window
    (csrc/mvwarp_uptaps.h), so the 512-channel up-sampled tensor never exists.  ``proj_mats`` stay the reference's matrices (built
    for the up-sampled map).  The blend runs in fp32 from the bf16 low-res values: one bf16 rounding (of the blended value) where
    the two-step path rounds twice.

    img_feats [B,V,C,H,W] bf16 CUDA, channel-contiguous; wpack from :func:`pack_conv1_weight`; bias fp32 [512] or None.
    Returns [B,512,Hg,Wg] (channels-last) in ``out_dtype``: fp32, or bf16 to feed :func:`mvdet_b200.map_head.conv3x3_cl`
    (map_classifier[2]) directly.  ``out`` re-uses a result tensor; ``rows=(y0, y1)`` computes only the grid
    rows [y0, y1) of it (y0 a multiple of 8), which is how the sharded path overlaps the reduce of one band with the GEMM
    of the next."""
    if not img_feats.is_cuda:
        raise _lib.MvwarpError("warp_conv1 is CUDA-only (no CPU fallback)")
    if img_feats.dtype != torch.bfloat16 or img_feats.dim() != 5 or img_feats.stride(2) != 1:
        raise ValueError("warp_conv1 needs bf16 channel-contiguous features [B,V,C,H,W]")
    B, V, Cc, Hs, Ws = img_feats.shape
    nblocks = V * (Cc // KB) * 9 + (1 if with_coord else 0)
    if wpack.dtype != torch.bfloat16 or tuple(wpack.shape) != (nblocks, N_OUT, KB) or not wpack.is_contiguous():
        raise ValueError(f"wpack must be a contiguous bf16 [{nblocks}, 512, 64] tensor from pack_conv1_weight")
    if upsample_shape is not None:
        if Hs > int(upsample_shape[0]) or Ws > int(upsample_shape[1]):
            raise ValueError(f"upsample_shape {tuple(upsample_shape)} is smaller than the feature map {(Hs, Ws)}")
        plan = get_plan_upsampled(proj_mats, (Hs, Ws), upsample_shape, grid_shape, img_feats.device)
    else:
        plan = get_plan(proj_mats, (Hs, Ws), grid_shape, img_feats.device)
    Hg, Wg = plan.grid_hw
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("warp_conv1 writes fp32 or bf16")
    if out is None:
        out = torch.empty((B, Hg, Wg, N_OUT), dtype=out_dtype, device=img_feats.device).permute(0, 3, 1, 2)
    elif tuple(out.shape) != (B, N_OUT, Hg, Wg) or out.dtype != out_dtype or out.stride(1) != 1:
        raise ValueError(f"out must be a channels-last {out_dtype} [B,512,Hg,Wg] tensor")
    row0, nrows = (int(rows[0]), int(rows[1]) - int(rows[0])) if rows is not None else (0, 0)
    bias_ptr = None
    if bias is not None:
        bias = bias.detach().to(device=img_feats.device, dtype=torch.float32).contiguous()
        bias_ptr = C.c_void_p(bias.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream(img_feats.device).cuda_stream)
    # scratch for the K-split of the last partial wave of tiles (include/mvwarp.h); comes from torch's caching allocator
    nbytes = int(_lib.load().mvwarp_warp_conv1_scratch_bytes(plan.handle, V, B, row0, nrows))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=img_feats.device) if nbytes else None
    rc = _lib.load().mvwarp_warp_conv1_fwd_ex(plan.handle, 0, V, C.c_void_p(img_feats.data_ptr()), _lib.i64x(img_feats.stride()),
                                              C.c_void_p(wpack.data_ptr()), bias_ptr, C.c_void_p(out.data_ptr()),
                                              _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32,
                                              _lib.i64x(out.stride()), B, Cc, int(relu), int(with_coord), row0, nrows,
                                              C.c_void_p(scratch.data_ptr()) if scratch is not None else None, nbytes, stream)
    _lib.check(rc, "mvwarp_warp_conv1_fwd")
    return out
