"""The rest of ``map_classifier`` behind the fused warp -> conv1: host side.

The reference's ground-plane head is (/root/reference/multiview_detector/models/persp_trans_detector.py:51-54, applied at :81)

    map_classifier = Sequential(Conv2d(V*512+2, 512, 3, padding=1), ReLU(),
                                Conv2d(512, 512, 3, padding=2, dilation=2), ReLU(),
                                Conv2d(512, 1, 3, padding=4, dilation=4, bias=False))

:class:`MapHead` runs it as three kernels of libmvwarp.so with bf16 activations and fp32 accumulation: the fused
warp -> conv1 (tcgen05, A operand produced by the warp), ``conv3x3_cl`` for the dilated 512 -> 512 convolution
(tcgen05, A operand fetched by TMA tensor copies) and ``conv3x3_head`` for the 512 -> 1 convolution (memory-bound, CUDA
cores).  Weights come straight from a reference checkpoint's state dict.  CUDA only, inference (forward) only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .conv1 import KB, N_OUT, conv1_from_state_dict, pack_conv1_weight, warp_conv1


def pack_conv3x3_weight(weight: torch.Tensor) -> torch.Tensor:
    """[512, Cin, 3, 3] -> bf16 [(Cin/64) * 9, 512, 64]: one swizzled K-major block per (channel block, tap)."""
    if weight.dim() != 4 or weight.shape[0] != N_OUT or tuple(weight.shape[2:]) != (3, 3) or weight.shape[1] % KB:
        raise ValueError(f"weight must be [512, Cin, 3, 3] with Cin % 64 == 0. Got {tuple(weight.shape)}")
    return pack_conv1_weight(weight, 1, weight.shape[1], with_coord=False)


def pack_head_weight(weight: torch.Tensor) -> torch.Tensor:
    """[1, Cin, 3, 3] -> bf16 [9, Cin] (tap-major)."""
    if weight.dim() != 4 or weight.shape[0] != 1 or tuple(weight.shape[2:]) != (3, 3):
        raise ValueError(f"weight must be [1, Cin, 3, 3]. Got {tuple(weight.shape)}")
    return weight.detach()[0].permute(1, 2, 0).reshape(9, weight.shape[1]).to(torch.bfloat16).contiguous()


def _check_cl_bf16(x: torch.Tensor, who: str):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{who}: x is not a torch.Tensor")
    if not x.is_cuda:
        raise _lib.MvwarpError(f"{who} is CUDA-only (no CPU fallback)")
    if x.dtype != torch.bfloat16 or x.dim() != 4 or x.stride(1) != 1:
        raise ValueError(f"{who} needs a bf16 channels-last [B,C,H,W] tensor")


def conv3x3_cl(x: torch.Tensor, wpack: torch.Tensor, bias: Optional[torch.Tensor] = None, dilation: int = 1,
               relu: bool = False, out_dtype: torch.dtype = torch.bfloat16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``[relu](conv2d(x, W, bias, padding=dilation, dilation=dilation))`` with 512 output channels on tcgen05 + TMA.
    x bf16 channels-last [B,Cin,H,W]; wpack from :func:`pack_conv3x3_weight`.  Returns channels-last [B,512,H,W]; ``out`` may be
    a channel slice of a wider channels-last tensor (16-byte aligned cell rows)."""
    _check_cl_bf16(x, "conv3x3_cl")
    B, Cin, H, W = x.shape
    if wpack.dtype != torch.bfloat16 or tuple(wpack.shape) != ((Cin // KB) * 9, N_OUT, KB) or not wpack.is_contiguous() or Cin % KB:
        raise ValueError(f"wpack must be a contiguous bf16 [{(Cin // KB) * 9}, 512, 64] tensor from pack_conv3x3_weight")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("conv3x3_cl writes fp32 or bf16")
    if out is None:
        out = torch.empty((B, H, W, N_OUT), dtype=out_dtype, device=x.device).permute(0, 3, 1, 2)
    elif tuple(out.shape) != (B, N_OUT, H, W) or out.dtype != out_dtype or out.stride(1) != 1:
        raise ValueError(f"out must be a channel-contiguous {out_dtype} [B,512,H,W] tensor")
    bias_ptr = None
    if bias is not None:
        bias = bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
        bias_ptr = C.c_void_p(bias.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    rc = _lib.load().mvwarp_conv3x3_fwd(C.c_void_p(x.data_ptr()), _lib.i64x(x.stride()), C.c_void_p(wpack.data_ptr()), bias_ptr,
                                        C.c_void_p(out.data_ptr()), _lib.i64x(out.stride()), B, H, W, Cin, int(dilation), int(relu),
                                        _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32, x.device.index, stream)
    _lib.check(rc, "mvwarp_conv3x3_fwd")
    return out


def conv3x3_head(x: torch.Tensor, w9: torch.Tensor, dilation: int = 1, two_pass: bool = True) -> torch.Tensor:
    """``conv2d(x, W[1,Cin,3,3], None, padding=dilation, dilation=dilation)`` -> fp32 [B,1,H,W]; w9 from :func:`pack_head_weight`.
    ``two_pass`` (default; needs Cin % 32 == 0, else the one-pass kernel runs): nine partial dot products per cell on the tensor cores
    into a ``[B*H*W, 9]`` fp32 scratch from torch's allocator, then a 9-term stencil -- every activation is read once."""
    _check_cl_bf16(x, "conv3x3_head")
    B, Cin, H, W = x.shape
    if w9.dtype != torch.bfloat16 or tuple(w9.shape) != (9, Cin) or not w9.is_contiguous():
        raise ValueError(f"w9 must be a contiguous bf16 [9, {Cin}] tensor from pack_head_weight")
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    lib = _lib.load()
    nbytes = int(lib.mvwarp_conv3x3_head_scratch_bytes(B, H, W)) if two_pass and Cin % 32 == 0 else 0
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if nbytes else None
    rc = lib.mvwarp_conv3x3_head_fwd_ex(C.c_void_p(x.data_ptr()), _lib.i64x(x.stride()), C.c_void_p(w9.data_ptr()),
                                        C.c_void_p(out.data_ptr()), _lib.i64x(out.stride()), B, H, W, Cin, int(dilation),
                                        C.c_void_p(scratch.data_ptr()) if scratch is not None else None, nbytes, x.device.index, stream)
    _lib.check(rc, "mvwarp_conv3x3_head_fwd")
    return out


class MapHead:
    """``map_classifier(warp_views(feats))`` of the reference model as three fused kernels (inference).

    Build it from a reference checkpoint (``torch.load`` of the file main.py:121 saves): keys ``map_classifier.0.weight``
    [512, V*C+2, 3, 3], ``.0.bias``, ``.2.weight`` [512, 512, 3, 3], ``.2.bias``, ``.4.weight`` [1, 512, 3, 3]."""

    def __init__(self, state_dict, num_views: int, channels: int = 512, device="cuda:0", prefix: str = "map_classifier."):
        def get(name):
            for k in (prefix + name, "module." + prefix + name):
                if k in state_dict:
                    return state_dict[k].to(device)
            raise KeyError(f"{prefix + name} not in the state dict")
        self.num_views, self.channels = int(num_views), int(channels)
        self.w1, self.b1 = conv1_from_state_dict(state_dict, num_views, channels, prefix + "0.", device)
        self.w2 = pack_conv3x3_weight(get("2.weight"))
        self.b2 = get("2.bias").float().contiguous()
        self.w3 = pack_head_weight(get("4.weight"))

    def __call__(self, img_feats: torch.Tensor, proj_mats: torch.Tensor, grid_shape: Sequence[int],
                 upsample_shape: Optional[Sequence[int]] = None) -> torch.Tensor:
        """img_feats [B,V,C,H,W] bf16 channels-last -> map_result fp32 [B,1,Hg,Wg] (before the reference's final
        same-size F.interpolate, persp_trans_detector.py:82, which is the identity).  With ``upsample_shape`` img_feats is the
        trunk's low-res output and the 512-channel F.interpolate of persp_trans_detector.py:65 is fused into the first kernel."""
        a1 = warp_conv1(img_feats, proj_mats, grid_shape, self.w1, self.b1, relu=True, out_dtype=torch.bfloat16,
                        upsample_shape=upsample_shape)
        a2 = conv3x3_cl(a1, self.w2, self.b2, dilation=2, relu=True, out_dtype=torch.bfloat16)
        return conv3x3_head(a2, self.w3, dilation=4)
