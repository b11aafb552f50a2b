"""Reference-side wiring, shipped as code (INTEGRATION.md shows the same seams as snippets).

* :func:`install_kornia_shim` -- seam 1, zero edits to the reference: ``kornia.geometry.transform.warp_perspective`` and the alias
  ``kornia.warp_perspective`` (persp_trans_detector.py:69, image_proj_variant.py:69, res_proj_variant.py:72,
  no_joint_conv_variant.py:68) become :func:`mvdet_b200.ops.warp_perspective`.  Works with or without kornia installed.
* :func:`patch_detector` -- seam 2: a ``PerspTransDetector`` instance (or anything with the same attributes) gets a ``forward``
  that keeps the reference's trunk / heads (persp_trans_detector.py:63-66, 81-82) and replaces lines 68-69, 75 and 77 -- the
  per-camera matrix staging, ``warp_perspective`` x V, the list append, ``torch.cat`` with ``coord_map`` -- by ONE ``warp_views`` call.

Nothing here computes: the warp is :func:`mvdet_b200.warp_views` (CUDA only, no fallback).  ``warp_fn`` exists so that the CPU
tests can stand the oracle in for the CUDA op and compare against the unmodified reference forward.
"""
from __future__ import annotations

import sys
import types
from typing import Callable, Optional

import torch
import torch.nn.functional as F

_ATTRS = ("num_cam", "base_pt1", "base_pt2", "img_classifier", "map_classifier", "proj_mats", "upsample_shape",
          "reducedgrid_shape")


def install_kornia_shim(warp: Optional[Callable] = None) -> types.ModuleType:
    """Point ``kornia.geometry.transform.warp_perspective`` and ``kornia.warp_perspective`` at the B200 op.

    If kornia is importable its functions are patched in place; otherwise stub modules ``kornia``, ``kornia.geometry`` and
    ``kornia.geometry.transform`` are registered in ``sys.modules`` so that the reference's ``import kornia``
    (persp_trans_detector.py:7) succeeds.  Returns the ``kornia`` module object.  Call it before importing the reference models."""
    if warp is None:
        from .ops import warp_perspective as warp
    try:
        import kornia                                   # noqa: F401  (the real package, if present)
        import kornia.geometry.transform as kgt
    except ImportError:
        kornia = sys.modules.get("kornia") or types.ModuleType("kornia")
        geometry = getattr(kornia, "geometry", None) or types.ModuleType("kornia.geometry")
        kgt = getattr(geometry, "transform", None) or types.ModuleType("kornia.geometry.transform")
        kornia.geometry, geometry.transform = geometry, kgt
        sys.modules.update({"kornia": kornia, "kornia.geometry": geometry, "kornia.geometry.transform": kgt})
    kgt.warp_perspective = warp
    kornia.warp_perspective = warp
    return kornia


class FusedHeadCache:
    """The packed-weight :class:`mvdet_b200.map_head.MapHead` of a model's ``map_classifier``, re-packed lazily whenever the weights
    change.  The reference trainer alternates ``train()`` with ``test()`` under ``torch.no_grad()`` (trainer.py:87-95) and ``main.py:125``
    calls ``load_state_dict`` after the model is built, so a snapshot taken at patch time would silently go stale after the first
    optimizer step or checkpoint load.  The cache is keyed on ``(data_ptr, _version, device)`` of every parameter of ``map_classifier``:
    in-place updates (optimizer steps, ``load_state_dict``'s ``copy_``) bump ``_version``; ``.to(device)`` / re-assignment changes
    ``data_ptr``.  ``factory(state_dict)`` builds the head (tests inject a stand-in)."""

    def __init__(self, model, factory: Optional[Callable] = None):
        self.model, self.factory = model, factory
        self._key, self._head = None, None
        self.rebuilds = 0

    def _signature(self):
        return tuple((name, p.data_ptr(), p._version, str(p.device)) for name, p in self.model.map_classifier.state_dict(keep_vars=True).items())

    def get(self):
        key = self._signature()
        if key != self._key:
            mc = self.model.map_classifier
            sd = {"map_classifier." + k: v.detach() for k, v in mc.state_dict().items()}
            if self.factory is not None:
                self._head = self.factory(sd)
            else:
                from .map_head import MapHead
                conv = mc[0]
                self._head = MapHead(sd, self.model.num_cam, (conv.in_channels - 2) // self.model.num_cam, device=conv.weight.device)
            self._key = key
            self.rebuilds += 1
        return self._head

    def __call__(self, *args, **kwargs):
        return self.get()(*args, **kwargs)


def _onto(t: torch.Tensor, module: torch.nn.Module) -> torch.Tensor:
    """``t`` on the device of ``module``'s parameters -- the reference's ``.to('cuda:0')`` hops (:63-66, 81) without the hard-coded name."""
    p = next(module.parameters(), None)
    return t if p is None else t.to(p.device)


def _split_img_classifier(img_classifier):
    """(first 1x1 convolution, the rest) of the reference's img_classifier = Conv2d(512,64,1), ReLU, Conv2d(64,2,1) (:48-49)."""
    first = img_classifier[0] if isinstance(img_classifier, torch.nn.Sequential) and len(img_classifier) > 0 else None
    ok = (isinstance(first, torch.nn.Conv2d) and first.kernel_size == (1, 1) and first.stride == (1, 1) and
          first.padding == (0, 0) and first.groups == 1)
    if not ok:
        raise TypeError("fuse_upsample needs img_classifier to start with a 1x1 Conv2d (the reference's does, "
                        "persp_trans_detector.py:48): a 1x1 convolution commutes with the bilinear up-sample")
    return first, img_classifier[1:]


def warped_forward(model, imgs: torch.Tensor, visualize: bool = False, warp_fn: Optional[Callable] = None,
                   fuse_upsample: bool = False, fused_head=None, fused_conv1: bool = False):
    """``PerspTransDetector.forward`` (persp_trans_detector.py:57-87) with lines 68-69, 75, 77 replaced by one ``warp_views`` call.

    Returns ``(map_result, imgs_result)`` like the reference.  Sub-modules run where the reference put them (each on the device of
    its own parameters; the reference pins all of them to ``cuda:0``); the warp and ``map_classifier`` run on ``map_classifier``'s
    device.  ``visualize`` is the reference's matplotlib debugging switch and is not supported here.

    ``fuse_upsample=True`` additionally removes the 512-channel ``F.interpolate`` of line 65 (a 1.86 GB tensor per Wildtrack frame that
    the reference writes, reads twice and back-propagates through): the warp reads the trunk's low-res output through
    :func:`mvdet_b200.ops.warp_views_upsampled` (up-sample composed into the plan), and ``img_classifier`` up-samples AFTER its first
    1x1 convolution -- 64 channels instead of 512; exact in real arithmetic because a 1x1 convolution (bias included: bilinear
    weights sum to one) commutes with the per-channel linear interpolation.  Results agree with the reference order to rounding
    (1e-5 relative in the tests), not bit for bit.  ``warp_fn`` then has the signature of ``warp_views_upsampled``.

    ``fused_head`` (a :class:`mvdet_b200.map_head.MapHead`; see :func:`patch_detector`) replaces the warp AND ``map_classifier``
    (:68-81) by three fused kernels whenever autograd is off: fused warp -> conv1 on tcgen05, dilated conv by TMA, 512 -> 1 head,
    bf16 activations with fp32 accumulation.  With autograd on the un-fused path runs (the fused kernels are forward-only) --
    unless ``fused_conv1`` is set: then ``map_classifier[0]`` + ReLU run as :func:`mvdet_b200.conv1.warp_conv1_trainable` (fused
    tcgen05 forward on bf16 features; backward = own ReLU-mask / data-gradient / weight-gradient kernels + the deterministic CSR
    warp backward, no cuDNN; BASELINE.json configs[2]) and the rest of ``map_classifier`` follows in torch.  With ``fuse_upsample``
    both fused paths read the low-res features (``upsample_shape=`` of ``warp_conv1``): line 65 disappears for the map branch."""
    if visualize:
        raise NotImplementedError("visualize=True is the reference's matplotlib debugging path; call the original forward for it")
    list_fn = None
    if warp_fn is None:
        # the CUDA op takes the per-camera tensors as they are (mvwarp_warp_views_fwd_list): no torch.stack, i.e. no 1.86 GB copy per
        # Wildtrack frame in the forward and none in the backward.  A caller-supplied warp_fn (the CPU tests' oracle) keeps the
        # stacked-tensor signature.
        from .ops import warp_views_list as list_fn
        if fuse_upsample:
            from .ops import warp_views_upsampled as warp_fn
        else:
            from .ops import warp_views as warp_fn
    B, N = imgs.shape[:2]
    if N != model.num_cam:
        raise AssertionError(f"imgs carries {N} views, the model has {model.num_cam}")          # the reference asserts (:59)
    head_first, head_rest = _split_img_classifier(model.img_classifier) if fuse_upsample else (None, None)
    feats, imgs_result = [], []
    for cam in range(model.num_cam):                                                            # :62
        f = model.base_pt1(_onto(imgs[:, cam], model.base_pt1))                                 # :63
        f = model.base_pt2(_onto(f, model.base_pt2))                                            # :64
        if fuse_upsample:
            h = F.interpolate(head_first(_onto(f, head_first)), model.upsample_shape, mode="bilinear")   # :65-66 reordered: 1x1 conv, then up-sample
            imgs_result.append(head_rest(h))
        else:
            f = F.interpolate(f, model.upsample_shape, mode="bilinear")                         # :65
            imgs_result.append(model.img_classifier(_onto(f, model.img_classifier)))            # :66-67
        feats.append(_onto(f, model.map_classifier))
    proj_mats = torch.stack([m for m in model.proj_mats])                                       # float64 CPU (:29-30); op applies .float() (:68)
    fused_now = (fused_head is not None and not torch.is_grad_enabled()) or fused_conv1
    if list_fn is not None and not fused_now:
        world_features = list_fn(feats, proj_mats, model.reducedgrid_shape,
                                 upsample_shape=model.upsample_shape if fuse_upsample else None)   # :68-69, 75, 77 (+ :65 when fused)
        map_result = model.map_classifier(world_features)                                       # :81
        return F.interpolate(map_result, model.reducedgrid_shape, mode="bilinear"), imgs_result # :82
    if fused_now and warp_fn is not None and feats[0].is_cuda:
        # the fused tcgen05 kernels read bf16 channel-contiguous [B,V,C,H,W]: convert, transpose (NCHW trunks) and stack in ONE pass
        # per camera instead of torch.stack (fp32 read + write) followed by .to(bf16) (another read + write)
        Bf, Cf, Hf, Wf = feats[0].shape
        x = torch.empty((Bf, len(feats), Hf, Wf, Cf), dtype=torch.bfloat16, device=feats[0].device).permute(0, 1, 4, 2, 3)
        for v, f in enumerate(feats):
            x[:, v].copy_(f)
        img_feats = x
    else:
        img_feats = torch.stack(feats, dim=1)                                                   # [B,V,C,H,W]; keeps channels-last
    if fused_head is not None and not torch.is_grad_enabled():
        x = img_feats.to(torch.bfloat16)
        if x.stride(2) != 1:                                                                    # NCHW trunk: one transposing copy
            x = x.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        up = {"upsample_shape": tuple(model.upsample_shape)} if fuse_upsample else {}            # :65 fused in as well
        map_result = fused_head(x, proj_mats, model.reducedgrid_shape, **up)                    # :68-81
        return F.interpolate(map_result, model.reducedgrid_shape, mode="bilinear"), imgs_result # :82
    if fused_conv1:
        from .conv1 import warp_conv1_trainable
        x = img_feats.to(torch.bfloat16)
        if x.stride(2) != 1:
            x = x.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        conv = model.map_classifier[0]
        a1 = warp_conv1_trainable(x, conv.weight, conv.bias, proj_mats, model.reducedgrid_shape, relu=True,
                                  upsample_shape=tuple(model.upsample_shape) if fuse_upsample else None)   # (:65) :68-77 + map_classifier[0:2]
        map_result = model.map_classifier[2:](a1)                                               # :81, the rest
        return F.interpolate(map_result, model.reducedgrid_shape, mode="bilinear"), imgs_result # :82
    if fuse_upsample:
        world_features = warp_fn(img_feats, model.upsample_shape, proj_mats, model.reducedgrid_shape)   # :65 (for the warp) + :68-69, 75, 77
    else:
        world_features = warp_fn(img_feats, proj_mats, model.reducedgrid_shape)                 # :68-69, 75, 77
    map_result = model.map_classifier(world_features)                                           # :81
    map_result = F.interpolate(map_result, model.reducedgrid_shape, mode="bilinear")            # :82
    return map_result, imgs_result


def patch_detector(model, channels_last: bool = True, warp_fn: Optional[Callable] = None, fuse_upsample: bool = False,
                   fused_head: bool = False, fused_conv1: bool = False):
    """Give ``model`` (a ``PerspTransDetector``) the :func:`warped_forward`; returns the same object.

    ``channels_last=True`` also converts the model's parameters to ``torch.channels_last`` so that the trunk, ``F.interpolate`` and
    ``torch.stack`` hand the warp channel-contiguous features (the fast kernels; NCHW is accepted but ~3x slower) and cuDNN
    consumes the channels-last result directly.  Feed ``imgs.contiguous(memory_format=torch.channels_last)`` per camera for the
    full effect.  ``fuse_upsample``: see :func:`warped_forward`.  ``fused_head=True`` packs ``map_classifier``'s weights into a
    :class:`mvdet_b200.map_head.MapHead` behind a :class:`FusedHeadCache` (stored as ``model.fused_map_head``; it re-packs by itself
    when an optimizer step or ``load_state_dict`` changes the weights) and uses it under ``torch.no_grad()`` -- inference: 1.14 ms per Wildtrack frame for the warp + ``map_classifier`` against 2.01 ms un-fused.
    ``fused_conv1=True`` (training) runs ``map_classifier[0]`` + ReLU as the fused warp -> conv1 kernel with gradients, see
    :func:`warped_forward`.  Both combine with ``fuse_upsample``: the fused kernel then reads the trunk's low-res output and composes
    up-sample and warp per halo cell (``warp_conv1(..., upsample_shape=)``), so no 512-channel ``F.interpolate`` runs at all -- in
    training neither its backward: the CSR warp backward over the composed plan writes the low-res gradient.  The original forward
    stays reachable as ``model.reference_forward``."""
    missing = [a for a in _ATTRS if not hasattr(model, a)]
    if missing:
        raise TypeError(f"patch_detector needs a PerspTransDetector-like model; missing attributes: {missing}")
    if fused_conv1:
        mc = model.map_classifier
        ok = (isinstance(mc, torch.nn.Sequential) and len(mc) >= 2 and isinstance(mc[0], torch.nn.Conv2d) and
              isinstance(mc[1], torch.nn.ReLU) and mc[0].kernel_size == (3, 3) and mc[0].padding == (1, 1) and mc[0].out_channels == 512)
        if not ok:
            raise TypeError("fused_conv1 needs map_classifier to start with Conv2d(V*C+2, 512, 3, padding=1), ReLU (persp_trans_detector.py:51)")
    if fuse_upsample:
        _split_img_classifier(model.img_classifier)
    head = None
    if fused_head:
        mc = model.map_classifier
        conv = mc[0] if isinstance(mc, torch.nn.Sequential) and len(mc) == 5 else None
        if not isinstance(conv, torch.nn.Conv2d) or (conv.in_channels - 2) % model.num_cam:
            raise TypeError("fused_head needs the reference's map_classifier: Conv2d(V*C+2,512,3) ReLU Conv2d(512,512,3,d=2) ReLU Conv2d(512,1,3,d=4)")
        head = FusedHeadCache(model)          # packs lazily and re-packs when the weights change (optimizer step, load_state_dict)
        head.get()
        model.fused_map_head = head
    if channels_last:
        model.to(memory_format=torch.channels_last)
    if not hasattr(model, "reference_forward"):
        model.reference_forward = model.forward
    model.forward = types.MethodType(
        lambda self, imgs, visualize=False: warped_forward(self, imgs, visualize, warp_fn, fuse_upsample, head, fused_conv1), model)
    return model
