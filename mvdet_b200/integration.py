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


def _device_of(module: torch.nn.Module) -> torch.device:
    return next(module.parameters()).device


def warped_forward(model, imgs: torch.Tensor, visualize: bool = False, warp_fn: Optional[Callable] = None):
    """``PerspTransDetector.forward`` (persp_trans_detector.py:57-87) with lines 68-69, 75, 77 replaced by one ``warp_views`` call.

    Returns ``(map_result, imgs_result)`` like the reference.  Sub-modules run where the reference put them (each on the device of
    its own parameters; the reference pins all of them to ``cuda:0``); the warp and ``map_classifier`` run on ``map_classifier``'s
    device.  ``visualize`` is the reference's matplotlib debugging switch and is not supported here."""
    if visualize:
        raise NotImplementedError("visualize=True is the reference's matplotlib debugging path; call the original forward for it")
    if warp_fn is None:
        from .ops import warp_views as warp_fn
    B, N = imgs.shape[:2]
    if N != model.num_cam:
        raise AssertionError(f"imgs carries {N} views, the model has {model.num_cam}")          # the reference asserts (:59)
    d1, d2 = _device_of(model.base_pt1), _device_of(model.base_pt2)
    dh, dm = _device_of(model.img_classifier), _device_of(model.map_classifier)
    feats, imgs_result = [], []
    for cam in range(model.num_cam):                                                            # :62
        f = model.base_pt1(imgs[:, cam].to(d1))                                                 # :63
        f = model.base_pt2(f.to(d2))                                                            # :64
        f = F.interpolate(f, model.upsample_shape, mode="bilinear")                             # :65
        imgs_result.append(model.img_classifier(f.to(dh)))                                      # :66-67
        feats.append(f.to(dm))
    img_feats = torch.stack(feats, dim=1)                                                       # [B,V,C,Hs,Ws]; keeps channels-last
    proj_mats = torch.stack([m for m in model.proj_mats])                                       # float64 CPU (:29-30); op applies .float() (:68)
    world_features = warp_fn(img_feats, proj_mats, model.reducedgrid_shape)                     # :68-69, 75, 77
    map_result = model.map_classifier(world_features)                                           # :81
    map_result = F.interpolate(map_result, model.reducedgrid_shape, mode="bilinear")            # :82
    return map_result, imgs_result


def patch_detector(model, channels_last: bool = True, warp_fn: Optional[Callable] = None):
    """Give ``model`` (a ``PerspTransDetector``) the :func:`warped_forward`; returns the same object.

    ``channels_last=True`` also converts the model's parameters to ``torch.channels_last`` so that the trunk, ``F.interpolate`` and
    ``torch.stack`` hand the warp channel-contiguous features (the fast kernels; NCHW is accepted but ~3x slower) and cuDNN
    consumes the channels-last result directly.  Feed ``imgs.contiguous(memory_format=torch.channels_last)`` per camera for the
    full effect.  The original forward stays reachable as ``model.reference_forward``."""
    missing = [a for a in _ATTRS if not hasattr(model, a)]
    if missing:
        raise TypeError(f"patch_detector needs a PerspTransDetector-like model; missing attributes: {missing}")
    if channels_last:
        model.to(memory_format=torch.channels_last)
    if not hasattr(model, "reference_forward"):
        model.reference_forward = model.forward
    model.forward = types.MethodType(lambda self, imgs, visualize=False: warped_forward(self, imgs, visualize, warp_fn), model)
    return model
