"""CPU oracle for MVDet's perspective-transformation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or the
reported CPU baseline.  Nothing under ``mvdet_b200/`` imports it.

PARITY UNPINNED: the reference has no test, golden vector or fixture for this path
 and its arithmetic lives in the third-party dependency
``kornia==0.6.11`` (/root/reference/requirements.txt:1), which is neither vendored in
/root/reference nor installed in this image.  This file restates the published
algorithm of kornia 0.6.11's ``warp_perspective`` (function by function, same torch op
sequence, all on CPU in ``src.dtype``) on top of the *real*
``torch.nn.functional.grid_sample``; parity is anchored on the reference's own call
site /root/reference/multiview_detector/models/persp_trans_detector.py:62-77.
Independent evidence that the restatement is right lives in
tests/test_oracle.py (agreement with cv2.warpPerspective, float64 pixel-space
definition, identity / translation / linearity / adjoint properties) and in the
plain-C second implementation oracle/warp_ref.c.

The oracle is *defined on CPU*: torch.inverse / torch.bmm differ at ulp level between
CPU and CUDA, and a 1-ulp coordinate change moves bilinear weights by ~3e-5.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# kornia 0.6.11 glue, restated (kornia/geometry/conversions.py, linalg.py,
# transform/imgwarp.py, utils/grid.py, utils/helpers.py)
# --------------------------------------------------------------------------------------
def normal_transform_pixel(height: int, width: int, eps: float = 1e-14,
                           device=None, dtype=None) -> torch.Tensor:
    """kornia.geometry.conversions.normal_transform_pixel: pixel -> [-1, 1] (1x3x3)."""
    tr_mat = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]],
                          device=device, dtype=dtype)
    width_denom = eps if width == 1 else width - 1.0
    height_denom = eps if height == 1 else height - 1.0
    tr_mat[0, 0] = tr_mat[0, 0] * 2.0 / width_denom
    tr_mat[1, 1] = tr_mat[1, 1] * 2.0 / height_denom
    return tr_mat.unsqueeze(0)


def _torch_inverse_cast(x: torch.Tensor) -> torch.Tensor:
    """kornia.utils.helpers._torch_inverse_cast: invert in fp32/fp64, cast back."""
    dtype = x.dtype
    if dtype not in (torch.float32, torch.float64):
        dtype = torch.float32
    return torch.inverse(x.to(dtype)).to(x.dtype)


def normalize_homography(dst_pix_trans_src_pix: torch.Tensor,
                         dsize_src: Tuple[int, int], dsize_dst: Tuple[int, int]) -> torch.Tensor:
    """kornia.geometry.conversions.normalize_homography (note the parenthesisation)."""
    src_h, src_w = dsize_src
    dst_h, dst_w = dsize_dst
    src_norm_trans_src_pix = normal_transform_pixel(src_h, src_w).to(dst_pix_trans_src_pix)
    src_pix_trans_src_norm = _torch_inverse_cast(src_norm_trans_src_pix)
    dst_norm_trans_dst_pix = normal_transform_pixel(dst_h, dst_w).to(dst_pix_trans_src_pix)
    return dst_norm_trans_dst_pix @ (dst_pix_trans_src_pix @ src_pix_trans_src_norm)


def create_meshgrid(height: int, width: int, normalized_coordinates: bool = True,
                    device=None, dtype=torch.float32) -> torch.Tensor:
    """kornia.utils.grid.create_meshgrid -> 1xHxWx2, last dim (x, y)."""
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    base_grid = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)  # WxHx2
    return base_grid.permute(1, 0, 2).unsqueeze(0)


def convert_points_to_homogeneous(points: torch.Tensor) -> torch.Tensor:
    return F.pad(points, [0, 1], "constant", 1.0)


def convert_points_from_homogeneous(points: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    z_vec = points[..., -1:]
    mask = torch.abs(z_vec) > eps
    scale = torch.where(mask, 1.0 / (z_vec + eps), torch.ones_like(z_vec))
    return scale * points[..., :-1]


def transform_points(trans_01: torch.Tensor, points_1: torch.Tensor) -> torch.Tensor:
    """kornia.geometry.linalg.transform_points (bmm with the transposed matrix)."""
    shape_inp = list(points_1.shape)
    points_1 = points_1.reshape(-1, points_1.shape[-2], points_1.shape[-1])
    trans_01 = trans_01.reshape(-1, trans_01.shape[-2], trans_01.shape[-1])
    trans_01 = torch.repeat_interleave(trans_01, repeats=points_1.shape[0] // trans_01.shape[0], dim=0)
    points_1_h = convert_points_to_homogeneous(points_1)
    points_0_h = torch.bmm(points_1_h, trans_01.permute(0, 2, 1))
    points_0_h = torch.squeeze(points_0_h, dim=-1)
    points_0 = convert_points_from_homogeneous(points_0_h)
    shape_inp[-2] = points_0.shape[-2]
    shape_inp[-1] = points_0.shape[-1]
    return points_0.reshape(shape_inp)


def sampling_grid(M: torch.Tensor, src_hw: Tuple[int, int], dsize: Sequence[int]) -> torch.Tensor:
    """Steps 1-6 of SURVEY.md §8c: the normalised sampling grid [B,Hg,Wg,2] kornia hands to
    grid_sample for homography M [B,3,3] (source pixel -> destination pixel)."""
    B = M.shape[0]
    H, W = src_hw
    h_out, w_out = dsize
    dst_norm_trans_src_norm = normalize_homography(M, (H, W), (h_out, w_out))
    src_norm_trans_dst_norm = _torch_inverse_cast(dst_norm_trans_src_norm)
    grid = create_meshgrid(h_out, w_out, normalized_coordinates=True, device=M.device).to(M.dtype)
    grid = grid.repeat(B, 1, 1, 1)
    return transform_points(src_norm_trans_dst_norm[:, None, None], grid)


def warp_perspective(src: torch.Tensor, M: torch.Tensor, dsize: Sequence[int],
                     mode: str = "bilinear", padding_mode: str = "zeros",
                     align_corners: bool = True) -> torch.Tensor:
    """kornia.geometry.transform.warp_perspective (0.6.11), as called at
    /root/reference/multiview_detector/models/persp_trans_detector.py:69."""
    if not isinstance(src, torch.Tensor):
        raise TypeError(f"Input src type is not a torch.Tensor. Got {type(src)}")
    if not isinstance(M, torch.Tensor):
        raise TypeError(f"Input M type is not a torch.Tensor. Got {type(M)}")
    if not len(src.shape) == 4:
        raise ValueError(f"Input src must be a BxCxHxW tensor. Got {src.shape}")
    if not (len(M.shape) == 3 and M.shape[-2:] == (3, 3)):
        raise ValueError(f"Input M must be a Bx3x3 tensor. Got {M.shape}")
    _, _, H, W = src.size()
    grid = sampling_grid(M, (H, W), dsize)
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode=padding_mode)


# --------------------------------------------------------------------------------------
# The reference call site, restated
# --------------------------------------------------------------------------------------
def create_coord_map(img_size: Sequence[int]) -> torch.Tensor:
    """/root/reference/multiview_detector/models/persp_trans_detector.py:103-112 (with_r=False)."""
    H, W, _ = img_size
    grid_x, grid_y = np.meshgrid(np.arange(W), np.arange(H))
    grid_x = torch.from_numpy(grid_x / (W - 1) * 2 - 1).float()
    grid_y = torch.from_numpy(grid_y / (H - 1) * 2 - 1).float()
    return torch.stack([grid_x, grid_y], dim=0).unsqueeze(0)


def warp_views_oracle(img_feats: torch.Tensor, proj_mats: torch.Tensor,
                      grid_shape: Sequence[int]) -> torch.Tensor:
    """What persp_trans_detector.py:62-77 computes from the per-camera feature maps:
    per camera ``warp_perspective(feat[:, cam], proj_mats[cam].repeat(B).float(), grid_shape)``
    (:68-69), then ``torch.cat(views + [coord_map.repeat(B)], dim=1)`` (:77).

    img_feats [B,V,C,H,W] (CPU), proj_mats [V,3,3] (float64 as in the reference, or float32),
    returns [B, V*C+2, Hg, Wg] in img_feats.dtype.  Differentiable w.r.t. img_feats.
    """
    B, V = img_feats.shape[:2]
    grid_shape = [int(g) for g in grid_shape]
    coord_map = create_coord_map(grid_shape + [1])
    world_features = []
    for cam in range(V):
        proj_mat = proj_mats[cam].repeat([B, 1, 1]).float()              # :68 (fp64 -> fp32 here)
        if img_feats.dtype != torch.float32:
            proj_mat = proj_mat.to(img_feats.dtype)
        world_features.append(warp_perspective(img_feats[:, cam], proj_mat, grid_shape))  # :69
    return torch.cat(world_features + [coord_map.repeat([B, 1, 1, 1]).to(img_feats.dtype)], dim=1)  # :77


def sampling_grids(proj_mats: torch.Tensor, src_hw: Tuple[int, int],
                   grid_shape: Sequence[int]) -> torch.Tensor:
    """[V,Hg,Wg,2] fp32 grids for a rig, one oracle call per camera exactly as the call site does."""
    return torch.cat([sampling_grid(proj_mats[v].repeat([1, 1, 1]).float(), src_hw, grid_shape)
                      for v in range(proj_mats.shape[0])], dim=0)
