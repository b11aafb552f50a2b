"""Homography setup for the perspective-transformation path.

Host-side float64 numpy, run once per camera rig -- exactly where the reference does it:
  * :func:`imgcoord2worldgrid_matrices`  <- PerspTransDetector.get_imgcoord2worldgrid_matrices
    (/root/reference/multiview_detector/models/persp_trans_detector.py:89-101)
  * :func:`projection_matrices`          <- proj_mats assembly (persp_trans_detector.py:23-30)
  * :func:`create_coord_map`             <- persp_trans_detector.py:103-112
  * :class:`RigSpec` / :func:`synthetic_rig` -- the real Wildtrack / MultiviewX calibrations are
    not available offline, so benchmarks and tests push synthetic pin-hole cameras through the
    reference's own matrix pipeline (generator specified in SURVEY.md §8d).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch


def imgcoord2worldgrid_matrices(intrinsic_matrices: Sequence[np.ndarray],
                                extrinsic_matrices: Sequence[np.ndarray],
                                worldgrid2worldcoord_mat: np.ndarray) -> List[np.ndarray]:
    """Per camera: drop the world-z column of [R|t], chain with the grid->world matrix,
    invert, and swap the two output rows so that grid (i, j) becomes (x=col, y=row)."""
    swap_rows = np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]])
    out = []
    for K, E in zip(intrinsic_matrices, extrinsic_matrices):
        worldcoord2img = K @ np.delete(E, 2, 1)
        worldgrid2img = worldcoord2img @ worldgrid2worldcoord_mat
        out.append(swap_rows @ np.linalg.inv(worldgrid2img))
    return out


def projection_matrices(intrinsic_matrices, extrinsic_matrices, worldgrid2worldcoord_mat,
                        img_shape: Sequence[int], img_reduce: float, grid_reduce: float) -> torch.Tensor:
    """``proj_mats`` of the reference as one float64 [V,3,3] tensor: feature-map pixel
    (x, y, 1) -> reduced-grid pixel (x, y, w)."""
    upsample_shape = [int(x / img_reduce) for x in img_shape]
    img_zoom = np.diag(np.append(np.array(img_shape) / np.array(upsample_shape), [1]))
    map_zoom = np.diag(np.append(np.ones([2]) / grid_reduce, [1]))
    mats = imgcoord2worldgrid_matrices(intrinsic_matrices, extrinsic_matrices, worldgrid2worldcoord_mat)
    return torch.from_numpy(np.stack([map_zoom @ m @ img_zoom for m in mats]))


def create_coord_map(grid_shape: Sequence[int]) -> torch.Tensor:
    """[1,2,Hg,Wg] fp32: channel 0 = x/(Wg-1)*2-1, channel 1 = y/(Hg-1)*2-1, computed in
    float64 then rounded once to fp32 (so the corners are exactly -1 / +1)."""
    H, W = int(grid_shape[0]), int(grid_shape[1])
    xs = torch.from_numpy(np.arange(W) / (W - 1) * 2 - 1).float()
    ys = torch.from_numpy(np.arange(H) / (H - 1) * 2 - 1).float()
    return torch.stack([xs[None, :].expand(H, W), ys[:, None].expand(H, W)], dim=0).unsqueeze(0).contiguous()


@dataclass
class RigSpec:
    """Shape constants of one dataset configuration (SURVEY.md §2 'shape constants')."""
    name: str
    num_cam: int
    img_shape: tuple = (1080, 1920)          # calibration space, H x W
    worldgrid_shape: tuple = (480, 1440)     # N_row x N_col
    img_reduce: int = 4
    grid_reduce: int = 4
    channels: int = 512
    seed: int = 0

    @property
    def feat_shape(self):        # warp source = upsample_shape
        return tuple(int(x / self.img_reduce) for x in self.img_shape)

    @property
    def grid_shape(self):        # warp target = reducedgrid_shape
        return tuple(int(x / self.grid_reduce) for x in self.worldgrid_shape)


WILDTRACK = RigSpec("wildtrack", 7, worldgrid_shape=(480, 1440), seed=0)          # Wildtrack.py:20-21
MULTIVIEWX = RigSpec("multiviewx", 6, worldgrid_shape=(640, 1000), seed=1)        # MultiviewX.py:20-21
RIG32 = RigSpec("rig32", 32, worldgrid_shape=(480, 1440), grid_reduce=1, seed=2)  # BASELINE.json configs[3]
RIGS = {r.name: r for r in (WILDTRACK, MULTIVIEWX, RIG32)}


def synthetic_cameras(num_cam: int, seed: int):
    """Pin-hole cameras on an ellipse around the Wildtrack ground rectangle, looking inward.
    Draw order per camera: angle jitter, height, target x, target y, focal, cx, cy."""
    rng = np.random.default_rng(seed)
    Ks, Es = [], []
    for c in range(num_cam):
        ang = 2 * np.pi * c / num_cam + rng.uniform(-0.2, 0.2)
        height = rng.uniform(220, 420)
        pos = np.array([300 + 1100 * np.cos(ang), 900 + 2300 * np.sin(ang), height])
        target = np.array([300 + rng.uniform(-150, 150), 900 + rng.uniform(-400, 400), 0.0])
        focal = rng.uniform(1100, 1800)
        cx = 960 + rng.uniform(-30, 30)
        cy = 540 + rng.uniform(-60, 60)
        fwd = target - pos
        fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd])
        t = -R @ pos
        Ks.append(np.array([[focal, 0, cx], [0, focal, cy], [0, 0, 1.0]]))
        Es.append(np.concatenate([R, t[:, None]], axis=1))
    return Ks, Es


def synthetic_rig(spec: RigSpec) -> torch.Tensor:
    """float64 [V,3,3] proj_mats for a synthetic rig of ``spec`` (cm units, Wildtrack.py:25)."""
    w2w = np.array([[2.5, 0, -300], [0, 2.5, -900], [0, 0, 1.0]])
    Ks, Es = synthetic_cameras(spec.num_cam, spec.seed)
    return projection_matrices(Ks, Es, w2w, spec.img_shape, spec.img_reduce, spec.grid_reduce)
