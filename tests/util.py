"""Shared helpers for the parity tests."""
import torch

from mvdet_b200.geometry import RigSpec, synthetic_rig


def small_rig(num_cam=3, seed=5, src=(54, 96), grid=(24, 72)) -> tuple:
    """A scaled-down Wildtrack-like rig: same generator, 1/5 resolution.  Returns (proj_mats, src_hw, grid_hw)."""
    spec = RigSpec("small", num_cam, img_shape=(1080, 1920), worldgrid_shape=(480, 1440),
                   img_reduce=1080 // src[0], grid_reduce=480 // grid[0], seed=seed)
    assert spec.feat_shape == tuple(src) and spec.grid_shape == tuple(grid)
    return synthetic_rig(spec), spec.feat_shape, spec.grid_shape


def rel_errors(a: torch.Tensor, o: torch.Tensor):
    """(max-relative, l2-relative) error of a against oracle o (SURVEY.md 8d definitions)."""
    a = a.detach().double().cpu()
    o = o.detach().double().cpu()
    denom_max = o.abs().max().clamp_min(1e-30)
    denom_l2 = o.norm().clamp_min(1e-30)
    return float((a - o).abs().max() / denom_max), float((a - o).norm() / denom_l2)


def bit_equal_fraction(a: torch.Tensor, o: torch.Tensor) -> float:
    a = a.cpu().contiguous()
    o = o.cpu().contiguous()
    return float((a == o).double().mean())


def to_channels_last_5d(x: torch.Tensor) -> torch.Tensor:
    """[B,V,C,H,W] logical, channel-contiguous physical."""
    return x.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
