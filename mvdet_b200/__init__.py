"""mvdet_b200 -- B200-native perspective-warp hot path of MVDet (`warp_views`).

Only what the path needs: the CUDA library + C ABI (csrc/, include/mvwarp.h), the rig-constant plan
(plan.py), the PyTorch op and kornia-signature seam (ops.py), the homography setup mirrored from the
reference (geometry.py), the reference-side wiring shipped as code (integration.py: install_kornia_shim,
patch_detector) and view sharding across GPUs (sharding.py).
"""
from .geometry import (MULTIVIEWX, RIG32, RIGS, WILDTRACK, RigSpec, create_coord_map, imgcoord2worldgrid_matrices,
                       projection_matrices, synthetic_rig)
from ._lib import MvwarpError


def __getattr__(name):  # ops import torch.library machinery lazily
    if name in ("warp_views", "warp_views_with_plan", "warp_views_upsampled", "warp_perspective", "grid_sample_views",
                "warp_views_fwd_into", "warp_views_bwd_into", "warp_views_list"):
        from . import ops
        return getattr(ops, name)
    if name in ("get_plan", "get_plan_from_grid", "get_plan_upsampled", "load_plan", "WarpPlan", "src_norm_trans_dst_norm", "host_grid", "host_csr"):
        from . import plan
        return getattr(plan, name)
    if name in ("install_kornia_shim", "patch_detector", "warped_forward"):
        from . import integration
        return getattr(integration, name)
    raise AttributeError(name)
