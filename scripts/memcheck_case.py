#!/usr/bin/env python
"""Small forward/backward cases for `compute-sanitizer --tool memcheck` (one tool per gpurun call): exercises the
channels-last kernels (all alignment paths, heavy rows, channel tails), the generic kernels and a fused up-sample plan."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvdet_b200 as mv
from mvdet_b200.geometry import RigSpec, synthetic_rig

spec = RigSpec("mc", 3, img_reduce=20, grid_reduce=20, seed=7)        # 54x96 -> 24x72
pm = synthetic_rig(spec)
for dtype in (torch.float32, torch.bfloat16):
    for C, layout in ((512, "cl"), (132, "cl"), (20, "cl"), (5, "nchw"), (64, "nchw")):
        x = torch.randn(2, 3, C, *spec.feat_shape).to(dtype)
        if layout == "cl":
            x = x.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        x = x.cuda().requires_grad_(True)
        out = mv.warp_views(x, pm, spec.grid_shape)
        (g,) = torch.autograd.grad(out, x, torch.randn_like(out))
        xl = torch.randn(1, 3, C, 18, 32).to(dtype)
        if layout == "cl":
            xl = xl.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        xl = xl.cuda().requires_grad_(True)
        o2 = mv.warp_views_upsampled(xl, spec.feat_shape, pm, spec.grid_shape)
        (g2,) = torch.autograd.grad(o2, xl, torch.randn_like(o2))
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all() and torch.isfinite(g.float()).all() and torch.isfinite(g2.float()).all()
print("memcheck cases done")
