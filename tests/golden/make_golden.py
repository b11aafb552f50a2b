#!/usr/bin/env python
"""Generates tests/golden/warp_small.npz from the CPU oracle (oracle/kornia_warp.py).

The reference itself cannot produce golden vectors here: its arithmetic is kornia==0.6.11, which is not
installed, and it ships no fixtures of its own.  These vectors therefore freeze the ORACLE's
output (torch 2.11.0 CPU) so that drift of the oracle, of torch, or of the host CPU's kernels is noticed.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mvdet_b200.geometry import RigSpec, synthetic_rig  # noqa: E402
from oracle import kornia_warp as K  # noqa: E402


def main():
    spec = RigSpec("golden", 3, img_reduce=40, grid_reduce=40, seed=13)   # 27x48 features -> 12x36 grid
    pm = synthetic_rig(spec)
    B, V, C = 2, spec.num_cam, 5
    x = torch.randn(B, V, C, *spec.feat_shape, generator=torch.Generator().manual_seed(0))
    xr = x.clone().requires_grad_(True)
    out = K.warp_views_oracle(xr, pm, spec.grid_shape)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (gx,) = torch.autograd.grad(out, xr, go)
    grid = K.sampling_grids(pm, spec.feat_shape, spec.grid_shape)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_small.npz")
    np.savez_compressed(path, proj_mats=pm.numpy(), feats=x.numpy(), out=out.detach().numpy(), grad_out=go.numpy(),
                        grad_feats=gx.numpy(), grid=grid.numpy(), feat_shape=np.array(spec.feat_shape),
                        grid_shape=np.array(spec.grid_shape), torch_version=np.array(torch.__version__))
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
