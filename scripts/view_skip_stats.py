#!/usr/bin/env python
"""How much of the fused conv1's K work the view skipping removes (CPU only): for every pair of consecutive 16 x 8 cell tiles and every
view, is there a halo cell (18 x 10 per tile) of either tile with at least one in-range bilinear tap?  If not, the kernel drops the
view's 9 * C/64 K blocks for that pair (mvwarp_conv1.cu, warp_conv1_direct_kernel).  Prints the skippable share per rig."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvdet_b200 import plan as P
from mvdet_b200.geometry import RIGS, synthetic_rig

for spec in RIGS.values():
    pm = synthetic_rig(spec)
    (Hs, Ws), (Hg, Wg) = spec.feat_shape, spec.grid_shape
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, (Hs, Ws), (Hg, Wg)), (Hg, Wg))
    x0 = torch.floor((g[..., 0] + 1) * ((Ws - 1) / 2))
    y0 = torch.floor((g[..., 1] + 1) * ((Hs - 1) / 2))
    valid = (x0 >= -1) & (x0 <= Ws - 1) & (y0 >= -1) & (y0 <= Hs - 1)          # some tap of the cell is inside the view's image
    V, tx, ty = valid.shape[0], (Wg + 15) // 16, (Hg + 7) // 8
    seen = torch.zeros(V, ty, tx, dtype=torch.bool)
    for j in range(ty):
        for i in range(tx):
            ya, yb, xa, xb = max(0, j * 8 - 1), min(Hg, j * 8 + 9), max(0, i * 16 - 1), min(Wg, i * 16 + 17)
            seen[:, j, i] = valid[:, ya:yb, xa:xb].reshape(V, -1).any(1)
    flat = seen.reshape(V, -1)
    m = flat.shape[1] // 2 * 2
    pair_unseen = (~flat[:, 0:m:2]) & (~flat[:, 1:m:2])
    print(f"{spec.name}: valid cells {valid.float().mean():.3f}, tiles {flat.shape[1]}, (tile, view) unseen {(~flat).float().mean():.3f}, "
          f"(tile pair, view) unseen = skipped K share {pair_unseen.float().mean():.3f}")
