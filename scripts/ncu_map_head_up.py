#!/usr/bin/env python
"""Workload for the ncu capture of the inference head over LOW-RES features: MapHead(upsample_shape=) = warp_conv1_direct_kernel<UPS>
(+ conv1_tail_reduce_kernel) -> conv3x3_tc2_kernel -> head_partial_kernel -> head_stencil_kernel, Wildtrack shape.
usage: ncu_map_head_up.py [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from mvdet_b200 import map_head as MH
from mvdet_b200.geometry import RIGS, synthetic_rig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = RIGS["wildtrack"]
pm = synthetic_rig(spec)
V, C = spec.num_cam, 512
Hs, Ws = spec.feat_shape
dev = "cuda:0"
torch.manual_seed(0)
low = torch.randn(B, V, Hs // 3, Ws // 3, C, device=dev).bfloat16().permute(0, 1, 4, 2, 3)
ref_head = nn.Sequential(nn.Conv2d(C * V + 2, 512, 3, padding=1), nn.ReLU(), nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                         nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False)).to(dev)
head = MH.MapHead({"map_classifier." + k: v for k, v in ref_head.state_dict().items()}, V, C, device=dev)
with torch.no_grad():
    for _ in range(iters):
        out = head(low, pm, spec.grid_shape, upsample_shape=spec.feat_shape)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    for _ in range(iters):
        out = head(low, pm, spec.grid_shape, upsample_shape=spec.feat_shape)
e1.record()
torch.cuda.synchronize()
print(f"map head from low-res features, B={B}: {e0.elapsed_time(e1) / iters:.4f} ms per frame batch; out {tuple(out.shape)} finite {bool(torch.isfinite(out).all())}")
