#!/usr/bin/env python
"""BASELINE.json configs[2]: Wildtrack training step, bf16 -- warp forward + deterministic CSR backward around map_classifier[0].
Times forward + backward (gradients w.r.t. the features, conv1's weight and bias) for
  fused:   mvdet_b200.conv1.warp_conv1_trainable (fused tcgen05 forward; backward = warp_views + cuDNN conv backward + CSR warp backward)
  unfused: mvdet_b200.warp_views + F.conv2d + relu, autograd throughout (the drop-in op alone)
Prints one JSON line.   usage: train_step_bench.py [B] [iters]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import mvdet_b200 as mv
from mvdet_b200 import conv1
from mvdet_b200.geometry import RIGS, synthetic_rig


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    spec = RIGS["wildtrack"]
    pm = synthetic_rig(spec)
    V, C = spec.num_cam, 512
    Hs, Ws = spec.feat_shape; Hg, Wg = spec.grid_shape
    dev = "cuda:0"
    torch.manual_seed(0)
    feats = torch.randn(B, V, Hs, Ws, C, device=dev).bfloat16().permute(0, 1, 4, 2, 3).requires_grad_(True)
    weight = (torch.randn(512, V * C + 2, 3, 3, device=dev) * 0.02).requires_grad_(True)
    bias = torch.randn(512, device=dev).requires_grad_(True)
    go = torch.randn(B, Hg, Wg, 512, device=dev).permute(0, 3, 1, 2)
    wpack = conv1.pack_conv1_weight(weight, V, C)
    w16 = weight.detach().bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    b16 = bias.detach().bfloat16().requires_grad_(True)
    go16 = go.bfloat16()

    def fused():
        out = conv1.warp_conv1_trainable(feats, weight, bias, pm, (Hg, Wg), wpack=wpack)
        return torch.autograd.grad(out, [feats, weight, bias], go)

    def unfused():
        x = mv.warp_views(feats, pm, (Hg, Wg))
        out = F.relu(F.conv2d(x, w16, b16, padding=1))
        return torch.autograd.grad(out, [feats, w16, b16], go16)

    def timeit(fn):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    gf, gw, gb = fused()
    uf, uw, ub = unfused()
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    res = {"B": B, "fused_fwd_bwd_ms": timeit(fused), "unfused_fwd_bwd_ms": timeit(unfused),
           "grad_feats_l2_diff": rel(gf, uf), "grad_weight_l2_diff": rel(gw, uw), "grad_bias_l2_diff": rel(gb, ub)}
    res["frames_per_s_fused"] = B / res["fused_fwd_bwd_ms"] * 1e3
    res["frames_per_s_unfused"] = B / res["unfused_fwd_bwd_ms"] * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
