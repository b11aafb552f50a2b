#!/usr/bin/env python
"""north_star (c): fused warp -> conv1 (tcgen05 implicit GEMM) against the unfused pair (warp_views + cuDNN conv2d).
Wildtrack shape, bf16 features, random weights.  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import mvdet_b200 as mv
from mvdet_b200 import conv1
from mvdet_b200.geometry import RIGS, synthetic_rig

def timeit(fn, iters):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    spec = RIGS["wildtrack"]
    pm = synthetic_rig(spec)
    V, C = spec.num_cam, 512
    Hs, Ws = spec.feat_shape; Hg, Wg = spec.grid_shape
    dev = "cuda:0"
    feats = torch.randn(B, V, Hs, Ws, C, device=dev).bfloat16().permute(0, 1, 4, 2, 3)
    weight = (torch.randn(512, V * C + 2, 3, 3, device=dev) * 0.02)
    bias = torch.randn(512, device=dev)
    wpack = conv1.pack_conv1_weight(weight, V, C)
    w_cl = weight.bfloat16().contiguous(memory_format=torch.channels_last)
    b16 = bias.bfloat16()
    flops = 2.0 * B * Hg * Wg * 512 * (V * C + 2) * 9

    def unfused():
        x = mv.warp_views(feats, pm, (Hg, Wg))
        return F.relu(F.conv2d(x, w_cl, b16, padding=1))
    def warp_only():
        return mv.warp_views(feats, pm, (Hg, Wg))
    def fused():
        return conv1.warp_conv1(feats, pm, (Hg, Wg), wpack, bias, relu=True)

    ref = unfused().float()
    out = fused()
    err = float((out - ref).abs().max() / ref.abs().max())     # cuDNN output is bf16: ~4e-3
    t_unf, t_warp, t_fused = timeit(unfused, iters), timeit(warp_only, iters), timeit(fused, iters)
    print(json.dumps({"B": B, "unfused_ms": t_unf, "warp_ms": t_warp, "cudnn_conv_ms": t_unf - t_warp, "fused_ms": t_fused,
                      "fused_tflops": flops / t_fused / 1e9, "cudnn_conv_tflops": flops / (t_unf - t_warp) / 1e9,
                      "max_rel_diff_vs_cudnn_bf16": err}))

if __name__ == "__main__":
    main()
