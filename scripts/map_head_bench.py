#!/usr/bin/env python
"""the whole ground-plane head, map_classifier(warp_views(feats)) (persp_trans_detector.py:51-54,
:68-81), as MapHead's three kernels against the unfused pipeline (warp_views + three cuDNN convolutions, bf16
channels-last).  Wildtrack shape, random weights.  Prints one JSON line.   usage: map_head_bench.py [B] [iters]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
import mvdet_b200 as mv
from mvdet_b200 import map_head as MH
from mvdet_b200.conv1 import warp_conv1
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
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    spec = RIGS["wildtrack"]
    pm = synthetic_rig(spec)
    V, C = spec.num_cam, 512
    Hs, Ws = spec.feat_shape; Hg, Wg = spec.grid_shape
    dev = "cuda:0"
    torch.manual_seed(0)
    feats = torch.randn(B, V, Hs, Ws, C, device=dev).bfloat16().permute(0, 1, 4, 2, 3)
    ref_head = nn.Sequential(nn.Conv2d(C * V + 2, 512, 3, padding=1), nn.ReLU(),
                             nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                             nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False)).to(dev)
    sd = {"map_classifier." + k: v for k, v in ref_head.state_dict().items()}
    head = MH.MapHead(sd, V, C, device=dev)
    cudnn_head = ref_head.bfloat16().to(memory_format=torch.channels_last)

    a1 = warp_conv1(feats, pm, (Hg, Wg), head.w1, head.b1, relu=True, out_dtype=torch.bfloat16)
    a2 = MH.conv3x3_cl(a1, head.w2, head.b2, dilation=2, relu=True)

    with torch.no_grad():
        def unfused():
            return cudnn_head(mv.warp_views(feats, pm, (Hg, Wg)))
        def fused():
            return head(feats, pm, (Hg, Wg))
        ref = unfused().float()
        out = fused()
        err = float((out - ref).abs().max() / ref.abs().max())
        t = {
            "unfused_ms": timeit(unfused, iters),
            "map_head_ms": timeit(fused, iters),
            "conv1_fused_ms": timeit(lambda: warp_conv1(feats, pm, (Hg, Wg), head.w1, head.b1, relu=True, out_dtype=torch.bfloat16), iters),
            "conv2_tma_ms": timeit(lambda: MH.conv3x3_cl(a1, head.w2, head.b2, dilation=2, relu=True), iters),
            "conv2_cudnn_ms": timeit(lambda: F.relu(cudnn_head[2](a1)), iters),
            "conv3_head_ms": timeit(lambda: MH.conv3x3_head(a2, head.w3, dilation=4), iters),
            "conv3_head_one_pass_ms": timeit(lambda: MH.conv3x3_head(a2, head.w3, dilation=4, two_pass=False), iters),
            "conv3_cudnn_ms": timeit(lambda: cudnn_head[4](a2), iters),
        }
    flops2 = 2.0 * B * Hg * Wg * 512 * 512 * 9
    t.update({"B": B, "conv2_tma_tflops": flops2 / t["conv2_tma_ms"] / 1e9, "conv2_cudnn_tflops": flops2 / t["conv2_cudnn_ms"] / 1e9,
              "conv3_head_gbs": (B * Hg * Wg * 512 * 2 + B * Hg * Wg * 4) / t["conv3_head_ms"] / 1e6,
              "frames_per_s_map_head": B / t["map_head_ms"] * 1e3, "frames_per_s_unfused": B / t["unfused_ms"] * 1e3,
              "max_rel_diff_vs_cudnn_bf16": err})
    print(json.dumps(t))


if __name__ == "__main__":
    main()
