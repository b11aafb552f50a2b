#!/usr/bin/env python
"""Where the B=8 bf16 training step (BASELINE.json configs[2]) spends its time: every piece of warp_conv1_trainable timed alone
with CUDA events.  usage: train_step_breakdown.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvdet_b200 import conv1, ops, plan as P
from mvdet_b200.map_head import conv3x3_cl
from mvdet_b200.geometry import RIGS, synthetic_rig


def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    spec = RIGS["wildtrack"]; pm = synthetic_rig(spec)
    V, C = spec.num_cam, 512
    Hs, Ws = spec.feat_shape; Hg, Wg = spec.grid_shape
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    feats = torch.randn(B, V, Hs, Ws, C, device=dev).bfloat16().permute(0, 1, 4, 2, 3)
    weight = torch.randn(512, V * C + 2, 3, 3, device=dev) * 0.02
    bias = torch.randn(512, device=dev)
    go = torch.randn(B, Hg, Wg, 512, device=dev).permute(0, 3, 1, 2)
    wpack = conv1.pack_conv1_weight(weight, V, C)
    plan = P.get_plan(pm, (Hs, Ws), (Hg, Wg), dev)
    out = conv1.warp_conv1(feats, pm, (Hg, Wg), wpack, bias, relu=True)
    g16 = (go * (out > 0)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    packs = conv1._dgrad_packs(weight, V, C)
    gx_buf = torch.empty((B, Hg, Wg, V * C), dtype=torch.bfloat16, device=dev)
    gx = gx_buf.permute(0, 3, 1, 2)
    g_feats = torch.empty((B, V, Hs, Ws, C), dtype=torch.bfloat16, device=dev).permute(0, 1, 4, 2, 3)

    def dgrad():
        for v, wp in enumerate(packs):
            conv3x3_cl(g16v, wp, None, dilation=1, relu=False, out_dtype=torch.bfloat16, out=gx[:, v * C:(v + 1) * C])

    Xp, Gp = conv1._padded_buffers(dev, B, Hg, Wg, (V * C + 2 + 7) // 8 * 8)
    g16v = Gp[:, :Hg, :Wg].permute(0, 3, 1, 2)
    go_cl = go.contiguous(memory_format=torch.channels_last)

    res = {"B": B,
           "fwd_fused_ms": timeit(lambda: conv1.warp_conv1(feats, pm, (Hg, Wg), wpack, bias, relu=True)),
           "relu_bwd_cast_kernel_ms": timeit(lambda: conv1.relu_bwd_cast(go_cl, out, g16v, True)),
           "torch_mask_cast_biasgrad_ms": timeit(lambda: ((go * (out > 0)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last), (go * (out > 0)).sum(dim=(0, 2, 3)))),
           "dgrad_tcgen05_ms": timeit(dgrad),
           "wgrad_warp_plus_gemms_ms": timeit(lambda: conv1.conv1_wgrad_gemm(plan, feats, Xp, Gp)),
           "warp_bwd_ms": timeit(lambda: ops.warp_views_bwd_into(plan, gx, g_feats)),
           }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
