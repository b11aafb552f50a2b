#!/usr/bin/env python
"""Model-level effect of the drop-in: everything of PerspTransDetector.forward AFTER the trunk (persp_trans_detector.py:65-82) --
F.interpolate(512 ch -> 270x480), img_classifier, the warp + concat, map_classifier -- forward + backward at the Wildtrack shape,
from a synthetic trunk output [B,7,512,90,160].  Four wirings of the same modules (random weights, same seeds):
  torch      the reference's loop with ATen grid_sample on a precomputed grid + torch.cat (NCHW, fp32)  [no kornia overhead]
  seam2      mvdet_b200.integration.patch_detector(model)                       (channels-last, fp32)
  seam2+up   patch_detector(model, fuse_upsample=True)                           (channels-last, fp32)
  seam2 nchw+up   patch_detector(model, fuse_upsample=True, channels_last=False): the UNMODIFIED NCHW model; the op transposes the
                  low-res maps once and returns a channels-last result
  seam2+up bf16   the same under torch.autocast(bfloat16) for the convolutions, features in bf16
  seam2 conv1 bf16  patch_detector(model, fused_conv1=True) under autocast: fused tcgen05 conv1 forward, backward without cuDNN
  seam2 conv1+up bf16  the same with fuse_upsample=True: the fused kernels read the low-res trunk output, no 512-channel F.interpolate
Then INFERENCE (torch.no_grad, bf16 autocast): seam2, seam2 + fused_head, seam2 + fused_head + fuse_upsample.
Prints one JSON line.   usage: detector_head_bench.py [B] [iters]"""
import copy, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
from mvdet_b200 import integration, plan as P
from mvdet_b200.geometry import RIGS, create_coord_map, synthetic_rig

DEV = torch.device("cuda", 0)


class Head(nn.Module):
    """The reference's post-trunk modules under the reference's attribute names (persp_trans_detector.py:46-54); base_pt1/2 = identity."""

    def __init__(self, spec, pm):
        super().__init__()
        V = spec.num_cam
        self.num_cam = V
        self.base_pt1, self.base_pt2 = nn.Identity(), nn.Identity()
        self.img_classifier = nn.Sequential(nn.Conv2d(512, 64, 1), nn.ReLU(), nn.Conv2d(64, 2, 1, bias=False))
        self.map_classifier = nn.Sequential(nn.Conv2d(512 * V + 2, 512, 3, padding=1), nn.ReLU(),
                                            nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                                            nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False))
        self.proj_mats = [pm[v] for v in range(V)]
        self.upsample_shape, self.reducedgrid_shape = list(spec.feat_shape), list(spec.grid_shape)
        self.coord_map = create_coord_map(spec.grid_shape)

    def forward(self, feats_low, visualize=False):           # the reference's loop, ATen kernels
        B = feats_low.shape[0]
        world, res = [], []
        for cam in range(self.num_cam):
            f = F.interpolate(feats_low[:, cam], self.upsample_shape, mode="bilinear")
            res.append(self.img_classifier(f))
            world.append(F.grid_sample(f, self.grid[cam:cam + 1].expand(B, -1, -1, -1), mode="bilinear", padding_mode="zeros",
                                       align_corners=True))
        world = torch.cat(world + [self.coord_map.repeat([B, 1, 1, 1]).to(f.device)], dim=1)
        return F.interpolate(self.map_classifier(world), self.reducedgrid_shape, mode="bilinear"), res


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    spec = RIGS["wildtrack"]
    pm = synthetic_rig(spec)
    torch.manual_seed(0)
    ref = Head(spec, pm).to(DEV)
    Hs, Ws = spec.feat_shape
    ref.grid = P.get_plan(pm, (Hs, Ws), spec.grid_shape, DEV).device_grid().to(DEV)
    feats_low = torch.randn(B, spec.num_cam, 512, Hs // 3, Ws // 3, device=DEV)

    def run(model, x, autocast=False):
        x = x.detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            m, res = model(x)
            loss = m.float().square().mean() + sum(r.float().square().mean() for r in res)
        loss.backward()
        for p in model.parameters():
            p.grad = None
        return m

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {"B": B, "iters": iters}
    m_ref = run(ref, feats_low)
    out["torch_ms"] = round(timeit(lambda: run(ref, feats_low)), 3)
    x_cl = feats_low.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    for name, kw, ac in (("seam2", {}, False), ("seam2_fused_upsample", {"fuse_upsample": True}, False),
                         ("seam2_nchw_model_fused_upsample", {"fuse_upsample": True, "channels_last": False}, False),
                         ("seam2_fused_upsample_bf16", {"fuse_upsample": True}, True),
                         ("seam2_fused_conv1_bf16", {"fused_conv1": True}, True),
                         ("seam2_fused_conv1_fused_upsample_bf16", {"fused_conv1": True, "fuse_upsample": True}, True)):
        model = copy.deepcopy(ref)
        kw = dict({"channels_last": True}, **kw)
        integration.patch_detector(model, **kw)
        x = feats_low if not kw["channels_last"] else (x_cl.bfloat16() if ac else x_cl)
        m = run(model, x, ac)
        out[name + "_ms"] = round(timeit(lambda: run(model, x, ac)), 3)
        out[name + "_map_rel_diff"] = float((m.float() - m_ref).abs().max() / m_ref.abs().max())
        del model
        torch.cuda.empty_cache()
    def infer(model, x):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(x)[0]

    for name, kw in (("infer_seam2", {}), ("infer_fused_head", {"fused_head": True}),
                     ("infer_fused_head_fused_upsample", {"fused_head": True, "fuse_upsample": True})):
        model = copy.deepcopy(ref).eval()
        integration.patch_detector(model, channels_last=True, **kw)
        x = x_cl.bfloat16()
        m = infer(model, x)
        out[name + "_ms"] = round(timeit(lambda: infer(model, x)), 3)
        out[name + "_map_rel_diff"] = float((m.float() - m_ref).abs().max() / m_ref.abs().max())
        del model
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
