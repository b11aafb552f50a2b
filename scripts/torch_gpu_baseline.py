#!/usr/bin/env python
"""What the reference's own GPU kernels do on this B200, for context next to bench.py's numbers.

The reference's path on a GPU is kornia glue (about 25 tiny launches + two synchronising torch.inverse per camera) in front of
ATen's grid_sampler_2d, then torch.cat with coord_map, and autograd's CatBackward + grid_sampler_2d_backward (atomicAdd scatter)
-- persp_trans_detector.py:68-69,75,77.  kornia is not installed here, so this script times the part that does the work and
flatters the reference: V x F.grid_sample(feats_v, grid_v) with the sampling grid PRECOMPUTED (this library's plan grid, the
same bits as the kornia chain), torch.cat with the coordinate map, and torch.autograd.grad through both.  No kornia overhead,
no per-call matrix upload.  It does not import oracle/ and is not used by any test or by bench.py.

usage: torch_gpu_baseline.py [wildtrack|multiviewx] [B] [iters]      -> one JSON line
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import mvdet_b200 as mv
from mvdet_b200 import plan as P
from mvdet_b200.geometry import RIGS, create_coord_map, synthetic_rig


def main():
    rig = sys.argv[1] if len(sys.argv) > 1 else "wildtrack"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    spec = RIGS[rig]
    dev = torch.device("cuda", 0)
    pm = synthetic_rig(spec)
    V, C = spec.num_cam, 512
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    plan = P.get_plan(pm, (Hs, Ws), (Hg, Wg), dev)
    grid = plan.device_grid().to(dev)                                   # [V,Hg,Wg,2], the kornia chain's bits
    coord = create_coord_map((Hg, Wg)).to(dev)
    torch.manual_seed(0)
    res = {"rig": rig, "B": B, "iters": iters}
    for layout in ("nchw", "channels_last"):
        if layout == "nchw":
            feats = [torch.randn(B, C, Hs, Ws, device=dev).requires_grad_(True) for _ in range(V)]
        else:
            feats = [torch.randn(B, Hs, Ws, C, device=dev).permute(0, 3, 1, 2).requires_grad_(True) for _ in range(V)]
        go = torch.randn(B, V * C + 2, Hg, Wg, device=dev)
        if layout == "channels_last":
            go = go.contiguous(memory_format=torch.channels_last)

        def fwd():
            outs = [F.grid_sample(feats[v], grid[v:v + 1].expand(B, -1, -1, -1), mode="bilinear", padding_mode="zeros",
                                  align_corners=True) for v in range(V)]
            return torch.cat(outs + [coord.repeat(B, 1, 1, 1)], dim=1)

        def fwdbwd():
            out = fwd()
            return torch.autograd.grad(out, feats, go)

        def timeit(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        with torch.no_grad():
            t_f = timeit(fwd)
        t_fb = timeit(fwdbwd)
        res[layout] = {"fwd_ms": round(t_f, 4), "fwd_bwd_ms": round(t_fb, 4), "frames_per_s_fwd_bwd": round(B / t_fb * 1e3, 1)}
        # ours, same tensors stacked, same layout
        x = torch.stack([f.detach() for f in feats], dim=1).requires_grad_(True)

        def ours_fwd():
            return mv.warp_views_with_plan(x, plan)

        def ours_fwdbwd():
            return torch.autograd.grad(ours_fwd(), x, go)

        with torch.no_grad():
            o_f = timeit(ours_fwd)
        o_fb = timeit(ours_fwdbwd)
        res[layout]["ours_fwd_ms"] = round(o_f, 4)
        res[layout]["ours_fwd_bwd_ms"] = round(o_fb, 4)
        with torch.no_grad():
            a, b = fwd(), ours_fwd()
        res[layout]["max_abs_diff_vs_ours"] = float((a - b).abs().max())
        del feats, go, x, a, b
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
