"""Worker of tests/test_gpu_multi.py: run under torchrun on >= 2 GPUs.  Every exchange variant of the view-sharded conv1
(NCCL all-reduce, overlapped band-wise all-reduce, fused epilogue reduce-scatter over peer memory, copy-engine reduce-scatter)
against the UNSHARDED result: relu(conv3x3(bf16(oracle warp of all cameras)) + bias) computed on rank 0's CPU."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvdet_b200 import sharding as S  # noqa: E402
from mvdet_b200.geometry import RigSpec, synthetic_rig  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # a scaled-down version of BASELINE.json configs[3]: 2 cameras per rank + one extra on rank 0, 54x96 features -> 48x144 grid.
    # With 4 ranks the grid shrinks to 24x72 = 3 tile rows, so that the last rank owns an EMPTY row band (more ranks than tile rows).
    V, Cc, B = 2 * world + 1, 64, 2
    spec = RigSpec("dist", V, img_shape=(1080, 1920), worldgrid_shape=(480, 1440), img_reduce=20, grid_reduce=10 if world <= 2 else 20, seed=51)
    pm = synthetic_rig(spec)
    Hg, Wg = spec.grid_shape
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, V, Cc, *spec.feat_shape, generator=g).bfloat16()          # same on every rank
    weight = torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05
    bias = torch.randn(512, generator=g)
    shard = S.ViewShard.for_rank(V)
    local_x = x[:, shard.first_view:shard.first_view + shard.count].permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3).to(dev)
    wpack = S.pack_conv1_weight_shard(weight, shard, Cc).to(dev)
    bias_d = bias.to(dev)
    res = {}
    full = S.reduce_partials(S.conv1_partial_fused(local_x, pm, (Hg, Wg), wpack, bias_d, shard), relu=True)
    banded = S.conv1_fused_allreduce(local_x, pm, (Hg, Wg), wpack, bias_d, shard, bands=3)
    inbox = S.Conv1Inbox((Hg, Wg), B, dev)
    y0, y1 = inbox.my_rows
    band_p2p = S.conv1_fused_reduce_scatter(local_x, pm, (Hg, Wg), wpack, bias_d, shard, inbox).clone()
    band_ce = S.conv1_fused_reduce_scatter_ce(local_x, pm, (Hg, Wg), wpack, bias_d, shard, inbox).clone()
    band_ce2 = S.conv1_fused_reduce_scatter_ce(local_x, pm, (Hg, Wg), wpack, bias_d, shard, inbox)
    torch.cuda.synchronize()
    if rank == 0:
        from oracle import kornia_warp as K
        X = K.warp_views_oracle(x.float(), pm, (Hg, Wg)).bfloat16().float()
        ref = F.relu(F.conv2d(X, weight.bfloat16().float(), bias, padding=1))
    else:
        ref = torch.empty(B, 512, Hg, Wg)
    ref = ref.to(dev)
    dist.broadcast(ref, src=0)
    scale = float(ref.abs().max())
    amax = lambda t: float(t.abs().max()) if t.numel() else 0.0          # an empty band has nothing to compare
    errs = {"allreduce": amax(full - ref) / scale,
            "allreduce_banded": amax(banded - ref) / scale,
            "reduce_scatter_p2p": amax(band_p2p - ref[:, :, y0:y1]) / scale,
            "reduce_scatter_ce": amax(band_ce - ref[:, :, y0:y1]) / scale}
    t = torch.tensor(list(errs.values()) + [0.0 if torch.equal(band_ce, band_ce2) else 1.0], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        vals = t.tolist()
        res = dict(zip(errs.keys(), vals[:-1]))
        res.update({"world": world, "views": V, "rows_of_rank0": [y0, y1], "ce_bitwise_reproducible": vals[-1] == 0.0})
        print("DIST_CONV1 " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
