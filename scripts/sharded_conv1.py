#!/usr/bin/env python
"""View-sharded warp + map_classifier[0] partial sums + NCCL all-reduce (BASELINE.json configs[3]).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/sharded_conv1.py

Each rank warps its own cameras (mvdet_b200.warp_views, no collective), convolves its channel slice with the matching
slice of conv1's weight (cuDNN; the bias and the coord-map term on the coord owner), and the [B,512,Hg,Wg] partial sums
are combined with one NCCL all-reduce over NVLink, ReLU after.  Part 1 checks the result against the unsharded
computation (oracle warp + F.conv2d, fp32, small channel count); part 2 times the three phases at full size in bf16.
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvdet_b200 import sharding as S  # noqa: E402
from mvdet_b200.geometry import RIG32, WILDTRACK, RigSpec, synthetic_rig  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}

    # ---- part 1: parity with the unsharded reference computation (fp32, C=16, Wildtrack shape) ----------------
    torch.backends.cudnn.allow_tf32 = False   # cuDNN convolutions default to TF32; the parity check wants fp32
    spec = WILDTRACK
    pm = synthetic_rig(spec)
    V, C, Cout, B = spec.num_cam, 16, 32, 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, V, C, *spec.feat_shape, generator=g)
    weight = torch.randn(Cout, V * C + 2, 3, 3, generator=g) * 0.05
    bias = torch.randn(Cout, generator=g)
    shard = S.ViewShard.for_rank(V)
    local_x = x[:, shard.first_view:shard.first_view + shard.count].permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3).to(dev)
    out_local = S.warp_views_local(local_x, pm, spec.grid_shape, shard)
    partial = S.conv1_partial(out_local, weight.to(dev), bias.to(dev), shard, C)
    reduced = S.reduce_partials(partial, relu=True)
    if rank == 0:
        from oracle import kornia_warp as K
        ref = F.relu(F.conv2d(K.warp_views_oracle(x, pm, spec.grid_shape), weight, bias, padding=1))
        err = float((reduced.cpu() - ref).abs().max() / ref.abs().max())
        res["parity_max_rel_err_fp32"] = err
        assert err <= 1e-4, err   # cuDNN fp32 conv (TF32 off) vs CPU conv: summation order only

    # ---- part 2: timing at full size, bf16: 4 cameras per rank of the full-resolution rig --------------------
    per_rank = 4
    spec = RigSpec("rig32", per_rank * world, worldgrid_shape=RIG32.worldgrid_shape, grid_reduce=1, seed=2)
    pm = synthetic_rig(spec)
    V, C, Cout, B = spec.num_cam, 512, 512, 1
    shard = S.ViewShard.for_rank(V)
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    feats = torch.randn(B, shard.count, Hs, Ws, C, device=dev, dtype=torch.bfloat16).permute(0, 1, 4, 2, 3)
    weight = (torch.randn(Cout, V * C + 2, 3, 3, device=dev, dtype=torch.bfloat16) * 0.02).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(Cout, device=dev, dtype=torch.bfloat16)
    dist.broadcast(bias, src=world - 1)      # one model: every rank holds the same bias (the weight slices are disjoint anyway)
    wpack = S.pack_conv1_weight_shard(weight.float(), shard, C)
    bias32 = bias.float()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times, ftimes = [], []
    for it in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        out_local = S.warp_views_local(feats, pm, (Hg, Wg), shard)
        ev[1].record()
        partial = S.conv1_partial(out_local, weight, bias, shard, C)
        ev[3].record()
        torch.cuda.synchronize()
        t_conv = ev[1].elapsed_time(ev[3])
        partial = partial.float().contiguous()
        torch.cuda.synchronize()
        dist.barrier()                       # time the exchange itself, not the skew between the ranks' convolutions
        torch.cuda.synchronize()
        ev[2].record()
        reduced = S.reduce_partials(partial, relu=True)
        ev[3].record()
        torch.cuda.synchronize()
        times.append([ev[0].elapsed_time(ev[1]), t_conv, ev[2].elapsed_time(ev[3])])
        ref_partial = partial.float() if it == 0 else None
        del out_local, partial, reduced
        fe[0].record()
        fpart = S.conv1_partial_fused(feats, pm, (Hg, Wg), wpack, bias32, shard)    # warp + conv1 partial in one kernel
        fe[1].record()
        torch.cuda.synchronize()
        ftimes.append(fe[0].elapsed_time(fe[1]))
        del fpart
    # whole pipeline: band-wise fused GEMM with the all-reduce of each band overlapped on a side stream
    ptimes = []
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        pe[0].record()
        full = S.conv1_fused_allreduce(feats, pm, (Hg, Wg), wpack, bias32, shard, bands=6)
        pe[1].record(); torch.cuda.synchronize()
        ptimes.append(pe[0].elapsed_time(pe[1]))
        if it == 0:
            seq = S.reduce_partials(S.conv1_partial_fused(feats, pm, (Hg, Wg), wpack, bias32, shard), relu=True)
            res_pipe_err = float((full - seq).abs().max() / seq.abs().max())
            del seq
        del full
    # fused GEMM + reduce-scatter over NVLink peer memory: each rank ends with its row band of relu(conv1); no collective call
    inbox = S.Conv1Inbox((Hg, Wg), B, dev)
    rtimes, gtimes = [], []
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        pe[0].record()
        band = S.conv1_fused_reduce_scatter(feats, pm, (Hg, Wg), wpack, bias32, shard, inbox)
        pe[1].record()
        # optional: every rank gets the whole map (bf16, what map_classifier[2] reads): all-gather of the bands, padded to equal size
        mine = torch.zeros((B, inbox.band_rows_max, Wg, 512), dtype=torch.bfloat16, device=dev)
        mine[:, :band.shape[2]] = band.permute(0, 2, 3, 1)
        gathered = torch.empty((world,) + tuple(mine.shape), dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(gathered, mine)
        pe[2].record(); torch.cuda.synchronize()
        rtimes.append(pe[0].elapsed_time(pe[1])); gtimes.append(pe[1].elapsed_time(pe[2]))
        if it == 0:
            seq = S.reduce_partials(S.conv1_partial_fused(feats, pm, (Hg, Wg), wpack, bias32, shard), relu=True)
            y0, y1 = inbox.my_rows
            res_p2p_err = float((band - seq[:, :, y0:y1]).abs().max() / seq.abs().max())
            del seq
        del band, mine, gathered
    # the same exchange on the copy engines (band-wise kernel launches in rotated owner order, P2P cudaMemcpyAsync on a side stream)
    ctimes = []
    work = torch.empty((B, Hg, Wg, 512), dtype=torch.float32, device=dev)
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        pe[0].record()
        band = S.conv1_fused_reduce_scatter_ce(feats, pm, (Hg, Wg), wpack, bias32, shard, inbox, workspace=work)
        pe[1].record(); torch.cuda.synchronize()
        ctimes.append(pe[0].elapsed_time(pe[1]))
        if it == 0:
            seq = S.reduce_partials(S.conv1_partial_fused(feats, pm, (Hg, Wg), wpack, bias32, shard), relu=True)
            y0, y1 = inbox.my_rows
            res_ce_err = float((band - seq[:, :, y0:y1]).abs().max() / seq.abs().max())
            del seq
        del band
    del work
    rt = torch.tensor([sum(rtimes[1:]) / len(rtimes[1:]), sum(gtimes[1:]) / len(gtimes[1:]), res_p2p_err,
                       sum(ctimes[1:]) / len(ctimes[1:]), res_ce_err], device=dev)
    dist.all_reduce(rt, op=dist.ReduceOp.MAX)
    pt = torch.tensor([sum(ptimes[1:]) / len(ptimes[1:])], device=dev)
    dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    ft = torch.tensor([sum(ftimes[1:]) / len(ftimes[1:])], device=dev)
    dist.all_reduce(ft, op=dist.ReduceOp.MAX)
    t = torch.tensor(times[1:], device=dev).mean(0)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        flops = 2.0 * Hg * Wg * Cout * (shard.count * C + 2) * 9
        res.update({"rig": f"{V} cameras x 512ch x {Hs}x{Ws} -> {Hg}x{Wg}, bf16, {per_rank} cameras per rank",
                    "warp_ms": float(t[0]), "conv1_partial_ms": float(t[1]), "allreduce_relu_ms": float(t[2]),
                    "conv1_partial_tflops": flops / (float(t[1]) * 1e-3) / 1e12,
                    "fused_pipelined_with_allreduce_ms": float(pt[0]), "pipelined_vs_sequential_max_rel_diff": res_pipe_err,
                    "fused_reduce_scatter_p2p_ms": float(rt[0]), "p2p_allgather_bf16_ms": float(rt[1]),
                    "p2p_vs_allreduce_max_rel_diff": float(rt[2]),
                    "fused_reduce_scatter_copy_engine_ms": float(rt[3]), "copy_engine_vs_allreduce_max_rel_diff": float(rt[4]),
                    "fused_warp_conv1_partial_ms": float(ft[0]), "fused_tflops": flops / (float(ft[0]) * 1e-3) / 1e12,
                    "allreduce_bytes": B * Cout * Hg * Wg * 4,
                    "allreduce_busbw_gbs": 2 * (world - 1) / world * B * Cout * Hg * Wg * 4 / (float(t[2]) * 1e-3) / 1e9})
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
