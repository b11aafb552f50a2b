"""View sharding of the perspective-warp path across the GPUs of one box.

Each camera's warp reads only its own feature map and writes a disjoint channel slice
``[v*C, (v+1)*C)`` of the concat (/root/reference/multiview_detector/models/persp_trans_detector.py:62-77),
so cameras partition across ranks with NO collective on the warp itself: a rank keeps its channel
slice.  The one real exchange on this path is downstream, when ``map_classifier[0]``
(persp_trans_detector.py:51,81) consumes the concat: a convolution is linear in its input channels,
so every rank convolves its own slice with the matching slice of the weight and the partial sums are
combined with one NCCL sum-reduce over NVLink; the bias and the coord_map term are added by one rank
and the ReLU runs after the reduce.

One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .geometry import create_coord_map


def view_partition(num_views: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced blocks ``(first_view, count)`` per rank; the first ``V % R`` ranks take one extra.
    Ranks beyond the number of views get an empty block (7 Wildtrack cameras on 8 GPUs leave one idle)."""
    if num_views <= 0 or world_size <= 0:
        raise ValueError("num_views and world_size must be positive")
    base, extra = divmod(num_views, world_size)
    out, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((start, n))
        start += n
    return out


@dataclass(frozen=True)
class ViewShard:
    """What one rank owns of a V-camera rig."""
    rank: int
    world_size: int
    num_views: int
    first_view: int
    count: int

    @property
    def owns_coord_map(self) -> bool:
        """The rank that carries the two coord_map channels (and conv1's bias): the last rank that has views,
        because the coord channels sit right after its slice in the reference's concat order."""
        return self.count > 0 and self.first_view + self.count == self.num_views

    def channel_slice(self, C: int) -> slice:
        """This rank's slice of the reference's V*C+2 concat channels (coord channels included for their owner)."""
        stop = (self.first_view + self.count) * C + (2 if self.owns_coord_map else 0)
        return slice(self.first_view * C, stop)

    @staticmethod
    def for_rank(num_views: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> "ViewShard":
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        first, count = view_partition(num_views, world_size)[rank]
        return ViewShard(rank, world_size, num_views, first, count)


def _cuda_warp(local_feats, local_mats, grid_shape, with_coord_map):
    from .ops import warp_views  # CUDA only; raises when the extension is missing
    return warp_views(local_feats, local_mats, grid_shape, with_coord_map=with_coord_map)


def warp_views_local(local_feats: torch.Tensor, proj_mats: torch.Tensor, grid_shape: Sequence[int],
                     shard: ViewShard, warp_fn: Optional[Callable] = None) -> torch.Tensor:
    """This rank's channel slice of ``warp_views``: ``[B, count*C (+2 on the coord owner), Hg, Wg]``.

    ``local_feats`` [B, count, C, H, W] are the features of this rank's cameras (they are produced per
    camera by the trunk, so they are born sharded); ``proj_mats`` [V,3,3] is the full rig.
    ``warp_fn(feats, mats, grid_shape, with_coord_map)`` defaults to the CUDA op; the CPU tests inject the oracle.
    """
    if local_feats.shape[1] != shard.count:
        raise ValueError(f"rank {shard.rank} owns {shard.count} views, got features for {local_feats.shape[1]}")
    warp_fn = warp_fn or _cuda_warp
    B = local_feats.shape[0]
    Hg, Wg = int(grid_shape[0]), int(grid_shape[1])
    if shard.count == 0:
        return local_feats.new_zeros((B, 0, Hg, Wg))
    mats = proj_mats[shard.first_view:shard.first_view + shard.count]
    return warp_fn(local_feats, mats, grid_shape, shard.owns_coord_map)


def gather_views(local_out: torch.Tensor, shard: ViewShard, C: int, group=None) -> torch.Tensor:
    """All-gather the rank slices into the reference's full ``[B, V*C+2, Hg, Wg]`` concat (test / debugging aid:
    the production path never materialises it on one GPU)."""
    if shard.world_size == 1:
        return local_out
    parts = view_partition(shard.num_views, shard.world_size)
    sizes = [n * C + (2 if (n > 0 and f + n == shard.num_views) else 0) for f, n in parts]
    B, _, Hg, Wg = local_out.shape
    widest = max(sizes)                      # all_gather wants equal shapes: pad the narrower slices
    mine = local_out.new_zeros((B, widest, Hg, Wg))
    mine[:, :local_out.shape[1]] = local_out
    bufs = [torch.empty_like(mine) for _ in sizes]
    dist.all_gather(bufs, mine, group=group)
    return torch.cat([b[:, :s] for b, s in zip(bufs, sizes)], dim=1)


def conv1_partial(local_out: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], shard: ViewShard,
                  C: int) -> torch.Tensor:
    """This rank's partial sum of ``map_classifier[0]`` = Conv2d(V*C+2, Cout, 3, padding=1)
    (persp_trans_detector.py:51): its slice of the concat convolved with the matching input-channel slice
    of ``weight`` [Cout, V*C+2, 3, 3]; the coord-map owner also contributes the coord term and the bias."""
    B, _, Hg, Wg = local_out.shape
    if shard.count == 0:
        return local_out.new_zeros((B, weight.shape[0], Hg, Wg))
    w = weight[:, shard.channel_slice(C)]
    return F.conv2d(local_out, w, bias if shard.owns_coord_map else None, padding=1)


def reduce_partials(partial: torch.Tensor, group=None, relu: bool = True) -> torch.Tensor:
    """Sum the conv1 partials over the view shards (NCCL all-reduce over NVLink / NVSwitch), then ReLU."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return F.relu_(partial) if relu else partial


def coord_map_for(grid_shape: Sequence[int], batch: int, like: torch.Tensor) -> torch.Tensor:
    return create_coord_map(grid_shape).to(device=like.device, dtype=like.dtype).repeat(batch, 1, 1, 1)


def pack_conv1_weight_shard(weight: torch.Tensor, shard: ViewShard, C: int) -> torch.Tensor:
    """This rank's slice of ``map_classifier[0].weight`` [512, V*C+2, 3, 3] packed for the fused kernel
    (:func:`mvdet_b200.conv1.pack_conv1_weight`); the coord-map owner also carries the two coord channels."""
    from .conv1 import pack_conv1_weight
    return pack_conv1_weight(weight[:, shard.channel_slice(C)], shard.count, C, with_coord=shard.owns_coord_map)


def conv1_partial_fused(local_feats: torch.Tensor, proj_mats: torch.Tensor, grid_shape: Sequence[int],
                        wpack_local: torch.Tensor, bias: Optional[torch.Tensor], shard: ViewShard) -> torch.Tensor:
    """:func:`conv1_partial` without materialising the rank's warped slice: warp + conv1 as one tcgen05 implicit GEMM
    (north_star (c)); bf16 channel-contiguous features in, fp32 partial sums out (no ReLU: that follows the reduce)."""
    from .conv1 import warp_conv1
    if shard.count == 0:
        B = local_feats.shape[0]
        return torch.zeros((B, 512, int(grid_shape[0]), int(grid_shape[1])), dtype=torch.float32, device=local_feats.device)
    mats = proj_mats[shard.first_view:shard.first_view + shard.count]
    return warp_conv1(local_feats, mats, grid_shape, wpack_local, bias if shard.owns_coord_map else None, relu=False,
                      with_coord=shard.owns_coord_map)


def conv1_fused_allreduce(local_feats: torch.Tensor, proj_mats: torch.Tensor, grid_shape: Sequence[int],
                          wpack_local: torch.Tensor, bias: Optional[torch.Tensor], shard: ViewShard,
                          bands: int = 4, relu: bool = True, group=None) -> torch.Tensor:
    """relu(conv1(concat of all ranks' warps) + bias) on every rank: the fused warp->conv1 kernel computes this rank's
    partial sums band by band (bands of grid rows, multiples of 8), and the NCCL all-reduce (+ ReLU) of band i runs on a
    side stream while the tensor cores work on band i+1 -- the exchange hides behind the GEMM except for the last band."""
    from .conv1 import warp_conv1
    B = local_feats.shape[0]
    Hg, Wg = int(grid_shape[0]), int(grid_shape[1])
    dev = local_feats.device
    phys = torch.empty((B, Hg, Wg, 512), dtype=torch.float32, device=dev)   # the channels-last storage of the result
    out = phys.permute(0, 3, 1, 2)
    distributed = dist.is_initialized() and dist.get_world_size(group) > 1
    rows_per_band = max(8, ((Hg + bands - 1) // bands + 7) // 8 * 8)
    edges = list(range(0, Hg, rows_per_band)) + [Hg]
    mats = proj_mats[shard.first_view:shard.first_view + shard.count]
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev) if distributed else None
    for y0, y1 in zip(edges[:-1], edges[1:]):
        if shard.count:
            warp_conv1(local_feats, mats, grid_shape, wpack_local, bias if shard.owns_coord_map else None, relu=False,
                       with_coord=shard.owns_coord_map, out=out, rows=(y0, y1))
        else:
            phys[:, y0:y1].zero_()
        if distributed:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for b in range(B):           # rows [y0, y1) of one batch element are one contiguous run of memory
                    dist.all_reduce(phys[b, y0:y1], op=dist.ReduceOp.SUM, group=group)
                if relu:
                    phys[:, y0:y1].relu_()
        elif relu:
            phys[:, y0:y1].relu_()
    if distributed:
        main.wait_stream(side)
        phys.record_stream(side)
    return out


# ------------------------------------------------------------------------------------------------------
# Fused GEMM + reduce-scatter over NVLink peer memory (no collective call on the data path)
# ------------------------------------------------------------------------------------------------------
def row_bands(grid_rows: int, world_size: int) -> List[int]:
    """Band edges ``[r0, r1, ..., Hg]``: rank r owns the grid rows ``[edges[r], edges[r+1])``.  Bands are whole tile rows
    (8 grid rows) of the fused kernel, balanced like :func:`view_partition`."""
    tile_rows = (int(grid_rows) + 7) // 8
    edges, acc = [0], 0
    for _, n in view_partition(tile_rows, world_size):
        acc += n
        edges.append(min(int(grid_rows), 8 * acc))
    return edges


class Conv1Inbox:
    """Symmetric-memory inboxes for :func:`conv1_fused_reduce_scatter`: on every rank one fp32 buffer
    ``[world, B, band_rows_max, Wg, 512]`` (slot s = the partial sums rank s computed for MY rows), mapped into every
    process of the group so that the fused kernel's epilogue can store into its peers' inboxes directly."""

    def __init__(self, grid_shape: Sequence[int], batch: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("the fused reduce-scatter addresses at most 8 peers (one NVSwitch box)")
        self.Hg, self.Wg, self.B = int(grid_shape[0]), int(grid_shape[1]), int(batch)
        self.edges = row_bands(self.Hg, self.world)
        self.band_rows_max = max(b - a for a, b in zip(self.edges[:-1], self.edges[1:]))
        self.buf = symm.empty((self.world, self.B, self.band_rows_max, self.Wg, 512), dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buf, group)
        self.buf.zero_()                       # slots of ranks without cameras are never written: they must read as zero
        self.handle.barrier(channel=0)
        # every rank's inbox as mapped into this process (get_buffer accounts for the tensor's offset inside the allocation)
        self.peers = [self.handle.get_buffer(r, tuple(self.buf.shape), torch.float32) for r in range(self.world)]
        self.peer_ptrs = [int(t.data_ptr()) for t in self.peers]
        self.side_stream = torch.cuda.Stream(device)
        if self.peer_ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory: my own inbox is mapped at an unexpected address")

    @property
    def my_rows(self) -> Tuple[int, int]:
        return self.edges[self.rank], self.edges[self.rank + 1]


def _reduce_my_inbox(inbox: "Conv1Inbox", bias: Optional[torch.Tensor], relu: bool, out_dtype: torch.dtype) -> torch.Tensor:
    """Sum the slots of this rank's inbox over its own rows (+ bias, ReLU) in one pass: [B, 512, rows, Wg] channels-last."""
    import ctypes as C
    from . import _lib
    y0, y1 = inbox.my_rows
    rows = y1 - y0
    dev = inbox.buf.device
    band = torch.empty((inbox.B, rows, inbox.Wg, 512), dtype=out_dtype, device=dev)
    if rows == 0:                      # more ranks than tile rows of the grid: this rank owns no band
        return band.permute(0, 3, 1, 2)
    bias_ptr = None
    if bias is not None:
        bias = bias.detach().to(device=dev, dtype=torch.float32).contiguous()
        bias_ptr = C.c_void_p(bias.data_ptr())
    b_pitch = inbox.band_rows_max * inbox.Wg * 512
    rc = _lib.load().mvwarp_inbox_reduce(C.c_void_p(inbox.buf.data_ptr()), inbox.world, b_pitch * inbox.B, b_pitch, inbox.B,
                                         rows * inbox.Wg * 512, bias_ptr, int(relu), C.c_void_p(band.data_ptr()),
                                         _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32, dev.index,
                                         C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "mvwarp_inbox_reduce")
    return band.permute(0, 3, 1, 2)


def conv1_fused_reduce_scatter(local_feats: torch.Tensor, proj_mats: torch.Tensor, grid_shape: Sequence[int],
                               wpack_local: torch.Tensor, bias: Optional[torch.Tensor], shard: ViewShard, inbox: Conv1Inbox,
                               relu: bool = True, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """This rank's row band ``[B, 512, rows, Wg]`` of ``relu(conv1(concat of all ranks' warps) + bias)``.

    The fused warp -> conv1 kernel computes this rank's partial sums for the WHOLE map, and its epilogue stores each tile
    straight into the inbox of the rank that owns the tile's rows (peer-to-peer stores over NVLink / NVSwitch riding along with
    the tensor-core work: ``mvwarp_warp_conv1_fwd_p2p``) -- the reduce-scatter costs no kernel, no SMs and no collective call.
    After one barrier every rank sums the slots of its own inbox in rank order (deterministic), adds the bias, applies the ReLU.
    A second barrier keeps the next call from overwriting an inbox that is still being read."""
    import ctypes as C
    from . import _lib
    from .plan import get_plan
    B, nv, Cc, Hs, Ws = local_feats.shape
    if B != inbox.B or (int(grid_shape[0]), int(grid_shape[1])) != (inbox.Hg, inbox.Wg):
        raise ValueError("the inbox was sized for another batch / grid")
    if shard.count:
        if local_feats.dtype != torch.bfloat16 or local_feats.stride(2) != 1:
            raise ValueError("conv1_fused_reduce_scatter needs bf16 channel-contiguous features")
        mats = proj_mats[shard.first_view:shard.first_view + shard.count]
        plan = get_plan(mats, (Hs, Ws), grid_shape, local_feats.device)
        ptrs = (C.c_uint64 * inbox.world)(*inbox.peer_ptrs)
        rows = (C.c_int32 * (inbox.world + 1))(*inbox.edges)
        stream = C.c_void_p(torch.cuda.current_stream(local_feats.device).cuda_stream)
        rc = _lib.load().mvwarp_warp_conv1_fwd_p2p(plan.handle, 0, nv, C.c_void_p(local_feats.data_ptr()), _lib.i64x(local_feats.stride()),
                                                   C.c_void_p(wpack_local.data_ptr()), B, Cc, int(shard.owns_coord_map), inbox.world,
                                                   inbox.rank, ptrs, rows, inbox.band_rows_max, stream)
        _lib.check(rc, "mvwarp_warp_conv1_fwd_p2p")
    inbox.handle.barrier(channel=0)            # every rank's stores into my inbox have landed
    band = _reduce_my_inbox(inbox, bias, relu, out_dtype)          # slots added in rank order (deterministic)
    inbox.handle.barrier(channel=1)            # my inbox has been read: peers may start the next step
    return band


def conv1_fused_reduce_scatter_ce(local_feats: torch.Tensor, proj_mats: torch.Tensor, grid_shape: Sequence[int],
                                  wpack_local: torch.Tensor, bias: Optional[torch.Tensor], shard: ViewShard, inbox: Conv1Inbox,
                                  relu: bool = True, workspace: Optional[torch.Tensor] = None,
                                  out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Same result as :func:`conv1_fused_reduce_scatter`, with the exchange on the COPY ENGINES instead of in the epilogue.

    The fused kernel runs band by band -- the bands are the ranks' row bands, visited in the rotated order r+1, r+2, ..., r --
    and writes its partial sums locally at full speed; as soon as band j is done, a side stream pushes it into slot r of owner
    j's inbox with peer-to-peer ``cudaMemcpyAsync`` (copy engines over NVLink: no SMs, so the tensor cores are not disturbed,
    and thanks to the rotation every owner receives from exactly one sender at a time).  The rank's own band comes last and is
    written straight into its own inbox slot, so the last remote copy hides behind it; what remains exposed is one barrier
    and the sum over the slots."""
    from .conv1 import warp_conv1
    B, nv, Cc, Hs, Ws = local_feats.shape
    Hg, Wg = inbox.Hg, inbox.Wg
    dev = local_feats.device
    R, r = inbox.world, inbox.rank
    if workspace is None:
        workspace = torch.empty((B, Hg, Wg, 512), dtype=torch.float32, device=dev)
    main = torch.cuda.current_stream(dev)
    side = inbox.side_stream
    mats = proj_mats[shard.first_view:shard.first_view + shard.count]
    peers = inbox.peers
    for k in range(R):
        j = (r + 1 + k) % R
        y0, y1 = inbox.edges[j], inbox.edges[j + 1]
        if y1 <= y0 or not shard.count:
            continue
        if j == r:        # my own band: the kernel writes into my inbox slot directly (the slot is [B, band_rows_max, Wg, 512])
            # warp_conv1 addresses rows of a whole [B,512,Hg,Wg] map: hand it a view whose row y0 is the slot's row 0
            base = inbox.buf[r]                                           # [B, band_rows_max, Wg, 512]
            whole = torch.as_strided(base, (B, Hg, Wg, 512), base.stride(), storage_offset=base.storage_offset() - y0 * Wg * 512)
            warp_conv1(local_feats, mats, grid_shape, wpack_local, None, relu=False, with_coord=shard.owns_coord_map,
                       out=whole.permute(0, 3, 1, 2), rows=(y0, y1))
        else:
            warp_conv1(local_feats, mats, grid_shape, wpack_local, None, relu=False, with_coord=shard.owns_coord_map,
                       out=workspace.permute(0, 3, 1, 2), rows=(y0, y1))
            side.wait_stream(main)
            with torch.cuda.stream(side):
                peers[j][r, :, :y1 - y0].copy_(workspace[:, y0:y1], non_blocking=True)
    main.wait_stream(side)
    workspace.record_stream(side)
    inbox.handle.barrier(channel=0)
    band = _reduce_my_inbox(inbox, bias, relu, out_dtype)
    inbox.handle.barrier(channel=1)
    return band
