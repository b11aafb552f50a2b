"""View sharding (N > 1) on CPU: world_size-2 and -3 gloo groups.  The per-rank compute is stood in by
the oracle (there is no GPU here); what is under test is the host logic: the partition, the channel
slices, the coord-map owner, the all-gather order and the conv1 partial-sum reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from mvdet_b200 import sharding as S
from oracle import kornia_warp as K
from tests.util import small_rig


def test_view_partition_properties():
    for V in (1, 6, 7, 32):
        for R in (1, 2, 3, 4, 8):
            parts = S.view_partition(V, R)
            assert len(parts) == R and sum(n for _, n in parts) == V
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(R - 1))
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
            owners = [S.ViewShard(r, R, V, *parts[r]).owns_coord_map for r in range(R)]
            assert sum(owners) == 1
    assert S.view_partition(7, 8)[7] == (7, 0)
    assert S.view_partition(32, 8) == [(4 * r, 4) for r in range(8)]
    with pytest.raises(ValueError):
        S.view_partition(0, 2)


def _oracle_warp(feats, mats, grid_shape, with_coord_map):
    out = K.warp_views_oracle(feats, mats, grid_shape)
    return out if with_coord_map else out[:, :-2]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, V, C, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        pm, src_hw, grid_hw = small_rig(num_cam=V, seed=21)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(2, V, C, *src_hw, generator=g)                     # same on every rank
        weight = torch.randn(6, V * C + 2, 3, 3, generator=g) * 0.1
        bias = torch.randn(6, generator=g)
        shard = S.ViewShard.for_rank(V)
        local = x[:, shard.first_view:shard.first_view + shard.count]
        out_local = S.warp_views_local(local, pm, grid_hw, shard, warp_fn=_oracle_warp)
        full = S.gather_views(out_local, shard, C)
        partial = S.conv1_partial(out_local, weight, bias, shard, C)
        reduced = S.reduce_partials(partial.clone(), relu=True)
        if rank == 0:
            ref_full = K.warp_views_oracle(x, pm, grid_hw)
            ref_conv = F.relu(F.conv2d(ref_full, weight, bias, padding=1))
            ret["gather_equal"] = bool(torch.equal(full, ref_full))
            ret["conv_err"] = float((reduced - ref_conv).abs().max() / ref_conv.abs().max())
            ret["slice"] = (shard.channel_slice(C).start, shard.channel_slice(C).stop)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,V", [(2, 3), (2, 4), (3, 2)])
def test_sharded_warp_and_conv1_reduce_gloo(world, V):
    C = 4
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, V, C, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert ret["gather_equal"] is True            # rank-order concat of the slices == the reference concat
        assert ret["conv_err"] <= 1e-5                # sum of per-shard conv1 partials == unsharded conv1
        assert ret["slice"][0] == 0


def test_row_bands_are_whole_tile_rows_and_cover_the_grid():
    """Row ownership of the fused reduce-scatter (sharding.row_bands): bands of whole 8-row tile rows, balanced, covering [0, Hg)."""
    from mvdet_b200.sharding import row_bands
    assert row_bands(480, 8) == [0, 64, 128, 192, 256, 312, 368, 424, 480]      # 60 tile rows: 8,8,8,8,7,7,7,7
    assert row_bands(120, 8) == [0, 16, 32, 48, 64, 80, 96, 112, 120]           # 15 tile rows
    assert row_bands(24, 3) == [0, 8, 16, 24]
    assert row_bands(20, 4) == [0, 8, 16, 20, 20]                               # 3 tile rows on 4 ranks: the last owns nothing
    for Hg in (1, 7, 8, 9, 120, 160, 480, 1000):
        for world in (1, 2, 3, 8):
            e = row_bands(Hg, world)
            assert len(e) == world + 1 and e[0] == 0 and e[-1] == Hg
            assert all(b >= a for a, b in zip(e[:-1], e[1:]))
            assert all(a % 8 == 0 for a in e[:-1] if a < Hg)
            sizes = [(b - a + 7) // 8 for a, b in zip(e[:-1], e[1:])]
            assert max(sizes) - min(sizes) <= 1
