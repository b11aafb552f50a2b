"""tcgen05 / TMEM building blocks (north_star (c)): the validation GEMM against torch.matmul."""
import ctypes as C

import pytest
import torch

from mvdet_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K", [(128, 64), (256, 192), (1024, 512)])
def test_tcgen05_gemm_matches_matmul(M, K):
    lib = _lib.load_dev()
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g).bfloat16().cuda()
    B = torch.randn(256, K, generator=g).bfloat16().cuda()
    D = torch.full((M, 256), float("nan"), device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mvwarp_debug_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), M, 256, K, stream),
               "mvwarp_debug_gemm")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = float((D - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err          # bf16 products are exact in fp32; only the summation order differs


def _conv1_reference(x_bf16, pm, grid_hw, weight, bias, relu):
    """conv3x3(bf16(warp_views_oracle(feats))) with bf16-rounded weights, fp32 accumulate -- what the fused kernel computes."""
    import torch.nn.functional as F
    from oracle import kornia_warp as K
    X = K.warp_views_oracle(x_bf16.float(), pm, grid_hw).bfloat16().float()
    out = F.conv2d(X, weight.bfloat16().float(), bias, padding=1)
    return F.relu(out) if relu else out


@pytest.mark.parametrize("V,Cc,relu", [(1, 64, False), (2, 128, True)])
def test_fused_warp_conv1_matches_reference(V, Cc, relu):
    from mvdet_b200 import conv1
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=31)      # 54x96 -> 24x72: partial tiles in x
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, V, Cc, *src_hw, generator=g).bfloat16()
    weight = torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05
    bias = torch.randn(512, generator=g)
    ref = _conv1_reference(x, pm, grid_hw, weight, bias, relu)
    wpack = conv1.pack_conv1_weight(weight, V, Cc).cuda()
    out = conv1.warp_conv1(to_channels_last_5d(x).cuda(), pm, grid_hw, wpack, bias, relu=relu)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    emax, el2 = rel_errors(out, ref)
    assert emax <= 1e-4 and el2 <= 1e-4, (emax, el2)


def test_fused_conv1_view_shards_sum_to_the_full_convolution():
    """Two view shards (emulated one after the other on one GPU): the fused partial sums add up to conv1 of the concat."""
    from mvdet_b200 import conv1, sharding as S
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    V, Cc = 3, 64
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=33)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, V, Cc, *src_hw, generator=g).bfloat16()
    weight = torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05
    bias = torch.randn(512, generator=g)
    ref = _conv1_reference(x, pm, grid_hw, weight, bias, relu=True)
    total = None
    for rank in range(2):
        shard = S.ViewShard.for_rank(V, rank=rank, world_size=2)
        local = to_channels_last_5d(x[:, shard.first_view:shard.first_view + shard.count]).cuda()
        wp = S.pack_conv1_weight_shard(weight, shard, Cc).cuda()
        part = S.conv1_partial_fused(local, pm, grid_hw, wp, bias.cuda(), shard)
        total = part if total is None else total + part
    out = torch.relu(total)
    emax, el2 = rel_errors(out, ref)
    assert emax <= 1e-4 and el2 <= 1e-4, (emax, el2)


def test_fused_conv1_rejects_what_it_cannot_run():
    from mvdet_b200 import conv1
    from tests.util import small_rig
    pm, src_hw, grid_hw = small_rig(num_cam=1, seed=31)
    x = torch.randn(1, 1, 64, *src_hw).cuda()                         # fp32: not supported
    w = conv1.pack_conv1_weight(torch.randn(512, 66, 3, 3), 1, 64).cuda()
    with pytest.raises(ValueError):
        conv1.warp_conv1(x, pm, grid_hw, w)
    with pytest.raises(ValueError):
        conv1.pack_conv1_weight(torch.randn(512, 50, 3, 3), 1, 48)   # C % 64 != 0


@pytest.mark.parametrize("V,Cc", [(2, 64), (1, 512)])
def test_fused_conv1_trainable_gradients(V, Cc):
    """warp_conv1_trainable: fused forward, unfused backward -- gradients against autograd through the oracle."""
    import torch.nn.functional as F
    from mvdet_b200 import conv1
    from oracle import kornia_warp as K
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=35)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, V, Cc, *src_hw, generator=g).bfloat16()
    weight = (torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05)
    bias = torch.randn(512, generator=g)
    go = torch.randn(1, 512, *grid_hw, generator=g)
    # reference: fp32 autograd through oracle warp (on bf16-rounded inputs / weights) + conv2d + relu
    xr = x.float().requires_grad_(True)
    wr = weight.bfloat16().float().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    ref = F.relu(F.conv2d(K.warp_views_oracle(xr, pm, grid_hw), wr, br, padding=1))
    gxr, gwr, gbr = torch.autograd.grad(ref, [xr, wr, br], go)
    xd = to_channels_last_5d(x).cuda().requires_grad_(True)
    wd = weight.cuda().requires_grad_(True)
    bd = bias.cuda().requires_grad_(True)
    out = conv1.warp_conv1_trainable(xd, wd, bd, pm, grid_hw)
    gx, gw, gb = torch.autograd.grad(out, [xd, wd, bd], go.cuda())
    # the reference keeps the warped tensor and every gradient in fp32; the CUDA path rounds the warped tensor, grad_out,
    # the convolution's input gradient and the feature gradient to bf16 (4 roundings of 2^-9 each, plus ReLU-mask flips
    # where the pre-activation is within that noise of zero): agreement is at the few-percent level in l2
    assert max(rel_errors(out, ref)) <= 2e-2
    assert rel_errors(gx.float(), gxr)[1] <= 5e-2
    assert rel_errors(gw, gwr)[1] <= 5e-2
    assert rel_errors(gb, gbr)[1] <= 5e-2


# ---- the rest of map_classifier ---------------------------------------------------------
@pytest.mark.parametrize("B,Cin,H,W,d,relu,odt", [(1, 64, 8, 16, 1, False, torch.float32), (2, 128, 24, 72, 2, True, torch.bfloat16),
                                                  (1, 512, 30, 45, 2, True, torch.float32), (1, 64, 19, 33, 4, False, torch.float32)])
def test_tcgen05_tma_conv3x3_matches_conv2d(B, Cin, H, W, d, relu, odt):
    """conv3x3_cl (A operand by TMA boxes, zero padding by out-of-bounds fill) against F.conv2d in fp32 on the same
    bf16-rounded operands; partial tiles in both directions, dilation 1 / 2 / 4."""
    import torch.nn.functional as F
    from mvdet_b200 import map_head as MH
    from tests.util import rel_errors
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16()
    w = (torch.randn(512, Cin, 3, 3, generator=g) * 0.05)
    bias = torch.randn(512, generator=g)
    ref = F.conv2d(x.float(), w.bfloat16().float(), bias, padding=d, dilation=d)
    if relu:
        ref = F.relu(ref)
    out = MH.conv3x3_cl(x.cuda().contiguous(memory_format=torch.channels_last), MH.pack_conv3x3_weight(w).cuda(), bias, dilation=d,
                        relu=relu, out_dtype=odt)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and out.dtype == odt and out.stride(1) == 1
    emax, el2 = rel_errors(out.float(), ref)
    tol = 1e-4 if odt == torch.float32 else 6e-3      # bf16 output: one rounding of 2^-9 relative on each value
    assert emax <= tol and el2 <= tol, (emax, el2)


@pytest.mark.parametrize("two_pass", [True, False])
@pytest.mark.parametrize("B,Cin,H,W,d", [(2, 512, 21, 35, 4), (1, 96, 9, 70, 2), (3, 256, 17, 16, 1), (1, 40, 12, 20, 4)])
def test_conv3x3_head_matches_conv2d(two_pass, B, Cin, H, W, d):
    """map_classifier[4] (512 -> 1, dilation 4, no bias): the two-pass path (nine partial dot products per cell by mma.sync, then a
    9-term stencil; needs Cin % 32 == 0 -- Cin = 40 falls back) and the one-pass kernel against F.conv2d in fp32.  Cell counts that
    are no multiple of the 256-cell CTA / 16-cell MMA tile, Cin that is no multiple of the 128-channel trip."""
    import torch.nn.functional as F
    from mvdet_b200 import map_head as MH
    from tests.util import rel_errors
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16()
    w = torch.randn(1, Cin, 3, 3, generator=g) * 0.05
    ref = F.conv2d(x.float(), w.bfloat16().float(), None, padding=d, dilation=d)
    xd = x.cuda().contiguous(memory_format=torch.channels_last)
    out = MH.conv3x3_head(xd, MH.pack_head_weight(w).cuda(), dilation=d, two_pass=two_pass)
    assert out.shape == ref.shape
    emax, el2 = rel_errors(out, ref)
    assert emax <= 1e-5 and el2 <= 1e-5, (emax, el2)
    assert torch.equal(out, MH.conv3x3_head(xd, MH.pack_head_weight(w).cuda(), dilation=d, two_pass=two_pass))   # deterministic


def test_map_head_matches_the_reference_module():
    """MapHead (fused warp->conv1 -> TMA conv -> head) against the reference's map_classifier built as a torch module
    (persp_trans_detector.py:51-54) on the oracle's warp, with the same bf16 roundings of the activations."""
    import torch.nn as nn
    import torch.nn.functional as F
    from mvdet_b200.map_head import MapHead
    from oracle import kornia_warp as K
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    V, Cc = 2, 64
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=37)
    torch.manual_seed(0)
    ref_head = nn.Sequential(nn.Conv2d(Cc * V + 2, 512, 3, padding=1), nn.ReLU(),
                             nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                             nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False))
    sd = {"map_classifier." + k: v for k, v in ref_head.state_dict().items()}
    x = torch.randn(2, V, Cc, *src_hw, generator=torch.Generator().manual_seed(1)).bfloat16()
    bf = lambda t: t.bfloat16().float()
    with torch.no_grad():
        X = bf(K.warp_views_oracle(x.float(), pm, grid_hw))
        a1 = bf(F.relu(F.conv2d(X, bf(ref_head[0].weight), ref_head[0].bias, padding=1)))
        a2 = bf(F.relu(F.conv2d(a1, bf(ref_head[2].weight), ref_head[2].bias, padding=2, dilation=2)))
        ref = F.conv2d(a2, bf(ref_head[4].weight), None, padding=4, dilation=4)
    head = MapHead(sd, V, Cc, device="cuda:0")
    out = head(to_channels_last_5d(x).cuda(), pm, grid_hw)
    assert out.shape == ref.shape == (2, 1, *grid_hw)
    # the two paths round the same activations to bf16 after sums taken in different orders: values within half an ulp of a
    # rounding boundary flip, each flip is 2^-8 relative on one activation -> a few 1e-3 on the final map
    emax, el2 = rel_errors(out, ref)
    assert emax <= 1e-2 and el2 <= 5e-3, (emax, el2)


def test_operand_swizzle_follows_absolute_address():
    """The property warp_conv1_direct_kernel rests on: an A operand written with chunk ^= (address >> 7) & 7 is
    read correctly from any 128-byte-aligned start and with any stride between its 8-row groups, base-offset field 0."""
    lib = _lib.load_dev()
    g = torch.Generator().manual_seed(0)
    M, K = 128, 128
    A = torch.randn(M, K, generator=g).bfloat16().cuda()
    B = torch.randn(256, K, generator=g).bfloat16().cuda()
    ref = A.float() @ B.float().T
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for sbo, start in [(1024, 0), (1024, 3), (1280, 0), (1280, 1), (1280, 11), (1280, 21), (1536, 5)]:
        D = torch.full((M, 256), float("nan"), device="cuda")
        _lib.check(lib.mvwarp_debug_gemm_layout(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), M, 256, K,
                                                sbo, start, 0, stream), "mvwarp_debug_gemm_layout")
        torch.cuda.synchronize()
        assert float((D - ref).abs().max() / ref.abs().max()) <= 1e-5, (sbo, start)
    D = torch.empty(M, 256, device="cuda")
    assert lib.mvwarp_debug_gemm_layout(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), M, 256, K, 1000, 0, 0,
                                        stream) == _lib.EINVAL


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_fused_conv1_kernel_generations_agree(mode, monkeypatch):
    """Single-CTA kernel (0), CTA pairs with copied tap stages (1) and the default direct kernel (2) against the same reference,
    on a shape with an odd tile count (padding CTA in pair mode), partial tiles, the coord block and a K-split tail; bf16 output too."""
    from mvdet_b200 import conv1
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    if mode != "2":
        # the superseded generations live in the development build, whose knobs are read once per process: run those in a child
        import subprocess, sys, os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        from mvdet_b200.build import build_library
        env = dict(os.environ, MVWARP_CONV1_PAIR=mode, MVWARP_LIB=build_library(dev=True), PYTHONPATH=root)
        code = ("import pytest, sys; sys.exit(pytest.main(['-q', '-x', '-m', 'gpu', "
                "'tests/test_gpu_umma.py::test_fused_conv1_kernel_generations_agree[2]']))")
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=900)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return
    V, Cc = 3, 64
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=41, grid=(24, 72))     # 3 x 5 tiles = 15 per frame
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, V, Cc, *src_hw, generator=g).bfloat16()
    weight = torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05
    bias = torch.randn(512, generator=g)
    ref = _conv1_reference(x, pm, grid_hw, weight, bias, relu=True)
    wpack = conv1.pack_conv1_weight(weight, V, Cc).cuda()
    xd = to_channels_last_5d(x).cuda()
    out = conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True)
    emax, el2 = rel_errors(out, ref)
    assert emax <= 1e-4 and el2 <= 1e-4, (mode, emax, el2)
    out16 = conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True, out_dtype=torch.bfloat16)
    assert out16.dtype == torch.bfloat16 and torch.equal(out16, out.bfloat16())
    again = conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True)
    assert torch.equal(again, out)                                         # deterministic (fixed K-split reduce order)
    rows = conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True, rows=(8, 24))
    emax, _ = rel_errors(rows[:, :, 8:24], ref[:, :, 8:24])
    assert emax <= 1e-4


def test_fused_conv1_p2p_epilogue_single_rank():
    """mvwarp_warp_conv1_fwd_p2p with this process as two of three 'ranks' of the row ownership (the inbox pointers all point into
    one local buffer): the raw partial sums land in the right slot / band / cell, and slot sums reproduce the ordinary output."""
    from mvdet_b200 import conv1
    from mvdet_b200.plan import get_plan
    from mvdet_b200.sharding import row_bands
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    V, Cc, B = 2, 64, 2
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=43)                   # 24 x 72 grid: 3 tile rows
    Hg, Wg = grid_hw
    g = torch.Generator().manual_seed(0)
    x = to_channels_last_5d(torch.randn(B, V, Cc, *src_hw, generator=g).bfloat16()).cuda()
    weight = torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05
    wpack = conv1.pack_conv1_weight(weight, V, Cc).cuda()
    ref = conv1.warp_conv1(x, pm, grid_hw, wpack, None, relu=False)       # [B,512,Hg,Wg] fp32, no bias
    world = 3
    edges = row_bands(Hg, world)
    assert edges == [0, 8, 16, 24]
    band_max = 8
    inboxes = torch.full((world, world, B, band_max, Wg, 512), float("nan"), device="cuda")   # [owner][src][b][y][x][c]
    plan = get_plan(pm, src_hw, grid_hw, "cuda:0")
    lib = _lib.load()
    ptrs = (C.c_uint64 * world)(*[inboxes[o].data_ptr() for o in range(world)])
    rows = (C.c_int32 * (world + 1))(*edges)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    src = 1
    _lib.check(lib.mvwarp_warp_conv1_fwd_p2p(plan.handle, 0, V, C.c_void_p(x.data_ptr()), _lib.i64x(x.stride()), C.c_void_p(wpack.data_ptr()),
                                             B, Cc, 1, world, src, ptrs, rows, band_max, stream), "mvwarp_warp_conv1_fwd_p2p")
    torch.cuda.synchronize()
    for o in range(world):
        got = inboxes[o, src, :, :edges[o + 1] - edges[o]].permute(0, 3, 1, 2)      # [B,512,rows,Wg]
        emax, _ = rel_errors(got, ref[:, :, edges[o]:edges[o + 1]])
        assert emax <= 1e-5, (o, emax)
        others = inboxes[o, [s for s in range(world) if s != src]]
        assert bool(others.isnan().all())                                           # nothing written outside slot `src`


@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_inbox_reduce_matches_torch(odt):
    """mvwarp_inbox_reduce: slots added in slot order + bias + ReLU in one pass, over the first `rows` rows of every batch element."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    nslots, B, band_max, rows, Wg = 3, 2, 8, 5, 9
    inbox = torch.randn(nslots, B, band_max, Wg, 512, generator=g).cuda()
    bias = torch.randn(512, generator=g).cuda()
    out = torch.empty(B, rows, Wg, 512, dtype=odt, device="cuda")
    b_pitch = band_max * Wg * 512
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mvwarp_inbox_reduce(C.c_void_p(inbox.data_ptr()), nslots, b_pitch * B, b_pitch, B, rows * Wg * 512, C.c_void_p(bias.data_ptr()), 1,
                                       C.c_void_p(out.data_ptr()), _lib.BF16 if odt == torch.bfloat16 else _lib.F32, 0, stream), "mvwarp_inbox_reduce")
    ref = ((inbox[0, :, :rows] + inbox[1, :, :rows]) + inbox[2, :, :rows] + bias).relu()
    assert torch.equal(out, ref.to(odt))
    assert lib.mvwarp_inbox_reduce(C.c_void_p(inbox.data_ptr()), nslots, b_pitch * B, b_pitch, B, 100, None, 0, C.c_void_p(out.data_ptr()),
                                   _lib.F32, 0, stream) == _lib.EINVAL


def test_fused_conv1_ragged_grid_and_many_views():
    """Default (direct) kernel on a grid that is no multiple of the 16 x 8 tile in either direction (20 x 60) with 5 views
    (K-split parts of unequal view counts) and B = 3 (odd item count: a padding CTA in the last pair)."""
    from mvdet_b200 import conv1
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    V, Cc, B = 5, 64, 3
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=47, grid=(20, 60))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, V, Cc, *src_hw, generator=g).bfloat16()
    weight = torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05
    bias = torch.randn(512, generator=g)
    ref = _conv1_reference(x, pm, grid_hw, weight, bias, relu=True)
    out = conv1.warp_conv1(to_channels_last_5d(x).cuda(), pm, grid_hw, conv1.pack_conv1_weight(weight, V, Cc).cuda(), bias, relu=True)
    assert out.shape == (B, 512, 20, 60)
    emax, el2 = rel_errors(out, ref)
    assert emax <= 1e-4 and el2 <= 1e-4, (emax, el2)


def test_patch_detector_fused_head_inference():
    """integration.patch_detector(model, fused_head=True): under no_grad the warp + map_classifier run as MapHead's three kernels and
    agree with the model's own un-fused forward to bf16 rounding; with autograd on the un-fused path runs and gradients flow."""
    import torch.nn as nn
    from mvdet_b200 import integration
    from tests.util import small_rig, rel_errors
    V, Cc = 2, 64
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=37)

    class Det(nn.Module):
        def __init__(self):
            super().__init__()
            self.num_cam = V
            self.base_pt1, self.base_pt2 = nn.Conv2d(3, Cc, 3, padding=1), nn.Sequential(nn.ReLU(), nn.Conv2d(Cc, Cc, 1))
            self.img_classifier = nn.Sequential(nn.Conv2d(Cc, 8, 1), nn.ReLU(), nn.Conv2d(8, 2, 1, bias=False))
            self.map_classifier = nn.Sequential(nn.Conv2d(Cc * V + 2, 512, 3, padding=1), nn.ReLU(),
                                                nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                                                nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False))
            self.proj_mats = [pm[v] for v in range(V)]
            self.upsample_shape, self.reducedgrid_shape = list(src_hw), list(grid_hw)

        def forward(self, imgs, visualize=False):
            raise RuntimeError("replaced by patch_detector")

    torch.manual_seed(0)
    model = integration.patch_detector(Det().cuda().eval(), fused_head=True)
    assert model.fused_map_head is not None
    imgs = torch.randn(2, V, 3, 18, 32, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        fused, res_f = model(imgs)
    plain, res_p = model(imgs)                              # autograd on: the un-fused path (cuDNN, fp32 / TF32)
    assert fused.shape == plain.shape == (2, 1, *grid_hw) and not fused.requires_grad and plain.requires_grad
    emax, el2 = rel_errors(fused, plain)
    assert emax <= 3e-2 and el2 <= 2e-2, (emax, el2)        # bf16 activations + weights against fp32
    assert all(torch.equal(a, b) for a, b in zip(res_f, res_p))
    plain.square().sum().backward()
    assert model.map_classifier[0].weight.grad is not None and model.base_pt1.weight.grad is not None

    # fused_conv1: the training path through warp_conv1_trainable agrees with the un-fused autograd path to bf16 rounding
    for p_ in model.parameters():
        p_.grad = None
    g_plain = torch.autograd.grad(model(imgs)[0].square().sum(), [model.map_classifier[0].weight, model.base_pt1.weight])
    trained = integration.patch_detector(model, fused_conv1=True)
    out_f, _ = trained(imgs)
    assert out_f.requires_grad and rel_errors(out_f, plain)[1] <= 2e-2
    g_fused = torch.autograd.grad(out_f.square().sum(), [model.map_classifier[0].weight, model.base_pt1.weight])
    for a, b in zip(g_fused, g_plain):
        assert rel_errors(a, b)[1] <= 6e-2, rel_errors(a, b)


# ---- parity at the benchmark's own sizes (VERDICT r01 weak #1 ii-iv) -------------------------------------------------
def _wildtrack_conv1_case(B, seed=0):
    from mvdet_b200.geometry import WILDTRACK, synthetic_rig
    spec = WILDTRACK
    pm = synthetic_rig(spec)
    g = torch.Generator().manual_seed(seed)
    # realistic smooth features (3x bilinear up-sample of noise, persp_trans_detector.py:65) keep the CPU side cheap to generate
    low = torch.randn(B * spec.num_cam, 512, 90, 160, generator=g)
    x = torch.nn.functional.interpolate(low, spec.feat_shape, mode="bilinear").view(B, spec.num_cam, 512, *spec.feat_shape).bfloat16()
    weight = torch.randn(512, spec.num_cam * 512 + 2, 3, 3, generator=g) * 0.02
    bias = torch.randn(512, generator=g)
    return spec, pm, x, weight, bias


@pytest.mark.parametrize("B", [1, 2])
def test_fused_conv1_at_the_wildtrack_shape(B):
    """warp_conv1 at V=7, C=512, 270x480 -> 120x360 (B=1: 345 tiles = 2.33 waves -> the last-wave K split over views and the
    56-block K loop run exactly as in bench.py; B=2: 690 tiles, CTA pairs + another split).  Reference: the CPU oracle's warp
    rounded to bf16, then conv3x3 in fp32 (TF32 off) on the GPU for all 512 output channels and on the CPU for 16 of them."""
    import torch.nn.functional as F
    from mvdet_b200 import conv1
    from oracle import kornia_warp as K
    from tests.util import to_channels_last_5d, rel_errors
    spec, pm, x, weight, bias = _wildtrack_conv1_case(B)
    X = torch.cat([K.warp_views_oracle(x[b:b + 1].float(), pm, spec.grid_shape) for b in range(B)]).bfloat16().float()
    w16 = weight.bfloat16().float()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.relu(F.conv2d(X.cuda(), w16.cuda(), bias.cuda(), padding=1))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    sel = torch.arange(0, 512, 32)                                   # 16 output channels on the CPU: ties the GPU reference down
    ref_cpu = F.relu(F.conv2d(X, w16[sel], bias[sel], padding=1))
    assert rel_errors(ref[:, sel], ref_cpu)[0] <= 2e-5
    wpack = conv1.pack_conv1_weight(weight, spec.num_cam, 512).cuda()
    xd = to_channels_last_5d(x).cuda()
    out = conv1.warp_conv1(xd, pm, spec.grid_shape, wpack, bias, relu=True)
    torch.cuda.synchronize()
    assert out.shape == (B, 512, *spec.grid_shape)
    emax, el2 = rel_errors(out, ref)
    assert emax <= 1e-4 and el2 <= 1e-4, (emax, el2)
    emax, el2 = rel_errors(out[:, sel], ref_cpu)
    assert emax <= 1e-4 and el2 <= 1e-4, (emax, el2)
    assert torch.equal(conv1.warp_conv1(xd, pm, spec.grid_shape, wpack, bias, relu=True), out)     # deterministic


def _rounded_reference_grads(x16, weight, bias, go, pm, grid_hw, mask):
    """The backward warp_conv1_trainable computes, restated on the CPU oracle with the SAME bf16 rounding points
    (warped tensor, grad_out, the convolution's input gradient, the feature gradient / weight gradient outputs) and the
    SAME ReLU mask, everything else in fp32: only summation order is left to differ."""
    import torch.nn.functional as F
    from oracle import kornia_warp as K
    bf = lambda t: t.bfloat16().float()
    xr = x16.float().requires_grad_(True)
    Xo = K.warp_views_oracle(xr, pm, grid_hw)
    X = bf(Xo.detach())
    w16 = bf(weight)
    g16 = bf(go * mask)
    gX = bf(torch.nn.grad.conv2d_input(X.shape, w16, g16, padding=1))
    gW = torch.nn.grad.conv2d_weight(X, w16.shape, g16, padding=1)        # the GEMM weight gradient returns fp32 (cuDNN's: bf16)
    gb = (go * mask).sum(dim=(0, 2, 3))
    (gx,) = torch.autograd.grad(Xo, xr, gX)
    return bf(gx), gW, gb


@pytest.mark.parametrize("V,Cc", [(2, 64), (1, 512)])
def test_fused_conv1_trainable_gradients_same_roundings(V, Cc):
    """Tight version of the gradient test: against a reference with the same four bf16 roundings and the same ReLU mask the
    gradients agree to 2e-3 (a wrong tap, a dropped view or a transposed weight block would be off by O(1))."""
    import torch.nn.functional as F
    from mvdet_b200 import conv1
    from oracle import kornia_warp as K
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=35)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, V, Cc, *src_hw, generator=g).bfloat16()
    weight = (torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05)
    bias = torch.randn(512, generator=g)
    go = torch.randn(1, 512, *grid_hw, generator=g)
    xd = to_channels_last_5d(x).cuda().requires_grad_(True)
    wd = weight.cuda().requires_grad_(True)
    bd = bias.cuda().requires_grad_(True)
    out = conv1.warp_conv1_trainable(xd, wd, bd, pm, grid_hw)
    gx, gw, gb = torch.autograd.grad(out, [xd, wd, bd], go.cuda())
    mask = (out.detach().cpu() > 0).float()
    ref_out = F.relu(F.conv2d(K.warp_views_oracle(x.float(), pm, grid_hw).bfloat16().float(), weight.bfloat16().float(), bias, padding=1))
    assert float(((ref_out > 0).float() == mask).float().mean()) >= 0.9999
    gxr, gwr, gbr = _rounded_reference_grads(x, weight, bias, go, pm, grid_hw, mask)
    assert max(rel_errors(out, ref_out)) <= 1e-4
    assert rel_errors(gx.float(), gxr)[1] <= 2e-3, rel_errors(gx.float(), gxr)
    assert rel_errors(gw, gwr)[1] <= 2e-3, rel_errors(gw, gwr)
    assert rel_errors(gb, gbr)[1] <= 1e-5, rel_errors(gb, gbr)


def test_patch_detector_fused_head_follows_weight_updates():
    """after an optimizer step / load_state_dict the no_grad (fused) forward must use the NEW weights."""
    import torch.nn as nn
    from mvdet_b200 import integration
    from tests.util import small_rig, rel_errors
    V, Cc = 2, 64
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=37)

    class Det(nn.Module):
        def __init__(self):
            super().__init__()
            self.num_cam = V
            self.base_pt1, self.base_pt2 = nn.Conv2d(3, Cc, 3, padding=1), nn.Sequential(nn.ReLU(), nn.Conv2d(Cc, Cc, 1))
            self.img_classifier = nn.Sequential(nn.Conv2d(Cc, 8, 1), nn.ReLU(), nn.Conv2d(8, 2, 1, bias=False))
            self.map_classifier = nn.Sequential(nn.Conv2d(Cc * V + 2, 512, 3, padding=1), nn.ReLU(),
                                                nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                                                nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False))
            self.proj_mats = [pm[v] for v in range(V)]
            self.upsample_shape, self.reducedgrid_shape = list(src_hw), list(grid_hw)

        def forward(self, imgs, visualize=False):
            raise RuntimeError("replaced by patch_detector")

    torch.manual_seed(0)
    model = integration.patch_detector(Det().cuda(), fused_head=True)
    imgs = torch.randn(1, V, 3, 18, 32, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        before, _ = model(imgs)
    opt = torch.optim.SGD(model.map_classifier.parameters(), lr=0.5)
    model(imgs)[0].square().sum().backward()
    opt.step()
    with torch.no_grad():
        after, _ = model(imgs)                              # fused path, must see the stepped weights
    plain, _ = model(imgs)                                  # autograd on: un-fused path on the live parameters
    assert rel_errors(after, plain.detach())[1] <= 2e-2, rel_errors(after, plain.detach())
    assert rel_errors(before, plain.detach())[1] > 5e-2     # the step really changed the output
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k in sd:
        if k.startswith("map_classifier.4"):
            sd[k] = sd[k] * 2
    model.load_state_dict(sd)                               # main.py:125
    with torch.no_grad():
        doubled, _ = model(imgs)
    assert rel_errors(doubled, 2 * after)[1] <= 1e-3


def test_conv1_training_backward_at_the_wildtrack_shape(monkeypatch):
    """The cuDNN-free backward of warp_conv1_trainable (fused ReLU/bf16/bias kernel, data gradient on the TMA tcgen05 convolution, weight
    gradient as batched GEMMs on shifted views) at the benchmark's own size -- V=7, C=512, 270x480 -> 120x360, B=1 -- against
    aten.convolution_backward (cuDNN) fed the same bf16 operands: same four rounding points, so only summation order differs."""
    from mvdet_b200 import conv1
    from tests.util import to_channels_last_5d, rel_errors
    spec, pm, x, weight, bias = _wildtrack_conv1_case(1, seed=3)
    xd = to_channels_last_5d(x).cuda().requires_grad_(True)
    wd = weight.cuda().requires_grad_(True)
    bd = bias.cuda().requires_grad_(True)
    go = torch.randn(1, *spec.grid_shape, 512, generator=torch.Generator().manual_seed(5)).permute(0, 3, 1, 2).cuda()
    wpack = conv1.pack_conv1_weight(wd, spec.num_cam, 512)

    def grads():
        out = conv1.warp_conv1_trainable(xd, wd, bd, pm, spec.grid_shape, wpack=wpack)
        return [g.float() for g in torch.autograd.grad(out, [xd, wd, bd], go)]

    gx, gw, gb = grads()                                              # own kernels + GEMMs
    monkeypatch.setenv("MVWARP_CUDNN_DGRAD", "1")
    monkeypatch.setenv("MVWARP_CUDNN_WGRAD", "1")
    rx, rw, rb = grads()                                              # cuDNN for both convolution gradients
    assert rel_errors(gx, rx)[1] <= 3e-3, rel_errors(gx, rx)          # both round the input gradient to bf16 after different sums
    assert rel_errors(gw, rw)[1] <= 3e-3, rel_errors(gw, rw)          # GEMM path: fp32 out; cuDNN: bf16 out (2^-9 per element)
    assert rel_errors(gb, rb)[1] <= 1e-6, rel_errors(gb, rb)
    monkeypatch.delenv("MVWARP_CUDNN_DGRAD"); monkeypatch.delenv("MVWARP_CUDNN_WGRAD")
    gx2, gw2, gb2 = grads()
    assert torch.equal(gx, gx2) and torch.equal(gb, gb2)              # deterministic where we own the kernels
    conv1.clear_caches()


# ---- fused conv1 over fused up-sample plans (VERDICT r01 "missing" #5 / "next" #7): F.interpolate of :65 inside the kernel -------
def _conv1_upsampled_reference(low_bf16, up_hw, pm, grid_hw, weight, bias, relu, device="cpu"):
    """conv3x3(bf16(warp_views_oracle(F.interpolate(low, up_hw, bilinear)))) with bf16-rounded weights, fp32 accumulate: the
    reference's persp_trans_detector.py:65, :68-77 and map_classifier[0] on the bf16 low-res features, ONE rounding of the warped
    value (the up-sampled tensor stays fp32) -- what the fused kernel computes."""
    import torch.nn.functional as F
    from oracle import kornia_warp as K
    B, V, Cc, hl, wl = low_bf16.shape
    up = F.interpolate(low_bf16.float().reshape(B * V, Cc, hl, wl), tuple(up_hw), mode="bilinear").reshape(B, V, Cc, *up_hw)
    X = torch.cat([K.warp_views_oracle(up[b:b + 1], pm, grid_hw) for b in range(B)]).bfloat16().float()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out = F.conv2d(X.to(device), weight.bfloat16().float().to(device), bias.to(device), padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return F.relu(out) if relu else out


# tolerance of the up-sampled variant: the kernel composes the two bilinear steps into one fp32 blend, the reference blends twice;
# the fp32 values differ by ~1e-6 relative, which flips the bf16 rounding of about one warped value in a thousand by one bf16 ulp
# (2^-8): a few 1e-4 of the output's scale after the K-sum.  A wrong or dropped tap shows up at >= 1e-2.
UP_TOL = (2e-3, 5e-4)


@pytest.mark.parametrize("V,Cc,low_hw,grid,B", [(2, 128, (18, 32), (24, 72), 2), (5, 64, (27, 48), (20, 60), 3), (1, 64, (54, 96), (24, 72), 1),
                                                (3, 64, (20, 37), (24, 72), 1)])
def test_fused_conv1_upsampled_matches_two_step_reference(V, Cc, low_hw, grid, B):
    """warp_conv1(upsample_shape=): 3x, 2x, 1:1 and a non-integer ratio; ragged grid + five views (unequal K-split parts) + odd item count."""
    from mvdet_b200 import conv1
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=53, grid=grid)
    g = torch.Generator().manual_seed(0)
    low = torch.randn(B, V, Cc, *low_hw, generator=g).bfloat16()
    weight = torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05
    bias = torch.randn(512, generator=g)
    ref = _conv1_upsampled_reference(low, src_hw, pm, grid_hw, weight, bias, relu=True)
    wpack = conv1.pack_conv1_weight(weight, V, Cc).cuda()
    xd = to_channels_last_5d(low).cuda()
    out = conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True, upsample_shape=src_hw)
    torch.cuda.synchronize()
    assert out.shape == ref.shape == (B, 512, *grid_hw)
    emax, el2 = rel_errors(out, ref)
    assert emax <= UP_TOL[0] and el2 <= UP_TOL[1], (emax, el2)
    # deterministic; bf16 output = the rounded fp32 output; a band of rows equals the same rows of the whole map
    assert torch.equal(out, conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True, upsample_shape=src_hw))
    out16 = conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True, upsample_shape=src_hw, out_dtype=torch.bfloat16)
    assert rel_errors(out16, out)[0] <= 2 ** -8
    band = torch.zeros_like(out)
    conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True, upsample_shape=src_hw, out=band, rows=(8, 16))
    assert rel_errors(band[:, :, 8:16], out[:, :, 8:16])[0] <= 1e-5 and float(band[:, :, :8].abs().max()) == 0.0
    # 1:1 "up-sample": every window is a single texel, the result is the plain fused kernel's
    if tuple(low_hw) == tuple(src_hw):
        plain = conv1.warp_conv1(xd, pm, grid_hw, wpack, bias, relu=True)
        assert rel_errors(out, plain)[0] <= 1e-5


def test_fused_conv1_upsampled_rejects_a_down_sample():
    from mvdet_b200 import conv1
    from tests.util import small_rig, to_channels_last_5d
    pm, src_hw, grid_hw = small_rig(num_cam=1, seed=3)
    x = to_channels_last_5d(torch.zeros(1, 1, 64, src_hw[0] + 2, src_hw[1]).bfloat16()).cuda()
    wpack = conv1.pack_conv1_weight(torch.zeros(512, 66, 3, 3), 1, 64).cuda()
    with pytest.raises(ValueError):
        conv1.warp_conv1(x, pm, grid_hw, wpack, upsample_shape=src_hw)


def test_fused_conv1_upsampled_at_the_wildtrack_shape():
    """V=7, C=512, 90x160 -> (270x480) -> 120x360, B=1: the benchmark's own size (K split of the last wave, 56 patches per tile).
    Reference: CPU oracle warp of the fp32 up-sampled features, rounded to bf16, conv3x3 in fp32 on the GPU; and the plain fused
    kernel on the bf16-rounded up-sampled features (two roundings: looser)."""
    import torch.nn.functional as F
    from mvdet_b200 import conv1
    from mvdet_b200.geometry import WILDTRACK, synthetic_rig
    from tests.util import to_channels_last_5d, rel_errors
    spec = WILDTRACK
    pm = synthetic_rig(spec)
    g = torch.Generator().manual_seed(5)
    low = torch.randn(1, spec.num_cam, 512, 90, 160, generator=g).bfloat16()
    weight = torch.randn(512, spec.num_cam * 512 + 2, 3, 3, generator=g) * 0.02
    bias = torch.randn(512, generator=g)
    ref = _conv1_upsampled_reference(low, spec.feat_shape, pm, spec.grid_shape, weight, bias, relu=True, device="cuda")
    wpack = conv1.pack_conv1_weight(weight, spec.num_cam, 512).cuda()
    xd = to_channels_last_5d(low).cuda()
    out = conv1.warp_conv1(xd, pm, spec.grid_shape, wpack, bias, relu=True, upsample_shape=spec.feat_shape)
    torch.cuda.synchronize()
    emax, el2 = rel_errors(out, ref)
    assert emax <= UP_TOL[0] and el2 <= UP_TOL[1], (emax, el2)
    assert torch.equal(out, conv1.warp_conv1(xd, pm, spec.grid_shape, wpack, bias, relu=True, upsample_shape=spec.feat_shape))
    up16 = F.interpolate(xd[0], spec.feat_shape, mode="bilinear")[None]              # bf16 in, bf16 out: a second rounding
    two_step = conv1.warp_conv1(up16, pm, spec.grid_shape, wpack, bias, relu=True)
    assert rel_errors(out, two_step)[1] <= 5e-3


def test_upsampled_map_head_and_detector_paths(monkeypatch):
    """MapHead(upsample_shape=) and patch_detector(fuse_upsample=True, fused_head=True / fused_conv1=True): no 512-channel
    F.interpolate on the map branch, in inference and in training; results against the model's own un-fused forward."""
    import torch.nn as nn
    from mvdet_b200 import integration
    from tests.util import small_rig, rel_errors
    V, Cc = 2, 64
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=37)

    class Det(nn.Module):
        def __init__(self):
            super().__init__()
            self.num_cam = V
            self.base_pt1, self.base_pt2 = nn.Conv2d(3, Cc, 3, padding=1), nn.Sequential(nn.ReLU(), nn.Conv2d(Cc, Cc, 1))
            self.img_classifier = nn.Sequential(nn.Conv2d(Cc, 8, 1), nn.ReLU(), nn.Conv2d(8, 2, 1, bias=False))
            self.map_classifier = nn.Sequential(nn.Conv2d(Cc * V + 2, 512, 3, padding=1), nn.ReLU(),
                                                nn.Conv2d(512, 512, 3, padding=2, dilation=2), nn.ReLU(),
                                                nn.Conv2d(512, 1, 3, padding=4, dilation=4, bias=False))
            self.proj_mats = [pm[v] for v in range(V)]
            self.upsample_shape, self.reducedgrid_shape = list(src_hw), list(grid_hw)

        def forward(self, imgs, visualize=False):
            raise RuntimeError("replaced by patch_detector")

    torch.manual_seed(0)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    base = Det().cuda().eval()
    imgs = torch.randn(2, V, 3, 18, 32, generator=torch.Generator().manual_seed(1))     # trunk output 18x32, up-sampled to 54x96
    plain_model = integration.patch_detector(base)                                       # un-fused: F.interpolate + warp_views + cuDNN
    plain, res_p = plain_model(imgs)
    g_plain = torch.autograd.grad(plain.square().sum(), [base.map_classifier[0].weight, base.base_pt1.weight])
    model = integration.patch_detector(base, fuse_upsample=True, fused_head=True)
    with torch.no_grad():
        fused, res_f = model(imgs)
    assert fused.shape == plain.shape == (2, 1, *grid_hw)
    emax, el2 = rel_errors(fused, plain)
    assert emax <= 3e-2 and el2 <= 2e-2, (emax, el2)        # bf16 activations + weights against fp32
    for a, b in zip(res_f, res_p):                           # img_classifier: up-sample after the 1x1 convolution (commutes)
        assert rel_errors(a, b)[0] <= 1e-4
    trained = integration.patch_detector(base, fuse_upsample=True, fused_conv1=True)
    out_f, _ = trained(imgs)
    assert out_f.requires_grad and rel_errors(out_f, plain)[1] <= 2e-2
    g_fused = torch.autograd.grad(out_f.square().sum(), [base.map_classifier[0].weight, base.base_pt1.weight])
    for a, b in zip(g_fused, g_plain):
        assert rel_errors(a, b)[1] <= 6e-2, rel_errors(a, b)


@pytest.mark.parametrize("V,Cc", [(2, 64), (1, 512)])
def test_fused_conv1_trainable_upsampled_gradients_same_roundings(V, Cc):
    """warp_conv1_trainable(upsample_shape=): forward AND backward on the trunk's low-res map.  Reference: autograd through
    F.interpolate -> oracle warp (fp32), with the training path's bf16 roundings (warped value, weight, masked gradient, data gradient,
    feature gradient) and the same ReLU mask.  C = 512 takes the tcgen05 data gradient; C = 64 the comparison path.  The low-res
    feature gradient, the weight gradient (re-materialised warp through the composed forward CSR) and the bias gradient must agree
    to 2e-3 -- a wrong composed tap or a gradient left at high resolution would be off by O(1)."""
    import torch.nn.functional as F
    from mvdet_b200 import conv1
    from oracle import kornia_warp as K
    from tests.util import small_rig, to_channels_last_5d, rel_errors
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=35)
    low_hw = (src_hw[0] // 3, src_hw[1] // 3)
    g = torch.Generator().manual_seed(0)
    low = torch.randn(1, V, Cc, *low_hw, generator=g).bfloat16()
    weight = (torch.randn(512, V * Cc + 2, 3, 3, generator=g) * 0.05)
    bias = torch.randn(512, generator=g)
    go = torch.randn(1, 512, *grid_hw, generator=g)
    xd = to_channels_last_5d(low).cuda().requires_grad_(True)
    wd = weight.cuda().requires_grad_(True)
    bd = bias.cuda().requires_grad_(True)
    out = conv1.warp_conv1_trainable(xd, wd, bd, pm, grid_hw, upsample_shape=src_hw)
    gx, gw, gb = torch.autograd.grad(out, [xd, wd, bd], go.cuda())
    assert gx.shape == xd.shape                                           # the gradient comes back at LOW resolution
    mask = (out.detach().cpu() > 0).float()
    bf = lambda t: t.bfloat16().float()
    xr = low.float().requires_grad_(True)
    up = torch.stack([F.interpolate(xr[:, v], src_hw, mode="bilinear") for v in range(V)], dim=1)      # persp_trans_detector.py:65
    Xo = K.warp_views_oracle(up, pm, grid_hw)
    X, w16 = bf(Xo.detach()), bf(weight)
    ref_out = F.relu(F.conv2d(X, w16, bias, padding=1))
    assert float(((ref_out > 0).float() == mask).float().mean()) >= 0.999
    g16 = bf(go * mask)
    gX = bf(torch.nn.grad.conv2d_input(X.shape, w16, g16, padding=1))
    gwr = torch.nn.grad.conv2d_weight(X, w16.shape, g16, padding=1)
    gbr = (go * mask).sum(dim=(0, 2, 3))
    (gxr,) = torch.autograd.grad(Xo, xr, gX)
    assert rel_errors(out, ref_out)[1] <= 5e-4
    assert rel_errors(gx.float(), bf(gxr))[1] <= 2e-3, rel_errors(gx.float(), bf(gxr))
    assert rel_errors(gw, gwr)[1] <= 2e-3, rel_errors(gw, gwr)
    assert rel_errors(gb, gbr)[1] <= 1e-5, rel_errors(gb, gbr)
