"""Parity of the CUDA path (through the C ABI / the torch op) with the CPU oracle.  Needs a B200.

Tolerances: fp32  max|a-o|/max|o| <= 1e-5 and ||a-o||/||o|| <= 1e-5 on the output and
on the gradient; bf16  |a-o| <= 2^-8 |o| + 1e-6 max|o| against the fp32 oracle run on the bf16-rounded
inputs.  The forward is additionally expected to be BIT-IDENTICAL to the oracle in fp32 (same
coordinate chain, same FMA order); the fraction is asserted where stated.
"""
import pytest
import torch

import mvdet_b200 as mv
from mvdet_b200 import ops, plan as P
from mvdet_b200.geometry import MULTIVIEWX, RIG32, WILDTRACK, RigSpec, synthetic_rig
from oracle import kornia_warp as K
from tests.util import bit_equal_fraction, rel_errors, small_rig, to_channels_last_5d

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _feats(B, V, C, hw, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, V, C, *hw, generator=g).to(dtype)


def _layout(x, layout):
    return to_channels_last_5d(x) if layout == "cl" else x.contiguous()


@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX, RIG32], ids=lambda s: s.name)
def test_device_coordinates_bit_identical_to_oracle(spec):
    pm = synthetic_rig(spec)
    plan = P.get_plan(pm, spec.feat_shape, spec.grid_shape, DEV)
    dev_grid = plan.device_grid()
    og = K.sampling_grids(pm, spec.feat_shape, spec.grid_shape)
    assert torch.equal(dev_grid.view(torch.int32), og.view(torch.int32))
    assert plan.stats["inline_coords"] == 1


@pytest.mark.parametrize("layout", ["cl", "nchw"])
@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX], ids=lambda s: s.name)
def test_forward_parity_full_rig(spec, layout):
    """Full-size rig geometry (every cell / texel of the real shapes), C reduced so the oracle takes seconds."""
    pm = synthetic_rig(spec)
    B, C = 1, 16
    x = _feats(B, spec.num_cam, C, spec.feat_shape)
    ref = K.warp_views_oracle(x, pm, spec.grid_shape)
    out = mv.warp_views(_layout(x, layout).to(DEV), pm, spec.grid_shape)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert out.stride(1) == (1 if layout == "cl" else spec.grid_shape[0] * spec.grid_shape[1])
    emax, el2 = rel_errors(out, ref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)
    assert bit_equal_fraction(out, ref) == 1.0          # "bit-for-bit" (BASELINE.json configs[1])


@pytest.mark.parametrize("layout", ["cl", "nchw"])
@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX], ids=lambda s: s.name)
def test_backward_parity_full_rig(spec, layout):
    pm = synthetic_rig(spec)
    B, C = 1, 8
    x = _feats(B, spec.num_cam, C, spec.feat_shape)
    xo = x.clone().requires_grad_(True)
    ref = K.warp_views_oracle(xo, pm, spec.grid_shape)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (gref,) = torch.autograd.grad(ref, xo, go)

    xd = _layout(x, layout).to(DEV).requires_grad_(True)
    out = mv.warp_views(xd, pm, spec.grid_shape)
    god = go.to(DEV)
    if layout == "cl":
        god = god.contiguous(memory_format=torch.channels_last)
    (gx,) = torch.autograd.grad(out, xd, god)
    assert gx.shape == x.shape
    emax, el2 = rel_errors(gx, gref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)
    # determinism: a second backward is bitwise identical (fixed-order CSR gather, no atomics)
    out2 = mv.warp_views(xd, pm, spec.grid_shape)
    (gx2,) = torch.autograd.grad(out2, xd, god)
    assert torch.equal(gx, gx2)


def test_full_channel_width_matches_oracle_on_one_view():
    """C = 512 (the real width) on one Wildtrack camera, forward and backward."""
    spec = WILDTRACK
    pm = synthetic_rig(spec)[2:3]
    x = _feats(1, 1, 512, spec.feat_shape)
    xo = x.clone().requires_grad_(True)
    ref = K.warp_views_oracle(xo, pm, spec.grid_shape)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (gref,) = torch.autograd.grad(ref, xo, go)
    xd = to_channels_last_5d(x).to(DEV).requires_grad_(True)
    out = mv.warp_views(xd, pm, spec.grid_shape)
    assert bit_equal_fraction(out, ref) == 1.0
    (gx,) = torch.autograd.grad(out, xd, go.to(DEV).contiguous(memory_format=torch.channels_last))
    emax, el2 = rel_errors(gx, gref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)


@pytest.mark.parametrize("layout", ["cl", "nchw"])
def test_bf16_within_documented_bound(layout):
    pm, src_hw, grid_hw = small_rig(num_cam=3)
    B, V, C = 2, 3, 64
    x = _feats(B, V, C, src_hw).bfloat16()
    ref = K.warp_views_oracle(x.float(), pm, grid_hw)          # fp32 oracle on the bf16-rounded inputs
    out = mv.warp_views(_layout(x, layout).to(DEV), pm, grid_hw)
    assert out.dtype == torch.bfloat16
    bound = 2.0 ** -8 * ref.abs() + 1e-6 * ref.abs().max()
    assert bool(((out.float().cpu() - ref).abs() <= bound).all())
    # gradient: same bound against the fp32 oracle gradient of bf16-rounded grad_out
    xo = x.float().requires_grad_(True)
    ref2 = K.warp_views_oracle(xo, pm, grid_hw)
    go = torch.randn(ref2.shape, generator=torch.Generator().manual_seed(1)).bfloat16()
    (gref,) = torch.autograd.grad(ref2, xo, go.float())
    xd = _layout(x, layout).to(DEV).requires_grad_(True)
    outd = mv.warp_views(xd, pm, grid_hw)
    god = go.to(DEV)
    if layout == "cl":
        god = god.contiguous(memory_format=torch.channels_last)
    (gx,) = torch.autograd.grad(outd, xd, god)
    bound = 2.0 ** -8 * gref.abs() + 1e-6 * gref.abs().max()
    assert bool(((gx.float().cpu() - gref).abs() <= bound).all())


@pytest.mark.parametrize("B,V,C", [(1, 1, 1), (2, 3, 3), (3, 2, 20), (1, 4, 132), (2, 1, 260)])
@pytest.mark.parametrize("layout", ["cl", "nchw"])
def test_ragged_shapes_forward_backward(B, V, C, layout):
    """Channel counts off the vector width (generic kernel), batch > 1, single view."""
    pm, src_hw, grid_hw = small_rig(num_cam=V, seed=11 + V)
    x = _feats(B, V, C, src_hw)
    xo = x.clone().requires_grad_(True)
    ref = K.warp_views_oracle(xo, pm, grid_hw)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (gref,) = torch.autograd.grad(ref, xo, go)
    xd = _layout(x, layout).to(DEV).requires_grad_(True)
    out = mv.warp_views(xd, pm, grid_hw)
    assert bit_equal_fraction(out, ref) == 1.0
    (gx,) = torch.autograd.grad(out, xd, go.to(DEV))
    emax, el2 = rel_errors(gx, gref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)


def test_non_contiguous_view_of_larger_buffer():
    pm, src_hw, grid_hw = small_rig(num_cam=2)
    big = _feats(2, 2, 24, (src_hw[0] + 3, src_hw[1] + 5)).to(DEV)
    x = big[:, :, 4:20, 1:1 + src_hw[0], 2:2 + src_hw[1]]
    assert not x.is_contiguous()
    ref = K.warp_views_oracle(x.cpu().contiguous(), pm, grid_hw)
    out = mv.warp_views(x, pm, grid_hw)
    assert bit_equal_fraction(out, ref) == 1.0


def test_identity_and_translation_properties():
    x = _feats(2, 1, 8, (17, 23))
    eye = torch.eye(3, dtype=torch.float64)[None]
    out = mv.warp_views(x.to(DEV), eye, (17, 23), with_coord_map=False)
    assert float((out.cpu() - x[:, 0]).abs().max()) < 2e-5
    shift = torch.tensor([[[1.0, 0, 3], [0, 1, 2], [0, 0, 1]]], dtype=torch.float64)
    out = mv.warp_views(x.to(DEV), shift, (17, 23), with_coord_map=False).cpu()
    expect = torch.zeros_like(x[:, 0])
    expect[..., 2:, 3:] = x[:, 0, :, :-2, :-3]
    assert float((out - expect).abs().max()) < 5e-5


def test_adjoint_identity_at_full_size():
    """<warp(x), g> == <x, warp^T(g)> at the real Wildtrack size (C=512, 7 views): oracle-free check of bwd."""
    spec = WILDTRACK
    pm = synthetic_rig(spec)
    x = torch.randn(1, spec.num_cam, spec.feat_shape[0], spec.feat_shape[1], 512, device=DEV).permute(0, 1, 4, 2, 3)
    x.requires_grad_(True)
    out = mv.warp_views(x, pm, spec.grid_shape)
    g = torch.randn(1, spec.grid_shape[0], spec.grid_shape[1], out.shape[1], device=DEV).permute(0, 3, 1, 2)
    (gx,) = torch.autograd.grad(out, x, g)
    lhs = float((out[:, :-2].double() * g[:, :-2].double()).sum())
    rhs = float((x.detach().double() * gx.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * (out[:, :-2].double().norm() * g.double().norm()).item()
    # coord channels exact
    cm = mv.create_coord_map(spec.grid_shape)
    assert torch.equal(out[:, -2:].cpu(), cm)


def test_linearity_at_full_size():
    spec = MULTIVIEWX
    pm = synthetic_rig(spec)
    shape = (1, spec.num_cam, spec.feat_shape[0], spec.feat_shape[1], 128)
    a = torch.randn(shape, device=DEV).permute(0, 1, 4, 2, 3)
    b = torch.randn(shape, device=DEV).permute(0, 1, 4, 2, 3)
    fa = mv.warp_views(a, pm, spec.grid_shape, with_coord_map=False)
    fb = mv.warp_views(b, pm, spec.grid_shape, with_coord_map=False)
    fab = mv.warp_views(2 * a - 3 * b, pm, spec.grid_shape, with_coord_map=False)
    assert float((fab - (2 * fa - 3 * fb)).abs().max()) < 1e-4


def test_kornia_signature_seam_matches_oracle():
    """ops.warp_perspective == oracle warp_perspective for the exact call of persp_trans_detector.py:68-69."""
    spec = WILDTRACK
    pm = synthetic_rig(spec)
    B, C = 2, 12
    for cam in (0, 3):
        src = _feats(B, 1, C, spec.feat_shape, seed=cam)[:, 0]
        proj_mat = pm[cam].repeat([B, 1, 1]).float()
        ref = K.warp_perspective(src, proj_mat, list(spec.grid_shape))
        out = ops.warp_perspective(src.to(DEV), proj_mat.to(DEV), list(spec.grid_shape))
        assert out.shape == ref.shape
        assert bit_equal_fraction(out, ref) == 1.0


def test_table_mode_matches_grid_sample():
    """Explicit-grid plans reproduce F.grid_sample for an arbitrary grid (incl. out-of-range and NaN cells)."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    V, C, Hs, Ws, Hg, Wg = 2, 8, 20, 30, 9, 13
    grid = torch.rand(V, Hg, Wg, 2) * 2.6 - 1.3
    grid[0, 0, 0] = float("nan")
    grid[1, 2, 3, 0] = float("inf")
    x = _feats(1, V, C, (Hs, Ws))
    ref = torch.cat([F.grid_sample(x[:, v], grid[v:v + 1], mode="bilinear", padding_mode="zeros", align_corners=True)
                     for v in range(V)], dim=1)
    out = ops.grid_sample_views(to_channels_last_5d(x).to(DEV), grid).cpu()
    same = (out == ref) | (out.isnan() & ref.isnan())
    assert bool(same.all())


def test_errors_and_no_cpu_fallback():
    pm, src_hw, grid_hw = small_rig()
    x = _feats(1, 3, 4, src_hw)
    with pytest.raises(mv.MvwarpError):
        mv.warp_views(x, pm, grid_hw)                       # CPU tensor: no fallback
    with pytest.raises(ValueError):
        mv.warp_views(x.to(DEV)[0], pm, grid_hw)            # not 5-D
    with pytest.raises(ValueError):
        mv.warp_views(x.to(DEV), pm[:2], grid_hw)           # view count mismatch
    with pytest.raises(TypeError):
        mv.warp_views(x.to(DEV).half(), pm, grid_hw)
    with pytest.raises(TypeError):
        ops.warp_perspective(x[0].numpy(), pm[:1], grid_hw)
    with pytest.raises(ValueError):
        ops.warp_perspective(x.to(DEV)[0], pm[0], grid_hw)


def test_empty_batch_returns_empty_tensors():
    """B = 0: no launch, empty tensors of the right shapes, backward too.  (kornia 0.6.11's transform_points divides by the batch
    size, so the reference path raises ZeroDivisionError here; returning empty tensors is a harmless superset.)"""
    pm, src_hw, grid_hw = small_rig()
    x = to_channels_last_5d(_feats(1, 3, 4, src_hw)).to(DEV)[:0].requires_grad_(True)
    with pytest.raises(ZeroDivisionError):
        K.warp_views_oracle(torch.zeros(0, 3, 4, *src_hw), pm, grid_hw)
    out = mv.warp_views(x, pm, grid_hw)
    assert tuple(out.shape) == (0, 3 * 4 + 2, *grid_hw)
    (gx,) = torch.autograd.grad(out, x, torch.zeros_like(out))
    assert tuple(gx.shape) == tuple(x.shape)
    y = ops.warp_perspective(torch.zeros(0, 4, *src_hw, device=DEV), pm[:1].float().expand(0, 3, 3).to(DEV), grid_hw)
    assert tuple(y.shape) == (0, 4, *grid_hw)


def test_rig32_slice_forward_backward():
    """The 32-camera full-resolution rig (480x1440 grid, CSR rows up to 1360 taps), 4 of its views, small C."""
    spec = RIG32
    pm = synthetic_rig(spec)[[0, 9, 17, 30]]
    C = 4
    x = _feats(1, 4, C, spec.feat_shape)
    xo = x.clone().requires_grad_(True)
    ref = K.warp_views_oracle(xo, pm, spec.grid_shape)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (gref,) = torch.autograd.grad(ref, xo, go)
    xd = to_channels_last_5d(x).to(DEV).requires_grad_(True)
    out = mv.warp_views(xd, pm, spec.grid_shape)
    assert bit_equal_fraction(out, ref) == 1.0
    (gx,) = torch.autograd.grad(out, xd, go.to(DEV).contiguous(memory_format=torch.channels_last))
    emax, el2 = rel_errors(gx, gref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)


# ---- fused up-sample + warp ---------------------------------------------------------
def _two_step_oracle(x_low, pm, up_hw, grid_hw):
    import torch.nn.functional as F
    V = x_low.shape[1]
    up = torch.stack([F.interpolate(x_low[:, v], up_hw, mode="bilinear") for v in range(V)], dim=1)   # persp_trans_detector.py:65
    return K.warp_views_oracle(up, pm, grid_hw)


@pytest.mark.parametrize("layout", ["cl", "nchw"])
@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX], ids=lambda s: s.name)
def test_upsampled_plan_matches_two_step_reference(spec, layout):
    """warp_views_upsampled(feat[90x160]) == warp_views(F.interpolate(feat, [270,480])) at the real shapes."""
    pm = synthetic_rig(spec)
    B, C = 1, 8
    low_hw = (90, 160)
    x = _feats(B, spec.num_cam, C, low_hw)
    xo = x.clone().requires_grad_(True)
    ref = _two_step_oracle(xo, pm, spec.feat_shape, spec.grid_shape)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (gref,) = torch.autograd.grad(ref, xo, go)
    xd = _layout(x, layout).to(DEV).requires_grad_(True)
    out = mv.warp_views_upsampled(xd, spec.feat_shape, pm, spec.grid_shape)
    assert out.shape == ref.shape
    emax, el2 = rel_errors(out, ref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)
    assert torch.equal(out[:, -2:].cpu(), ref[:, -2:])                 # coord_map channels exact
    god = go.to(DEV)
    if layout == "cl":
        god = god.contiguous(memory_format=torch.channels_last)
    (gx,) = torch.autograd.grad(out, xd, god)
    assert gx.shape == x.shape
    emax, el2 = rel_errors(gx, gref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)
    out2 = mv.warp_views_upsampled(xd, spec.feat_shape, pm, spec.grid_shape)
    (gx2,) = torch.autograd.grad(out2, xd, god)
    assert torch.equal(gx, gx2) and torch.equal(out, out2)           # deterministic


def test_upsampled_plan_full_width_and_bf16():
    spec = WILDTRACK
    pm = synthetic_rig(spec)[1:3]
    low_hw = (90, 160)
    x = _feats(2, 2, 512, low_hw)
    ref = _two_step_oracle(x, pm, spec.feat_shape, spec.grid_shape)
    out = mv.warp_views_upsampled(to_channels_last_5d(x).to(DEV), spec.feat_shape, pm, spec.grid_shape)
    emax, el2 = rel_errors(out, ref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)
    xb = x.bfloat16()
    refb = _two_step_oracle(xb.float(), pm, spec.feat_shape, spec.grid_shape)
    outb = mv.warp_views_upsampled(to_channels_last_5d(xb).to(DEV), spec.feat_shape, pm, spec.grid_shape)
    bound = 2.0 ** -8 * refb.abs() + 1e-5 * refb.abs().max()
    assert bool(((outb.float().cpu() - refb).abs() <= bound).all())


def test_upsampled_plan_ragged_channels():
    pm, src_hw, grid_hw = small_rig(num_cam=2, seed=9)
    low_hw = (src_hw[0] // 3, src_hw[1] // 3)
    for C in (3, 20, 132):
        x = _feats(2, 2, C, low_hw)
        xo = x.clone().requires_grad_(True)
        ref = _two_step_oracle(xo, pm, src_hw, grid_hw)
        go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
        (gref,) = torch.autograd.grad(ref, xo, go)
        xd = to_channels_last_5d(x).to(DEV).requires_grad_(True)
        out = mv.warp_views_upsampled(xd, src_hw, pm, grid_hw)
        (gx,) = torch.autograd.grad(out, xd, go.to(DEV))
        assert max(rel_errors(out, ref)) <= TOL and max(rel_errors(gx, gref)) <= TOL


# ---- ablation variants through the kornia-signature seam ---------------------------
@pytest.mark.parametrize("C,strided", [(3, False), (1, True), (512, False)], ids=["image_proj_C3", "res_proj_C1", "no_joint_conv_C512"])
def test_ablation_variant_shapes(C, strided):
    """The calls pinned by tests/test_reference_integration.py::test_ablation_variants_call_the_same_seam:
    image_proj_variant.py:69 warps the 3-channel image, res_proj_variant.py:72 warps ``img_res[:, 1].unsqueeze(1)`` (one
    channel, a strided slice of a 2-channel tensor), no_joint_conv_variant.py:68 the 512-channel features; all NCHW."""
    spec = WILDTRACK
    pm = synthetic_rig(spec)
    B, cam = 2, 4
    if strided:
        full = _feats(B, 1, 2, spec.feat_shape, seed=2)[:, 0]            # img_res [B,2,H,W]
        src = full[:, 1].unsqueeze(1)
        src_dev = full.to(DEV)[:, 1].unsqueeze(1)
        assert not src_dev.is_contiguous()
    else:
        src = _feats(B, 1, C, spec.feat_shape, seed=2)[:, 0] if C <= 3 else _feats(1, 1, C, spec.feat_shape, seed=2)[:, 0]
        src_dev = src.to(DEV)
    Bc = src.shape[0]
    proj_mat = pm[cam].repeat([Bc, 1, 1]).float()
    so = src.clone().requires_grad_(True)
    ref = K.warp_perspective(so, proj_mat, list(spec.grid_shape))
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(4))
    (gref,) = torch.autograd.grad(ref, so, go)
    sd = src_dev.detach().requires_grad_(True)
    out = ops.warp_perspective(sd, proj_mat.to(DEV), list(spec.grid_shape))
    assert out.shape == ref.shape == (Bc, C, *spec.grid_shape)
    assert bit_equal_fraction(out, ref) == 1.0
    (gx,) = torch.autograd.grad(out, sd, go.to(DEV))
    emax, el2 = rel_errors(gx, gref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)


def test_image_proj_variant_full_resolution_grid():
    """C = 3 onto the un-reduced 480x1440 ground grid (SURVEY.md 8f rank 4's small-C edge at grid_reduce = 1)."""
    spec = RIG32
    pm = synthetic_rig(spec)[[3, 20]]
    x = _feats(1, 2, 3, spec.feat_shape, seed=6)
    ref = K.warp_views_oracle(x, pm, spec.grid_shape)
    out = mv.warp_views(x.to(DEV), pm, spec.grid_shape)
    assert bit_equal_fraction(out, ref) == 1.0


# ---- plan persistence -----------------------------------------------------------------
@pytest.mark.parametrize("kind", ["plain", "upsampled"])
def test_saved_plan_reproduces_the_built_plan(tmp_path, kind):
    pm, src_hw, grid_hw = small_rig()
    low = (src_hw[0] // 3, src_hw[1] // 3)
    if kind == "plain":
        plan = P.get_plan(pm, src_hw, grid_hw, DEV)
        feat_hw = src_hw
    else:
        plan = P.get_plan_upsampled(pm, low, src_hw, grid_hw, DEV)
        feat_hw = low
    path = str(tmp_path / "rig.mvwplan")
    plan.save(path)
    T = P.src_norm_trans_dst_norm(pm, src_hw, grid_hw)
    loaded = P.load_plan(path, DEV, expect_matrices=T)
    assert loaded.stats == plan.stats
    assert torch.equal(loaded.matrices(), T)
    x = to_channels_last_5d(_feats(2, 3, 24, feat_hw)).to(DEV).requires_grad_(True)
    a = ops.warp_views_with_plan(x, plan)
    b = ops.warp_views_with_plan(x, loaded)
    assert torch.equal(a, b)
    go = torch.randn_like(a)
    (ga,) = torch.autograd.grad(a, x, go)
    (gb,) = torch.autograd.grad(b, x, go)
    assert torch.equal(ga, gb)
    # wrong rig, corrupt payload
    with pytest.raises(mv.MvwarpError):
        P.load_plan(path, DEV, expect_matrices=T.flip(0))
    blob = bytearray(open(path, "rb").read())
    blob[-100] ^= 0x40
    bad = str(tmp_path / "bad.mvwplan")
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(mv.MvwarpError, match="checksum"):
        P.load_plan(bad, DEV)
    open(bad, "wb").write(bytes(blob[:len(blob) // 2]))
    with pytest.raises(mv.MvwarpError, match="truncated"):
        P.load_plan(bad, DEV)
    # a well-formed file with a VALID checksum but out-of-range indices must be refused too.  Poison the last CSR entry's
    # column (the entries of the last-carved map end the blob up to alignment padding) and re-seal the checksum.
    import struct
    raw = bytearray(open(path, "rb").read())
    d_bytes, = struct.unpack_from("<Q", raw, 80)                       # PlanFileHeader: d_bytes at 80, blob_fnv1a at 88
    start = len(raw) - d_bytes
    words = struct.unpack_from(f"<{d_bytes // 4}I", raw, start)
    # find some entry column: any word equal to a valid cell id followed by a float weight in (0, 1] -- use the known layout instead:
    from mvdet_b200 import _lib as L
    st = plan.stats
    grid = P.host_grid(T, grid_hw)
    _, _, _, cells, _ = P.host_csr(grid, src_hw, with_map=True, lowres_hw=low if kind == "upsampled" else None)
    first = int(cells[0])
    hit = None
    for i in range(0, len(words) - 1):                                 # first (col, weight) pair matching the CSR's first entry
        if words[i] == first and 0 < struct.unpack("<f", struct.pack("<I", words[i + 1]))[0] <= 1.0:
            hit = i
            break
    assert hit is not None
    struct.pack_into("<I", raw, start + 4 * hit, 0x7fffffff)           # column far outside the grid
    def fnv(buf):
        h = 1469598103934665603
        n8 = len(buf) // 8
        for w in struct.unpack_from(f"<{n8}Q", buf, 0):
            h = ((h ^ w) * 1099511628211) & 0xffffffffffffffff
        for b in buf[8 * n8:]:
            h = ((h ^ b) * 1099511628211) & 0xffffffffffffffff
        return h
    struct.pack_into("<Q", raw, 88, fnv(bytes(raw[start:])))
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(mv.MvwarpError, match="out of range|inconsistent"):
        P.load_plan(bad, DEV)


def test_plan_disk_cache(tmp_path):
    pm, src_hw, grid_hw = small_rig(seed=11)
    P.clear_plan_cache()
    p1 = P.get_plan(pm, src_hw, grid_hw, DEV, cache_dir=str(tmp_path))
    files = list(tmp_path.iterdir())
    assert len(files) == 1 and files[0].suffix == ".mvwplan"
    P.clear_plan_cache()
    p2 = P.get_plan(pm, src_hw, grid_hw, DEV, cache_dir=str(tmp_path))       # comes from the file
    assert p2 is not p1 and p2.stats == p1.stats
    x = to_channels_last_5d(_feats(1, 3, 8, src_hw)).to(DEV)
    assert torch.equal(ops.warp_views_with_plan(x, p1), ops.warp_views_with_plan(x, p2))
    files[0].write_bytes(b"garbage")                                        # a broken cache entry is rebuilt, not trusted
    P.clear_plan_cache()
    p3 = P.get_plan(pm, src_hw, grid_hw, DEV, cache_dir=str(tmp_path))
    assert p3.stats == p1.stats and files[0].stat().st_size > 1000
    P.clear_plan_cache()


# ---- shipped wiring (mvdet_b200.integration) on a duck-typed detector -------------------------------------------
class _TinyDetector(torch.nn.Module):
    """Same attribute names and forward as PerspTransDetector (persp_trans_detector.py:14-87), small enough for a test.
    Its own forward is the reference's loop: per-camera warp_perspective (seam 1), list append, cat with coord_map."""

    def __init__(self, pm, src_hw, grid_hw, C=8):
        super().__init__()
        from mvdet_b200.geometry import create_coord_map
        self.num_cam = pm.shape[0]
        self.base_pt1 = torch.nn.Conv2d(3, C, 3, padding=1)
        self.base_pt2 = torch.nn.Sequential(torch.nn.ReLU(), torch.nn.Conv2d(C, C, 1))
        self.img_classifier = torch.nn.Sequential(torch.nn.Conv2d(C, 4, 1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 1, bias=False))
        self.map_classifier = torch.nn.Sequential(torch.nn.Conv2d(C * self.num_cam + 2, 4, 3, padding=1), torch.nn.ReLU(),
                                                  torch.nn.Conv2d(4, 1, 3, padding=2, dilation=2))
        self.proj_mats = [pm[v] for v in range(self.num_cam)]
        self.upsample_shape, self.reducedgrid_shape = list(src_hw), list(grid_hw)
        self.coord_map = create_coord_map(grid_hw)

    def forward(self, imgs, visualize=False):
        import torch.nn.functional as F
        B = imgs.shape[0]
        world_features, imgs_result = [], []
        for cam in range(self.num_cam):
            f = self.base_pt2(self.base_pt1(imgs[:, cam].to(DEV)))
            f = F.interpolate(f, self.upsample_shape, mode="bilinear")
            imgs_result.append(self.img_classifier(f))
            proj_mat = self.proj_mats[cam].repeat([B, 1, 1]).float().to(DEV)
            world_features.append(ops.warp_perspective(f, proj_mat, self.reducedgrid_shape))
        world = torch.cat(world_features + [self.coord_map.repeat([B, 1, 1, 1]).to(DEV)], dim=1)
        return F.interpolate(self.map_classifier(world), self.reducedgrid_shape, mode="bilinear"), imgs_result


@pytest.mark.parametrize("channels_last", [False, True])
def test_patch_detector_matches_the_per_camera_loop(channels_last, monkeypatch):
    """Seam 2 as shipped code against seam 1 on the GPU: one warp_views call == V warp_perspective calls + cat + coord_map,
    forward (same bits in NCHW; to rounding when the model switches to channels-last convolutions) and parameter gradients."""
    import copy
    from mvdet_b200 import integration
    pm, src_hw, grid_hw = small_rig()
    torch.manual_seed(0)
    ref = _TinyDetector(pm, src_hw, grid_hw).to(DEV)
    new = integration.patch_detector(copy.deepcopy(ref), channels_last=channels_last)
    imgs = torch.randn(2, 3, 3, 18, 32, generator=torch.Generator().manual_seed(1))
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)     # cuDNN's channels-last algorithms otherwise run TF32 (3.6e-4 here)
    map_ref, img_ref = ref(imgs)
    map_new, img_new = new(imgs)
    assert map_new.shape == map_ref.shape == (2, 1, *grid_hw)
    if channels_last:
        emax, _ = rel_errors(map_new, map_ref)
        assert emax <= 1e-4, emax
    else:
        assert torch.equal(map_new, map_ref)
        assert all(torch.equal(a, b) for a, b in zip(img_new, img_ref))
    map_ref.square().sum().backward()
    map_new.square().sum().backward()
    for (name, a), (_, b) in zip(new.named_parameters(), ref.named_parameters()):
        if b.grad is None:
            assert a.grad is None, name
            continue
        emax, _ = rel_errors(a.grad, b.grad)
        assert emax <= 1e-4, (name, emax)


def test_patch_detector_fused_upsample_on_gpu(monkeypatch):
    """fuse_upsample=True on the GPU: warp_views_upsampled on the trunk's low-res output + img_classifier up-sampling after its 1x1
    convolution, against the reference order (interpolate 8 channels, per-camera warp, cat): forward and parameter gradients."""
    import copy
    from mvdet_b200 import integration
    pm, src_hw, grid_hw = small_rig()
    torch.manual_seed(0)
    ref = _TinyDetector(pm, src_hw, grid_hw).to(DEV)
    new = integration.patch_detector(copy.deepcopy(ref), channels_last=True, fuse_upsample=True)
    imgs = torch.randn(2, 3, 3, 18, 32, generator=torch.Generator().manual_seed(1))
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    map_ref, img_ref = ref(imgs)
    map_new, img_new = new(imgs)
    assert rel_errors(map_new, map_ref)[0] <= 1e-4
    assert all(rel_errors(a, b)[0] <= 1e-4 for a, b in zip(img_new, img_ref))
    (map_ref.square().sum() + sum(r.square().sum() for r in img_ref)).backward()
    (map_new.square().sum() + sum(r.square().sum() for r in img_new)).backward()
    for (name, a), (_, b) in zip(new.named_parameters(), ref.named_parameters()):
        emax, _ = rel_errors(a.grad, b.grad)
        assert emax <= 1e-4, (name, emax)


def test_forward_and_backward_are_cuda_graph_capturable():
    """The launches take the current stream, allocate only through torch and never synchronise once the plan is cached, so a
    forward + backward can be captured in a CUDA graph and replayed on new inputs (launch-bound B=1 loops)."""
    pm, src_hw, grid_hw = small_rig()
    plan = P.get_plan(pm, src_hw, grid_hw, DEV)
    x = to_channels_last_5d(_feats(1, 3, 8, src_hw)).to(DEV)
    go = torch.randn(1, 3 * 8 + 2, *grid_hw, generator=torch.Generator().manual_seed(2)).to(DEV).contiguous(memory_format=torch.channels_last)
    out = torch.empty_like(go)
    gx = torch.empty_like(x)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                           # warm-up outside the capture
        ops.warp_views_fwd_into(plan, x, out, 0, 0, 3 * 8)
        ops.warp_views_bwd_into(plan, go, gx, 0, 0)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.warp_views_fwd_into(plan, x, out, 0, 0, 3 * 8)
        ops.warp_views_bwd_into(plan, go, gx, 0, 0)
    x2 = to_channels_last_5d(_feats(1, 3, 8, src_hw) * 3.0 + 1.0).to(DEV)
    go2 = go * -2.0
    x.copy_(x2)
    go.copy_(go2)
    out.zero_()
    gx.zero_()
    graph.replay()
    torch.cuda.synchronize()
    ref_out = torch.empty_like(out)
    ref_gx = torch.empty_like(gx)
    ops.warp_views_fwd_into(plan, x2, ref_out, 0, 0, 3 * 8)
    ops.warp_views_bwd_into(plan, go2, ref_gx, 0, 0)
    assert torch.equal(out, ref_out) and torch.equal(gx, ref_gx)


def test_realistic_upsampled_features_at_the_wildtrack_shape():
    """SURVEY.md 8(d)'s "realistic" inputs: features that are a 3x bilinear up-sample of a low-res map (persp_trans_detector.py:65)
    instead of white noise, two Wildtrack cameras at full size: forward bit-identical, gradient within 1e-5, bf16 inside its bound."""
    import torch.nn.functional as F
    spec = WILDTRACK
    pm = synthetic_rig(spec)[[1, 4]]
    Hs, Ws = spec.feat_shape
    V, C = 2, 32
    low = torch.randn(V, C, Hs // 3, Ws // 3, generator=torch.Generator().manual_seed(7))
    x = F.interpolate(low, (Hs, Ws), mode="bilinear")[None]             # [1,V,C,Hs,Ws]
    xo = x.clone().requires_grad_(True)
    ref = K.warp_views_oracle(xo, pm, spec.grid_shape)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (gref,) = torch.autograd.grad(ref, xo, go)
    xd = to_channels_last_5d(x).to(DEV).requires_grad_(True)
    out = mv.warp_views(xd, pm, spec.grid_shape)
    assert bit_equal_fraction(out, ref) == 1.0
    (gx,) = torch.autograd.grad(out, xd, go.to(DEV).contiguous(memory_format=torch.channels_last))
    emax, el2 = rel_errors(gx, gref)
    assert emax <= TOL and el2 <= TOL, (emax, el2)
    xb = x.bfloat16()
    refb = K.warp_views_oracle(xb.float(), pm, spec.grid_shape)
    outb = mv.warp_views(to_channels_last_5d(xb).to(DEV), pm, spec.grid_shape)
    bound = 2.0 ** -8 * refb.abs() + 1e-6 * refb.abs().max()
    assert bool(((outb.float().cpu() - refb).abs() <= bound).all())


# ---- list variant: one tensor per camera, no stack copy (VERDICT r01 missing #3) -------------------------------------
@pytest.mark.parametrize("layout", ["cl", "nchw", "mixed"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_warp_views_list_equals_warp_views(layout, dtype):
    """warp_views_list(per-camera tensors) == warp_views(stacked tensor), bit for bit, forward and backward; channels-last and
    NCHW cameras go through ONE launch (pointer table in the kernel parameters), mixed formats through per-view launches."""
    pm, src_hw, grid_hw = small_rig(num_cam=4, seed=61)
    B, V, Cc = 2, 4, 40
    x = _feats(B, V, Cc, src_hw).to(dtype)
    cams = []
    for v in range(V):
        t = x[:, v].to(DEV)
        if layout == "cl" or (layout == "mixed" and v % 2 == 0):
            t = t.contiguous(memory_format=torch.channels_last)
        else:
            t = t.contiguous()
        cams.append(t.requires_grad_(True))
    out = ops.warp_views_list(cams, pm, grid_hw)
    stacked = (to_channels_last_5d(x) if layout == "cl" else x.contiguous()).to(DEV).requires_grad_(True)
    ref = ops.warp_views(stacked, pm, grid_hw)
    assert out.shape == ref.shape and torch.equal(out, ref)
    assert (out.stride(1) == 1) == (layout != "nchw")
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3)).to(dtype).to(DEV)
    grads = torch.autograd.grad(out, cams, go)
    (gref,) = torch.autograd.grad(ref, stacked, go)
    for v in range(V):
        assert grads[v].shape == cams[v].shape
        assert torch.equal(grads[v], gref[:, v]), (layout, v)
    # only some cameras need a gradient (frozen trunk for the others): autograd drops the rest
    cams2 = [c.detach().requires_grad_(v == 1) for v, c in enumerate(cams)]
    (g1,) = torch.autograd.grad(ops.warp_views_list(cams2, pm, grid_hw), [cams2[1]], go)
    assert torch.equal(g1, gref[:, 1])


def test_warp_views_list_upsampled_and_errors():
    pm, src_hw, grid_hw = small_rig(num_cam=3, seed=63)
    low = (src_hw[0] // 3, src_hw[1] // 3)
    x = _feats(1, 3, 32, low)
    cams = [x[:, v].to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for v in range(3)]
    out = ops.warp_views_list(cams, pm, grid_hw, upsample_shape=src_hw)
    xs = to_channels_last_5d(x).to(DEV).requires_grad_(True)
    ref = ops.warp_views_upsampled(xs, src_hw, pm, grid_hw)
    assert torch.equal(out, ref)
    go = torch.randn_like(ref)
    grads = torch.autograd.grad(out, cams, go)
    (gref,) = torch.autograd.grad(ref, xs, go)
    assert all(torch.equal(grads[v], gref[:, v]) for v in range(3))
    with pytest.raises(ValueError):
        ops.warp_views_list(cams[:2], pm, grid_hw)                   # 2 tensors, 3 matrices
    with pytest.raises(ValueError):
        ops.warp_views_list([cams[0], cams[1][:, :16], cams[2]], pm, grid_hw)
    with pytest.raises(mv.MvwarpError):
        ops.warp_views_list([c.detach().cpu() for c in cams], pm, grid_hw)


def test_patched_detector_uses_no_stack_copy(monkeypatch):
    """integration.warped_forward hands the per-camera tensors to warp_views_list: torch.stack must not be called on the features."""
    from mvdet_b200 import integration
    pm, src_hw, grid_hw = small_rig(num_cam=3, seed=5)
    model = integration.patch_detector(_TinyDetector(pm, src_hw, grid_hw).to(DEV))
    imgs = torch.randn(2, 3, 3, src_hw[0] // 3, src_hw[1] // 3, generator=torch.Generator().manual_seed(2))
    real_stack = torch.stack
    sizes = []

    def spy(tensors, dim=0, **kw):
        sizes.append(max(t.numel() for t in tensors))
        return real_stack(tensors, dim=dim, **kw)

    monkeypatch.setattr(torch, "stack", spy)
    out, _ = model(imgs)
    out.sum().backward()
    assert max(sizes) <= 9, sizes                                    # only the 3x3 matrices get stacked


# ---- host-resident features: only the touched texels cross the bus -----------------------------
@pytest.mark.parametrize("kind", ["plain", "upsampled"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_copy_touched_texels_round_trip(kind, dtype):
    """Pinned host features -> device through the gather kernel: the warp of the sparsely staged tensor equals the warp of the dense
    one bit for bit (every texel the warp reads was sent); device gradient -> zeroed pinned host tensor through the scatter equals
    the dense gradient bit for bit (every non-zero row came back, the rest is zero)."""
    pm, src_hw, grid_hw = small_rig(num_cam=3, seed=71)
    low = (src_hw[0] // 3, src_hw[1] // 3)
    B, V, Cc = 2, 3, 24
    if kind == "plain":
        plan, feat_hw = P.get_plan(pm, src_hw, grid_hw, DEV), src_hw
    else:
        plan, feat_hw = P.get_plan_upsampled(pm, low, src_hw, grid_hw, DEV), low
    x = _feats(B, V, Cc, feat_hw).to(dtype)
    h_x = torch.empty((B, V, *feat_hw, Cc), dtype=dtype).pin_memory().permute(0, 1, 4, 2, 3)
    h_x.copy_(x)
    d_sparse = torch.full((B, V, *feat_hw, Cc), float("nan"), dtype=dtype, device=DEV).permute(0, 1, 4, 2, 3)
    ops.copy_touched_texels(plan, h_x, d_sparse)
    base, ids = ops.touched_texels(plan)
    assert int(base[-1]) == plan.stats["unique_texels"] == ids.numel()
    mask = torch.zeros(V, feat_hw[0] * feat_hw[1], dtype=torch.bool)
    for v in range(V):
        seg = ids[int(base[v]):int(base[v + 1])].long()
        assert bool((seg[1:] > seg[:-1]).all())
        mask[v, seg] = True
    mask = mask.view(1, V, 1, *feat_hw)
    got = d_sparse.cpu()
    assert bool(got.isnan().logical_not().eq(mask.expand_as(got)).all())          # exactly the touched texels were written
    d_sparse = torch.nan_to_num(d_sparse, nan=0.0)                                # untouched texels are never read; any finite value does
    d_dense = to_channels_last_5d(x).to(DEV).requires_grad_(True)
    a = ops.warp_views_with_plan(d_dense, plan)
    b = ops.warp_views_with_plan(d_sparse, plan)
    assert torch.equal(a, b)
    go = torch.randn(a.shape, generator=torch.Generator().manual_seed(4)).to(dtype).to(DEV)
    (g,) = torch.autograd.grad(a, d_dense, go)
    h_g = torch.zeros((B, V, *feat_hw, Cc), dtype=dtype).pin_memory().permute(0, 1, 4, 2, 3)
    ops.copy_touched_texels(plan, g, h_g)
    torch.cuda.synchronize()
    assert torch.equal(h_g, g.cpu())
    # view sub-range + argument checks
    d2 = torch.zeros_like(d_sparse[:, 1:])
    ops.copy_touched_texels(plan, h_x[:, 1:], d2, view0=1)
    assert torch.equal(torch.nan_to_num(d2.cpu()), torch.where(mask[:, 1:].expand_as(got[:, 1:]), h_x[:, 1:], torch.zeros_like(h_x[:, 1:])))
    with pytest.raises(mv.MvwarpError):
        ops.copy_touched_texels(plan, x.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3), d_sparse)   # pageable host memory
    with pytest.raises(ValueError):
        ops.copy_touched_texels(plan, torch.empty((B, V, Cc, *feat_hw), dtype=dtype).pin_memory(), d_sparse)   # not channel-contiguous


# ---- NCHW trunks through fused up-sample plans: one transpose of the low-res map, then the channel-contiguous kernels --------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_nchw_lowres_route_equals_channels_last(dtype):
    """warp_views_upsampled / warp_views_list on NCHW low-res maps (what an unmodified reference trunk emits) give the same bits as
    on channel-contiguous ones, forward and backward, and the gradient comes back in NCHW; odd sizes exercise the tile edges."""
    pm, src_hw, grid_hw = small_rig(num_cam=3, seed=81)
    low = (src_hw[0] // 3, src_hw[1] // 3)                               # 18 x 32
    B, V, Cc = 2, 3, 40
    x = _feats(B, V, Cc, low).to(dtype)
    x_nchw = x.contiguous().to(DEV).requires_grad_(True)
    x_cl = to_channels_last_5d(x).to(DEV).requires_grad_(True)
    a = ops.warp_views_upsampled(x_nchw, src_hw, pm, grid_hw)
    b = ops.warp_views_upsampled(x_cl, src_hw, pm, grid_hw)
    assert torch.equal(a, b) and a.stride(1) == 1 and b.stride(1) == 1   # channels-last memory format either way
    go = torch.randn(a.shape, generator=torch.Generator().manual_seed(9)).to(dtype).to(DEV)
    (ga,) = torch.autograd.grad(a, x_nchw, go)
    (gb,) = torch.autograd.grad(b, x_cl, go)
    assert ga.is_contiguous() and torch.equal(ga, gb)
    cams = [x[:, v].contiguous().to(DEV).requires_grad_(True) for v in range(V)]
    c = ops.warp_views_list(cams, pm, grid_hw, upsample_shape=src_hw, channels_last_out=True)
    assert torch.equal(c, b) and c.stride(1) == 1
    gl = torch.autograd.grad(c, cams, go)
    assert all(g.is_contiguous() and torch.equal(g, gb[:, v]) for v, g in enumerate(gl))
    # non-contiguous batch / view strides take per-image transposes
    big = torch.zeros(B, V + 1, Cc, *low, dtype=dtype, device=DEV)
    big[:, :V] = x_nchw.detach()
    assert torch.equal(ops.warp_views_upsampled(big[:, :V], src_hw, pm, grid_hw), a)


def test_layout_transpose_round_trip():
    from mvdet_b200 import _lib as L
    import ctypes as C
    for dtype, code in ((torch.float32, L.F32), (torch.bfloat16, L.BF16)):
        x = torch.randn(3, 37, 19 * 23, generator=torch.Generator().manual_seed(1)).to(dtype).to(DEV)      # ragged tiles
        y = torch.empty(3, 19 * 23, 37, dtype=dtype, device=DEV)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(L.load().mvwarp_layout_transpose(C.c_void_p(x.data_ptr()), 0, C.c_void_p(y.data_ptr()), 0, 3, 37, 19 * 23, code, 1, 0, stream), "t")
        assert torch.equal(y, x.transpose(1, 2).contiguous())
        z = torch.empty_like(x)
        L.check(L.load().mvwarp_layout_transpose(C.c_void_p(y.data_ptr()), 0, C.c_void_p(z.data_ptr()), 0, 3, 37, 19 * 23, code, 0, 0, stream), "t")
        assert torch.equal(z, x)


@pytest.mark.parametrize("stage_kb", [2, 8, 96])
def test_staged_upsample_forward_slabs_and_fallback(stage_kb):
    """fwd_table_stage_kernel (development build, MVWARP_TBL_VARIANT=3: TMA-staged texel rows; measured slower than the default tile
    kernel and therefore not shipped as the default) with staging buffers of MVWARP_STAGE_KB: 2 KB forces the gather fallback for most
    blocks, 8 KB forces several channel slabs per block, 96 KB stages whole rows -- every path must give the bits of the default
    kernel.  Knobs are read once per process, so each setting runs in a child process."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from mvdet_b200.build import build_library
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from mvdet_b200 import ops
from tests.util import small_rig, to_channels_last_5d
for dtype in (torch.float32, torch.bfloat16):
    for V, Cc, seed in ((3, 512, 91), (2, 384, 92)):
        pm, src_hw, grid_hw = small_rig(num_cam=V, seed=seed)
        low = (src_hw[0] // 3, src_hw[1] // 3)
        x = to_channels_last_5d(torch.randn(2, V, Cc, *low, generator=torch.Generator().manual_seed(seed)).to(dtype)).cuda()
        out = ops.warp_views_upsampled(x, src_hw, pm, grid_hw)
        torch.save(out.cpu(), sys.argv[1] + f"_{dtype}_{V}.pt")
print("ok")
""" % root
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        outs = {}
        # "default" = the CSR tile kernel (MVWARP_TBL_VARIANT=2), whose entry order the staged kernel follows; the shipped default for
        # these shapes is fwd_up_quad_kernel, which composes its weights in-kernel and agrees to rounding, not bit for bit
        dev_lib = build_library(dev=True)
        for tag, env_extra in (("default", {"MVWARP_LIB": dev_lib, "MVWARP_TBL_VARIANT": "2"}),
                               ("knob", {"MVWARP_LIB": dev_lib, "MVWARP_STAGE_KB": str(stage_kb), "MVWARP_TBL_VARIANT": "3"})):
            env = dict(os.environ, PYTHONPATH=root, **env_extra)
            res = subprocess.run([sys.executable, "-c", code, os.path.join(td, tag)], capture_output=True, text=True, env=env, cwd=root, timeout=600)
            assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
            outs[tag] = sorted(f for f in os.listdir(td) if f.startswith(tag))
        assert len(outs["default"]) == len(outs["knob"]) == 4
        for a, b in zip(outs["default"], outs["knob"]):
            ta, tb = torch.load(os.path.join(td, a)), torch.load(os.path.join(td, b))
            assert torch.equal(ta, tb), (stage_kb, a)


def test_quad_upsample_forward_agrees_with_the_shipped_kernel():
    """fwd_up_quad_kernel (development build, MVWARP_TBL_VARIANT=8: 2 x 2 quads of cells, 3 x 3 low-res windows composed in-kernel by
    csrc/mvwarp_uptaps.h; measured slower than the CSR tile kernel and therefore not shipped) computes the same operator as the shipped
    forward with weights composed in fp32 instead of double: agreement to rounding, fp32 and bf16, partial tiles, odd concat pitch."""
    import os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from mvdet_b200.build import build_library
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from mvdet_b200 import ops
from tests.util import small_rig, to_channels_last_5d
for dtype in (torch.float32, torch.bfloat16):
    for V, Cc, seed in ((3, 512, 91), (2, 96, 92)):
        pm, src_hw, grid_hw = small_rig(num_cam=V, seed=seed)
        low = (src_hw[0] // 3, src_hw[1] // 3)
        x = to_channels_last_5d(torch.randn(2, V, Cc, *low, generator=torch.Generator().manual_seed(seed)).to(dtype)).cuda()
        out = ops.warp_views_upsampled(x, src_hw, pm, grid_hw)
        torch.save(out.float().cpu(), sys.argv[1] + f"_{dtype}_{V}.pt")
print("ok")
""" % root
    dev_lib = build_library(dev=True)
    with tempfile.TemporaryDirectory() as td:
        outs = {}
        for tag, env_extra in (("shipped", {}), ("quad", {"MVWARP_LIB": dev_lib, "MVWARP_TBL_VARIANT": "8"})):
            env = dict(os.environ, PYTHONPATH=root, **env_extra)
            res = subprocess.run([sys.executable, "-c", code, os.path.join(td, tag)], capture_output=True, text=True, env=env, cwd=root, timeout=600)
            assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
            outs[tag] = sorted(f for f in os.listdir(td) if f.startswith(tag))
        assert len(outs["shipped"]) == len(outs["quad"]) == 4
        for a, b in zip(outs["shipped"], outs["quad"]):
            ta, tb = torch.load(os.path.join(td, a)), torch.load(os.path.join(td, b))
            tol = 1e-5 if "float32" in a else 2.0 ** -7          # bf16: the two fp32 results may round to neighbouring bf16 values
            assert float((ta - tb).abs().max()) <= tol * float(ta.abs().max()), (a, float((ta - tb).abs().max()))
            assert not torch.equal(ta, torch.zeros_like(ta))


# ---- memory safety without compute-sanitizer (closed on this pool): canary bands around every written tensor --------------
@pytest.mark.parametrize("layout", ["cl", "nchw"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kernels_do_not_write_outside_their_tensors(layout, dtype):
    """Forward and backward (plain and fused up-sample plans, list variant, touched-texel copy) into tensors carved out of the middle
    of a larger canary-filled allocation, on ragged shapes (C = 20: vector tails; 54x96 -> 24x72: partial tiles): every byte of the
    canary bands before and after the tensor must survive, and every element of the tensor itself must have been written."""
    pm, src_hw, grid_hw = small_rig(num_cam=3, seed=97)
    low = (src_hw[0] // 3, src_hw[1] // 3)
    B, V, Cc = 2, 3, 20
    GUARD = 4096

    def carve(shape5_or_4, cl):
        """tensor of the logical shape (channels-last physical if cl) in the middle of a canary-filled buffer"""
        n = 1
        for s in shape5_or_4:
            n *= s
        buf = torch.full((n + 2 * GUARD,), 12345.0, dtype=dtype, device=DEV)
        core = buf[GUARD:GUARD + n]
        core.fill_(float("nan"))
        if len(shape5_or_4) == 5:
            b, v, c, h, w = shape5_or_4
            t = core.view(b, v, h, w, c).permute(0, 1, 4, 2, 3) if cl else core.view(b, v, c, h, w)
        else:
            b, c, h, w = shape5_or_4
            t = core.view(b, h, w, c).permute(0, 3, 1, 2) if cl else core.view(b, c, h, w)
        return buf, t

    def check(buf, t, what):
        assert bool((buf[:GUARD] == 12345.0).all()) and bool((buf[-GUARD:] == 12345.0).all()), f"{what}: canary band overwritten"
        assert not bool(t.isnan().any()), f"{what}: part of the tensor was never written"

    cl = layout == "cl"
    for kind in ("plain", "upsampled"):
        if kind == "plain":
            plan, fhw = P.get_plan(pm, src_hw, grid_hw, DEV), src_hw
        else:
            plan, fhw = P.get_plan_upsampled(pm, low, src_hw, grid_hw, DEV), low
        x = _layout(_feats(B, V, Cc, fhw).to(dtype), layout).to(DEV)
        obuf, out = carve((B, V * Cc + 2, *grid_hw), cl)
        ops.warp_views_fwd_into(plan, x, out, 0, 0, V * Cc)
        torch.cuda.synchronize()
        check(obuf, out, f"{kind} forward")
        go = torch.randn_like(out)
        gbuf, gx = carve((B, V, Cc, *fhw), cl)
        ops.warp_views_bwd_into(plan, go, gx, 0, 0)
        torch.cuda.synchronize()
        check(gbuf, gx, f"{kind} backward")
        # list variant: per-camera gradients, each in its own canary buffer
        cams = [x[:, v].contiguous(memory_format=torch.channels_last) if cl else x[:, v].contiguous() for v in range(V)]
        obuf2, out2 = carve((B, V * Cc + 2, *grid_hw), cl)
        ops._launch_fwd_list(plan, cams, out2, V * Cc)
        torch.cuda.synchronize()
        check(obuf2, out2, f"{kind} forward (list)")
        assert torch.equal(out2, out)
        pairs = [carve((B, Cc, *fhw), cl) for _ in range(V)]
        ops._launch_bwd_list(plan, go, [t for _, t in pairs])
        torch.cuda.synchronize()
        for v, (bf, t) in enumerate(pairs):
            check(bf, t, f"{kind} backward (list, view {v})")
            assert torch.equal(t, gx[:, v])
        if cl:   # touched-texel copy into a canary-filled destination: only touched rows change, bands survive
            dbuf, dst = carve((B, V, Cc, *fhw), True)
            dst.fill_(0)
            ops.copy_touched_texels(plan, x, dst)
            torch.cuda.synchronize()
            check(dbuf, dst, f"{kind} touched copy")


def test_staged_nchw_kernels_give_the_default_bits():
    """fwd_nchw_stage_kernel / gather_nchw_stage_kernel (development build, MVWARP_NCHW_VARIANT=2: source footprint staged in shared
    memory; measured 2-3x SLOWER than the shipped planar kernels and therefore not the default) must still be correct: same bits as
    the default kernels on NCHW tensors, forward and backward, plain and fused up-sample plans."""
    import os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from mvdet_b200.build import build_library
    code = r"""
import sys, torch
sys.path.insert(0, %r)
from mvdet_b200 import ops, plan as P
from tests.util import small_rig
res = {}
for dtype in (torch.float32, torch.bfloat16):
    for V, Cc, seed in ((3, 70, 93), (2, 16, 94)):
        pm, src_hw, grid_hw = small_rig(num_cam=V, seed=seed)
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(2, V, Cc, *src_hw, generator=g).to(dtype).cuda().requires_grad_(True)
        out = ops.warp_views(x, pm, grid_hw)
        go = torch.randn(out.shape, generator=g).to(dtype).cuda()
        (gx,) = torch.autograd.grad(out, x, go)
        plan = P.get_plan_upsampled(pm, (src_hw[0] // 3, src_hw[1] // 3), src_hw, grid_hw, "cuda:0")
        xl = torch.randn(2, V, Cc, src_hw[0] // 3, src_hw[1] // 3, generator=g).to(dtype).cuda()
        o2 = torch.empty(2, V * Cc + 2, *grid_hw, dtype=dtype, device="cuda")
        ops.warp_views_fwd_into(plan, xl, o2, 0, 0, V * Cc)           # raw planar launch (no transposing route)
        g2 = torch.empty_like(xl)
        ops.warp_views_bwd_into(plan, torch.randn(o2.shape, generator=g).to(dtype).cuda(), g2, 0, 0)
        res[f"{dtype}_{V}"] = [t.cpu() for t in (out.detach(), gx, o2, g2)]
torch.save(res, sys.argv[1])
print("ok")
""" % root
    with tempfile.TemporaryDirectory() as td:
        files = {}
        for tag, extra in (("default", {}), ("staged", {"MVWARP_LIB": build_library(dev=True), "MVWARP_NCHW_VARIANT": "2"})):
            files[tag] = os.path.join(td, tag + ".pt")
            res = subprocess.run([sys.executable, "-c", code, files[tag]], capture_output=True, text=True,
                                 env=dict(os.environ, PYTHONPATH=root, **extra), cwd=root, timeout=600)
            assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
        a, b = torch.load(files["default"]), torch.load(files["staged"])
        assert a.keys() == b.keys() and len(a) == 4
        for k in a:
            for i, (ta, tb) in enumerate(zip(a[k], b[k])):
                assert torch.equal(ta, tb), (k, i)
