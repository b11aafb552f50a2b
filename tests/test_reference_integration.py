"""Wiring test against the UNMODIFIED reference model (CPU, this container only: /root/reference is not on
the GPU box).  kornia and matplotlib are not installed, so they are stubbed in sys.modules; the kornia stub's
warp_perspective is the oracle.  What is pinned: the arguments the reference passes at
persp_trans_detector.py:68-69, that replacing :68-69,75,77 by one warp_views-shaped call leaves map_result
unchanged, that mvdet_b200.geometry reproduces the reference's proj_mats / coord_map bit for bit, and that the
kornia-signature seam (mvdet_b200.ops.warp_perspective) accepts exactly that call."""
import inspect
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "multiview_detector")),
                                reason="reference tree not present (GPU box)")


class _Base:
    pass


class _Dataset:
    pass


def _fake_dataset(num_cam=3):
    from mvdet_b200.geometry import synthetic_cameras
    Ks, Es = synthetic_cameras(num_cam, seed=3)
    scale = np.diag([0.2, 0.2, 1.0])                       # 1080x1920 calibration -> 216x384 images
    base = _Base()
    base.intrinsic_matrices = [scale @ K for K in Ks]
    base.extrinsic_matrices = Es
    base.worldgrid2worldcoord_mat = np.array([[2.5, 0, -300], [0, 2.5, -900], [0, 0, 1.0]])
    ds = _Dataset()
    ds.base, ds.num_cam = base, num_cam
    ds.img_shape, ds.reducedgrid_shape = [216, 384], [24, 72]
    ds.img_reduce, ds.grid_reduce = 4, 20
    return ds


@pytest.fixture()
def reference_model(monkeypatch):
    from oracle import kornia_warp as K
    calls = []

    def recording_warp(src, M, dsize, *args, **kwargs):
        calls.append((tuple(src.shape), src.dtype, src.is_contiguous(), tuple(M.shape), M.dtype, dsize, args, kwargs))
        return K.warp_perspective(src, M, dsize, *args, **kwargs)

    kornia = types.ModuleType("kornia")
    kornia.geometry = types.ModuleType("kornia.geometry")
    kornia.geometry.transform = types.ModuleType("kornia.geometry.transform")
    kornia.geometry.transform.warp_perspective = recording_warp
    kornia.warp_perspective = recording_warp
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    for name, mod in {"kornia": kornia, "kornia.geometry": kornia.geometry,
                      "kornia.geometry.transform": kornia.geometry.transform,
                      "matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot}.items():
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.syspath_prepend(REF)
    # the reference pins everything to 'cuda:0' (persp_trans_detector.py:37-54,63-81): redirect to CPU
    t_to, m_to = torch.Tensor.to, torch.nn.Module.to
    fix = lambda a: tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: t_to(self, *fix(a), **k))
    monkeypatch.setattr(torch.nn.Module, "to", lambda self, *a, **k: m_to(self, *fix(a), **k))
    for m in [k for k in sys.modules if k.startswith("multiview_detector")]:
        monkeypatch.delitem(sys.modules, m)
    from multiview_detector.models.persp_trans_detector import PerspTransDetector
    torch.manual_seed(0)
    model = PerspTransDetector(_fake_dataset(), arch="resnet18").eval()
    model._calls = calls
    return model, calls


def test_reference_forward_through_the_seams(reference_model):
    from mvdet_b200 import geometry, ops
    from oracle import kornia_warp as K
    model, calls = reference_model
    ds = _fake_dataset()
    B = 2
    imgs = torch.randn(B, ds.num_cam, 3, 144, 256, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        map_ref, img_res = model(imgs)
    assert map_ref.shape == (B, 1, 24, 72) and len(img_res) == ds.num_cam

    # (1) the call the seam must accept: positional (src, M, dsize), fp32 NCHW-contiguous, M = [B,3,3] fp32, dsize a list
    assert len(calls) == ds.num_cam
    for shape, dtype, contig, mshape, mdtype, dsize, args, kwargs in calls:
        assert shape == (B, 512, 54, 96) and dtype == torch.float32 and contig
        assert mshape == (B, 3, 3) and mdtype == torch.float32 and dsize == [24, 72] and not args and not kwargs
    seam = inspect.signature(ops.warp_perspective)
    ref_sig = inspect.signature(K.warp_perspective)
    assert list(seam.parameters) == list(ref_sig.parameters)
    assert [p.default for p in seam.parameters.values()] == [p.default for p in ref_sig.parameters.values()]

    # (2) proj_mats and coord_map rebuilt by mvdet_b200.geometry == the reference's, bit for bit
    pm = geometry.projection_matrices(ds.base.intrinsic_matrices, ds.base.extrinsic_matrices,
                                      ds.base.worldgrid2worldcoord_mat, ds.img_shape, ds.img_reduce, ds.grid_reduce)
    assert torch.equal(pm, torch.stack(model.proj_mats))
    assert torch.equal(geometry.create_coord_map(ds.reducedgrid_shape), model.coord_map)

    # (3) seam 2: lines 68-69, 75, 77 replaced by one warp_views-shaped call (oracle standing in for the CUDA op)
    with torch.no_grad():
        feats = []
        for cam in range(ds.num_cam):
            f = model.base_pt2(model.base_pt1(imgs[:, cam]))
            feats.append(F.interpolate(f, model.upsample_shape, mode="bilinear"))
        world = K.warp_views_oracle(torch.stack(feats, dim=1), torch.stack(model.proj_mats), model.reducedgrid_shape)
        assert world.shape == (B, 512 * ds.num_cam + 2, 24, 72)
        map_new = F.interpolate(model.map_classifier(world), model.reducedgrid_shape, mode="bilinear")
    assert torch.equal(map_new, map_ref)


@pytest.mark.parametrize("variant,cls,shape,contig", [
    ("image_proj_variant", "ImageProjVariant", (2, 3, 54, 96), True),       # warps the RGB image itself: C = 3
    ("res_proj_variant", "ResProjVariant", (2, 1, 54, 96), False),          # warps img_res[:, 1].unsqueeze(1): C = 1, a strided slice
    ("no_joint_conv_variant", "NoJointConvVariant", (2, 512, 54, 96), True),
])
def test_ablation_variants_call_the_same_seam(reference_model, variant, cls, shape, contig):
    """the three ablation models reach the warp through the alias kornia.warp_perspective
    (image_proj_variant.py:69, res_proj_variant.py:72, no_joint_conv_variant.py:68) with C = 3 / 1 / 512; pin what they
    pass, so the GPU tests of the seam (tests/test_gpu_parity.py::test_ablation_variant_shapes) use the same calls."""
    import importlib
    _, calls = reference_model                      # fixture installs the stubs and the cuda->cpu redirection
    mod = importlib.import_module("multiview_detector.models." + variant)
    ds = _fake_dataset()
    torch.manual_seed(0)
    model = getattr(mod, cls)(ds, arch="resnet18").eval()
    del calls[:]
    imgs = torch.randn(2, ds.num_cam, 3, 144, 256, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        map_res, _ = model(imgs)
    assert map_res.shape == (2, 1, 24, 72)
    assert len(calls) == ds.num_cam
    for got_shape, dtype, is_contig, mshape, mdtype, dsize, args, kwargs in calls:
        assert got_shape == shape and dtype == torch.float32 and is_contig == contig
        assert mshape == (2, 3, 3) and mdtype == torch.float32 and dsize == [24, 72] and not args and not kwargs


def test_patch_detector_reproduces_the_reference_forward(reference_model):
    """mvdet_b200.integration.patch_detector (seam 2 as shipped code): the patched forward, with the oracle standing in for the
    CUDA op, returns exactly what the unmodified reference forward returns -- map_result and every img_res, bit for bit."""
    from mvdet_b200 import integration
    from oracle import kornia_warp as K
    model, calls = reference_model
    ds = _fake_dataset()
    imgs = torch.randn(2, ds.num_cam, 3, 144, 256, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        map_ref, img_ref = model(imgs)
    n_ref_calls = len(calls)
    assert n_ref_calls == ds.num_cam
    seen = []

    def warp_fn(img_feats, proj_mats, grid_shape):
        seen.append((tuple(img_feats.shape), proj_mats.dtype, tuple(proj_mats.shape), list(grid_shape)))
        return K.warp_views_oracle(img_feats, proj_mats, grid_shape)

    patched = integration.patch_detector(model, channels_last=False, warp_fn=warp_fn)
    assert patched is model
    with torch.no_grad():
        map_new, img_new = model(imgs)
    assert len(calls) == n_ref_calls                       # the per-camera kornia call is gone ...
    assert seen == [((2, ds.num_cam, 512, 54, 96), torch.float64, (ds.num_cam, 3, 3), [24, 72])]   # ... one warp_views call instead
    assert torch.equal(map_new, map_ref)
    assert len(img_new) == len(img_ref) and all(torch.equal(a, b) for a, b in zip(img_new, img_ref))
    with torch.no_grad():
        map_again, _ = model.reference_forward(imgs)       # the original forward stays reachable
    assert torch.equal(map_again, map_ref) and len(calls) == 2 * n_ref_calls
    with pytest.raises(NotImplementedError):
        model(imgs, visualize=True)
    with pytest.raises(AssertionError):
        model(imgs[:, :2])
    with pytest.raises(TypeError):
        integration.patch_detector(torch.nn.Linear(2, 2))


def test_patch_detector_channels_last_keeps_the_result(reference_model):
    """channels_last=True only changes memory formats: same map_result to rounding (CPU convolutions take other kernels)."""
    from mvdet_b200 import integration
    from oracle import kornia_warp as K
    model, _ = reference_model
    ds = _fake_dataset()
    imgs = torch.randn(1, ds.num_cam, 3, 144, 256, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        map_ref, _ = model(imgs)
    layouts = []

    def warp_fn(img_feats, proj_mats, grid_shape):
        layouts.append(img_feats.stride(2) == 1)
        return K.warp_views_oracle(img_feats, proj_mats, grid_shape)

    integration.patch_detector(model, channels_last=True, warp_fn=warp_fn)
    with torch.no_grad():
        map_new, _ = model(imgs.contiguous())
    assert layouts == [True]                               # trunk + F.interpolate + torch.stack hand over channel-contiguous features
    assert float((map_new - map_ref).abs().max()) <= 1e-4 * float(map_ref.abs().max())


def test_patch_detector_fused_upsample_agrees_to_rounding(reference_model):
    """fuse_upsample=True: the warp reads the trunk's low-res output (two-step oracle standing in for warp_views_upsampled) and
    img_classifier up-samples after its first 1x1 convolution -- the same map_result / img_res as the reference order to rounding."""
    from mvdet_b200 import integration
    from oracle import kornia_warp as K
    model, calls = reference_model
    ds = _fake_dataset()
    imgs = torch.randn(2, ds.num_cam, 3, 144, 256, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        map_ref, img_ref = model(imgs)
    seen = []

    def warp_up(img_feats_low, upsample_shape, proj_mats, grid_shape):
        seen.append((tuple(img_feats_low.shape), list(upsample_shape)))
        V = img_feats_low.shape[1]
        up = torch.stack([F.interpolate(img_feats_low[:, v], upsample_shape, mode="bilinear") for v in range(V)], dim=1)  # :65
        return K.warp_views_oracle(up, proj_mats, grid_shape)

    integration.patch_detector(model, channels_last=False, warp_fn=warp_up, fuse_upsample=True)
    with torch.no_grad():
        map_new, img_new = model(imgs)
    assert seen == [((2, ds.num_cam, 512, 18, 32), [54, 96])]          # the 512-channel up-sample is gone from the model's forward
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(map_new, map_ref) <= 1e-5
    assert all(rel(a, b) <= 1e-5 for a, b in zip(img_new, img_ref))
    model.img_classifier = torch.nn.Sequential(torch.nn.Conv2d(512, 2, 3, padding=1))
    with pytest.raises(TypeError):
        integration.patch_detector(model, fuse_upsample=True)          # a 3x3 first layer does not commute with the up-sample


def test_fused_head_over_lowres_features_on_the_reference_model(reference_model):
    """fuse_upsample=True together with a fused head (integration.warped_forward, the wiring behind patch_detector(fuse_upsample=True,
    fused_head=True)): under no_grad the head receives the trunk's LOW-RES features in bf16 plus upsample_shape -- lines 65 and 68-81 of
    persp_trans_detector.py collapse into one call -- and with a stand-in head that does the reference's arithmetic (F.interpolate,
    oracle warp, the model's own map_classifier) the unmodified reference model returns the same map_result to bf16 rounding."""
    from mvdet_b200 import integration
    from oracle import kornia_warp as K
    model, calls = reference_model
    ds = _fake_dataset()
    imgs = torch.randn(1, ds.num_cam, 3, 144, 256, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        map_ref, img_ref = model(imgs)
    seen = []

    def head(img_feats_low, proj_mats, grid_shape, upsample_shape=None):
        seen.append((tuple(img_feats_low.shape), img_feats_low.dtype, tuple(upsample_shape)))
        V = img_feats_low.shape[1]
        up = torch.stack([F.interpolate(img_feats_low[:, v].float(), upsample_shape, mode="bilinear") for v in range(V)], dim=1)   # :65
        return model.map_classifier(K.warp_views_oracle(up, proj_mats, grid_shape))                                              # :68-81

    with torch.no_grad():
        map_new, img_new = integration.warped_forward(model, imgs, fuse_upsample=True, fused_head=head,
                                                      warp_fn=lambda *a, **k: pytest.fail("the un-fused warp must not run"))
    assert seen == [((1, ds.num_cam, 512, 18, 32), torch.bfloat16, (54, 96))]
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(map_new, map_ref) <= 3e-2                                  # bf16 features
    assert all(rel(a, b) <= 1e-5 for a, b in zip(img_new, img_ref))       # img branch: fp32, up-sample after the 1x1 convolution
    # with autograd on and no fused_conv1 the un-fused path runs (the fused head is forward-only)
    ran = []

    def warp_up(img_feats_low, upsample_shape, proj_mats, grid_shape):
        ran.append(tuple(img_feats_low.shape))
        V = img_feats_low.shape[1]
        up = torch.stack([F.interpolate(img_feats_low[:, v], upsample_shape, mode="bilinear") for v in range(V)], dim=1)
        return K.warp_views_oracle(up, proj_mats, grid_shape)

    map_g, _ = integration.warped_forward(model, imgs, fuse_upsample=True, fused_head=head, warp_fn=warp_up)
    assert ran == [(1, ds.num_cam, 512, 18, 32)] and len(seen) == 1 and map_g.requires_grad
