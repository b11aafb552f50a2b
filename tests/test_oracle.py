"""Oracle self-checks (CPU).  The reference pins nothing for this path, so the
restated kornia glue is validated against independent definitions of the same warp:
cv2.warpPerspective, the float64 pixel-space formula, and algebraic properties."""
import numpy as np
import pytest
import torch

from mvdet_b200.geometry import MULTIVIEWX, WILDTRACK, create_coord_map, synthetic_rig
from oracle import kornia_warp as K
from tests.util import small_rig


def _smooth_image(h, w):
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    return (np.sin(xs / 17.0) * np.cos(ys / 11.0) + 0.002 * xs + 0.001 * ys).astype(np.float32)


@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX], ids=lambda s: s.name)
def test_oracle_matches_cv2(spec):
    cv2 = pytest.importorskip("cv2")
    pm = synthetic_rig(spec)
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    img = _smooth_image(Hs, Ws)
    src = torch.from_numpy(img)[None, None]
    for v in range(spec.num_cam):
        ours = K.warp_perspective(src, pm[v][None].float(), [Hg, Wg])[0, 0].numpy()
        ref = cv2.warpPerspective(img, pm[v].numpy(), (Wg, Hg), flags=cv2.INTER_LINEAR,
                                  borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        # interior pixels only: away from the zero-padding fade, where cv2's border handling differs
        grid = K.sampling_grid(pm[v][None].float(), (Hs, Ws), (Hg, Wg))[0]
        ix = (grid[..., 0] + 1) / 2 * (Ws - 1)
        iy = (grid[..., 1] + 1) / 2 * (Hs - 1)
        interior = ((ix > 1) & (ix < Ws - 2) & (iy > 1) & (iy < Hs - 2)).numpy()
        assert interior.sum() > 1000
        # cv2 quantises bilinear weights to 1/32 -> 3e-3 on this image
        assert np.abs(ours - ref)[interior].max() < 3e-3


@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX], ids=lambda s: s.name)
def test_oracle_matches_float64_bilinear_of_scipy(spec):
    """A second independent implementation with EXACT (float64) bilinear weights: scipy.ndimage.map_coordinates(order=1) sampled
    at the float64 pixel-space coordinates M^-1 (xd, yd, 1).  cv2 pins the definition to its 1/32 weight quantisation (3e-3);
    this pins it to the fp32 coordinate noise of the kornia chain (<= 4e-4 px, times the image gradient <= 0.1 per px)."""
    ndi = pytest.importorskip("scipy.ndimage")
    pm = synthetic_rig(spec)
    Hs, Ws = spec.feat_shape
    Hg, Wg = spec.grid_shape
    img = _smooth_image(Hs, Ws)
    src = torch.from_numpy(img)[None, None]
    yd, xd = np.mgrid[0:Hg, 0:Wg].astype(np.float64)
    worst = 0.0
    for v in range(spec.num_cam):
        ours = K.warp_perspective(src, pm[v][None].float(), [Hg, Wg])[0, 0].numpy()
        Minv = np.linalg.inv(pm[v].numpy())
        q = np.stack([xd, yd, np.ones_like(xd)], -1) @ Minv.T
        px, py = q[..., 0] / q[..., 2], q[..., 1] / q[..., 2]
        ref = ndi.map_coordinates(img.astype(np.float64), [py, px], order=1, mode="constant", cval=0.0)
        interior = (px > 1) & (px < Ws - 2) & (py > 1) & (py < Hs - 2)
        assert interior.sum() > 1000
        worst = max(worst, float(np.abs(ours - ref)[interior].max()))
    assert worst < 5e-5, worst      # measured: 3e-6 .. 2.2e-5 over the 13 cameras


def test_oracle_grid_matches_float64_pixel_space():
    """normalise -> invert -> transform == direct M^-1 (xd, yd, 1) in float64."""
    pm = synthetic_rig(WILDTRACK)
    Hs, Ws = WILDTRACK.feat_shape
    Hg, Wg = WILDTRACK.grid_shape
    for v in range(WILDTRACK.num_cam):
        grid = K.sampling_grid(pm[v][None].double(), (Hs, Ws), (Hg, Wg))[0]
        ix = (grid[..., 0] + 1) / 2 * (Ws - 1)
        iy = (grid[..., 1] + 1) / 2 * (Hs - 1)
        Minv = torch.linalg.inv(pm[v])
        yd, xd = torch.meshgrid(torch.arange(Hg, dtype=torch.float64), torch.arange(Wg, dtype=torch.float64), indexing="ij")
        q = torch.stack([xd, yd, torch.ones_like(xd)], -1) @ Minv.T
        px, py = q[..., 0] / q[..., 2], q[..., 1] / q[..., 2]
        # Inside the image.  kornia builds the mesh grid in fp32 even for a float64 warp, so the float64
        # route still carries ~1e-4 px of mesh rounding (measured 3e-5..1.2e-4); direction, (x, y) order
        # and the align_corners convention are what this pins.
        near = (px > -1) & (px < Ws) & (py > -1) & (py < Hs)
        assert int(near.sum()) > 10000
        assert float((ix - px)[near].abs().max()) < 1e-3
        assert float((iy - py)[near].abs().max()) < 1e-3


def test_identity_homography_is_identity():
    torch.manual_seed(0)
    src = torch.randn(2, 3, 17, 23)
    M = torch.eye(3)[None].repeat(2, 1, 1)
    out = K.warp_perspective(src, M, [17, 23])
    assert float((out - src).abs().max()) < 2e-5


def test_integer_translation_shifts_with_zero_fill():
    torch.manual_seed(0)
    src = torch.randn(1, 2, 16, 20)
    M = torch.tensor([[[1.0, 0, 3], [0, 1, 2], [0, 0, 1]]])   # dst = src + (3, 2)
    out = K.warp_perspective(src, M, [16, 20])
    expect = torch.zeros_like(src)
    expect[..., 2:, 3:] = src[..., :-2, :-3]
    assert float((out - expect).abs().max()) < 5e-5


def test_linearity_in_src():
    pm, src_hw, grid_hw = small_rig()
    torch.manual_seed(0)
    a = torch.randn(1, 3, 4, *src_hw)
    b = torch.randn(1, 3, 4, *src_hw)
    fa, fb, fab = (K.warp_views_oracle(t, pm, grid_hw) for t in (a, b, 2 * a - 3 * b))
    lin = 2 * fa - 3 * fb
    lin[:, -2:] = fab[:, -2:]   # coord channels are not linear in src
    assert float((fab - lin).abs().max()) < 1e-4


def test_adjoint_identity_of_oracle_gradient():
    """<warp(x), g> == <x, warp^T(g)>: the autograd gradient used as grad oracle is the true adjoint."""
    pm, src_hw, grid_hw = small_rig()
    torch.manual_seed(0)
    x = torch.randn(2, 3, 4, *src_hw, dtype=torch.float64, requires_grad=True)
    out = K.warp_views_oracle(x, pm.double(), grid_hw)
    g = torch.randn_like(out)
    (gx,) = torch.autograd.grad(out, x, g)
    lhs = float((out[:, :-2] * g[:, :-2]).sum().detach())
    rhs = float((x * gx).sum().detach())
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))


def test_coord_map_and_channel_order():
    pm, src_hw, grid_hw = small_rig()
    B, V, C = 2, 3, 4
    torch.manual_seed(0)
    x = torch.randn(B, V, C, *src_hw)
    out = K.warp_views_oracle(x, pm, grid_hw)
    assert out.shape == (B, V * C + 2, *grid_hw)
    cm = out[:, -2:]
    assert float(cm[0, 0, 0, 0]) == -1.0 and float(cm[0, 0, 0, -1]) == 1.0      # x channel
    assert float(cm[0, 1, 0, 0]) == -1.0 and float(cm[0, 1, -1, 0]) == 1.0      # y channel
    assert torch.equal(cm[0], cm[1])
    assert torch.equal(cm[:1], create_coord_map(grid_hw))
    for v in range(V):   # view-major channel order
        single = K.warp_perspective(x[:, v], pm[v].repeat(B, 1, 1).float(), list(grid_hw))
        assert torch.equal(out[:, v * C:(v + 1) * C], single)


def test_kornia_error_conventions():
    with pytest.raises(TypeError):
        K.warp_perspective(np.zeros((1, 1, 4, 4)), torch.eye(3)[None], [4, 4])
    with pytest.raises(TypeError):
        K.warp_perspective(torch.zeros(1, 1, 4, 4), np.eye(3), [4, 4])
    with pytest.raises(ValueError):
        K.warp_perspective(torch.zeros(1, 4, 4), torch.eye(3)[None], [4, 4])
    with pytest.raises(ValueError):
        K.warp_perspective(torch.zeros(1, 1, 4, 4), torch.eye(3), [4, 4])


def test_create_meshgrid_arithmetic_matches_a_vendored_copy_of_kornia():
    """The only piece of kornia source that exists in this image: transformers ships a copy of ``kornia.utils.grid.create_meshgrid``
    ("Copied from kornia library : kornia/kornia/utils/grid.py:26", transformers/models/efficientloftr/modeling_efficientloftr.py).
    Its 1-D coordinate arithmetic -- ``linspace(0, n-1, n)``, then ``(xs / (n - 1) - 0.5) * 2`` -- is kornia's; the copy builds the 2-D
    grid with the meshgrid arguments swapped (it returns [1, W, H, (y, x)]), so the comparison is made on the per-axis values, which
    is where the fp32 rounding the kernels must reproduce (mvwarp_coords.h mesh_coord) lives.  This pins step 4 of SURVEY.md 8c
    against third-party-held kornia code; the rest of the chain stays pinned by cv2 / scipy / the plain-C restatement only."""
    mod = pytest.importorskip("transformers.models.efficientloftr.modeling_efficientloftr")
    for h, w in [(120, 360), (160, 250), (480, 1440), (24, 72), (5, 7), (2, 2)]:
        theirs = mod.create_meshgrid(h, w, normalized_coordinates=True, dtype=torch.float32)     # [1, W, H, 2], last dim (y, x)
        ours = K.create_meshgrid(h, w, normalized_coordinates=True)                              # [1, H, W, 2], last dim (x, y)
        assert theirs.shape == (1, w, h, 2) and ours.shape == (1, h, w, 2)
        xs_theirs, ys_theirs = theirs[0, :, 0, 1], theirs[0, 0, :, 0]
        assert torch.equal(ours[0, 0, :, 0], xs_theirs) and torch.equal(ours[0, :, 0, 1], ys_theirs)
        assert torch.equal(ours[0, :, :, 0], xs_theirs[None, :].expand(h, w))
        assert torch.equal(ours[0, :, :, 1], ys_theirs[:, None].expand(h, w))
