"""Randomised property tests (hypothesis) of the host side of the path: the coordinate chain must reproduce the oracle's
bits for ANY homography and size -- not only the three fixed rigs -- and the inverse CSR map must be the exact transpose of
the forward taps.  CPU only; sizes are small so the whole file runs in seconds."""
import numpy as np
import pytest
import torch

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st          # noqa: E402

from mvdet_b200 import plan as P                                   # noqa: E402
from mvdet_b200.geometry import RigSpec, synthetic_rig             # noqa: E402
from oracle import kornia_warp as K                                # noqa: E402

SETTINGS = dict(max_examples=25, deadline=None, derandomize=True)


@pytest.fixture(autouse=True)
def _single_threaded_oracle():
    """The oracle is the kornia chain on ONE CPU thread -- the reference forces OMP_NUM_THREADS=1 (main.py:3).  That matters in one
    corner these tests found: when MKL has >= ~3 threads per bmm batch item (B*Hg = 2 at 6-8 threads, <= 5 at 16) it splits each
    tiny GEMM across threads and the columns just past a multiple of 48 (Wg = 97..106, 148, 152, 196 ...) come out 1 ulp
    different; B*Hg = 1 and B*Hg >= 6 (swept to 480, 3-16 threads) are thread-count independent -- every real grid."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _bits(t):
    return t.contiguous().view(torch.int32)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), num_cam=st.integers(1, 4), hs=st.integers(2, 40), ws=st.integers(2, 60),
       hg=st.integers(1, 30), wg=st.sampled_from([1, 2, 7, 31, 44, 45, 64, 97]))
def test_host_grid_bits_equal_oracle_on_random_camera_rigs(seed, num_cam, hs, ws, hg, wg):
    """Synthetic pin-hole rigs pushed through the reference's matrix pipeline, arbitrary (also degenerate 1-wide) sizes;
    Wg straddles torch.bmm's small-matrix switch (9*Wg < 400 <=> Wg <= 44)."""
    spec = RigSpec("rnd", num_cam, img_shape=(1080, 1920), worldgrid_shape=(480, 1440), img_reduce=4, grid_reduce=4, seed=seed)
    pm = synthetic_rig(spec)
    T = P.src_norm_trans_dst_norm(pm, (hs, ws), (hg, wg))
    g = P.host_grid(T, (hg, wg))
    og = K.sampling_grids(pm, (hs, ws), (hg, wg))
    assert torch.equal(_bits(g), _bits(og))


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), scale=st.sampled_from([1e-3, 1.0, 1e3]), hg=st.integers(2, 12), wg=st.sampled_from([3, 44, 45, 50]))
def test_host_grid_bits_equal_oracle_on_arbitrary_matrices(seed, scale, hg, wg):
    """Arbitrary (not camera-like) normalised transforms, including ones whose homogeneous z crosses zero inside the grid:
    the |z| > 1e-8 special case, sign flips, infinities and NaNs must come out exactly as in the torch op sequence."""
    rng = np.random.default_rng(seed)
    T = torch.from_numpy((rng.standard_normal((2, 3, 3)) * scale).astype(np.float32))
    T[1, 2] = torch.tensor([0.5, -0.25, 0.0])                       # z = 0 along a line through the grid
    g = P.host_grid(T, (hg, wg))
    grid = K.create_meshgrid(hg, wg, normalized_coordinates=True).repeat(2, 1, 1, 1)
    og = K.transform_points(T[:, None, None], grid)
    same = _bits(g) == _bits(og)
    both_nan = torch.isnan(g) & torch.isnan(og)                     # NaN payload bits may differ; NaN-ness must not
    assert bool((same | both_nan).all())


def _dense_forward_matrix(grid_v, src_hw):
    """[cells, texels] float64 matrix of the oracle's forward for one view: autograd through F.grid_sample."""
    Hs, Ws = src_hw
    Hg, Wg = grid_v.shape[:2]
    eye = torch.eye(Hs * Ws, dtype=torch.float64).reshape(1, Hs * Ws, Hs, Ws)
    out = torch.nn.functional.grid_sample(eye, grid_v[None].double(), mode="bilinear", padding_mode="zeros", align_corners=True)
    return out[0].reshape(Hs * Ws, Hg * Wg).T                       # [cell, texel]


@settings(max_examples=12, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10_000), hs=st.integers(2, 9), ws=st.integers(2, 11), hg=st.integers(2, 8), wg=st.integers(2, 10))
def test_inverse_csr_is_the_transpose_of_the_forward_taps(seed, hs, ws, hg, wg):
    """CSR rows (texel -> cells, weights) against the dense forward matrix of grid_sample on the same grid: same sparsity,
    weights equal to fp32 rounding, every texel exactly one row, entries of a row in ascending cell order.  (1-wide grids are left
    out: kornia's mesh is 0/0 = NaN there, the forward is 0*NaN and the map is empty -- tests/test_host.py covers NaN grids.)"""
    spec = RigSpec("rnd", 2, img_shape=(1080, 1920), worldgrid_shape=(480, 1440), img_reduce=4, grid_reduce=4, seed=seed)
    pm = synthetic_rig(spec)
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, (hs, ws), (hg, wg)), (hg, wg))
    stt, tex_of, row_ptr, cells, weights = P.host_csr(g, (hs, ws), with_map=True)
    base = 0
    total = 0
    for v in range(2):
        A = _dense_forward_matrix(g[v], (hs, ws))                   # [cells, texels]
        mine = torch.zeros_like(A)
        rp = row_ptr[v].long()
        n = int(rp[-1])
        assert torch.equal(torch.sort(tex_of[v].long()).values, torch.arange(hs * ws))
        for r in range(hs * ws):
            seg = slice(base + int(rp[r]), base + int(rp[r + 1]))
            c = cells[seg].long()
            assert bool((c[1:] >= c[:-1]).all())
            mine[:, int(tex_of[v][r])].index_add_(0, c, weights[seg].double())
        base += n
        total += n
        assert float((mine - A).abs().max()) <= 2e-6
        # entries with zero weight are legal (a tap exactly on a texel), but no entry may exist where the forward has none
        assert bool(((mine != 0) <= (A.abs() > 0) | (mine.abs() <= 2e-6)).all())
    assert total == stt["nnz"]


@settings(max_examples=15, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10_000), hl=st.integers(1, 6), wl=st.integers(1, 7), fy=st.sampled_from([1, 2, 3, 2.5, 1.4]),
       fx=st.sampled_from([1, 2, 3, 1.75]), hg=st.integers(2, 7), wg=st.integers(2, 9))
def test_composed_upsample_map_equals_the_two_step_matrix(seed, hl, wl, fy, fx, hg, wg):
    """Fused up-sample plans for ANY low-res size and (also non-integer, also 1:1) up-sample ratio: the composed map on the
    low-res image against the dense matrix of  grid_sample(F.interpolate(., bilinear, align_corners=False))  in float64."""
    import torch.nn.functional as F
    hs, ws = max(2, int(round(hl * fy))), max(2, int(round(wl * fx)))
    spec = RigSpec("rnd", 1, img_shape=(1080, 1920), worldgrid_shape=(480, 1440), img_reduce=4, grid_reduce=4, seed=seed)
    pm = synthetic_rig(spec)
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, (hs, ws), (hg, wg)), (hg, wg))
    stt, tex_of, row_ptr, cells, weights = P.host_csr(g, (hs, ws), with_map=True, lowres_hw=(hl, wl))
    nlow = hl * wl
    eye = torch.eye(nlow, dtype=torch.float64).reshape(1, nlow, hl, wl)
    # ATen computes the up-sample's source index and lambdas in the tensor's dtype; the plan mirrors the fp32 formulas, so build the
    # reference up-sample in fp32 (exact on a one-hot basis: each value is a product of two fp32 lambdas) and warp it in float64
    up = F.interpolate(eye.float(), (hs, ws), mode="bilinear").double()
    A = F.grid_sample(up, g[0][None].double(), mode="bilinear", padding_mode="zeros", align_corners=True)[0].reshape(nlow, hg * wg).T
    mine = torch.zeros_like(A)
    rp = row_ptr[0].long()
    assert torch.equal(torch.sort(tex_of[0].long()).values, torch.arange(nlow))
    for r in range(nlow):
        seg = slice(int(rp[r]), int(rp[r + 1]))
        mine[:, int(tex_of[0][r])].index_add_(0, cells[seg].long(), weights[seg].double())
    assert int(rp[-1]) == stt["nnz"]
    assert float((mine - A).abs().max()) <= 2e-6, float((mine - A).abs().max())


@settings(max_examples=20, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10_000), hl=st.integers(1, 6), wl=st.integers(1, 7), fy=st.sampled_from([1, 2, 3, 2.5, 1.4]),
       fx=st.sampled_from([1, 2, 3, 1.75]), hg=st.integers(2, 7), wg=st.integers(2, 9))
def test_upsample_windows_equal_the_two_step_matrix(seed, hl, wl, fy, fx, hg, wg):
    """The 3 x 3 low-res windows of the fused warp -> conv1 kernel (csrc/mvwarp_uptaps.h, run on the host by mvwarp_host_upsample_windows)
    for ANY low-res size and up-sample ratio >= 1 (integer, non-integer, 1:1): as a dense matrix they equal
    grid_sample(F.interpolate(., bilinear, align_corners=False)) in float64, every live slot addresses a texel of the map, and
    they agree with the plan builder's composed CSR (weights composed in double) to fp32 rounding."""
    import torch.nn.functional as F
    hs, ws = max(hl, max(2, int(round(hl * fy)))), max(wl, max(2, int(round(wl * fx))))
    spec = RigSpec("rnd", 1, img_shape=(1080, 1920), worldgrid_shape=(480, 1440), img_reduce=4, grid_reduce=4, seed=seed)
    pm = synthetic_rig(spec)
    T = P.src_norm_trans_dst_norm(pm, (hs, ws), (hg, wg))
    g = P.host_grid(T, (hg, wg))
    origin, mask, wx, wy = P.host_upsample_windows(T, (hl, wl), (hs, ws), (hg, wg))
    nlow = hl * wl
    eye = torch.eye(nlow, dtype=torch.float64).reshape(1, nlow, hl, wl)
    up = F.interpolate(eye.float(), (hs, ws), mode="bilinear").double()
    A = F.grid_sample(up, g[0][None].double(), mode="bilinear", padding_mode="zeros", align_corners=True)[0].reshape(nlow, hg * wg).T
    mine = torch.zeros_like(A)
    for cell in range(hg * wg):
        y, x = divmod(cell, wg)
        for j in range(3):
            for i in range(3):
                if (int(mask[0, y, x]) >> (j * 3 + i)) & 1:
                    yy, xx = int(origin[0, y, x, 1]) + j, int(origin[0, y, x, 0]) + i
                    assert 0 <= yy < hl and 0 <= xx < wl
                    mine[cell, yy * wl + xx] += float(wy[0, y, x, j]) * float(wx[0, y, x, i])
    assert float((mine - A).abs().max()) <= 2e-6, float((mine - A).abs().max())
    # the composed CSR of the plan (the backward's weights) describes the same operator
    stt, tex_of, row_ptr, cells, weights = P.host_csr(g, (hs, ws), with_map=True, lowres_hw=(hl, wl))
    csr = torch.zeros_like(A)
    rp = row_ptr[0].long()
    for r in range(nlow):
        seg = slice(int(rp[r]), int(rp[r + 1]))
        csr[:, int(tex_of[0][r])].index_add_(0, cells[seg].long(), weights[seg].double())
    assert float((mine - csr).abs().max()) <= 2e-6
