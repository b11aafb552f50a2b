"""The plain-C restatement (oracle/warp_ref.c) against the torch oracle: two independent
implementations of the same published algorithm must agree bit for bit on the forward."""
import ctypes as C
import os
import subprocess

import pytest
import torch

from mvdet_b200 import plan as P
from oracle import kornia_warp as K
from tests.util import small_rig

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


@pytest.fixture(scope="module")
def cref():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return C.CDLL(os.path.join(ORACLE_DIR, "libwarp_ref.so"))


def _p(t):
    return C.cast(t.data_ptr(), C.POINTER(C.c_float))


def test_c_restatement_matches_torch_oracle(cref):
    pm, (Hs, Ws), (Hg, Wg) = small_rig()
    V, N, Cc = pm.shape[0], 2, 3
    torch.manual_seed(0)
    x = torch.randn(N, V, Cc, Hs, Ws)
    xo = x.clone().requires_grad_(True)
    ref = K.warp_views_oracle(xo, pm, (Hg, Wg))
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (gref,) = torch.autograd.grad(ref, xo, go)
    for v in range(V):
        # 3x3 part from torch (kornia's own ops); per-cell chain, sampling and scatter in C
        T = K._torch_inverse_cast(K.normalize_homography(pm[v][None].float(), (Hs, Ws), (Hg, Wg)))[0].contiguous()
        grid = torch.empty(Hg, Wg, 2)
        cref.warp_ref_grid(_p(T), Hg, Wg, _p(grid))
        og = K.sampling_grid(pm[v][None].float(), (Hs, Ws), (Hg, Wg))[0]
        assert torch.equal(grid.view(torch.int32), og.view(torch.int32))
        src = x[:, v].contiguous()
        out = torch.empty(N, Cc, Hg, Wg)
        cref.warp_ref_forward(_p(src), N, Cc, Hs, Ws, _p(grid), Hg, Wg, _p(out))
        assert torch.equal(out, ref[:, v * Cc:(v + 1) * Cc])
        g = go[:, v * Cc:(v + 1) * Cc].contiguous()
        gin = torch.empty(N, Cc, Hs, Ws)
        cref.warp_ref_backward(_p(g), N, Cc, Hs, Ws, _p(grid), Hg, Wg, _p(gin))
        assert float((gin - gref[:, v]).abs().max()) <= 1e-5 * float(gref[:, v].abs().max())
    cm = torch.empty(2, Hg, Wg)
    cref.warp_ref_coord_map(Hg, Wg, _p(cm))
    assert torch.equal(cm, ref[0, -2:])


def test_product_host_grid_matches_c_restatement(cref):
    pm, src_hw, (Hg, Wg) = small_rig(num_cam=2, seed=3)
    T = P.src_norm_trans_dst_norm(pm, src_hw, (Hg, Wg))
    mine = P.host_grid(T, (Hg, Wg))
    for v in range(2):
        grid = torch.empty(Hg, Wg, 2)
        cref.warp_ref_grid(_p(T[v].contiguous()), Hg, Wg, _p(grid))
        assert torch.equal(grid.view(torch.int32), mine[v].view(torch.int32))
