"""Golden vectors (tests/golden/warp_small.npz, made by tests/golden/make_golden.py from the oracle)."""
import os

import numpy as np
import pytest
import torch

from mvdet_b200 import plan as P
from oracle import kornia_warp as K
from tests.util import rel_errors, to_channels_last_5d

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_small.npz"))


def _t(name):
    return torch.from_numpy(G[name])


def test_oracle_reproduces_golden_vectors():
    pm, x = _t("proj_mats"), _t("feats").requires_grad_(True)
    grid_shape = tuple(int(v) for v in G["grid_shape"])
    out = K.warp_views_oracle(x, pm, grid_shape)
    (gx,) = torch.autograd.grad(out, x, _t("grad_out"))
    # same torch build: bit-identical; another build / CPU may round the 3x3 inverse differently
    assert max(rel_errors(out.detach(), _t("out"))) <= 1e-5
    assert max(rel_errors(gx, _t("grad_feats"))) <= 1e-5
    if str(G["torch_version"]) == torch.__version__:
        assert torch.equal(out.detach(), _t("out"))


def test_host_coordinate_chain_reproduces_golden_grid():
    pm = _t("proj_mats")
    src_hw = tuple(int(v) for v in G["feat_shape"])
    grid_shape = tuple(int(v) for v in G["grid_shape"])
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, src_hw, grid_shape), grid_shape)
    assert float((g - _t("grid")).abs().max()) <= 1e-5
    if str(G["torch_version"]) == torch.__version__:
        assert torch.equal(g.view(torch.int32), _t("grid").view(torch.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["cl", "nchw"])
def test_cuda_path_reproduces_golden_vectors(layout):
    import mvdet_b200 as mv
    pm = _t("proj_mats")
    grid_shape = tuple(int(v) for v in G["grid_shape"])
    x = _t("feats")
    xd = (to_channels_last_5d(x) if layout == "cl" else x).to("cuda:0").requires_grad_(True)
    out = mv.warp_views(xd, pm, grid_shape)
    (gx,) = torch.autograd.grad(out, xd, _t("grad_out").to("cuda:0"))
    assert max(rel_errors(out.detach(), _t("out"))) <= 1e-5
    assert max(rel_errors(gx, _t("grad_feats"))) <= 1e-5
