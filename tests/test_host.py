"""Host-side logic of the product (CPU): the C-ABI library loads and exports what include/mvwarp.h
declares, the host coordinate chain is bit-identical to the oracle's torch op sequence, and the
inverse CSR map is the exact adjoint of the oracle's forward."""
import ctypes
import os
import re

import pytest
import torch

from mvdet_b200 import _lib, plan as P
from mvdet_b200.geometry import MULTIVIEWX, RIG32, WILDTRACK, create_coord_map, synthetic_rig
from oracle import kornia_warp as K
from tests.util import small_rig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mvwarp.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    dev_part = "".join(re.findall(r"#ifdef MVWARP_DEV(.*?)#endif", header, flags=re.S))
    header = re.sub(r"#ifdef MVWARP_DEV.*?#endif", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mvwarp_[a-z_0-9]+)\s*\(", header))
    declared_dev = set(re.findall(r"\b(mvwarp_[a-z_0-9]+)\s*\(", dev_part))
    assert len(declared) >= 10
    lib = ctypes.CDLL(os.path.join(ROOT, "mvdet_b200", "libmvwarp.so"))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mvwarp.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert declared_dev == set(_lib.DEV_SIGNATURES)
    for name in declared_dev:                     # validation hooks live in the development build only
        assert not hasattr(lib, name), f"{name} is a development hook but the product library exports it"
    assert _lib.load().mvwarp_version() == 1


def test_stats_struct_layout_matches_header():
    # int32 x6, uint64 x3, uint32 x4, int32 x2, uint32 x2, uint64: 24 + 24 + 16 + 8 + 8 + 8
    assert ctypes.sizeof(_lib.Stats) == 88


def test_bad_arguments_return_error_codes_not_crashes():
    lib = _lib.load()
    assert lib.mvwarp_host_grid(None, 1, 4, 4, None) == _lib.EINVAL
    assert b"host_grid" in lib.mvwarp_last_error()
    st = _lib.Stats()
    assert lib.mvwarp_host_csr(None, 1, 4, 4, 4, 4, ctypes.byref(st), None, None, None, None, 0) == _lib.EINVAL
    handle = ctypes.c_void_p()
    assert lib.mvwarp_plan_create(None, 1, 4, 4, 4, 4, 0, ctypes.byref(handle)) == _lib.EINVAL
    assert lib.mvwarp_plan_destroy(None) == _lib.OK
    with pytest.raises(_lib.MvwarpError):
        _lib.check(_lib.EINVAL, "demo")


def test_every_entry_point_rejects_null_arguments_without_a_gpu():
    """include/mvwarp.h: "returns a negative MVWARP_E* code on failure, and never throws or aborts across the boundary" -- every
    compute entry point, called with null pointers / impossible sizes on a box with no GPU, must come back with EINVAL and a message."""
    lib = _lib.load()
    i5, i4 = _lib.i64x([1] * 5), _lib.i64x([1] * 4)
    st, h = _lib.Stats(), ctypes.c_void_p()
    nine = ctypes.cast((ctypes.c_float * 9)(), ctypes.POINTER(ctypes.c_float))
    calls = {
        "warp_views_fwd": lambda: lib.mvwarp_warp_views_fwd(None, 0, 1, None, i5, None, i4, 0, -1, 1, 4, 0, None),
        "warp_views_bwd": lambda: lib.mvwarp_warp_views_bwd(None, 0, 1, None, i4, 0, None, i5, 1, 4, 0, None),
        "plan_stats": lambda: lib.mvwarp_plan_stats(None, ctypes.byref(st)),
        "plan_save": lambda: lib.mvwarp_plan_save(None, b"/tmp/x"),
        "plan_load": lambda: lib.mvwarp_plan_load(b"/nonexistent/x.mvwplan", 0, ctypes.byref(h)),
        "plan_matrices": lambda: lib.mvwarp_plan_matrices(None, None),
        "plan_device_grid": lambda: lib.mvwarp_plan_device_grid(None, None, None),
        "warp_conv1_fwd": lambda: lib.mvwarp_warp_conv1_fwd(None, 0, 1, None, i5, None, None, None, i4, 1, 64, 1, 1, 0, 0, None),
        "warp_conv1_fwd_ex": lambda: lib.mvwarp_warp_conv1_fwd_ex(None, 0, 1, None, i5, None, None, None, 0, i4, 1, 64, 1, 1, 0, 0,
                                                                  None, 0, None),
        "conv3x3_fwd": lambda: lib.mvwarp_conv3x3_fwd(None, i4, None, None, None, i4, 1, 8, 8, 64, 1, 1, 0, 0, None),
        "conv3x3_head_fwd": lambda: lib.mvwarp_conv3x3_head_fwd(None, i4, None, None, i4, 1, 8, 8, 64, 1, 0, None),
        "conv3x3_head_fwd_ex": lambda: lib.mvwarp_conv3x3_head_fwd_ex(None, i4, None, None, i4, 1, 8, 8, 64, 1, None, 0, 0, None),
        "host_upsample_windows": lambda: lib.mvwarp_host_upsample_windows(nine, 1, 2, 2, 4, 4, 4, 4, None, None, None, None),
        "inbox_reduce": lambda: lib.mvwarp_inbox_reduce(None, 2, 0, 0, 1, 512, None, 0, None, 0, 0, None),
        "warp_conv1_fwd_p2p": lambda: lib.mvwarp_warp_conv1_fwd_p2p(None, 0, 1, None, i5, None, 1, 64, 1, 2, 0, None, None, 8, None),
        "warp_views_fwd_list": lambda: lib.mvwarp_warp_views_fwd_list(None, 0, 1, None, i4, None, i4, 0, -1, 1, 4, 0, None),
        "warp_views_bwd_list": lambda: lib.mvwarp_warp_views_bwd_list(None, 0, 1, None, i4, 0, None, i4, 1, 4, 0, None),
        "relu_bwd_cast": lambda: lib.mvwarp_relu_bwd_cast(None, i4, None, i4, None, i4, None, None, 0, 1, 4, 4, 512, 0, None),
        "layout_transpose": lambda: lib.mvwarp_layout_transpose(None, 0, None, 0, 1, 4, 4, 0, 1, 0, None),
        "copy_touched_texels": lambda: lib.mvwarp_copy_touched_texels(None, 0, 1, None, i5, None, i5, 1, 4, 0, None),
        "plan_touched": lambda: lib.mvwarp_plan_touched(None, None, None, 0),
        "plan_create": lambda: lib.mvwarp_plan_create(nine, 1, 0, 4, 4, 4, 0, ctypes.byref(h)),
        "plan_create_upsampled": lambda: lib.mvwarp_plan_create_upsampled(nine, 1, 0, 0, 4, 4, 4, 4, 0, ctypes.byref(h)),
        "plan_create_from_grid": lambda: lib.mvwarp_plan_create_from_grid(None, 1, 4, 4, 4, 4, 0, ctypes.byref(h)),
    }
    for name, call in calls.items():
        assert call() == _lib.EINVAL, name
        assert len(lib.mvwarp_last_error()) > 0, name
    assert lib.mvwarp_warp_conv1_scratch_bytes(None, 1, 1, 0, 0) == 0
    assert lib.mvwarp_conv3x3_head_scratch_bytes(0, 8, 8) == 0 and lib.mvwarp_conv3x3_head_scratch_bytes(2, 8, 8) == 2 * 64 * 9 * 4


def test_matrix_chain_matches_oracle_bits():
    for spec in (WILDTRACK, MULTIVIEWX):
        pm = synthetic_rig(spec)
        T = P.src_norm_trans_dst_norm(pm, spec.feat_shape, spec.grid_shape)
        for v in range(spec.num_cam):
            ref = K._torch_inverse_cast(K.normalize_homography(pm[v][None].float(), spec.feat_shape, spec.grid_shape))[0]
            assert torch.equal(T[v], ref)


@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX, RIG32], ids=lambda s: s.name)
def test_host_grid_bit_identical_to_oracle(spec):
    pm = synthetic_rig(spec)
    T = P.src_norm_trans_dst_norm(pm, spec.feat_shape, spec.grid_shape)
    g = P.host_grid(T, spec.grid_shape)
    og = K.sampling_grids(pm, spec.feat_shape, spec.grid_shape)
    assert torch.equal(g.view(torch.int32), og.view(torch.int32))


# U / nnz / max_row measured independently during the survey (SURVEY.md 7 hard part 4, 8d)
KNOWN = {"wildtrack": (237509, 980278, 95), "multiviewx": (174222, 688362, 84), "rig32": (1702864, 67136331, 1360)}


@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX, RIG32], ids=lambda s: s.name)
def test_csr_statistics_match_survey(spec):
    pm = synthetic_rig(spec)
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, spec.feat_shape, spec.grid_shape), spec.grid_shape)
    st = P.host_csr(g, spec.feat_shape)
    assert (st["unique_texels"], st["nnz"], st["max_row"]) == KNOWN[spec.name]


def _csr_backward(tex_of, row_ptr, cells, weights, grad_out_v, ntex):
    """grad_in[c, tex_of[row]] = sum_k w_k * grad_out[c, cell_k], rows in CSR order (dense torch emulation)."""
    C = grad_out_v.shape[0]
    lens = (row_ptr[1:] - row_ptr[:-1]).long()
    rows = torch.repeat_interleave(tex_of.long(), lens)
    contrib = grad_out_v.reshape(C, -1)[:, cells.long()] * weights[None]
    out = torch.zeros(C, ntex, dtype=grad_out_v.dtype)
    out.index_add_(1, rows, contrib)
    return out


def test_csr_is_adjoint_of_oracle_forward():
    pm, src_hw, grid_hw = small_rig()
    V, C = pm.shape[0], 5
    T = P.src_norm_trans_dst_norm(pm, src_hw, grid_hw)
    g = P.host_grid(T, grid_hw)
    st, tex_of, row_ptr, cells, weights = P.host_csr(g, src_hw, with_map=True)
    assert int(row_ptr[:, -1].sum()) == st["nnz"]
    ntex_ = src_hw[0] * src_hw[1]
    assert torch.equal(torch.sort(tex_of[0].long()).values, torch.arange(ntex_))   # a permutation of the texels
    torch.manual_seed(0)
    x = torch.randn(1, V, C, *src_hw, requires_grad=True)
    out = K.warp_views_oracle(x, pm, grid_hw)
    go = torch.randn_like(out)
    (gx,) = torch.autograd.grad(out, x, go)
    ntex = src_hw[0] * src_hw[1]
    base = 0
    for v in range(V):
        n = int(row_ptr[v, -1])
        mine = _csr_backward(tex_of[v], row_ptr[v], cells[base:base + n], weights[base:base + n], go[0, v * C:(v + 1) * C], ntex)
        base += n
        ref = gx[0, v].reshape(C, ntex)
        assert float((mine - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    # row entries are sorted by cell (fixed summation order = determinism)
    r = row_ptr[0].long()
    c0 = cells[: int(r[-1])].long()
    for t in torch.nonzero(r[1:] - r[:-1] > 1)[:200, 0]:
        seg = c0[r[t]:r[t + 1]]
        assert bool((seg[1:] >= seg[:-1]).all())


def test_csr_handles_degenerate_grids():
    # all out of range, NaN and inf coordinates: no taps, no crash
    g = torch.full((1, 4, 5, 2), 7.0)
    g[0, 0, 0] = float("nan")
    g[0, 0, 1] = float("inf")
    g[0, 0, 2] = -float("inf")
    st = P.host_csr(g, (6, 6))
    assert st["nnz"] == 0 and st["unique_texels"] == 0 and st["valid_cells"] == 0
    # exact corner hit: the four taps of the last texel, three of them outside
    g = torch.ones((1, 1, 1, 2))
    st, tex_of, row_ptr, cells, weights = P.host_csr(g, (6, 6), with_map=True)
    assert st["nnz"] == 1 and float(weights[0]) == 1.0 and int(row_ptr[0, -1]) == 1
    assert int(tex_of[0][int(torch.nonzero(row_ptr[0, 1:] - row_ptr[0, :-1])[0, 0])]) == 35   # the last texel


def test_backward_task_decomposition_covers_every_row_once():
    """Plan-time load balancing: light rows are packed into warp tasks, heavy rows get a CTA each."""
    pm = synthetic_rig(WILDTRACK)
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, WILDTRACK.feat_shape, WILDTRACK.grid_shape), WILDTRACK.grid_shape)
    st, tex_of, row_ptr, cells, weights = P.host_csr(g, WILDTRACK.feat_shape, with_map=True)
    lens = (row_ptr[:, 1:] - row_ptr[:, :-1]).long()
    heavy = int((lens > 31).sum())
    assert st["heavy_rows"] == heavy and heavy > 0
    # rows of zero segments (all-empty runs of 8 texels) are in no task; what is left weighs 1 per row + 1 per entry
    light_weight = int((lens[lens <= 31] + 1).sum()) - st["zero_rows"]
    # every task weighs at most 32 (rows + entries): the task count is bounded both ways (a task ends at every heavy row and at
    # every zero segment that interrupts a run)
    assert st["bwd_tasks"] >= light_weight / 32
    assert st["bwd_tasks"] <= light_weight / 8 + 2 * heavy + 7 * 2040 + st["zero_segments"]


@pytest.mark.parametrize("Wg", [2, 44, 45, 64])
def test_host_grid_follows_torch_bmm_dispatch(Wg):
    """torch.bmm switches from its scalar small-matrix kernel to MKL at 9*Wg >= 400; the chain mirrors both."""
    pm, src_hw, _ = small_rig(num_cam=2, seed=4)
    grid_hw = (7, Wg)
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, src_hw, grid_hw), grid_hw)
    og = K.sampling_grids(pm, src_hw, grid_hw)
    assert torch.equal(g.view(torch.int32), og.view(torch.int32))


def test_coord_map_matches_reference_formula():
    cm = create_coord_map((120, 360))
    assert torch.equal(cm, K.create_coord_map([120, 360, 1]))


def test_composed_upsample_map_is_adjoint_of_two_step_oracle():
    """Fused up-sample plan: the composed CSR map on the low-res image reproduces the gradient of
    warp(F.interpolate(x, bilinear)) -- the reference's persp_trans_detector.py:65 followed by :68-69."""
    import torch.nn.functional as F
    pm, src_hw, grid_hw = small_rig()
    low_hw = (src_hw[0] // 3, src_hw[1] // 3)
    V, C = pm.shape[0], 4
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, src_hw, grid_hw), grid_hw)
    st, tex_of, row_ptr, cells, weights = P.host_csr(g, src_hw, with_map=True, lowres_hw=low_hw)
    assert st["lowres_h"] == low_hw[0] and st["fwd_tasks"] > 0
    torch.manual_seed(0)
    x = torch.randn(1, V, C, *low_hw, requires_grad=True)
    up = torch.stack([F.interpolate(x[:, v], src_hw, mode="bilinear") for v in range(V)], dim=1)
    out = K.warp_views_oracle(up, pm, grid_hw)
    go = torch.randn_like(out)
    (gx,) = torch.autograd.grad(out, x, go)
    nlow = low_hw[0] * low_hw[1]
    base = 0
    for v in range(V):
        n = int(row_ptr[v, -1])
        mine = _csr_backward(tex_of[v], row_ptr[v], cells[base:base + n], weights[base:base + n], go[0, v * C:(v + 1) * C], nlow)
        base += n
        ref = gx[0, v].reshape(C, nlow)
        assert float((mine - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    # at most 16 low-res taps per cell, typically far fewer
    assert st["nnz"] <= 16 * st["valid_cells"]


def test_conv1_weight_packing_layout():
    """pack_conv1_weight (host side of the fused warp->conv1 kernel): block order (view, channel block, tap), K-major rows,
    16-byte chunks XOR-swizzled by (row & 7), coord channels of all nine taps in the last block."""
    from mvdet_b200.conv1 import pack_conv1_weight
    V, C = 2, 128
    g = torch.Generator().manual_seed(0)
    w = torch.randn(512, V * C + 2, 3, 3, generator=g)
    packed = pack_conv1_weight(w, V, C)
    assert packed.shape == (V * (C // 64) * 9 + 1, 512, 64) and packed.dtype == torch.bfloat16
    wb = w.bfloat16()
    for (v, cb, tap, n, k) in [(0, 0, 0, 0, 0), (1, 1, 5, 17, 33), (0, 1, 8, 511, 63), (1, 0, 4, 260, 9)]:
        block = (v * (C // 64) + cb) * 9 + tap
        chunk, j = k // 8, k % 8
        pos = chunk ^ (n & 7)
        assert packed[block, n, pos * 8 + j] == wb[n, v * C + cb * 64 + k, tap // 3, tap % 3]
    last = packed[-1]
    for (n, tap, ch) in [(0, 0, 0), (5, 8, 1), (300, 4, 0)]:
        k = tap * 2 + ch
        pos = (k // 8) ^ (n & 7)
        assert last[n, pos * 8 + k % 8] == wb[n, V * C + ch, tap // 3, tap % 3]
    # everything beyond k = 18 in the coord block is zero
    unsw = torch.stack([last[n].reshape(8, 8)[[c ^ (n & 7) for c in range(8)]].reshape(64) for n in range(16)])
    assert bool((unsw[:, 18:] == 0).all())
    with pytest.raises(ValueError):
        pack_conv1_weight(w, V, C + 64)


def test_conv1_weight_roundtrip_and_reference_state_dict():
    """the reference checkpoint layout (main.py:121/125: map_classifier.0.weight [512, V*C+2, 3, 3],
    .bias [512]) goes into the packed GEMM layout and comes back exactly (bf16 values)."""
    from mvdet_b200.conv1 import conv1_from_state_dict, conv1_to_state_dict, pack_conv1_weight, unpack_conv1_weight
    V, C = 3, 64
    g = torch.Generator().manual_seed(3)
    sd = {"map_classifier.0.weight": torch.randn(512, V * C + 2, 3, 3, generator=g),
          "map_classifier.0.bias": torch.randn(512, generator=g),
          "map_classifier.2.weight": torch.randn(4, 4)}
    wpack, bias = conv1_from_state_dict(sd, V, C)
    assert torch.equal(wpack, pack_conv1_weight(sd["map_classifier.0.weight"], V, C))
    assert torch.equal(unpack_conv1_weight(wpack, V, C), sd["map_classifier.0.weight"].bfloat16())
    back = conv1_to_state_dict(wpack, bias, V, C)
    assert torch.equal(back["map_classifier.0.weight"], sd["map_classifier.0.weight"].bfloat16().float())
    assert torch.equal(back["map_classifier.0.bias"], sd["map_classifier.0.bias"])
    # DataParallel-style keys, view shards without the coord block
    wp2, _ = conv1_from_state_dict({"module." + k: v for k, v in sd.items()}, V, C)
    assert torch.equal(wp2, wpack)
    shard = pack_conv1_weight(sd["map_classifier.0.weight"][:, C:2 * C], 1, C, with_coord=False)
    assert torch.equal(unpack_conv1_weight(shard, 1, C, with_coord=False), sd["map_classifier.0.weight"][:, C:2 * C].bfloat16())
    with pytest.raises(KeyError):
        conv1_from_state_dict({}, V, C)
    with pytest.raises(ValueError):
        unpack_conv1_weight(wpack, V + 1, C)


def test_plan_load_refuses_bad_files(tmp_path):
    """mvwarp_plan_load must fail cleanly (error code + message, no CUDA call) on missing, foreign and truncated files."""
    import ctypes as C
    from mvdet_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.mvwarp_plan_load(str(tmp_path / "missing.mvwplan").encode(), 0, C.byref(h)) == _lib.EINVAL
    assert b"cannot open" in lib.mvwarp_last_error()
    junk = tmp_path / "junk.mvwplan"
    junk.write_bytes(b"not a plan" * 100)
    assert lib.mvwarp_plan_load(str(junk).encode(), 0, C.byref(h)) == _lib.EINVAL
    assert b"not a plan file" in lib.mvwarp_last_error()
    short = tmp_path / "short.mvwplan"
    short.write_bytes(b"MVWPLAN1" + b"\0" * 16)
    assert lib.mvwarp_plan_load(str(short).encode(), 0, C.byref(h)) == _lib.EINVAL
    assert not h.value
    assert lib.mvwarp_plan_save(None, b"/tmp/x") == _lib.EINVAL
    assert lib.mvwarp_plan_matrices(None, None) == _lib.EINVAL


def test_map_head_weight_packing():
    """map_head.pack_conv3x3_weight / pack_head_weight: layouts of map_classifier[2] / [4] (persp_trans_detector.py:53-54)."""
    from mvdet_b200.map_head import pack_conv3x3_weight, pack_head_weight
    g = torch.Generator().manual_seed(2)
    w2 = torch.randn(512, 128, 3, 3, generator=g)
    p2 = pack_conv3x3_weight(w2)
    assert p2.shape == (2 * 9, 512, 64) and p2.dtype == torch.bfloat16
    wb = w2.bfloat16()
    for (cb, tap, n, k) in [(0, 0, 0, 0), (1, 7, 300, 41), (0, 4, 511, 63)]:
        pos = (k // 8) ^ (n & 7)
        assert p2[cb * 9 + tap, n, pos * 8 + k % 8] == wb[n, cb * 64 + k, tap // 3, tap % 3]
    w3 = torch.randn(1, 512, 3, 3, generator=g)
    p3 = pack_head_weight(w3)
    assert p3.shape == (9, 512) and p3.dtype == torch.bfloat16
    assert p3[5, 17] == w3.bfloat16()[0, 17, 1, 2]
    with pytest.raises(ValueError):
        pack_conv3x3_weight(torch.randn(512, 100, 3, 3))
    with pytest.raises(ValueError):
        pack_head_weight(torch.randn(2, 512, 3, 3))


def test_install_kornia_shim_without_kornia(monkeypatch):
    """Seam 1 as shipped code: with no kornia installed the shim registers stub modules, so the reference's `import kornia`
    and both spellings of the call (persp_trans_detector.py:69, image_proj_variant.py:69) reach the given function."""
    import importlib
    import sys
    from mvdet_b200 import integration
    if importlib.util.find_spec("kornia") is not None:
        pytest.skip("kornia is installed here; the stub branch is not taken")
    for name in ("kornia", "kornia.geometry", "kornia.geometry.transform"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    marker = lambda src, M, dsize, *a, **k: ("warped", tuple(dsize))
    try:
        mod = integration.install_kornia_shim(warp=marker)
        import kornia
        import kornia.geometry.transform as kgt
        assert kornia is mod and kgt.warp_perspective is marker and kornia.warp_perspective is marker
        assert kornia.geometry.transform.warp_perspective(None, None, [3, 4]) == ("warped", (3, 4))
        integration.install_kornia_shim(warp=len)          # idempotent: re-pointing re-uses the registered modules
        assert sys.modules["kornia"] is mod and kornia.warp_perspective is len
    finally:
        for name in ("kornia", "kornia.geometry", "kornia.geometry.transform"):
            sys.modules.pop(name, None)


def test_install_kornia_shim_defaults_to_the_cuda_seam(monkeypatch):
    import sys
    from mvdet_b200 import integration, ops
    for name in ("kornia", "kornia.geometry", "kornia.geometry.transform"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    try:
        mod = integration.install_kornia_shim()
        assert mod.geometry.transform.warp_perspective is ops.warp_perspective
    finally:
        for name in ("kornia", "kornia.geometry", "kornia.geometry.transform"):
            sys.modules.pop(name, None)


def test_fused_head_cache_repacks_when_the_weights_change():
    """patch_detector(fused_head=True) must not serve a weight snapshot taken at patch time.  The cache re-packs
    after an optimizer step (in-place update), after load_state_dict (copy_) and after a device / storage change, and only then."""
    import torch.nn as nn
    from mvdet_b200.integration import FusedHeadCache

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.num_cam = 1
            self.map_classifier = nn.Sequential(nn.Conv2d(6, 4, 3, padding=1), nn.ReLU(), nn.Conv2d(4, 1, 3, padding=1, bias=False))

    model = M()
    cache = FusedHeadCache(model, factory=lambda sd: {k: v.clone() for k, v in sd.items()})
    h0 = cache.get()
    assert cache.get() is h0 and cache.rebuilds == 1                       # unchanged weights: no re-pack
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    model.map_classifier(torch.randn(1, 6, 5, 5)).sum().backward()
    opt.step()
    h1 = cache.get()
    assert h1 is not h0 and cache.rebuilds == 2
    assert torch.equal(h1["map_classifier.0.weight"], model.map_classifier[0].weight.detach())
    assert not torch.equal(h0["map_classifier.0.weight"], h1["map_classifier.0.weight"])
    sd = {k: torch.zeros_like(v) for k, v in model.state_dict().items()}
    model.load_state_dict(sd)                                               # main.py:125
    h2 = cache.get()
    assert cache.rebuilds == 3 and float(h2["map_classifier.2.weight"].abs().sum()) == 0.0
    model.map_classifier[0].weight.data = model.map_classifier[0].weight.data.clone()   # new storage (what .to(device) does)
    cache.get()
    assert cache.rebuilds == 4
    assert cache.get() is cache.get() and cache.rebuilds == 4


def test_plan_table_holds_no_strong_references():
    """ops._plans_by_id must not keep plans alive (it is a WeakValueDictionary)."""
    import weakref
    from mvdet_b200 import ops
    assert isinstance(ops._plans_by_id, weakref.WeakValueDictionary)


def test_weight_gradient_as_gemms_on_shifted_views_is_the_convolution_weight_gradient():
    """The identity conv1_wgrad_gemm rests on, checked on the CPU in fp32: with X in the interior of a zero-bordered
    buffer Xp[B,H+2,W+2,P] and g in the top-left corner of an equally bordered Gp, tap (ky,kx) is the constant shift ky*(W+2)+kx of the
    flattened position index, so gW[:,:,ky,kx] = Gp[:K].T @ Xp[shift:shift+K] -- including the overlapping batched view used for kx."""
    import torch.nn.functional as F
    g0 = torch.Generator().manual_seed(0)
    B, Cin, Cout, H, W = 2, 5, 4, 6, 9
    X = torch.randn(B, Cin, H, W, generator=g0)
    g = torch.randn(B, Cout, H, W, generator=g0)
    ref = torch.nn.grad.conv2d_weight(X, (Cout, Cin, 3, 3), g, padding=1)
    P_ = 8
    Xp = torch.zeros(B, H + 2, W + 2, P_)
    Xp[:, 1:H + 1, 1:W + 1, :Cin] = X.permute(0, 2, 3, 1)
    Gp = torch.zeros(B, H + 2, W + 2, Cout)
    Gp[:, :H, :W] = g.permute(0, 2, 3, 1)
    Xf, Gf = Xp.view(-1, P_), Gp.view(-1, Cout)
    K = Xf.shape[0] - (2 * (W + 2) + 2)
    Gt = Gf[:K].t()
    rows = []
    for ky in range(3):
        Xk = torch.as_strided(Xf, (3, K, P_), (P_, P_, 1), Xf.storage_offset() + ky * (W + 2) * P_)   # the three kx taps: one bmm
        rows.append(torch.bmm(Gt.unsqueeze(0).expand(3, Cout, K), Xk))
    gw = torch.stack(rows).view(3, 3, Cout, P_)[..., :Cin].permute(2, 3, 0, 1)
    assert torch.allclose(gw, ref, rtol=1e-5, atol=1e-5)
    assert float(torch.stack(rows).view(3, 3, Cout, P_)[..., Cin:].abs().max()) == 0.0     # pad channels stay zero


def test_data_gradient_weights_are_the_transposed_flipped_convolution():
    """pack_conv1_dgrad_weight: the gradient w.r.t. view v's warped channels is conv3x3(g, W'_v, padding=1) with
    W'_v[c,n,ky,kx] = W[n, v*C+c, 2-ky, 2-kx] -- checked against torch's conv2d_input on the CPU through unpack_conv1_weight."""
    import torch.nn.functional as F
    from mvdet_b200 import conv1
    g0 = torch.Generator().manual_seed(1)
    V, Cc, H, W = 2, 512, 5, 7
    weight = (torch.randn(512, V * Cc + 2, 3, 3, generator=g0) * 0.05).bfloat16().float()
    g = torch.randn(1, 512, H, W, generator=g0)
    ref = torch.nn.grad.conv2d_input((1, V * Cc + 2, H, W), weight, g, padding=1)
    packs = conv1.pack_conv1_dgrad_weight(weight, V, Cc)
    assert len(packs) == V
    for v, wp in enumerate(packs):
        wt = conv1.unpack_conv1_weight(wp, 1, 512, with_coord=False).float()        # [c, n, 3, 3]
        got = F.conv2d(g, wt, None, padding=1)
        assert torch.allclose(got, ref[:, v * Cc:(v + 1) * Cc], rtol=1e-4, atol=1e-4)
    with pytest.raises(ValueError):
        conv1.pack_conv1_dgrad_weight(torch.randn(512, 2 * 64 + 2, 3, 3), 2, 64)   # the tcgen05 convolution has 512 outputs per view


def test_row_bands_with_more_ranks_than_tile_rows():
    from mvdet_b200 import sharding as S
    assert S.row_bands(24, 4) == [0, 8, 16, 24, 24]           # 3 tile rows over 4 ranks: the last rank owns an empty band
    assert S.row_bands(20, 2) == [0, 16, 20]
    e = S.row_bands(480, 8)
    assert e[0] == 0 and e[-1] == 480 and all(b >= a and a % 8 == 0 for a, b in zip(e[:-1], e[1:]))


def test_new_ops_refuse_cpu_tensors_and_bad_arguments_without_a_gpu():
    """warp_views_list / copy_touched_texels: CUDA only, kornia-style argument errors first -- all checkable on a box with no GPU."""
    import mvdet_b200 as mv
    from mvdet_b200 import ops
    pm, src_hw, grid_hw = small_rig(num_cam=2, seed=3)
    cams = [torch.zeros(1, 4, *src_hw) for _ in range(2)]
    with pytest.raises(mv.MvwarpError):
        ops.warp_views_list(cams, pm, grid_hw)                 # CPU tensors: no fallback
    with pytest.raises(ValueError):
        ops.warp_views_list([], pm, grid_hw)
    with pytest.raises(TypeError):
        ops.warp_views_list(["x", cams[0]], pm, grid_hw)
    assert _lib.load().mvwarp_layout_transpose(None, 0, None, 0, 1, 4, 4, 0, 1, 0, None) == _lib.EINVAL
    assert _lib.load().mvwarp_relu_bwd_scratch_floats(0, 0) == 0


def _apply_upsample_windows(low, origin, mask, wx, wy):
    """out[v,c,y,x] = sum_{j,i<3} wy[j] wx[i] low[v,c,oy+j,ox+i] over the live slots, in float64 (low [V,C,hl,wl])."""
    V, Cc, hl, wl = low.shape
    out = torch.zeros(V, Cc, *origin.shape[1:3], dtype=torch.float64)
    for v in range(V):
        for j in range(3):
            for i in range(3):
                live = ((mask[v] >> (j * 3 + i)) & 1).bool()
                yy, xx = origin[v, ..., 1].long() + j, origin[v, ..., 0].long() + i
                # a live slot always addresses a texel of the low-res map (the kernel gathers it without a bounds test)
                assert bool(((yy[live] >= 0) & (yy[live] < hl) & (xx[live] >= 0) & (xx[live] < wl)).all())
                w = wy[v, ..., j].double() * wx[v, ..., i].double() * live
                out[v] += low[v].double()[:, yy.clamp(0, hl - 1), xx.clamp(0, wl - 1)] * w
    return out


@pytest.mark.parametrize("low_hw", [(18, 32), (27, 48), (54, 96), (20, 37), (1, 1)])
def test_fused_conv1_upsample_windows_equal_the_two_step_oracle(low_hw):
    """Fused warp -> conv1 over fused up-sample plans: the 3 x 3 low-res window the kernel's producers compose per halo cell
    (csrc/mvwarp_uptaps.h, run here on the host through mvwarp_host_upsample_windows -- the same source) reproduces
    warp(F.interpolate(low, bilinear)) -- persp_trans_detector.py:65 followed by :68-69 -- for integer, non-integer and 1:1 ratios."""
    import torch.nn.functional as F
    pm, src_hw, grid_hw = small_rig()
    V, Cc = pm.shape[0], 3
    T = P.src_norm_trans_dst_norm(pm, src_hw, grid_hw)
    origin, mask, wx, wy = P.host_upsample_windows(T, low_hw, src_hw, grid_hw)
    torch.manual_seed(1)
    low = torch.randn(V, Cc, *low_hw)
    mine = _apply_upsample_windows(low, origin, mask, wx, wy)
    up = F.interpolate(low, src_hw, mode="bilinear")
    ref = K.warp_views_oracle(up.reshape(1, V, Cc, *src_hw), pm, grid_hw)[0, :V * Cc].reshape(V, Cc, *grid_hw)
    assert float(ref.abs().max()) > 0.5                                    # the rig sees the grid
    assert float((mine - ref.double()).abs().max()) <= 1e-5 * float(ref.abs().max())
    # cells the warp leaves empty have no live slot, and every other cell has at least one
    empty = (ref == 0).all(dim=1)
    assert bool((mask[empty] == 0).all()) and bool((mask[~empty] != 0).all())


def test_upsample_windows_reject_a_down_sample():
    pm, src_hw, grid_hw = small_rig()
    T = P.src_norm_trans_dst_norm(pm, src_hw, grid_hw)
    with pytest.raises(_lib.MvwarpError):
        P.host_upsample_windows(T, (src_hw[0] + 1, src_hw[1]), src_hw, grid_hw)


@pytest.mark.parametrize("spec", [WILDTRACK, MULTIVIEWX], ids=lambda s: s.name)
def test_zero_segments_cover_the_area_outside_the_footprints(spec):
    """Backward work decomposition: runs of 8 consecutive texels (one image row of one 8x8 tile) with no tap at all are zero segments
    -- streamed by dedicated CTAs, in no task.  The builder checks on the host that every row is in exactly one task, heavy row or
    zero segment (a plan that left a row out would raise here); the statistics must add up, for plain and fused up-sample plans."""
    pm = synthetic_rig(spec)
    g = P.host_grid(P.src_norm_trans_dst_norm(pm, spec.feat_shape, spec.grid_shape), spec.grid_shape)
    total = spec.num_cam * spec.feat_shape[0] * spec.feat_shape[1]
    st = P.host_csr(g, spec.feat_shape)
    assert st["zero_rows"] <= total - st["unique_texels"]
    assert st["zero_rows"] >= 0.5 * total                         # most of a feature map lies outside the ground rectangle
    assert st["zero_segments"] * 8 >= st["zero_rows"] > st["zero_segments"]
    low = (spec.feat_shape[0] // 3, spec.feat_shape[1] // 3)
    st_up = P.host_csr(g, spec.feat_shape, lowres_hw=low)
    assert 0 < st_up["zero_rows"] <= spec.num_cam * low[0] * low[1] - st_up["unique_texels"]
