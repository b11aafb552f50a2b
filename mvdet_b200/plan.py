"""Rig-constant warp plan (host side).

The reference recomputes, on every forward and for every camera, the chain
``proj_mat.float() -> normalize_homography -> inverse -> meshgrid -> transform_points``
(/root/reference/multiview_detector/models/persp_trans_detector.py:68-69, kornia 0.6.11 internals).
None of it depends on the features, so here it runs once per camera rig:

  * :func:`src_norm_trans_dst_norm` does the 3x3 part with the same torch CPU fp32 ops kornia uses
    (so the matrix bits are the reference's), and
  * libmvwarp.so does the per-cell part (in-kernel, mvwarp_coords.h) and builds the inverse CSR
    map for the deterministic backward.

Plans are cached per (matrix bytes, shapes, device).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import threading
import weakref
from typing import Dict, Sequence, Tuple

import torch

from . import _lib


def _pixel_to_norm(height: int, width: int) -> torch.Tensor:
    """[-1,1] normalisation of pixel coordinates, fp32, built the way kornia builds it
    (entries scaled in place: ``1 * 2.0 / (size - 1)`` in fp32; 1e-14 guards size 1)."""
    m = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
    m[0, 0] = m[0, 0] * 2.0 / (1e-14 if width == 1 else width - 1.0)
    m[1, 1] = m[1, 1] * 2.0 / (1e-14 if height == 1 else height - 1.0)
    return m.unsqueeze(0)


def src_norm_trans_dst_norm(proj_mats: torch.Tensor, src_hw: Sequence[int], dst_hw: Sequence[int]) -> torch.Tensor:
    """fp32 [V,3,3]: normalised destination coords -> normalised source coords, per camera.

    ``proj_mats`` [V,3,3] maps source pixel -> destination pixel (float64 in the reference; it is
    rounded to fp32 first, as persp_trans_detector.py:68 does).  Same op order as kornia's
    ``normalize_homography`` followed by ``_torch_inverse_cast``:
    ``inverse(N_dst @ (M @ inverse(N_src)))``, all on CPU in fp32.
    """
    if proj_mats.dim() != 3 or tuple(proj_mats.shape[-2:]) != (3, 3):
        raise ValueError(f"Input M must be a Bx3x3 tensor. Got {tuple(proj_mats.shape)}")
    M = proj_mats.detach().to("cpu").float()
    n_src = _pixel_to_norm(int(src_hw[0]), int(src_hw[1]))
    n_dst = _pixel_to_norm(int(dst_hw[0]), int(dst_hw[1]))
    out = []
    for v in range(M.shape[0]):           # one camera at a time, exactly like the call site's loop
        m = M[v:v + 1]
        dst_norm_trans_src_norm = n_dst @ (m @ torch.inverse(n_src))
        out.append(torch.inverse(dst_norm_trans_src_norm))
    return torch.cat(out, dim=0).contiguous()


def _fptr(t: torch.Tensor):
    return C.cast(t.data_ptr(), C.POINTER(C.c_float))


def host_grid(T: torch.Tensor, dst_hw: Sequence[int]) -> torch.Tensor:
    """Normalised sampling grid [V,Hg,Wg,2] computed by the library's host code (no GPU needed)."""
    lib = _lib.load()
    T = T.contiguous().float()
    V, (Hg, Wg) = T.shape[0], (int(dst_hw[0]), int(dst_hw[1]))
    grid = torch.empty(V, Hg, Wg, 2, dtype=torch.float32)
    _lib.check(lib.mvwarp_host_grid(_fptr(T), V, Hg, Wg, _fptr(grid)), "mvwarp_host_grid")
    return grid


def host_csr(grid: torch.Tensor, src_hw: Sequence[int], with_map: bool = False, lowres_hw: Sequence[int] = None):
    """Statistics (and optionally the arrays ``tex_of, row_ptr, cells, weights``) of the inverse CSR map of
    ``grid`` [V,Hg,Wg,2]; host only.  Row r of view v is the texel ``tex_of[v, r]`` (tile-major order).
    With ``lowres_hw`` the map is the composition with the bilinear up-sample ``lowres_hw -> src_hw`` and its
    rows are the texels of the low-res image."""
    lib = _lib.load()
    grid = grid.contiguous().float()
    V, Hg, Wg, _ = grid.shape
    Hs, Ws = int(src_hw[0]), int(src_hw[1])
    hl, wl = (int(lowres_hw[0]), int(lowres_hw[1])) if lowres_hw is not None else (0, 0)
    st = _lib.Stats()
    if not with_map:
        _lib.check(lib.mvwarp_host_csr_upsampled(_fptr(grid), V, hl, wl, Hs, Ws, Hg, Wg, C.byref(st), None, None, None,
                                                 None, 0), "mvwarp_host_csr")
        return st.as_dict()
    nrow = hl * wl if hl else Hs * Ws
    cap = (16 if hl else 4) * V * Hg * Wg
    tex_of = torch.empty(V * nrow, dtype=torch.int32)
    row_ptr = torch.empty(V * (nrow + 1), dtype=torch.int32)
    cells = torch.empty(cap, dtype=torch.int32)
    weights = torch.empty(cap, dtype=torch.float32)
    u32 = C.POINTER(C.c_uint32)
    _lib.check(lib.mvwarp_host_csr_upsampled(_fptr(grid), V, hl, wl, Hs, Ws, Hg, Wg, C.byref(st),
                                             C.cast(tex_of.data_ptr(), u32), C.cast(row_ptr.data_ptr(), u32),
                                             C.cast(cells.data_ptr(), u32), _fptr(weights), cap), "mvwarp_host_csr")
    d = st.as_dict()
    return d, tex_of.view(V, nrow), row_ptr.view(V, nrow + 1), cells[:d["nnz"]], weights[:d["nnz"]]


def host_upsample_windows(T: torch.Tensor, lowres_hw: Sequence[int], src_hw: Sequence[int], dst_hw: Sequence[int]):
    """The composed 3 x 3 low-res windows of ``warp(F.interpolate(low, src_hw, bilinear))`` that the fused warp -> conv1 kernel uses
    on fused up-sample plans, computed on the host by the kernel's own source (``mvwarp_host_upsample_windows``; no GPU needed).
    ``T`` [V,3,3] from :func:`src_norm_trans_dst_norm`.  Returns ``origin`` int32 [V,Hg,Wg,2] (ox, oy), ``mask`` int64 [V,Hg,Wg]
    (bit j*3+i: slot live), ``wx``, ``wy`` fp32 [V,Hg,Wg,3]:  ``out[v,y,x] = sum_{j,i} wy[j] * wx[i] * low[v, oy+j, ox+i]``."""
    lib = _lib.load()
    T = T.contiguous().float()
    V, (Hg, Wg) = T.shape[0], (int(dst_hw[0]), int(dst_hw[1]))
    origin = torch.empty(V, Hg, Wg, 2, dtype=torch.int32)
    mask = torch.empty(V, Hg, Wg, dtype=torch.int32)
    wx = torch.empty(V, Hg, Wg, 3, dtype=torch.float32)
    wy = torch.empty(V, Hg, Wg, 3, dtype=torch.float32)
    _lib.check(lib.mvwarp_host_upsample_windows(_fptr(T), V, int(lowres_hw[0]), int(lowres_hw[1]), int(src_hw[0]), int(src_hw[1]), Hg, Wg,
                                                C.cast(origin.data_ptr(), C.POINTER(C.c_int32)),
                                                C.cast(mask.data_ptr(), C.POINTER(C.c_uint32)), _fptr(wx), _fptr(wy)),
               "mvwarp_host_upsample_windows")
    return origin, mask.long() & 0x1ff, wx, wy


class WarpPlan:
    """Opaque device-resident plan for one camera rig (wraps mvwarp_plan_t*)."""

    def __init__(self, handle, device: torch.device, key):
        self._handle = handle
        self.device = device
        self.key = key
        st = _lib.Stats()
        _lib.check(_lib.load().mvwarp_plan_stats(handle, C.byref(st)), "mvwarp_plan_stats")
        self.stats: Dict[str, int] = st.as_dict()
        self._finalizer = weakref.finalize(self, WarpPlan._destroy, handle)

    @staticmethod
    def _destroy(handle):
        try:
            _lib.load().mvwarp_plan_destroy(handle)
        except Exception:  # interpreter shutdown
            pass

    @property
    def handle(self):
        return self._handle

    @property
    def V(self): return self.stats["V"]

    @property
    def src_hw(self) -> Tuple[int, int]: return self.stats["Hs"], self.stats["Ws"]

    @property
    def grid_hw(self) -> Tuple[int, int]: return self.stats["Hg"], self.stats["Wg"]

    @property
    def upsampled(self) -> bool: return self.stats["lowres_h"] > 0

    @property
    def feat_hw(self) -> Tuple[int, int]:
        """Spatial size of the feature tensors the kernels read / write gradients for."""
        return (self.stats["lowres_h"], self.stats["lowres_w"]) if self.upsampled else self.src_hw

    def matrices(self) -> torch.Tensor:
        """fp32 [V,3,3] ``src_norm_trans_dst_norm`` matrices the plan was built from (not for explicit-grid plans)."""
        out = torch.empty(self.V, 3, 3, dtype=torch.float32)
        _lib.check(_lib.load().mvwarp_plan_matrices(self._handle, _fptr(out)), "mvwarp_plan_matrices")
        return out

    def save(self, path: str) -> None:
        """Write the device-resident plan to ``path`` (store it next to the calibration, so a
        start-up costs one file read + one H2D copy instead of the host builder: 0.1 s Wildtrack, 1.7 s 32-camera rig)."""
        _lib.check(_lib.load().mvwarp_plan_save(self._handle, os.fsencode(path)), "mvwarp_plan_save")

    def device_grid(self) -> torch.Tensor:
        """Normalised grid [V,Hg,Wg,2] as the DEVICE computes it (test hook)."""
        Hg, Wg = self.grid_hw
        out = torch.empty(self.V, Hg, Wg, 2, dtype=torch.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.load().mvwarp_plan_device_grid(self._handle, _fptr(out), C.c_void_p(stream)), "mvwarp_plan_device_grid")
        return out

    # ---- algorithmic (compulsory) HBM bytes per frame, SURVEY.md 8(d) -------------------------------
    def fwd_bytes(self, C_: int, elem: int, with_coord: bool = True) -> int:
        Hg, Wg = self.grid_hw
        return elem * C_ * self.stats["unique_texels"] + elem * (self.V * C_ + (2 if with_coord else 0)) * Hg * Wg

    def bwd_bytes(self, C_: int, elem: int) -> int:
        Hg, Wg = self.grid_hw
        Hs, Ws = self.feat_hw
        return elem * self.V * C_ * Hg * Wg + elem * self.V * C_ * Hs * Ws


_cache: Dict[tuple, WarpPlan] = {}
_cache_lock = threading.Lock()


def _device_index(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.MvwarpError(f"warp plans live on CUDA devices only (got {device}); there is no CPU path")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def load_plan(path: str, device, expect_matrices: torch.Tensor = None) -> WarpPlan:
    """Re-create a plan saved with :meth:`WarpPlan.save` on ``device``.  The library refuses truncated / corrupt files
    and files from an incompatible build.  ``expect_matrices`` (fp32 [V,3,3] ``src_norm_trans_dst_norm``) makes the
    load fail unless the stored rig is bit-for-bit that one."""
    device = _device_index(device)
    handle = C.c_void_p()
    _lib.check(_lib.load().mvwarp_plan_load(os.fsencode(path), device.index, C.byref(handle)), "mvwarp_plan_load")
    plan = WarpPlan(handle, device, ("file", os.path.abspath(path), device.index))
    if expect_matrices is not None:
        want = expect_matrices.detach().to("cpu").float().contiguous()
        if tuple(want.shape) != (plan.V, 3, 3) or not torch.equal(plan.matrices(), want):
            raise _lib.MvwarpError(f"{path} holds the plan of a different camera rig")
    return plan


def _disk_cached(key: tuple, T: torch.Tensor, shapes: tuple, device: torch.device, cache_dir, build):
    """Plan from ``cache_dir`` (or $MVWARP_PLAN_DIR) if a matching file is there, else built and stored there."""
    cache_dir = cache_dir or os.environ.get("MVWARP_PLAN_DIR")
    if not cache_dir:
        return build()
    name = hashlib.sha1(repr(key[:-1]).encode()).hexdigest()[:20] + ".mvwplan"
    path = os.path.join(cache_dir, name)
    if os.path.exists(path):
        try:
            plan = load_plan(path, device, expect_matrices=T)
            st = plan.stats
            if (st["Hs"], st["Ws"], st["Hg"], st["Wg"], st["lowres_h"], st["lowres_w"]) == shapes:
                plan.key = key
                return plan
        except _lib.MvwarpError:
            pass                      # stale or foreign file: rebuild and overwrite
    plan = build()
    os.makedirs(cache_dir, exist_ok=True)
    plan.save(path)
    return plan


def get_plan(proj_mats: torch.Tensor, src_hw: Sequence[int], grid_shape: Sequence[int], device, cache_dir: str = None) -> WarpPlan:
    """Plan for homographies ``proj_mats`` [V,3,3] (source pixel -> grid pixel), cached in the process and, with
    ``cache_dir`` (or $MVWARP_PLAN_DIR), on disk."""
    device = _device_index(device)
    src_hw = (int(src_hw[0]), int(src_hw[1]))
    dst_hw = (int(grid_shape[0]), int(grid_shape[1]))
    m32 = proj_mats.detach().to("cpu").float().contiguous()
    key = ("mat", m32.numpy().tobytes(), src_hw, dst_hw, device.index)
    with _cache_lock:
        plan = _cache.get(key)
    if plan is not None:
        return plan
    T = src_norm_trans_dst_norm(m32, src_hw, dst_hw)

    def build():
        handle = C.c_void_p()
        _lib.check(_lib.load().mvwarp_plan_create(_fptr(T), T.shape[0], src_hw[0], src_hw[1], dst_hw[0], dst_hw[1],
                                                  device.index, C.byref(handle)), "mvwarp_plan_create")
        return WarpPlan(handle, device, key)

    plan = _disk_cached(key, T, (*src_hw, *dst_hw, 0, 0), device, cache_dir, build)
    with _cache_lock:
        plan = _cache.setdefault(key, plan)
    return plan


def get_plan_upsampled(proj_mats: torch.Tensor, lowres_hw: Sequence[int], upsample_hw: Sequence[int],
                       grid_shape: Sequence[int], device, cache_dir: str = None) -> WarpPlan:
    """Plan for ``warp(F.interpolate(feats[.., hl, wl], upsample_hw, mode='bilinear'))`` as ONE sparse operator on
    the low-res feature map (persp_trans_detector.py:65 + :68-69).  ``proj_mats`` are the
    reference's matrices for the up-sampled (``upsample_hw``) feature map, unchanged."""
    device = _device_index(device)
    low = (int(lowres_hw[0]), int(lowres_hw[1]))
    src_hw = (int(upsample_hw[0]), int(upsample_hw[1]))
    dst_hw = (int(grid_shape[0]), int(grid_shape[1]))
    m32 = proj_mats.detach().to("cpu").float().contiguous()
    key = ("up", m32.numpy().tobytes(), low, src_hw, dst_hw, device.index)
    with _cache_lock:
        plan = _cache.get(key)
    if plan is not None:
        return plan
    T = src_norm_trans_dst_norm(m32, src_hw, dst_hw)

    def build():
        handle = C.c_void_p()
        _lib.check(_lib.load().mvwarp_plan_create_upsampled(_fptr(T), T.shape[0], low[0], low[1], src_hw[0], src_hw[1],
                                                            dst_hw[0], dst_hw[1], device.index, C.byref(handle)),
                   "mvwarp_plan_create_upsampled")
        return WarpPlan(handle, device, key)

    plan = _disk_cached(key, T, (*src_hw, *dst_hw, *low), device, cache_dir, build)
    with _cache_lock:
        plan = _cache.setdefault(key, plan)
    return plan


def get_plan_from_grid(grid: torch.Tensor, src_hw: Sequence[int], device) -> WarpPlan:
    """Plan for an explicit normalised grid [V,Hg,Wg,2] (grid_sample semantics, align_corners=True)."""
    device = _device_index(device)
    g = grid.detach().to("cpu").float().contiguous()
    if g.dim() != 4 or g.shape[-1] != 2:
        raise ValueError(f"grid must be [V,Hg,Wg,2]. Got {tuple(grid.shape)}")
    src_hw = (int(src_hw[0]), int(src_hw[1]))
    key = ("grid", g.numpy().tobytes(), src_hw, tuple(g.shape), device.index)
    with _cache_lock:
        plan = _cache.get(key)
    if plan is not None:
        return plan
    handle = C.c_void_p()
    _lib.check(_lib.load().mvwarp_plan_create_from_grid(_fptr(g), g.shape[0], src_hw[0], src_hw[1], g.shape[1], g.shape[2],
                                                        device.index, C.byref(handle)), "mvwarp_plan_create_from_grid")
    plan = WarpPlan(handle, device, key)
    with _cache_lock:
        plan = _cache.setdefault(key, plan)
    return plan


def clear_plan_cache():
    with _cache_lock:
        _cache.clear()
