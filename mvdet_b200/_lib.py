"""ctypes binding of libmvwarp.so (include/mvwarp.h).  There is no fallback: if the CUDA library
is missing or fails to load, every op raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

from .build import LIB_PATH as _DEFAULT_LIB_PATH

# development: MVWARP_LIB=/path/to/variant.so selects another build of the same ABI (kernel tuning runs)
LIB_PATH = os.environ.get("MVWARP_LIB", _DEFAULT_LIB_PATH)

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
F32, BF16 = 0, 1
MAX_LIST_VIEWS = 64      # MVWARP_MAX_LIST_VIEWS (include/mvwarp.h)


class Stats(C.Structure):
    _fields_ = [("V", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32), ("Hg", C.c_int32), ("Wg", C.c_int32),
                ("inline_coords", C.c_int32), ("unique_texels", C.c_uint64), ("nnz", C.c_uint64),
                ("valid_cells", C.c_uint64), ("max_row", C.c_uint32), ("bwd_tasks", C.c_uint32), ("heavy_rows", C.c_uint32),
                ("fwd_tasks", C.c_uint32), ("lowres_h", C.c_int32), ("lowres_w", C.c_int32), ("zero_segments", C.c_uint32),
                ("zero_rows", C.c_uint32), ("device_bytes", C.c_uint64)]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_F32P = C.POINTER(C.c_float)
_U32P = C.POINTER(C.c_uint32)

# name -> (restype, argtypes); must list every symbol include/mvwarp.h declares (tests check this)
SIGNATURES = {
    "mvwarp_version": (C.c_int, []),
    "mvwarp_last_error": (C.c_char_p, []),
    "mvwarp_host_grid": (C.c_int, [_F32P, C.c_int, C.c_int, C.c_int, _F32P]),
    "mvwarp_host_csr": (C.c_int, [_F32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Stats),
                                  _U32P, _U32P, _U32P, _F32P, C.c_uint64]),
    "mvwarp_host_csr_upsampled": (C.c_int, [_F32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(Stats), _U32P, _U32P, _U32P, _F32P, C.c_uint64]),
    "mvwarp_host_upsample_windows": (C.c_int, [_F32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(C.c_int32), _U32P, _F32P, _F32P]),
    "mvwarp_plan_create_upsampled": (C.c_int, [_F32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.POINTER(_P)]),
    "mvwarp_plan_create": (C.c_int, [_F32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "mvwarp_plan_create_from_grid": (C.c_int, [_F32P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(_P)]),
    "mvwarp_plan_destroy": (C.c_int, [_P]),
    "mvwarp_plan_save": (C.c_int, [_P, C.c_char_p]),
    "mvwarp_plan_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "mvwarp_plan_matrices": (C.c_int, [_P, _F32P]),
    "mvwarp_plan_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "mvwarp_plan_device_grid": (C.c_int, [_P, _F32P, _P]),
    "mvwarp_warp_views_fwd": (C.c_int, [_P, C.c_int, C.c_int, _P, _I64P, _P, _I64P, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_warp_conv1_fwd": (C.c_int, [_P, C.c_int, C.c_int, _P, _I64P, _P, _P, _P, _I64P, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_warp_conv1_fwd_ex": (C.c_int, [_P, C.c_int, C.c_int, _P, _I64P, _P, _P, _P, C.c_int, _I64P, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "mvwarp_warp_conv1_fwd_p2p": (C.c_int, [_P, C.c_int, C.c_int, _P, _I64P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.c_int, _P]),
    "mvwarp_inbox_reduce": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int64, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "mvwarp_warp_conv1_scratch_bytes": (C.c_int64, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mvwarp_conv3x3_fwd": (C.c_int, [_P, _I64P, _P, _P, _P, _I64P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, _P]),
    "mvwarp_conv3x3_head_fwd": (C.c_int, [_P, _I64P, _P, _P, _I64P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_conv3x3_head_scratch_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "mvwarp_conv3x3_head_fwd_ex": (C.c_int, [_P, _I64P, _P, _P, _I64P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int64,
                                             C.c_int, _P]),
    "mvwarp_warp_views_fwd_list": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), _I64P, _P, _I64P, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_warp_views_bwd_list": (C.c_int, [_P, C.c_int, C.c_int, _P, _I64P, C.c_int, C.POINTER(_P), _I64P,
                                             C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_relu_bwd_scratch_floats": (C.c_int64, [C.c_int, C.c_int]),
    "mvwarp_relu_bwd_cast": (C.c_int, [_P, _I64P, _P, _I64P, _P, _I64P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_layout_transpose": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_copy_touched_texels": (C.c_int, [_P, C.c_int, C.c_int, _P, _I64P, _P, _I64P, C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_plan_touched": (C.c_int, [_P, _U32P, _U32P, C.c_uint64]),
    "mvwarp_warp_views_bwd": (C.c_int, [_P, C.c_int, C.c_int, _P, _I64P, C.c_int, _P, _I64P,
                                        C.c_int, C.c_int, C.c_int, _P]),
}

# exported by the development build only (mvdet_b200/libmvwarp_dev.so, -DMVWARP_DEV; include/mvwarp.h #ifdef MVWARP_DEV)
DEV_SIGNATURES = {
    "mvwarp_debug_gemm": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "mvwarp_debug_gemm_layout": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
}

_lib = None
_dev_lib = None
_lock = threading.Lock()


class MvwarpError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises MvwarpError when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MvwarpError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m mvdet_b200.build` "
                "(or __graft_entry__.build()). There is no CPU or PyTorch fallback for warp_views.")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise MvwarpError(f"failed to load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.mvwarp_version() != 1:
            raise MvwarpError(f"libmvwarp.so ABI version {lib.mvwarp_version()} != 1: rebuild")
        _lib = lib
    return _lib


def load_dev():
    """The development build (superseded kernel generations, MVWARP_* tuning knobs, mvwarp_debug_* hooks) as a SEPARATE ctypes
    handle.  Used by tests and scripts only; nothing in the package calls it."""
    global _dev_lib
    if _dev_lib is not None:
        return _dev_lib
    from .build import DEV_LIB_PATH, build_library
    with _lock:
        if _dev_lib is None:
            path = os.environ.get("MVWARP_DEV_LIB") or (DEV_LIB_PATH if os.path.exists(DEV_LIB_PATH) else build_library(dev=True))
            lib = C.CDLL(path)
            for name, (res, args) in {**SIGNATURES, **DEV_SIGNATURES}.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            _dev_lib = lib
    return _dev_lib


def check(rc: int, what: str):
    if rc != OK:
        msg = load().mvwarp_last_error()
        raise MvwarpError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def i64x(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])
