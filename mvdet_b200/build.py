"""Build libmvwarp.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI).

Two libraries come out of the same sources:
  * ``libmvwarp.so``      the product: default kernels only, no environment knobs, no debug exports;
  * ``libmvwarp_dev.so``  ``-DMVWARP_DEV``: additionally the superseded kernel generations, the tuning knobs
                          (``MVWARP_*`` environment variables, read once) and the ``mvwarp_debug_*`` validation hooks that
                          tests/test_gpu_umma.py and the scripts under scripts/ use.  Nothing in the package loads it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmvwarp.so")
DEV_LIB_PATH = os.path.join(PKG_DIR, "libmvwarp_dev.so")
OBJ_DIR = os.path.join(PKG_DIR, "csrc", "_obj")
SOURCES = ["mvwarp_plan.cu", "mvwarp_kernels.cu", "mvwarp_conv1.cu", "mvwarp_conv2.cu", "mvwarp_conv1_bwd.cu", "mvwarp_host_io.cu", "mvwarp_layout.cu"]
HEADERS = ["mvwarp_coords.h", "mvwarp_internal.h", "mvwarp_umma.h", "mvwarp_uptaps.h", os.path.join("..", "..", "include", "mvwarp.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # -ffp-contract=off: the host plan builder must round exactly like the device intrinsics (mvwarp_coords.h)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmvwarp.so cannot be built (set NVCC=/path/to/nvcc)")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale(dev: bool = False) -> bool:
    lib = DEV_LIB_PATH if dev else LIB_PATH
    if not os.path.exists(lib):
        return True
    built = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in _sources() + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force: bool = False, verbose: bool = False, dev: bool = False) -> str:
    """Compile the CUDA sources into mvdet_b200/libmvwarp.so (``dev=True``: libmvwarp_dev.so); returns the path.
    Translation units are compiled in parallel, then linked."""
    lib = DEV_LIB_PATH if dev else LIB_PATH
    if not force and not is_stale(dev):
        return lib
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    defs = ["-DMVWARP_DEV"] if dev else []

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".dev.o" if dev else ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defs, "-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, cmd, res

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _sources()))
    for obj, cmd, res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stdout + res.stderr)
    cmd = [nvcc, "-shared", "-o", lib, *[obj for obj, _, _ in results]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return lib


if __name__ == "__main__":
    import sys
    # python -m mvdet_b200.build [--dev | --all] [-v]
    if "--all" in sys.argv:
        print(build_library(force=True, verbose="-v" in sys.argv))
        print(build_library(force=True, verbose="-v" in sys.argv, dev=True))
    else:
        print(build_library(force=True, verbose="-v" in sys.argv, dev="--dev" in sys.argv))
