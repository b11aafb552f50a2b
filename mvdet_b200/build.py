"""Build libmvwarp.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmvwarp.so")
SOURCES = ["mvwarp_plan.cu", "mvwarp_kernels.cu", "mvwarp_conv1.cu", "mvwarp_conv2.cu"]
HEADERS = ["mvwarp_coords.h", "mvwarp_internal.h", "mvwarp_umma.h", os.path.join("..", "..", "include", "mvwarp.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # -ffp-contract=off: the host plan builder must round exactly like the device intrinsics (mvwarp_coords.h)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmvwarp.so cannot be built (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources into mvdet_b200/libmvwarp.so; returns the path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
