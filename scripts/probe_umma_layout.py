#!/usr/bin/env python
"""Which non-canonical A-operand layouts does tcgen05.mma read correctly?  (mvwarp_debug_gemm_layout)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvdet_b200 import _lib
lib = _lib.load_dev()
g = torch.Generator().manual_seed(0)
M, K = 128, 128
A = torch.randn(M, K, generator=g).bfloat16().cuda()
B = torch.randn(256, K, generator=g).bfloat16().cuda()
ref = A.float() @ B.float().T
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for sbo in (1024, 1280, 2304):
    for start in (0, 1, 2, 10, 11):
        for bo in (0, 1):
            D = torch.full((M, 256), float("nan"), device="cuda")
            rc = lib.mvwarp_debug_gemm_layout(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), M, 256, K, sbo, start, bo, stream)
            if rc:
                print(sbo, start, bo, "rc", rc, lib.mvwarp_last_error().decode()); continue
            torch.cuda.synchronize()
            err = float((D - ref).abs().max() / ref.abs().max())
            print(f"sbo={sbo:5d} start_cells={start:2d} base_offset_field={bo}: max rel err {err:.3e} {'OK' if err < 1e-5 else 'WRONG'}")
