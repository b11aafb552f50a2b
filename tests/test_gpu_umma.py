"""tcgen05 / TMEM building blocks (north_star (c)): the validation GEMM against torch.matmul."""
import ctypes as C

import pytest
import torch

from mvdet_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K", [(128, 64), (256, 192), (1024, 512)])
def test_tcgen05_gemm_matches_matmul(M, K):
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g).bfloat16().cuda()
    B = torch.randn(256, K, generator=g).bfloat16().cuda()
    D = torch.full((M, 256), float("nan"), device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mvwarp_debug_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), M, 256, K, stream),
               "mvwarp_debug_gemm")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = float((D - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err          # bf16 products are exact in fp32; only the summation order differs
