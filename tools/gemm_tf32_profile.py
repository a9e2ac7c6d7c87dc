"""Launches the 3xTF32 tcgen05 GEMM once per shape (for `ncu --set full`): c3 B_p = Lq^T A and dense 8192^3."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import _cabi  # noqa: E402

lib = _cabi.load()
st = torch.cuda.current_stream().cuda_stream


def gemm(batch, M, N, K, tA, tB, tri=0, tril_out=0):
    A = torch.randn(batch, K, M, dtype=torch.float64, device="cuda") if tA else torch.randn(batch, M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(batch, N, K, dtype=torch.float64, device="cuda") if tB else torch.randn(batch, K, N, dtype=torch.float64, device="cuda")
    C = torch.empty(batch, M, N, dtype=torch.float64, device="cuda")
    ws = torch.empty(lib.gpfx_gemm_tf32x3_workspace_bytes(batch, M, N, K) // 4 + 16, dtype=torch.float32, device="cuda")
    _cabi.check(lib.gpfx_gemm_tf32x3(batch, M, N, K, 1.0, A.data_ptr(), A.shape[2], A.shape[1] * A.shape[2], int(tA), B.data_ptr(),
                                     B.shape[2], B.shape[1] * B.shape[2], int(tB), 0.0, C.data_ptr(), N, M * N, tri, tril_out,
                                     ws.data_ptr(), st))
    torch.cuda.synchronize()


gemm(32, 1024, 8192, 1024, True, False, tri=2)
gemm(1, 8192, 8192, 1024, False, True)
print("ok")
