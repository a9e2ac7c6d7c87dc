"""Launches the hot DMMA GEMM shapes a few times (for ncu): c2 B_p = Lq^T A, c3-like, dense."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import _cabi  # noqa: E402

lib = _cabi.load()
dev = "cuda"


def run(P, M, N, tri, transA, reps=1):
    Lq = torch.tril(torch.randn(P, M, M, dtype=torch.float64, device=dev))
    A = torch.randn(M, N, dtype=torch.float64, device=dev)
    out = torch.empty(P, M, N, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(reps):
        _cabi.check(lib.gpfx_gemm(P, M, N, M, 1.0, Lq.data_ptr(), M, M * M, int(transA), A.data_ptr(), N, 0, 0, 0.0,
                                  out.data_ptr(), N, M * N, tri, 0, st))
    torch.cuda.synchronize()


run(8, 256, 8192, _cabi.TRI_UPPER_A, True)
run(8, 1024, 8192, _cabi.TRI_UPPER_A, True)
run(1, 4096, 4096, 0, False)
print("ok")


def run_nt(P, M, N, reps=1):
    """dLq_p = tril(A Bs_p^T): [M x N] x [M x N]^T, k-contiguous on both sides, lower-triangular output."""
    A = torch.randn(M, N, dtype=torch.float64, device=dev)
    Bs = torch.randn(P, M, N, dtype=torch.float64, device=dev)
    out = torch.empty(P, M, M, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(reps):
        _cabi.check(lib.gpfx_gemm(P, M, M, N, 1.0, A.data_ptr(), N, 0, 0, Bs.data_ptr(), N, M * N, 1, 0.0,
                                  out.data_ptr(), M, M * M, 0, 1, st))
    torch.cuda.synchronize()


if "--nt" in sys.argv:
    run_nt(8, 1024, 8192)
    print("nt ok")
