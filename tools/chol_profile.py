"""Launches the batched Cholesky + inverse on K matrices of size M (for ncu / timing)."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from gpflux_b200 import ops

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
A = torch.randn(K, M, M, dtype=torch.float64, device="cuda")
A = A @ A.transpose(1, 2) + M * torch.eye(M, dtype=torch.float64, device="cuda")
for _ in range(3):
    L, Li, info = ops.cholesky_inverse(A)
torch.cuda.synchronize()
print(float((L[0] @ L[0].T - A[0]).abs().max()), float((Li[0] @ L[0] - torch.eye(M, dtype=torch.float64, device="cuda")).abs().max()))
