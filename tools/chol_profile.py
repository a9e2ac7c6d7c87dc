import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gpflux_b200 import ops
torch.manual_seed(0)
M = 256
A = torch.randn(1, M, M, dtype=torch.float64, device="cuda")
A = A @ A.transpose(1, 2) + M * torch.eye(M, dtype=torch.float64, device="cuda")
for _ in range(3):
    L, Li, info = ops.cholesky_inverse(A)
torch.cuda.synchronize()
print(float((L[0] @ L[0].T - A[0]).abs().max()))
