"""Launches the non-GEMM hot kernels once each at representative sizes (for `ncu --set full`):
fused kernel-matrix backward (c2 Kuf), TMA-staged sampling, whitened KL fwd/bwd (c3 layer),
scale_cols / dA_init via one c2 hidden-layer fwd+bwd, Gaussian VE."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
rng = np.random.default_rng(0)

# a8: sampling, 64M elements (TMA-staged path)
n = 1 << 26
m = torch.randn(n, dtype=torch.float64, device=dev)
v = torch.rand(n, dtype=torch.float64, device=dev) + 0.1
e = torch.randn(n, dtype=torch.float64, device=dev)
for _ in range(1):
    ops.reparam_sample(m, v, e)
del m, v, e

# a9: whitened KL, c3 layer shape
M, P = 1024, 32
q_mu = torch.randn(M, P, dtype=torch.float64, device=dev)
q_sqrt = torch.tril(torch.randn(P, M, M, dtype=torch.float64, device=dev)) + 2 * torch.eye(M, dtype=torch.float64, device=dev)
g = torch.ones((), dtype=torch.float64, device=dev)
for _ in range(1):
    ops.kl_white(q_mu, q_sqrt)
    ops.kl_white_bwd(q_mu, q_sqrt, g)
del q_mu, q_sqrt

# a14: one c2 hidden layer fwd+bwd (M=256, D=8, P=8, N=8192): fused kmat backward, scale_cols, dA_init, GEMMs
M, D, P, N = 256, 8, 8, 8192
X = torch.randn(N, D, dtype=torch.float64, device=dev)
Z = torch.randn(1, M, D, dtype=torch.float64, device=dev).requires_grad_(True)
ls = (torch.rand(1, D, dtype=torch.float64, device=dev) + 1.0).requires_grad_(True)
var = torch.ones(1, dtype=torch.float64, device=dev).requires_grad_(True)
qm = (0.1 * torch.randn(M, P, dtype=torch.float64, device=dev)).requires_grad_(True)
qs = (torch.eye(M, dtype=torch.float64, device=dev).repeat(P, 1, 1) * 0.5
      + 0.01 * torch.tril(torch.randn(P, M, M, dtype=torch.float64, device=dev), -1)).requires_grad_(True)
eps = torch.randn(N, P, dtype=torch.float64, device=dev)
for _ in range(1):
    f, fm, fv, kl = ops.svgp_layer(X, Z, ls, var, qm, qs, mean_add=None, eps=eps, kernel="se", whiten=True, jitter=1e-6)
    (f.sum() + kl).backward()
torch.cuda.synchronize()
print("ok")
