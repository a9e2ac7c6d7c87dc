import os, sys
sys.path.insert(0, os.getcwd())
import torch
from gpflux_b200 import ops
torch.manual_seed(0)
M, N, D = 256, 8192, 8
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3
for ls_v, var_v, xs in [(2.0, 1.0, 1.0), (2.0, 1e-6, 1.0), (1.0, 1.0, 1.0), (2.0, 1.0, 0.1), (0.5, 1.0, 1.0), (2.0, 1.0, 3.0)]:
    Z = xs * torch.randn(1, M, D, dtype=torch.float64, device="cuda")
    X = xs * torch.randn(N, D, dtype=torch.float64, device="cuda")
    ls = torch.full((1, D), ls_v, dtype=torch.float64, device="cuda")
    var = torch.full((1,), var_v, dtype=torch.float64, device="cuda")
    us = t(lambda: ops.kernel_matrix("se", Z, X, ls, var))
    K = ops.kernel_matrix("se", Z, X, ls, var)
    print(f"ls={ls_v} var={var_v} xscale={xs}: {us:.1f} us/launch (incl. alloc), K range [{float(K.min()):.3e}, {float(K.max()):.3e}]")
