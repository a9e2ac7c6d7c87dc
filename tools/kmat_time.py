"""Device time of kmat_fwd (Kuf, M=256, N=8192, D=8) for different kernel variances / data scales,
measured inside a CUDA graph of 50 launches (no host launch overhead)."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

from gpflux_b200 import _cabi

lib = _cabi.load()
torch.manual_seed(0)
M, N, D = 256, 8192, 8
out = torch.empty(1, M, N, dtype=torch.float64, device="cuda")


def run(ls_v, var_v, xs, zfromx):
    X = xs * torch.randn(N, D, dtype=torch.float64, device="cuda")
    Z = (X[:M].clone() if zfromx else xs * torch.randn(M, D, dtype=torch.float64, device="cuda"))[None].contiguous()
    ls = torch.full((1, D), ls_v, dtype=torch.float64, device="cuda")
    var = torch.full((1,), var_v, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        def launch():
            _cabi.check(lib.gpfx_kernel_matrix(0, 1, M, N, D, Z.data_ptr(), 0, X.data_ptr(), 1, ls.data_ptr(), var.data_ptr(),
                                               0, 0.0, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        launch()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50):
                launch()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
    print(f"ls={ls_v} var={var_v} xscale={xs} Z_from_X={zfromx}: {a.elapsed_time(b) / 50 * 1e3:.2f} us per launch")


for cfg in [(2.0, 1.0, 1.0, True), (2.0, 1e-6, 1.0, True), (2.0, 1.0, 1.0, False), (2.0, 1e-6, 1.0, False), (1.0, 1.0, 1.0, False), (2.0, 1e-3, 1.0, True)]:
    run(*cfg)
