"""Device time of every GEMM launch of one c3 hidden layer's forward + backward (M=1024, P=K=32, D=32,
8192 rows), in launch order, through the library's per-launch profile (CUDA events around each launch).
    gpurun -- python tools/layer_gemm_times.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpflux_b200 import _cabi  # noqa: E402
from gpflux_b200 import gpflow_compat as gfc  # noqa: E402
from gpflux_b200 import ops as gops  # noqa: E402
from gpflux_b200 import workloads as wl  # noqa: E402

lib = _cabi.load()
dev = torch.device("cuda", 0)
model = wl.build_c3(layers=2).to(dev)
hl = model.f_layers[0]
D = int(hl._inducing_points()[0].Z.shape[1])
kind, Z, ls, var = hl._packed(D)
base = [t.detach() for t in (Z, ls, var, hl.q_mu.value(), hl.q_sqrt.value())]
N, P = 8192, base[3].shape[1]
X = torch.randn(N, D, dtype=torch.float64, device=dev)
eps = torch.randn(N, P, dtype=torch.float64, device=dev)
w = torch.randn(N, P, dtype=torch.float64, device=dev)


def step():
    args = [t.clone().requires_grad_(True) for t in base]
    Xa = X.clone().requires_grad_(True)
    f, m_, v_, kl_ = gops.svgp_layer(Xa, *args, mean_add=Xa if P == D else None, eps=eps, kernel=kind,
                                     whiten=bool(hl.whiten), jitter=gfc.default_jitter())
    ((f * w).sum() + kl_).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
lib.gpfx_gemm_profile_read(None, 0)
lib.gpfx_gemm_profile_enable(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
step()
b.record()
torch.cuda.synchronize()
lib.gpfx_gemm_profile_enable(0)
buf = (_cabi.GemmSample * 4096)()
n = lib.gpfx_gemm_profile_read(buf, 4096)
print(f"layer fwd+bwd {a.elapsed_time(b):.2f} ms (profiled: launches serialised by their events)")
for i in range(n):
    s = buf[i]
    if s.ms < 1.0:
        continue
    op = ("T" if s.transA else "N") + ("T" if s.transB else "N")
    print(f"  {op} b={s.batch} M={s.M} N={s.N} K={s.K} tri={s.tri} tril_out={s.tril_out} cfg={s.cfg}: {s.ms:.3f} ms")
