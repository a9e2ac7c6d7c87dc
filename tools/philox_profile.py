"""Launches the in-kernel-noise sampling kernel (philox_normal_kernel) once at 2^26 elements, for
`ncu --set full -k regex:philox` (which pipe bounds it: FP64 Box-Muller vs HBM)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import ops  # noqa: E402

n = 1 << 26
m = torch.randn(n, dtype=torch.float64, device="cuda")
v = torch.rand(n, dtype=torch.float64, device="cuda") + 0.1
ops.reparam_sample_philox(m, v, seed=1, offset=0)
torch.cuda.synchronize()
