"""Per-k scaled NT product (gpfx_gemm_kscaled) on the c3 dLq shape: device time against the unscaled product
and against scale_cols + unscaled product.  gpurun -- python tools/gemm_kscale_check.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import _cabi  # noqa: E402

lib = _cabi.load()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def time_launch(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(0.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


batch, M, K = 32, 1024, 8192
A = torch.randn(batch, M, K, dtype=torch.float64, device=dev)
B = torch.randn(batch, M, K, dtype=torch.float64, device=dev)
s = torch.randn(batch, K, dtype=torch.float64, device=dev)
C = torch.empty(batch, M, M, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
res = {}
res["unscaled_ms"] = time_launch(lambda: _cabi.check(lib.gpfx_gemm_splitk(
    batch, M, M, K, 1.0, A.data_ptr(), K, M * K, 0, B.data_ptr(), K, M * K, 1, 0.0, C.data_ptr(), M, M * M, 0, 1, 1, None, st)))
res["kscaled_ms"] = time_launch(lambda: _cabi.check(lib.gpfx_gemm_kscaled(
    batch, M, M, K, 1.0, A.data_ptr(), K, M * K, B.data_ptr(), K, M * K, s.data_ptr(), K, C.data_ptr(), M, M * M, 1, 1, None, st)))
res["scale_pass_ms"] = time_launch(lambda: B.mul_(s[:, None, :]))
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_kscale_check.json", "w"))
