"""Times the DMMA GEMM tile configurations on the shapes of the SVGP path (CUDA events, L2 flushed
between launches).  Third sweep: split-K factor x tile configuration on the M x M = (M x N)(N x M) products.  Writes gpurun_out/gemm_sweep3.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import _cabi  # noqa: E402

lib = _cabi.load()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def time_launch(fn, reps=8):
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


def case(name, batch, M, N, K, tri, tA, tB, sB_zero=False, tril_out=0, sA_zero=False, splits=1, cfgs=(0, 3, 4, -1)):
    na = 1 if sA_zero else batch
    A = torch.randn(na, K, M, dtype=torch.float64, device=dev) if tA else torch.randn(na, M, K, dtype=torch.float64, device=dev)
    if tri & (_cabi.TRI_LOWER_A | _cabi.TRI_UPPER_A):
        A = torch.tril(A)
    nb = 1 if sB_zero else batch
    B = torch.randn(nb, N, K, dtype=torch.float64, device=dev) if tB else torch.randn(nb, K, N, dtype=torch.float64, device=dev)
    C = torch.empty(batch, M, N, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lda = A.shape[2]
    ldb = B.shape[2]

    ws = torch.empty(max(1, splits * batch * M * N), dtype=torch.float64, device=dev)

    def fn():
        _cabi.check(lib.gpfx_gemm_splitk(batch, M, N, K, 1.0, A.data_ptr(), lda, 0 if sA_zero else A.shape[1] * A.shape[2], int(tA),
                                         B.data_ptr(), ldb, 0 if sB_zero else B.shape[1] * B.shape[2], int(tB), 0.0, C.data_ptr(),
                                         N, M * N, tri, tril_out, splits, ws.data_ptr(), st))

    dense = 2.0 * batch * M * N * K
    algo = dense / 2 if (tri or tril_out) else dense
    out = {"case": name}
    for cfg in cfgs:
        lib.gpfx_tune_gemm_config(cfg)
        ms = time_launch(fn)
        out[f"cfg{cfg}_ms"] = ms
        out[f"cfg{cfg}_algo_tflops"] = algo / ms / 1e9
    lib.gpfx_tune_gemm_config(-1)
    return out


res = []
for sp in (8, 12, 16, 20, 24, 29, 32):
    res.append(case(f"c2 dL = tril(dKuf A^T) P=1 splits={sp}", 1, 256, 256, 8192, 0, False, True, tril_out=1, splits=sp, cfgs=(1, 5, 6, -1)))
for sp in (13, 16, 20, 26):
    res.append(case(f"c2 dLq = tril(A Bs^T) P=8 splits={sp}", 8, 256, 256, 8192, 0, False, True, tril_out=1, sA_zero=True, splits=sp, cfgs=(1, 5, -1)))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_sweep3.json", "w"), indent=1)
for r in res:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if not k.endswith("tflops")})
