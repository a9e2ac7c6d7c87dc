"""Measures the FP64 ceiling of the box (cuBLAS DGEMM through torch.matmul, same protocol as the
bf16 entry of MEASURED_PEAKS.json: best of 10 = burst, back-to-back for 4 s = sustained) and the
throughput of gpflux_b200's own DMMA GEMM on the shapes the SVGP path uses.  Writes
gpurun_out/fp64_peak.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def time_fn(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    from gpflux_b200 import ops
    from gpflux_b200._cabi import TRI_LOWER_A, TRI_UPPER_A

    out = {"gpu": torch.cuda.get_device_name(0)}
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    B = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.empty_like(A)
    ms = time_fn(lambda: torch.matmul(A, B, out=C))
    out["cublas_dgemm_8192_burst_tflops"] = 2 * n**3 / ms / 1e9
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    k = 0
    while time.perf_counter() - t0 < 4.0:
        for _ in range(4):
            torch.matmul(A, B, out=C)
        k += 4
        torch.cuda.synchronize()
    b.record()
    b.synchronize()
    out["cublas_dgemm_8192_sustained_tflops"] = 2 * n**3 * k / a.elapsed_time(b) / 1e9

    res = []
    for (bt, M, N, K, tri, name) in [
        (1, 4096, 4096, 4096, 0, "dense 4096^3"),
        (1, 8192, 8192, 8192, 0, "dense 8192^3"),
        (1, 256, 8192, 256, TRI_LOWER_A, "c2 A=Linv*Kuf"),
        (8, 256, 8192, 256, TRI_UPPER_A, "c2 B_p=Lq^T*A (transA)"),
        (32, 1024, 8192, 1024, TRI_LOWER_A, "c3 A=Linv*Kuf"),
        (32, 1024, 8192, 1024, TRI_UPPER_A, "c3 B_p=Lq^T*A (transA)"),
    ]:
        tA = bool(tri & TRI_UPPER_A)
        Am = torch.randn(bt, K, M, dtype=torch.float64, device="cuda") if tA else torch.randn(bt, M, K, dtype=torch.float64, device="cuda")
        Am = torch.tril(Am) if tri else Am
        Bm = torch.randn(bt, K, N, dtype=torch.float64, device="cuda")
        ms = time_fn(lambda: ops.gemm(Am, Bm, transA=tA, tri=tri), iters=5, warm=2)
        dense = 2.0 * bt * M * N * K
        algo = dense / 2 if tri else dense
        # ops.gemm allocates + zero-fills C each call; time that separately and subtract
        ms0 = time_fn(lambda: torch.zeros((bt, M, N), dtype=torch.float64, device="cuda"), iters=5, warm=2)
        res.append({"case": name, "ms": ms, "ms_alloc": ms0, "algorithmic_tflops": algo / max(ms - ms0, 1e-6) / 1e9,
                    "dense_equiv_tflops": dense / max(ms - ms0, 1e-6) / 1e9})
        del Am, Bm
    out["gpfx_gemm"] = res

    # HBM: reparam sample on 1 Gi doubles? keep it modest: 256 Mi elements x 4 arrays x 8 B = 8.6 GB
    nel = 1 << 28
    m = torch.randn(nel, dtype=torch.float64, device="cuda")
    v = torch.rand(nel, dtype=torch.float64, device="cuda") + 0.1
    e = torch.randn(nel, dtype=torch.float64, device="cuda")
    ms = time_fn(lambda: ops.reparam_sample(m, v, e), iters=5, warm=2)
    out["reparam_sample_gbs_incl_alloc"] = 4 * 8 * nel / ms / 1e6
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/fp64_peak.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
