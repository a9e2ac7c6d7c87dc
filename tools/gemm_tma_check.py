"""TMA-fed DMMA GEMM (gemm_f64_tma.cu) against the cp.async kernel and torch.matmul: exact agreement on
layouts / triangular flags / batch strides / tails, then device timings (CUDA events, L2 flushed) on the
c2 / c3 products and a dense 8192^3.  Writes gpurun_out/gemm_tma_check.json.
    gpurun -- python tools/gemm_tma_check.py [--time-only] [--modes 1,2]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import _cabi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--time-only", action="store_true")
ap.add_argument("--check-only", action="store_true")
ap.add_argument("--modes", default="1,2,-1", help="0 = cp.async kernel, 1/2 = TMA variants, -1 = automatic")
args = ap.parse_args()
MODES = [int(x) for x in args.modes.split(",")]

lib = _cabi.load()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
TRI_LOWER_A, TRI_UPPER_A, TRI_LOWER_B, TRI_UPPER_B = 1, 2, 4, 8


def make(batch, M, N, K, tri, tA, tB, sA_zero=False, sB_zero=False, pad=0):
    g = torch.Generator(device=dev).manual_seed(1234 + M + 3 * N + 7 * K)
    na, nb = (1 if sA_zero else batch), (1 if sB_zero else batch)
    opA = torch.randn(na, M, K, dtype=torch.float64, device=dev, generator=g)
    opB = torch.randn(nb, K, N, dtype=torch.float64, device=dev, generator=g)
    if tri & TRI_LOWER_A:
        opA = torch.tril(opA)
    if tri & TRI_UPPER_A:
        opA = torch.triu(opA)
    if tri & TRI_LOWER_B:
        opB = torch.tril(opB)
    if tri & TRI_UPPER_B:
        opB = torch.triu(opB)
    A = opA.transpose(1, 2).contiguous() if tA else opA.contiguous()
    B = opB.transpose(1, 2).contiguous() if tB else opB.contiguous()
    if pad:  # leading dimension larger than the row length
        A = torch.nn.functional.pad(A, (0, pad)).contiguous()
        B = torch.nn.functional.pad(B, (0, pad)).contiguous()
    return opA, opB, A, B


def run(mode, batch, M, N, K, tri, tA, tB, A, B, C, sA_zero, sB_zero, tril_out, splits, ws, alpha=1.0, beta=0.0):
    lib.gpfx_tune_gemm_config(100 + mode if mode >= 0 else -1)
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.gpfx_gemm_splitk(batch, M, N, K, alpha, A.data_ptr(), A.shape[2], 0 if sA_zero else A.shape[1] * A.shape[2],
                                     int(tA), B.data_ptr(), B.shape[2], 0 if sB_zero else B.shape[1] * B.shape[2], int(tB), beta,
                                     C.data_ptr(), N, M * N, tri, tril_out, splits, ws.data_ptr(), st))
    lib.gpfx_tune_gemm_config(-1)


def check(name, batch, M, N, K, tri=0, tA=False, tB=False, sA_zero=False, sB_zero=False, tril_out=0, splits=1, pad=0,
          beta=0.0):
    opA, opB, A, B = make(batch, M, N, K, tri, tA, tB, sA_zero, sB_zero, pad)
    ref = torch.matmul(opA, opB).expand(batch, M, N)
    if tril_out:
        ref = torch.tril(ref)
    C0 = torch.randn(batch, M, N, dtype=torch.float64, device=dev)
    ws = torch.empty(max(1, splits * batch * M * N), dtype=torch.float64, device=dev)
    ref = 0.5 * ref + beta * C0
    outs = {}
    for mode in [0] + MODES:
        C = C0.clone()
        run(mode, batch, M, N, K, tri, tA, tB, A, B, C, sA_zero, sB_zero, tril_out, splits, ws, alpha=0.5, beta=beta)
        torch.cuda.synchronize()
        outs[mode] = C
    scale = float(ref.abs().max())
    line = {"case": name}
    ok = True
    for mode, C in outs.items():
        err = float((C - ref).abs().max()) / scale
        dif = float((C - outs[0]).abs().max()) / scale
        line[f"mode{mode}_err_vs_torch"] = err
        line[f"mode{mode}_diff_vs_cpasync"] = dif
        ok = ok and err < 1e-12 and bool(torch.isfinite(C).all())
    line["ok"] = ok
    print(("ok   " if ok else "FAIL ") + json.dumps(line))
    return line


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


def timing(name, batch, M, N, K, tri=0, tA=False, tB=False, sA_zero=False, sB_zero=False, tril_out=0, splits=1):
    opA, opB, A, B = make(batch, M, N, K, tri, tA, tB, sA_zero, sB_zero)
    del opA, opB
    C = torch.empty(batch, M, N, dtype=torch.float64, device=dev)
    ws = torch.empty(max(1, splits * batch * M * N), dtype=torch.float64, device=dev)
    dense = 2.0 * batch * M * N * K
    algo = dense / 2 if (tri or tril_out) else dense
    line = {"case": name, "algorithmic_gflop": algo / 1e9}
    for mode in [0] + MODES:
        ms = time_launch(lambda: run(mode, batch, M, N, K, tri, tA, tB, A, B, C, sA_zero, sB_zero, tril_out, splits, ws))
        line[f"mode{mode}_ms"] = round(ms, 4)
        line[f"mode{mode}_tflops"] = round(algo / ms / 1e9, 2)
    print(json.dumps(line))
    return line


res = {"checks": [], "timings": []}
if not args.time_only:
    C = res["checks"]
    for tA in (False, True):
        for tB in (False, True):
            C.append(check(f"dense tA={tA} tB={tB}", 2, 256, 384, 200, 0, tA, tB))
            C.append(check(f"dense tails tA={tA} tB={tB}", 3, 128 + 16 * 3, 256 + 16 * 5, 136 + 3, 0, tA, tB, pad=2))
    C.append(check("k-major tails (M, N not multiples of 16)", 2, 137, 201, 78, 0, False, True))
    C.append(check("lower-A NN", 3, 512, 640, 512, TRI_LOWER_A, False, False))
    C.append(check("upper-A TN", 3, 512, 640, 512, TRI_UPPER_A, True, False))
    C.append(check("upper-A TN broadcast B", 4, 256, 1024, 256, TRI_UPPER_A, True, False, sB_zero=True))
    C.append(check("lower-A x lower-B NN", 2, 512, 512, 512, TRI_LOWER_A | TRI_LOWER_B, False, False))
    C.append(check("lower-A NT upper-B", 2, 384, 384, 384, TRI_LOWER_A | TRI_UPPER_B, False, True))
    C.append(check("tril-out NT", 3, 512, 512, 1000, 0, False, True, tril_out=1))
    C.append(check("tril-out NT broadcast A, beta", 3, 256, 256, 2048, 0, False, True, sA_zero=True, tril_out=1, beta=1.0))
    C.append(check("tril-out NT split-K", 2, 256, 256, 4096, 0, False, True, tril_out=1, splits=5))
    C.append(check("upper-A TN split-K", 1, 256, 512, 256, TRI_UPPER_A, True, False, splits=3))
    C.append(check("beta NN lower-A", 2, 256, 512, 256, TRI_LOWER_A, False, False, beta=1.0))
    res["all_ok"] = all(c["ok"] for c in C)
    print("ALL OK" if res["all_ok"] else "SOME CHECKS FAILED")

if not args.check_only:
    T = res["timings"]
    T.append(timing("c3 A = Linv Kuf (NN lower-A) 32x1024x8192x1024", 32, 1024, 8192, 1024, TRI_LOWER_A, False, False))
    T.append(timing("c3 B = Lq^T A (TN upper-A) 32x1024x8192x1024", 32, 1024, 8192, 1024, TRI_UPPER_A, True, False))
    T.append(timing("c3 dLq = tril(A Bs^T) (NT) 32x1024x1024x8192", 32, 1024, 1024, 8192, 0, False, True, tril_out=1))
    T.append(timing("c2 B = Lq^T A (TN upper-A, shared A) 8x256x8192x256", 8, 256, 8192, 256, TRI_UPPER_A, True, False, sB_zero=True))
    T.append(timing("c2 A = Linv Kuf (NN lower-A) 1x256x8192x256", 1, 256, 8192, 256, TRI_LOWER_A, False, False))
    T.append(timing("dense NN batch 32 x 1024x8192x1024", 32, 1024, 8192, 1024))
    T.append(timing("dense NN batch 32 x 1024x8192x512", 32, 1024, 8192, 512))
    T.append(timing("dense NN batch 32 x 1024x8192x128", 32, 1024, 8192, 128))
    T.append(timing("dense NN 8192^3", 1, 8192, 8192, 8192))
    T.append(timing("dense NT 8192^3", 1, 8192, 8192, 8192, 0, False, True))
    T.append(timing("dense TN 8192^3", 1, 8192, 8192, 8192, 0, True, False))
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    ms = time_launch(lambda: torch.matmul(a, b))
    res["cublas_dgemm_8192_tflops"] = round(2 * 8192.0 ** 3 / ms / 1e9, 2)
    print("cuBLAS DGEMM 8192^3:", res["cublas_dgemm_8192_tflops"], "TF")

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_tma_check.json", "w"), indent=1)
