"""3xTF32 tcgen05 GEMM (csrc/gemm_tf32.cu, the float32 opt-in's contraction) against torch.matmul in float64:
relative error per layout / shape, then device timings on the c3 products beside the FP64 DMMA kernel.
Writes gpurun_out/gemm_tf32_check.json.   gpurun -- timeout 120 python tools/gemm_tf32_check.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import _cabi  # noqa: E402

lib = _cabi.load()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def make(batch, M, N, K, tA, tB, tri=0):
    g = torch.Generator(device=dev).manual_seed(7 + M + 3 * N + 5 * K)
    opA = torch.randn(batch, M, K, dtype=torch.float64, device=dev, generator=g)
    opB = torch.randn(batch, K, N, dtype=torch.float64, device=dev, generator=g)
    if tri & 1:
        opA = torch.tril(opA)
    if tri & 2:
        opA = torch.triu(opA)
    A = opA.transpose(1, 2).contiguous() if tA else opA.contiguous()
    B = opB.transpose(1, 2).contiguous() if tB else opB.contiguous()
    return opA, opB, A, B


def run_tf32(batch, M, N, K, tA, tB, A, B, C, ws, alpha=1.0, beta=0.0, tri=0, tril_out=0):
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.gpfx_gemm_tf32x3(batch, M, N, K, alpha, A.data_ptr(), A.shape[2], A.shape[1] * A.shape[2], int(tA), B.data_ptr(),
                                     B.shape[2], B.shape[1] * B.shape[2], int(tB), beta, C.data_ptr(), N, M * N, tri, tril_out,
                                     ws.data_ptr(), st))


def run_f64(batch, M, N, K, tA, tB, A, B, C, alpha=1.0, beta=0.0, tri=0, tril_out=0):
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.gpfx_gemm(batch, M, N, K, alpha, A.data_ptr(), A.shape[2], A.shape[1] * A.shape[2], int(tA), B.data_ptr(),
                              B.shape[2], B.shape[1] * B.shape[2], int(tB), beta, C.data_ptr(), N, M * N, tri, tril_out, st))


def check(name, batch, M, N, K, tA, tB, tri=0, tril_out=0, beta=0.0):
    opA, opB, A, B = make(batch, M, N, K, tA, tB, tri)
    ref = torch.matmul(opA, opB)
    if tril_out:
        ref = torch.tril(ref)
    C0 = torch.randn(batch, M, N, dtype=torch.float64, device=dev)
    ref = 0.5 * ref + beta * C0
    C = C0.clone()
    ws = torch.empty(lib.gpfx_gemm_tf32x3_workspace_bytes(batch, M, N, K) // 4 + 16, dtype=torch.float32, device=dev)
    run_tf32(batch, M, N, K, tA, tB, A, B, C, ws, 0.5, beta, tri, tril_out)
    torch.cuda.synchronize()
    # error relative to the size of the sums involved: sqrt(K) for unit-variance operands
    scale = float(ref.abs().max())
    err = float((C - ref).abs().max()) / scale
    ok = err < 2e-5 and bool(torch.isfinite(C).all())
    print(("ok   " if ok else "FAIL ") + f"{name}: rel err {err:.3e}")
    return {"case": name, "rel_err": err, "ok": ok}


def time_launch(fn, reps=5):
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


def timing(name, batch, M, N, K, tA, tB, tri=0, tril_out=0):
    opA, opB, A, B = make(batch, M, N, K, tA, tB, tri)
    del opA, opB
    C = torch.empty(batch, M, N, dtype=torch.float64, device=dev)
    ws = torch.empty(lib.gpfx_gemm_tf32x3_workspace_bytes(batch, M, N, K) // 4 + 16, dtype=torch.float32, device=dev)
    ms32 = time_launch(lambda: run_tf32(batch, M, N, K, tA, tB, A, B, C, ws, tri=tri, tril_out=tril_out))
    ms64 = time_launch(lambda: run_f64(batch, M, N, K, tA, tB, A, B, C, tri=tri, tril_out=tril_out))
    algo = 2.0 * batch * M * N * K * (0.5 if (tri or tril_out) else 1.0)
    line = {"case": name, "tf32x3_ms_incl_split": round(ms32, 3), "f64_dmma_ms": round(ms64, 3),
            "tf32x3_effective_tflops": round(algo / ms32 / 1e9, 1), "f64_tflops": round(algo / ms64 / 1e9, 1)}
    print(json.dumps(line))
    return line


res = {"checks": [], "timings": []}
C_ = res["checks"]
C_.append(check("NT 128x128x32 (one tile, one k block)", 1, 128, 128, 32, False, True))
C_.append(check("NT 128x128x256", 1, 128, 128, 256, False, True))
C_.append(check("NT 256x384x200 batch 2", 2, 256, 384, 200, False, True))
C_.append(check("NN 256x384x200 batch 2", 2, 256, 384, 200, False, False))
C_.append(check("TN 256x384x200 batch 2", 2, 256, 384, 200, True, False))
C_.append(check("TT tails 176x336x139 batch 3", 3, 176, 336, 139, True, True))
C_.append(check("lower-A NN 512x640x512", 3, 512, 640, 512, False, False, tri=1))
C_.append(check("upper-A TN 512x640x512", 3, 512, 640, 512, True, False, tri=2))
C_.append(check("tril NT 512x512x1000 beta", 3, 512, 512, 1000, False, True, tril_out=1, beta=1.0))
C_.append(check("NT K=8192", 2, 256, 256, 8192, False, True))
res["all_ok"] = all(c["ok"] for c in C_)
print("ALL OK" if res["all_ok"] else "SOME CHECKS FAILED")
if res["all_ok"] or "--time" in sys.argv:
    T = res["timings"]
    T.append(timing("c3 A = Linv Kuf (NN lower-A) 32x1024x8192x1024", 32, 1024, 8192, 1024, False, False, tri=1))
    T.append(timing("c3 B = Lq^T A (TN upper-A)", 32, 1024, 8192, 1024, True, False, tri=2))
    T.append(timing("c3 dLq = tril(A Bs^T) (NT)", 32, 1024, 1024, 8192, False, True, tril_out=1))
    T.append(timing("dense NT 8192^3", 1, 8192, 8192, 8192, False, True))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_tf32_check.json", "w"), indent=1)
