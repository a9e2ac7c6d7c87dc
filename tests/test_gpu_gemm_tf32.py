"""gpfx_gemm_tf32x3 - the tcgen05 (kind::tf32, TMEM accumulators) contraction of the float32 opt-in - against
torch.matmul in float64.  Tolerance 2e-5 relative to max|C| (3xTF32 carries 22 significand bits per operand;
float32 accumulation is cut into k-chunks of 1024 that are added in float64); the north-star bound for the
opt-in is 1e-4."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-5

CASES = [
    # name, batch, M, N, K, tA, tB, tri, tril_out, beta
    ("NT one tile, one k block", 1, 128, 128, 32, False, True, 0, 0, 0.0),
    ("NT 256x384x200", 2, 256, 384, 200, False, True, 0, 0, 0.0),
    ("NN 256x384x200", 2, 256, 384, 200, False, False, 0, 0, 0.0),
    ("TN 256x384x200", 2, 256, 384, 200, True, False, 0, 0, 0.0),
    ("TT ragged 176x336x139", 3, 176, 336, 139, True, True, 0, 0, 0.0),
    ("lower-A NN", 3, 512, 640, 512, False, False, 1, 0, 0.0),
    ("upper-A TN", 3, 512, 640, 512, True, False, 2, 0, 0.0),
    ("tril output NT, beta", 3, 512, 512, 1000, False, True, 0, 1, 1.0),
    ("NT K = 8192 (k-chunked accumulation)", 2, 256, 256, 8192, False, True, 0, 0, 0.0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tf32x3_gemm_matches_float64_matmul(case):
    from gpflux_b200 import _cabi

    name, batch, M, N, K, tA, tB, tri, tril_out, beta = case
    lib = _cabi.load()
    g = torch.Generator(device="cuda").manual_seed(7 + M + 3 * N + 5 * K)
    opA = torch.randn(batch, M, K, dtype=torch.float64, device="cuda", generator=g)
    opB = torch.randn(batch, K, N, dtype=torch.float64, device="cuda", generator=g)
    if tri & 1:
        opA = torch.tril(opA)
    if tri & 2:
        opA = torch.triu(opA)
    A = opA.transpose(1, 2).contiguous() if tA else opA.contiguous()
    B = opB.transpose(1, 2).contiguous() if tB else opB.contiguous()
    ref = torch.matmul(opA, opB)
    if tril_out:
        ref = torch.tril(ref)
    C0 = torch.randn(batch, M, N, dtype=torch.float64, device="cuda")
    ref = 0.5 * ref + beta * C0
    C = C0.clone()
    ws = torch.empty(lib.gpfx_gemm_tf32x3_workspace_bytes(batch, M, N, K) // 4 + 16, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.gpfx_gemm_tf32x3(batch, M, N, K, 0.5, A.data_ptr(), A.shape[2], A.shape[1] * A.shape[2], int(tA), B.data_ptr(),
                                     B.shape[2], B.shape[1] * B.shape[2], int(tB), beta, C.data_ptr(), N, M * N, tri, tril_out,
                                     ws.data_ptr(), st))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(C).all())
    err = float((C - ref).abs().max()) / float(ref.abs().max())
    assert err <= TOL, f"{name}: relative error {err:.3e}"
    assert err > 1e-12, f"{name}: suspiciously exact - did the float64 kernel run instead?"
