"""The TMA-fed persistent DMMA GEMM (csrc/gemm_f64_tma.cu) against torch.matmul and against the cp.async
kernel on the same inputs: every operand layout (NN / NT / TN / TT), ragged M / N / K, padded leading
dimensions, shared (stride-0) operands, triangular operand flags, tril outputs (dead tiles must still be
written as beta*C), split-K, beta, and both kernel variants.  The two kernels accumulate k in the same
order, so they agree exactly except under split-K."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TRI_LOWER_A, TRI_UPPER_A, TRI_LOWER_B, TRI_UPPER_B = 1, 2, 4, 8


def _make(batch, M, N, K, tri, tA, tB, sA0, sB0, pad, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    na, nb = (1 if sA0 else batch), (1 if sB0 else batch)
    opA = torch.randn(na, M, K, dtype=torch.float64, device="cuda", generator=g)
    opB = torch.randn(nb, K, N, dtype=torch.float64, device="cuda", generator=g)
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
    if pad:
        A = torch.nn.functional.pad(A, (0, pad)).contiguous()
        B = torch.nn.functional.pad(B, (0, pad)).contiguous()
    return opA, opB, A, B


def _run(lib, mode, batch, M, N, K, tri, tA, tB, A, B, C, sA0, sB0, tril_out, splits, ws, alpha, beta):
    from gpflux_b200 import _cabi

    lib.gpfx_tune_gemm_config(100 + mode)  # 100: cp.async kernel, 101 / 102: TMA variants
    try:
        st = torch.cuda.current_stream().cuda_stream
        _cabi.check(lib.gpfx_gemm_splitk(batch, M, N, K, alpha, A.data_ptr(), A.shape[2], 0 if sA0 else A.shape[1] * A.shape[2],
                                         int(tA), B.data_ptr(), B.shape[2], 0 if sB0 else B.shape[1] * B.shape[2], int(tB),
                                         beta, C.data_ptr(), N, M * N, tri, tril_out, splits, ws.data_ptr(), st))
    finally:
        lib.gpfx_tune_gemm_config(-1)


CASES = [
    # name, batch, M, N, K, tri, tA, tB, shared A, shared B, tril_out, splits, pad, beta
    ("NN", 2, 256, 384, 200, 0, False, False, False, False, 0, 1, 0, 0.0),
    ("NT", 2, 256, 384, 200, 0, False, True, False, False, 0, 1, 0, 0.0),
    ("TN", 2, 256, 384, 200, 0, True, False, False, False, 0, 1, 0, 0.0),
    ("TT", 2, 256, 384, 200, 0, True, True, False, False, 0, 1, 0, 0.0),
    ("NN tails + padded ld", 3, 176, 336, 139, 0, False, False, False, False, 0, 1, 2, 0.0),
    ("TT tails + padded ld", 3, 176, 336, 139, 0, True, True, False, False, 0, 1, 2, 0.0),
    ("NT, M and N not multiples of 16", 2, 137, 201, 78, 0, False, True, False, False, 0, 1, 0, 0.0),
    ("lower-A NN (A = Linv Kuf)", 3, 512, 640, 512, TRI_LOWER_A, False, False, False, False, 0, 1, 0, 0.0),
    ("upper-A TN (B = Lq^T A)", 3, 512, 640, 512, TRI_UPPER_A, True, False, False, False, 0, 1, 0, 0.0),
    ("upper-A TN, shared B (c2)", 4, 256, 1024, 256, TRI_UPPER_A, True, False, False, True, 0, 1, 0, 0.0),
    ("lower-A x lower-B NN", 2, 512, 512, 512, TRI_LOWER_A | TRI_LOWER_B, False, False, False, False, 0, 1, 0, 0.0),
    ("lower-A NT upper-B", 2, 384, 384, 384, TRI_LOWER_A | TRI_UPPER_B, False, True, False, False, 0, 1, 0, 0.0),
    ("tril output NT (dLq)", 3, 512, 512, 1000, 0, False, True, False, False, 1, 1, 0, 0.0),
    ("tril output NT, shared A, beta", 3, 256, 256, 2048, 0, False, True, True, False, 1, 1, 0, 1.0),
    ("tril output NT split-K", 2, 256, 256, 4096, 0, False, True, False, False, 1, 5, 0, 0.0),
    ("upper-A TN split-K", 1, 256, 512, 256, TRI_UPPER_A, True, False, False, False, 0, 3, 0, 0.0),
    ("lower-A NN beta (dA)", 2, 256, 512, 256, TRI_LOWER_A, False, False, False, False, 0, 1, 0, 1.0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tma_gemm_matches_torch_and_cp_async(case):
    from gpflux_b200 import _cabi

    name, batch, M, N, K, tri, tA, tB, sA0, sB0, tril_out, splits, pad, beta = case
    lib = _cabi.load()
    opA, opB, A, B = _make(batch, M, N, K, tri, tA, tB, sA0, sB0, pad, 1000 + M + 3 * N + 7 * K)
    ref = torch.matmul(opA, opB).expand(batch, M, N)
    if tril_out:
        ref = torch.tril(ref)
    C0 = torch.randn(batch, M, N, dtype=torch.float64, device="cuda")
    ref = 0.5 * ref + beta * C0
    ws = torch.empty(max(1, splits * batch * M * N), dtype=torch.float64, device="cuda")
    outs = {}
    for mode in (0, 1, 2):
        C = C0.clone()
        _run(lib, mode, batch, M, N, K, tri, tA, tB, A, B, C, sA0, sB0, tril_out, splits, ws, 0.5, beta)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(C).all()), f"{name}: mode {mode} produced non-finite values"
        outs[mode] = C
    scale = float(ref.abs().max())
    for mode, C in outs.items():
        assert float((C - ref).abs().max()) <= 1e-12 * scale, f"{name}: mode {mode} vs torch.matmul"
    if splits == 1:
        assert torch.equal(outs[1], outs[0]) and torch.equal(outs[2], outs[0]), f"{name}: TMA vs cp.async kernel"


def test_tma_kernel_is_selected_for_the_bench_products():
    """Large products go through the TMA kernel by default (profile cfg ids 101 / 102), small grids do not."""
    from gpflux_b200 import _cabi

    lib = _cabi.load()
    A = torch.tril(torch.randn(8, 256, 256, dtype=torch.float64, device="cuda"))
    B = torch.randn(8, 256, 4096, dtype=torch.float64, device="cuda")
    C = torch.empty(8, 256, 4096, dtype=torch.float64, device="cuda")
    E = torch.randn(8, 1024, 2048, dtype=torch.float64, device="cuda")
    F = torch.empty(8, 1024, 1024, dtype=torch.float64, device="cuda")
    G = torch.empty(1, 256, 256, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib.gpfx_gemm_profile_read(None, 0)
    lib.gpfx_gemm_profile_enable(1)
    try:
        # lower-A NN (A = Linv Kuf shape)
        _cabi.check(lib.gpfx_gemm(8, 256, 4096, 256, 1.0, A.data_ptr(), 256, 256 * 256, 0, B.data_ptr(), 4096, 256 * 4096, 0, 0.0,
                                  C.data_ptr(), 4096, 256 * 4096, TRI_LOWER_A, 0, st))
        # tril output NT (dLq shape)
        _cabi.check(lib.gpfx_gemm(8, 1024, 1024, 2048, 1.0, E.data_ptr(), 2048, 1024 * 2048, 0, E.data_ptr(), 2048, 1024 * 2048, 1,
                                  0.0, F.data_ptr(), 1024, 1024 * 1024, 0, 1, st))
        # a single small product: four tiles cannot fill the machine, the cp.async kernel keeps it
        _cabi.check(lib.gpfx_gemm(1, 256, 256, 256, 1.0, A.data_ptr(), 256, 0, 0, A.data_ptr(), 256, 0, 1, 0.0, G.data_ptr(), 256, 0,
                                  0, 0, st))
        torch.cuda.synchronize()
    finally:
        lib.gpfx_gemm_profile_enable(0)
    buf = (_cabi.GemmSample * 8)()
    assert lib.gpfx_gemm_profile_read(buf, 8) == 3
    assert buf[0].cfg == 102  # triangular A: 64x128 tiles, two CTAs per SM
    assert buf[1].cfg == 101  # tril output: 128x128 tiles
    assert buf[2].cfg < 100
    assert float((C - torch.matmul(A, B)).abs().max()) <= 1e-10
    ref = torch.tril(torch.matmul(E, E.transpose(1, 2)))
    assert float((F - ref).abs().max()) <= 1e-12 * float(ref.abs().max())


@pytest.mark.parametrize("batch,M,K,tril_out,splits,shared_a", [
    (8, 1024, 512, 1, 1, False),   # dLq of a separate-kernel layer: tril output, static masks on the diagonal tiles
    (40, 384, 2000 - 2000 % 16, 0, 1, True),   # dense output, shared A (stride 0)
    (4, 512, 4096, 1, 5, False),   # split-K
])
def test_per_k_scale_inside_the_nt_product(batch, M, K, tril_out, splits, shared_a):
    """C = alpha * A diag(s) B^T with the scale applied to the operand fragments (gpfx_gemm_kscaled): against
    torch on an explicitly scaled copy, and against the unscaled kernel fed the scaled copy."""
    from gpflux_b200 import _cabi

    lib = _cabi.load()
    g = torch.Generator(device="cuda").manual_seed(batch * 7 + K)
    A = torch.randn(1 if shared_a else batch, M, K, dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn(batch, M, K, dtype=torch.float64, device="cuda", generator=g)
    s = torch.randn(batch, K, dtype=torch.float64, device="cuda", generator=g)
    Bs = (B * s[:, None, :]).contiguous()
    ref = 0.75 * torch.matmul(A, Bs.transpose(1, 2)).expand(batch, M, M)
    if tril_out:
        ref = torch.tril(ref)
    ws = torch.empty(max(1, splits * batch * M * M), dtype=torch.float64, device="cuda")
    C = torch.full((batch, M, M), float("nan"), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.gpfx_gemm_kscaled(batch, M, M, K, 0.75, A.data_ptr(), K, 0 if shared_a else M * K, B.data_ptr(), K, M * K,
                                      s.data_ptr(), K, C.data_ptr(), M, M * M, tril_out, splits, ws.data_ptr(), st))
    C2 = torch.full((batch, M, M), float("nan"), dtype=torch.float64, device="cuda")
    _cabi.check(lib.gpfx_gemm_splitk(batch, M, M, K, 0.75, A.data_ptr(), K, 0 if shared_a else M * K, 0, Bs.data_ptr(), K, M * K, 1,
                                     0.0, C2.data_ptr(), M, M * M, 0, tril_out, splits, ws.data_ptr(), st))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(C).all())
    scale = float(ref.abs().max())
    assert float((C - ref).abs().max()) <= 1e-12 * scale
    # same products, same order: the scaled-copy route rounds s*B once as well, so the two agree exactly
    assert torch.equal(C, C2)


def test_per_k_scale_refuses_shapes_the_tma_kernel_does_not_take():
    from gpflux_b200 import _cabi

    lib = _cabi.load()
    A = torch.randn(2, 64, 64, dtype=torch.float64, device="cuda")
    s = torch.randn(2, 64, dtype=torch.float64, device="cuda")
    C = torch.empty(2, 64, 64, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.gpfx_gemm_kscaled(2, 64, 64, 64, 1.0, A.data_ptr(), 64, 4096, A.data_ptr(), 64, 4096, s.data_ptr(), 64,
                               C.data_ptr(), 64, 4096, 0, 1, None, st)
    assert rc == -7  # GPFX_E_UNSUPPORTED
