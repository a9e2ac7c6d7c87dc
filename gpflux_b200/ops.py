"""torch-tensor front end of the C-ABI (``include/gpfx.h``).

PyTorch is plumbing here: it owns device memory, the CUDA stream and autograd bookkeeping; every
numerical operation below is a call into libgpflux_b200.so with raw device pointers.  There is no
CPU or eager-PyTorch fallback - tensors that are not CUDA float64 raise.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Optional, Tuple

import torch

from . import _cabi
from ._cabi import KERNEL_IDS, SOLVER_IDS, LayerDesc, check

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Optional[torch.Tensor], name: str, shape=None) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (gpflux_b200 has no CPU path), got {t.device}")
    if t.dtype != torch.float64:
        # mirrors DeepGP's dtype check (reference gpflux/models/deep_gp.py:145-149)
        raise ValueError(f"{name}: expected dtype float64 (gpflow default_float), got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.contiguous()


def _fp32_optin(keep_outputs: bool = False):
    """float32 opt-in (gpflow.config.set_default_float(np.float32), read at reference
    gp_layer.py:163,213,220): when it is on, float32 tensor arguments are widened to float64 (an
    autograd-tracked cast, so gradients come back as float32) and float64 results are rounded to float32.
    ``keep_outputs``: leave the results alone (opaque float64 context buffers)."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            from .gpflow_compat import default_float

            if default_float() is not torch.float32:
                return fn(*args, **kwargs)

            def up(x):
                if isinstance(x, torch.Tensor):
                    return x.double() if x.dtype == torch.float32 else x
                if isinstance(x, (tuple, list)):
                    return type(x)(up(v) for v in x)
                return x

            def down(x):
                if isinstance(x, torch.Tensor):
                    return x.float() if x.dtype == torch.float64 else x
                if isinstance(x, (tuple, list)):
                    return type(x)(down(v) for v in x)
                return x

            out = fn(*up(args), **{k: up(v) for k, v in kwargs.items()})
            return out if keep_outputs else down(out)
        return wrapper
    return deco


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_workspaces = {}

# Cholesky status sink: a GPLayer points this at its persistent int32 [K] device buffer around its calls;
# the forward entry points then copy the factorisation status there with gpfx_layer_status_async (device to
# device, stream-ordered, no sync) so that a non-PD Kuu can be raised lazily (GPLayer.check_status()).
_status_sink = None


class status_sink:
    def __init__(self, buf):
        self.buf = buf

    def __enter__(self):
        global _status_sink
        self.prev, _status_sink = _status_sink, self.buf
        return self

    def __exit__(self, *exc):
        global _status_sink
        _status_sink = self.prev
        return False


def _record_status(lib, desc, ctx_buf):
    sink = _status_sink
    if sink is not None and sink.numel() >= desc.K:
        check(lib.gpfx_layer_status_async(C.byref(desc), _ptr(ctx_buf), _ptr(sink), _stream()))


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Grow-only scratch buffer per device (all users are ordered on the current stream)."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def _kernel_matrix_raw(kind: str, U, V, lengthscales, variance, jitter: Optional[float] = None):
    """out[k] = k_k(U[k], V[k]) (+ jitter I).  U [K,nu,D]; V [K,nv,D] or shared [nv,D];
    lengthscales [K,D]; variance [K]."""
    lib = _cabi.load()
    U = _chk(U, "U")
    V = _chk(V, "V")
    K, nu, D = U.shape
    v_shared = V.dim() == 2
    nv = V.shape[-2]
    ls = _chk(lengthscales, "lengthscales", (K, D))
    var = _chk(variance, "variance", (K,))
    out = torch.empty((K, nu, nv), dtype=torch.float64, device=U.device)
    check(lib.gpfx_kernel_matrix(KERNEL_IDS[kind], K, nu, nv, D, _ptr(U), 0, _ptr(V), int(v_shared),
                                 _ptr(ls), _ptr(var), int(jitter is not None), float(jitter or 0.0),
                                 _ptr(out), _stream()))
    return out



class _KernelMatrixFn(torch.autograd.Function):
    """Kernel matrix with the hand-written backward (gpfx_kernel_matrix_bwd)."""

    @staticmethod
    def forward(ctx, kind, U, V, ls, var, jitter):
        ctx.kind = kind
        ctx.save_for_backward(U, V, ls, var)
        return _kernel_matrix_raw(kind, U, V, ls, var, jitter)

    @staticmethod
    def backward(ctx, g):
        U, V, ls, var = ctx.saved_tensors
        dU, dV, dls, dvar = kernel_matrix_bwd(ctx.kind, U, V, ls, var, g.contiguous())
        return None, dU, dV, dls, dvar, None


@_fp32_optin()
def kernel_matrix(kind: str, U, V, lengthscales, variance, jitter: Optional[float] = None):
    """out[k] = k_k(U[k], V[k]) (+ jitter I): U [K,nu,D]; V [K,nv,D] or [nv,D] (shared); lengthscales [K,D];
    variance [K].  Differentiable w.r.t. U, V, lengthscales and variance."""
    ts = [t for t in (U, V, lengthscales, variance) if isinstance(t, torch.Tensor)]
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts):
        return _KernelMatrixFn.apply(kind, _chk(U, "U"), _chk(V, "V"), _chk(lengthscales, "lengthscales"),
                                     _chk(variance, "variance"), jitter)
    return _kernel_matrix_raw(kind, U, V, lengthscales, variance, jitter)


@_fp32_optin()
def kernel_matrix_bwd(kind: str, U, V, lengthscales, variance, dK):
    """-> dU [K,nu,D], dV ([nv,D] if V shared else [K,nv,D]), dlengthscales [K,D], dvariance [K]."""
    lib = _cabi.load()
    U = _chk(U, "U")
    V = _chk(V, "V")
    K, nu, D = U.shape
    v_shared = V.dim() == 2
    nv = V.shape[-2]
    ls = _chk(lengthscales, "lengthscales", (K, D))
    var = _chk(variance, "variance", (K,))
    dK = _chk(dK, "dK", (K, nu, nv)).clone()
    dU = torch.empty_like(U)
    dV = torch.empty_like(V)
    dls = torch.empty_like(ls)
    dvar = torch.empty_like(var)
    scratch = _workspace(lib.gpfx_kernel_matrix_bwd_scratch_bytes(K, nu, nv, D), U.device)
    check(lib.gpfx_kernel_matrix_bwd(KERNEL_IDS[kind], K, nu, nv, D, _ptr(U), _ptr(V), int(v_shared),
                                     _ptr(ls), _ptr(var), _ptr(dK), _ptr(dU), _ptr(dV), _ptr(dls),
                                     _ptr(dvar), _ptr(scratch), _stream()))
    return dU, dV, dls, dvar


def _cholesky_inverse_raw(A) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """A [K,M,M] SPD -> (L, L^-1, info[K] int32)."""
    lib = _cabi.load()
    A = _chk(A, "A")
    K, M, _ = A.shape
    L = torch.empty_like(A)
    Linv = torch.empty_like(A)
    info = torch.empty((K,), dtype=torch.int32, device=A.device)
    ws = _workspace(lib.gpfx_cholesky_workspace_bytes(K, M), A.device)
    check(lib.gpfx_cholesky_inverse(K, M, _ptr(A), _ptr(L), _ptr(Linv), _ptr(info), _ptr(ws), _stream()))
    return L, Linv, info



def cholesky_backward(L, Linv, dL):
    """dA = sym(L^-T Phi(L^T dL) L^-1): the adjoint of A -> chol(A) (gpfx_cholesky_backward)."""
    lib = _cabi.load()
    L = _chk(L, "L")
    K, M, _ = L.shape
    Linv = _chk(Linv, "Linv", (K, M, M))
    dL = _chk(dL, "dL", (K, M, M))
    dA = torch.empty_like(L)
    ws = _workspace(lib.gpfx_cholesky_backward_workspace_bytes(K, M), L.device)
    check(lib.gpfx_cholesky_backward(K, M, _ptr(L), _ptr(Linv), _ptr(dL), _ptr(dA), _ptr(ws), _stream()))
    return dA


class _CholInvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        L, Linv, info = _cholesky_inverse_raw(A)
        ctx.save_for_backward(L, Linv)
        ctx.mark_non_differentiable(info)
        return L, Linv, info

    @staticmethod
    def backward(ctx, dL, dLinv, _dinfo):
        L, Linv = ctx.saved_tensors
        g = torch.zeros_like(L) if dL is None else torch.tril(dL)
        if dLinv is not None:
            # Linv = L^-1  ->  dL -= tril(L^-T dLinv L^-T)
            t = _gemm_raw(Linv, torch.tril(dLinv).contiguous(), transA=True)
            g = g - torch.tril(_gemm_raw(t, Linv, transB=True))
        return cholesky_backward(L, Linv, g.contiguous())


@_fp32_optin()
def cholesky_inverse(A) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """A [K,M,M] SPD -> (L, L^-1, info[K] int32).  Differentiable w.r.t. A."""
    if torch.is_grad_enabled() and A.requires_grad:
        return _CholInvFn.apply(_chk(A, "A"))
    return _cholesky_inverse_raw(A)


@_fp32_optin()
def natgrad_step(q_mu, q_sqrt, dq_mu, dq_sqrt, gamma: float):
    """One natural-gradient step (XiNat) for q(u) = N(q_mu, q_sqrt q_sqrt^T) given the ordinary
    gradients of the loss -> (q_mu_new [M,P], q_sqrt_new [P,M,M], info [3,P] int32).  Replaces
    gpflow.optimizers.NaturalGradient._natgrad_apply_gradients (keras_natgrad.py:187-191)."""
    lib = _cabi.load()
    q_mu = _chk(q_mu, "q_mu")
    M, P = q_mu.shape
    q_sqrt = _chk(q_sqrt, "q_sqrt", (P, M, M))
    dq_mu = _chk(dq_mu, "dq_mu", (M, P))
    dq_sqrt = _chk(dq_sqrt, "dq_sqrt", (P, M, M))
    q_mu_new = torch.empty_like(q_mu)
    q_sqrt_new = torch.empty_like(q_sqrt)
    info = torch.empty((3, P), dtype=torch.int32, device=q_mu.device)
    ws = _workspace(lib.gpfx_natgrad_workspace_bytes(M, P), q_mu.device)
    check(lib.gpfx_natgrad_step(M, P, float(gamma), _ptr(q_mu), _ptr(q_sqrt), _ptr(dq_mu), _ptr(dq_sqrt),
                                _ptr(q_mu_new), _ptr(q_sqrt_new), _ptr(info), _ptr(ws), _stream()))
    return q_mu_new, q_sqrt_new, info


def _gemm_raw(A, B, *, transA=False, transB=False, alpha=1.0, beta=0.0, C_in=None, tri=0, tril_out=False):
    """Batched C = alpha op(A) op(B) + beta C on the FP64 tensor pipe.  A, B: [b, r, c]."""
    lib = _cabi.load()
    A = _chk(A, "A")
    B = _chk(B, "B")
    b = A.shape[0]
    M, Kd = (A.shape[2], A.shape[1]) if transA else (A.shape[1], A.shape[2])
    N = B.shape[1] if transB else B.shape[2]
    Cout = torch.zeros((b, M, N), dtype=torch.float64, device=A.device) if C_in is None else _chk(C_in, "C").clone()
    check(lib.gpfx_gemm(b, M, N, Kd, float(alpha), _ptr(A), A.shape[2], A.shape[1] * A.shape[2], int(transA),
                        _ptr(B), B.shape[2], B.shape[1] * B.shape[2], int(transB), float(beta), _ptr(Cout), N,
                        M * N, int(tri), int(tril_out), _stream()))
    return Cout



class _GemmFn(torch.autograd.Function):
    """C = alpha op(A) op(B) + beta C_in with the transposed products as backward (all on the DMMA GEMM)."""

    @staticmethod
    def forward(ctx, A, B, C_in, transA, transB, alpha, beta, tri, tril_out):
        ctx.save_for_backward(A, B)
        ctx.cfg = (transA, transB, alpha, beta, tri, tril_out, C_in is not None)
        return _gemm_raw(A, B, transA=transA, transB=transB, alpha=alpha, beta=beta, C_in=C_in, tri=tri,
                         tril_out=tril_out)

    @staticmethod
    def backward(ctx, G):
        A, B = ctx.saved_tensors
        transA, transB, alpha, beta, tri, tril_out, has_c = ctx.cfg
        G = (torch.tril(G) if tril_out else G).contiguous()
        dA = dB = dC = None
        if ctx.needs_input_grad[0]:
            # d op(A) = alpha G op(B)^T, masked to the declared triangle of op(A)
            if transA:
                dA = _gemm_raw(B, G, transA=transB, transB=True, alpha=alpha)      # (d op(A))^T = op(B) G^T
                if tri & _cabi.TRI_LOWER_A:
                    dA = torch.triu(dA)
                if tri & _cabi.TRI_UPPER_A:
                    dA = torch.tril(dA)
            else:
                dA = _gemm_raw(G, B, transB=not transB, alpha=alpha)
                if tri & _cabi.TRI_LOWER_A:
                    dA = torch.tril(dA)
                if tri & _cabi.TRI_UPPER_A:
                    dA = torch.triu(dA)
        if ctx.needs_input_grad[1]:
            # d op(B) = alpha op(A)^T G, masked to the declared triangle of op(B)
            if transB:
                dB = _gemm_raw(G, A, transA=True, transB=transA, alpha=alpha)      # (d op(B))^T = G^T op(A)
                if tri & _cabi.TRI_LOWER_B:
                    dB = torch.triu(dB)
                if tri & _cabi.TRI_UPPER_B:
                    dB = torch.tril(dB)
            else:
                dB = _gemm_raw(A, G, transA=not transA, alpha=alpha)
                if tri & _cabi.TRI_LOWER_B:
                    dB = torch.tril(dB)
                if tri & _cabi.TRI_UPPER_B:
                    dB = torch.triu(dB)
        if has_c and ctx.needs_input_grad[2]:
            dC = beta * G
        return dA, dB, dC, None, None, None, None, None, None


@_fp32_optin()
def gemm(A, B, *, transA=False, transB=False, alpha=1.0, beta=0.0, C_in=None, tri=0, tril_out=False):
    """Batched C = alpha op(A) op(B) + beta C_in on the FP64 tensor pipe.  A, B: [b, r, c].  ``tri`` declares
    triangular operands (their zero triangle is skipped).  Differentiable w.r.t. A, B and C_in."""
    ts = [t for t in (A, B, C_in) if isinstance(t, torch.Tensor)]
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts):
        return _GemmFn.apply(_chk(A, "A"), _chk(B, "B"), None if C_in is None else _chk(C_in, "C"), bool(transA),
                             bool(transB), float(alpha), float(beta), int(tri), bool(tril_out))
    return _gemm_raw(A, B, transA=transA, transB=transB, alpha=alpha, beta=beta, C_in=C_in, tri=tri, tril_out=tril_out)


@_fp32_optin()
def reparam_sample(mean, var, eps):
    """f = mean + sqrt(var) * eps (reference gpflux/layers/gp_layer.py:339-363)."""
    lib = _cabi.load()
    mean = _chk(mean, "mean")
    var = _chk(var, "var", mean.shape)
    eps = _chk(eps, "eps", mean.shape)
    f = torch.empty_like(mean)
    check(lib.gpfx_reparam_sample(_ptr(mean), _ptr(var), _ptr(eps), _ptr(f), mean.numel(), _stream()))
    return f


_U64 = (1 << 64) - 1


def philox_raw(seed: int, offset: int, n_pairs: int, device="cuda"):
    """[n_pairs, 4] words of the Philox4x32-10 stream behind ``randn_philox`` (int64 tensor holding
    uint32 values; what the bit-exact test compares with the oracle)."""
    lib = _cabi.load()
    out = torch.empty((n_pairs, 4), dtype=torch.int32, device=device)
    check(lib.gpfx_philox_raw(seed & _U64, offset & _U64, _ptr(out), n_pairs, _stream()))
    return out.to(torch.int64) & 0xFFFFFFFF


def randn_philox(shape, seed: int, offset: int = 0, device="cuda"):
    """N(0,1) float64 tensor drawn in one kernel from (seed, offset); a pure function of its
    arguments, so the same call regenerates the same noise (e.g. in a backward pass)."""
    lib = _cabi.load()
    out = torch.empty(shape, dtype=torch.float64, device=device)
    check(lib.gpfx_randn_philox(seed & _U64, offset & _U64, _ptr(out), out.numel(), _stream()))
    return out


def reparam_sample_philox(mean, var, seed: int, offset: int = 0, return_eps: bool = False):
    """f = mean + sqrt(var) * eps with eps generated inside the kernel (no eps array is read).
    Equal in distribution to the reference's TFP draw (gp_layer.py:358-363), not stream-identical."""
    lib = _cabi.load()
    mean = _chk(mean, "mean")
    var = _chk(var, "var", mean.shape)
    f = torch.empty_like(mean)
    eps = torch.empty_like(mean) if return_eps else None
    check(lib.gpfx_reparam_sample_philox(_ptr(mean), _ptr(var), seed & _U64, offset & _U64, _ptr(f),
                                         _ptr(eps) if return_eps else None, mean.numel(), _stream()))
    return (f, eps) if return_eps else f


@_fp32_optin()
def kl_white(q_mu, q_sqrt):
    """Whitened KL[q(v) || N(0, I)] (reference gpflux/layers/gp_layer.py:303-311)."""
    lib = _cabi.load()
    q_mu = _chk(q_mu, "q_mu")
    M, P = q_mu.shape
    q_sqrt = _chk(q_sqrt, "q_sqrt", (P, M, M))
    kl = torch.empty((), dtype=torch.float64, device=q_mu.device)
    scratch = _workspace(lib.gpfx_kl_scratch_bytes(M, P), q_mu.device)
    check(lib.gpfx_kl_white(M, P, _ptr(q_mu), _ptr(q_sqrt), _ptr(kl), _ptr(scratch), _stream()))
    return kl


@_fp32_optin()
def kl_white_bwd(q_mu, q_sqrt, g_kl):
    """-> (dq_mu, dq_sqrt) of g_kl * KL."""
    lib = _cabi.load()
    q_mu = _chk(q_mu, "q_mu")
    M, P = q_mu.shape
    q_sqrt = _chk(q_sqrt, "q_sqrt", (P, M, M))
    g_kl = _chk(g_kl.reshape(()), "g_kl")
    dq_mu = torch.empty_like(q_mu)
    dq_sqrt = torch.empty_like(q_sqrt)
    check(lib.gpfx_kl_white_bwd(M, P, _ptr(q_mu), _ptr(q_sqrt), _ptr(g_kl), _ptr(dq_mu), _ptr(dq_sqrt), _stream()))
    return dq_mu, dq_sqrt


# --------------------------------------------------------------------------------------------
# the SVGP layer (autograd)
# --------------------------------------------------------------------------------------------
def _make_desc(N, M, D, P, K, kernel: str, whiten: bool, jitter: float, solver: str = "gemm") -> LayerDesc:
    if kernel not in KERNEL_IDS:
        raise NotImplementedError(f"kernel {kernel!r} is not supported by the B200 SVGP path")
    if solver not in SOLVER_IDS:
        raise ValueError(f"solver must be one of {sorted(SOLVER_IDS)}, got {solver!r}")
    desc = LayerDesc(N, M, D, P, K, KERNEL_IDS[kernel], int(bool(whiten)), SOLVER_IDS[solver], float(jitter))
    _prepare_contraction(desc)
    return desc


_tf32_state = {"mode": 0, "ws": None}


def _prepare_contraction(desc: LayerDesc) -> None:
    """float32 opt-in: switch the library's contraction mode to 3xTF32 (gpfx_set_contraction) and hand it a
    float32 workspace large enough for this layer's products; back to FP64 DMMA when the opt-in is off.  Runs
    before every ctx / workspace size query of a layer call (the column-partial layout depends on the mode).
    ``GPFX_FP32_TF32=0`` keeps the float32 opt-in on the FP64 contraction path."""
    import os

    from .gpflow_compat import default_float

    want = 1 if (default_float() is torch.float32 and os.environ.get("GPFX_FP32_TF32", "1") != "0") else 0
    st = _tf32_state
    if want == 0:
        if st["mode"] != 0:
            check(_cabi.load().gpfx_set_contraction(0, None, 0))
            st["mode"] = 0
        return
    lib = _cabi.load()
    need = int(lib.gpfx_layer_tf32_workspace_bytes(C.byref(desc)))
    ws = st["ws"]
    if st["mode"] != 1 or ws is None or ws.numel() < need or ws.device != torch.device("cuda", torch.cuda.current_device()):
        if ws is None or ws.numel() < need or ws.device != torch.device("cuda", torch.cuda.current_device()):
            ws = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device="cuda")
            st["ws"] = ws
        check(lib.gpfx_set_contraction(1, ws.data_ptr(), ws.numel()))
        st["mode"] = 1


class _SVGPLayerFn(torch.autograd.Function):
    """(X, Z, lengthscales, variance, q_mu, q_sqrt, mean_add, eps) -> (fsample, fmean, fvar, kl)."""

    @staticmethod
    def forward(ctx, X, Z, ls, var, q_mu, q_sqrt, mean_add, eps, kernel, whiten, jitter, solver="gemm"):
        lib = _cabi.load()
        X = _chk(X, "X")
        N, D = X.shape
        Z = _chk(Z, "Z")
        K, M, Dz = Z.shape
        if Dz != D:
            raise ValueError(f"Z has input dim {Dz}, X has {D}")
        q_mu = _chk(q_mu, "q_mu")
        P = q_mu.shape[1]
        q_mu = _chk(q_mu, "q_mu", (M, P))
        q_sqrt = _chk(q_sqrt, "q_sqrt", (P, M, M))
        ls = _chk(ls, "lengthscales", (K, D))
        var = _chk(var, "variance", (K,))
        mean_add = _chk(mean_add, "mean_add", (N, P))
        eps = _chk(eps, "eps", (N, P))
        desc = _make_desc(N, M, D, P, K, kernel, whiten, jitter, solver)
        dev = X.device
        ctx_buf = torch.empty(lib.gpfx_layer_ctx_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
        ws = _workspace(lib.gpfx_layer_workspace_bytes(C.byref(desc)), dev)
        fmean = torch.empty((N, P), dtype=torch.float64, device=dev)
        fvar = torch.empty_like(fmean)
        fsample = torch.empty_like(fmean)
        kl = torch.empty((), dtype=torch.float64, device=dev)
        check(lib.gpfx_layer_forward(C.byref(desc), _ptr(X), _ptr(Z), _ptr(ls), _ptr(var), _ptr(q_mu),
                                     _ptr(q_sqrt), _ptr(mean_add), _ptr(eps), _ptr(fmean), _ptr(fvar),
                                     _ptr(fsample), _ptr(kl), _ptr(ctx_buf), _ptr(ws), _stream()))
        _record_status(lib, desc, ctx_buf)
        ctx.desc = desc
        ctx.ctx_buf = ctx_buf
        ctx.has_mean_add = mean_add is not None
        ctx.has_eps = eps is not None
        ctx.save_for_backward(X, Z, ls, var, q_mu, eps if eps is not None else X.new_empty(0), fvar)
        return fsample, fmean, fvar, kl

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_f, g_mean, g_var, g_kl):
        lib = _cabi.load()
        if ctx.ctx_buf is None:
            raise RuntimeError("gpflux_b200 SVGP layer: backward called twice (ctx is consumed)")
        X, Z, ls, var, q_mu, eps, fvar = ctx.saved_tensors
        eps = eps if ctx.has_eps else None
        desc = ctx.desc
        dev = X.device
        N, M, D, P, K = desc.N, desc.M, desc.D, desc.P, desc.K
        g_f = _chk(g_f, "g_fsample")
        g_mean = _chk(g_mean, "g_fmean")
        g_var = _chk(g_var, "g_fvar")
        g_kl = torch.zeros((), dtype=torch.float64, device=dev) if g_kl is None else _chk(g_kl, "g_kl")
        dX = torch.empty((N, D), dtype=torch.float64, device=dev)
        dZ = torch.empty((K, M, D), dtype=torch.float64, device=dev)
        dls = torch.empty((K, D), dtype=torch.float64, device=dev)
        dvar = torch.empty((K,), dtype=torch.float64, device=dev)
        dq_mu = torch.empty((M, P), dtype=torch.float64, device=dev)
        dq_sqrt = torch.empty((P, M, M), dtype=torch.float64, device=dev)
        dmean = torch.empty((N, P), dtype=torch.float64, device=dev)
        ws = _workspace(lib.gpfx_layer_workspace_bytes(C.byref(desc)), dev)
        check(lib.gpfx_layer_backward(C.byref(desc), _ptr(X), _ptr(Z), _ptr(ls), _ptr(var), _ptr(q_mu),
                                      _ptr(eps), _ptr(fvar), _ptr(g_f), _ptr(g_mean), _ptr(g_var), _ptr(g_kl),
                                      _ptr(dX), _ptr(dZ), _ptr(dls), _ptr(dvar), _ptr(dq_mu), _ptr(dq_sqrt),
                                      _ptr(dmean), _ptr(ctx.ctx_buf), _ptr(ws), _stream()))
        ctx.ctx_buf = None
        # eps gets no gradient (it is the reparameterisation noise)
        return (dX, dZ, dls, dvar, dq_mu, dq_sqrt, dmean if ctx.has_mean_add else None, None, None, None, None, None)


@_fp32_optin()
def svgp_layer(X, Z, lengthscales, variance, q_mu, q_sqrt, *, mean_add=None, eps=None, kernel="se",
               whiten=True, jitter=DEFAULT_JITTER, solver="gemm"):
    """One GPLayer evaluation on the B200 path -> (fsample, fmean, fvar, kl).

    Replaces GPLayer.predict + prior_kl + distribution.sample()
    (reference gpflux/layers/gp_layer.py:226-268, 303-311, 339-363).  ``eps=None`` returns
    fsample == fmean.  ``solver``: "gemm" (A = L^-1 Kuf as an explicit-inverse DMMA GEMM, the fast default)
    or "trsm" (blocked forward substitution with L, the reference-faithful form; see include/gpfx.h)."""
    return _SVGPLayerFn.apply(X, Z, lengthscales, variance, q_mu, q_sqrt, mean_add, eps, kernel, whiten, jitter, solver)


# --------------------------------------------------------------------------------------------
# the same layer as two autograd nodes: parameter-only "factor" + data "cond"
# --------------------------------------------------------------------------------------------
class _FactorFn(torch.autograd.Function):
    """(Z, lengthscales, variance, q_mu, q_sqrt) -> (fctx [opaque bytes], link [K,Mp,Mp], kl).

    ``link`` is the differentiable edge to the cond node: its gradient is dLoss/dL (the Cholesky
    factor's cotangent), which cond_backward produces and factor_backward consumes.  Its values are
    never read (L itself lives inside fctx)."""

    @staticmethod
    def forward(ctx, Z, ls, var, q_mu, q_sqrt, kernel, whiten, jitter):
        lib = _cabi.load()
        Z = _chk(Z, "Z")
        K, M, D = Z.shape
        q_mu = _chk(q_mu, "q_mu")
        P = q_mu.shape[1]
        q_mu = _chk(q_mu, "q_mu", (M, P))
        q_sqrt = _chk(q_sqrt, "q_sqrt", (P, M, M))
        ls = _chk(ls, "lengthscales", (K, D))
        var = _chk(var, "variance", (K,))
        desc = _make_desc(1, M, D, P, K, kernel, whiten, jitter)  # the factor part does not depend on N
        dev = Z.device
        fctx = torch.empty(lib.gpfx_layer_factor_ctx_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
        ws = _workspace(lib.gpfx_layer_factor_workspace_bytes(C.byref(desc)), dev)
        nL = lib.gpfx_layer_dL_bytes(C.byref(desc)) // 8
        Mp = int(round((nL // K) ** 0.5))
        link = torch.empty((K, Mp, Mp), dtype=torch.float64, device=dev)
        kl = torch.empty((), dtype=torch.float64, device=dev)
        check(lib.gpfx_layer_factor_forward(C.byref(desc), _ptr(Z), _ptr(ls), _ptr(var), _ptr(q_mu), _ptr(q_sqrt),
                                            _ptr(kl), _ptr(fctx), _ptr(ws), _stream()))
        _record_status(lib, desc, fctx)
        ctx.desc = desc
        ctx.fctx = fctx
        ctx.save_for_backward(Z, ls, var, q_mu)
        ctx.mark_non_differentiable(fctx)
        # non-whitened: the cond node differentiates w.r.t. the effective (alpha, C) = (L^-1 q_mu,
        # L^-1 Lq) held in fctx; these two links carry those cotangents back to this node, which maps
        # them to q_mu / q_sqrt (and L).  Whitened: the cond node's q_mu / q_sqrt gradients go straight
        # to the parameters and the links are empty.
        if whiten:
            link_a = torch.empty((0,), dtype=torch.float64, device=dev)
            link_c = torch.empty((0,), dtype=torch.float64, device=dev)
        else:
            link_a = torch.empty((M, P), dtype=torch.float64, device=dev)
            link_c = torch.empty((P, M, M), dtype=torch.float64, device=dev)
        return fctx, link, kl, link_a, link_c

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _g_fctx, g_link, g_kl, g_link_a, g_link_c):
        lib = _cabi.load()
        Z, ls, var, q_mu = ctx.saved_tensors
        desc = ctx.desc
        dev = Z.device
        M, D, P, K = desc.M, desc.D, desc.P, desc.K
        dZ = torch.empty((K, M, D), dtype=torch.float64, device=dev)
        dls = torch.empty((K, D), dtype=torch.float64, device=dev)
        dvar = torch.empty((K,), dtype=torch.float64, device=dev)
        acc = 0
        if desc.whiten:
            dq_mu = torch.empty((M, P), dtype=torch.float64, device=dev)
            dq_sqrt = torch.empty((P, M, M), dtype=torch.float64, device=dev)
        else:
            # in/out buffers: data-part cotangents of (alpha, C) in, gradients of (q_mu, q_sqrt) out
            dq_mu = torch.zeros((M, P), dtype=torch.float64, device=dev) if g_link_a is None else g_link_a.contiguous().clone()
            dq_sqrt = (torch.zeros((P, M, M), dtype=torch.float64, device=dev) if g_link_c is None
                       else g_link_c.contiguous().clone())
            acc = 2
        if g_link is None:
            # KL-only use of the factor (no conditional consumed it): dL = 0
            nL = lib.gpfx_layer_dL_bytes(C.byref(desc)) // 8
            g_link = torch.zeros(nL, dtype=torch.float64, device=dev)
        g_link = _chk(g_link, "dL")
        g_link.record_stream(torch.cuda.current_stream())  # produced on the cond node's stream
        g_kl = None if g_kl is None else _chk(g_kl, "g_kl")
        ws = _workspace(lib.gpfx_layer_factor_workspace_bytes(C.byref(desc)), dev)
        check(lib.gpfx_layer_factor_backward(C.byref(desc), _ptr(Z), _ptr(ls), _ptr(var), _ptr(q_mu), _ptr(g_kl),
                                             _ptr(g_link), _ptr(dZ), _ptr(dls), _ptr(dvar), _ptr(dq_mu),
                                             _ptr(dq_sqrt), acc, _ptr(ctx.fctx), _ptr(ws), _stream()))
        return dZ, dls, dvar, dq_mu, dq_sqrt, None, None, None


class _CondFn(torch.autograd.Function):
    """(X, Z, lengthscales, variance, q_mu, q_sqrt, link, fctx, mean_add, eps) -> (fsample, fmean, fvar).
    q_sqrt and link are taken for gradient routing only (their values live in fctx)."""

    @staticmethod
    def forward(ctx, X, Z, ls, var, q_mu, q_sqrt, link, link_a, link_c, fctx, mean_add, eps, kernel, whiten, jitter,
                solver="gemm"):
        lib = _cabi.load()
        X = _chk(X, "X")
        N, D = X.shape
        Z = _chk(Z, "Z")
        K, M, Dz = Z.shape
        if Dz != D:
            raise ValueError(f"Z has input dim {Dz}, X has {D}")
        q_mu = _chk(q_mu, "q_mu")
        P = q_mu.shape[1]
        ls = _chk(ls, "lengthscales", (K, D))
        var = _chk(var, "variance", (K,))
        mean_add = _chk(mean_add, "mean_add", (N, P))
        eps = _chk(eps, "eps", (N, P))
        desc = _make_desc(N, M, D, P, K, kernel, whiten, jitter, solver)
        dev = X.device
        if fctx.numel() != lib.gpfx_layer_factor_ctx_bytes(C.byref(desc)):
            raise ValueError("factor context does not match this layer's shape")
        cctx = torch.empty(lib.gpfx_layer_cond_ctx_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
        ws = _workspace(lib.gpfx_layer_cond_workspace_bytes(C.byref(desc)), dev)
        fmean = torch.empty((N, P), dtype=torch.float64, device=dev)
        fvar = torch.empty_like(fmean)
        fsample = torch.empty_like(fmean)
        check(lib.gpfx_layer_cond_forward(C.byref(desc), _ptr(X), _ptr(Z), _ptr(ls), _ptr(var), _ptr(q_mu),
                                          _ptr(fctx), _ptr(mean_add), _ptr(eps), _ptr(fmean), _ptr(fvar),
                                          _ptr(fsample), _ptr(cctx), _ptr(ws), _stream()))
        ctx.desc = desc
        ctx.cctx = cctx
        ctx.fctx = fctx
        ctx.link_shape = tuple(link.shape)
        ctx.has_mean_add = mean_add is not None
        ctx.has_eps = eps is not None
        ctx.save_for_backward(X, Z, ls, var, q_mu, eps if eps is not None else X.new_empty(0), fvar)
        return fsample, fmean, fvar

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_f, g_mean, g_var):
        lib = _cabi.load()
        if ctx.cctx is None:
            raise RuntimeError("gpflux_b200 SVGP layer: backward called twice (ctx is consumed)")
        X, Z, ls, var, q_mu, eps, fvar = ctx.saved_tensors
        eps = eps if ctx.has_eps else None
        desc = ctx.desc
        dev = X.device
        N, M, D, P, K = desc.N, desc.M, desc.D, desc.P, desc.K
        g_f = _chk(g_f, "g_fsample")
        g_mean = _chk(g_mean, "g_fmean")
        g_var = _chk(g_var, "g_fvar")
        dX = torch.empty((N, D), dtype=torch.float64, device=dev)
        dZ = torch.empty((K, M, D), dtype=torch.float64, device=dev)
        dls = torch.empty((K, D), dtype=torch.float64, device=dev)
        dvar = torch.empty((K,), dtype=torch.float64, device=dev)
        dq_mu = torch.empty((M, P), dtype=torch.float64, device=dev)
        dq_sqrt = torch.empty((P, M, M), dtype=torch.float64, device=dev)
        dL = torch.empty(ctx.link_shape, dtype=torch.float64, device=dev)
        dmean = torch.empty((N, P), dtype=torch.float64, device=dev)
        ws = _workspace(lib.gpfx_layer_cond_workspace_bytes(C.byref(desc)), dev)
        check(lib.gpfx_layer_cond_backward(C.byref(desc), _ptr(X), _ptr(Z), _ptr(ls), _ptr(var), _ptr(q_mu),
                                           _ptr(eps), _ptr(fvar), _ptr(g_f), _ptr(g_mean), _ptr(g_var), _ptr(dX),
                                           _ptr(dZ), _ptr(dls), _ptr(dvar), _ptr(dq_mu), _ptr(dq_sqrt), _ptr(dL),
                                           _ptr(dmean), _ptr(ctx.fctx), _ptr(ctx.cctx), _ptr(ws), _stream()))
        ctx.cctx = None
        dm = dmean if ctx.has_mean_add else None
        if desc.whiten:
            return (dX, dZ, dls, dvar, dq_mu, dq_sqrt, dL, None, None, None, dm, None, None, None, None, None)
        # non-whitened: dq_mu / dq_sqrt are cotangents of (alpha, C) and go back through the factor node
        return (dX, dZ, dls, dvar, None, None, dL, dq_mu, dq_sqrt, None, dm, None, None, None, None, None)


@_fp32_optin(keep_outputs=True)
def svgp_factor(Z, lengthscales, variance, q_mu, q_sqrt, *, kernel="se", whiten=True, jitter=DEFAULT_JITTER):
    """Parameter-only part of a GPLayer on the current stream -> (fctx, link, kl, link_a, link_c)."""
    return _FactorFn.apply(Z, lengthscales, variance, q_mu, q_sqrt, kernel, whiten, jitter)


@_fp32_optin()
def svgp_cond(X, Z, lengthscales, variance, q_mu, q_sqrt, factor, *, mean_add=None, eps=None, kernel="se",
              whiten=True, jitter=DEFAULT_JITTER, solver="gemm"):
    """Data part of a GPLayer given ``factor = svgp_factor(...)`` -> (fsample, fmean, fvar)."""
    fctx, link, link_a, link_c = factor[0], factor[1], factor[3], factor[4]
    return _CondFn.apply(X, Z, lengthscales, variance, q_mu, q_sqrt, link, link_a, link_c, fctx, mean_add, eps,
                         kernel, whiten, jitter, solver)


# --------------------------------------------------------------------------------------------
# LatentVariableLayer arithmetic (autograd)
# --------------------------------------------------------------------------------------------
class _LatentFn(torch.autograd.Function):
    """(X, mean, std, prior_mean, prior_std, eps) -> (concat(X, mean + std*eps), sum KL[q||p])."""

    @staticmethod
    def forward(ctx, X, mean, std, prior_mean, prior_std, eps):
        lib = _cabi.load()
        X = _chk(X, "X")
        N, D = X.shape
        eps = _chk(eps, "eps")
        W = eps.shape[1]
        eps = _chk(eps, "eps", (N, W))
        per_row = mean.dim() == 2
        mean = _chk(mean, "mean", (N, W) if per_row else (W,))
        std = _chk(std, "std", tuple(mean.shape))
        pm = _chk(prior_mean, "prior_mean", (W,))
        ps = _chk(prior_std, "prior_std", (W,))
        out = torch.empty((N, D + W), dtype=torch.float64, device=X.device)
        kl = torch.empty((), dtype=torch.float64, device=X.device)
        scratch = _workspace(lib.gpfx_latent_scratch_bytes(N, W), X.device)
        check(lib.gpfx_latent_forward(N, D, W, _ptr(X), _ptr(mean), _ptr(std), int(per_row), _ptr(pm), _ptr(ps),
                                      _ptr(eps), _ptr(out), _ptr(kl), _ptr(scratch), _stream()))
        ctx.per_row = per_row
        ctx.dims = (N, D, W)
        ctx.save_for_backward(mean, std, pm, ps, eps)
        return out, kl

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out, g_kl):
        lib = _cabi.load()
        mean, std, pm, ps, eps = ctx.saved_tensors
        N, D, W = ctx.dims
        dev = eps.device
        g_out = _chk(g_out if g_out is not None else torch.zeros((N, D + W), dtype=torch.float64, device=dev), "g_out")
        g_kl = None if g_kl is None else _chk(g_kl, "g_kl")
        dX = torch.empty((N, D), dtype=torch.float64, device=dev)
        if not ctx.per_row:
            return g_out[:, :D].contiguous(), None, None, None, None, None
        dmean = torch.empty((N, W), dtype=torch.float64, device=dev)
        dstd = torch.empty((N, W), dtype=torch.float64, device=dev)
        check(lib.gpfx_latent_backward(N, D, W, _ptr(mean), _ptr(std), _ptr(pm), _ptr(ps), _ptr(eps), _ptr(g_out),
                                       _ptr(g_kl), _ptr(dX), _ptr(dmean), _ptr(dstd), _stream()))
        return dX, dmean, dstd, None, None, None


@_fp32_optin()
def latent_sample_and_kl(X, mean, std, prior_mean, prior_std, eps):
    """LatentVariableLayer arithmetic (reference gpflux/layers/latent_variable_layer.py:140-228):
    -> (concat(X, mean + std * eps) [N, D+W], sum_n KL[N(mean, std^2) || prior])."""
    return _LatentFn.apply(X, mean, std, prior_mean, prior_std, eps)


# --------------------------------------------------------------------------------------------
# Gaussian likelihood (autograd)
# --------------------------------------------------------------------------------------------
class _GaussianVEFn(torch.autograd.Function):
    """(Fmu, Fvar, Y, noise_variance) -> sum of variational expectations (scalar)."""

    @staticmethod
    def forward(ctx, Fmu, Fvar, Y, noise_variance):
        lib = _cabi.load()
        Fmu = _chk(Fmu, "Fmu")
        Fvar = _chk(Fvar, "Fvar", Fmu.shape)
        Y = _chk(Y, "Y", Fmu.shape)
        nv = _chk(noise_variance.reshape(()), "noise_variance")
        n = Fmu.numel()
        ve = torch.empty((), dtype=torch.float64, device=Fmu.device)
        scratch = _workspace(lib.gpfx_gaussian_ve_scratch_bytes(n), Fmu.device)
        check(lib.gpfx_gaussian_ve(_ptr(Fmu), _ptr(Fvar), _ptr(Y), _ptr(nv), n, _ptr(ve), None, None, None,
                                   None, _ptr(scratch), _stream()))
        ctx.save_for_backward(Fmu, Fvar, Y, nv)
        ctx.nv_shape = noise_variance.shape
        return ve

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _cabi.load()
        Fmu, Fvar, Y, nv = ctx.saved_tensors
        n = Fmu.numel()
        g = _chk(g, "g")
        dFmu = torch.empty_like(Fmu)
        dFvar = torch.empty_like(Fvar)
        dnoise = torch.empty((), dtype=torch.float64, device=Fmu.device)
        ve = torch.empty((), dtype=torch.float64, device=Fmu.device)
        scratch = _workspace(lib.gpfx_gaussian_ve_scratch_bytes(n), Fmu.device)
        check(lib.gpfx_gaussian_ve(_ptr(Fmu), _ptr(Fvar), _ptr(Y), _ptr(nv), n, _ptr(ve), _ptr(g), _ptr(dFmu),
                                   _ptr(dFvar), _ptr(dnoise), _ptr(scratch), _stream()))
        return dFmu, dFvar, None, dnoise.reshape(ctx.nv_shape)


@_fp32_optin()
def gaussian_variational_expectations_sum(Fmu, Fvar, Y, noise_variance):
    """sum_{n,q} E_q[log N(y | f, noise)] (reference gpflux/layers/likelihood_layer.py:84-90)."""
    return _GaussianVEFn.apply(Fmu, Fvar, Y, noise_variance)


# --------------------------------------------------------------------------------------------
# random Fourier features / Wilson sampling
# --------------------------------------------------------------------------------------------
def _rff_forward(X, W, bias, ls, var, concat):
    lib = _cabi.load()
    K, L, D = W.shape
    N = X.shape[0]
    out = torch.empty((K, N, 2 * L if concat else L), dtype=torch.float64, device=X.device)
    check(lib.gpfx_rff_features(K, N, D, L, int(concat), _ptr(X), _ptr(W), _ptr(bias), _ptr(ls), _ptr(var),
                                _ptr(out), _stream()))
    return out


class _RffFn(torch.autograd.Function):
    """Feature map with the hand-written backward (gpfx_rff_backward): gradients w.r.t. X, the
    lengthscales and the variance; W and the bias are constants, as in the reference."""

    @staticmethod
    def forward(ctx, X, W, bias, ls, var, concat):
        ctx.save_for_backward(X, W, bias if bias is not None else X.new_zeros(0), ls, var)
        ctx.concat = concat
        return _rff_forward(X, W, bias, ls, var, concat)

    @staticmethod
    def backward(ctx, g):
        lib = _cabi.load()
        X, W, bias, ls, var = ctx.saved_tensors
        K, L, D = W.shape
        N = X.shape[0]
        g = g.contiguous()
        dX, dls, dvar = torch.empty_like(X), torch.empty_like(ls), torch.empty_like(var)
        scratch = torch.empty(lib.gpfx_rff_backward_scratch_bytes(K, N, D, L) // 8, dtype=torch.float64, device=X.device)
        check(lib.gpfx_rff_backward(K, N, D, L, int(ctx.concat), _ptr(X), _ptr(W), _ptr(None if ctx.concat else bias),
                                    _ptr(ls), _ptr(var), _ptr(g), _ptr(dX), _ptr(dls), _ptr(dvar), _ptr(scratch),
                                    _stream()))
        return dX, None, None, dls, dvar, None


@_fp32_optin()
def rff_features(X, W, bias, lengthscales, variance, concat: bool = False):
    """FourierFeaturesBase.call (reference fourier_features/base.py:59-76).  X [N,D]; W [K,L,D];
    bias [K,L] (cosine variant) or None; lengthscales [K,D]; variance [K] -> [K,N,L] or [K,N,2L].
    Differentiable w.r.t. X, lengthscales and variance."""
    X = _chk(X, "X")
    W = _chk(W, "W")
    K, L, D = W.shape
    if X.shape[1] != D:
        raise ValueError(f"X has input dim {X.shape[1]}, W has {D}")
    ls = _chk(lengthscales, "lengthscales", (K, D))
    var = _chk(variance, "variance", (K,))
    bias = None if concat else _chk(bias, "bias", (K, L))
    if torch.is_grad_enabled() and (X.requires_grad or ls.requires_grad or var.requires_grad):
        return _RffFn.apply(X, W.detach(), None if bias is None else bias.detach(), ls, var, bool(concat))
    return _rff_forward(X, W, bias, ls, var, concat)


@_fp32_optin(keep_outputs=True)
def wilson_setup(kernel: str, Z, lengthscales, variance, q_mu, q_sqrt, W, b, coeffs, eps_w, eps_u, *, whiten: bool,
                 jitter: float):
    """Set-up of S Matheron-rule draws in one C-ABI call (reference sample.py:158-181): returns the prior
    weights w [S,L,P] and the function-space update weights v [S,M,P] that ``wilson_eval`` consumes.
    Z [M,D]; lengthscales [D]; variance [1]; q_mu [M,P]; q_sqrt [P,M,M]; W [L,D]; b [L]; coeffs [L];
    eps_w [S,L,P]; eps_u [S,P,M]."""
    lib = _cabi.load()
    Z = _chk(Z, "Z")
    M, D = Z.shape
    eps_w = _chk(eps_w, "eps_w")
    S, L, P = eps_w.shape
    ls = _chk(lengthscales, "lengthscales", (D,))
    var = _chk(variance, "variance", (1,))
    q_mu = _chk(q_mu, "q_mu", (M, P))
    q_sqrt = _chk(q_sqrt, "q_sqrt", (P, M, M))
    W = _chk(W, "W", (L, D))
    b = _chk(b, "b", (L,))
    coeffs = _chk(coeffs, "coeffs", (L,))
    eps_u = _chk(eps_u, "eps_u", (S, P, M))
    w = torch.empty((S, L, P), dtype=torch.float64, device=Z.device)
    v = torch.empty((S, M, P), dtype=torch.float64, device=Z.device)
    info = torch.zeros((2,), dtype=torch.int32, device=Z.device)
    scratch = torch.empty(lib.gpfx_wilson_setup_scratch_bytes(S, M, D, P, L) // 8, dtype=torch.float64, device=Z.device)
    check(lib.gpfx_wilson_setup(KERNEL_IDS[kernel], int(whiten), S, M, D, P, L, float(jitter), _ptr(Z), _ptr(ls), _ptr(var),
                                _ptr(q_mu), _ptr(q_sqrt), _ptr(W), _ptr(b), _ptr(coeffs), _ptr(eps_w), _ptr(eps_u),
                                _ptr(w), _ptr(v), _ptr(info), _ptr(scratch), _stream()))
    return w, v, info


@_fp32_optin()
def wilson_eval(kernel: str, X, W, b, lengthscales, variance, Z, w, v, mean_add=None):
    """Fused WilsonSample.__call__ (reference gpflux/sampling/sample.py:184-198) for S draws.
    X [N,D] (shared by all draws) or [S,N,D]; W [L,D]; b [L]; lengthscales [D]; variance [1];
    Z [M,D]; w [S,L,P]; v [S,M,P]; mean_add [S,N,P] or None -> [S,N,P]."""
    lib = _cabi.load()
    X = _chk(X, "X")
    w = _chk(w, "w")
    S, L, P = w.shape
    x_shared = X.dim() == 2
    N, D = X.shape[-2], X.shape[-1]
    if not x_shared and X.shape[0] != S:
        raise ValueError("X must be [N,D] or [S,N,D]")
    W = _chk(W, "W", (L, D))
    b = _chk(b.reshape(-1), "b", (L,))
    Z = _chk(Z, "Z")
    M = Z.shape[0]
    Z = _chk(Z, "Z", (M, D))
    v = _chk(v, "v", (S, M, P))
    ls = _chk(lengthscales.reshape(-1), "lengthscales", (D,))
    var = _chk(variance.reshape(-1), "variance", (1,))
    mean_add = _chk(mean_add, "mean_add", (S, N, P))
    out = torch.empty((S, N, P), dtype=torch.float64, device=X.device)
    check(lib.gpfx_wilson_eval(KERNEL_IDS[kernel], S, N, D, P, L, M, _ptr(X), int(x_shared), _ptr(W), _ptr(b),
                               _ptr(ls), _ptr(var), _ptr(Z), _ptr(w), _ptr(v), _ptr(mean_add), _ptr(out), _stream()))
    return out
