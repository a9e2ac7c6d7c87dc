"""ctypes binding of ``include/gpfx.h`` (libgpflux_b200.so).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libgpflux_b200.so"

KERNEL_IDS = {"se": 0, "matern12": 1, "matern32": 2, "matern52": 3}
SOLVER_IDS = {"gemm": 0, "trsm": 1}

TRI_LOWER_A, TRI_UPPER_A, TRI_LOWER_B, TRI_UPPER_B = 1, 2, 4, 8

E_NOT_PD = -4


class GpfxError(RuntimeError):
    pass


class LayerDesc(C.Structure):
    _fields_ = [
        ("N", C.c_int32),
        ("M", C.c_int32),
        ("D", C.c_int32),
        ("P", C.c_int32),
        ("K", C.c_int32),
        ("kernel", C.c_int32),
        ("whiten", C.c_int32),
        ("solver", C.c_int32),
        ("jitter", C.c_double),
    ]


_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_ull = C.c_ulonglong
_d = C.c_double
_sz = C.c_size_t
_dp = C.POINTER(LayerDesc)

# name -> (restype, argtypes); must list every symbol declared in include/gpfx.h
SIGNATURES = {
    "gpfx_version": (_i, []),
    "gpfx_last_error": (C.c_char_p, []),
    "gpfx_create": (_i, [_i, C.POINTER(C.c_void_p)]),
    "gpfx_destroy": (_i, [_p]),
    "gpfx_set_current": (_i, [_p]),
    "gpfx_get_current": (C.c_void_p, []),
    "gpfx_handle_last_error": (C.c_char_p, [_p]),
    "gpfx_kernel_launch_count": (_ll, []),
    "gpfx_tune_gemm_config": (None, [_i]),
    "gpfx_layer_ctx_bytes": (_sz, [_dp]),
    "gpfx_layer_workspace_bytes": (_sz, [_dp]),
    "gpfx_layer_forward": (_i, [_dp] + [_p] * 15),
    "gpfx_layer_backward": (_i, [_dp] + [_p] * 21),
    "gpfx_layer_status": (_i, [_dp, _p, _p, _p]),
    "gpfx_layer_status_async": (_i, [_dp, _p, _p, _p]),
    "gpfx_layer_factor_ctx_bytes": (_sz, [_dp]),
    "gpfx_layer_cond_ctx_bytes": (_sz, [_dp]),
    "gpfx_layer_factor_workspace_bytes": (_sz, [_dp]),
    "gpfx_layer_cond_workspace_bytes": (_sz, [_dp]),
    "gpfx_layer_dL_bytes": (_sz, [_dp]),
    "gpfx_layer_factor_forward": (_i, [_dp] + [_p] * 9),
    "gpfx_layer_factor_backward": (_i, [_dp] + [_p] * 11 + [_i] + [_p] * 3),
    "gpfx_layer_cond_forward": (_i, [_dp] + [_p] * 14),
    "gpfx_layer_cond_backward": (_i, [_dp] + [_p] * 22),
    "gpfx_kernel_matrix": (_i, [_i, _i, _i, _i, _i, _p, _i, _p, _i, _p, _p, _i, _d, _p, _p]),
    "gpfx_kernel_matrix_bwd_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "gpfx_kernel_matrix_bwd": (_i, [_i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_cholesky_workspace_bytes": (_sz, [_i, _i]),
    "gpfx_cholesky_inverse": (_i, [_i, _i, _p, _p, _p, _p, _p, _p]),
    "gpfx_cholesky_backward_workspace_bytes": (_sz, [_i, _i]),
    "gpfx_cholesky_backward": (_i, [_i, _i, _p, _p, _p, _p, _p, _p]),
    "gpfx_gemm": (_i, [_i, _i, _i, _i, _d, _p, _i, _ll, _i, _p, _i, _ll, _i, _d, _p, _i, _ll, _i, _i, _p]),
    "gpfx_gemm_splitk": (_i, [_i, _i, _i, _i, _d, _p, _i, _ll, _i, _p, _i, _ll, _i, _d, _p, _i, _ll, _i, _i, _i, _p, _p]),
    "gpfx_gemm_kscaled": (_i, [_i, _i, _i, _i, _d, _p, _i, _ll, _p, _i, _ll, _p, _ll, _p, _i, _ll, _i, _i, _p, _p]),
    "gpfx_set_contraction": (_i, [_i, _p, _sz]),
    "gpfx_get_contraction": (_i, []),
    "gpfx_layer_tf32_workspace_bytes": (_sz, [_dp]),
    "gpfx_gemm_tf32x3_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "gpfx_gemm_tf32x3": (_i, [_i, _i, _i, _i, _d, _p, _i, _ll, _i, _p, _i, _ll, _i, _d, _p, _i, _ll, _i, _i, _p, _p]),
    "gpfx_reparam_sample": (_i, [_p, _p, _p, _p, _ll, _p]),
    "gpfx_philox_raw": (_i, [_ull, _ull, _p, _ll, _p]),
    "gpfx_randn_philox": (_i, [_ull, _ull, _p, _ll, _p]),
    "gpfx_reparam_sample_philox": (_i, [_p, _p, _ull, _ull, _p, _p, _ll, _p]),
    "gpfx_kl_scratch_bytes": (_sz, [_i, _i]),
    "gpfx_kl_white": (_i, [_i, _i, _p, _p, _p, _p, _p]),
    "gpfx_kl_white_tril": (_i, [_i, _i, _p, _p, _p, _p, _p, _p]),
    "gpfx_kl_white_bwd": (_i, [_i, _i, _p, _p, _p, _p, _p, _p]),
    "gpfx_gaussian_ve_scratch_bytes": (_sz, [_ll]),
    "gpfx_gaussian_ve": (_i, [_p, _p, _p, _p, _ll, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_latent_scratch_bytes": (_sz, [_ll, _i]),
    "gpfx_latent_forward": (_i, [_ll, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_latent_backward": (_i, [_ll, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_rff_features": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_rff_backward_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "gpfx_rff_backward": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_wilson_setup_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "gpfx_wilson_setup": (_i, [_i, _i, _i, _i, _i, _i, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_natgrad_workspace_bytes": (_sz, [_i, _i]),
    "gpfx_natgrad_step": (_i, [_i, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_wilson_eval": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gpfx_comm_unique_id_bytes": (_i, []),
    "gpfx_comm_nccl_version": (_i, []),
    "gpfx_comm_unique_id": (_i, [_p]),
    "gpfx_comm_init": (_i, [_p, _i, _i, C.POINTER(_p)]),
    "gpfx_comm_destroy": (_i, [_p]),
    "gpfx_allreduce_grads": (_i, [_p, _p, _ll, _i, _p]),
    "gpfx_gemm_profile_enable": (None, [_i]),
    "gpfx_gemm_profile_read": (_i, [_p, _i]),
}



class GemmSample(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("M", "N", "K", "batch", "nseg", "tri", "transA", "transB", "tril_out",
                                         "splits", "cfg", "reserved")] + [("ms", C.c_double)]


_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the C-ABI library (once).  Raises GpfxError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise GpfxError(
            f"{_LIB_PATH} is missing: build it with `python -m gpflux_b200.build` "
            "(or __graft_entry__.build()).  gpflux_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().gpfx_last_error().decode("utf-8", "replace")
        raise GpfxError(f"gpfx call failed with status {status}: {msg}")
