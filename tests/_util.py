"""Shared helpers for the parity tests (oracle = checker, CUDA path = thing under test)."""
from __future__ import annotations

import numpy as np
import torch

REL_TOL = 1e-9  # BASELINE.json north_star: 1e-9 relative in float64


def assert_close(got, ref, scale=None, tol=REL_TOL, what=""):
    """|got - ref| <= tol * max(|ref|, scale) elementwise (SURVEY §7 hard part 3: variances are
    differences of O(scale) numbers, so the tolerance is relative to that scale)."""
    got = got.detach().cpu().numpy() if hasattr(got, "detach") else np.asarray(got)
    ref = ref.detach().cpu().numpy() if hasattr(ref, "detach") else np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.all(np.isfinite(got)), f"{what}: non-finite values"
    if scale is None:
        scale = float(np.max(np.abs(ref))) if ref.size else 1.0
    bound = tol * np.maximum(np.abs(ref), scale)
    err = np.abs(got - ref)
    worst = float(np.max(err / np.maximum(bound, 1e-300))) if ref.size else 0.0
    assert worst <= 1.0, f"{what}: max err/bound = {worst:.3g} (max abs err {float(err.max()):.3g}, scale {scale:.3g})"


def cond_tol(spec, base=REL_TOL):
    """Tolerance for one layer case: the north-star 1e-9, widened to cond(Kuu) * eps where Kuu is so
    ill-conditioned that the oracle's own answer moves by more than 1e-9 between mathematically
    equivalent formulations (e.g. 50 inducing points on a 1-D input: cond ~ 4e7, and switching the
    oracle from the expansion form of the squared distance to the direct form changes its
    lengthscale gradient by 4e-10 relative; see DESIGN.md "Tolerance and conditioning")."""
    from oracle import svgp_torch as ot

    cond = max(float(torch.linalg.cond(ot.Kuu(spec, k))) for k in range(spec["Z"].shape[0]))
    return max(base, cond * 2.220446049250313e-16)


def spec_to_cuda(spec, device="cuda"):
    out = {}
    for k, v in spec.items():
        out[k] = v.detach().to(device) if isinstance(v, torch.Tensor) else v
    return out
