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


def spec_to_cuda(spec, device="cuda"):
    out = {}
    for k, v in spec.items():
        out[k] = v.detach().to(device) if isinstance(v, torch.Tensor) else v
    return out
