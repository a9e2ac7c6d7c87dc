"""The CUDA path against the independent pin (tests/golden/make_independent.py: mpmath 50-digit dense
closed form + finite-difference gradients), without the oracle in between."""
import pytest
import torch

from _independent import PARAMS, case_names, load_cases
from _util import assert_close
from test_gpu_layer import _run_gpu

pytestmark = pytest.mark.gpu
CASES = {c["name"]: c for c in load_cases()}
TOL = 1e-9


@pytest.mark.parametrize("name", case_names())
def test_cuda_layer_matches_independent_closed_form(name):
    c = CASES[name]
    N, P = c["fmean"].shape
    zero = torch.zeros((N, P), dtype=torch.float64)
    f, m, v, kl, grads = _run_gpu(c["spec"], c["X"], None, zero, c["w_m"], c["w_v"], c["w_kl"])
    assert_close(m, c["fmean"], scale=1.0, tol=TOL, what="fmean")
    assert_close(v, c["fvar"], scale=float(c["spec"]["variance"].max()), tol=TOL, what="fvar")
    assert_close(kl, c["kl"], tol=TOL, what="kl")
    for k in PARAMS + ("X",):
        assert_close(grads[k], c["grads"][k], tol=TOL, what=f"grad {k}")
