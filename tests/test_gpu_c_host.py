"""The C ABI driven by a host that is neither Python nor torch: examples/c_host/svgp_layer_step.c (plain C99 +
the CUDA runtime API) is compiled here with gcc, run on files of seeded inputs, and its outputs - layer
outputs and every gradient - are checked against the CPU oracle at the same tolerance as the Python-hosted
parity tests.  Reference call site: gpflux/layers/gp_layer.py:226-363 (GPLayer.call) under a GradientTape."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from _util import assert_close, cond_tol
from test_gpu_layer import PARAMS, _run_oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
KERNEL_ID = {"se": 0, "matern12": 1, "matern32": 2, "matern52": 3}


@pytest.fixture(scope="module")
def c_host(tmp_path_factory):
    gcc = shutil.which("gcc")
    if gcc is None or not os.path.exists(os.path.join(CUDA_HOME, "include", "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    from gpflux_b200 import _cabi

    _cabi.load()  # fails loudly if the library is not built
    libdir = os.path.join(ROOT, "gpflux_b200", "lib")
    exe = str(tmp_path_factory.mktemp("c_host") / "svgp_layer_step")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA_HOME, "include"),
           os.path.join(ROOT, "examples", "c_host", "svgp_layer_step.c"), "-L" + libdir, "-lgpflux_b200",
           "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart", "-Wl,-rpath," + libdir,
           "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64"), "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


@pytest.mark.parametrize("N,M,D,P,K,kernel", [
    (256, 64, 4, 3, 3, "se"),        # separate kernels
    (1000, 256, 8, 8, 1, "se"),      # config-2 hidden layer shape (reduced N), shared kernel
    (515, 130, 3, 2, 2, "matern32"),
])
def test_c_host_layer_step_matches_oracle(c_host, tmp_path, N, M, D, P, K, kernel):
    from oracle import svgp_torch as ot

    rng = np.random.default_rng(4000 + N + M)
    spec = ot.make_spec(rng, M, D, P, K=K, kernel=kernel, mean="zero")
    X = torch.from_numpy(rng.standard_normal((N, D)))
    eps, w_f, w_m, w_v = (torch.from_numpy(rng.standard_normal((N, P))) for _ in range(4))
    w_kl = 0.37
    ref = _run_oracle(spec, X, eps, w_f, w_m, w_v, w_kl)

    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([N, M, D, P, K, KERNEL_ID[kernel], 1 if spec.get("whiten", True) else 0, 0], dtype=np.int32).tofile(f)
        arrays = [np.array([spec["jitter"]]), X, spec["Z"], spec["lengthscales"], spec["variance"], spec["q_mu"], spec["q_sqrt"],
                  eps, w_f, w_m, w_v, np.array([w_kl])]
        for a in arrays:
            np.ascontiguousarray(a.numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float64).tofile(f)
    p = subprocess.run([c_host, fin, fout], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr + p.stdout
    assert "kernels launched" in p.stdout

    out = np.fromfile(fout, dtype=np.float64)
    shapes = [("fmean", (N, P)), ("fvar", (N, P)), ("fsample", (N, P)), ("kl", ()), ("X", (N, D)), ("Z", (K, M, D)),
              ("lengthscales", (K, D)), ("variance", (K,)), ("q_mu", (M, P)), ("q_sqrt", (P, M, M)), ("dmean", (N, P))]
    assert out.size == sum(int(np.prod(s)) for _, s in shapes)
    got, off = {}, 0
    for name, s in shapes:
        n = int(np.prod(s))
        got[name] = torch.from_numpy(out[off:off + n].reshape(s).copy())
        off += n

    tol = cond_tol(spec, 1e-9)
    assert_close(got["fmean"], ref[1], scale=1.0, tol=tol, what="fmean")
    assert_close(got["fvar"], ref[2], scale=float(spec["variance"].max()), tol=tol, what="fvar")
    assert_close(got["fsample"], ref[0], scale=1.0, tol=tol, what="fsample")
    assert_close(got["kl"], ref[3], tol=tol, what="kl")
    for k in PARAMS + ("X",):
        assert_close(got[k], ref[4][k], tol=tol, what=f"grad {k}")
    # dmean = dLoss / d fmean: what the host would back-propagate into a mean function
    assert_close(got["dmean"], w_f + w_m, tol=tol, what="dmean")
