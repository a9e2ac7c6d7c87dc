"""CPU tests of the host logic: the C-ABI library loads and exports every symbol include/gpfx.h
declares (no compute calls), builder/compat semantics mirror the reference's, the product path
refuses to run without CUDA, and the data-parallel plumbing works over gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from gpflux_b200 import _cabi
    from gpflux_b200.build import build_library

    build_library()
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "gpfx.h")).read()
    declared = set(re.findall(r"\b(gpfx_[A-Za-z0-9_]+)\s*\(", header))
    assert "gpfx_layer_forward" in declared and "gpfx_layer_backward" in declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in gpfx.h but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_cabi.SIGNATURES) == declared
    assert lib.gpfx_version() == 100


def test_size_queries_do_not_need_a_gpu():
    import ctypes as C

    from gpflux_b200 import _cabi

    lib = _cabi.load()
    d = _cabi.LayerDesc(8192, 256, 8, 8, 1, 0, 1, 0, 1e-6)
    ctx = lib.gpfx_layer_ctx_bytes(C.byref(d))
    ws = lib.gpfx_layer_workspace_bytes(C.byref(d))
    # ctx holds Lq [P,M,M], L/Linv [K,Mp,Mp], Kuf/A [K,M,N], B [P,M,N]
    assert ctx >= 8 * (8 * 256 * 256 + 2 * 256 * 256 + 2 * 256 * 8192 + 8 * 256 * 8192)
    assert ws > 0
    bad = _cabi.LayerDesc(8192, 256, 8, 8, 3, 0, 1, 0, 1e-6)  # K must be 1 or P
    assert lib.gpfx_layer_ctx_bytes(C.byref(bad)) == 0
    assert b"K must be 1 or P" in lib.gpfx_last_error()
    white = _cabi.LayerDesc(100, 20, 2, 2, 1, 0, 1, 0, 1e-6)
    nonwhite = _cabi.LayerDesc(100, 20, 2, 2, 1, 0, 0, 0, 1e-6)
    # the non-whitened layer keeps alpha = L^-1 q_mu [M,P] next to the factor
    assert lib.gpfx_layer_factor_ctx_bytes(C.byref(nonwhite)) > lib.gpfx_layer_factor_ctx_bytes(C.byref(white))
    badk = _cabi.LayerDesc(100, 20, 2, 2, 1, 9, 1, 0, 1e-6)
    assert lib.gpfx_layer_ctx_bytes(C.byref(badk)) == 0


def test_product_path_has_no_cpu_fallback():
    from gpflux_b200 import ops

    x = torch.zeros(4, 2, dtype=torch.float64)
    with pytest.raises(ValueError, match="CUDA"):
        ops.svgp_layer(x, x[None], x[:1], x[0, :1], x, x[None])
    with pytest.raises(ValueError, match="CUDA"):
        ops.kl_white(x, x[None])


def test_oracle_is_not_imported_by_the_product():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gpflux_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_helpers_mirror_reference_semantics():
    """reference tests/gpflux/test_helpers.py:47-180, test_runtime_checks.py:51-82."""
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200 import helpers
    from gpflux_b200.exceptions import GPLayerIncompatibilityException
    from gpflux_b200.layers import GPLayer
    from gpflux_b200.runtime_checks import verify_compatibility

    k = helpers.construct_basic_kernel(gf.SquaredExponential(), output_dim=3, share_hyperparams=True)
    assert isinstance(k, gf.SharedIndependent) and k.num_latent_gps == 3
    k = helpers.construct_basic_kernel(gf.SquaredExponential(), output_dim=3, share_hyperparams=False)
    assert isinstance(k, gf.SeparateIndependent) and len(k.kernels) == 3
    assert k.kernels[0] is not k.kernels[1]
    k2 = helpers.construct_basic_kernel([gf.SquaredExponential(), gf.Matern32()])
    assert isinstance(k2, gf.SeparateIndependent) and k2.num_latent_gps == 2

    with pytest.warns(UserWarning):
        iv = helpers.construct_basic_inducing_variables(7, 2, output_dim=3, share_variables=True)
    assert isinstance(iv, gf.SharedIndependentInducingVariables) and iv.num_inducing == 7
    z = np.zeros((3, 7, 2))
    iv = helpers.construct_basic_inducing_variables(7, 2, output_dim=3, share_variables=False, z_init=z)
    assert isinstance(iv, gf.SeparateIndependentInducingVariables) and len(iv.inducing_variable_list) == 3
    with pytest.raises(ValueError):
        helpers.construct_basic_inducing_variables(7, 2, output_dim=3, share_variables=False, z_init=np.zeros((7, 2)))
    iv = helpers.construct_basic_inducing_variables([5, 5], 2, z_init=[np.zeros((5, 2))] * 2)
    assert len(iv.inducing_variable_list) == 2

    X = np.random.default_rng(0).standard_normal((30, 4))
    assert isinstance(helpers.construct_mean_function(X, 4, 4), gf.Identity)
    lin = helpers.construct_mean_function(X, 4, 2)
    assert isinstance(lin, gf.Linear) and tuple(lin.A.shape) == (4, 2) and not lin.A.trainable
    pad = helpers.construct_mean_function(X, 4, 6)
    np.testing.assert_array_equal(pad.A.numpy(), np.concatenate([np.eye(4), np.zeros((4, 2))], 1))

    layer = helpers.construct_gp_layer(100, 11, 4, 2, z_init=np.zeros((11, 4)))
    assert tuple(layer.q_mu.shape) == (11, 2) and tuple(layer.q_sqrt.shape) == (2, 11, 11)
    np.testing.assert_array_equal(layer.q_mu.numpy(), 0.0)
    np.testing.assert_array_equal(layer.q_sqrt.numpy()[1], np.eye(11))
    np.testing.assert_allclose(layer.kernel.kernel.lengthscales.numpy(), 2.0)

    # verify_compatibility
    kern = gf.SharedIndependent(gf.SquaredExponential(), 2)
    ivs = gf.SharedIndependentInducingVariables(gf.InducingPoints(np.zeros((5, 3))))
    assert verify_compatibility(kern, gf.Zero(), ivs) == (5, 2)
    with pytest.raises(GPLayerIncompatibilityException):
        verify_compatibility(gf.SquaredExponential(), gf.Zero(), ivs)
    with pytest.raises(GPLayerIncompatibilityException):
        verify_compatibility(kern, gf.Zero(), gf.InducingPoints(np.zeros((5, 3))))
    sep = gf.SeparateIndependentInducingVariables([gf.InducingPoints(np.zeros((5, 3)))] * 3)
    with pytest.raises(GPLayerIncompatibilityException):
        verify_compatibility(kern, gf.Zero(), sep)
    # plain kernel + InducingPoints is accepted when num_latent_gps is given (gp_layer.py:193-209)
    with pytest.warns(UserWarning):
        lay = GPLayer(gf.SquaredExponential(), gf.InducingPoints(np.zeros((5, 3))), 10, gf.Zero(), num_latent_gps=4)
    assert tuple(lay.q_mu.shape) == (5, 4)
    with pytest.raises(GPLayerIncompatibilityException):
        GPLayer(gf.SquaredExponential(), gf.InducingPoints(np.zeros((5, 3))), 10, gf.Zero())


def test_parameter_transforms():
    from gpflux_b200.gpflow_compat import Parameter

    p = Parameter([0.5, 2.0, 40.0], transform="positive")
    np.testing.assert_allclose(p.numpy(), [0.5, 2.0, 40.0], rtol=1e-14)
    p.assign([1.0, 1.0, 1.0])
    np.testing.assert_allclose(p.numpy(), 1.0, rtol=1e-14)
    t = Parameter(np.ones((1, 3, 3)), transform="triangular")
    np.testing.assert_array_equal(t.numpy()[0], np.tril(np.ones((3, 3))))
    with pytest.raises(ValueError):
        Parameter([-1.0], transform="positive")


def test_deep_gp_builder_and_num_data_validation():
    from gpflux_b200 import gpflow_compat as gf
    from gpflux_b200.architectures import Config, build_constant_input_dim_deep_gp
    from gpflux_b200.helpers import construct_gp_layer
    from gpflux_b200.models import DeepGP

    X = np.random.default_rng(0).standard_normal((300, 3))
    m = build_constant_input_dim_deep_gp(X, 3, Config(10, 1e-5, 0.01))
    assert len(m.f_layers) == 3 and m.num_data == 300
    assert m.f_layers[0].num_latent_gps == 3 and m.f_layers[-1].num_latent_gps == 1
    np.testing.assert_allclose(m.f_layers[0].q_sqrt.numpy()[0], 1e-5 * np.eye(10))
    np.testing.assert_allclose(m.f_layers[-1].q_sqrt.numpy()[0], np.eye(10))
    np.testing.assert_allclose(m.f_layers[0].kernel.kernel.variance.numpy(), 1e-6, rtol=1e-10)
    np.testing.assert_allclose(m.f_layers[-1].kernel.kernel.variance.numpy(), 1.0, rtol=1e-12)
    a = construct_gp_layer(100, 5, 2, 2, z_init=np.zeros((5, 2)))
    b = construct_gp_layer(200, 5, 2, 1, z_init=np.zeros((5, 2)))
    with pytest.raises(ValueError):
        DeepGP([a, b], gf.Gaussian(0.1))
    assert DeepGP([a], gf.Gaussian(0.1)).num_data == 100


def test_shard_rows():
    from gpflux_b200.distributed import shard_rows

    for n, w in [(65536, 8), (10, 3), (7, 8)]:
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from gpflux_b200.distributed import FlatGradients, allreduce_scalar_mean
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
torch.manual_seed(0)
lin = torch.nn.Linear(3, 2).double()
flat = FlatGradients(lin.parameters())
x = torch.full((4, 3), float(rank + 1), dtype=torch.float64)
flat.zero_()
lin(x).sum().backward()
flat.allreduce_mean()
# weight grad = sum_rows x = 4 * (rank + 1) per entry; mean over ranks = 4 * 1.5 = 6
assert torch.allclose(lin.weight.grad, torch.full((2, 3), 6.0, dtype=torch.float64)), lin.weight.grad
assert torch.allclose(lin.bias.grad, torch.full((2,), 4.0, dtype=torch.float64))
assert lin.weight.grad.data_ptr() == flat.flat.data_ptr()  # still a view: one collective, no copies
m = allreduce_scalar_mean(torch.tensor(float(rank), dtype=torch.float64))
assert float(m) == 0.5
dist.destroy_process_group()
print("ok", rank)
"""


def test_gradient_allreduce_gloo_world2(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out
        assert "ok" in out


def test_tf_shim_sources_reference_declared_entry_points():
    """The TensorFlow custom-op shim cannot be compiled here (no TensorFlow); at least every gpfx_*
    symbol it calls must be declared in include/gpfx.h, and its Python half must import cleanly
    without TensorFlow (the import is deferred to load())."""
    import importlib
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "gpfx.h")).read()
    declared = set(re.findall(r"\b(gpfx_[a-z_0-9]+)\s*\(", header))
    src = open(os.path.join(root, "gpflux_b200", "tf_shim", "gpfx_tf_ops.cc")).read()
    used = set(re.findall(r"\b(gpfx_[a-z_0-9]+)\s*\(", src))
    assert used and used <= declared, used - declared
    mod = importlib.import_module("gpflux_b200.tf_shim.gpfx_tf_ops")
    assert mod.KERNEL_ID["SquaredExponential"] == 0
    from gpflux_b200.tf_shim.build import build_tf_shim

    assert build_tf_shim() is None  # TensorFlow absent -> no-op


def test_flat_gradients_buckets_are_persistent_views():
    """FlatGradients: .grad tensors are views of one persistent buffer, grouped in buckets, and survive
    zero_() / optimizer.zero_grad(set_to_none=True) without being re-pointed (what a captured CUDA graph
    relies on; ADVICE r1 bench.py:285)."""
    import torch

    from gpflux_b200.distributed import FlatGradients, local_loss_weight

    a, b = torch.nn.Linear(3, 2).double(), torch.nn.Linear(2, 1).double()
    flat = FlatGradients([a.parameters(), b.parameters(), list(a.parameters())])  # duplicates dropped
    assert flat.bucket_numels() == [8, 3] and flat.numel() == 11
    ptrs = [p.grad.data_ptr() for p in flat.params]
    for step in range(3):
        flat.zero_()
        x = torch.full((4, 3), float(step + 1), dtype=torch.float64)
        b(a(x)).sum().backward()
        flat.allreduce_mean()  # world of one: no-op
        assert [p.grad.data_ptr() for p in flat.params] == ptrs
        assert torch.equal(flat.flat[:6].view(2, 3), a.weight.grad)
        assert float(a.weight.grad.abs().sum()) > 0
        for p in flat.params:
            p.grad = None  # what optimizer.zero_grad(set_to_none=True) does
    assert local_loss_weight(10, 40, 4) == 1.0 and local_loss_weight(11, 40, 4) == 1.1


def _run_bench(args, env_extra=None, timeout=300):
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                       env=env, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    return [json.loads(ln) for ln in lines]


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU oracle timed on the host cores) emits exactly one JSON line with the
    product arm's metric / unit / config keys, `impl`, a `cpu_baseline` describing the run and a host-only `e2e`."""
    out = _run_bench(["--impl", "reference", "--config", "c2", "--steps", "1", "--warmup", "0"])
    assert len(out) == 1
    d = out[0]
    assert d["impl"] == "reference" and d["metric"] == "deepgp_elbo_fwd_bwd_rows_per_s" and d["unit"] == "rows/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("c2")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the reference arm; the other ranks print nothing and exit 0."""
    out = _run_bench(["--impl", "reference", "--config", "c2", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                     env_extra={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out == []


def test_header_is_plain_c_and_the_c_host_example_links(tmp_path):
    """include/gpfx.h compiles as C99 (no C++ in the boundary) and examples/c_host links against the built
    library; running it needs a GPU (tests/test_gpu_c_host.py)."""
    import os
    import shutil
    import subprocess

    import pytest

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    gcc = shutil.which("gcc")
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    from gpflux_b200 import _cabi

    _cabi.load()
    libdir = os.path.join(root, "gpflux_b200", "lib")
    probe = tmp_path / "probe.c"
    probe.write_text('#include "gpfx.h"\nint main(void) { gpfx_layer_desc d; d.N = 0; (void)d; return gpfx_version() >= 100 ? 0 : 1; }\n')
    exe = str(tmp_path / "probe")
    p = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), str(probe),
                        "-L" + libdir, "-lgpflux_b200", "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert subprocess.run([exe]).returncode == 0  # gpfx_version() needs no device
    p = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), "-I" + os.path.join(cuda, "include"),
                        os.path.join(root, "examples", "c_host", "svgp_layer_step.c"), "-L" + libdir, "-lgpflux_b200",
                        "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-o", str(tmp_path / "svgp_layer_step")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
