"""BASELINE configs 1 and 4 (the two small ones), fwd+bwd step time on one GPU, captured in a CUDA graph
like bench.py:
  c1: single-layer GPLayer SVGP, SquaredExponential, M = 50, the reference's tests/snelson1d.npz (N = 200);
  c4: LatentVariableLayer arithmetic (posterior sample + KL + concat, latent dim 2) + GPLayer(M = 512,
      D = 1 + 2, P = 1) conditional-density model on a minibatch of 8192 of 2M synthetic rows; the
      per-row posterior mean/std arrive as inputs (an amortised encoder is host-side).
Writes gpurun_out/bench_c1_c4.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import gpflow_compat as gf  # noqa: E402
from gpflux_b200 import ops  # noqa: E402
from gpflux_b200.layers import GPLayer, LikelihoodLayer  # noqa: E402
from gpflux_b200.models import DeepGP  # noqa: E402


def time_graph(step, iters=50):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def c1():
    d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "snelson1d.npz"))
    X, Y = torch.from_numpy(d["X"]).cuda(), torch.from_numpy(d["Y"]).cuda()
    M = 50
    kernel = gf.SharedIndependent(gf.SquaredExponential(variance=0.7, lengthscales=0.6), 1)
    iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(np.linspace(0, 6, M)[:, None]))
    layer = GPLayer(kernel, iv, num_data=200, mean_function=gf.Zero())
    model = DeepGP([layer], gf.Gaussian(0.08)).cuda()
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        for p in params:
            p.grad = None
        model.loss(X, Y).backward()

    ms = time_graph(step)
    return {"config": "c1: GPLayer SVGP, SE, M=50, snelson1d N=200 (full batch)", "ms_per_step": ms, "rows_per_s": 200 / ms * 1e3}


def c4():
    rng = np.random.default_rng(0)
    N, W, M = 8192, 2, 512
    X = torch.from_numpy(rng.uniform(0, 1, (N, 1))).cuda()
    Y = torch.from_numpy(np.sin(6 * X.cpu().numpy()) + 0.1 * rng.standard_normal((N, 1))).cuda()
    mean = torch.from_numpy(0.1 * rng.standard_normal((N, W))).cuda().requires_grad_(True)
    std = torch.from_numpy(rng.uniform(0.5, 1.0, (N, W))).cuda().requires_grad_(True)
    pm, ps = torch.zeros(W, dtype=torch.float64, device="cuda"), torch.ones(W, dtype=torch.float64, device="cuda")
    kernel = gf.SharedIndependent(gf.SquaredExponential(variance=1.0, lengthscales=np.ones(1 + W)), 1)
    iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(rng.uniform(-1, 1, (M, 1 + W))))
    layer = GPLayer(kernel, iv, num_data=2_000_000, mean_function=gf.Zero())
    layer.q_mu.assign(0.1 * rng.standard_normal((M, 1)))
    lik = LikelihoodLayer(gf.Gaussian(0.1))
    layer, lik = layer.cuda(), lik.cuda()
    params = [p for p in list(layer.parameters()) + list(lik.parameters()) if p.requires_grad] + [mean, std]

    def step():
        for p in params:
            p.grad = None
        eps = torch.randn((N, W), dtype=torch.float64, device="cuda")
        h, kl_w = ops.latent_sample_and_kl(X, mean, std, pm, ps, eps)
        f = layer(h, training=True)
        lik(f, targets=Y, training=True)
        loss = kl_w / 2_000_000 + layer.losses[0] + lik.losses[0]
        loss.backward()

    ms = time_graph(step)
    flops = 3 * (2 * M * M * 3 + 2 * M * N * 3 + M**3 / 3 + 2 * M * M * N + 4 * M * N + 2 * M * N)
    return {"config": "c4: latent (W=2) + GPLayer(M=512, D=3, P=1) CDE, minibatch 8192 of 2M", "ms_per_step": ms,
            "rows_per_s": N / ms * 1e3, "algorithmic_tflops": flops / ms / 1e9}


if __name__ == "__main__":
    res = [c1(), c4()]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_c1_c4.json", "w"), indent=1)
    for r in res:
        print(json.dumps(r))
