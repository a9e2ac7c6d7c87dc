"""The library baseline SURVEY §8(d) asks for beside the CPU one: the same c2 training step written
with stock torch-CUDA float64 operators (cuBLAS matmul, cuSOLVER / MAGMA cholesky and
triangular solve, elementwise kernels, autograd for the backward pass), un-fused - what a user gets
from a straight GPU translation of the reference's TensorFlow graph.  It shares no code with the
product path or with oracle/ (written here from the equations of SURVEY §2a) and is measurement
infrastructure only.

Workload = bench.py's: 2 layers (P = 8 then P = 1), M = 256, D = 8, 8192-row minibatch, SE kernel,
whitened q(u), Identity / Zero mean, Gaussian likelihood; forward + backward to every parameter.
Timed with CUDA events after warm-up, eager and as a replayed CUDA graph (launch overhead removed).
Usage: python tools/bench_torch_cuda_baseline.py [--steps 50]  ->  gpurun_out/torch_cuda_baseline.json"""
import argparse
import json
import math
import os

import numpy as np
import torch


def se_kernel(A, B, ls, var):
    As, Bs = A / ls, B / ls
    d2 = (As * As).sum(-1)[:, None] + (Bs * Bs).sum(-1)[None, :] - 2.0 * As @ Bs.T
    return var * torch.exp(-0.5 * d2.clamp_min(0.0))


def layer_forward(X, p, eps, identity_mean):
    M = p["Z"].shape[0]
    Kuu = se_kernel(p["Z"], p["Z"], p["ls"], p["var"]) + 1e-6 * torch.eye(M, dtype=X.dtype, device=X.device)
    L = torch.linalg.cholesky(Kuu)
    Kuf = se_kernel(p["Z"], X, p["ls"], p["var"])  # [M, N]
    A = torch.linalg.solve_triangular(L, Kuf, upper=False)
    mean = A.T @ p["q_mu"]
    if identity_mean:
        mean = mean + X
    Lq = torch.tril(p["q_sqrt"])  # [P, M, M]
    B = Lq.transpose(-1, -2) @ A  # [P, M, N]
    fvar = p["var"] - (A * A).sum(0)[:, None] + (B * B).sum(1).T
    P = p["q_mu"].shape[1]
    diag = torch.diagonal(Lq, dim1=-2, dim2=-1)
    kl = 0.5 * ((p["q_mu"] ** 2).sum() - M * P - torch.log(diag * diag).sum() + (Lq * Lq).sum())
    f = mean if eps is None else mean + torch.sqrt(fvar) * eps
    return f, mean, fvar, kl


def loss_fn(params, noise, X, Y, eps, num_data):
    f, mean, fvar, kl0 = layer_forward(X, params[0], eps, True)
    _, mean, fvar, kl1 = layer_forward(f, params[1], None, False)
    ve = -0.5 * math.log(2.0 * math.pi) - 0.5 * torch.log(noise) - 0.5 * ((Y - mean) ** 2 + fvar) / noise
    return -ve.mean() + (kl0 + kl1) / num_data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    N, M, D = 8192, 256, 8
    X = torch.from_numpy(rng.standard_normal((N, D))).to(dev)
    Y = torch.sin(X.sum(1, keepdim=True)) + 0.1 * torch.from_numpy(rng.standard_normal((N, 1))).to(dev)
    Z0 = rng.standard_normal((M, D))

    def make(P, qs):
        p = {"Z": torch.from_numpy(Z0.copy()), "ls": torch.full((D,), math.sqrt(D)), "var": torch.tensor(1.0),
             "q_mu": torch.from_numpy(0.1 * rng.standard_normal((M, P))),
             "q_sqrt": torch.from_numpy(np.stack([qs * np.eye(M) for _ in range(P)]))}
        return {k: v.double().to(dev).requires_grad_(True) for k, v in p.items()}

    params = [make(D, 1e-5), make(1, 1.0)]
    noise = torch.tensor(1e-2, dtype=torch.float64, device=dev, requires_grad=True)
    leaves = [v for p in params for v in p.values()] + [noise]
    eps = torch.randn(N, D, dtype=torch.float64, device=dev)

    def step():
        for v in leaves:
            v.grad = None
        loss = loss_fn(params, noise, X, Y, eps, 1_000_000)
        loss.backward()
        return loss

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(np.min(ts))

    loss = step()
    assert torch.isfinite(loss), loss
    eager_med, eager_best = timed(step, args.steps, args.warmup)
    out = {"workload": "c2 step (2 layers, M=256, D=8, P=8/1, 8192 rows) with stock torch-CUDA float64 ops + autograd, un-fused",
           "torch": torch.__version__, "device": torch.cuda.get_device_name(0), "loss": float(loss.detach()),
           "eager_ms_median": eager_med, "eager_ms_best": eager_best, "eager_rows_per_s": N / eager_med * 1e3}
    # the same step replayed as one CUDA graph (removes the eager launch overhead); linalg ops that
    # synchronise inside make capture fail on some builds - reported, not fatal
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        for v in leaves:
            v.grad = None
        with torch.cuda.graph(g):
            step()
        graph_med, graph_best = timed(g.replay, args.steps, args.warmup)
        out.update(graph_ms_median=graph_med, graph_ms_best=graph_best, graph_rows_per_s=N / graph_med * 1e3)
    except Exception as e:  # noqa: BLE001
        out["graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/torch_cuda_baseline.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
