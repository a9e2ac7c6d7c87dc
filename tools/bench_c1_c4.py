"""BASELINE configs 1 and 4 (the two small ones), fwd+bwd step time on one GPU, captured in a CUDA graph
like bench.py:
  c1: single-layer GPLayer SVGP, SquaredExponential, M = 50, the reference's tests/snelson1d.npz (N = 200);
  c4: LatentVariableLayer arithmetic (posterior sample + KL + concat, latent dim 2) + GPLayer(M = 512,
      D = 1 + 2, P = 1) conditional-density model on a minibatch of 8192 of 2M synthetic rows; the
      per-row posterior mean/std arrive as inputs (an amortised encoder is host-side).
One GPU:  python tools/bench_c1_c4.py
N GPUs:   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              tools/bench_c1_c4.py
          (data parallel, weak scaling: every rank steps its own minibatch, the parameter gradients are averaged
          by the library's NCCL all-reduce inside the captured step; time = max over ranks)
Writes gpurun_out/bench_c1_c4[_nN].json."""
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

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
COMM = None


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
    if WORLD > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    b.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device="cuda")
    if WORLD > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    del g
    return float(ms)


def flat_grads(params):
    """Persistent flat gradient buffer; with more than one rank its mean over ranks is taken by
    gpfx_allreduce_grads (NCCL) at the end of every step."""
    from gpflux_b200.distributed import FlatGradients

    return FlatGradients([list(params)], comm=COMM, overlap=False)


def c1():
    d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "snelson1d.npz"))
    X, Y = torch.from_numpy(d["X"]).cuda(), torch.from_numpy(d["Y"]).cuda()
    M = 50
    kernel = gf.SharedIndependent(gf.SquaredExponential(variance=0.7, lengthscales=0.6), 1)
    iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(np.linspace(0, 6, M)[:, None]))
    layer = GPLayer(kernel, iv, num_data=200, mean_function=gf.Zero())
    model = DeepGP([layer], gf.Gaussian(0.08)).cuda()
    flat = flat_grads(p for p in model.parameters() if p.requires_grad)

    def step():
        flat.zero_()
        model.loss(X, Y).backward()
        flat.allreduce_mean()

    ms = time_graph(step)
    return {"config": "c1: GPLayer SVGP, SE, M=50, snelson1d N=200 (full batch on every rank)", "n_gpus": WORLD,
            "ms_per_step": ms, "rows_per_s": WORLD * 200 / ms * 1e3}


def c4():
    rng = np.random.default_rng(0)
    drng = np.random.default_rng(1000 + RANK)  # same parameters on every rank, different rows
    N, W, M = 8192, 2, 512
    X = torch.from_numpy(drng.uniform(0, 1, (N, 1))).cuda()
    Y = torch.from_numpy(np.sin(6 * X.cpu().numpy()) + 0.1 * drng.standard_normal((N, 1))).cuda()
    mean = torch.from_numpy(0.1 * drng.standard_normal((N, W))).cuda().requires_grad_(True)
    std = torch.from_numpy(drng.uniform(0.5, 1.0, (N, W))).cuda().requires_grad_(True)
    pm, ps = torch.zeros(W, dtype=torch.float64, device="cuda"), torch.ones(W, dtype=torch.float64, device="cuda")
    kernel = gf.SharedIndependent(gf.SquaredExponential(variance=1.0, lengthscales=np.ones(1 + W)), 1)
    iv = gf.SharedIndependentInducingVariables(gf.InducingPoints(rng.uniform(-1, 1, (M, 1 + W))))
    layer = GPLayer(kernel, iv, num_data=2_000_000, mean_function=gf.Zero())
    layer.q_mu.assign(0.1 * rng.standard_normal((M, 1)))
    lik = LikelihoodLayer(gf.Gaussian(0.1))
    layer, lik = layer.cuda(), lik.cuda()
    flat = flat_grads(p for p in list(layer.parameters()) + list(lik.parameters()) if p.requires_grad)

    def step():
        flat.zero_()
        mean.grad = std.grad = None  # per-row variational parameters: local to the rank that owns the rows
        eps = torch.randn((N, W), dtype=torch.float64, device="cuda")
        h, kl_w = ops.latent_sample_and_kl(X, mean, std, pm, ps, eps)
        f = layer(h, training=True)
        lik(f, targets=Y, training=True)
        loss = kl_w / 2_000_000 + layer.losses[0] + lik.losses[0]
        loss.backward()
        flat.allreduce_mean()

    ms = time_graph(step)
    flops = 3 * (2 * M * M * 3 + 2 * M * N * 3 + M**3 / 3 + 2 * M * M * N + 4 * M * N + 2 * M * N)
    return {"config": "c4: latent (W=2) + GPLayer(M=512, D=3, P=1) CDE, minibatch 8192 per GPU of 2M", "n_gpus": WORLD,
            "ms_per_step": ms, "rows_per_s": WORLD * N / ms * 1e3, "algorithmic_tflops_per_gpu": flops / ms / 1e9}


if __name__ == "__main__":
    torch.cuda.set_device(LOCAL_RANK)
    if WORLD > 1:
        from gpflux_b200.distributed import GpfxComm

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", LOCAL_RANK))
        COMM = GpfxComm()
    res = [c1(), c4()]
    if RANK == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        name = "bench_c1_c4.json" if WORLD == 1 else f"bench_c1_c4_n{WORLD}.json"
        json.dump(res, open(os.path.join("gpurun_out", name), "w"), indent=1)
        for r in res:
            print(json.dumps(r))
    torch.cuda.synchronize()
    if WORLD > 1:
        COMM.destroy()
        torch.distributed.destroy_process_group()
