"""One GPU's shard of BASELINE config 3: 5-layer DeepGP, M=1024, SeparateIndependent kernels with
P=32 outputs on the 4 hidden layers (+ a P=1 last layer), D=32, 8192 rows (= 65536 / 8 GPUs).
Times forward+backward with CUDA events and reports algorithmic TFLOP/s (SURVEY §8d flops).
Usage: python tools/bench_c3.py [--layers 5] [--M 1024] [--P 32] [--N 8192] [--steps 3]
Under torchrun (WORLD_SIZE > 1) every rank evaluates its own 8192-row shard of the 65536-row
minibatch and the step includes the NCCL all-reduce of the flat FP64 gradient buffer (1.07 GB);
time = max over ranks, rows/s = all ranks' rows / that time (weak scaling, the c3 configuration)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gpflux_b200 import gpflow_compat as gf  # noqa: E402
from gpflux_b200.layers import GPLayer, LikelihoodLayer  # noqa: E402
from gpflux_b200.models import DeepGP  # noqa: E402


def f_fwd(N, M, D, P, K):
    return (2 * K * M * M * D + 2 * K * M * N * D + K * M**3 / 3 + K * M * M * N + 2 * M * N * P
            + 2 * K * M * N + P * M * M * N + 2 * P * M * N + 2 * P * (M * M + 2 * M))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--M", type=int, default=1024)
    ap.add_argument("--P", type=int, default=32)
    ap.add_argument("--N", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-overlap", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rng = np.random.default_rng(0)
    D = P = args.P
    M, N = args.M, args.N
    layers = []
    flops = 0.0
    for i in range(args.layers):
        last = i == args.layers - 1
        Pl = 1 if last else P
        kernels = [gf.SquaredExponential(variance=1.0 if last else 1e-2, lengthscales=rng.uniform(1.0, 3.0, D) * np.sqrt(D) / 2)
                   for _ in range(Pl)]
        ivs = [gf.InducingPoints(rng.standard_normal((M, D))) for _ in range(Pl)]
        kern = gf.SeparateIndependent(kernels) if Pl > 1 else gf.SharedIndependent(kernels[0], 1)
        iv = gf.SeparateIndependentInducingVariables(ivs) if Pl > 1 else gf.SharedIndependentInducingVariables(ivs[0])
        layer = GPLayer(kern, iv, num_data=1_000_000, mean_function=gf.Zero() if last else gf.Identity(), name=f"gp_{i}")
        layer.q_mu.assign(0.1 * rng.standard_normal((M, Pl)))
        q = np.stack([1e-2 * np.eye(M) + 1e-3 * np.tril(rng.standard_normal((M, M)), -1) for _ in range(Pl)])
        layer.q_sqrt.assign(q)
        layers.append(layer)
        flops += 3 * f_fwd(N, M, D, Pl, Pl)
    model = DeepGP(layers, LikelihoodLayer(gf.Gaussian(0.01))).cuda()
    if args.no_overlap:
        model.overlap_factorization = False
    from gpflux_b200.distributed import FlatGradients

    drng = np.random.default_rng(100 + rank)  # same parameters on every rank, different rows
    X = torch.from_numpy(drng.standard_normal((N, D))).cuda()
    Y = torch.from_numpy(drng.standard_normal((N, 1))).cuda()
    flat = FlatGradients(model.parameters())

    def step():
        flat.zero_()
        loss = model.loss(X, Y)
        loss.backward()
        flat.allreduce_mean()
        return loss

    loss = step()
    torch.cuda.synchronize()
    assert torch.isfinite(loss), loss
    ts = []
    for _ in range(args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        b.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    ms = min(ts)
    out = {"workload": f"c3: {args.layers} layers, M={M}, P={P} separate, D={D}, {N} rows per GPU x {world} GPU(s)"
                       + (" + all-reduce of the flat gradient buffer" if world > 1 else ""),
           "n_gpus": world, "global_batch": N * world, "grad_doubles": flat.numel(),
           "ms_per_step": ms, "rows_per_s": N * world / ms * 1e3, "algorithmic_tflops_per_gpu": flops / ms / 1e9,
           "algorithmic_flops_per_step_per_gpu": flops, "loss": float(loss),
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9, "all_ms": ts}
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/bench_c3_n{world}.json", "w"), indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
