"""BASELINE config 5 (efficient posterior sampling): throughput of the fused Wilson-sample kernel
``gpfx_wilson_eval`` - L = 10 000 random Fourier features, M = 256 inducing points, S draws,
N query rows - for D = P = 1 and D = P = 8.  Reports rows*draws/s and cos evaluations/s, and checks
one small slice against the CPU oracle.  Writes gpurun_out/bench_c5.json.
Under torchrun (``python -m torch.distributed.run --nproc-per-node N ... tools/bench_c5.py --literal``) only the
literal config runs: its 10^6 query rows are partitioned over the ranks (every rank holds the same draws, no
data-path collective), time = max over ranks -> gpurun_out/bench_c5_nN.json ("scaling": "strong")."""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpflux_b200 import ops  # noqa: E402


def run(N, S, L, M, D, P, reps=3):
    rng = np.random.default_rng(0)
    dev = "cuda"
    X = torch.from_numpy(rng.standard_normal((N, D))).to(dev)
    W = torch.from_numpy(rng.standard_normal((L, D))).to(dev)
    b = torch.from_numpy(rng.uniform(0, 2 * math.pi, (L,))).to(dev)
    ls = torch.from_numpy(rng.uniform(0.8, 1.5, (D,))).to(dev)
    var = torch.tensor([1.3], dtype=torch.float64, device=dev)
    Z = torch.from_numpy(rng.standard_normal((M, D))).to(dev)
    w = torch.from_numpy(rng.standard_normal((S, L, P))).to(dev)
    v = torch.from_numpy(rng.standard_normal((S, M, P))).to(dev)
    out = ops.wilson_eval("se", X, W, b, ls, var, Z, w, v)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = ops.wilson_eval("se", X, W, b, ls, var, Z, w, v)
        e.record()
        e.synchronize()
        ts.append(a.elapsed_time(e))
    ms = min(ts)
    # oracle check on a slice (first 64 rows of draw 0)
    from oracle import svgp_torch as ot

    spec = {"kernel": "se", "Z": Z.cpu()[None], "lengthscales": ls.cpu()[None], "variance": var.cpu(), "q_mu": torch.zeros(M, P, dtype=torch.float64),
            "mean": "zero"}
    ref = ot.wilson_eval(spec, W.cpu(), b.cpu()[None], w[0].cpu(), v[0].cpu(), X[:64].cpu())
    err = float((out[0, :64].cpu() - ref).abs().max() / ref.abs().max())
    return {"N": N, "S": S, "L": L, "M": M, "D": D, "P": P, "ms": ms, "rows_draws_per_s": N * S / ms * 1e3,
            "cos_per_s": N * S * L / ms * 1e3, "flops_per_s": N * S * (L * (2 * D + 2 * P) + M * (3 * D + 2 * P)) / ms * 1e3,
            "rel_err_vs_oracle": err}


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))


def literal(N=1_000_000, S=64, L=10000, M=256, layers=3, DP=1, chunk=250_000):
    """The literal BASELINE config 5: a 3-layer DGP, 10^6 query rows x 64 draws, 10k features per layer;
    draw s of layer i is evaluated at draw s of layer i-1 (sample_dgp).  Rows are processed in chunks so the
    [S, chunk, P] intermediates stay small; returns the device time of the whole job."""
    rng = np.random.default_rng(1)
    dev = "cuda"
    par = []
    for _ in range(layers):
        par.append(dict(
            W=torch.from_numpy(rng.standard_normal((L, DP))).to(dev), b=torch.from_numpy(rng.uniform(0, 2 * math.pi, (L,))).to(dev),
            ls=torch.from_numpy(rng.uniform(0.8, 1.5, (DP,))).to(dev), var=torch.tensor([1.0], dtype=torch.float64, device=dev),
            Z=torch.from_numpy(rng.standard_normal((M, DP))).to(dev), w=torch.from_numpy(rng.standard_normal((S, L, DP))).to(dev),
            v=torch.from_numpy(0.1 * rng.standard_normal((S, M, DP))).to(dev)))
    from gpflux_b200.distributed import shard_rows

    r0, r1 = shard_rows(N, RANK, WORLD)  # this rank's rows of the query set
    X = torch.from_numpy(rng.standard_normal((N, DP))[r0:r1]).to(dev)
    for p in par[:1]:  # warm-up launch (kernel attributes, workspace)
        ops.wilson_eval("se", X[:1024].contiguous(), p["W"], p["b"], p["ls"], p["var"], p["Z"], p["w"], p["v"])
    torch.cuda.synchronize()
    if WORLD > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for c0 in range(0, r1 - r0, chunk):
        f = X[c0:c0 + chunk].contiguous()
        for p in par:
            f = ops.wilson_eval("se", f, p["W"], p["b"], p["ls"], p["var"], p["Z"], p["w"], p["v"])  # [S, chunk, P]
    e.record()
    e.synchronize()
    t = torch.tensor([a.elapsed_time(e)], dtype=torch.float64, device=dev)
    if WORLD > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t)
    return {"literal_c5": True, "n_gpus": WORLD, "scaling": "strong", "rows_per_gpu": r1 - r0, "layers": layers, "N": N, "S": S,
            "L": L, "M": M, "D": DP, "P": DP, "ms": ms,
            "rows_draws_per_s": N * S / ms * 1e3, "cos_per_s": layers * N * S * L / ms * 1e3, "finite": bool(torch.isfinite(f).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=262144)
    ap.add_argument("--S", type=int, default=4)
    ap.add_argument("--literal", action="store_true", help="also run the literal config 5 (3 layers, 10^6 rows x 64 draws)")
    args = ap.parse_args()
    torch.cuda.set_device(LOCAL_RANK)
    if WORLD > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", LOCAL_RANK))
        res = [literal(), literal(DP=8, S=8)]
    else:
        res = [run(args.N, args.S, 10000, 256, 1, 1), run(args.N, args.S, 10000, 256, 8, 8), run(args.N // 4, args.S, 10000, 256, 32, 32)]
        if args.literal:
            res.append(literal())
            res.append(literal(DP=8, S=8))
    if RANK == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        name = "bench_c5.json" if WORLD == 1 else f"bench_c5_n{WORLD}.json"
        json.dump(res, open(os.path.join("gpurun_out", name), "w"), indent=1)
        for r in res:
            print(json.dumps(r))
    if WORLD > 1:
        torch.cuda.synchronize()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
