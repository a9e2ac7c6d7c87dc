"""Achieved HBM bandwidth of the streaming (HBM-bound) kernels of the SVGP path - the sampling (a8),
whitened-KL (a9) and Gaussian variational-expectation (f-1) kernels - at c3-per-GPU and larger
sizes, with CUDA events on the launching stream and an L2 flush between launches.  Algorithmic
bytes: a8 32 B/element (3 reads + 1 write); a9 8 P (M^2 + M) B; VE 24 B/row (+ 2 x 8 B gradients).
Writes gpurun_out/elementwise_bw.json."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


def main():
    from gpflux_b200 import ops

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6454.6))
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    out = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": hbm, "cases": []}

    def add(name, nbytes, fn):
        med, best = timed(fn, flush)
        out["cases"].append({"kernel": name, "algorithmic_bytes": nbytes, "ms_median": med, "ms_best": best,
                             "gbs_median": nbytes / med / 1e6, "gbs_best": nbytes / best / 1e6,
                             "frac_of_hbm_peak": nbytes / med / 1e6 / hbm})
        print(out["cases"][-1])

    # a8: f = mean + sqrt(var) eps
    for n in (8192 * 32, 1 << 22, 1 << 26):
        m = torch.randn(n, dtype=torch.float64, device=dev)
        v = torch.rand(n, dtype=torch.float64, device=dev) + 0.1
        e = torch.randn(n, dtype=torch.float64, device=dev)
        add(f"reparam_sample n={n}" + (" (TMA-staged)" if n >= (1 << 20) else " (c3 layer: launch-latency bound)"),
            32 * n, lambda: ops.reparam_sample(m, v, e))
        # the same map with eps generated in the kernel: 24 B / element of traffic, plus the FP64
        # log / sincospi / sqrt of Box-Muller (whichever of the two bounds it is what gets measured)
        add(f"reparam_sample_philox n={n} (no eps array; 24 B/elt)", 24 * n,
            lambda: ops.reparam_sample_philox(m, v, seed=1, offset=0))
        add(f"torch.randn + reparam_sample n={n} (what the Philox kernel replaces; 40 B/elt)", 40 * n,
            lambda: ops.reparam_sample(m, v, torch.randn(n, dtype=torch.float64, device=dev)))
        del m, v, e
    # a9: whitened KL, c3 layer (P=32, M=1024) and c2 hidden layer
    for (M, P) in ((1024, 32), (256, 8), (2048, 32)):
        q_mu = torch.randn(M, P, dtype=torch.float64, device=dev)
        q_sqrt = torch.tril(torch.randn(P, M, M, dtype=torch.float64, device=dev)) + 2 * torch.eye(M, dtype=torch.float64, device=dev)
        # the kernel reads only the lower triangle: 8 P (M (M+1)/2 + M) bytes actually needed
        add(f"kl_white M={M} P={P} (lower triangle read)", 8 * P * (M * (M + 1) // 2 + M), lambda: ops.kl_white(q_mu, q_sqrt))
        g = torch.ones((), dtype=torch.float64, device=dev)
        add(f"kl_white_bwd M={M} P={P} (tril read + full write)", 8 * P * (M * (M + 1) // 2 + M * M + 2 * M),
            lambda: ops.kl_white_bwd(q_mu, q_sqrt, g))
        del q_mu, q_sqrt
    # f-1: Gaussian variational expectations with gradients
    for n in (65536, 1 << 24):
        Fmu = torch.randn(n, 1, dtype=torch.float64, device=dev)
        Fvar = torch.rand(n, 1, dtype=torch.float64, device=dev) + 0.1
        Y = torch.randn(n, 1, dtype=torch.float64, device=dev)
        nv = torch.tensor(0.1, dtype=torch.float64, device=dev)
        add(f"gaussian_ve n={n} (fwd + dFmu, dFvar)", 40 * n,
            lambda: ops.gaussian_variational_expectations_sum(Fmu, Fvar, Y, nv))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/elementwise_bw.json", "w"), indent=1)


if __name__ == "__main__":
    main()
