#!/usr/bin/env python
"""bench.py - DeepGP ELBO forward+backward throughput (minibatch rows/s) on B200.

Default workload = BASELINE.json configs[2] ("c3", the configuration the north-star target is quoted
on): 5-layer DeepGP, M=1024 inducing points, SeparateIndependent P=32 outputs on the four hidden
layers (+ a P=1 last layer), D=32, float64, 8192 minibatch rows per GPU (weak scaling: 65536 rows at
8 GPUs).  ``--config c2`` selects configs[1] (2-layer ``build_constant_input_dim_deep_gp``, M=256,
D=8, minibatch 8192/GPU) - the smaller configuration.

One "step" = forward through all layers + likelihood, backward to every parameter, and (N>1) the
NCCL all-reduce of the flat FP64 gradient buffer, one bucket per layer, overlapped with the backward
of the layers below.  No optimiser step (the metric is ELBO fwd+bwd).

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU oracle (the GPflow algorithm
restated in torch float64 - TensorFlow/GPflow are not installable here) on the host
cores for the same metric and config, each step on a bounded row sample of the minibatch.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

_RESULT_FD = None


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "deepgp_elbo_fwd_bwd_rows_per_s"
UNIT = "rows/s"
POOL_BATCHES = 8  # c3: minibatches of synthetic rows per rank that the steps cycle through
CPU_SAMPLE_ROWS = {"c3": 2048, "c2": 8192}  # rows per reference-arm step (c2: the whole minibatch)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 10 (c3) / 2000 (c2): >= 2 s timed")
    ap.add_argument("--warmup", type=int, default=None, help="default: 3 (c3) / 20 (c2)")
    ap.add_argument("--impl", default="gpflux_b200", choices=["gpflux_b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c2"])
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="capture the step in a CUDA graph (default: c2 yes, c3 no - 0.3 s steps are not launch-bound)")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--no-overlap", action="store_true", help="run the factor part of every layer in line")
    ap.add_argument("--batch", type=int, default=8192, help="minibatch rows per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=30.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--skip-extras", action="store_true", help="headline timing only (no roofline / hbm / cpu legs)")
    ap.add_argument("--watchdog-seconds", type=int, default=2400,
                    help="dump every thread's Python stack to stderr and exit(3) if the run lasts longer (0 = off)")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 10 if a.config == "c3" else 2000
    if a.warmup is None:
        a.warmup = 3 if a.config == "c3" else 20
    if a.graph is None:
        a.graph = a.config == "c2"
    return a


def base_config(name: str, batch: int, world: int) -> dict:
    """The ``config`` object, identical in the product arm and the reference arm."""
    from gpflux_b200 import workloads as wl

    c = wl.C3 if name == "c3" else wl.C2
    return {"workload": wl.C3_WORKLOAD if name == "c3" else wl.C2_WORKLOAD,
            "global_batch": batch * world, "rows_per_gpu": batch, "layers": c["layers"], "M": c["M"], "D": c["D"],
            "P": c["P"], "parallelism": f"dp{world}"}


def build_workload(name: str, batch: int, rank: int):
    """-> (model on CPU, X_pool, Y_pool, num_data); pools are host numpy arrays of whole minibatches."""
    from gpflux_b200 import workloads as wl

    if name == "c2":
        X, Y = wl.make_data_c2()
        return wl.build_c2(X), X, Y, wl.N_DATA_C2
    model = wl.build_c3()
    X, Y = wl.make_data_c3(POOL_BATCHES * batch, seed=100 + rank)  # same parameters on every rank, different rows
    return model, X, Y, wl.N_DATA_C3


# ------------------------------------------------------------------------------------------ CPU leg
def cpu_oracle_step_seconds(specs, noise, Xb, Yb, num_data):
    """One fwd+bwd of the oracle port on the host cores -> seconds."""
    import torch

    from oracle import svgp_torch as ot

    sp = [ot.spec_requires_grad(s) for s in specs]
    nv = noise.clone().requires_grad_(True)
    eps = [torch.randn(Xb.shape[0], s["q_mu"].shape[1], dtype=torch.float64) for s in sp[:-1]] + [None]
    t0 = time.perf_counter()
    loss, _ = ot.deep_gp_loss(sp, Xb, Yb, eps, nv, num_data)
    loss.backward()
    return time.perf_counter() - t0


def run_reference(args):
    """The reference arm: the CPU oracle (port of the GPflow algorithm the reference calls), all host
    threads, K timed steps after W warm-up steps, each step a bounded row sample of the minibatch."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.from_model import noise_variance_from_model, specs_from_model

    torch.set_num_threads(os.cpu_count() or 1)
    model, X, Y, num_data = build_workload(args.config, args.batch, 0)
    specs, noise = specs_from_model(model), noise_variance_from_model(model)
    rows = min(CPU_SAMPLE_ROWS[args.config], args.batch)
    n_slices = X.shape[0] // rows
    times = []
    for i in range(args.warmup + args.steps):
        sl = slice((i % n_slices) * rows, (i % n_slices) * rows + rows)
        dt = cpu_oracle_step_seconds(specs, noise, torch.from_numpy(X[sl]), torch.from_numpy(Y[sl]), num_data)
        if i >= args.warmup:
            times.append(dt)
    ms = statistics.mean(times) * 1e3
    val = rows / (ms / 1e3)
    cores = torch.get_num_threads()
    sample = (f"{len(times)} steps, each the full {len(specs)}-layer fwd+bwd on {rows} of the {args.batch} minibatch rows "
              f"(oracle/svgp_torch.py, torch float64 CPU + autograd, {cores} threads, {ms:.0f} ms/step)")
    if rows < args.batch:
        sample += ("; the parameter-only work (Kuu, Cholesky, KL and their backward) does not shrink with the row sample, "
                   "so rows/s on the sample understates the CPU rate on the whole minibatch - see cpu_baseline.full_step_estimate "
                   "of the product arm")
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args.config, args.batch, args.gpus),
        "note": "GPflow/TensorFlow are absent from this image: the reference arm is the oracle port of the GPflow algorithm",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------ GPU leg
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


TRI_NAMES = {0: "dense", 1: "lower-A", 2: "upper-A", 4: "lower-B", 8: "upper-B", 5: "lower-A x lower-B"}


def gemm_algorithmic_flops(s) -> float:
    """Algorithmic flops of one GEMM launch: 2 b nseg M N K, a triangular operand or a lower-triangular
    output counted once (x 1/2) - the SURVEY §8(d) convention (B_p = Lq_p^T A is P M^2 N)."""
    f = 2.0 * s["batch"] * s["nseg"] * s["M"] * s["N"] * s["K"]
    if s["tri"]:
        f *= 0.5
    if s["tril_out"]:
        f *= 0.5
    return f


def read_gemm_profile(lib, _cabi):
    n_max = 16384
    buf = (_cabi.GemmSample * n_max)()
    n = lib.gpfx_gemm_profile_read(buf, n_max)
    groups = {}
    for i in range(n):
        s = buf[i]
        key = (s.M, s.N, s.K, s.batch, s.nseg, s.tri, s.transA, s.transB, s.tril_out, s.splits, s.cfg)
        g = groups.setdefault(key, [0, 0.0])
        g[0] += 1
        g[1] += s.ms
    out = []
    for key, (cnt, ms) in groups.items():
        d = dict(zip(("M", "N", "K", "batch", "nseg", "tri", "transA", "transB", "tril_out", "splits", "cfg"), key))
        d["launches"] = cnt
        d["total_ms"] = ms
        d["launch_ms"] = ms / cnt
        d["algorithmic_flops_per_launch"] = gemm_algorithmic_flops(d)
        d["tflops"] = d["algorithmic_flops_per_launch"] / d["launch_ms"] / 1e9
        out.append(d)
    out.sort(key=lambda d: -d["total_ms"])
    return out


def gemm_label(d) -> str:
    op = ("T" if d["transA"] else "N") + ("T" if d["transB"] else "N")
    tri = TRI_NAMES.get(d["tri"], f"tri={d['tri']}")
    extra = (", tril output" if d["tril_out"] else "") + (f", split-K {d['splits']}" if d["splits"] > 1 else "") \
        + (f", {d['nseg']} k-segments" if d["nseg"] > 1 else "")
    if d["cfg"] >= 100:  # TMA-fed persistent kernel (gemm_f64_tma.cu): 101 = 128x128 tile, 102 = 64x128 tile, 2 CTAs/SM
        kern = f"gemm_f64_tma_kernel<{'128x128' if d['cfg'] == 101 else '64x128'}> (UTMALDG ring + mbarriers, DMMA.8x8x4)"
    else:
        kern = f"gemm_f64_kernel cfg {d['cfg']} (cp.async, DMMA.8x8x4)"
    return f"{kern} {op} batch={d['batch']} M={d['M']} N={d['N']} K={d['K']} {tri}{extra}"


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from gpflux_b200 import _cabi
    from gpflux_b200.distributed import FlatGradients, GpfxComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - gpflux_b200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = GpfxComm()  # the library's own NCCL communicator (gpfx_allreduce_grads)
    lib = _cabi.load()

    B = args.batch
    cfg_name = args.config
    model, X, Y, num_data = build_workload(cfg_name, B, rank)
    specs_cpu = noise_cpu = None
    if world == 1 and not args.skip_extras:
        from oracle.from_model import noise_variance_from_model, specs_from_model

        specs_cpu, noise_cpu = specs_from_model(model), noise_variance_from_model(model)
    model = model.to(dev)
    if args.no_overlap:
        model.overlap_factorization = False
    flat = FlatGradients.for_model(model, comm=comm, overlap=True)
    D_in = X.shape[1]
    Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
    Xh, Yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
    if cfg_name == "c2":
        n_slices = X.shape[0] // (B * world)

        def batch_slice(step):
            g = (step % n_slices) * B * world + rank * B
            return slice(g, g + B)
    else:
        n_slices = X.shape[0] // B

        def batch_slice(step):
            g = (step % n_slices) * B
            return slice(g, g + B)

    static_X = torch.empty((B, D_in), dtype=torch.float64, device=dev)
    static_Y = torch.empty((B, 1), dtype=torch.float64, device=dev)
    loss_out = torch.zeros((), dtype=torch.float64, device=dev)

    def fwd_bwd():
        flat.zero_()
        loss = model.loss(static_X, static_Y)
        loss.backward()
        flat.allreduce_mean()  # buckets not yet reduced by the hooks + join of the communication stream
        loss_out.copy_(loss.detach())

    # eager warm-up (also sets kernel attributes, sizes the workspace) + per-step launch count
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(2):
            static_X.copy_(Xd[batch_slice(i)])
            static_Y.copy_(Yd[batch_slice(i)])
            c0 = lib.gpfx_kernel_launch_count()
            fwd_bwd()
            launches_per_step = lib.gpfx_kernel_launch_count() - c0
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    if not torch.isfinite(loss_out):
        raise SystemExit(f"bench.py: non-finite loss {float(loss_out)}")

    holder = {"graph": None}
    graph = None
    graph_err = None
    if args.graph:
        for overlap in ((True, False) if model.overlap_factorization else (False,)):
            model.overlap_factorization = overlap
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fwd_bwd()
                graph = holder["graph"] = g
                break
            except Exception as e:  # retry without multi-stream overlap, then fall back to eager launches
                graph_err = f"overlap={overlap}: {type(e).__name__}: {e}"[:300]
                torch.cuda.synchronize()
        g = None  # `holder` and `graph` are the only references from here on (see teardown)

    def run_step():
        if holder["graph"] is not None:
            holder["graph"].replay()
        else:
            fwd_bwd()

    def teardown():
        # A live CUDA graph holds references to the communicator's captured collectives and ncclCommDestroy waits
        # for them: every reference to the graph must be gone first (seen as a hang of c2 at N > 1).  The destroy
        # calls run on a helper thread with a deadline so that a teardown problem can never hold the box: the
        # result line has been written by then.
        import gc
        import threading

        holder["graph"] = None
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            done = threading.Event()

            def _destroy():
                try:
                    torch.cuda.set_device(local_rank)  # (a new thread starts on device 0)
                    comm.destroy()
                    dist.destroy_process_group()
                finally:
                    done.set()

            threading.Thread(target=_destroy, daemon=True).start()
            if not done.wait(30.0):
                sys.stderr.write(f"bench.py: rank {rank}: communicator teardown did not return within 30 s; exiting without it\n")
                sys.stderr.flush()
                os._exit(0)

    def step_resident(i):
        sl = batch_slice(i)
        static_X.copy_(Xd[sl])
        static_Y.copy_(Yd[sl])
        run_step()

    def step_e2e(i):
        sl = batch_slice(i)
        static_X.copy_(Xh[sl], non_blocking=True)
        static_Y.copy_(Yh[sl], non_blocking=True)
        run_step()
        return float(loss_out.item())  # device -> host read of the step's loss

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    profile_in_step = holder["graph"] is None and rank == 0 and not args.skip_extras
    if profile_in_step:
        lib.gpfx_gemm_profile_read(None, 0)
        lib.gpfx_gemm_profile_enable(1)
    barrier()
    for i in range(args.steps):
        flush.fill_(float(i))  # evict L2 between timed iterations (untimed)
        ev[i][0].record()
        step_resident(args.warmup + i)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    gemm_groups = None
    if profile_in_step:
        lib.gpfx_gemm_profile_enable(0)
        gemm_groups = read_gemm_profile(lib, _cabi)
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    # end-to-end: pinned host -> device copies and the loss read-back inside the timed region (wall clock,
    # no L2 flush: the step's own working set, 2-45 GB, is far larger than the 126 MB L2)
    e2e_warm = min(2, args.warmup)
    for i in range(e2e_warm):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / float(t.item())

    peak_mem_gb = torch.cuda.max_memory_allocated() / 1e9
    has_graph = graph is not None
    del graph
    # c2 (graph replay: events cannot be read out of a graph): three extra eager steps after the timed region are
    # profiled for the roofline leg.  They contain the gradient all-reduce, so EVERY rank runs them (rank 0 alone
    # would wait for peers that have already left); only rank 0 reads the profile.
    prof_steps, prof_where = args.steps, "the timed steps themselves"
    if has_graph or args.skip_extras:  # (the same on every rank; rank 0 has no in-step profile exactly then)
        if rank == 0:
            lib.gpfx_gemm_profile_read(None, 0)
            lib.gpfx_gemm_profile_enable(1)
        for i in range(3):
            flush.fill_(0.0)
            static_X.copy_(Xd[batch_slice(i)])
            static_Y.copy_(Yd[batch_slice(i)])
            fwd_bwd()
        torch.cuda.synchronize()
        if rank == 0:
            lib.gpfx_gemm_profile_enable(0)
            gemm_groups = read_gemm_profile(lib, _cabi)
        prof_steps, prof_where = 3, "3 eager steps after the timed region (the timed steps are CUDA-graph replays)"
    if rank != 0:
        teardown()
        return

    from gpflux_b200 import workloads as wl

    cfg = wl.C3 if cfg_name == "c3" else wl.C2
    flops_step = wl.algorithmic_flops_per_step(cfg, B)
    config = base_config(cfg_name, B, world)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "run": {"cuda_graph": has_graph, "cuda_graph_error": graph_err,
                "overlap_factorization": bool(model.overlap_factorization),
                "l2": "256 MB flush buffer written between timed iterations (untimed); e2e leg: no flush, 2 warm-up steps",
                "gradient_allreduce": (f"{flat.numel()} f64 in {len(flat.buckets)} per-layer buckets "
                                       f"{flat.bucket_numels()}, gpfx_allreduce_grads (NCCL {lib.gpfx_comm_nccl_version()}) "
                                       "on a communication stream, overlapped with the backward of earlier layers"
                                       if world > 1 else "none (1 GPU)"),
                "peak_mem_gb": peak_mem_gb, "timed_region_s": ms_total / 1e3},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (D_in + 1) * 8, "d2h_bytes_per_step": 8},
        "gpu_launches": int(launches_per_step) * args.steps,
        "gpu_launches_per_step": int(launches_per_step),
        "algorithmic_tflops_per_gpu": flops_step / (ms_per_step * 1e-3) / 1e12,
    }
    if args.skip_extras:
        emit(out)
        teardown()
        return

    # ---- FP64 peak: MEASURED_PEAKS.json has no FP64 entry, so cuBLAS DGEMM 8192^3 is measured here
    # (torch.matmul float64, best of 5 after one warm-up: the "burst" protocol of MEASURED_PEAKS.json)
    n = 8192
    Ga = torch.randn(n, n, dtype=torch.float64, device=dev)
    Gb = torch.randn(n, n, dtype=torch.float64, device=dev)
    Gc = torch.empty_like(Ga)
    best = 1e30
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        torch.matmul(Ga, Gb, out=Gc)
        b.record()
        b.synchronize()
        if i > 0:
            best = min(best, a.elapsed_time(b))
    fp64_peak = 2.0 * n**3 / best / 1e9
    del Ga, Gb, Gc
    out["fp64_fraction_whole_step"] = out["algorithmic_tflops_per_gpu"] / fp64_peak

    # ---- roofline of the dominant kernel.  c3 (eager): every DMMA GEMM launch of the TIMED steps was
    # bracketed by CUDA events on its own stream (gpfx_gemm_profile_*); the launches are grouped by shape
    # and the group with the largest share of the step is the dominant kernel.  c2: the three eager steps above.
    traffic = None
    traffic_src = None
    try:  # dram bytes of the dominant kernel from the committed ncu --set full capture, if one matches
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_dominant_r02.json")))
        traffic_src = "profiles/ncu_dominant_r02.json"
    except Exception:
        ncu = {}
    roofline = None
    if gemm_groups:
        top = gemm_groups[0]
        key = f"{cfg_name}:{top['batch']}x{top['M']}x{top['N']}x{top['K']}:tri{top['tri']}:t{top['transA']}{top['transB']}:lo{top['tril_out']}"
        if key in ncu:
            traffic = ncu[key].get("dram_bytes_per_launch")
        roofline = {
            "bound": "tensor", "achieved": top["tflops"], "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": top["tflops"] / fp64_peak, "traffic": traffic,
            "traffic_source": traffic_src if traffic is not None else None,
            "kernel": gemm_label(top), "kernel_key": key,
            "launch_ms": top["launch_ms"], "launches_per_step": top["launches"] / prof_steps,
            "share_of_step": top["total_ms"] / prof_steps / ms_per_step,
            "algorithmic_flops_per_launch": top["algorithmic_flops_per_launch"],
            "measured": f"CUDA events around every launch on its own stream, {prof_where}",
            "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul float64) measured in this run; MEASURED_PEAKS.json has no FP64 entry",
        }
        out["gemm_kernels"] = [
            {"kernel": gemm_label(d), "launches_per_step": d["launches"] / prof_steps, "launch_ms": d["launch_ms"],
             "tflops": d["tflops"], "frac": d["tflops"] / fp64_peak, "share_of_step": d["total_ms"] / prof_steps / ms_per_step}
            for d in gemm_groups[:8]]
        out["gemm_share_of_step"] = sum(d["total_ms"] for d in gemm_groups) / prof_steps / ms_per_step
    out["roofline"] = roofline

    # ---- HBM-bound kernels of the path (BASELINE metric: "... % FP64 TC peak & HBM GB/s"): the sampling
    # kernel a8 (TMA-staged, 2^26 elements, 32 B/element) and the whitened-KL forward/backward a9/a14 at the
    # c3 layer shape (P = 32, M = 1024), CUDA events, L2 flushed between launches; peak = MEASURED_PEAKS.json
    hbm = None
    try:
        from gpflux_b200 import ops

        hbm_peak, hbm_src = 6454.6, "fallback 6454.6 GB/s (MEASURED_PEAKS.json absent)"
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        stream = torch.cuda.current_stream().cuda_stream

        def timed(fn, reps=10):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush.fill_(0.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            return statistics.median(ts)

        n_el = 1 << 26
        hm = torch.randn(n_el, dtype=torch.float64, device=dev)
        hv = torch.rand(n_el, dtype=torch.float64, device=dev) + 0.1
        he = torch.randn(n_el, dtype=torch.float64, device=dev)
        hf = torch.empty_like(hm)
        ms_s = timed(lambda: _cabi.check(lib.gpfx_reparam_sample(hm.data_ptr(), hv.data_ptr(), he.data_ptr(), hf.data_ptr(),
                                                                  n_el, stream)))
        del hm, hv, he, hf
        Mk, Pk = 1024, 32
        kq = torch.randn(Mk, Pk, dtype=torch.float64, device=dev)
        ks = torch.tril(torch.randn(Pk, Mk, Mk, dtype=torch.float64, device=dev)) + 2 * torch.eye(Mk, dtype=torch.float64, device=dev)
        kg = torch.ones((), dtype=torch.float64, device=dev)
        ms_k = timed(lambda: ops.kl_white_bwd(kq, ks, kg))
        ms_kf = timed(lambda: ops.kl_white(kq, ks))
        # ... and as the layer runs it: the same pass also materialises Lq = tril(q_sqrt) for the GEMMs
        # (gpfx_layer_factor_forward's first kernel), i.e. reads the triangle and writes the full [P,M,M]
        kf_scr = torch.empty(lib.gpfx_kl_scratch_bytes(Mk, Pk) // 8, dtype=torch.float64, device=dev)
        kf_Lq = torch.empty_like(ks)
        kf_out = torch.empty(1, dtype=torch.float64, device=dev)
        ms_kfl = timed(lambda: _cabi.check(lib.gpfx_kl_white_tril(Mk, Pk, kq.data_ptr(), ks.data_ptr(), kf_Lq.data_ptr(),
                                                             kf_out.data_ptr(), kf_scr.data_ptr(), stream)))
        del kf_Lq
        del kq, ks
        b_s = 32.0 * n_el
        b_k = 8.0 * Pk * (Mk * (Mk + 1) // 2 + Mk * Mk + 2 * Mk)
        b_kf = 8.0 * Pk * (Mk * (Mk + 1) // 2 + Mk)
        b_kfl = b_kf + 8.0 * Pk * Mk * Mk
        hbm = {"peak_gbs": hbm_peak, "peak_source": hbm_src,
               "kernels": [
                   {"kernel": "reparam_sample_tma_kernel (a8, 2^26 elements)", "algorithmic_bytes": b_s, "launch_ms": ms_s,
                    "achieved_gbs": b_s / ms_s / 1e6, "frac": b_s / ms_s / 1e6 / hbm_peak},
                   {"kernel": "kl_white_bwd_kernel (a9 backward, P=32, M=1024; includes host launch of one op)", "algorithmic_bytes": b_k,
                    "launch_ms": ms_k, "achieved_gbs": b_k / ms_k / 1e6, "frac": b_k / ms_k / 1e6 / hbm_peak},
                   {"kernel": "kl_tril_kernel (a9 forward, P=32, M=1024, lower triangle only; includes host launch of one op)",
                    "algorithmic_bytes": b_kf, "launch_ms": ms_kf, "achieved_gbs": b_kf / ms_kf / 1e6,
                    "frac": b_kf / ms_kf / 1e6 / hbm_peak},
                   {"kernel": "kl_tril_kernel as the layer runs it (a9 forward + Lq = tril(q_sqrt) written for the GEMMs, P=32, M=1024)",
                    "algorithmic_bytes": b_kfl, "launch_ms": ms_kfl, "achieved_gbs": b_kfl / ms_kfl / 1e6,
                    "frac": b_kfl / ms_kfl / 1e6 / hbm_peak}]}
    except Exception as e:  # the headline numbers do not depend on this section
        hbm = {"error": f"{type(e).__name__}: {e}"[:200]}
    out["hbm_kernels"] = hbm

    # ---- library baseline ("torch-CUDA float64, cuBLAS/cuSOLVER, unfused ... the library baseline to
    # beat"): the same workload with stock torch operators + autograd on this GPU (tools/torch_cuda_baseline.py).
    # c3's autograd graph of a whole 8192-row minibatch would hold ~60 GB of intermediates, so it is timed on
    # two row samples and extrapolated like the CPU leg; c2 runs whole minibatches.
    library = None
    if world == 1:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import torch_cuda_baseline as tcb

            holder["graph"] = None  # the replayed graph's buffers are not needed any more
            torch.cuda.synchronize()
            lp, lnoise = tcb.params_from_model(model, dev)
            if cfg_name == "c2":
                msl = tcb.step_ms(lp, lnoise, Xd[:B], Yd[:B], num_data, steps=10, warmup=3)
                library = {"value": B / (msl / 1e3), "unit": UNIT, "ms_per_step": msl,
                           "sample": f"10 eager steps of {B} rows (whole minibatch)"}
            else:
                n1, n2 = 1024, 2048
                t1 = tcb.step_ms(lp, lnoise, Xd[:n1], Yd[:n1], num_data, steps=2, warmup=1)
                t2 = tcb.step_ms(lp, lnoise, Xd[:n2], Yd[:n2], num_data, steps=2, warmup=1)
                bb = max((t2 - t1) / (n2 - n1), 1e-9)
                aa = max(t1 - bb * n1, 0.0)
                t_full = aa + bb * B
                library = {"value": B / (t_full / 1e3), "unit": UNIT, "ms_per_step": t_full,
                           "sample": f"eager fwd+bwd of the full 5-layer model on {n1} rows ({t1:.1f} ms) and {n2} rows ({t2:.1f} ms); "
                                     f"t(n) = {aa:.1f} ms + {bb:.4f} ms/row extrapolated to the {B}-row minibatch "
                                     "(its autograd graph would not fit beside the product's buffers)"}
            library["what"] = ("stock torch-CUDA float64 operators (cuBLAS matmul, cuSOLVER cholesky / triangular solve, "
                               "elementwise kernels, autograd backward), un-fused, same parameters, same GPU, same run")
            library["speedup_of_product"] = library["ms_per_step"] / ms_per_step
            del lp, lnoise
            torch.cuda.empty_cache()
        except Exception as e:  # the headline numbers do not depend on this section
            library = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
    out["library_baseline"] = library

    # ---- float32 opt-in (never the headline dtype): one hidden layer of the workload, fwd + bwd through the public
    # op, float64 vs gpflow_compat.set_default_float(np.float32) (float32 interface; B_p / dLq / dA contractions as
    # 3xTF32 tcgen05 products with TMEM accumulators, the cond(Kuu)-amplified products and everything else float64)
    fp32 = None
    if world == 1:
        try:
            import numpy as _np

            from gpflux_b200 import gpflow_compat as gfc
            from gpflux_b200 import ops as gops

            hl = model.f_layers[0]
            Dl = int(hl._inducing_points()[0].Z.shape[1])
            kind_l, Zl, lsl, varl = hl._packed(Dl)
            base = [t.detach() for t in (Zl, lsl, varl, hl.q_mu.value(), hl.q_sqrt.value())]
            Xl = Xd[:B].detach()
            Pl = base[3].shape[1]
            epsl = torch.randn(B, Pl, dtype=torch.float64, device=dev)
            wl_ = torch.randn(B, Pl, dtype=torch.float64, device=dev)

            def layer_ms(dt, reps=3):
                ts = []
                for i in range(reps + 1):
                    args = [t.to(dt).clone().requires_grad_(True) for t in base]
                    Xa = Xl.to(dt).clone().requires_grad_(True)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    f, m_, v_, kl_ = gops.svgp_layer(Xa, *args, mean_add=Xa if Pl == Dl else None, eps=epsl.to(dt), kernel=kind_l,
                                                     whiten=bool(hl.whiten), jitter=gfc.default_jitter())
                    ((f * wl_.to(dt)).sum() + kl_).backward()
                    b.record()
                    b.synchronize()
                    if i > 0:
                        ts.append(a.elapsed_time(b))
                    out_ = (f.detach().double(), args[4].grad.detach().double())
                return statistics.mean(ts), out_

            ms64, o64 = layer_ms(torch.float64)
            gfc.set_default_float(_np.float32)
            try:
                ms32, o32 = layer_ms(torch.float32)
                mode32 = int(lib.gpfx_get_contraction())
            finally:
                gfc.set_default_float(_np.float64)
                gops._prepare_contraction(gops._make_desc(1, 1, 1, 1, 1, "se", True, 1e-6))  # back to FP64 DMMA
            rel = lambda x, y: float((x - y).abs().max() / y.abs().max())  # noqa: E731
            fp32 = {"what": "first hidden layer of the workload (fwd + bwd through ops.svgp_layer): float64 vs the float32 opt-in",
                    "layer_ms_f64": ms64, "layer_ms_f32_optin": ms32, "speedup": ms64 / ms32, "contraction_mode": mode32,
                    "contraction": "B_p, dLq, dA: 3xTF32 on tcgen05 (kind::tf32, TMEM accumulators); A, dKuf, dL and all other kernels float64",
                    "max_rel_diff_fsample": rel(o32[0], o64[0]), "max_rel_diff_dq_sqrt": rel(o32[1], o64[1]), "tolerance": 1e-4}
            del base, Xl, epsl, wl_, o64, o32
            torch.cuda.empty_cache()
        except Exception as e:  # the headline numbers do not depend on this section
            fp32 = {"error": f"{type(e).__name__}: {e}"[:300]}
            try:
                from gpflux_b200 import gpflow_compat as gfc
                import numpy as _np

                gfc.set_default_float(_np.float64)
            except Exception:
                pass
    out["fp32_optin"] = fp32

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu_baseline = None
    if world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        cores = torch.get_num_threads()
        if cfg_name == "c2":
            ts = []
            t_budget = time.perf_counter()
            i = 0
            while True:
                sl = batch_slice(i)
                dt = cpu_oracle_step_seconds(specs_cpu, noise_cpu, torch.from_numpy(X[sl]), torch.from_numpy(Y[sl]), num_data)
                if i >= 1:
                    ts.append(dt)
                i += 1
                if ts and (time.perf_counter() - t_budget > args.cpu_seconds or len(ts) >= 50):
                    break
            msc = statistics.mean(ts) * 1e3
            cpu_baseline = {"value": B / (msc / 1e3), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{len(ts)} steps of {B} rows (whole minibatch), oracle/svgp_torch.py "
                                      f"(torch float64 CPU, autograd backward), {msc:.1f} ms/step"}
        else:
            # two sample sizes -> t(n) = a + b n (a: the parameter-only work, which does not shrink with the
            # sample) -> estimate for the whole 8192-row minibatch, reported beside the measured sample rate
            n1, n2 = 256, 1024
            t1 = cpu_oracle_step_seconds(specs_cpu, noise_cpu, torch.from_numpy(X[:n1]), torch.from_numpy(Y[:n1]), num_data)
            t2 = cpu_oracle_step_seconds(specs_cpu, noise_cpu, torch.from_numpy(X[:n2]), torch.from_numpy(Y[:n2]), num_data)
            bb = max((t2 - t1) / (n2 - n1), 1e-9)
            aa = max(t1 - bb * n1, 0.0)
            t_full = aa + bb * B
            cpu_baseline = {"value": n2 / t2, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"one fwd+bwd step of the full 5-layer model on {n2} of the {B} minibatch rows ({t2:.1f} s) and one "
                                      f"on {n1} rows ({t1:.1f} s), oracle/svgp_torch.py (torch float64 CPU, autograd backward)",
                            "full_step_estimate": {"rows_per_s": B / t_full, "seconds_per_step": t_full,
                                                   "method": f"t(n) = a + b n fitted to the two samples: a = {aa:.2f} s "
                                                             f"(parameter-only work), b = {bb * 1e3:.3f} ms/row"}}
    out["cpu_baseline"] = cpu_baseline
    emit(out)
    teardown()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. NCCL's version banner)
    # is diverted to stderr, and the result is written to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.watchdog_seconds > 0 and args.impl != "reference":  # a hung collective must not hold the box until the caller's own limit
        import faulthandler

        faulthandler.dump_traceback_later(args.watchdog_seconds, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
