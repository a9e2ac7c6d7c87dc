"""Kernel time by name for one eager step of a bench.py workload (torch.profiler / CUPTI; no graph).
Usage: python tools/c3_timeline.py [c3|c2] -> gpurun_out/timeline_<cfg>.txt"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
model, X, Y, num_data = bench.build_workload(cfg, 8192, 0)
model = model.cuda()
Xd, Yd = torch.from_numpy(X[:8192]).cuda(), torch.from_numpy(Y[:8192]).cuda()


def step():
    for p in model.parameters():
        p.grad = None
    loss = model.loss(Xd, Yd)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step()
    b.record()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
spans = []  # (start_us, end_us, short name)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name
        m = re.search(r"(gemm_f64\w*_kernel<[^>]*>|\w+_kernel(<[^>]*>)?)", n)
        nm = m.group(1) if m else n[:60]
        agg[nm][0] += 1
        dur = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        agg[nm][1] += dur
        spans.append((e.time_range.start, e.time_range.start + dur, nm))
tot = sum(v[1] for v in agg.values())
lines = [f"{cfg}: one eager step = {a.elapsed_time(b):.2f} ms (under CUPTI); summed kernel time {tot / 1e3:.2f} ms"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    lines.append(f"{v[1] / 1e3:9.3f} ms {v[0]:5d}  {k}")
# exposed time: what runs while none of the large GEMM launches (> 2 ms) is on the device
big = sorted((a0, b0) for a0, b0, _ in spans if b0 - a0 > 2000.0)
merged = []
for a0, b0 in big:
    if merged and a0 <= merged[-1][1]:
        merged[-1][1] = max(merged[-1][1], b0)
    else:
        merged.append([a0, b0])
t0, t1 = min(s0 for s0, _, _ in spans), max(e0 for _, e0, _ in spans)
covered = sum(b0 - a0 for a0, b0 in merged)
lines.append(f"large GEMM launches cover {covered / 1e3:.2f} ms of the {(t1 - t0) / 1e3:.2f} ms between first and last kernel; "
             f"exposed (no large GEMM running): {(t1 - t0 - covered) / 1e3:.2f} ms, by kernel:")
exposed = collections.defaultdict(float)
for a0, b0, nm in spans:
    if b0 - a0 > 2000.0:
        continue
    vis = b0 - a0
    for c0, c1 in merged:
        lo, hi = max(a0, c0), min(b0, c1)
        if hi > lo:
            vis -= hi - lo
    exposed[nm] += max(vis, 0.0)
for k, v in sorted(exposed.items(), key=lambda kv: -kv[1])[:25]:
    lines.append(f"{v / 1e3:9.3f} ms exposed  {k}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"timeline_{cfg}.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
