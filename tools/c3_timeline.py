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
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name
        m = re.search(r"(gemm_f64_kernel<[^>]*>|\w+_kernel(<[^>]*>)?)", n)
        nm = m.group(1) if m else n[:60]
        agg[nm][0] += 1
        agg[nm][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
lines = [f"{cfg}: one eager step = {a.elapsed_time(b):.2f} ms (under CUPTI); summed kernel time {tot / 1e3:.2f} ms"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    lines.append(f"{v[1] / 1e3:9.3f} ms {v[0]:5d}  {k}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"timeline_{cfg}.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
