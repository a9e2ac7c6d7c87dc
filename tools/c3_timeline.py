"""Kernel time by name for one eager c3-shard step (torch.profiler / CUPTI; no graph)."""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "--steps", "1"]
import runpy

from torch.profiler import ProfilerActivity, profile

mod = runpy.run_path(os.path.join(os.path.dirname(__file__), "bench_c3.py"), run_name="c3mod")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    mod["main"]()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name
        m = re.search(r"(gemm_f64_kernel<[^>]*>|\w+_kernel(<\d+>)?)", n)
        nm = m.group(1) if m else n[:50]
        agg[nm][0] += 1
        agg[nm][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1] / 1e3:9.2f} ms {v[0]:5d}  {k}")
print(f"total kernel time {tot / 1e3:.1f} ms (two steps: warm-up + timed)")
