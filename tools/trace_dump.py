"""Prints the kernel timeline (start, duration, stream, grid, name) of gpurun_out/graph_trace.json
written by tools/graph_timeline.py."""
import json
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/graph_trace.json"
d = json.load(open(path))
ev = [e for e in d["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
lo = float(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
for e in ev:
    n = e["name"]
    m = re.search(r"(gemm_f64_kernel<[^>]*>|\w+_kernel(<\d+>)?)", n)
    nm = m.group(1) if m else n[:50]
    a = e.get("args", {})
    if lo <= e["ts"] - t0 <= hi:
        print(f"{e['ts'] - t0:8.1f} {e['dur']:7.1f} s{a.get('stream', '?'):>3} g{str(a.get('grid', '')):<14} {nm[:60]}")
