"""Kernel timeline of one CUDA-graph replay of the c2 training step (torch.profiler / CUPTI):
span, busy time, idle gaps and the top kernels as they actually run inside the graph (ncu serialises
launches and adds cold-cache overhead, so it cannot show this).  Writes gpurun_out/graph_timeline.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    X, Y = bench.make_data()
    model = bench.build_model(X).to(dev)
    if "--no-overlap" in sys.argv:
        model.overlap_factorization = False
    B = bench.BATCH
    sX = torch.from_numpy(X[:B]).to(dev)
    sY = torch.from_numpy(Y[:B]).to(dev)

    def fwd_bwd():
        for p in model.parameters():
            p.grad = None
        loss = model.loss(sX, sY)
        loss.backward()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fwd_bwd()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd_bwd()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        g.replay()
        torch.cuda.synchronize()
    os.makedirs("gpurun_out", exist_ok=True)
    prof.export_chrome_trace("gpurun_out/graph_trace.json")
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    if not ks:
        print("no kernel events captured")
        return
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    # union of busy intervals
    busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
    for s_, e_, _ in ks[1:]:
        if s_ > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    agg = {}
    for s_, e_, n in ks:
        n = n.split("(")[0][-60:]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += e_ - s_
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]
    out = {"kernels": len(ks), "span_us": t1 - t0, "busy_us": busy, "idle_us": (t1 - t0) - busy,
           "sum_kernel_us": sum(e_ - s_ for s_, e_, _ in ks),
           "top": [{"name": n, "count": c, "us": u} for n, (c, u) in top]}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/graph_timeline.json", "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "top"}))
    for t in out["top"]:
        print(f"{t['us']:9.1f} us {t['count']:4d}  {t['name']}")


if __name__ == "__main__":
    main()
