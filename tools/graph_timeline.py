"""Kernel timeline of one CUDA-graph replay of a bench.py training step (torch.profiler / CUPTI): every
kernel in start order with its duration and stream, then span, busy time, idle gaps and the top kernels as
they actually run inside the graph (ncu serialises launches and adds cold-cache overhead, so it cannot show
this).  Usage: python tools/graph_timeline.py [c2|c3] [--no-overlap] -> gpurun_out/graph_timeline_<cfg>.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    cfg = next((a for a in sys.argv[1:] if not a.startswith("-")), "c2")
    dev = torch.device("cuda", 0)
    model, X, Y, _ = bench.build_workload(cfg, 8192, 0)
    model = model.to(dev)
    if "--no-overlap" in sys.argv:
        model.overlap_factorization = False
    sX = torch.from_numpy(X[:8192]).to(dev)
    sY = torch.from_numpy(Y[:8192]).to(dev)

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

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", -1)) for e in evs),
                key=lambda t: t[0])
    if not ks:
        print("no kernel events captured")
        return
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    lines = []
    for s_, e_, n, sid in ks:
        short = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("gpfx::", "").replace("at::native::", "")
        short = short.split("(")[0][:70]
        lines.append(f"{s_ - t0:9.1f} {e_ - s_:8.1f}  s{sid:<4} {short}")
    busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
    for s_, e_, _, _ in ks[1:]:
        if s_ > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    agg = {}
    for s_, e_, n, _ in ks:
        n = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("gpfx::", "").replace("at::native::", "").split("(")[0][:70]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += e_ - s_
    lines.append(f"# {cfg}: {len(ks)} kernels, span {t1 - t0:.1f} us, busy {busy:.1f} us, idle {(t1 - t0) - busy:.1f} us, "
                 f"summed kernel time {sum(e_ - s_ for s_, e_, _, _ in ks):.1f} us")
    for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        lines.append(f"# {u:9.1f} us {c:4d}  {n}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"graph_timeline_{cfg}.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-27:]))


if __name__ == "__main__":
    main()
