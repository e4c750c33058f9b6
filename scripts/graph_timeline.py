"""Kernel timeline of one replay of the captured training-step graph (the bench workload), from torch.profiler (CUPTI):
per-kernel (name, stream, start, duration) -> CSV, plus a summary: span, per-stream busy time, how much of the span has
1 / 2 / 3+ kernels in flight, idle gaps, and the kernels that cover most of the span while running ALONE.

    python scripts/graph_timeline.py [--size 128] [--out gpurun_out/timeline.csv]"""
import argparse
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("mic::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\(.*$", "", name)
    return name[:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--out", default="gpurun_out/timeline.csv")
    a = ap.parse_args()
    from micformer_b200 import _native
    from micformer_b200.models.MICFormer_self import Head
    from micformer_b200.loss.dice import MDiceLoss
    from micformer_b200.optim import FusedAdam
    _native.set_gemm_mode(1)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = Head(embed_dim=48, num_classes=8, window_size=(2, 2, 2)).to(dev).train()
    crit = MDiceLoss()
    opt = FusedAdam(model.parameters(), lr=1e-4, weight_decay=0.0)
    x = torch.randn(2, 2, a.size, a.size, a.size, device=dev)
    lab = (torch.rand(2, 8, a.size, a.size, a.size, device=dev) > 0.5).float()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), lab)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    opt.zero_grad(set_to_none=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = crit(model(x), lab)
        loss.backward()
        opt.step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        graph.replay()
        torch.cuda.synchronize()
    evs = []
    try:
        for ke in prof.profiler.kineto_results.events():
            if ke.device_type() != torch.autograd.DeviceType.CUDA:
                continue
            nm = ke.name()
            if nm.startswith("Memcpy") or nm.startswith("Memset"):
                nm = nm.split(" ")[0]
            s0 = ke.start_ns() / 1e3
            evs.append((s0, s0 + ke.duration_ns() / 1e3, short(nm), ke.device_resource_id()))
    except Exception as ex:  # older / newer profiler API: fall back to the FunctionEvent view
        print("kineto_results path failed:", ex)
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range is not None:
                evs.append((ev.time_range.start, ev.time_range.end, short(ev.name), getattr(ev, "device_resource_id", -1)))
    evs.sort()
    if not evs:
        print("no CUDA events recorded")
        return
    t0 = evs[0][0]
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["start_us", "dur_us", "stream", "kernel"])
        for s, e, n, st in evs:
            w.writerow([f"{s - t0:.2f}", f"{e - s:.2f}", st, n])
    span = max(e for _, e, _, _ in evs) - t0
    print(f"{len(evs)} device activities, span {span / 1e3:.3f} ms")
    busy = collections.Counter()
    for s, e, n, st in evs:
        busy[st] += e - s
    for st, b in busy.most_common():
        print(f"  stream {st}: busy {b / 1e3:.3f} ms ({sum(1 for v in evs if v[3] == st)} activities)")
    # concurrency sweep
    pts = []
    for i, (s, e, n, st) in enumerate(evs):
        pts.append((s, 1, i))
        pts.append((e, -1, i))
    pts.sort(key=lambda p: (p[0], p[1]))
    level = collections.Counter()
    alone = collections.Counter()
    active = set()
    last = pts[0][0]
    for t, d, i in pts:
        dt = t - last
        if dt > 0:
            level[min(len(active), 4)] += dt
            if len(active) == 1:
                alone[evs[next(iter(active))][2]] += dt
        last = t
        if d > 0:
            active.add(i)
        else:
            active.discard(i)
    for k in sorted(level):
        print(f"  {k}{'+' if k == 4 else ' '} kernels in flight: {level[k] / 1e3:.3f} ms ({100 * level[k] / span:.1f} %)")
    print("  time with exactly ONE kernel in flight, by kernel:")
    for n, t in alone.most_common(25):
        print(f"    {t / 1e3:7.3f} ms  {n}")
    tot = collections.Counter()
    cnt = collections.Counter()
    for s, e, n, st in evs:
        tot[n] += e - s
        cnt[n] += 1
    print("  device time by kernel (in-graph, overlapped):")
    for n, t in tot.most_common(30):
        print(f"    {t / 1e3:7.3f} ms  {cnt[n]:5d} x {t / cnt[n]:7.1f} us  {n}")


if __name__ == "__main__":
    main()
