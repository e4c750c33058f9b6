"""Which lines of this package launch torch's own elementwise kernels (fill / copy / add ...) inside one training step?
One eager step of the bench workload under torch.profiler with Python stacks; prints, per aten op that launched a CUDA
kernel, the call sites inside micformer_b200/ ranked by launch count.

    python scripts/plumbing_audit.py [--size 128] [--arena 0]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--arena", type=int, default=0)
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
    if a.arena:
        from micformer_b200.arena import GradArena
        arena = GradArena.for_model(model)
        opt.attach_arena(arena)
    x = torch.randn(2, 2, a.size, a.size, a.size, device=dev)
    lab = (torch.rand(2, 8, a.size, a.size, a.size, device=dev) > 0.5).float()

    def step():
        opt.zero_grad(set_to_none=not a.arena)
        if a.arena:
            arena.zero()
        loss = crit(model(x), lab)
        loss.backward()
        opt.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    # cpu op events that own at least one CUDA kernel, keyed by (op name, first frame inside this repository)
    sites = collections.Counter()
    shapes = collections.Counter()
    kern = collections.Counter()
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith("aten::"):
            continue
        nk = sum(1 for k in ev.kernels)
        if nk == 0:
            continue
        # only leaf aten ops (children of other aten ops would double count): the profiler attaches kernels to the leaf
        frame = next((s for s in ev.stack if "micformer_b200" in s or "bench.py" in s or "scripts/" in s), "(autograd engine / no python frame)")
        sites[(ev.name, frame.strip())] += nk
        kern[ev.name] += nk
        shapes[(ev.name, str(ev.input_shapes)[:90], (ev.stack[0].strip()[-70:] if ev.stack else ""))] += nk
    print("aten ops that launched kernels in one step:", dict(kern))
    for (name, frame), n in sites.most_common(60):
        print(f"{n:5d}  {name:28s} {frame}")
    print("by input shapes (and innermost recorded frame):")
    for (name, shp, fr), n in shapes.most_common(70):
        print(f"{n:5d}  {name:14s} {shp:90s} {fr}")


if __name__ == "__main__":
    main()
