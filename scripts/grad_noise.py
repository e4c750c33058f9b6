"""Run-to-run and mode-to-mode gradient differences per parameter tensor (TINY config, 64^3): which tensors carry the
tensor-core mode's gradient error, and how much of it is nondeterministic (atomics)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import _native as N
from micformer_b200.models.MICFormer_self import Head, MicFormer
from micformer_b200.loss.dice import MDiceLoss
from oracle import micformer_oracle as O

cfg = O.TINY
sd = O.synth_state_dict(cfg, seed=3)
head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
head.swin = MicFormer(window_size=cfg.window_size, in_chans=1, embed_dim=cfg.embed_dim, depths=list(cfg.depths), num_heads=list(cfg.num_heads))
head.load_state_dict(sd, strict=True)
head = head.cuda().eval()
x, lab = O.synth_inputs(1, 64, cfg.num_classes, seed=5)
x, lab = x.cuda(), lab.cuda()


def grads(mode):
    N.set_gemm_mode(mode)
    for p in head.parameters():
        p.grad = None
    MDiceLoss()(head(x), lab).backward()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in head.named_parameters() if p.grad is not None}


g0 = grads(0)
g1a, g1b = grads(1), grads(1)
g0b = grads(0)
gl2 = float(sum((g.double() ** 2).sum() for g in g0.values()) ** 0.5)
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-5 * gl2))
rows = [(rel(g1a[k], g0[k]), rel(g1a[k], g1b[k]), rel(g0b[k], g0[k]), k) for k in g0]
rows.sort(reverse=True)
print("worst mode1-vs-mode0 | run-to-run mode1 | run-to-run mode0 | tensor")
for r in rows[:25]:
    print(f"{r[0]:.3e}  {r[1]:.3e}  {r[2]:.3e}  {r[3]}")
rows.sort(key=lambda r: -r[1])
print("worst run-to-run (mode 1):")
for r in rows[:8]:
    print(f"{r[0]:.3e}  {r[1]:.3e}  {r[2]:.3e}  {r[3]}")
