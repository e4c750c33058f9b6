"""Which tensor-core-mode kernels are nondeterministic beyond fp32 reordering?  Each op runs twice on identical inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, fused, _native as N
N.set_gemm_mode(1)
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
R = lambda *s: torch.randn(*s, generator=g).to(dev)


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def twice(name, fn):
    a = [t.clone() for t in fn()]
    worst = 0.0
    for _ in range(4):
        b = fn()
        worst = max(worst, max(rel(x, y) for x, y in zip(b, a)))
    print(f"{name:58s} run-to-run {worst:.3e}")


for (M, Nn, K) in [(4096, 24, 24), (4096, 96, 24), (4096, 24, 96), (65536, 48, 192), (1024, 192, 768), (128, 384, 1536)]:
    X, W, b, dY = R(M, K), R(Nn, K) * 0.1, R(Nn), R(M, Nn)
    twice(f"linear_fwd M{M} N{Nn} K{K}", lambda: [ops.linear_fwd(X, K, W, b, M, Nn, K)])
    twice(f"linear_bwd_data M{M} N{Nn} K{K}", lambda: [ops.linear_bwd_data(dY, Nn, W, M, Nn, K)])
    twice(f"linear_bwd_weight M{M} N{Nn} K{K}", lambda: list(ops.linear_bwd_weight(dY, Nn, X, K, M, Nn, K)))
for (B, S, C0, C1, Co) in [(1, 16, 24, 24, 16), (2, 32, 48, 48, 16), (1, 8, 96, 96, 16)]:
    x0, x1 = R(B, S, S, S, C0), R(B, S, S, S, C1)
    cw = R(27, C0 + C1, Co) * 0.05
    cwk = cw.permute(0, 2, 1).contiguous()
    cb = R(Co)
    dy = R(B, S, S, S, Co)
    def f():
        out = torch.empty(B * S ** 3, Co, device=dev)
        ops.conv3_fwd(x0, x1, cw, cwk, cb, out, B, (S, S, S), Co, False)
        return [out]
    twice(f"conv3 fwd B{B} S{S} Cin{C0 + C1}", f)
    def fb():
        d0, d1 = torch.zeros_like(x0), torch.zeros_like(x1)
        ops.conv3_bwd_data(dy, cw, d0, True, d1, True, B, (S, S, S), Co, False)
        return [d0, d1]
    twice(f"conv3 bwd_data B{B} S{S} Cin{C0 + C1}", fb)
    def fw():
        dw, db = torch.zeros_like(cw), torch.zeros_like(cb)
        ops.conv3_bwd_weight(dy, x0, x1, dw, db, B, (S, S, S), Co, False)
        return [dw, db]
    twice(f"conv3 bwd_weight B{B} S{S} Cin{C0 + C1}", fw)
