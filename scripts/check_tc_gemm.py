"""Diagnostic: tcgen05 TF32 GEMM (gemm mode 1) vs fp64 reference for all three operand-layout combinations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N

torch.manual_seed(0)
dev = "cuda"
def rel(a, b): return float((a.double().cpu() - b).abs().max() / b.abs().max())
cases = [(256, 48, 32), (300, 144, 48), (1000, 192, 48), (128, 384, 1536), (65536, 48, 192), (1024, 768, 192), (517, 96, 24), (64, 1536, 96), (4096, 96, 96)]
for (M, Nn, K) in cases:
    for kn in (False, True):
        x = torch.randn(M, K); w = (torch.randn(K, Nn) if kn else torch.randn(Nn, K)) * 0.1; b = torch.randn(Nn); dy = torch.randn(M, Nn)
        wm = (w if kn else w.t()).double()
        xd, wd, bd, dyd = x.to(dev), w.to(dev), b.to(dev), dy.to(dev)
        out = {}
        for mode in (0, 1):
            N.set_gemm_mode(mode)
            y = ops.linear_fwd(xd, K, wd, bd, M, Nn, K, w_is_kn=kn)
            dx = ops.linear_bwd_data(dyd, Nn, wd, M, Nn, K, w_is_kn=kn)
            dW, db = ops.linear_bwd_weight(dyd, Nn, xd, K, M, Nn, K, w_is_kn=kn)
            torch.cuda.synchronize()
            out[mode] = (rel(y, x.double() @ wm + b.double()), rel(dx, dy.double() @ wm.t()),
                         rel(dW, (x.double().t() @ dy.double()) if kn else (dy.double().t() @ x.double())))
        print(f"M{M} N{Nn} K{K} kn={int(kn)}  simt fwd/dx/dw {out[0][0]:.1e} {out[0][1]:.1e} {out[0][2]:.1e} | tc {out[1][0]:.1e} {out[1][1]:.1e} {out[1][2]:.1e}", flush=True)
# epilogues in TC mode
N.set_gemm_mode(1)
M, Nn, K = 2048, 192, 48
x = torch.randn(M, K); w = torch.randn(Nn, K) * 0.1; b = torch.randn(Nn); res = torch.randn(M, Nn)
xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
pre = torch.empty(M, Nn, device=dev)
yg = ops.linear_fwd(xd, K, wd, bd, M, Nn, K, act=True, pre=pre)
ref = x.double() @ w.double().t() + b.double()
print("gelu epi", rel(pre, ref), rel(yg, torch.nn.functional.gelu(ref)))
rs = torch.tensor([0.5, 2.0])
yr = ops.linear_fwd(xd, K, wd, bd, M, Nn, K, res=res.to(dev), rowscale=rs.to(dev), rps=M // 2)
print("res+rowscale epi", rel(yr, res.double() + ref * rs.double().repeat_interleave(M // 2)[:, None]))
dy = torch.randn(M, Nn); dyd = dy.to(dev)
dW, db = ops.linear_bwd_weight(dyd, Nn, xd, K, M, Nn, K, rowscale=rs.to(dev), rps=M // 2)
sdy = dy.double() * rs.double().repeat_interleave(M // 2)[:, None]
print("bwd_weight rowscale", rel(dW, sdy.t() @ x.double()), rel(db, sdy.sum(0)))
# timing
import time
for (M, Nn, K) in [(65536, 192, 48), (65536, 48, 192), (1024, 192, 192), (1024, 768, 192), (8192, 384, 96)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(Nn, K, device=dev); b = torch.randn(Nn, device=dev); dy = torch.randn(M, Nn, device=dev)
    for mode in (0, 1):
        N.set_gemm_mode(mode)
        res = []
        for fn in (lambda: ops.linear_fwd(x, K, w, b, M, Nn, K), lambda: ops.linear_bwd_data(dy, Nn, w, M, Nn, K), lambda: ops.linear_bwd_weight(dy, Nn, x, K, M, Nn, K)):
            for _ in range(3): fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 20 * 1e3)
        print(f"time M{M} N{Nn} K{K} mode{mode}: fwd {res[0]:.1f} us  bwd_data {res[1]:.1f} us  bwd_weight(+colsum) {res[2]:.1f} us")
