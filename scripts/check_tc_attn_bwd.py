"""tcgen05 window-attention backward (csrc/window_attn_tc_bwd.cu) vs fp64 autograd and vs the exact CUDA-core kernel,
then its time on BASELINE config 4 (4096 windows x 343 tokens x 96 ch x 3 heads) beside the CUDA-core kernel's."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
from oracle import micformer_oracle as O
dev = "cuda"


def ref(qkv, do, C, heads, B, pd, ws):
    hd = C // heads
    x = qkv.double().requires_grad_(True)
    g = x.view(B, *pd, 3 * C)
    q, k, v = (O.window_partition(g[..., i * C:(i + 1) * C].contiguous(), ws) for i in range(3))
    Bw, Nt, _ = q.shape
    sp = lambda t: t.view(Bw, Nt, heads, hd).permute(0, 2, 1, 3)
    o = (((sp(q) * hd ** -0.5) @ sp(k).transpose(-2, -1)).softmax(-1) @ sp(v)).transpose(1, 2).reshape(Bw, Nt, C)
    o = O.window_reverse(o, ws, B, *pd).reshape(-1, C)
    (o * do.double()).sum().backward()
    return x.grad


cases = [(1, (7, 7, 7), (7, 7, 7), 96, 3), (2, (7, 14, 14), (7, 7, 7), 96, 3), (2, (8, 8, 8), (4, 8, 8), 64, 2),
         (1, (6, 6, 6), (6, 6, 6), 32, 1), (1, (5, 5, 6), (5, 5, 6), 64, 2), (1, (14, 14, 14), (7, 7, 7), 192, 6)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cases = cases[:2]
for (B, pd, ws, C, heads) in cases:
    P = B * pd[0] * pd[1] * pd[2]
    gen = torch.Generator().manual_seed(0)
    qkv = torch.randn(P, 3 * C, generator=gen)
    do = torch.randn(P, C, generator=gen)
    gref = ref(qkv, do, C, heads, B, pd, ws)
    qd, dd = qkv.to(dev), do.to(dev)
    out = {}
    for mode in (0, 1):
        N.set_gemm_mode(mode)
        o, lse = ops.window_attn_fwd(qd, C, heads, B, pd, ws)
        d = ops.window_attn_bwd(qd, o, dd, lse, C, heads, B, pd, ws)
        torch.cuda.synchronize()
        out[mode] = d.cpu().double()
    errs = []
    for mode in (0, 1):
        e = [float((out[mode][:, i * C:(i + 1) * C] - gref[:, i * C:(i + 1) * C]).abs().max() / gref[:, i * C:(i + 1) * C].abs().max())
             for i in range(3)]
        errs.append(e)
    print(f"B{B} grid{pd} win{ws} C{C} h{heads}: simt dq/dk/dv {errs[0][0]:.1e} {errs[0][1]:.1e} {errs[0][2]:.1e}   "
          f"tc dq/dk/dv {errs[1][0]:.1e} {errs[1][1]:.1e} {errs[1][2]:.1e}  nan {int(torch.isnan(out[1]).sum())}", flush=True)

Bw, C, heads = 4096, 96, 3
qkv = torch.randn(Bw * 343, 3 * C, device=dev)
do = torch.randn(Bw * 343, C, device=dev)
flops = 2.0 * 5 * Bw * heads * 343 * 343 * 32          # five 343 x 343 x 32 products (the kernel runs eight: both orientations)
for mode in (1, 0):
    N.set_gemm_mode(mode)
    o, lse = ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
    for _ in range(2): ops.window_attn_bwd(qkv, o, do, lse, C, heads, Bw, (7, 7, 7), (7, 7, 7))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 5 if mode == 1 else 2
    for _ in range(n): ops.window_attn_bwd(qkv, o, do, lse, C, heads, Bw, (7, 7, 7), (7, 7, 7))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"config4 backward mode{mode}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s algorithmic  "
          f"({(3 + 2 + 3) * Bw * 343 * C * 4 / ms / 1e6:.0f} GB/s algorithmic q,k,v,o,do in + dq,dk,dv out)", flush=True)
