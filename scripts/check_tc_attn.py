import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
dev = "cuda"
def ref_attn(qkv, C, heads, B, pd, ws):
    from oracle import micformer_oracle as O
    hd = C // heads
    g = qkv.view(B, *pd, 3 * C).double()
    q = O.window_partition(g[..., :C].contiguous(), ws); k = O.window_partition(g[..., C:2*C].contiguous(), ws); v = O.window_partition(g[..., 2*C:].contiguous(), ws)
    Bw, Nt, _ = q.shape
    sp = lambda t: t.view(Bw, Nt, heads, hd).permute(0, 2, 1, 3)
    a = ((sp(q) * hd ** -0.5) @ sp(k).transpose(-2, -1)).softmax(-1)
    o = (a @ sp(v)).transpose(1, 2).reshape(Bw, Nt, C)
    return O.window_reverse(o, ws, B, *pd).reshape(-1, C)
for (B, pd, ws, C, heads) in [(1, (7, 7, 7), (7, 7, 7), 96, 3), (2, (7, 14, 14), (7, 7, 7), 96, 3), (2, (8, 8, 8), (4, 8, 8), 64, 2), (1, (21, 21, 21), (7, 7, 7), 192, 6)]:
    P = B * pd[0] * pd[1] * pd[2]
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(P, 3 * C, generator=g)
    ref = ref_attn(qkv, C, heads, B, pd, ws)
    qd = qkv.to(dev)
    res = {}
    for mode in (0, 1):
        N.set_gemm_mode(mode)
        o, lse = ops.window_attn_fwd(qd, C, heads, B, pd, ws)
        torch.cuda.synchronize()
        res[mode] = (float((o.cpu().double() - ref).abs().max() / ref.abs().max()), lse.cpu())
    print(f"B{B} grid{pd} win{ws} C{C} h{heads}: simt err {res[0][0]:.2e}  tc err {res[1][0]:.2e}  lse diff {float((res[0][1]-res[1][1]).abs().max()):.2e}", flush=True)
# BASELINE.json config 4: 4096 windows x 343 tokens x 96 ch x 3 heads (pre-partitioned windows = a (4096,7,7,7) grid)
Bw, C, heads = 4096, 96, 3
qkv = torch.randn(Bw * 343, 3 * C, device=dev)
flops = 4.0 * Bw * heads * 343 * 343 * 32
for mode in (1, 0):
    N.set_gemm_mode(mode)
    for _ in range(2): ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"config4 mode{mode}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  ({4 * Bw * 343 * C * 4 / ms / 1e6:.0f} GB/s algorithmic Q+K+V+O)")
