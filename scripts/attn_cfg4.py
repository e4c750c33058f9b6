"""BASELINE.json config 4 only (4096 windows x 343 tokens x 96 ch x 3 heads): a short ncu target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
N.set_gemm_mode(1)
Bw, C, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 96, 3
qkv = torch.randn(Bw * 343, 3 * C, device="cuda")
for _ in range(3):
    o, lse = ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.window_attn_fwd(qkv, C, heads, Bw, (7, 7, 7), (7, 7, 7))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"config4 x{Bw}: {ms:.3f} ms  {4.0 * Bw * heads * 343 * 343 * 32 / ms / 1e9:.1f} TFLOP/s")
if len(sys.argv) > 2 and sys.argv[2] == "bwd":          # the same shape's tcgen05 backward
    do = torch.randn(Bw * 343, C, device="cuda")
    for _ in range(2): ops.window_attn_bwd(qkv, o, do, lse, C, heads, Bw, (7, 7, 7), (7, 7, 7))
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.window_attn_bwd(qkv, o, do, lse, C, heads, Bw, (7, 7, 7), (7, 7, 7))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"config4 backward x{Bw}: {ms:.3f} ms  {10.0 * Bw * heads * 343 * 343 * 32 / ms / 1e9:.1f} TFLOP/s (algorithmic, 5 products)")
