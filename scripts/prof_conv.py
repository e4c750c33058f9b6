"""Event-time (and give ncu a short target for) the 3x3x3 conv kernels at the model's shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
dev = "cuda"
N.set_gemm_mode(1)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
shapes = [("s0 conv_offset", 2, (32, 32, 32), 48, 48, 16, False), ("s1 conv_offset", 2, (16, 16, 16), 96, 96, 16, False),
          ("out_conv", 2, (128, 128, 128), 24, 0, 8, True)]
def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / max(reps, 1) * 1e3 + 1e-9
for name, B, dims, C0, C1, Co, ncdhw in shapes:
    D, H, W = dims
    P = B * D * H * W
    Cin = C0 + C1
    x0 = torch.randn(B, D, H, W, C0, device=dev); x1 = torch.randn(B, D, H, W, C1, device=dev) if C1 else None
    wt = torch.randn(27, Cin, Co, device=dev) * 0.1; wk = wt.permute(0, 2, 1).contiguous(); bias = torch.randn(Co, device=dev)
    y = torch.empty((B, Co, D, H, W) if ncdhw else (B, D, H, W, Co), device=dev)
    dy = torch.randn_like(y)
    dx0 = torch.zeros(B, D, H, W, C0, device=dev); dx1 = torch.zeros(B, D, H, W, C1, device=dev) if C1 else None
    dwt = torch.zeros_like(wt); db = torch.zeros(Co, device=dev)
    flops = 2.0 * 27 * Cin * Co * P
    t_f = timeit(lambda: ops.conv3_fwd(x0, x1, wt, wk, bias, y, B, dims, Co, ncdhw))
    t_d = timeit(lambda: ops.conv3_bwd_data(dy, wt, dx0, True, dx1, True, B, dims, Co, ncdhw))
    t_w = timeit(lambda: ops.conv3_bwd_weight(dy, x0, x1, dwt, db, B, dims, Co, ncdhw))
    by = 4.0 * (P * Cin + P * Co)
    print(f"{name}: fwd {t_f:.1f} us ({flops/t_f/1e6:.1f} TF, {by/t_f/1e3:.0f} GB/s)  bwd_data {t_d:.1f} us ({flops/t_d/1e6:.1f} TF)  "
          f"bwd_weight {t_w:.1f} us ({flops/t_w/1e6:.1f} TF)", flush=True)
