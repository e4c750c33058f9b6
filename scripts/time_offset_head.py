"""mic_offset_head_bwd timing vs grid size / data (the window-7 step shows 0.45 ms per call, the train config 9 us)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import _native as N
dev = "cuda"
def run(B, D, H, W, scale=1.0, tag=""):
    P = B * D * H * W
    h = torch.randn(P, 16, device=dev) * scale; dpos = torch.randn(P, 3, device=dev)
    g = torch.ones(16, device=dev); b = torch.zeros(16, device=dev); w3 = torch.randn(3, 16, device=dev) * 0.1
    dh = torch.empty(P, 16, device=dev); dg = torch.zeros(16, device=dev); db = torch.zeros(16, device=dev); dw = torch.zeros(3, 16, device=dev)
    f = lambda: N.call("mic_offset_head_bwd", N.ptr(dpos), N.ptr(h), N.ptr(g), N.ptr(b), N.ptr(w3), N.ptr(dh), N.ptr(dg), N.ptr(db), N.ptr(dw), B, D, H, W, 16, 1e-5)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"P={P:8d} ({B},{D},{H},{W}) scale {scale:g} {tag}: {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us", flush=True)
run(2, 32, 32, 32); run(2, 35, 35, 35); run(2, 35, 35, 35, 1e-3); run(2, 35, 35, 35, 1e3); run(2, 21, 21, 21); run(2, 14, 14, 14); run(2, 7, 7, 7)
run(2, 40, 40, 40); run(2, 64, 64, 64)
