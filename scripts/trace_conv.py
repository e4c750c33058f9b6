import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
lib = N.load(); lib.mic_debug_conv_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"; N.set_gemm_mode(1)
B, dims, C0, C1, Co = 2, (32, 32, 32), 48, 48, 16
D, H, W = dims
x0 = torch.randn(B, D, H, W, C0, device=dev); x1 = torch.randn(B, D, H, W, C1, device=dev)
wt = torch.randn(27, C0 + C1, Co, device=dev) * 0.1; wk = wt.permute(0, 2, 1).contiguous(); bias = torch.randn(Co, device=dev)
y = torch.empty(B, D, H, W, Co, device=dev); dy = torch.randn_like(y)
dx0 = torch.zeros_like(x0); dx1 = torch.zeros_like(x1)
def run(tag, fn):
    for _ in range(3): fn()
    buf = torch.zeros(6 * 256, dtype=torch.int64, device=dev)
    lib.mic_debug_conv_trace(buf.data_ptr()); torch.cuda.synchronize()
    fn(); torch.cuda.synchronize(); lib.mic_debug_conv_trace(None)
    t = buf.view(6, 256).cpu().double()
    t0 = t[0, 0]
    rel = lambda v: (v - t0) / 1e3
    n = int((t[2] > 0).sum()); m = int((t[4] > 0).sum())
    print(f"--- {tag}: {n} plane-chunks, {m} output plane-chunks; epilogue wait {rel(t[5,0]):.2f} -> {rel(t[5,1]):.2f}, end {rel(t[5,2]):.2f} us")
    print("plane-chunk: issued / landed / published (us)")
    print("  ".join(f"{i}:{rel(t[0,i]):.1f}/{rel(t[1,i]):.1f}/{rel(t[2,i]):.1f}" for i in range(n)))
    print("mma: ready / issued (us)")
    print("  ".join(f"{i}:{rel(t[3,i]):.1f}/{rel(t[4,i]):.1f}" for i in range(m)))
run("fwd s0", lambda: ops.conv3_fwd(x0, x1, wt, wk, bias, y, B, dims, Co, False))
run("bwd_data s0", lambda: ops.conv3_bwd_data(dy, wt, dx0, True, dx1, True, B, dims, Co, False))
