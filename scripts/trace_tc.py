import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
lib = N.load(); lib.mic_debug_tc_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"; N.set_gemm_mode(1)
for (M, Nn, K) in [(1024, 192, 192), (1024, 192, 768), (1024, 768, 192), (8192, 96, 384), (128, 384, 1536)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(Nn, K, device=dev); b = torch.randn(Nn, device=dev)
    for _ in range(3): ops.linear_fwd(x, K, w, b, M, Nn, K)
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
    lib.mic_debug_tc_trace(buf.data_ptr())
    torch.cuda.synchronize()
    ops.linear_fwd(x, K, w, b, M, Nn, K)
    torch.cuda.synchronize()
    lib.mic_debug_tc_trace(None)
    t = buf.view(4096, 16).cpu()
    n = int((t[:, 0] > 0).sum())
    t = t[:n].double()
    t0 = t[:, 0].min()
    print(f"--- M{M} N{Nn} K{K}: {n} CTAs, kernel span {float(t[:, 9].max() - t0)/1e3:.1f} us")
    names = ["start", "alloc+sync", "tma0 issued", "tma all issued", "mma first full", "mma committed", "epi start", "epi done", "final sync", "dealloc", "first tmem_ld"]
    for cta in (0, 1, n // 2, n - 1):
        rel = [(float(t[cta, i] - t[cta, 0]) / 1e3) for i in range(11)]
        print(f"cta {cta}: start@{float(t[cta,0]-t0)/1e3:.1f}us  " + "  ".join(f"{nm}={v:.2f}" for nm, v in zip(names[1:], rel[1:])))
    d = (t[:, 1:11] - t[:, 0:1]) / 1e3
    print("mean us since start:", "  ".join(f"{nm}={float(d[:, i].mean()):.2f}" for i, nm in enumerate(names[1:])))
