import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
lib = N.load(); lib.mic_debug_tc_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"; N.set_gemm_mode(1)
names = ["tma0", "mma_gotbuf", "mma_first_ready", "mma_commit", "epi_wait", "epi_got", "epi_done"]
def run(tag, fn):
    for _ in range(3): fn()
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
    lib.mic_debug_tc_trace(buf.data_ptr()); torch.cuda.synchronize()
    fn(); torch.cuda.synchronize(); lib.mic_debug_tc_trace(None)
    t = buf.view(-1, 8).cpu().double()
    n = int((t[:, 1] > 0).sum())
    t0 = t[0, 0]
    print(f"--- {tag}: CTA 0 ran {n} tiles")
    for i in list(range(min(n, 5))) + ([n - 1] if n > 5 else []):
        print(f"tile {i}: " + "  ".join(f"{nm}={(float(t[i, k] - t0)) / 1e3:7.2f}" for k, nm in enumerate(names)))
M = 65536
for (Nn, K, kind) in [(192, 48, "gelu"), (48, 48, "res"), (48, 192, "res"), (96, 48, "plain")]:
    x = torch.randn(M, K, device=dev); w = torch.randn(Nn, K, device=dev); b = torch.randn(Nn, device=dev)
    res = torch.randn(M, Nn, device=dev); pre = torch.empty(M, Nn, device=dev)
    if kind == "gelu": fn = lambda: ops.linear_fwd(x, K, w, b, M, Nn, K, act=True, pre=pre)
    elif kind == "res": fn = lambda: ops.linear_fwd(x, K, w, b, M, Nn, K, res=res)
    else: fn = lambda: ops.linear_fwd(x, K, w, b, M, Nn, K)
    run(f"fwd M{M} N{Nn} K{K} {kind}", fn)
dy = torch.randn(M, 48, device=dev); w = torch.randn(48, 192, device=dev)
run("bwd_data M65536 N48 K192", lambda: ops.linear_bwd_data(dy, 48, w, M, 48, 192))
xx = torch.randn(M, 48, device=dev)
run("bwd_weight M65536 N48 K48", lambda: ops.linear_bwd_weight(dy, 48, xx, 48, M, 48, 48))
