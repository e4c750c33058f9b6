"""Per-phase %globaltimer trace of the fused attention backward kernel (first 4 CTAs, first 7 tiles each)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from micformer_b200 import fused, _native as N
N.set_gemm_mode(1)
lib = N.load()
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
C, B_, D_ = 48, 2, 32
x = torch.randn(B_, D_, D_, D_, C, generator=g).to(dev)
dy = torch.randn(B_, D_, D_, D_, C, generator=g).to(dev)
src = torch.randn(B_, D_, D_, D_, C, generator=g).to(dev)
u = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * 0.3).to(dev)
qw, kvw, pw = u(C, C), u(2 * C, C), u(C, C)
n1w, n1b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
qb, kvb = torch.zeros(C, device=dev), torch.zeros(2 * C, device=dev)
img = fused.attn_images(qw, kvw, pw); img.refresh()
gb = [torch.zeros_like(t) for t in (n1w, n1b, qw, qb, kvw, kvb, pw, qb)]
buf = torch.zeros(4 * 8 * 32, dtype=torch.int64, device=dev)
for cross in (False, True):
    s_ = src if cross else None
    for _ in range(2):
        fused.attn_block_bwd(dy, x, s_, img, n1w, n1b, qb, kvb, None, 3, 1e-5, *gb)
    torch.cuda.synchronize()
    lib.mic_debug_t5_trace(ctypes.c_void_p(buf.data_ptr()))
    buf.zero_()
    fused.attn_block_bwd(dy, x, s_, img, n1w, n1b, qb, kvb, None, 3, 1e-5, *gb)
    torch.cuda.synchronize()
    lib.mic_debug_t5_trace(ctypes.c_void_p(0))
    t = buf.cpu().view(4, 8, 32)
    names = ["tile start", "staged", "q/k/v/do ready", "o written", "dk/dv done", "dWp done", "dq/dkv written", "dxn ready", "tile end"]
    print("cross" if cross else "self")
    for cta in range(2):
        t0 = int(t[cta, 0, 0])
        for n in range(4):
            row = [int(t[cta, n, k]) - t0 for k in range(9)]
            if row[8] <= 0: continue
            print(f"  cta {cta} tile {n}: " + "  ".join(f"{names[k]} {row[k] / 1000:.2f}" for k in range(9)))
        print(f"  cta {cta}: loop end {(int(t[cta, 7, 0]) - t0) / 1000:.2f} us, flushed {(int(t[cta, 7, 1]) - t0) / 1000:.2f} us")
