"""Timing of the fused block kernels vs the unfused tcgen05 path at the train config's stage-0 shape (T=65536, C=48)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, fused, _native as N
N.set_gemm_mode(1)
dev = torch.device("cuda")
T, C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 48
g = torch.Generator().manual_seed(0)
x = torch.randn(T, C, generator=g).to(dev)
n2w, n2b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
f1w = ((torch.rand(4 * C, C, generator=g) * 2 - 1) * 0.3).to(dev); f1b = torch.zeros(4 * C, device=dev)
f2w = ((torch.rand(C, 4 * C, generator=g) * 2 - 1) * 0.3).to(dev); f2b = torch.zeros(C, device=dev)
dy = torch.randn(T, C, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


img = fused.mlp_images(f1w, f2w)
t_img = timeit(img.refresh)
t_fused = timeit(lambda: fused.mlp_block_fwd(x, img, n2w, n2b, f1b, f2b, None, 1, 1e-5))
dims = (1, 1, 1, T)
t_unf = timeit(lambda: ops._mlp_fwd(x.view(1, 1, 1, T, C), n2w, n2b, f1w, f1b, f2w, f2b, None, dims))
print(f"T={T} C={C}: weight images {t_img:.1f} us | MLP fwd fused {t_fused:.1f} us vs unfused (LN+fc1+fc2) {t_unf:.1f} us")
if hasattr(fused, "mlp_block_bwd"):
    y, saved = ops._mlp_fwd(x.view(1, 1, 1, T, C), n2w, n2b, f1w, f1b, f2w, f2b, None, dims)
    def unf_bwd():
        with ops.zero_arena(12 * C * C + 128 * C + 4096, x), ops.side_branch() as sb:
            ops._mlp_bwd(sb, dy, x.view(1, 1, 1, T, C), saved, n2w, f1w, f2w, None, dims)
    t_ub = timeit(unf_bwd)
    gw = [torch.zeros_like(t) for t in (n2w, n2b, f1w, f1b, f2w, f2b)]
    t_fb = timeit(lambda: fused.mlp_block_bwd(dy, x, img, n2w, n2b, f1b, None, 1, 1e-5, *gw))
    print(f"           MLP bwd fused {t_fb:.1f} us vs unfused {t_ub:.1f} us")

# ---- attention half-block (B=2, 32^3 grid, C=48, 3 heads): fused vs the unfused kernel sequence of SelfBlockFn
if C == 48 and T == 65536:
    B_, D_ = 2, 32
    xg = x.view(B_, D_, D_, D_, C)
    u = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * 0.3).to(dev)
    qw, kvw, pw = u(C, C), u(2 * C, C), u(C, C)
    qb, kvb, pb = torch.zeros(C, device=dev), torch.zeros(2 * C, device=dev), torch.zeros(C, device=dev)
    aimg = fused.attn_images(qw, kvw, pw)
    aimg.refresh()
    src = torch.randn(B_, D_, D_, D_, C, generator=g).to(dev)
    for cross in (False, True):
        s_ = src if cross else None
        t_f = timeit(lambda: fused.attn_block_fwd(xg, s_, aimg, n2w, n2b, qb, kvb, pb, None, 3, 1e-5))
        names = (n2w, n2b, qw, qb, kvw, kvb, pw, pb)
        gb = [torch.zeros_like(t) for t in names]
        t_b = timeit(lambda: fused.attn_block_bwd(dy.view_as(xg), xg, s_, aimg, n2w, n2b, qb, kvb, None, 3, 1e-5, *gb))
        print(f"           attention half ({'cross' if cross else 'self'}): fwd fused {t_f:.1f} us, bwd fused {t_b:.1f} us")
