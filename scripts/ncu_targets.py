"""One launch of each kernel the round-2 DESIGN.md discusses, as a short `ncu --set full` target:
fused stage-0 block kernels (T = 65536, C = 48), the one-shot GEMM in its four roles at a stage-2 shape, the offset conv's
three kernels at stage 0, LayerNorm backward at stage 0, the large-window attention forward (BASELINE config 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, fused, _native as N
N.set_gemm_mode(1)
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).to(dev)
u = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * 0.3).to(dev)

# ---- fused stage-0 kernels
B, D, C, heads = 2, 32, 48, 3
T = B * D ** 3
x = r(B, D, D, D, C); dy = r(B, D, D, D, C)
n1w, n1b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
qw, kvw, pw = u(C, C), u(2 * C, C), u(C, C)
qb, kvb, pb = torch.zeros(C, device=dev), torch.zeros(2 * C, device=dev), torch.zeros(C, device=dev)
f1w, f2w = u(4 * C, C), u(C, 4 * C); f1b, f2b = torch.zeros(4 * C, device=dev), torch.zeros(C, device=dev)
aimg = fused.attn_images(qw, kvw, pw); mimg = fused.mlp_images(f1w, f2w)
fused.refresh_all([aimg, mimg])
x1 = fused.attn_block_fwd(x, None, aimg, n1w, n1b, qb, kvb, pb, None, heads, 1e-5)
y = fused.mlp_block_fwd(x1, mimg, n1w, n1b, f1b, f2b, None, D ** 3, 1e-5)
ga = [torch.zeros_like(t) for t in (n1w, n1b, qw, qb, kvw, kvb, pw, pb)]
fused.attn_block_bwd(dy, x, None, aimg, n1w, n1b, qb, kvb, None, heads, 1e-5, *ga)
gm = [torch.zeros_like(t) for t in (n1w, n1b, f1w, f1b, f2w, f2b)]
fused.mlp_block_bwd(dy.view(T, C), x1.view(T, C), mimg, n1w, n1b, f1b, None, 1, 1e-5, *gm)

# ---- one-shot GEMM, stage-2 shape (T = 1024, C = 192): forward split-bf16, forward split-K, data gradient, weight gradient (+ bias)
M, K, Nn = 1024, 192, 768
a = r(M, K); w1 = u(Nn, K); b1 = r(Nn); w2 = u(K, Nn); b2 = r(K); res = r(M, K)
h = ops.linear_fwd(a, K, w1, b1, M, Nn, K)                              # fc1-like: K = 192
ops.linear_fwd(h, Nn, w2, b2, M, K, Nn, res=res)                        # fc2-like: K = 768 -> split-K + residual init
dh = r(M, Nn)
ops.linear_bwd_data(dh, Nn, w1, M, Nn, K)                               # reduction over 768
ops.linear_bwd_weight(dh, Nn, a, K, M, Nn, K)                           # dW (768 x 192) + db folded in

# ---- offset conv at stage 0 + LayerNorm backward
HC = 16
xa = r(B, D, D, D, C)
xn, mean, rstd = ops.ln_fwd(x, None, n1w, n1b, (B, D, D, D))
cw = r(27, 2 * C, HC) * 0.05; cwk = cw.permute(0, 2, 1).contiguous(); cb = r(HC)
h16 = torch.empty(T, HC, device=dev)
ops.conv3_fwd(xn, xa, cw, cwk, cb, h16, B, (D, D, D), HC, False)
dh16 = r(T, HC); dxn = torch.empty_like(x); dxa = torch.zeros_like(x)
ops.conv3_bwd_data(dh16, cw, dxn, False, dxa, True, B, (D, D, D), HC, False)
dcw = torch.zeros_like(cw); dcb = torch.zeros(HC, device=dev)
ops.conv3_bwd_weight(dh16, xn, xa, dcw, dcb, B, (D, D, D), HC, False)
ops.ln_bwd(dy, x, None, n1w, mean, rstd, dxn, None, (B, D, D, D))

# ---- BASELINE config 4: 4096 windows x 343 tokens x 96 channels x 3 heads
qkv = r(4096 * 343, 3 * 96)
o_a, lse_a = ops.window_attn_fwd(qkv, 96, 3, 4096, (7, 7, 7), (7, 7, 7))
# ---- the same shape's tcgen05 backward
ops.window_attn_bwd(qkv, o_a, r(4096 * 343, 96), lse_a, 96, 3, 4096, (7, 7, 7), (7, 7, 7))
# ---- out_conv weight gradient (row-reuse mma.sync kernel): 2 x 128^3, 24 -> 8 classes, NCDHW dlogits
dwo = torch.zeros(27, 24, 8, device="cuda"); dbo = torch.zeros(8, device="cuda")
ops.conv3_bwd_weight(r(2, 8, 128, 128, 128), r(2, 128, 128, 128, 24), None, dwo, dbo, 2, (128, 128, 128), 8, True)
torch.cuda.synchronize()
print("done")
