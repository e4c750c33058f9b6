"""Stand-alone device time of the non-GEMM kernels at the train config's shapes (batch 2, 128^3): LayerNorm fwd/bwd, deformable
sampling, offset head, the offset conv's three kernels, small-window attention.  Median of 10, L2 flushed between runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from micformer_b200 import ops, _native as N  # noqa: E402

N.set_gemm_mode(1)
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10, cold=True):
    """device time per call: 10 back-to-back calls captured in one CUDA graph (no host launch gaps); ``cold``: a 256 MB memset
    between the calls (its own time, measured the same way, is subtracted)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()

    def graph_time(body):
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            body()
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    def body_fn():
        for _ in range(n):
            if cold:
                flush.zero_()
            fn()

    def body_flush():
        for _ in range(n):
            flush.zero_()

    t = graph_time(body_fn)
    if cold:
        t -= graph_time(body_flush)
    return t / n


def r(*s):
    return torch.randn(*s, device=dev)


for (S, C, heads) in [(32, 48, 3), (16, 96, 6), (8, 192, 12), (4, 384, 24)]:
    B = 2
    dims = (B, S, S, S)
    T = B * S ** 3
    x = r(B, S, S, S, C); g = r(C); b = r(C); dy = r(B, S, S, S, C); dres = r(B, S, S, S, C)
    y, mean, rstd = ops.ln_fwd(x, None, g, b, dims)
    t_f = timeit(lambda: ops.ln_fwd(x, None, g, b, dims))
    t_b = timeit(lambda: ops.ln_bwd(dy, x, None, g, mean, rstd, dres, None, dims))
    t_bw = timeit(lambda: ops.ln_bwd(dy, x, None, g, mean, rstd, dres, None, dims), cold=False)
    mb = T * C * 4 / 1e6
    print(f"T={T:6d} C={C:3d}: ln_fwd {t_f:6.1f} us ({2 * mb / t_f * 1e-3 * 1e3:.0f} GB/s)  ln_bwd {t_b:6.1f} us cold / {t_bw:6.1f} warm "
          f"({4 * mb / t_b:.0f} GB/s algorithmic)")
    # small-window attention on (P, 3C) rows
    qkv = r(T, 3 * C)
    ws = (2, 2, 2)
    o, lse = ops.window_attn_fwd(qkv, C, heads, B, (S, S, S), ws)
    do = r(T, C)
    t_af = timeit(lambda: ops.window_attn_fwd(qkv, C, heads, B, (S, S, S), ws))
    t_ab = timeit(lambda: ops.window_attn_bwd(qkv, o, do, lse, C, heads, B, (S, S, S), ws))
    print(f"                 window attention 2^3 ({heads} heads): fwd {t_af:6.1f} us  bwd {t_ab:6.1f} us   (algorithmic {4 * mb / t_af:.0f} / {8 * mb / t_ab:.0f} GB/s)")
    # offset branch of a cross block
    HC = 16
    P = T
    xa = r(B, S, S, S, C); xn = y
    cw = r(27, 2 * C, HC) * 0.05; cwk = cw.permute(0, 2, 1).contiguous(); cb = r(HC)
    h16 = torch.empty(P, HC, device=dev)
    t_cf = timeit(lambda: ops.conv3_fwd(xn, xa, cw, cwk, cb, h16, B, (S, S, S), HC, False))
    lnw, lnb, w3 = r(HC), r(HC), r(3, HC) * 0.1
    pos = torch.empty(P, 3, device=dev)
    f_oh = lambda: N.call("mic_offset_head_fwd", N.ptr(h16), N.ptr(lnw), N.ptr(lnb), N.ptr(w3), N.ptr(pos), B, S, S, S, HC, 1e-5)
    t_oh = timeit(f_oh)
    samp = torch.empty(P, C, device=dev)
    f_ds = lambda: N.call("mic_deform_sample_fwd", N.ptr(xa), N.ptr(pos), N.ptr(samp), B, S, S, S, S, S, S, C)
    t_ds = timeit(f_ds)
    dsamp = r(P, C); dxa = torch.zeros(B, S, S, S, C, device=dev); dpos = torch.empty(P, 3, device=dev)
    f_dsb = lambda: N.call("mic_deform_sample_bwd", N.ptr(dsamp), N.ptr(xa), N.ptr(pos), N.ptr(dxa), N.ptr(dpos), B, S, S, S, S, S, S, C)
    t_dsb = timeit(f_dsb)
    dh16 = torch.empty(P, HC, device=dev); dlnw = torch.zeros(HC, device=dev); dlnb = torch.zeros(HC, device=dev); dw3 = torch.zeros(3, HC, device=dev)
    f_ohb = lambda: N.call("mic_offset_head_bwd", N.ptr(dpos), N.ptr(h16), N.ptr(lnw), N.ptr(lnb), N.ptr(w3), N.ptr(dh16), N.ptr(dlnw),
                           N.ptr(dlnb), N.ptr(dw3), B, S, S, S, HC, 1e-5)
    t_ohb = timeit(f_ohb)
    dcw = torch.zeros_like(cw); dcb = torch.zeros(HC, device=dev)
    t_cw = timeit(lambda: ops.conv3_bwd_weight(dh16, xn, xa, dcw, dcb, B, (S, S, S), HC, False))
    dcn = torch.zeros(HC, 2 * C, 3, 3, 3, device=dev)
    t_cwn = timeit(lambda: ops.conv3_bwd_weight(dh16, xn, xa, dcn, dcb, B, (S, S, S), HC, False, True))
    dxn = torch.empty(B, S, S, S, C, device=dev)
    t_cd = timeit(lambda: ops.conv3_bwd_data(dh16, cw, dxn, False, dxa, True, B, (S, S, S), HC, False))
    print(f"                 offset branch: conv fwd {t_cf:6.1f}  head fwd {t_oh:6.1f}  sample fwd {t_ds:6.1f} | sample bwd {t_dsb:6.1f}  head bwd {t_ohb:6.1f}  "
          f"conv dW {t_cw:6.1f} (native layout {t_cwn:6.1f})  conv bwd-data {t_cd:6.1f} us")
