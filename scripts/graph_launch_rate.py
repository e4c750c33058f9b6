"""How fast does a replayed CUDA graph start small kernels?  N kernels spread over k parallel branches (streams), tiny torch
elementwise kernels vs this package's tcgen05 GEMM (programmatic dependent launch, 4 tensor maps in the parameters) vs LayerNorm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from micformer_b200 import ops, _native as N
N.set_gemm_mode(1)
dev = torch.device("cuda")


def bench(make_fn, nk=512, branches=(1, 2, 4, 8)):
    out = []
    for k in branches:
        fns = [make_fn(i) for i in range(k)]
        streams = [torch.cuda.Stream() for _ in range(k)]
        def body():
            cur = torch.cuda.current_stream()
            for st in streams:
                st.wait_stream(cur)
            for j in range(nk // k):
                for st, fn in zip(streams, fns):
                    with torch.cuda.stream(st):
                        fn()
            for st in streams:
                cur.wait_stream(st)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        out.append((k, ts[2] / nk))
    return out


def mk_add(i):
    x = torch.zeros(1024, device=dev)
    return lambda: x.add_(1.0)


def mk_gemm(i):
    M, K, Nn = 1024, 192, 192
    x = torch.randn(M, K, device=dev); w = torch.randn(Nn, K, device=dev); b = torch.randn(Nn, device=dev)
    y = torch.empty(M, Nn, device=dev)
    return lambda: ops.linear_fwd(x, K, w, b, M, Nn, K, out=y, ldy=Nn)


def mk_ln(i):
    x = torch.randn(2, 8, 8, 8, 192, device=dev); g = torch.randn(192, device=dev); b = torch.randn(192, device=dev)
    return lambda: ops.ln_fwd(x, None, g, b, (2, 8, 8, 8))


for name, mk in (("torch add_ (1024 floats)", mk_add), ("LayerNorm fwd T=1024 C=192", mk_ln), ("tcgen05 GEMM 1024x192x192", mk_gemm)):
    res = bench(mk)
    print(f"{name:30s} us per kernel (whole graph time / kernels): " + "  ".join(f"{k} branches: {t:.2f}" for k, t in res))
for pdl in ("0",):
    os.environ["MICFORMER_PDL"] = pdl
