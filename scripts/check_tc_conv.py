import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from micformer_b200 import _native as N
dev = "cuda"
def run(B, D, H, W, C0, C1, Co, ncdhw):
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B, D, H, W, C0, generator=g); x1 = torch.randn(B, D, H, W, C1, generator=g) if C1 else None
    w = torch.randn(Co, C0 + C1, 3, 3, 3, generator=g) * 0.1; b = torch.randn(Co, generator=g)
    xin = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.conv3d(xin.permute(0, 4, 1, 2, 3).double(), w.double(), b.double(), padding=1)
    wk = w.permute(2, 3, 4, 0, 1).reshape(27, Co, C0 + C1).contiguous().to(dev)
    wt = w.permute(2, 3, 4, 1, 0).reshape(27, C0 + C1, Co).contiguous().to(dev)
    x0d = x0.to(dev); x1d = x1.to(dev) if C1 else None; bd = b.to(dev)
    shape = (B, Co, D, H, W) if ncdhw else (B, D, H, W, Co)
    y = torch.full(shape, float("nan"), device=dev)
    ok = N.try_call("mic_conv3_tc_fwd", N.ptr(x0d), C0, N.ptr(x1d), C1, N.ptr(wk), N.ptr(bd), N.ptr(y), B, D, H, W, Co, int(ncdhw))
    torch.cuda.synchronize()
    yk = y.cpu().double() if ncdhw else y.cpu().double().permute(0, 4, 1, 2, 3)
    err = float((yk - ref).abs().max() / ref.abs().max()) if ok else None
    ys = torch.empty(shape, device=dev)
    def tc(): N.try_call("mic_conv3_tc_fwd", N.ptr(x0d), C0, N.ptr(x1d), C1, N.ptr(wk), N.ptr(bd), N.ptr(y), B, D, H, W, Co, int(ncdhw))
    def simt(): N.call("mic_conv3_fwd", N.ptr(x0d), C0, N.ptr(x1d), C1, N.ptr(wt), N.ptr(bd), N.ptr(ys), B, D, H, W, D, H, W, Co, int(ncdhw))
    ts = []
    for fn in (tc, simt):
        for _ in range(2): fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 10 * 1e3)
    print(f"B{B} {D}x{H}x{W} C{C0}+{C1}->{Co} ncdhw={int(ncdhw)}: taken={ok} rel err {err}  tc {ts[0]:.1f} us  simt {ts[1]:.1f} us", flush=True)
run(1, 4, 16, 8, 16, 0, 16, False)
run(2, 6, 16, 16, 24, 24, 16, False)
run(2, 32, 32, 32, 48, 48, 16, False)
run(2, 16, 16, 16, 96, 96, 16, False)
run(1, 32, 32, 32, 24, 0, 8, True)
run(2, 128, 128, 128, 24, 0, 8, True)
