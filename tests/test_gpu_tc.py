"""tcgen05 / TMA tensor-core path (gemm mode 1: fp32 I/O, operands read as TF32, fp32 accumulation in TMEM).

TF32 keeps 10 mantissa bits, so single GEMMs are held to 2e-3 of the output scale; the end-to-end bar is
BASELINE.json's: logits within 1e-3 relative of the fp32 CPU path, argmax equal wherever the top-2 margin
exceeds twice the measured error."""
import pytest
import torch
import torch.nn.functional as F

from oracle import micformer_oracle as O
from helpers import max_rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture()
def tc_mode():
    from micformer_b200 import _native as N
    prev = N.get_gemm_mode()
    N.set_gemm_mode(1)
    yield
    N.set_gemm_mode(prev)


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("M,N,K", [(256, 48, 32), (300, 144, 48), (1000, 192, 48), (128, 384, 1536), (517, 96, 24),
                                   (64, 1536, 96), (8192, 96, 96), (130, 80, 40),
                                   (40000, 144, 48), (33333, 80, 200)])     # >= 2 tiles per SM: persistent kernel
@pytest.mark.parametrize("w_is_kn", [False, True])
def test_tc_linear_all_layouts(tc_mode, M, N, K, w_is_kn):
    from micformer_b200 import ops
    x = _rand(M, K, seed=1); w = (_rand(K, N, seed=2) if w_is_kn else _rand(N, K, seed=2)) * 0.1
    b = _rand(N, seed=3); dy = _rand(M, N, seed=4)
    wm = (w if w_is_kn else w.t()).double()
    xd, wd, bd, dyd = x.to(DEV), w.to(DEV), b.to(DEV), dy.to(DEV)
    before = ops.N.launch_count()
    y = ops.linear_fwd(xd, K, wd, bd, M, N, K, w_is_kn=w_is_kn)
    dx = ops.linear_bwd_data(dyd, N, wd, M, N, K, w_is_kn=w_is_kn)
    dW, db = ops.linear_bwd_weight(dyd, N, xd, K, M, N, K, w_is_kn=w_is_kn)
    assert ops.N.launch_count() > before
    assert max_rel(y.cpu(), x.double() @ wm + b.double()) < 2e-3
    assert max_rel(dx.cpu(), dy.double() @ wm.t()) < 2e-3
    assert max_rel(dW.cpu(), (x.double().t() @ dy.double()) if w_is_kn else (dy.double().t() @ x.double())) < 2e-3
    assert max_rel(db.cpu(), dy.double().sum(0)) < 1e-5


@pytest.mark.parametrize("M", [2048, 65536])       # one-shot kernel / persistent kernel
def test_tc_epilogues_and_strided_views(tc_mode, M):
    from micformer_b200 import ops
    N, K = 192, 48
    x = _rand(M, K, seed=1); w = _rand(N, K, seed=2) * 0.1; b = _rand(N, seed=3); res = _rand(M, N, seed=4)
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    ref = x.double() @ w.double().t() + b.double()
    pre = torch.empty(M, N, device=DEV)
    yg = ops.linear_fwd(xd, K, wd, bd, M, N, K, act=True, pre=pre)
    assert max_rel(pre.cpu(), ref) < 2e-3 and max_rel(yg.cpu(), F.gelu(ref)) < 2e-3
    rs = torch.tensor([0.5, 2.0])
    yr = ops.linear_fwd(xd, K, wd, bd, M, N, K, res=res.to(DEV), rowscale=rs.to(DEV), rps=M // 2)
    assert max_rel(yr.cpu(), res.double() + ref * rs.double().repeat_interleave(M // 2)[:, None]) < 2e-3
    # write into a column slice of a wider buffer (q | kv layout) without touching the neighbours
    buf = torch.full((M, 3 * N), 7.0, device=DEV)
    ops.linear_fwd(xd, K, wd, bd, M, N, K, out=buf, out_col=N, ldy=3 * N)
    assert max_rel(buf[:, N:2 * N].cpu(), ref) < 2e-3
    assert bool((buf[:, :N] == 7.0).all()) and bool((buf[:, 2 * N:] == 7.0).all())
    # accumulate epilogue (TMA reduce-add)
    acc = res.to(DEV).clone()
    ops.linear_fwd(xd, K, wd, None, M, N, K, out=acc, ldy=N, accumulate=True)
    assert max_rel(acc.cpu(), res.double() + (ref - b.double())) < 2e-3
    # GELU' epilogue of backward-data and DropPath-scaled weight gradient
    dy = _rand(M, N, seed=5); dyd = dy.to(DEV)
    g = ops.linear_bwd_data(dyd, N, wd, M, N, K, gelu_pre=xd)
    xg = x.double().clone().requires_grad_(True)
    F.gelu(xg).backward(dy.double() @ w.double())
    assert max_rel(g.cpu(), xg.grad) < 2e-3
    dW, db = ops.linear_bwd_weight(dyd, N, xd, K, M, N, K, rowscale=rs.to(DEV), rps=M // 2)
    sdy = dy.double() * rs.double().repeat_interleave(M // 2)[:, None]
    assert max_rel(dW.cpu(), sdy.t() @ x.double()) < 2e-3 and max_rel(db.cpu(), sdy.sum(0)) < 1e-5


@pytest.mark.parametrize("S", [64, 128])
def test_tc_whole_model_logits_within_baseline_bar(tc_mode, S):
    from micformer_b200.models.MICFormer_self import Head
    from micformer_b200.loss.dice import MDiceLoss
    cfg = O.TRAIN
    sd = O.synth_state_dict(cfg, seed=7)
    x, lab = O.synth_inputs(1, S, cfg.num_classes, seed=9)
    head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
    head.load_state_dict(sd); head = head.to(DEV).eval()
    y = head(x.to(DEV))
    loss = MDiceLoss()(y, lab.to(DEV))
    loss.backward()
    logits, loss_ref, grads = O.train_step(x, lab, sd, cfg)
    yc = y.detach().cpu()
    err = float((yc - logits).abs().max())
    rel = err / float(logits.abs().max())
    print(f"TF32 logits rel err {rel:.2e}")
    assert rel < 1e-3
    top2 = logits.topk(2, dim=1).values
    mism = (yc.argmax(1) != logits.argmax(1)) & ((top2[:, 0] - top2[:, 1]) > 2 * err)
    assert int(mism.sum()) == 0
    assert abs(float(loss) - float(loss_ref)) < 1e-4
    # gradients: global relative error (per-tensor TF32 noise is larger on the tiny offset-net gradients)
    num = sum(float((p.grad.cpu().double() - grads[k].double()).pow(2).sum()) for k, p in head.named_parameters() if grads[k] is not None)
    den = sum(float(grads[k].double().pow(2).sum()) for k in grads if grads[k] is not None)
    assert (num / den) ** 0.5 < 2e-2


@pytest.mark.parametrize("B,pd,ws,C,heads", [(1, (7, 7, 7), (7, 7, 7), 96, 3), (2, (7, 14, 14), (7, 7, 7), 96, 3),
                                             (2, (8, 8, 8), (4, 8, 8), 64, 2), (1, (14, 7, 21), (7, 7, 7), 192, 6),
                                             (1, (8, 6, 12), (4, 6, 6), 64, 2),      # 144 tokens: single-pipeline kernel
                                             (2, (10, 6, 7), (5, 6, 7), 32, 1),      # 210 tokens: 192 + 18 key split
                                             (3, (21, 21, 21), (7, 7, 7), 96, 3)])   # 243 items: several per CTA
def test_tc_window_attention_large_windows(tc_mode, B, pd, ws, C, heads):
    """tcgen05 FlashAttention-style kernel (343-token windows, head_dim 32): TMA 5-D window gather, S/P in TMEM."""
    from micformer_b200 import ops
    P = B * pd[0] * pd[1] * pd[2]
    hd = C // heads
    qkv = _rand(P, 3 * C, seed=3)
    g = qkv.view(B, *pd, 3 * C).double()
    q, k, v = (O.window_partition(g[..., i * C:(i + 1) * C].contiguous(), ws) for i in range(3))
    Bw, Nt, _ = q.shape
    sp = lambda t: t.view(Bw, Nt, heads, hd).permute(0, 2, 1, 3)
    s = (sp(q) * hd ** -0.5) @ sp(k).transpose(-2, -1)
    o_ref = O.window_reverse((s.softmax(-1) @ sp(v)).transpose(1, 2).reshape(Bw, Nt, C), ws, B, *pd).reshape(-1, C)
    lse_ref = O.window_reverse(torch.logsumexp(s, -1).permute(0, 2, 1).contiguous(), ws, B, *pd).reshape(-1, heads)
    o, lse = ops.window_attn_fwd(qkv.to(DEV), C, heads, B, pd, ws)
    assert max_rel(o.cpu(), o_ref) < 3e-3          # TF32 (nearest) operands on N(0,1) scores up to |s| ~ 20
    assert float((lse.cpu().double() - lse_ref).abs().max()) < 3e-3
    # the CUDA-core backward consumes the tensor-core forward's output and log-sum-exp
    do = _rand(P, C, seed=4)
    qkv_r = qkv.double().requires_grad_(True)
    gr = qkv_r.view(B, *pd, 3 * C)
    q2, k2, v2 = (O.window_partition(gr[..., i * C:(i + 1) * C].contiguous(), ws) for i in range(3))
    o2 = O.window_reverse((((sp(q2) * hd ** -0.5) @ sp(k2).transpose(-2, -1)).softmax(-1) @ sp(v2)).transpose(1, 2)
                          .reshape(Bw, Nt, C), ws, B, *pd).reshape(-1, C)
    (o2 * do.double()).sum().backward()
    dqkv = ops.window_attn_bwd(qkv.to(DEV), o, do.to(DEV), lse, C, heads, B, pd, ws)
    assert max_rel(dqkv.cpu(), qkv_r.grad) < 5e-3


def test_tc_w7_model_logits(tc_mode):
    from micformer_b200.models.MICFormer_self import Head
    cfg = O.W7
    sd = O.synth_state_dict(cfg, seed=7)
    x, lab = O.synth_inputs(1, 64, cfg.num_classes, seed=9)
    head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
    head.load_state_dict(sd); head = head.to(DEV).eval()
    with torch.no_grad():
        y = head(x.to(DEV)).cpu()
        ref = O.head_forward(x, sd, cfg)
    rel = float((y - ref).abs().max() / ref.abs().max())
    print(f"TF32 w7 logits rel err {rel:.2e}")
    assert rel < 1e-3


@pytest.mark.parametrize("C0,C1,Co,dims,ncdhw", [(48, 48, 16, (8, 32, 32), False),      # conv_offset, stage-0 planes
                                                 (96, 96, 16, (16, 16, 16), False),     # conv_offset, stage 1
                                                 (24, 0, 8, (12, 32, 16), True),        # out_conv (NCDHW logits / dlogits)
                                                 (32, 0, 16, (5, 16, 8), False),
                                                 (192, 192, 16, (8, 8, 8), False),      # stage 2: half-empty footprint, K split
                                                 (384, 384, 16, (4, 4, 4), False),      # stage 3
                                                 (24, 24, 16, (7, 7, 7), False),        # odd sizes (window-7 padded grids)
                                                 (16, 0, 8, (3, 20, 9), True),
                                                 (24, 0, 8, (9, 10, 16), False),        # Co = 8 channels-last
                                                 (40, 24, 8, (6, 7, 12), True)])        # two sources, two 32-channel chunks, ragged bricks
def test_tc_conv3_fwd_bwd(tc_mode, C0, C1, Co, dims, ncdhw):
    """tcgen05 implicit-GEMM 3x3x3 conv: forward, backward-data (mirrored taps through the same kernel, accumulate
    epilogue) and backward-weight against F.conv3d autograd in fp64."""
    from micformer_b200 import _native as N, ops
    from helpers import rel_err
    B = 2
    D, H, W = dims
    Cin = C0 + C1
    x0 = _rand(B, D, H, W, C0, seed=1).double().requires_grad_(True)
    x1 = _rand(B, D, H, W, C1, seed=2).double().requires_grad_(True) if C1 else None
    w = (_rand(Co, Cin, 3, 3, 3, seed=3) * 0.1).double().requires_grad_(True)
    b = _rand(Co, seed=4).double().requires_grad_(True)
    xin = torch.cat([x0, x1], -1) if C1 else x0
    y_ref = F.conv3d(xin.permute(0, 4, 1, 2, 3), w, b, padding=1)
    gy = _rand(*y_ref.shape, seed=5).double()
    (y_ref * gy).sum().backward()
    wt = w.detach().float().permute(2, 3, 4, 1, 0).reshape(27, Cin, Co).contiguous().to(DEV)
    wk = w.detach().float().permute(2, 3, 4, 0, 1).reshape(27, Co, Cin).contiguous().to(DEV)
    x0d = x0.detach().float().to(DEV)
    x1d = x1.detach().float().to(DEV) if C1 else None
    y = torch.empty((B, Co, D, H, W) if ncdhw else (B, D, H, W, Co), device=DEV)
    N.call("mic_conv3_tc_fwd", N.ptr(x0d), C0, N.ptr(x1d), C1, N.ptr(wk), N.ptr(b.detach().float().to(DEV)), N.ptr(y), B, D,
           H, W, Co, int(ncdhw))
    yk = y.cpu() if ncdhw else y.cpu().permute(0, 4, 1, 2, 3)
    assert max_rel(yk, y_ref.detach()) < 2e-3
    dy = (gy if ncdhw else gy.permute(0, 2, 3, 4, 1)).float().contiguous().to(DEV)
    dx0 = torch.full((B, D, H, W, C0), 1.0, device=DEV)         # acc0=1: accumulates onto the ones
    dx1 = torch.full((B, D, H, W, max(C1, 4)), 7.0, device=DEV)  # acc1=0: overwritten
    N.call("mic_conv3_tc_bwd_data", N.ptr(dy), N.ptr(wt), N.ptr(dx0), C0, 1, N.ptr(dx1) if C1 else None, C1, 0, B, D, H, W,
           Co, int(ncdhw))
    assert rel_err(dx0.cpu() - 1.0, x0.grad) < 2e-3
    if C1:
        assert rel_err(dx1.cpu(), x1.grad) < 2e-3
    dwt = torch.zeros_like(wt)
    dbias = torch.zeros(Co, device=DEV)
    ops.conv3_bwd_weight(dy, x0d, x1d, dwt, dbias, B, (D, H, W), Co, ncdhw)
    dw_ref = w.grad.permute(2, 3, 4, 1, 0).reshape(27, Cin, Co)
    assert rel_err(dwt.cpu(), dw_ref) < 2e-3 and rel_err(dbias.cpu(), b.grad) < 1e-4


def test_unpatch_view_tail_gemms(tc_mode):
    """mic_linear_unpatch_view: the decoder tail's GEMMs address the (B, 4D, 4H, 4W, ch) grid directly (5-D TMA maps) --
    same numbers as GEMM + mic_block_permute (reference M:1033-1037), forward and both backward GEMMs; a grid that is not
    32 cells wide is refused loudly instead of being read as a plain matrix."""
    from micformer_b200 import ops
    B, D, H, W, Ch, K = 1, 2, 4, 32, 8, 48
    T = B * D * H * W
    Nout = 64 * Ch
    x = _rand(T, K, seed=1).to(DEV); w = (_rand(K, Nout, seed=2) * 0.2).to(DEV); b = _rand(Nout, seed=3).to(DEV)
    view = (Ch, D, H, W)
    # forward: rows -> permute  vs  direct
    rows = ops.linear_fwd(x, K, w, b, T, Nout, K, w_is_kn=True)
    fine_ref = torch.empty(B, 4 * D, 4 * H, 4 * W, Ch, device=DEV)
    ops.block_permute(rows, fine_ref, B, D, H, W, 4, Ch, 64 * D * H * W * Ch, False)
    fine = torch.full_like(fine_ref, 7.0)
    ops.linear_fwd(x, K, w, b, T, Nout, K, w_is_kn=True, out=fine, ldy=Nout, view=view)
    assert torch.equal(fine, fine_ref)
    # backward: dY given on the fine grid
    dfine = _rand(B, 4 * D, 4 * H, 4 * W, Ch, seed=4).to(DEV)
    drows = torch.empty(T, Nout, device=DEV)
    ops.block_permute(dfine, drows, B, D, H, W, 4, Ch, 64 * D * H * W * Ch, True)
    dx_ref = ops.linear_bwd_data(drows, Nout, w, T, Nout, K, w_is_kn=True)
    dx = ops.linear_bwd_data(dfine, Nout, w, T, Nout, K, w_is_kn=True, view=view)
    assert max_rel(dx.cpu(), dx_ref.cpu().double()) < 1e-6
    dw_ref, db_ref = ops.linear_bwd_weight(drows, Nout, x, K, T, Nout, K, w_is_kn=True)
    dw, db = ops.linear_bwd_weight(dfine, Nout, x, K, T, Nout, K, w_is_kn=True, view=view)
    assert max_rel(dw.cpu(), dw_ref.cpu().double()) < 1e-5
    # db: exact after summing the 64 block positions of a channel (a bias tiled over the block positions)
    assert max_rel(db.view(64, Ch).sum(0).cpu(), db_ref.view(64, Ch).sum(0).cpu().double()) < 1e-5
    # 16 cells wide: refused, and the pending view does not leak into the next call
    with pytest.raises(RuntimeError):
        ops.linear_fwd(x[:T // 2], K, w, b, T // 2, Nout, K, w_is_kn=True, out=fine, ldy=Nout, view=(Ch, D, H, 16))
    rows2 = ops.linear_fwd(x, K, w, b, T, Nout, K, w_is_kn=True)
    assert torch.equal(rows2, rows)
