"""GPU parity tests: the sm_100a kernels (through the C ABI) against the CPU oracle and the golden vectors.

Tolerances: everything here is fp32 arithmetic; the bar from BASELINE.json is 1e-3 relative on logits.  The
exact-fp32 kernels are held to much tighter bounds (1e-4 .. 1e-5, reduction-order noise only)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import micformer_oracle as O
from helpers import load_golden, block_weights, block_inputs, rel_err, max_rel

pytestmark = pytest.mark.gpu

VEC, META = load_golden()
BLOCKS = [k for k in META["cases"] if k.startswith(("self_", "cross_"))]
DEV = "cuda"


def _ops():
    from micformer_b200 import ops
    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("C0,C1,pad", [(48, 0, False), (32, 0, True), (24, 24, False), (768, 0, False), (16, 0, True)])
def test_layernorm_fwd_bwd(C0, C1, pad):
    ops = _ops()
    B, D, H, W = 2, 3, 4, 5
    pd = (4, 6, 7) if pad else (D, H, W)
    x0 = _rand(B, D, H, W, C0, seed=1).requires_grad_(True)
    x1 = _rand(B, D, H, W, C1, seed=2).requires_grad_(True) if C1 else None
    g = (1 + 0.1 * _rand(C0 + C1, seed=3)).requires_grad_(True)
    b = (0.1 * _rand(C0 + C1, seed=4)).requires_grad_(True)
    gy = _rand(B, *pd, C0 + C1, seed=5)
    xin = torch.cat([x0, x1], -1) if C1 else x0
    y = O.layer_norm(xin, g, b)
    y = F.pad(y, (0, 0, 0, pd[2] - W, 0, pd[1] - H, 0, pd[0] - D))
    (y * gy).sum().backward()
    dres = _rand(B, D, H, W, C0, seed=6)
    yk, mean, rstd = ops.ln_fwd(x0.detach().to(DEV), x1.detach().to(DEV) if C1 else None, g.detach().to(DEV),
                                b.detach().to(DEV), (B, D, H, W), pd)
    assert max_rel(yk.cpu(), y.detach()) < 1e-5
    dx0, dx1, dg, db = ops.ln_bwd(gy.to(DEV), x0.detach().to(DEV), x1.detach().to(DEV) if C1 else None, g.detach().to(DEV),
                                  mean, rstd, dres.to(DEV), None, (B, D, H, W), pd)
    assert rel_err(dx0.cpu(), x0.grad + dres) < 1e-5
    if C1:
        assert rel_err(dx1.cpu(), x1.grad) < 1e-5
    assert rel_err(dg.cpu(), g.grad) < 1e-5 and rel_err(db.cpu(), b.grad) < 1e-5


@pytest.mark.parametrize("M,N,K", [(300, 144, 48), (129, 20, 24), (1000, 192, 48), (64, 768, 1536), (517, 3, 16)])
@pytest.mark.parametrize("w_is_kn", [False, True])
def test_linear_fwd_bwd(M, N, K, w_is_kn):
    ops = _ops()
    x = _rand(M, K, seed=1)
    w = _rand(K, N, seed=2) * 0.1 if w_is_kn else _rand(N, K, seed=2) * 0.1
    b = _rand(N, seed=3)
    dy = _rand(M, N, seed=4)
    wm = w if w_is_kn else w.t()
    y_ref = x @ wm + b
    xd, wd, bd, dyd = x.to(DEV), w.to(DEV), b.to(DEV), dy.to(DEV)
    y = ops.linear_fwd(xd, K, wd, bd, M, N, K, w_is_kn=w_is_kn)
    assert max_rel(y.cpu(), y_ref) < 1e-5
    # gelu epilogue + saved pre-activation
    pre = torch.empty(M, N, device=DEV)
    yg = ops.linear_fwd(xd, K, wd, bd, M, N, K, w_is_kn=w_is_kn, act=True, pre=pre)
    assert max_rel(pre.cpu(), y_ref) < 1e-5 and max_rel(yg.cpu(), F.gelu(y_ref)) < 1e-5
    # residual + per-sample scale (2 samples)
    if M % 2 == 0:
        res = _rand(M, N, seed=5)
        rs = torch.tensor([0.0, 1.25])
        yr = ops.linear_fwd(xd, K, wd, bd, M, N, K, w_is_kn=w_is_kn, res=res.to(DEV), rowscale=rs.to(DEV), rps=M // 2)
        ref = res + y_ref * rs.repeat_interleave(M // 2)[:, None]
        assert max_rel(yr.cpu(), ref) < 1e-5
    dx = ops.linear_bwd_data(dyd, N, wd, M, N, K, w_is_kn=w_is_kn)
    assert max_rel(dx.cpu(), dy @ wm.t()) < 1e-5
    dxg = ops.linear_bwd_data(dyd, N, wd, M, N, K, w_is_kn=w_is_kn, gelu_pre=xd)
    xg = x.clone().requires_grad_(True)
    F.gelu(xg).backward(dy @ wm.t())
    assert max_rel(dxg.cpu(), xg.grad) < 1e-5
    dW, db = ops.linear_bwd_weight(dyd, N, xd, K, M, N, K, w_is_kn=w_is_kn)
    dW_ref = x.t() @ dy if w_is_kn else dy.t() @ x
    assert rel_err(dW.cpu(), dW_ref) < 1e-5 and rel_err(db.cpu(), dy.sum(0)) < 1e-5


def _attn_ref(qkv, C, heads, B, pd, ws):
    """oracle window attention core on a (P,3C) buffer laid out on the padded grid"""
    hd = C // heads
    g = qkv.view(B, *pd, 3 * C)
    q = O.window_partition(g[..., :C].contiguous(), ws)
    k = O.window_partition(g[..., C:2 * C].contiguous(), ws)
    v = O.window_partition(g[..., 2 * C:].contiguous(), ws)
    Bw, Nt, _ = q.shape
    sp = lambda t: t.view(Bw, Nt, heads, hd).permute(0, 2, 1, 3)
    a = ((sp(q) * hd ** -0.5) @ sp(k).transpose(-2, -1)).softmax(-1)
    o = (a @ sp(v)).transpose(1, 2).reshape(Bw, Nt, C)
    return O.window_reverse(o, ws, B, *pd).reshape(-1, C)


@pytest.mark.parametrize("C,heads,pd,ws", [(48, 3, (4, 4, 6), (2, 2, 2)), (96, 3, (7, 7, 14), (7, 7, 7)),
                                           (64, 2, (4, 4, 4), (4, 4, 4)), (24, 2, (2, 4, 2), (2, 2, 2)),
                                           (48, 6, (2, 2, 4), (2, 2, 2)), (72, 3, (2, 2, 2), (2, 2, 2))])
def test_window_attention_core(C, heads, pd, ws):
    ops = _ops()
    B = 2
    P = B * pd[0] * pd[1] * pd[2]
    qkv = _rand(P, 3 * C, seed=7).requires_grad_(True)
    do = _rand(P, C, seed=8)
    o_ref = _attn_ref(qkv, C, heads, B, pd, ws)
    (o_ref * do).sum().backward()
    qd = qkv.detach().to(DEV)
    o, lse = ops.window_attn_fwd(qd, C, heads, B, pd, ws)
    assert max_rel(o.cpu(), o_ref.detach()) < 2e-5
    dqkv = ops.window_attn_bwd(qd, o, do.to(DEV), lse, C, heads, B, pd, ws)
    assert rel_err(dqkv.cpu(), qkv.grad) < 2e-5


@pytest.mark.parametrize("C0,C1,Co,dims,pd,ncdhw", [(24, 24, 16, (4, 5, 6), (4, 6, 6), False),
                                                    (48, 48, 16, (8, 8, 8), (8, 8, 8), False),
                                                    (24, 0, 8, (8, 8, 12), (8, 8, 12), True),
                                                    (12, 0, 14, (5, 6, 7), (5, 6, 7), True),
                                                    (40, 40, 16, (3, 3, 3), (7, 7, 7), False)])
def test_conv3(C0, C1, Co, dims, pd, ncdhw):
    from micformer_b200 import _native as N
    B = 2
    D, H, W = dims
    x0 = _rand(B, D, H, W, C0, seed=1).requires_grad_(True)
    x1 = _rand(B, D, H, W, C1, seed=2).requires_grad_(True) if C1 else None
    w = (_rand(Co, C0 + C1, 3, 3, 3, seed=3) * 0.1).requires_grad_(True)
    b = _rand(Co, seed=4).requires_grad_(True)
    xin = torch.cat([x0, x1], -1) if C1 else x0
    xin = F.pad(xin, (0, 0, 0, pd[2] - W, 0, pd[1] - H, 0, pd[0] - D))
    y_ref = F.conv3d(xin.permute(0, 4, 1, 2, 3), w, b, padding=1)       # (B,Co,Dp,Hp,Wp)
    gy = _rand(*y_ref.shape, seed=5)
    (y_ref * gy).sum().backward()
    wt = w.detach().permute(2, 3, 4, 1, 0).reshape(27, C0 + C1, Co).contiguous().to(DEV)
    x0d = x0.detach().to(DEV)
    x1d = x1.detach().to(DEV) if C1 else None
    shape = (B, Co, *pd) if ncdhw else (B, *pd, Co)
    y = torch.empty(shape, device=DEV)
    N.call("mic_conv3_fwd", N.ptr(x0d), C0, N.ptr(x1d), C1, N.ptr(wt), N.ptr(b.detach().to(DEV)), N.ptr(y), B, D, H, W,
           *pd, Co, int(ncdhw))
    yk = y.cpu() if ncdhw else y.cpu().permute(0, 4, 1, 2, 3)
    assert max_rel(yk, y_ref.detach()) < 2e-5
    dy = (gy if ncdhw else gy.permute(0, 2, 3, 4, 1)).contiguous().to(DEV)
    dx0 = torch.full((B, D, H, W, C0), 1.0, device=DEV)         # acc0=1: accumulates onto the ones
    dx1 = torch.empty(B, D, H, W, max(C1, 1), device=DEV)
    N.call("mic_conv3_bwd_data", N.ptr(dy), N.ptr(wt), N.ptr(dx0), C0, 1, N.ptr(dx1) if C1 else None, C1, 0, B, D, H, W,
           *pd, Co, int(ncdhw))
    assert rel_err(dx0.cpu() - 1.0, x0.grad) < 2e-5
    if C1:
        assert rel_err(dx1.cpu(), x1.grad) < 2e-5
    dwt = torch.zeros_like(wt)
    dbias = torch.zeros(Co, device=DEV)
    N.call("mic_conv3_bwd_weight", N.ptr(dy), N.ptr(x0d), C0, N.ptr(x1d), C1, N.ptr(dwt), N.ptr(dbias), B, D, H, W, *pd, Co,
           int(ncdhw))
    dw_ref = w.grad.permute(2, 3, 4, 1, 0).reshape(27, C0 + C1, Co)
    assert rel_err(dwt.cpu(), dw_ref) < 2e-5 and rel_err(dbias.cpu(), b.grad) < 2e-5


def test_offset_head():
    from micformer_b200 import _native as N
    B, pd = 2, (3, 4, 5)
    P = B * 3 * 4 * 5
    h = _rand(P, 16, seed=1, scale=2.0).requires_grad_(True)
    g = (1 + 0.1 * _rand(16, seed=2)).requires_grad_(True)
    b = (0.1 * _rand(16, seed=3)).requires_grad_(True)
    w3 = (_rand(3, 16, seed=4) * 0.3).requires_grad_(True)
    pos_ref = F.linear(F.gelu(O.layer_norm(h, g, b)), w3).view(B, *pd, 3) + O.ref_points(*pd).unsqueeze(0)
    gp = _rand(B, *pd, 3, seed=5)
    (pos_ref * gp).sum().backward()
    hd, gd, bd, wd = (t.detach().to(DEV) for t in (h, g, b, w3))
    pos = torch.empty(P, 3, device=DEV)
    N.call("mic_offset_head_fwd", N.ptr(hd), N.ptr(gd), N.ptr(bd), N.ptr(wd), N.ptr(pos), B, *pd, 16, 1e-5)
    assert max_rel(pos.cpu().view(B, *pd, 3), pos_ref.detach()) < 1e-5
    dh = torch.empty_like(hd); dg = torch.zeros(16, device=DEV); db = torch.zeros(16, device=DEV)
    dw = torch.zeros(3, 16, device=DEV)
    N.call("mic_offset_head_bwd", N.ptr(gp.view(P, 3).contiguous().to(DEV)), N.ptr(hd), N.ptr(gd), N.ptr(bd), N.ptr(wd),
           N.ptr(dh), N.ptr(dg), N.ptr(db), N.ptr(dw), B, *pd, 16, 1e-5)
    assert rel_err(dh.cpu(), h.grad) < 2e-5 and rel_err(dg.cpu(), g.grad) < 2e-5
    assert rel_err(db.cpu(), b.grad) < 2e-5 and rel_err(dw.cpu(), w3.grad) < 2e-5


@pytest.mark.parametrize("dims,pd,scale", [((6, 7, 9), (6, 7, 9), 1.5), ((5, 5, 6), (7, 7, 7), 0.7), ((4, 4, 4), (4, 4, 4), 4.0)])
def test_deform_sample(dims, pd, scale):
    """vs the reference call (grid_sample) on the zero-padded source, fwd and bwd; covers far out-of-range offsets"""
    from micformer_b200 import _native as N
    B, C = 2, 8
    D, H, W = dims
    src = _rand(B, D, H, W, C, seed=21).requires_grad_(True)
    pos = (_rand(B, *pd, 3, seed=22) * scale).requires_grad_(True)
    srcp = F.pad(src, (0, 0, 0, pd[2] - W, 0, pd[1] - H, 0, pd[0] - D))
    out_ref = O.stn_sample(srcp.permute(0, 4, 1, 2, 3), pos.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
    go = _rand(B, *pd, C, seed=23)
    (out_ref * go).sum().backward()
    sd, pdv = src.detach().to(DEV), pos.detach().to(DEV)
    out = torch.empty(B, *pd, C, device=DEV)
    N.call("mic_deform_sample_fwd", N.ptr(sd), N.ptr(pdv), N.ptr(out), B, D, H, W, *pd, C)
    assert max_rel(out.cpu(), out_ref.detach()) < 1e-5
    dsrc = torch.zeros_like(sd); dpos = torch.empty_like(pdv)
    N.call("mic_deform_sample_bwd", N.ptr(go.to(DEV)), N.ptr(sd), N.ptr(pdv), N.ptr(dsrc), N.ptr(dpos), B, D, H, W, *pd, C)
    assert rel_err(dsrc.cpu(), src.grad) < 2e-5
    assert rel_err(dpos.cpu(), pos.grad) < 2e-4


def test_stn_module_matches_reference_golden():
    from micformer_b200.models.STN import SpatialTransformer
    g = torch.Generator().manual_seed(META["cases"]["stn"]["seed"])
    src = torch.randn(2, 5, 6, 7, 9, generator=g)
    flow = torch.randn(2, 3, 6, 7, 9, generator=g) * 1.5
    out = SpatialTransformer()(src.to(DEV), flow.to(DEV))
    assert max_rel(out.cpu(), VEC["stn/out"]) < 1e-5


def test_dice_loss_golden_and_grad():
    from micformer_b200.loss.dice import MDiceLoss
    g = torch.Generator().manual_seed(META["cases"]["dice"]["seed"])
    lg = torch.randn(2, 8, 6, 6, 6, generator=g) * 3
    lg[0, 0, 0, 0, :3] = torch.tensor([40.0, -40.0, 120.0])
    tg = F.one_hot(torch.randint(0, 8, (2, 6, 6, 6), generator=g), 8).permute(0, 4, 1, 2, 3).float()
    lgd = lg.to(DEV).requires_grad_(True)
    loss = MDiceLoss()(lgd, tg.to(DEV))
    loss.backward()
    assert abs(float(loss) - float(VEC["dice/loss"])) < 2e-6
    assert rel_err(lgd.grad.cpu(), VEC["dice/dlogits"]) < 1e-5


def test_dice_loss_soft_targets_and_odd_size():
    from micformer_b200.loss.dice import MDiceLoss
    lg = _rand(1, 3, 5, 7, 3, seed=1, scale=2.0).requires_grad_(True)
    tg = torch.rand(1, 3, 5, 7, 3, generator=torch.Generator().manual_seed(2))
    ref = O.mdice_loss(lg, tg)
    ref.backward()
    lgd = lg.detach().to(DEV).requires_grad_(True)
    loss = MDiceLoss()(lgd, tg.to(DEV))
    (2.0 * loss).backward()
    assert abs(float(loss) - float(ref)) < 2e-6 and rel_err(lgd.grad.cpu(), 2.0 * lg.grad) < 1e-5


# --------------------------------------------------------------------------------------------- block level
@pytest.mark.parametrize("name", BLOCKS)
def test_block_vs_reference_golden(name):
    from micformer_b200.models.MICFormer_self import CrossTransformerBlock3D, TransformerBlock3D
    c = META["cases"][name]
    cls = CrossTransformerBlock3D if c["cross"] else TransformerBlock3D
    blk = cls(dim=c["C"], num_heads=c["heads"], window_size=tuple(c["window"]), qkv_bias=True)
    blk.load_state_dict({k[4:]: v for k, v in block_weights(c["C"], c["cross"], c["seed"]).items()}, strict=True)
    blk = blk.to(DEV).eval()
    x, xa, gy = block_inputs(c["C"], c["dims"], c["seed"])
    x = x.to(DEV).requires_grad_(True); xa = xa.to(DEV).requires_grad_(True)
    y = blk(x, xa) if c["cross"] else blk(x)
    (y * gy.to(DEV)).sum().backward()
    assert max_rel(y.detach().cpu(), VEC[f"{name}/y"]) < 2e-5
    assert rel_err(x.grad.cpu(), VEC[f"{name}/dx"]) < 5e-5
    if c["cross"]:
        assert rel_err(xa.grad.cpu(), VEC[f"{name}/dxa"]) < 5e-5
    for k, v in blk.named_parameters():
        assert rel_err(v.grad.cpu(), VEC[f"{name}/grad/{k}"]) < 1e-4, k


def test_drop_path_train_mode_scales_branches():
    """timm DropPath semantics: per-sample Bernoulli(keep)/keep on each residual branch (M:419,424)."""
    from micformer_b200.models.MICFormer_self import TransformerBlock3D
    torch.manual_seed(0)
    blk = TransformerBlock3D(dim=24, num_heads=2, window_size=(2, 2, 2), qkv_bias=True, drop_path=0.5).to(DEV)
    x = _rand(4, 2, 2, 2, 24, seed=3).to(DEV)
    blk.eval()
    y_eval = blk(x)
    blk.train()
    seen = set()
    for _ in range(8):
        y = blk(x)
        for b in range(4):
            same_as_input = torch.allclose(y[b], x[b])
            seen.add(bool(same_as_input))
    assert seen == {True, False}          # some samples dropped both branches, some kept at least one
    assert not torch.allclose(y_eval, x)


# ------------------------------------------------------------------------------------------------ model level
def _build_head(cfg, sd):
    from micformer_b200.models.MICFormer_self import Head, MicFormer
    head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
    if tuple(cfg.depths) != (2, 2, 6, 2) or tuple(cfg.num_heads) != (3, 6, 12, 24):
        head.swin = MicFormer(window_size=cfg.window_size, in_chans=1, embed_dim=cfg.embed_dim, depths=list(cfg.depths),
                              num_heads=list(cfg.num_heads))
    head.load_state_dict(sd, strict=True)
    return head.to(DEV).eval()


def test_tiny64_whole_model_vs_reference_golden():
    from micformer_b200.loss.dice import MDiceLoss
    cfg = O.TINY
    head = _build_head(cfg, O.synth_state_dict(cfg, seed=3))
    x, lab = O.synth_inputs(1, 64, cfg.num_classes, seed=5)
    y = head(x.to(DEV))
    loss = MDiceLoss()(y, lab.to(DEV))
    loss.backward()
    yc = y.detach().cpu()
    assert max_rel(yc[:, :, ::4, ::4, ::4], VEC["tiny64/logits_sub"]) < 1e-4
    assert max_rel(yc[0, :, :6, :6, :6], VEC["tiny64/logits_corner"]) < 1e-4
    assert abs(float(loss) - float(VEC["tiny64/loss"])) < 2e-6
    hist = torch.bincount(yc.argmax(1).flatten(), minlength=cfg.num_classes).numpy()
    assert np.abs(hist - VEC["tiny64/argmax_hist"]).sum() <= 8
    grads = dict(head.named_parameters())
    for k, n in META["tiny64_grad_norms"].items():
        assert abs(float(grads[k].grad.norm()) - n) <= 1e-3 * n + 1e-9, k
    assert sorted(META["tiny64_no_grad"]) == sorted(k for k, v in grads.items() if v.grad is None)
    for key in VEC.files:
        if key.startswith("tiny64/grad/"):
            assert rel_err(grads[key[len("tiny64/grad/"):]].grad.cpu(), VEC[key]) < 1e-3, key


def test_train_config_default_init_known_answers_64():
    """torch.manual_seed(0) default init reproduces the reference's numbers through the CUDA path."""
    from micformer_b200.models.MICFormer_self import Head
    from micformer_b200.loss.dice import MDiceLoss
    ka = META["train64_default_init"]
    torch.manual_seed(0)
    head = Head(embed_dim=48, num_classes=8)
    assert abs(float(head.swin.patch_embed.proj.weight.sum()) - ka["patch_embed_weight_sum"]) < 1e-5
    head = head.to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 2, 64, 64, 64, generator=g)
    lab = F.one_hot(torch.randint(0, 8, (1, 64, 64, 64), generator=g), 8).permute(0, 4, 1, 2, 3).float()
    y = head(x.to(DEV))
    loss = MDiceLoss()(y, lab.to(DEV))
    loss.backward()
    yc = y.detach().cpu()
    assert max_rel(yc[:, :, ::8, ::8, ::8], VEC["train64/logits_sub"]) < 1e-4
    assert abs(float(yc.double().sum()) - ka["y_sum"]) < 1e-3 * ka["y_abs_sum"] * 1e-2
    assert abs(float(loss) - ka["loss"]) < 2e-6
    gl2 = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in head.parameters() if p.grad is not None)))
    assert abs(gl2 - ka["grad_l2"]) < 1e-3 * ka["grad_l2"]
    hist = torch.bincount(yc.argmax(1).flatten(), minlength=8).tolist()
    assert sum(abs(a - b) for a, b in zip(hist, ka["argmax_hist"])) <= 8


@pytest.mark.parametrize("cfgname,S", [("TRAIN", 128), ("W7", 64)])
def test_full_size_vs_oracle(cfgname, S):
    """BASELINE.json bar: logits within 1e-3 relative of the CPU path; argmax equal wherever the top-2 margin
    exceeds twice the measured logit error.  Oracle runs on the host cores (seconds)."""
    from micformer_b200.loss.dice import MDiceLoss
    cfg = getattr(O, cfgname)
    sd = O.synth_state_dict(cfg, seed=7)
    B = 1
    x, lab = O.synth_inputs(B, S, cfg.num_classes, seed=9)
    head = _build_head(cfg, sd)
    y = head(x.to(DEV))
    loss = MDiceLoss()(y, lab.to(DEV))
    loss.backward()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    logits, loss_ref, grads = O.train_step(x, lab, sd, cfg)
    yc = y.detach().cpu()
    err = float((yc - logits).abs().max())
    assert err / float(logits.abs().max()) < 1e-3
    top2 = logits.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    mism = (yc.argmax(1) != logits.argmax(1)) & (margin > 2 * err)
    assert int(mism.sum()) == 0
    assert abs(float(loss) - float(loss_ref)) < 1e-5
    # per-tensor relative error with an absolute floor tied to the global gradient norm: some tensors have
    # mathematically zero gradients (e.g. q of a window whose sampled keys are all identical) and hold only noise
    gl2 = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values() if g is not None)))
    worst, num_sq = 0.0, 0.0
    for k, p in head.named_parameters():
        if grads[k] is None:
            assert p.grad is None
            continue
        d = float((p.grad.cpu().double() - grads[k].double()).norm())
        e = d / (float(grads[k].double().norm()) + 1e-6 * gl2)
        # the offset net's gradient passes through d(trilinear)/d(position), which is discontinuous at voxel
        # boundaries: fp32 rounding differences flip a few cells at full size -> looser bound on that path only
        offset_path = "conv_offset" in k or ".norm1." in k
        assert e < (2e-2 if offset_path else 2e-3), (k, e)
        worst = max(worst, e)
        num_sq += d * d
    assert (num_sq ** 0.5) / gl2 < 1e-3


def test_state_dict_roundtrip_and_keys():
    from micformer_b200.models.MICFormer_self import Head
    head = Head(embed_dim=48, num_classes=8)
    assert list(head.state_dict().keys()) == list(O.param_shapes(O.TRAIN).keys())
    assert len(list(head.buffers())) == 0
