"""Fused multi-tensor Adam (mic_adam_step) against torch.optim.Adam, including a missing gradient, odd sizes, a
learning-rate change between steps and state_dict interchange."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(48, 1, 4, 4, 4), (48,), (192, 48), (3, 16, 1, 1, 1), (70001,), (7,), (16, 96, 3, 3, 3)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]


def test_fused_adam_matches_torch_adam():
    from micformer_b200.optim import FusedAdam
    a, b = _params(0), _params(0)
    oa = torch.optim.Adam(a, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    ob = FusedAdam(b, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    g = torch.Generator().manual_seed(5)
    for it in range(6):
        if it == 3:                      # per-iteration LR schedule (train_mmwhs_noPad.py:206-207)
            oa.param_groups[0]["lr"] = 3e-4
            ob.param_groups[0]["lr"] = 3e-4
        for i, (pa, pb) in enumerate(zip(a, b)):
            if i == 5 and it < 2:        # a parameter that receives no gradient in some steps
                pa.grad = None; pb.grad = None
                continue
            gr = torch.randn(pa.shape, generator=g).cuda()
            pa.grad = gr.clone(); pb.grad = gr.clone()
        oa.step(); ob.step()
    for pa, pb in zip(a, b):
        assert float((pa - pb).abs().max()) < 2e-6
    # optimizer checkpoints are interchangeable with torch.optim.Adam
    sd = ob.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    oc = torch.optim.Adam(_params(0), lr=1e-3)
    oc.load_state_dict(sd)
    assert float(oc.state_dict()["state"][2]["exp_avg"].sub(sd["state"][2]["exp_avg"]).abs().max()) == 0.0


def test_fused_adam_inside_cuda_graph():
    from micformer_b200.optim import FusedAdam
    a, b = _params(1), _params(1)
    oa = torch.optim.Adam(a, lr=1e-3, capturable=True)
    ob = FusedAdam(b, lr=1e-3)
    grads = [torch.randn_like(p) for p in a]
    for pa, pb, gr in zip(a, b, grads):
        pa.grad = gr.clone(); pb.grad = gr.clone()
    ob.step(); oa.step()                                   # warm-up (tables, lr on device)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ob.step()
    for _ in range(3):
        graph.replay()
    for _ in range(4):                                     # capture executed once (not run) + 3 replays = 3 steps
        pass
    for _ in range(3):
        oa.step()
    torch.cuda.synchronize()
    for pa, pb in zip(a, b):
        assert float((pa - pb).abs().max()) < 2e-6


@pytest.mark.parametrize("mode", [0, 1])
def test_grad_arena_direct_accumulation_matches_autograd(mode):
    """With a GradArena attached, backward writes straight into the flat gradient buffer: same gradients as the
    ordinary autograd path, accumulation across two backward passes, one memset to clear."""
    from micformer_b200 import _native as N
    from micformer_b200.arena import GradArena
    from micformer_b200.models.MICFormer_self import Head, MicFormer
    from micformer_b200.loss.dice import MDiceLoss
    from oracle import micformer_oracle as O
    prev = N.get_gemm_mode()
    N.set_gemm_mode(mode)
    try:
        cfg = O.TINY
        sd = O.synth_state_dict(cfg, seed=3)
        head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
        head.swin = MicFormer(window_size=cfg.window_size, in_chans=1, embed_dim=cfg.embed_dim, depths=list(cfg.depths),
                              num_heads=list(cfg.num_heads))
        head.load_state_dict(sd, strict=True)
        head = head.cuda().eval()
        x, lab = O.synth_inputs(1, 64, cfg.num_classes, seed=5)
        x, lab = x.cuda(), lab.cuda()
        MDiceLoss()(head(x), lab).backward()
        ref = {k: (p.grad.clone() if p.grad is not None else None) for k, p in head.named_parameters()}
        for p in head.parameters():
            p.grad = None
        arena = GradArena(head.parameters())
        MDiceLoss()(head(x), lab).backward()
        assert arena.attached()
        gl2 = float(sum((g.double() ** 2).sum() for g in ref.values() if g is not None) ** 0.5)
        # mode 0 is deterministic up to fp32 atomics order.  Mode 1: every kernel is (scripts/determinism.py: <= 5e-7 run to
        # run), but single-pass TF32 backward operands turn that reordering noise into rounding flips, and the offset branch
        # (d pos = differences of neighbouring voxels, scripts/grad_noise.py) amplifies them: up to 2.4e-2 run to run on
        # conv_offset / norm1 tensors whose norm is 1e-3 of the global gradient norm, <= 2e-3 elsewhere
        for k, p in head.named_parameters():
            tol = 1e-4 if mode == 0 else (6e-2 if ("conv_offset" in k or "norm1" in k) else 2e-2)
            if ref[k] is None:
                assert float(p.grad.abs().max()) == 0.0, k
            else:
                assert float((p.grad - ref[k]).norm() / (ref[k].norm() + 1e-5 * gl2)) < tol, k
        once = arena.flat.clone()
        MDiceLoss()(head(x), lab).backward()       # gradients accumulate, as with zero_grad(set_to_none=False)
        assert float((arena.flat - 2 * once).norm() / once.norm()) < (1e-4 if mode == 0 else 2e-2)
        arena.zero()
        assert float(arena.flat.abs().max()) == 0.0 and arena.attached()
    finally:
        N.set_gemm_mode(prev)
