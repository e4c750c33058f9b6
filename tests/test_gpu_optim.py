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
