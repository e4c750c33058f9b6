"""SURVEY 8(f) rows on the GPU: the data feeder's sample contract into the CUDA path (fp16 image, bool one-hot labels read
as bytes by the loss kernels), the reference script's training iteration through the compat import paths, and the
validation-time sliding-window inference + meandice around the real model (fp16 input under autocast)."""
import os
import sys

import pytest
import torch

from oracle import micformer_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_feeder_contract_into_the_loss_kernels():
    """dataset contract (dataset/MMWHS.py:392,414-425): image fp16 (2,S,S,S), label bool (8,S,S,S).  The loss consumes the
    bool labels directly (1 byte per label) and agrees with the float path and with the oracle."""
    from micformer_b200.data import SyntheticMMWHS, train_transform
    from micformer_b200.loss.dice import MDiceLoss, MDiceLoss_Val
    ds = SyntheticMMWHS(n=2, size=32, transform=None)
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(3)
    batch = [train_transform({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in ds[i].items()}, gen) for i in range(2)]
    img = torch.stack([b["image"] for b in batch])             # (2, 2, 32, 32, 32) float32 on the device
    lab = torch.stack([b["label"] for b in batch])             # bool
    assert img.is_cuda and img.dtype == torch.float32 and lab.dtype == torch.bool
    logits = torch.randn(2, 8, 32, 32, 32, generator=torch.Generator().manual_seed(5)).to(dev).requires_grad_(True)
    l_bool = MDiceLoss()(logits, lab)
    g_bool, = torch.autograd.grad(l_bool, logits)
    l_f32 = MDiceLoss()(logits, lab.float())
    g_f32, = torch.autograd.grad(l_f32, logits)
    assert abs(float(l_bool.detach()) - float(l_f32.detach())) < 1e-7 and torch.equal(g_bool, g_f32)
    ref = O.mdice_loss(logits.detach().cpu(), lab.float().cpu())
    assert abs(float(l_bool.detach()) - float(ref)) < 2e-6
    # MDiceLoss_Val = Dice term only (loss/dice.py:216-221)
    lv = MDiceLoss_Val()(logits, lab)
    pr = torch.sigmoid(logits.detach().cpu().double())
    t = lab.cpu().double()
    dice = sum(1 - (2 * (pr[:, i] * t[:, i]).sum() + 1) / ((pr[:, i] ** 2).sum() + (t[:, i] ** 2).sum() + 1) for i in range(8)) / 8
    assert abs(float(lv.detach()) - float(dice)) < 2e-6


def test_reference_training_iteration_through_compat_paths():
    """what train_mmwhs_noPad.py:92,108-114,148,177-207 does, with the script's own import lines resolved through compat/"""
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        for m in [k for k in sys.modules if k == "loss" or k.startswith("loss.") or k == "models" or k.startswith("models.")]:
            del sys.modules[m]
        from loss import MDiceLoss                                              # train_mmwhs_noPad.py:19
        from loss.dice import MDiceLoss_Val                                     # :20
        from MMWHS_pre.Multi_modal.SymCFNet.models.MICFormer_self import Head   # :26
        from micformer_b200.optim import FusedAdam
        torch.manual_seed(0)
        model_1 = Head(embed_dim=24, num_classes=8).cuda()                      # :92 (narrower, same depths / heads)
        criterion, criterian_val = MDiceLoss().cuda(), MDiceLoss_Val().cuda()   # :108-109
        metric = criterian_val.metric                                           # :110
        optimizer = FusedAdam(model_1.parameters(), lr=1e-4, weight_decay=0.0)  # :114 (drop-in for torch.optim.Adam)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, 300)  # :148
        x, lab = O.synth_inputs(1, 64, 8, seed=2)
        losses = []
        for _ in range(3):
            inputs, labels = x.float().cuda(), lab.float().cuda()               # :177-181
            optimizer.zero_grad()
            segs = model_1(inputs)
            loss_ = criterion(segs, labels)
            losses.append(loss_.item())
            loss_.backward()
            optimizer.step()
            scheduler.step()                                                    # every iteration (:206-207)
        assert all(l == l for l in losses) and losses[-1] < losses[0]
        model_1.eval()
        with torch.no_grad():
            d = metric(model_1(inputs), labels)
        assert len(d) == 1 and len(d[0]) == 8
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))


def test_sliding_window_inference_and_meandice_on_the_model():
    """utils.py:222-240 + train_mmwhs_noPad.py:283-306,392-407: 128^3 windows at 50 % overlap over a 160x128x128 volume,
    fp16 input under autocast; equals the manual average of the per-window predictions; meandice on device == on host."""
    from micformer_b200.inference import inference, evaluate, window_lattice
    from micformer_b200.models.MICFormer_self import Head
    torch.manual_seed(0)
    model = Head(embed_dim=24, num_classes=8).cuda().eval()
    g = torch.Generator().manual_seed(4)
    vol = torch.randn(1, 2, 160, 128, 128, generator=g).half().cuda()
    with torch.no_grad(), torch.autocast("cuda"):
        out = inference(vol, model)
    assert out.shape == (1, 8, 160, 128, 128)
    corners = window_lattice((160, 128, 128), (128, 128, 128), 0.5)
    assert corners == [(0, 0, 0), (32, 0, 0)]
    acc = torch.zeros_like(out)
    cnt = torch.zeros(1, 1, 160, 128, 128, device="cuda")
    with torch.no_grad():
        for (z, y, w) in corners:
            acc[:, :, z:z + 128] += model(vol[:, :, z:z + 128])
            cnt[:, :, z:z + 128] += 1
    assert float((out - acc / cnt).abs().max()) < 1e-5
    lab = torch.nn.functional.one_hot(torch.randint(0, 8, (1, 160, 128, 128), generator=g), 8).permute(0, 4, 1, 2, 3).cuda()
    # (softmax on the device vs on the host may order exact near-ties differently: a handful of voxels of 2.6 M)
    assert abs(float(evaluate(out, lab)) - float(evaluate(out.cpu(), lab.cpu()))) < 1e-4
