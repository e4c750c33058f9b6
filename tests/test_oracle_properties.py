"""Size-independent properties of the path, checked on the CPU oracle (they are what lets the GPU tests trust it at sizes
where no golden vector exists, and what the data-parallel sharding of SURVEY 8e rests on)."""
import pytest
import torch

from oracle import micformer_oracle as O


@pytest.mark.parametrize("dims,ws", [((4, 6, 8), (2, 2, 2)), ((7, 14, 21), (7, 7, 7)), ((4, 8, 8), (4, 8, 8)), ((6, 6, 4), (3, 2, 4))])
def test_window_partition_reverse_is_a_permutation(dims, ws):
    B, C = 2, 5
    x = torch.arange(B * dims[0] * dims[1] * dims[2] * C, dtype=torch.float32).view(B, *dims, C)
    w = O.window_partition(x, ws)
    assert w.shape == (B * (dims[0] // ws[0]) * (dims[1] // ws[1]) * (dims[2] // ws[2]), ws[0] * ws[1] * ws[2], C)
    assert torch.equal(O.window_reverse(w, ws, B, *dims), x)                 # round trip is exact
    assert torch.equal(w.flatten().sort().values, x.flatten())              # no element lost or duplicated
    # token (iz, iy, ix) of window (wz, wy, wx) is grid position (wz*wd+iz, wy*wh+iy, wx*ww+ix): the address rule the
    # CUDA kernels implement instead of materialising windows
    nwh, nww = dims[1] // ws[1], dims[2] // ws[2]
    wz, wy, wx, iz, iy, ix = 1 % (dims[0] // ws[0]), nwh - 1, nww - 1, ws[0] - 1, 0, ws[2] - 1
    win = (wz * nwh + wy) * nww + wx
    tok = (iz * ws[1] + iy) * ws[2] + ix
    assert torch.equal(w[win, tok], x[0, wz * ws[0] + iz, wy * ws[1] + iy, wx * ws[2] + ix])


def test_samples_are_independent_in_eval_mode():
    """Every op is per-sample (LayerNorm, windows, sampler): sharding the batch over ranks changes nothing but the loss
    sums.  This is the premise of the data-parallel partitioning (no data-path collective)."""
    cfg = O.TINY
    sd = O.synth_state_dict(cfg, seed=11)
    x, _ = O.synth_inputs(2, 64, cfg.num_classes, seed=12)
    with torch.no_grad():
        both = O.head_forward(x, sd, cfg)
        one = O.head_forward(x[1:2], sd, cfg)
    assert float((both[1:2] - one).abs().max()) < 1e-5 * float(both.abs().max())


def test_mdice_partial_sums_reproduce_the_loss_and_its_closed_form_gradient():
    """The fused loss kernel reduces 4 sums per channel (mdice_sums) and differentiates in closed form; both must equal
    autograd through the reference formulation (loss/dice.py:130-166)."""
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(2, 4, 6, 5, 7, generator=g) * 3).requires_grad_(True)
    target = (torch.rand(2, 4, 6, 5, 7, generator=g) > 0.7).float()
    loss = O.mdice_loss(logits, target)
    loss.backward()
    S = O.mdice_sums(logits.detach(), target)                # (C, 4): sum p t, sum p^2, sum t^2, sum bce
    C = target.shape[1]
    n = target.numel() / C
    inter, pp, tt, bce = S[:, 0], S[:, 1], S[:, 2], S[:, 3]
    den = pp + tt + 1.0
    from_sums = (0.7 * (1 - (2 * inter + 1.0) / den).sum() + 0.3 * (bce / n).sum()) / C
    assert abs(float(from_sums) - float(loss.detach())) < 1e-6
    p = torch.sigmoid(logits.detach())
    a = (2.0 / den).view(1, C, 1, 1, 1)
    b = (2 * (2 * inter + 1.0) / den ** 2).view(1, C, 1, 1, 1)
    dl = (0.7 * (-(a * target) + b * p) * p * (1 - p) + 0.3 * (p - target) / n) / C
    assert float((dl - logits.grad).abs().max()) < 1e-6 * max(1.0, float(logits.grad.abs().max()) * 1e3)


def test_drop_path_is_per_sample_and_unbiased():
    g = torch.Generator().manual_seed(0)
    x = torch.ones(4000, 3, 2)
    y = O._drop_path(x, 0.2, True, g)
    per_sample = y.view(4000, -1)
    assert bool(((per_sample == 0).all(1) | (per_sample == 1.25).all(1)).all())          # whole sample kept (1/keep) or dropped
    assert abs(float(y.mean()) - 1.0) < 0.03
    assert O._drop_path(x, 0.2, False, g) is x and O._drop_path(x, 0.0, True, g) is x      # identity in eval / rate 0


def test_window_clamp_and_pad_rules():
    assert O.get_window_size((4, 4, 4), (7, 7, 7)) == (4, 4, 4)
    assert O.get_window_size((8, 3, 9), (7, 7, 7)) == (7, 3, 7)
    x = torch.randn(1, 8, 8, 8, 3)
    xp = O._pad_to_window(x, (7, 7, 7))
    assert xp.shape == (1, 14, 14, 14, 3) and torch.equal(xp[:, :8, :8, :8], x) and float(xp[:, 8:].abs().max()) == 0.0
