"""The oracle (oracle/micformer_oracle.py) replayed against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  This is what pins the oracle; the CUDA tests then compare the
kernels with the oracle."""
import numpy as np
import pytest
import torch

from oracle import micformer_oracle as O
from helpers import load_golden, block_weights, block_inputs, rel_err, max_rel

VEC, META = load_golden()
BLOCKS = [k for k in META["cases"] if k.startswith(("self_", "cross_"))]


@pytest.mark.parametrize("name", BLOCKS)
def test_block_matches_reference(name):
    c = META["cases"][name]
    p = {k: v.requires_grad_(True) for k, v in block_weights(c["C"], c["cross"], c["seed"]).items()}
    x, xa, gy = block_inputs(c["C"], c["dims"], c["seed"])
    x.requires_grad_(True); xa.requires_grad_(True)
    if c["cross"]:
        y = O.cross_block(x, xa, p, "blk", c["heads"], tuple(c["window"]))
    else:
        y = O.self_block(x, p, "blk", c["heads"], tuple(c["window"]))
    (y * gy).sum().backward()
    assert max_rel(y.detach(), VEC[f"{name}/y"]) < 2e-6
    assert rel_err(x.grad, VEC[f"{name}/dx"]) < 1e-5
    if c["cross"]:
        assert rel_err(xa.grad, VEC[f"{name}/dxa"]) < 1e-5
    for k, v in p.items():
        assert rel_err(v.grad, VEC[f"{name}/grad/{k[4:]}"]) < 2e-5, k


@pytest.mark.parametrize("name", [n for n in BLOCKS if n.startswith("cross_")])
def test_closed_form_sampler_equals_grid_sample_path(name):
    c = META["cases"][name]
    p = block_weights(c["C"], True, c["seed"])
    x, xa, _ = block_inputs(c["C"], c["dims"], c["seed"])
    y = O.cross_block(x, xa, p, "blk", c["heads"], tuple(c["window"]), closed_form_stn=True)
    assert max_rel(y, VEC[f"{name}/y"]) < 5e-6


def test_stn_non_cubic():
    g = torch.Generator().manual_seed(META["cases"]["stn"]["seed"])
    src = torch.randn(2, 5, 6, 7, 9, generator=g)
    flow = torch.randn(2, 3, 6, 7, 9, generator=g) * 1.5
    ref = VEC["stn/out"]
    assert max_rel(O.stn_sample(src, flow), ref) < 1e-6
    cf = O.stn_sample_closed_form(src.permute(0, 2, 3, 4, 1).contiguous(), flow.permute(0, 2, 3, 4, 1).contiguous())
    assert max_rel(cf.permute(0, 4, 1, 2, 3), ref) < 5e-6


def test_ref_points_permuted_normalisers():
    # MICFormer_self.py:333-335 divides the z channel by H, y by W, x by D
    r = O.ref_points(4, 6, 8)
    assert abs(float(r[1, 0, 0, 0]) - ((1.5 / 6) * 2 - 1)) < 1e-6
    assert abs(float(r[0, 2, 0, 1]) - ((2.5 / 8) * 2 - 1)) < 1e-6
    assert abs(float(r[0, 0, 3, 2]) - ((3.5 / 4) * 2 - 1)) < 1e-6


def test_dice_loss_with_saturation():
    g = torch.Generator().manual_seed(META["cases"]["dice"]["seed"])
    lg = torch.randn(2, 8, 6, 6, 6, generator=g) * 3
    lg[0, 0, 0, 0, :3] = torch.tensor([40.0, -40.0, 120.0])
    lg.requires_grad_(True)
    tg = torch.nn.functional.one_hot(torch.randint(0, 8, (2, 6, 6, 6), generator=g), 8).permute(0, 4, 1, 2, 3).float()
    loss = O.mdice_loss(lg, tg)
    loss.backward()
    assert abs(float(loss) - float(VEC["dice/loss"])) < 1e-6
    assert rel_err(lg.grad, VEC["dice/dlogits"]) < 1e-6
    # the partial-sum formulation used by the fused kernel gives the same loss
    s = O.mdice_sums(lg.detach(), tg).double()
    n = lg[:, 0].numel()
    alt = (0.7 * (1 - (2 * s[:, 0] + 1) / (s[:, 1] + s[:, 2] + 1)).sum() + 0.3 * (s[:, 3] / n).sum()) / 8
    assert abs(float(alt) - float(VEC["dice/loss"])) < 1e-5


def test_tiny64_whole_model():
    cfg = O.TINY
    sd = O.synth_state_dict(cfg, seed=3)
    x, lab = O.synth_inputs(1, 64, cfg.num_classes, seed=5)
    logits, loss, grads = O.train_step(x, lab, sd, cfg)
    assert max_rel(logits[:, :, ::4, ::4, ::4], VEC["tiny64/logits_sub"]) < 5e-6
    assert max_rel(logits[0, :, :6, :6, :6], VEC["tiny64/logits_corner"]) < 5e-6
    assert abs(float(loss) - float(VEC["tiny64/loss"])) < 1e-6
    hist = torch.bincount(logits.argmax(1).flatten(), minlength=cfg.num_classes).numpy()
    assert np.abs(hist - VEC["tiny64/argmax_hist"]).sum() <= 4
    for k, n in META["tiny64_grad_norms"].items():
        assert abs(float(grads[k].norm()) - n) <= 2e-4 * n + 1e-9, k
    # concat_back_dim[0] is never used by the forward -> no gradient
    assert sorted(META["tiny64_no_grad"]) == sorted(k for k, g in grads.items() if g is None)
    for key in VEC.files:
        if key.startswith("tiny64/grad/"):
            assert rel_err(grads[key[len("tiny64/grad/"):]], VEC[key]) < 2e-4, key


def test_param_shapes_match_reference_structure():
    # 1626 tensors, 61,722,608 parameters for the train config
    s = O.param_shapes(O.TRAIN)
    assert len(s) == 1626
    assert sum(int(np.prod(v)) for v in s.values()) == 61_722_608
    assert sum(int(np.prod(v)) for v in O.param_shapes(O.Config(embed_dim=96, num_classes=14)).values()) == 231_137_342


@pytest.mark.reference
def test_oracle_equals_live_reference_train64():
    """Build container only: run the unmodified reference and the oracle side by side (default init)."""
    from _reference_loader import reference_available, load_reference
    if not reference_available():
        pytest.skip("/root/reference not present")
    M, _ = load_reference()
    torch.manual_seed(0)
    head = M.Head(embed_dim=48, num_classes=8).eval()
    ka = META["train64_default_init"]
    assert abs(float(head.swin.patch_embed.proj.weight.sum()) - ka["patch_embed_weight_sum"]) < 1e-5
    assert abs(ka["patch_embed_weight_sum"] - (-1.4963787198)) < 1e-5      # SURVEY §8(c) known answer
    assert abs(ka["loss"] - 0.67536324) < 2e-6 and abs(ka["y_sum"] - 3264.836295) < 0.05
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 2, 64, 64, 64, generator=g)
    with torch.no_grad():
        y_ref = head(x)
        y_or = O.head_forward(x, head.state_dict(), O.TRAIN)
    assert max_rel(y_or, y_ref) < 5e-6
    assert max_rel(y_or[:, :, ::8, ::8, ::8], VEC["train64/logits_sub"]) < 5e-6


@pytest.mark.reference
def test_oracle_equals_live_reference_on_odd_sizes():
    """PatchEmbed3D / PatchMerging padding (M:864-869, 551-555) and the decoder's trilinear resize
    (M:1018-1025) -- a 70 x 72 x 66 volume pads to 72 x 72 x 68, stages 18x18x17 -> 9x9x9 -> 5^3 -> 3^3, and every skip
    connection but the first needs the resize.  Forward and gradients of the restatement vs the unmodified reference."""
    import dataclasses
    from _reference_loader import reference_available, load_reference
    if not reference_available():
        pytest.skip("/root/reference not present")
    ref_models, ref_loss = load_reference()
    torch.manual_seed(0)
    m = ref_models.Head(embed_dim=24, num_classes=8).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 2, 70, 72, 66, generator=g)
    y = m(x)
    assert tuple(y.shape) == (1, 8, 72, 72, 68)
    lab = torch.nn.functional.one_hot(torch.randint(0, 8, (1, 72, 72, 68), generator=g), 8).permute(0, 4, 1, 2, 3).float()
    ref_loss.MDiceLoss()(y, lab).backward()
    cfg = dataclasses.replace(O.TRAIN, embed_dim=24)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    logits, loss, grads = O.train_step(x, lab, sd, cfg)
    assert float((logits - y.detach()).abs().max() / y.detach().abs().max()) < 1e-5
    gl2 = float(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None) ** 0.5)
    for k, p in m.named_parameters():
        if p.grad is None:
            assert grads[k] is None, k
        else:
            assert float((grads[k] - p.grad).norm() / (p.grad.norm() + 1e-6 * gl2)) < 1e-3, k
