"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE on CPU.

    python tests/golden/make_golden.py          # needs /root/reference (build container only)

The reference has no tests/fixtures of its own for this path, so these vectors are what
pins the oracle (oracle/micformer_oracle.py) and, through it, the CUDA kernels.  Inputs and weights are
regenerated from seeds at test time (oracle.synth_inputs / synth_state_dict use per-key CPU generators,
deterministic across machines); only the reference OUTPUTS are stored.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _reference_loader import load_reference  # noqa: E402
from oracle import micformer_oracle as O      # noqa: E402  (only for seeded inputs/weights + shapes)

torch.set_num_threads(8)
M, Dice = load_reference()


def block_weights(C, cross, seed):
    shapes = O._block_shapes("blk", C, cross, 16, 4.0)
    out = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed(O._key_seed(key, seed))
        if ".norm" in key and key.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".norm" in key:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.2
        out[key[len("blk."):]] = t
    return out


def block_case(name, C, heads, window, dims, cross, seed):
    """One TransformerBlock3D / CrossTransformerBlock3D fwd+bwd; returns dict of np arrays."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, *dims, C, generator=g).requires_grad_(True)
    xa = torch.randn(1, *dims, C, generator=g).requires_grad_(True)
    gy = torch.randn(1, *dims, C, generator=g)
    cls = M.CrossTransformerBlock3D if cross else M.TransformerBlock3D
    blk = cls(dim=C, num_heads=heads, window_size=window, qkv_bias=True).eval()
    blk.load_state_dict(block_weights(C, cross, seed), strict=True)
    y = blk(x, xa) if cross else blk(x)
    (y * gy).sum().backward()
    out = {f"{name}/y": y.detach().numpy(), f"{name}/dx": x.grad.numpy()}
    if cross:
        out[f"{name}/dxa"] = xa.grad.numpy()
    for k, v in blk.named_parameters():
        out[f"{name}/grad/{k}"] = v.grad.numpy()
    return out


def main():
    vec = {}
    meta = {"torch": torch.__version__, "reference": "fxxJuses/MICFormer @ fc74cf0e", "cases": {}}

    # ---- block-level cases (window 2 no pad; window 7 with zero-pad; clamped window) -------------
    cases = [
        ("self_w2", 24, 2, (2, 2, 2), (4, 6, 8), False, 11),
        ("cross_w2", 24, 2, (2, 2, 2), (4, 6, 8), True, 12),
        ("self_w7pad", 32, 1, (7, 7, 7), (8, 9, 10), False, 13),
        ("cross_w7pad", 32, 1, (7, 7, 7), (8, 9, 10), True, 14),
        ("cross_w7clamp", 64, 2, (7, 7, 7), (4, 4, 4), True, 15),
        ("cross_w2_d16", 48, 3, (2, 2, 2), (8, 8, 8), True, 16),
    ]
    for name, C, heads, window, dims, cross, seed in cases:
        vec.update(block_case(name, C, heads, window, dims, cross, seed))
        meta["cases"][name] = dict(C=C, heads=heads, window=window, dims=dims, cross=cross, seed=seed)

    # ---- STN on a non-cubic volume -----------------------------------------------------------------
    g = torch.Generator().manual_seed(21)
    src = torch.randn(2, 5, 6, 7, 9, generator=g)           # (B,C,D,H,W)
    flow = torch.randn(2, 3, 6, 7, 9, generator=g) * 1.5
    vec["stn/out"] = M.SpatialTransformer()(src, flow).numpy()
    meta["cases"]["stn"] = dict(seed=21, src=[2, 5, 6, 7, 9], flow_scale=1.5)

    # ---- MDiceLoss incl. saturated logits ----------------------------------------------------------
    g = torch.Generator().manual_seed(22)
    lg = (torch.randn(2, 8, 6, 6, 6, generator=g) * 3).requires_grad_(True)
    with torch.no_grad():
        lg[0, 0, 0, 0, :3] = torch.tensor([40.0, -40.0, 120.0])   # sigmoid saturates -> BCE log clamp
    tg = torch.nn.functional.one_hot(torch.randint(0, 8, (2, 6, 6, 6), generator=g), 8).permute(0, 4, 1, 2, 3).float()
    loss = Dice.MDiceLoss()(lg, tg)
    loss.backward()
    vec["dice/loss"] = loss.detach().numpy()
    vec["dice/dlogits"] = lg.grad.numpy()
    meta["cases"]["dice"] = dict(seed=22)

    # ---- whole model, tiny 4-stage config at 64^3, synthetic weights --------------
    cfg = O.TINY
    sd = O.synth_state_dict(cfg, seed=3)
    ref = M.MicFormer(window_size=cfg.window_size, in_chans=1, embed_dim=cfg.embed_dim, depths=list(cfg.depths),
                      num_heads=list(cfg.num_heads)).eval()
    head = M.Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
    head.swin = ref
    head.eval()
    head.load_state_dict(sd, strict=True)
    x, lab = O.synth_inputs(1, 64, cfg.num_classes, seed=5)
    y = head(x)
    loss = Dice.MDiceLoss()(y, lab)
    loss.backward()
    vec["tiny64/logits_sub"] = y.detach()[:, :, ::4, ::4, ::4].numpy()
    vec["tiny64/logits_corner"] = y.detach()[0, :, :6, :6, :6].numpy()
    vec["tiny64/loss"] = loss.detach().numpy()
    vec["tiny64/argmax_hist"] = torch.bincount(y.argmax(1).flatten(), minlength=cfg.num_classes).numpy()
    gn = {k: float(v.grad.norm()) for k, v in head.named_parameters() if v.grad is not None}
    meta["tiny64_grad_norms"] = gn
    meta["tiny64_no_grad"] = [k for k, v in head.named_parameters() if v.grad is None]
    for k in ("swin.patch_embed.proj.weight", "swin.layers.0.blocks1.0.conv_offset.0.bias",
              "swin.layers.0.blocks1.0.conv_offset.3.weight", "swin.layers.3.self_blocks2.0.self_attn.kv.bias",
              "swin.up_layers.3.blocks2.1.mlp.fc2.bias", "swin.concat_back_dim.2.bias", "out_conv.weight",
              "swin.norm2.weight", "swin.layers.1.downsample.norm.weight"):
        vec[f"tiny64/grad/{k}"] = dict(head.named_parameters())[k].grad.numpy()
    meta["cases"]["tiny64"] = dict(weights_seed=3, input_seed=5, S=64, B=1)

    # ---- train config, default init (torch.manual_seed(0)), 64^3: SURVEY §8(c) known answers --------
    torch.manual_seed(0)
    head = M.Head(embed_dim=48, num_classes=8).eval()
    gg = torch.Generator().manual_seed(1)
    x = torch.randn(1, 2, 64, 64, 64, generator=gg)
    lab = torch.nn.functional.one_hot(torch.randint(0, 8, (1, 64, 64, 64), generator=gg), 8).permute(0, 4, 1, 2, 3).float()
    y = head(x)
    loss = Dice.MDiceLoss()(y, lab)
    loss.backward()
    gl2 = float(torch.sqrt(sum((v.grad.double() ** 2).sum() for v in head.parameters() if v.grad is not None)))
    meta["train64_default_init"] = dict(
        patch_embed_weight_sum=float(head.swin.patch_embed.proj.weight.sum()),
        y_sum=float(y.double().sum()), y_abs_sum=float(y.double().abs().sum()),
        y0=[float(v) for v in y[0, :, 0, 0, 0]], loss=float(loss), grad_l2=gl2,
        argmax_hist=[int(v) for v in torch.bincount(y.argmax(1).flatten(), minlength=8)])
    vec["train64/logits_sub"] = y.detach()[:, :, ::8, ::8, ::8].numpy()

    np.savez_compressed(os.path.join(HERE, "golden_vectors.npz"), **vec)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    sz = os.path.getsize(os.path.join(HERE, "golden_vectors.npz"))
    print(f"wrote {len(vec)} arrays, {sz / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
