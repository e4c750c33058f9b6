"""BASELINE.json config 5 in miniature: the reference's training loop (Head -> MDiceLoss -> backward -> Adam ->
CosineAnnealingLR stepped EVERY ITERATION, train_mmwhs_noPad.py:114,148,185-207) run for N steps by this
package on cuda:0 and by the CPU oracle from the same initial state_dict on the same synthetic labelled volumes; prints
both loss curves and their largest relative deviation.

    python scripts/loss_curve_parity.py [--steps 100] [--size 64] [--batch 2] [--gemm-mode 1] [--droppath 1] [--out x.json]

--droppath 1 (default): both sides run in train mode (drop_path_rate 0.2 as the reference hard-codes, M:917,941) with the
SAME per-sample masks: they are drawn from one CPU generator in the reference's draw order and handed to both
implementations (micformer_b200/testing.py).  --droppath 0: eval-mode forward inside the training loop.
Needs a GPU; the oracle leg takes ~0.3 s per step at 64^3, ~4 s at 128^3 batch 2 on 16 threads."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def run(steps, size, batch, gemm_mode, lr=1e-4, epochs=300, droppath=True, cfgname="TRAIN", quiet=False):
    from oracle import micformer_oracle as O
    from micformer_b200 import _native
    from micformer_b200.models.MICFormer_self import Head, MicFormer
    from micformer_b200.loss.dice import MDiceLoss
    from micformer_b200.optim import FusedAdam
    from micformer_b200.testing import share_drop_path_masks

    cfg = getattr(O, cfgname)
    sd = O.synth_state_dict(cfg, seed=0)
    data = [O.synth_inputs(batch, size, cfg.num_classes, seed=100 + s) for s in range(min(steps, 8))]   # cycled

    # ---- this package on the GPU
    prev = _native.get_gemm_mode()
    _native.set_gemm_mode(gemm_mode)
    dev = torch.device("cuda", 0)
    head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
    if tuple(cfg.depths) != (2, 2, 6, 2) or tuple(cfg.num_heads) != (3, 6, 12, 24):
        head.swin = MicFormer(window_size=cfg.window_size, in_chans=1, embed_dim=cfg.embed_dim, depths=list(cfg.depths),
                              num_heads=list(cfg.num_heads))
    head.load_state_dict(sd, strict=True)
    head = head.to(dev).train(droppath)
    opt = FusedAdam(head.parameters(), lr=lr, weight_decay=0.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, epochs)
    crit = MDiceLoss()
    gen = torch.Generator().manual_seed(1234)
    ours, lrs = [], []
    for s in range(steps):
        x, lab = data[s % len(data)]
        if droppath:
            share_drop_path_masks(head, gen, batch, dev)
        opt.zero_grad(set_to_none=True)
        loss = crit(head(x.to(dev)), lab.to(dev))
        loss.backward()
        opt.step()
        lrs.append(opt.param_groups[0]["lr"])
        sched.step()                                   # every iteration, like the reference
        ours.append(float(loss.detach()))
    _native.set_gemm_mode(prev)

    # ---- the oracle on the host cores
    torch.set_num_threads(os.cpu_count() or 1)
    params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
    ropt = torch.optim.Adam(params.values(), lr=lr, weight_decay=0.0)
    rsched = torch.optim.lr_scheduler.CosineAnnealingLR(ropt, epochs)
    gen = torch.Generator().manual_seed(1234)
    ref = []
    for s in range(steps):
        x, lab = data[s % len(data)]
        ropt.zero_grad(set_to_none=True)
        loss = O.mdice_loss(O.head_forward(x, params, cfg, training=droppath, gen=gen), lab)
        loss.backward()
        ropt.step()
        rsched.step()
        ref.append(float(loss.detach()))

    dev_rel = [abs(o - r) / max(abs(r), 1e-12) for o, r in zip(ours, ref)]
    res = {"config": cfgname, "steps": steps, "size": size, "batch": batch, "gemm_mode": gemm_mode, "lr0": lr,
           "scheduler": f"CosineAnnealingLR(T_max={epochs}) stepped every iteration", "lr_last": lrs[-1],
           "droppath_shared_masks": bool(droppath), "max_rel_dev": max(dev_rel), "rel_dev_last": dev_rel[-1],
           "loss_first": [ours[0], ref[0]], "loss_last": [ours[-1], ref[-1]], "ours": ours, "reference": ref}
    if not quiet:
        print(json.dumps({k: v for k, v in res.items() if k not in ("ours", "reference")}))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--gemm-mode", type=int, default=1)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--epochs", type=int, default=300, help="T_max of the cosine schedule (the script's --epochs default)")
    ap.add_argument("--droppath", type=int, default=1)
    ap.add_argument("--config", default="TRAIN")
    ap.add_argument("--tol", type=float, default=None, help="max relative loss deviation (default 2e-3 tensor-core, 1e-4 exact)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = run(a.steps, a.size, a.batch, a.gemm_mode, a.lr, a.epochs, bool(a.droppath), a.config)
    tol = a.tol if a.tol is not None else (2e-3 if a.gemm_mode == 1 else 1e-4)
    res["tolerance"] = tol
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(res, f)
    sys.exit(0 if res["max_rel_dev"] <= tol else 1)


if __name__ == "__main__":
    main()
