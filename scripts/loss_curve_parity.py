"""BASELINE.json config 5 in miniature: the training loop (Head -> MDiceLoss -> backward -> Adam, train_mmwhs_noPad.py:
114,185-201) run for N steps by this package on cuda:0 and by the CPU oracle from the same initial state_dict on the same
synthetic labelled volumes; prints both loss curves and their largest relative deviation.

    python scripts/loss_curve_parity.py [--steps 100] [--size 64] [--batch 2] [--gemm-mode 1] [--out profiles/x.json]

DropPath is off on both sides (eval-mode forward inside a training loop): the two implementations draw their per-sample
masks from different random streams, so only the deterministic part of the step can be compared curve to curve.
Needs a GPU; not part of the test suites (the oracle leg alone takes ~0.3 s per step at 64^3, ~2 s at 128^3 on 16 threads)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--gemm-mode", type=int, default=1)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--tol", type=float, default=None, help="max relative loss deviation (default 2e-3 TF32, 1e-4 exact)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    from oracle import micformer_oracle as O
    from micformer_b200 import _native
    from micformer_b200.models.MICFormer_self import Head
    from micformer_b200.loss.dice import MDiceLoss
    from micformer_b200.optim import FusedAdam

    cfg = O.TRAIN
    sd = O.synth_state_dict(cfg, seed=0)
    data = [O.synth_inputs(a.batch, a.size, cfg.num_classes, seed=100 + s) for s in range(min(a.steps, 8))]   # cycled

    # ---- this package on the GPU
    _native.set_gemm_mode(a.gemm_mode)
    dev = torch.device("cuda", 0)
    head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
    head.load_state_dict(sd, strict=True)
    head = head.to(dev).eval()                      # eval(): DropPath off; gradients still flow
    opt = FusedAdam(head.parameters(), lr=a.lr, weight_decay=0.0)
    crit = MDiceLoss()
    ours = []
    for s in range(a.steps):
        x, lab = data[s % len(data)]
        opt.zero_grad(set_to_none=True)
        loss = crit(head(x.to(dev)), lab.to(dev))
        loss.backward()
        opt.step()
        ours.append(float(loss.detach()))

    # ---- the oracle on the host cores
    torch.set_num_threads(os.cpu_count() or 1)
    params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
    ropt = torch.optim.Adam(params.values(), lr=a.lr, weight_decay=0.0)
    ref = []
    for s in range(a.steps):
        x, lab = data[s % len(data)]
        ropt.zero_grad(set_to_none=True)
        loss = O.mdice_loss(O.head_forward(x, params, cfg, training=False), lab)
        loss.backward()
        ropt.step()
        ref.append(float(loss.detach()))

    dev_rel = [abs(o - r) / max(abs(r), 1e-12) for o, r in zip(ours, ref)]
    tol = a.tol if a.tol is not None else (2e-3 if a.gemm_mode == 1 else 1e-4)
    res = {"steps": a.steps, "size": a.size, "batch": a.batch, "gemm_mode": a.gemm_mode, "lr": a.lr,
           "max_rel_dev": max(dev_rel), "rel_dev_last": dev_rel[-1], "tolerance": tol, "loss_first": [ours[0], ref[0]],
           "loss_last": [ours[-1], ref[-1]], "ours": ours, "reference": ref}
    print(json.dumps({k: v for k, v in res.items() if k not in ("ours", "reference")}))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(res, f)
    sys.exit(0 if max(dev_rel) <= tol else 1)


if __name__ == "__main__":
    main()
