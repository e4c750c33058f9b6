"""Hooks for parity runs (tests/, scripts/): nothing here is on the product path."""
from __future__ import annotations

import torch


def share_drop_path_masks(head, gen: torch.Generator, batch: int, device) -> None:
    """Draw the DropPath masks of one training-mode forward from a CPU generator in exactly the order the reference (and
    the oracle, oracle/micformer_oracle.py:_drop_path) draws them -- per stage, per depth: self block CT, self block MR,
    cross block CT, cross block MR, two draws each (attention branch, MLP branch), none for blocks with rate 0 -- and
    hand them to the blocks, so that both implementations apply the same per-sample scales."""
    swin = head.swin if hasattr(head, "swin") else head
    swin.__dict__["_dp_external"] = True
    for layer in list(swin.layers) + list(swin.up_layers):
        for j in range(len(layer.blocks1)):
            for blk in (layer.self_blocks1[j], layer.self_blocks2[j], layer.blocks1[j], layer.blocks2[j]):
                rate = float(getattr(blk.drop_path, "drop_prob", 0.0))
                if rate == 0.0 or not swin.training:
                    blk.__dict__["_dp_scales"] = (None, None)
                    continue
                keep = 1.0 - rate
                draws = []
                for _ in range(2):
                    m = torch.empty((batch, 1, 1, 1, 1), dtype=torch.float32).bernoulli_(keep, generator=gen)
                    draws.append(m.div_(keep).view(batch).to(device))
                blk.__dict__["_dp_scales"] = tuple(draws)
