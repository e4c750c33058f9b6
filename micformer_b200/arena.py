"""Gradient arena: every parameter's ``.grad`` is a slice of ONE flat fp32 buffer.

The weight-gradient kernels of this package accumulate (TMA reduce-add, atomics), so with the arena attached
``backward()`` writes straight into ``p.grad`` (``ops._acc``): one memset per step instead of ~200 small zero fills, no
AccumulateGrad copies/adds, shared modules (the down/up-sample layers and ``concat_back_dim`` serve both modality
streams, reference M:689-707,1027-1030) accumulate in place, and the data-parallel all-reduce runs in place on the flat
buffer (``parallel.GradSync``).  Semantics are those of ``optimizer.zero_grad(set_to_none=False)``: gradients accumulate
across ``backward()`` calls until ``zero()``; code that drops the grads (``set_to_none=True``) simply falls back to
the ordinary autograd path.
"""
from __future__ import annotations

import torch


class GradArena:
    def __init__(self, params):
        ps = [p for p in params if p.requires_grad]
        if not ps:
            raise ValueError("GradArena: no trainable parameters")
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != ps[0].device:
                raise RuntimeError("GradArena: parameters must be contiguous float32 tensors on one device")
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + 63) // 64 * 64                 # 256-byte aligned slices (TMA reduce-add, float4 atomics)
        self.flat = torch.zeros(n, device=ps[0].device, dtype=torch.float32)
        self.params = ps
        self._views = []
        for p, o in zip(ps, offs):
            v = self.flat[o:o + p.numel()].view_as(p)
            p.grad = v
            p._mic_arena = True
            self._views.append(v)

    def attached(self) -> bool:
        """every parameter still points into the arena (nobody called zero_grad(set_to_none=True))"""
        return all(p.grad is v for p, v in zip(self.params, self._views))

    def reattach(self) -> None:
        for p, v in zip(self.params, self._views):
            p.grad = v

    def zero(self) -> None:
        if not self.attached():
            self.reattach()
        self.flat.zero_()
