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
    def __init__(self, params, segments=None):
        """``segments``: optional list of parameter lists that partition ``params``; the flat buffer is laid out segment by
        segment and ``self.segments`` holds their (start, end) offsets -- ``parallel.GradSync`` exchanges a segment as soon as
        the backward pass has finished producing it (``for_model``)."""
        if segments is not None:
            params = [p for seg in segments for p in seg]
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
        self.segments = None
        if segments is not None:
            self.segments, k = [], 0
            for seg in segments:
                cnt = sum(1 for p in seg if p.requires_grad)
                start = offs[k] if cnt else n
                k += cnt
                self.segments.append((start, offs[k] if k < len(offs) else n))
        self._views = []
        for p, o in zip(ps, offs):
            v = self.flat[o:o + p.numel()].view_as(p)
            p.grad = v
            p._mic_arena = True
            self._views.append(v)

    @classmethod
    def for_model(cls, head):
        """Head / MicFormer: [decoder + tail | encoder] -- the backward pass finishes the decoder's gradients (up_layers,
        concat_back_dim, norm2, reverse_patch_embedding, out_conv) when it reaches the bottleneck, long before the encoder's."""
        dec, enc = [], []
        for name, p in head.named_parameters():
            n = name[5:] if name.startswith("swin.") else name
            (dec if n.startswith(("up_layers.", "concat_back_dim.", "norm2.", "reverse_patch_embedding.", "out_conv.")) else enc).append(p)
        return cls(None, segments=[dec, enc])

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
