"""Data parallelism over volumes: one process per GPU, parameters replicated, ONE gradient all-reduce (mean) per
step over NCCL (NVLink 5 / NVSwitch) -- the path has no other exchange.  The reference itself is
single-GPU (train_mmwhs_noPad.py:413); this is what BASELINE.json's 1/2/4/8-GPU rows add.

``GradSync`` packs gradients into flat buckets (reverse registration order ~ the order backward produces them),
all-reduces each bucket asynchronously and scatters the averages back.  Parameters that never receive a gradient
(``swin.concat_back_dim.0.*`` is unused by the forward, reference M:1015-1016 / SURVEY F12) are skipped -- the
set must be identical on every rank, which is checked once.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params, bucket_bytes: int = 64 << 20, process_group=None, arena=None, reduce: str = "mean"):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.group = process_group
        self._plan = None      # list of lists of param indices
        self._flat = None
        self.arena = arena     # GradArena: the gradients already live in one flat buffer -> one in-place all-reduce
        if reduce not in ("mean", "sum"):
            raise ValueError("GradSync: reduce must be 'mean' (per-rank loss, the reference semantics) or 'sum' "
                             "(MDiceLoss(process_group=...): the loss already spans the global batch)")
        self.reduce = reduce
        self._early = None

    # ---- overlap: exchange the decoder's gradients while the encoder's backward is still running --------------------
    def enable_overlap(self, head) -> None:
        """With a segmented arena (``GradArena.for_model``): the decoder segment is all-reduced asynchronously as soon as the
        backward pass reaches the bottleneck (gradient hooks on the two ``norm`` outputs, installed by ``MicFormer._trunk``),
        the encoder segment in ``sync()``.  Works eagerly and inside CUDA-graph capture (the collective becomes a parallel
        branch of the graph)."""
        if self.arena is None or not getattr(self.arena, "segments", None):
            raise RuntimeError("GradSync.enable_overlap needs a segmented arena (GradArena.for_model)")
        swin = head.swin if hasattr(head, "swin") else head
        self._pending, self._early, self._hook_events = 0, None, []
        swin.__dict__["_bottleneck_grad_hook"] = self._on_bottleneck_grad

    def _on_bottleneck_grad(self, grad):
        if self.world <= 1:
            return grad
        from . import ops
        ops.flush_side_branches()                  # deferred weight-gradient branches of the decoder blocks join their streams first
        cur = torch.cuda.current_stream()
        self._pending += 1
        if self._pending < 2:                      # first of the two modality streams: remember where its backward stands
            ev = torch.cuda.Event()
            ev.record(cur)
            self._hook_events.append(ev)
            return grad
        for ev in self._hook_events:               # second one: both decoders' gradient kernels are enqueued
            cur.wait_event(ev)
        self._hook_events.clear()
        self._pending = 0
        a, b = self.arena.segments[0]
        self._early = self._reduce_(self.arena.flat[a:b], async_op=True)
        return grad

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _build_plan(self):
        live = [i for i, p in enumerate(self.params) if p.grad is not None]
        if dist.is_initialized() and self.world > 1:
            # every rank must skip the same parameters
            sig = torch.tensor([len(live), sum(live) % 1000003], dtype=torch.int64, device=self.params[0].device)
            lo, hi = sig.clone(), sig.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            if not torch.equal(lo, hi):
                raise RuntimeError("GradSync: ranks disagree on which parameters received gradients")
        plan, cur, size = [], [], 0
        for i in reversed(live):
            n = self.params[i].numel() * self.params[i].element_size()
            if cur and size + n > self.bucket_bytes:
                plan.append(cur); cur, size = [], 0
            cur.append(i); size += n
        if cur:
            plan.append(cur)
        self._plan = plan
        p0 = self.params[0]
        self._flat = [torch.empty(sum(self.params[i].numel() for i in b), device=p0.device, dtype=p0.dtype) for b in plan]

    def _reduce_(self, flat: torch.Tensor, async_op: bool = False):
        """in-place mean (or sum, ``reduce="sum"``: a loss that already spans the global batch) over the ranks.  NCCL
        averages inside the collective (ReduceOp.AVG): no separate 1/world pass over the 247 MB buffer."""
        if self.reduce == "sum":
            return dist.all_reduce(flat, group=self.group, async_op=async_op)
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        flat.mul_(1.0 / self.world)                      # gloo has no AVG
        return dist.all_reduce(flat, group=self.group, async_op=async_op)

    def sync(self) -> None:
        """Average ``.grad`` across ranks in place.  Call after ``backward()``."""
        if self.world <= 1:
            return
        if self.arena is not None and self.arena.attached():
            # parameters without a gradient (concat_back_dim.0) hold zeros on every rank: harmless to include
            early = getattr(self, "_early", None)
            if early is not None:                  # decoder segment already in flight: finish it, exchange the rest
                a, b = self.arena.segments[0]
                rest = self.arena.flat[b:] if a == 0 else torch.cat([self.arena.flat[:a], self.arena.flat[b:]])
                self._reduce_(rest)
                early.wait()
                self._early = None
            else:
                self._reduce_(self.arena.flat)
            return
        if self._plan is None:
            self._build_plan()
        works = []
        for flat, idxs in zip(self._flat, self._plan):
            grads = [self.params[i].grad for i in idxs]
            views = list(torch.split(flat, [g.numel() for g in grads]))
            torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
            works.append(self._reduce_(flat, async_op=True))
        for w, flat, idxs in zip(works, self._flat, self._plan):
            w.wait()
            grads = [self.params[i].grad for i in idxs]
            views = list(torch.split(flat, [g.numel() for g in grads]))
            for g, v in zip(grads, views):
                g.copy_(v.view(g.shape))                # handles non-contiguous .grad too

    def skipped(self) -> List[int]:
        live = set(i for b in (self._plan or []) for i in b)
        return [i for i in range(len(self.params)) if i not in live]
