"""Fused multi-tensor Adam on the C ABI (`mic_adam_step`): the reference's `torch.optim.Adam(lr=1e-4, weight_decay=0)`
stepped every iteration (MicFormer/train_mmwhs_noPad.py:114,201) as ONE kernel over all 1626
tensors.  `state_dict()` keeps torch.optim.Adam's layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) so optimizer
checkpoints written by the reference script (`utils.py:57-65`) load here and vice-versa.  The step count and the
learning rate live on the device, so `step()` is CUDA-graph capturable; per-iteration LR schedules
(`CosineAnnealingLR.step()` every iteration, :148,206-207) are honoured by re-reading `param_groups[0]['lr']`.
"""
from __future__ import annotations

import torch

from . import _native as N


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdam: one parameter group (the reference passes model.parameters())")
        self._tables = None
        self._grad_ptrs_host = None
        self._lr_on_device = None
        self._stage_busy = None        # event: the last non-blocking H2D copy of the pinned grad-pointer table
        self.arena = None

    def attach_arena(self, arena) -> None:
        """``micformer_b200.arena.GradArena``: ``zero_grad()`` then clears the flat buffer with one memset and keeps the
        parameters pointing into it (whatever ``set_to_none`` says)."""
        self.arena = arena

    def zero_grad(self, set_to_none: bool = True) -> None:
        if self.arena is not None:
            self.arena.zero()
            return
        super().zero_grad(set_to_none=set_to_none)

    def load_state_dict(self, state_dict) -> None:
        """torch.optim.Adam checkpoints (and our own) load here; the device tables that cache raw pointers of the moment
        buffers and the step counters are rebuilt from the loaded state on the next ``step()`` / ``set_lr()``."""
        super().load_state_dict(state_dict)
        self._tables = None
        self._grad_ptrs_host = None
        self._lr_on_device = None

    # ---- state in torch.optim.Adam's layout --------------------------------------------------------------
    def _init_state(self):
        g = self.param_groups[0]
        ps = [p for p in g["params"] if p.requires_grad]
        for p in ps:
            N.check_cuda_f32(p)
        dev = ps[0].device
        self._params = ps
        self._lr = torch.zeros(1, device=dev, dtype=torch.float32)
        self._steps = torch.zeros(len(ps), device=dev, dtype=torch.float32)     # one counter per parameter, like torch
        for i, p in enumerate(ps):
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            else:                                          # loaded from a torch.optim.Adam checkpoint (or our own)
                self._steps[i] = float(st.get("step", 0.0))
                for k in ("exp_avg", "exp_avg_sq"):
                    if st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                        st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()
            st["step"] = self._steps[i]                    # 0-dim view into the device counters
        chunk = N.load().mic_adam_chunk_elems()
        ct, ci = [], []
        for t, p in enumerate(ps):
            n = p.numel()
            for c in range((n + chunk - 1) // chunk):
                ct.append(t); ci.append(c)
        i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
        self._tables = dict(
            params=i64([p.data_ptr() for p in ps]), exp_avg=i64([self.state[p]["exp_avg"].data_ptr() for p in ps]),
            exp_avg_sq=i64([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]), sizes=i64([p.numel() for p in ps]),
            chunk_tensor=torch.tensor(ct, dtype=torch.int32, device=dev),
            chunk_index=torch.tensor(ci, dtype=torch.int32, device=dev), grads=torch.zeros(len(ps), dtype=torch.int64, device=dev))
        self._grad_stage = torch.zeros(len(ps), dtype=torch.int64).pin_memory()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self._tables is None:
            self._init_state()
        g = self.param_groups[0]
        t = self._tables
        ptrs = [0 if p.grad is None else p.grad.data_ptr() for p in self._params]
        if ptrs != self._grad_ptrs_host:
            for p in self._params:
                if p.grad is not None:
                    N.check_cuda_f32(p.grad)
            capturing = torch.cuda.is_current_stream_capturing()
            if self._stage_busy is not None and not capturing:
                self._stage_busy.synchronize()      # the previous async copy must have read the pinned table
            self._grad_stage.copy_(torch.tensor(ptrs, dtype=torch.int64))
            t["grads"].copy_(self._grad_stage, non_blocking=True)
            if not capturing:
                self._stage_busy = torch.cuda.Event()
                self._stage_busy.record()
            else:
                self._stage_busy = None             # a captured copy re-reads the pinned table at every replay: it stays as is
            self._grad_ptrs_host = ptrs
        lr = float(g["lr"])
        if lr != self._lr_on_device:
            self._lr.fill_(lr)
            self._lr_on_device = lr
        b1, b2 = g["betas"]
        N.call("mic_adam_step", N.ptr(t["params"]), N.ptr(t["grads"]), N.ptr(t["exp_avg"]), N.ptr(t["exp_avg_sq"]),
               N.ptr(t["sizes"]), N.ptr(t["chunk_tensor"]), N.ptr(t["chunk_index"]), int(t["chunk_tensor"].numel()),
               len(self._params), N.ptr(self._steps), N.ptr(self._lr), float(b1), float(b2), float(g["eps"]),
               float(g["weight_decay"]))
        return loss

    def set_lr(self, lr: float) -> None:
        """Device-side learning-rate update that does not break a captured graph (call between replays)."""
        self.param_groups[0]["lr"] = lr
        if self._tables is None:
            self._init_state()
        self._lr.fill_(float(lr))
        self._lr_on_device = float(lr)
