"""Host side of the fused tcgen05 block kernels (``csrc/block_mlp*.cu``, ``csrc/block_attn*.cu``).

The fused kernels read their weights as pre-swizzled bf16 hi/lo shared-memory IMAGES (``include/micformer_b200.h``,
"Fused block kernels").  ``WeightImages`` owns the image buffer of one module's weights and the job records that
``mic_weight_images`` turns into images; ``refresh_all`` converts every registered module of a model in ONE launch
(the parameters change every optimizer step, so the model refreshes once per forward; a block used stand-alone
refreshes its own images).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N

FUSED_C = (24, 48)           # channel counts the fused kernels are built for


def enabled() -> bool:
    """fused kernels are part of the tensor-core mode (gemm mode 1); MICFORMER_FUSED=0 keeps the unfused tcgen05 path"""
    import os
    return N.get_gemm_mode() == 1 and os.environ.get("MICFORMER_FUSED", "1") != "0"


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class WeightImages:
    """Images of a fixed list of weight views.  spec: (name, tensor, N, K, transpose, n_pad) with the image being the B
    operand B[n][k] = transpose ? W[k][n] : W[n][k] of a row-major fp32 weight ``tensor`` (leading dimension = its row
    stride)."""

    def __init__(self, specs: Sequence[Tuple[str, torch.Tensor, int, int, bool, int]]):
        self.specs = list(specs)
        dev = self.specs[0][1].device
        off = 0
        self.offsets = {}
        rows = []
        self.max_chunks = 0
        for name, t, n, k, tr, n_pad in self.specs:
            if not (t.is_cuda and t.dtype == torch.float32 and t.stride(-1) == 1):
                raise RuntimeError("WeightImages: weights must be CUDA float32 with a unit inner stride")
            panels = (k + 63) // 64
            nbytes = panels * n_pad * 128
            self.offsets[name] = (off, off + nbytes, nbytes)        # hi offset, lo offset, bytes per image
            rows.append([t.data_ptr(), off, off + nbytes, t.stride(0), n, k, int(tr), n_pad, panels])
            self.max_chunks = max(self.max_chunks, panels * n_pad * 8)
            off += 2 * nbytes
        self.buf = torch.empty(_ceil(off, 256), dtype=torch.uint8, device=dev)
        base = self.buf.data_ptr()
        for r in rows:
            r[1] += base
            r[2] += base
        self.jobs_host = rows
        self.jobs = torch.tensor(rows, dtype=torch.int64, device=dev)
        self._ptrs = [t.data_ptr() for _, t, *_ in self.specs]
        self.stamp = -1          # model-level refresh counter this buffer was last converted at

    def valid(self) -> bool:
        return [t.data_ptr() for _, t, *_ in self.specs] == self._ptrs

    def hi(self, name: str) -> int:
        return self.buf.data_ptr() + self.offsets[name][0]

    def lo(self, name: str) -> int:
        return self.buf.data_ptr() + self.offsets[name][1]

    def refresh(self) -> None:
        N.call("mic_weight_images", N.ptr(self.jobs), len(self.jobs_host), self.max_chunks)


def refresh_all(images: List[WeightImages]) -> None:
    """one launch for the images of all ``images`` (they must live on one device)"""
    if not images:
        return
    key = tuple(id(i) for i in images)
    cache = refresh_all._cache
    ent = cache.get(key)
    if ent is None:
        jobs = torch.cat([i.jobs for i in images], dim=0).contiguous()
        ent = cache[key] = (jobs, max(i.max_chunks for i in images), images)
        if len(cache) > 8:
            cache.pop(next(iter(cache)))
    jobs, mc, _ = ent
    N.call("mic_weight_images", N.ptr(jobs), jobs.shape[0], mc)


refresh_all._cache = {}


def mlp_images(fc1_w: torch.Tensor, fc2_w: torch.Tensor) -> WeightImages:
    """fc1 (4C, C), fc2 (C, 4C) -> the four images the fused MLP kernels read (forward: w1_nk, w2_nk; backward adds the
    transposed views w2_kn, w1_kn)."""
    hid, c = fc1_w.shape
    cp = _ceil(c, 16)
    hp = _ceil(hid, 64)                             # the backward walks the hidden axis in chunks of 64 rows
    return WeightImages([
        ("w1_nk", fc1_w, hid, c, False, hp),        # fc1:           B[n = hidden][k = c]      = W1[n][k]
        ("w2_nk", fc2_w, c, hid, False, cp),        # fc2:           B[n = c][k = hidden]      = W2[n][k]
        ("w2_kn", fc2_w, hid, c, True, hp),         # dh = dy W2:    B[n = hidden][k = c]      = W2[k][n]
        ("w1_kn", fc1_w, c, hid, True, cp),         # dxn = dh W1:   B[n = c][k = hidden]      = W1[k][n]
    ])


class ConvLayouts:
    """The two operand layouts the 3x3x3 conv kernels read -- ``tcio`` (27, Cin, Co) and ``toci`` (27, Co, Cin) -- of one
    nn.Conv3d(k=3) weight (Co, Cin, 3, 3, 3), kept in persistent buffers and rewritten once per model forward by
    ``mic_conv_weight_layouts`` (one launch for all convolutions of a model) instead of two permuted copies per block."""

    def __init__(self, weight: torch.Tensor):
        if not (weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous() and tuple(weight.shape[2:]) == (3, 3, 3)):
            raise RuntimeError("ConvLayouts: contiguous CUDA float32 (Co, Cin, 3, 3, 3) weight expected")
        co, cin = weight.shape[:2]
        self.weight = weight
        self.tcio = torch.empty(27, cin, co, device=weight.device)
        self.toci = torch.empty(27, co, cin, device=weight.device)
        self.elems = 27 * cin            # grid bound: 32-channel chunks of the widest job
        self.co = co
        self.job_host = [weight.data_ptr(), self.tcio.data_ptr(), self.toci.data_ptr(), cin, co]
        self.job = torch.tensor([self.job_host], dtype=torch.int64, device=weight.device)
        self._ptr = weight.data_ptr()
        self.stamp = -1

    def valid(self) -> bool:
        return self.weight.data_ptr() == self._ptr

    def refresh(self) -> None:
        N.call("mic_conv_weight_layouts", N.ptr(self.job), 1, self.elems, self.co)


def refresh_conv_layouts(convs: List["ConvLayouts"]) -> None:
    if not convs:
        return
    key = tuple(id(c) for c in convs)
    ent = refresh_conv_layouts._cache.get(key)
    if ent is None:
        jobs = torch.cat([c.job for c in convs], dim=0).contiguous()
        ent = refresh_conv_layouts._cache[key] = (jobs, (max(c.elems for c in convs), max(c.co for c in convs)), convs)
        if len(refresh_conv_layouts._cache) > 8:
            refresh_conv_layouts._cache.pop(next(iter(refresh_conv_layouts._cache)))
    jobs, (me, mco), _ = ent
    N.call("mic_conv_weight_layouts", N.ptr(jobs), jobs.shape[0], me, mco)


refresh_conv_layouts._cache = {}


class QkvCat:
    """q.weight (C, C) | kv.weight (2C, C) -> one (3C, C) weight and (3C,) bias, so that the unfused self-attention blocks
    run ONE q|k|v GEMM forward and ONE data-gradient GEMM backward (two launches fewer on each block's dependent chain).
    The buffers are persistent; all blocks of a model are refreshed by one multi-tensor copy per forward."""

    def __init__(self, qw, qb, kvw, kvb):
        c = qw.shape[0]
        self.w = torch.empty(3 * c, qw.shape[1], device=qw.device)
        self.b = torch.empty(3 * c, device=qw.device) if qb is not None else None
        self.srcs = [qw, kvw] + ([qb, kvb] if qb is not None else [])
        self.dsts = [self.w[:c], self.w[c:]] + ([self.b[:c], self.b[c:]] if qb is not None else [])
        self._ptrs = [t.data_ptr() for t in self.srcs]
        self.stamp = -1

    def valid(self) -> bool:
        return [t.data_ptr() for t in self.srcs] == self._ptrs

    def refresh(self) -> None:
        torch._foreach_copy_(self.dsts, self.srcs)


def refresh_qkv_cats(cats: List["QkvCat"]) -> None:
    if cats:
        torch._foreach_copy_([d for c in cats for d in c.dsts], [s for c in cats for s in c.srcs])


_epoch = 0        # bumped by every model-level refresh; an image set whose stamp equals it is current


def epoch() -> int:
    return _epoch


def model_refresh(images: List[WeightImages], convs: Optional[List[ConvLayouts]] = None,
                  cats: Optional[List[QkvCat]] = None) -> None:
    """Called once per model forward (the optimizer changed the weights): one launch converts all images, one more
    rewrites the operand layouts of all 3x3x3 conv weights."""
    global _epoch
    _epoch += 1
    if images:
        refresh_all(images)
        for i in images:
            i.stamp = _epoch
    if convs:
        refresh_conv_layouts(convs)
        for c in convs:
            c.stamp = _epoch
    if cats:
        refresh_qkv_cats(cats)
        for c in cats:
            c.stamp = _epoch


def ensure_current(img) -> None:
    """A block used stand-alone (no enclosing model refreshed this forward) converts its own images / layouts."""
    if img is not None and img.stamp != _epoch:
        img.refresh()


def mlp_supported(c: int, hid: int) -> bool:
    """resident-weight kernels for C in FUSED_C (stage 0), hidden-split streaming kernels for 64 <= C <= 384 (stages 1-3)"""
    return enabled() and hid == 4 * c and (c in FUSED_C or mlp_split(c))


def mlp_split(c: int) -> bool:
    """MICFORMER_FUSED_SPLIT=<min C> enables the hidden-split kernels for channel counts >= <min C> (0 / unset: off)"""
    import os
    lo = int(os.environ.get("MICFORMER_FUSED_SPLIT", "0") or 0)
    return lo > 0 and max(lo, 64) <= c <= 384 and c % 8 == 0 and (4 * c) % 64 == 0


def mlp_block_fwd(x: torch.Tensor, img: WeightImages, gamma, beta, b1, b2, rowscale: Optional[torch.Tensor], rps: int,
                  eps: float) -> torch.Tensor:
    """y = x + rowscale * fc2(GELU(fc1(LN(x))));  x (..., C) contiguous fp32"""
    c = x.shape[-1]
    t = x.numel() // c
    if c not in FUSED_C:          # deep stages: hidden-split kernel accumulates into a zeroed output
        y = torch.zeros_like(x)
        N.call("mic_mlp_split_fwd", N.ptr(x), N.ptr(y), N.ptr(gamma), N.ptr(beta), N.ptr(b1), N.ptr(b2), img.hi("w1_nk"),
               img.lo("w1_nk"), img.hi("w2_nk"), img.lo("w2_nk"), N.ptr(rowscale), int(rps), t, c, 4 * c, float(eps))
        return y
    y = torch.empty_like(x)
    N.call("mic_mlp_block_fwd", N.ptr(x), N.ptr(y), N.ptr(gamma), N.ptr(beta), N.ptr(b1), N.ptr(b2), img.hi("w1_nk"),
           img.lo("w1_nk"), img.hi("w2_nk"), img.lo("w2_nk"), N.ptr(rowscale), int(rps), t, c, float(eps))
    return y


def mlp_block_bwd(dy: torch.Tensor, x: torch.Tensor, img: WeightImages, gamma, beta, b1, rowscale: Optional[torch.Tensor],
                  rps: int, eps: float, dgamma, dbeta, dW1, db1, dW2, db2) -> torch.Tensor:
    """dx = dy + d(branch)/dx; the six parameter gradients are ACCUMULATED into the given buffers."""
    c = x.shape[-1]
    t = x.numel() // c
    dx = torch.empty_like(x)
    N.call("mic_mlp_block_bwd", N.ptr(dy), N.ptr(x), N.ptr(dx), N.ptr(gamma), N.ptr(beta), N.ptr(b1), img.hi("w1_nk"),
           img.lo("w1_nk"), img.hi("w2_kn"), img.lo("w2_kn"), img.hi("w1_kn"), img.lo("w1_kn"), N.ptr(rowscale), int(rps),
           N.ptr(dW1), N.ptr(db1), N.ptr(dW2), N.ptr(db2), N.ptr(dgamma), N.ptr(dbeta), t, c, float(eps))
    return dx


# ---------------------------------------------------------------------------------------------- attention half-block
FUSED_ATTN = ((48, 16), (48, 24))      # (C, head_dim) the fused attention kernels are built for


def attn_images(q_w: torch.Tensor, kv_w: torch.Tensor, p_w: torch.Tensor) -> WeightImages:
    """q (C, C), kv (2C, C), proj (C, C) -> forward images (wq_nk, wkv_nk, wp_nk) and the transposed views of the backward"""
    c = q_w.shape[0]
    cp = _ceil(c, 16)
    return WeightImages([
        ("wq_nk", q_w, c, c, False, cp),
        ("wkv_nk", kv_w, 2 * c, c, False, 2 * c),
        ("wp_nk", p_w, c, c, False, cp),
        ("wp_kn", p_w, c, c, True, cp),             # do  = dy Wp:        B[n = c_in][k = c_out] = Wp[k][n]
        ("wq_kn", q_w, c, c, True, cp),             # dxn = dq Wq
        ("wkv_kn", kv_w, c, 2 * c, True, cp),       # dsrc = [dk|dv] Wkv: B[n = c_in][k = 2C]
    ])


def attn_supported(c: int, heads: int, window, dims) -> bool:
    return (enabled() and heads > 0 and c % heads == 0 and (c, c // heads) in FUSED_ATTN and tuple(window) == (2, 2, 2)
            and all(d % 2 == 0 for d in dims))


def attn_block_fwd(x: torch.Tensor, kvsrc: Optional[torch.Tensor], img: WeightImages, gamma, beta, bq, bkv, bp,
                   rowscale: Optional[torch.Tensor], heads: int, eps: float) -> torch.Tensor:
    """x1 = x + rowscale * proj(window_attention(q(LN x), kv(kvsrc or LN x)));  x (B, D, H, W, C)"""
    b, d, h, w, c = x.shape
    y = torch.empty_like(x)
    N.call("mic_attn_block_fwd", N.ptr(x), N.ptr(kvsrc), N.ptr(y), N.ptr(gamma), N.ptr(beta), N.ptr(bq), N.ptr(bkv), N.ptr(bp),
           img.hi("wq_nk"), img.lo("wq_nk"), img.hi("wkv_nk"), img.lo("wkv_nk"), img.hi("wp_nk"), img.lo("wp_nk"),
           N.ptr(rowscale), b, d, h, w, c, heads, float(c // heads) ** -0.5, float(eps))
    return y


def attn_block_bwd(dy: torch.Tensor, x: torch.Tensor, kvsrc: Optional[torch.Tensor], img: WeightImages, gamma, beta, bq, bkv,
                   rowscale: Optional[torch.Tensor], heads: int, eps: float, dgamma, dbeta, dWq, dbq, dWkv, dbkv, dWp, dbp):
    """-> (dx, dkvsrc or None); the eight parameter gradients are ACCUMULATED into the given buffers."""
    import ctypes
    b, d, h, w, c = x.shape
    dx = torch.empty_like(x)
    dsrc = torch.empty_like(x) if kvsrc is not None else None
    names = ("wq_nk", "wkv_nk", "wp_kn", "wq_kn", "wkv_kn")
    arr = (ctypes.c_void_p * 10)(*[f(nm) for nm in names for f in (img.hi, img.lo)])
    N.call("mic_attn_block_bwd", N.ptr(x), N.ptr(kvsrc), N.ptr(dy), N.ptr(dx), N.ptr(dsrc), N.ptr(gamma), N.ptr(beta), N.ptr(bq),
           N.ptr(bkv), ctypes.cast(arr, ctypes.c_void_p), N.ptr(rowscale), N.ptr(dgamma), N.ptr(dbeta), N.ptr(dWq), N.ptr(dbq),
           N.ptr(dWkv), N.ptr(dbkv), N.ptr(dWp), N.ptr(dbp), b, d, h, w, c, heads, float(c // heads) ** -0.5, float(eps))
    return dx, dsrc


def mlp_split_bwd(dy: torch.Tensor, x: torch.Tensor, img: WeightImages, gamma, beta, b1, rowscale: Optional[torch.Tensor],
                  rps: int, eps: float, dW1, db1, dW2, db2):
    """deep-stage MLP backward up to the LayerNorm: -> (dxn, mean, rstd); fc1 / fc2 gradients are ACCUMULATED.  The caller
    finishes with the LayerNorm backward kernel (dx = dy + LN'(dxn), dgamma, dbeta)."""
    c = x.shape[-1]
    t = x.numel() // c
    dxn = torch.zeros_like(x)
    mean = torch.empty(t, device=x.device, dtype=torch.float32)
    rstd = torch.empty(t, device=x.device, dtype=torch.float32)
    N.call("mic_mlp_split_bwd", N.ptr(dy), N.ptr(x), N.ptr(dxn), N.ptr(mean), N.ptr(rstd), N.ptr(gamma), N.ptr(beta), N.ptr(b1),
           img.hi("w1_nk"), img.lo("w1_nk"), img.hi("w2_kn"), img.lo("w2_kn"), img.hi("w1_kn"), img.lo("w1_kn"),
           N.ptr(rowscale), int(rps), N.ptr(dW1), N.ptr(db1), N.ptr(dW2), N.ptr(db2), t, c, 4 * c, float(eps))
    return dxn, mean, rstd
