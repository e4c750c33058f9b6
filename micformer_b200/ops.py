"""Host-side operators: ``torch.autograd.Function``s that sequence the sm_100a kernels of the C ABI.

torch is plumbing here (device memory from its caching allocator, the current stream, autograd's tape for
ordering); every arithmetic step of the hot path is one of the ``mic_*`` kernels.  Nothing in this file has a
CPU or torch-op fallback.

Layout: activations are channels-last token grids ``(B, D, H, W, C)`` fp32 contiguous.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import _native as N
from . import fused as F_

LN_EPS = 1e-5
Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------------------
def window_geometry(dims: Sequence[int], window: Sequence[int]):
    """(use_window, padded dims) -- get_window_size clamp (reference M:135-145) + trailing pad (M:345-348)."""
    ws = tuple(d if d <= w else w for d, w in zip(dims, window))
    pdims = tuple(d + (w - d % w) % w for d, w in zip(dims, ws))
    return ws, pdims


def _empty(shape, like: Tensor) -> Tensor:
    return torch.empty(shape, device=like.device, dtype=torch.float32)


# Zero-initialised gradient buffers (atomically accumulated by the kernels) are carved out of ONE zero-filled
# arena per backward call: one fill kernel instead of ~20 per block.
_arena = None


class zero_arena:
    def __init__(self, nfloats: int, like: Tensor):
        self.n, self.like = int(nfloats), like

    def __enter__(self):
        global _arena
        self.prev = _arena
        _arena = [torch.zeros(self.n, device=self.like.device, dtype=torch.float32), 0]

    def __exit__(self, *exc):
        global _arena
        _arena = self.prev


def _zeros(shape, like: Tensor) -> Tensor:
    n = 1
    for d in (shape if isinstance(shape, (tuple, list)) else (shape,)):
        n *= d
    if _arena is not None:
        off = _arena[1]
        end = off + ((n + 63) // 64) * 64            # 256-byte aligned slices (TMA / float4 friendly)
        if end <= _arena[0].numel():
            _arena[1] = end
            return _arena[0][off:off + n].view(shape)
    return torch.zeros(shape, device=like.device, dtype=torch.float32)


# ----------------------------------------------------------------------------------------------------------
# gradient arena: when ``micformer_b200.arena.GradArena`` is attached, every parameter's ``.grad`` is a slice of one
# flat buffer that is cleared with ONE memset per step.  The weight-gradient kernels all accumulate (TMA reduce-add /
# atomics), so backward writes straight into ``p.grad`` and returns None for that input: no per-block zero fills, no
# AccumulateGrad copies / adds for shared modules, and the data-parallel all-reduce runs in place on the flat buffer.
# ----------------------------------------------------------------------------------------------------------
def _acc(p) -> Optional[Tensor]:
    """``p.grad`` when it is arena-backed (accumulate into it directly), else None"""
    if p is not None and getattr(p, "_mic_arena", False):
        g = p.grad
        if g is not None and g.is_contiguous() and g.dtype == torch.float32:
            return g
    return None


def _gret(p, t):
    """what backward returns for parameter ``p`` whose gradient was produced in ``t``"""
    return None if _acc(p) is not None else t


# ----------------------------------------------------------------------------------------------------------
# weight-gradient side branch: dW / db kernels do not feed the data-gradient chain, so each backward forks them
# onto an auxiliary stream (one per calling stream) and joins before returning.  Under CUDA-graph capture this
# becomes a parallel branch; buffers are allocated on the calling stream before the fork and the join precedes
# any later reuse, so no record_stream bookkeeping is needed.
# ----------------------------------------------------------------------------------------------------------
import os as _os

_AUX_ON = _os.environ.get("MICFORMER_AUX_STREAM", "1") != "0"
_Q_SIDE = _os.environ.get("MICFORMER_Q_SIDE", "1") != "0"      # unfused cross blocks: q GEMM / q data gradient beside the offset branch
_AUX = {}


_DEFER_ON = _os.environ.get("MICFORMER_DEFER_JOIN", "0") == "1"      # measured: no gain (21.32 vs 21.24 ms/step), off by default
_PENDING = {}          # calling stream -> (stream, event of its side branch's last kernel, tensors those kernels read)
_FLUSH_QUEUED = [False]


def flush_side_branches() -> None:
    """Join every deferred side branch into the stream that forked it (end of backward; before a gradient exchange)."""
    _FLUSH_QUEUED[0] = False
    here = torch.cuda.current_stream() if _PENDING else None
    for key in list(_PENDING):
        cur, ev, keep = _PENDING.pop(key)
        cur.wait_event(ev)
        # the engine callback runs on the caller's stream AFTER autograd has joined its leaf streams into it: join directly too
        if here is not None and here != cur:
            here.wait_event(ev)
        keep.clear()


def _inline(fn, *args, **kw):
    return fn(*args, **kw)


class side_branch:
    """with side_branch() as sb: sb.run(fn, *args) ... ; joins on exit.

    ``defer=True`` (block backward functions, whose parameters are used once per forward): the join is postponed by one
    block -- the exit only joins the PREVIOUS deferred branch of this stream, so the weight-gradient kernels of block k
    run under the data-gradient chain of block k-1; the last one is joined by an autograd-engine callback at the end of
    the backward pass (``flush_side_branches``).  The tensors the side kernels read stay referenced until their join."""

    def __init__(self, defer: bool = False):
        self.defer = defer and _DEFER_ON and _AUX_ON

    def __enter__(self):
        self.cur = torch.cuda.current_stream()
        self.used = False
        self.keep = []          # inputs of side-branch kernels stay referenced until the join: the caching
                                # allocator must not hand their memory to later main-branch kernels
        if _AUX_ON:
            key = (self.cur.device_index, self.cur.cuda_stream)
            aux = _AUX.get(key)
            if aux is None:
                aux = _AUX[key] = torch.cuda.Stream(device=self.cur.device)
            self.aux = aux
            self.key = key
        else:
            self.aux = None
        return self

    def run(self, fn, *args, **kw):
        if self.aux is None:
            return fn(*args, **kw)
        self.aux.wait_stream(self.cur)
        with torch.cuda.stream(self.aux):
            out = fn(*args, **kw)
        self.used = True
        return out

    def hold(self, *tensors):
        self.keep.extend(t for t in tensors if t is not None)

    def mark(self):
        """event after the side kernels launched so far (None without an auxiliary stream): ``wait(mark)`` joins just those"""
        return self.aux.record_event() if (self.aux is not None and self.used) else None

    def wait(self, ev) -> None:
        if ev is not None:
            self.cur.wait_event(ev)

    def __exit__(self, *exc):
        if self.aux is None or not self.used:
            self.keep.clear()
            return
        if self.defer and exc[0] is None:
            try:
                if not _FLUSH_QUEUED[0]:
                    torch.autograd.Variable._execution_engine.queue_callback(flush_side_branches)
                    _FLUSH_QUEUED[0] = True
            except RuntimeError:           # not inside an autograd backward pass: nothing would join the branch later
                self.cur.wait_stream(self.aux)
                self.keep.clear()
                return
            prev = _PENDING.pop(self.key, None)
            if prev is not None:
                self.cur.wait_event(prev[1])
                prev[2].clear()
            _PENDING[self.key] = (self.cur, self.aux.record_event(), self.keep)
            self.keep = []
            return
        prev = _PENDING.pop(self.key, None)      # an immediate join also covers anything deferred earlier on this stream
        if prev is not None:
            prev[2].clear()
        self.cur.wait_stream(self.aux)
        self.keep.clear()


# ----------------------------------------------------------------------------------------------------------
# thin kernel wrappers (no autograd)
# ----------------------------------------------------------------------------------------------------------
def ln_fwd(x0: Tensor, x1: Optional[Tensor], gamma: Tensor, beta: Tensor, dims, pdims=None):
    """x0 (B,D,H,W,C0) [| x1 (B,D,H,W,C1)] -> y (B,Dp,Hp,Wp,C0+C1) zero padded, mean, rstd (rows)."""
    B, D, H, W = dims
    Dp, Hp, Wp = pdims if pdims is not None else (D, H, W)
    C0 = x0.shape[-1]
    C1 = x1.shape[-1] if x1 is not None else 0
    y = _empty((B, Dp, Hp, Wp, C0 + C1), x0)
    mean = _empty((B * D * H * W,), x0)
    rstd = _empty((B * D * H * W,), x0)
    N.call("mic_layernorm_fwd", N.ptr(x0), C0, N.ptr(x1), C1, N.ptr(gamma), N.ptr(beta), N.ptr(y), N.ptr(mean),
           N.ptr(rstd), B, D, H, W, Dp, Hp, Wp, LN_EPS)
    return y, mean, rstd


def ln_bwd(dy: Tensor, x0: Tensor, x1: Optional[Tensor], gamma: Tensor, mean: Tensor, rstd: Tensor,
           dres0: Optional[Tensor], dres1: Optional[Tensor], dims, pdims=None, beta=None):
    """``beta``: the bias Parameter (only needed to find its arena gradient; dgamma goes to ``gamma``'s)."""
    B, D, H, W = dims
    Dp, Hp, Wp = pdims if pdims is not None else (D, H, W)
    C0 = x0.shape[-1]
    C1 = x1.shape[-1] if x1 is not None else 0
    dx0 = torch.empty_like(x0)
    dx1 = torch.empty_like(x1) if x1 is not None else None
    dgamma = _acc(gamma) if _acc(gamma) is not None else _zeros((C0 + C1,), x0)
    dbeta = _acc(beta) if _acc(beta) is not None else _zeros((C0 + C1,), x0)
    N.call("mic_layernorm_bwd", N.ptr(dy), N.ptr(x0), C0, N.ptr(x1), C1, N.ptr(gamma), N.ptr(mean), N.ptr(rstd),
           N.ptr(dres0), N.ptr(dres1), N.ptr(dx0), N.ptr(dx1), N.ptr(dgamma), N.ptr(dbeta), B, D, H, W, Dp, Hp, Wp)
    return dx0, dx1, _gret(gamma, dgamma), _gret(beta, dbeta)


def _view_ptr(t: Tensor, col: int) -> int:
    """pointer to column ``col`` of a 2-D row-major buffer"""
    return t.data_ptr() + 4 * col


def linear_fwd(X: Tensor, ldx: int, W: Tensor, bias: Optional[Tensor], M: int, Nout: int, K: int, *, out: Tensor = None,
               out_col: int = 0, ldy: int = None, w_is_kn: bool = False, ldw: int = None, w_col: int = 0, x_col: int = 0,
               act: bool = False, pre: Tensor = None, res: Tensor = None, rowscale: Tensor = None, rps: int = 0,
               accumulate: bool = False, view=None):
    if out is None:
        out = _empty((M, Nout), X)
        ldy = Nout
    if ldw is None:
        ldw = Nout if w_is_kn else K
    if view is not None:
        N.unpatch_view(*view)          # Y lives as the fine channels-last grid (decoder tail)
    N.call("mic_linear_fwd", _view_ptr(X, x_col), ldx, _view_ptr(W, w_col), ldw, int(w_is_kn), N.ptr(bias),
           _view_ptr(out, out_col), ldy, M, Nout, K, int(act), N.ptr(pre), Nout, N.ptr(res), Nout, N.ptr(rowscale), rps,
           int(accumulate))
    return out


def linear_bwd_data(dY: Tensor, lddy: int, W: Tensor, M: int, Nout: int, K: int, *, dy_col: int = 0, out: Tensor = None,
                    out_col: int = 0, lddx: int = None, w_is_kn: bool = False, ldw: int = None, w_col: int = 0,
                    gelu_pre: Tensor = None, rowscale: Tensor = None, rps: int = 0, accumulate: bool = False, view=None):
    if out is None:
        out = _empty((M, K), dY)
        lddx = K
    if ldw is None:
        ldw = Nout if w_is_kn else K
    if view is not None:
        N.unpatch_view(*view)          # dY lives as the fine channels-last grid (decoder tail)
    N.call("mic_linear_bwd_data", _view_ptr(dY, dy_col), lddy, _view_ptr(W, w_col), ldw, int(w_is_kn),
           _view_ptr(out, out_col), lddx, M, Nout, K, N.ptr(gelu_pre), K, N.ptr(rowscale), rps, int(accumulate))
    return out


def linear_bwd_weight(dY: Tensor, lddy: int, X: Tensor, ldx: int, M: int, Nout: int, K: int, *, dy_col: int = 0,
                      x_col: int = 0, w_is_kn: bool = False, dW: Tensor = None, lddw: int = None, dw_col: int = 0,
                      want_bias: bool = True, db: Tensor = None, rowscale: Tensor = None, rps: int = 0, view=None):
    """Returns (dW, db); dW/db are accumulated into when passed in.  ``view``: dY lives as the fine channels-last grid
    (decoder tail); db is then exact only after summing the 64 block positions of a channel (see the header)."""
    if dW is None:
        dW = _zeros((K, Nout) if w_is_kn else (Nout, K), dY)
        lddw = Nout if w_is_kn else K
    if db is None and want_bias:
        db = _zeros((Nout,), dY)
    if view is not None:
        N.unpatch_view(*view)
    N.call("mic_linear_bwd_weight", _view_ptr(dY, dy_col), lddy, _view_ptr(X, x_col), ldx, _view_ptr(dW, dw_col), lddw,
           int(w_is_kn), N.ptr(db) if want_bias else None, M, Nout, K, N.ptr(rowscale), rps)
    return dW, db


def linear_bwd_weight_side(sb: "side_branch", dY: Tensor, lddy: int, X: Tensor, ldx: int, M: int, Nout: int, K: int,
                           wp=None, bp=None, **kw):
    """linear_bwd_weight on the side branch; outputs are allocated (zeroed) on the calling stream first.
    ``wp`` / ``bp``: the weight / bias Parameters -- with a gradient arena attached the kernels accumulate straight into
    their ``.grad`` (row-major (Nout, K) weight, or a row slice of it selected by ``dw_row0``) and None is returned."""
    w_is_kn = kw.get("w_is_kn", False)
    row0 = kw.pop("dw_row0", 0)
    gw, gb = _acc(wp), _acc(bp)
    if gw is not None and kw.get("dW") is None and not w_is_kn:
        kw["dW"] = gw.view(-1, K)[row0:row0 + Nout]
        kw["lddw"] = K
    if gb is not None and kw.get("db") is None and kw.get("want_bias", True):
        kw["db"] = gb[row0:row0 + Nout]
    if kw.get("dW") is None:
        kw["dW"] = _zeros((K, Nout) if w_is_kn else (Nout, K), dY)
        kw["lddw"] = Nout if w_is_kn else K
    if kw.get("db") is None and kw.get("want_bias", True):
        kw["db"] = _zeros((Nout,), dY)
    sb.hold(dY, X, kw.get("rowscale"))
    dW, db = sb.run(linear_bwd_weight, dY, lddy, X, ldx, M, Nout, K, **kw)
    return (None if gw is not None and not w_is_kn else dW), (None if gb is not None else db)


def window_attn_fwd(qkv: Tensor, C: int, heads: int, B: int, pdims, ws):
    """qkv (P, 3C) rows on the padded grid -> o (P, C), lse (P, heads)."""
    Dp, Hp, Wp = pdims
    P = B * Dp * Hp * Wp
    hd = C // heads
    o = _empty((P, C), qkv)
    lse = _empty((P, heads), qkv)
    N.call("mic_window_attn_fwd", _view_ptr(qkv, 0), 3 * C, _view_ptr(qkv, C), _view_ptr(qkv, 2 * C), 3 * C, N.ptr(o), C,
           N.ptr(lse), B, Dp, Hp, Wp, heads, hd, ws[0], ws[1], ws[2], float(hd) ** -0.5)
    return o, lse


def window_attn_bwd(qkv: Tensor, o: Tensor, do: Tensor, lse: Tensor, C: int, heads: int, B: int, pdims, ws):
    Dp, Hp, Wp = pdims
    hd = C // heads
    dqkv = torch.empty_like(qkv)
    N.call("mic_window_attn_bwd", _view_ptr(qkv, 0), 3 * C, _view_ptr(qkv, C), _view_ptr(qkv, 2 * C), 3 * C, N.ptr(o),
           N.ptr(do), C, N.ptr(lse), _view_ptr(dqkv, 0), 3 * C, _view_ptr(dqkv, C), _view_ptr(dqkv, 2 * C), 3 * C, B, Dp, Hp,
           Wp, heads, hd, ws[0], ws[1], ws[2], float(hd) ** -0.5)
    return dqkv


def conv3_fwd(x0: Tensor, x1: Optional[Tensor], wt: Tensor, wk: Optional[Tensor], bias: Tensor, out: Tensor, B, dims,
              Co: int, ncdhw: bool) -> None:
    """3x3x3 conv forward on a (B, D, H, W) grid: tcgen05 implicit GEMM when the tensor-core mode is on and the
    geometry is taken (wk = weight as [27][Co][Cin]), else the fp32 CUDA-core kernel (wt = [27][Cin][Co])."""
    D, H, W = dims
    C0 = x0.shape[-1]
    C1 = x1.shape[-1] if x1 is not None else 0
    if wk is not None and N.get_gemm_mode() == 1:
        if N.try_call("mic_conv3_tc_fwd", N.ptr(x0), C0, N.ptr(x1), C1, N.ptr(wk), N.ptr(bias), N.ptr(out), B, D, H, W, Co,
                      int(ncdhw)):
            return
    N.call("mic_conv3_fwd", N.ptr(x0), C0, N.ptr(x1), C1, N.ptr(wt), N.ptr(bias), N.ptr(out), B, D, H, W, D, H, W, Co,
           int(ncdhw))


def conv3_bwd_data(dy: Tensor, wt: Tensor, dx0: Tensor, acc0: bool, dx1: Optional[Tensor], acc1: bool, B, dims, Co: int,
                   dy_ncdhw: bool) -> None:
    """dx0 | dx1 (+)= transposed 3x3x3 conv of dy with wt = [27][Cin][Co]: tcgen05 implicit GEMM (the forward kernel
    with mirrored taps) when the tensor-core mode is on and the geometry is taken, else the fp32 CUDA-core kernel."""
    D, H, W = dims
    C0 = dx0.shape[-1]
    C1 = dx1.shape[-1] if dx1 is not None else 0
    if N.get_gemm_mode() == 1:
        if N.try_call("mic_conv3_tc_bwd_data", N.ptr(dy), N.ptr(wt), N.ptr(dx0), C0, int(acc0), N.ptr(dx1), C1, int(acc1), B, D,
                      H, W, Co, int(dy_ncdhw)):
            return
    N.call("mic_conv3_bwd_data", N.ptr(dy), N.ptr(wt), N.ptr(dx0), C0, int(acc0), N.ptr(dx1), C1, int(acc1), B, D, H, W, D, H,
           W, Co, int(dy_ncdhw))


def conv3_bwd_weight(dy: Tensor, x0: Tensor, x1: Optional[Tensor], dwt: Tensor, dbias: Optional[Tensor], B, dims, Co: int,
                     dy_ncdhw: bool, native: bool = False) -> None:
    """dwt += , dbias[Co] += : TF32 mma.sync kernel in tensor-core mode, else the fp32 CUDA-core kernel.
    ``native``: dwt is the Conv3d parameter's own (Co, Cin, 3, 3, 3) gradient; otherwise the permuted [27][Cin][Co] buffer."""
    D, H, W = dims
    C0 = x0.shape[-1]
    C1 = x1.shape[-1] if x1 is not None else 0
    if N.get_gemm_mode() == 1:
        if N.try_call("mic_conv3_mma_bwd_weight", N.ptr(dy), N.ptr(x0), C0, N.ptr(x1), C1, N.ptr(dwt), N.ptr(dbias), B, D, H, W,
                      Co, int(dy_ncdhw), int(native)):
            return
    if native:       # the CUDA-core kernel writes [27][Cin][Co]: permute-add (exact-fp32 mode / shapes the mma kernel declines)
        tmp = torch.zeros(27, C0 + C1, Co, device=dy.device, dtype=torch.float32)
        N.call("mic_conv3_bwd_weight", N.ptr(dy), N.ptr(x0), C0, N.ptr(x1), C1, N.ptr(tmp), N.ptr(dbias), B, D, H, W, D, H, W, Co,
               int(dy_ncdhw))
        dwt.view(Co, C0 + C1, 27).add_(tmp.permute(2, 1, 0))
        return
    N.call("mic_conv3_bwd_weight", N.ptr(dy), N.ptr(x0), C0, N.ptr(x1), C1, N.ptr(dwt), N.ptr(dbias), B, D, H, W, D, H, W, Co,
           int(dy_ncdhw))


def pad_grid(x: Tensor, dims, pdims) -> Tensor:
    """zero-pad (B,D,H,W,C) -> (B,Dp,Hp,Wp,C)  (F.pad of xa, reference M:350)"""
    B, D, H, W = dims
    out = _empty((B, *pdims, x.shape[-1]), x)
    N.call("mic_crop_residual_bwd", N.ptr(x), None, N.ptr(out), B, D, H, W, *pdims, x.shape[-1])
    return out


def crop_add(res: Tensor, branch_p: Tensor, rowscale: Optional[Tensor], dims, pdims) -> Tensor:
    B, D, H, W = dims
    y = torch.empty_like(res)
    N.call("mic_crop_residual", N.ptr(res), N.ptr(branch_p), N.ptr(rowscale), N.ptr(y), B, D, H, W, *pdims, res.shape[-1])
    return y


def crop_bwd(dy: Tensor, rowscale: Optional[Tensor], dims, pdims) -> Tensor:
    B, D, H, W = dims
    out = _empty((B, *pdims, dy.shape[-1]), dy)
    N.call("mic_crop_residual_bwd", N.ptr(dy), N.ptr(rowscale), N.ptr(out), B, D, H, W, *pdims, dy.shape[-1])
    return out


# ----------------------------------------------------------------------------------------------------------
# shared pieces of the two transformer blocks
# ----------------------------------------------------------------------------------------------------------
def _proj_residual_fwd(x, o_p, pw, pb, s1, dims, pdims, padded):
    B, D, H, W = dims
    C = x.shape[-1]
    T = B * D * H * W
    if not padded:
        return linear_fwd(o_p, C, pw, pb, T, C, C, res=x, rowscale=s1, rps=D * H * W).view(x.shape)
    P = B * pdims[0] * pdims[1] * pdims[2]
    pr = linear_fwd(o_p, C, pw, pb, P, C, C)
    return crop_add(x, pr, s1, dims, pdims)


def _proj_residual_bwd(sb, dx1, o_p, pw, s1, dims, pdims, padded, pb=None):
    """-> do_p (P,C), dpw, dpb"""
    B, D, H, W = dims
    C = dx1.shape[-1]
    if not padded:
        T = B * D * H * W
        dpw, dpb = linear_bwd_weight_side(sb, dx1, C, o_p, C, T, C, C, wp=pw, bp=pb, rowscale=s1, rps=D * H * W)
        do_p = linear_bwd_data(dx1, C, pw, T, C, C, rowscale=s1, rps=D * H * W)
        return do_p, dpw, dpb
    P = B * pdims[0] * pdims[1] * pdims[2]
    dpr = crop_bwd(dx1, s1, dims, pdims)
    dpw, dpb = linear_bwd_weight_side(sb, dpr, C, o_p, C, P, C, C, wp=pw, bp=pb)
    do_p = linear_bwd_data(dpr, C, pw, P, C, C)
    return do_p, dpw, dpb


def _mlp_fwd(x1, n2w, n2b, f1w, f1b, f2w, f2b, s2, dims, img=None):
    """``img``: the Mlp's fused-kernel weight images -> LN + fc1 + GELU + fc2 + residual in ONE tcgen05 kernel that saves
    nothing (the backward recomputes from x1); None -> the unfused kernels with their saved intermediates."""
    B, D, H, W = dims
    C = x1.shape[-1]
    T = B * D * H * W
    Hd = f1w.shape[0]
    if img is not None:
        y = F_.mlp_block_fwd(x1, img, n2w, n2b, f1b, f2b, s2, D * H * W, LN_EPS)
        return y, (None, None, None, None, None)
    xn2, mean2, rstd2 = ln_fwd(x1, None, n2w, n2b, dims)
    hpre = _empty((T, Hd), x1)
    h = linear_fwd(xn2, C, f1w, f1b, T, Hd, C, act=True, pre=hpre)
    y = linear_fwd(h, Hd, f2w, f2b, T, C, Hd, res=x1, rowscale=s2, rps=D * H * W).view(x1.shape)
    return y, (xn2, mean2, rstd2, hpre, h)


def _mlp_bwd(sb, dy, x1, saved, n2w, f1w, f2w, s2, dims, n2b=None, f1b=None, f2b=None, img=None):
    """-> dx1 (= dy + grad through LN/MLP), dn2w, dn2b, df1w, df1b, df2w, df2b"""
    xn2, mean2, rstd2, hpre, h = saved
    B, D, H, W = dims
    C = x1.shape[-1]
    T = B * D * H * W
    Hd = f1w.shape[0]
    rps = D * H * W
    if img is not None and C in F_.FUSED_C:
        ps = (n2w, n2b, f1w, f1b, f2w, f2b)
        gs = [_acc(p) if _acc(p) is not None else _zeros(tuple(p.shape), x1) for p in ps]
        dx1 = F_.mlp_block_bwd(dy, x1, img, n2w, n2b, f1b, s2, rps, LN_EPS, *gs)
        return (dx1, *[_gret(p, g) for p, g in zip(ps, gs)])
    if img is not None:
        # deep stages: hidden-split kernel (fc1 / fc2 gradients + d LN(x)), then the LayerNorm backward kernel
        ps = (f1w, f1b, f2w, f2b)
        gs = [_acc(p) if _acc(p) is not None else _zeros(tuple(p.shape), x1) for p in ps]
        dxn, mean2, rstd2 = F_.mlp_split_bwd(dy, x1, img, n2w, n2b, f1b, s2, rps, LN_EPS, *gs)
        dx1, _, dn2w, dn2b = ln_bwd(dxn, x1, None, n2w, mean2, rstd2, dy, None, dims, beta=n2b)
        return (dx1, dn2w, dn2b, *[_gret(p, g) for p, g in zip(ps, gs)])
    df2w, df2b = linear_bwd_weight_side(sb, dy, C, h, Hd, T, C, Hd, wp=f2w, bp=f2b, rowscale=s2, rps=rps)
    dh = linear_bwd_data(dy, C, f2w, T, C, Hd, gelu_pre=hpre, rowscale=s2, rps=rps)
    df1w, df1b = linear_bwd_weight_side(sb, dh, Hd, xn2, C, T, Hd, C, wp=f1w, bp=f1b)
    dxn2 = linear_bwd_data(dh, Hd, f1w, T, Hd, C)
    dx1, _, dn2w, dn2b = ln_bwd(dxn2, x1, None, n2w, mean2, rstd2, dy, None, dims, beta=n2b)
    return dx1, dn2w, dn2b, df1w, df1b, df2w, df2b


class SelfBlockFn(torch.autograd.Function):
    """TransformerBlock3D.forward (reference M:473-524) as one tape node.

    args: x, s1, s2 (per-sample DropPath scales or None), heads, window, then 14 parameters:
    norm1.{w,b}, q.{w,b}, kv.{w,b}, proj.{w,b}, norm2.{w,b}, fc1.{w,b}, fc2.{w,b}."""

    @staticmethod
    def forward(ctx, x, s1, s2, heads, window, n1w, n1b, qw, qb, kvw, kvb, pw, pb, n2w, n2b, f1w, f1b, f2w, f2b, mlp_img=None,
                qkv_cat=None):
        N.check_cuda_f32(x, n1w, qw, kvw, pw, f1w, f2w)
        B, D, H, W, C = x.shape
        dims = (B, D, H, W)
        ws, pdims = window_geometry((D, H, W), window)
        padded = pdims != (D, H, W)
        P = B * pdims[0] * pdims[1] * pdims[2]
        xn_p, mean1, rstd1 = ln_fwd(x, None, n1w, n1b, dims, pdims)
        qkv = _empty((P, 3 * C), x)
        if qkv_cat is not None and (qkv_cat.b is not None) == (qb is not None):
            # q and kv read the same LayerNorm output: one GEMM against the concatenated (3C, C) weight (fused.QkvCat)
            linear_fwd(xn_p, C, qkv_cat.w, qkv_cat.b, P, 3 * C, C, out=qkv, out_col=0, ldy=3 * C)
            ctx.wqkv = qkv_cat.w
        else:
            linear_fwd(xn_p, C, qw, qb, P, C, C, out=qkv, out_col=0, ldy=3 * C)
            linear_fwd(xn_p, C, kvw, kvb, P, 2 * C, C, out=qkv, out_col=C, ldy=3 * C)
            ctx.wqkv = None
        o_p, lse = window_attn_fwd(qkv, C, heads, B, pdims, ws)
        x1 = _proj_residual_fwd(x, o_p, pw, pb, s1, dims, pdims, padded)
        y, mlp_saved = _mlp_fwd(x1, n2w, n2b, f1w, f1b, f2w, f2b, s2, dims, mlp_img)
        ctx.save_for_backward(x, xn_p, mean1, rstd1, qkv, o_p, lse, x1, *mlp_saved, n1w, qw, kvw, pw, n2w, f1w, f2w,
                              *( [s1] if s1 is not None else []), *([s2] if s2 is not None else []))
        ctx.meta = (dims, pdims, ws, padded, heads, s1 is not None, s2 is not None)
        ctx.biases = (n1b, qb, kvb, pb, n2b, f1b, f2b)
        ctx.mlp_img = mlp_img
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        dims, pdims, ws, padded, heads, has1, has2 = ctx.meta
        n1b, qb, kvb, pb, n2b, f1b, f2b = ctx.biases
        sv = list(ctx.saved_tensors)
        x, xn_p, mean1, rstd1, qkv, o_p, lse, x1 = sv[:8]
        mlp_saved = sv[8:13]
        n1w, qw, kvw, pw, n2w, f1w, f2w = sv[13:20]
        rest = sv[20:]
        s1 = rest.pop(0) if has1 else None
        s2 = rest.pop(0) if has2 else None
        B, D, H, W = dims
        C = x.shape[-1]
        P = B * pdims[0] * pdims[1] * pdims[2]
        dy = dy.contiguous()
        with zero_arena(12 * C * C + 128 * C + 4096, x), side_branch(defer=True) as sb:
            dx1, dn2w, dn2b, df1w, df1b, df2w, df2b = _mlp_bwd(sb, dy, x1, mlp_saved, n2w, f1w, f2w, s2, dims, n2b, f1b, f2b,
                                                               ctx.mlp_img)
            do_p, dpw, dpb = _proj_residual_bwd(sb, dx1, o_p, pw, s1, dims, pdims, padded, pb)
            dqkv = window_attn_bwd(qkv, o_p, do_p, lse, C, heads, B, pdims, ws)
            dqw, dqb = linear_bwd_weight_side(sb, dqkv, 3 * C, xn_p, C, P, C, C, wp=qw, bp=qb)
            dkvw, dkvb = linear_bwd_weight_side(sb, dqkv, 3 * C, xn_p, C, P, 2 * C, C, wp=kvw, bp=kvb, dy_col=C)
            if ctx.wqkv is not None:
                dxn_p = linear_bwd_data(dqkv, 3 * C, ctx.wqkv, P, 3 * C, C)       # reduction over q | k | v at once
            else:
                dxn_p = linear_bwd_data(dqkv, 3 * C, qw, P, C, C)
                linear_bwd_data(dqkv, 3 * C, kvw, P, 2 * C, C, dy_col=C, out=dxn_p, lddx=C, accumulate=True)
            dx, _, dn1w, dn1b = ln_bwd(dxn_p, x, None, n1w, mean1, rstd1, dx1, None, dims, pdims, beta=n1b)
        return (dx, None, None, None, None, dn1w, dn1b, dqw, dqb, dkvw, dkvb, dpw, dpb, dn2w, dn2b, df1w, df1b, df2w,
                df2b, None, None)


class CrossBlockFn(torch.autograd.Function):
    """CrossTransformerBlock3D.forward (reference M:339-426): LN(x) -> offset net on cat[LN(x), xa] -> deformable
    trilinear resampling of xa -> windowed cross attention (q from LN(x), k/v from the resampled xa) -> proj +
    residual -> LN -> MLP -> residual.

    args: x, xa, s1, s2, heads, window, then the parameters: norm1.{w,b}, q.{w,b}, kv.{w,b}, proj.{w,b},
    conv_offset.0 weight permuted to (27, 2C, 16) [and optionally to (27, 16, 2C) for the tcgen05 forward; None
    otherwise] and bias, conv_offset.1.norm.{w,b}, conv_offset.3 weight (3,16),
    norm2.{w,b}, fc1.{w,b}, fc2.{w,b}."""

    @staticmethod
    def forward(ctx, x, xa, s1, s2, heads, window, n1w, n1b, qw, qb, kvw, kvb, pw, pb, cw, cwk, cb, lnw, lnb, w3, n2w,
                n2b, f1w, f1b, f2w, f2b, mlp_img=None, cwp=None):
        N.check_cuda_f32(x, xa, n1w, qw, kvw, pw, cw, w3, f1w, f2w)
        B, D, H, W, C = x.shape
        dims = (B, D, H, W)
        ws, pdims = window_geometry((D, H, W), window)
        padded = pdims != (D, H, W)
        Dp, Hp, Wp = pdims
        P = B * Dp * Hp * Wp
        HC = cw.shape[-1]
        xn_p, mean1, rstd1 = ln_fwd(x, None, n1w, n1b, dims, pdims)
        xa_p = pad_grid(xa, dims, pdims) if padded else xa
        h16 = _empty((P, HC), x)
        qkv = _empty((P, 3 * C), x)
        with side_branch() as sb:
            # q depends only on LN(x): its GEMM runs beside the offset branch (conv -> head -> resampling) instead of after it
            sb.hold(xn_p, qkv)
            (sb.run if _Q_SIDE else _inline)(linear_fwd, xn_p, C, qw, qb, P, C, C, out=qkv, out_col=0, ldy=3 * C)
            conv3_fwd(xn_p, xa_p, cw, cwk, cb, h16, B, (Dp, Hp, Wp), HC, False)
            pos = _empty((P, 3), x)
            N.call("mic_offset_head_fwd", N.ptr(h16), N.ptr(lnw), N.ptr(lnb), N.ptr(w3), N.ptr(pos), B, Dp, Hp, Wp, HC, LN_EPS)
            samp = _empty((P, C), x)
            N.call("mic_deform_sample_fwd", N.ptr(xa_p), N.ptr(pos), N.ptr(samp), B, Dp, Hp, Wp, Dp, Hp, Wp, C)
            linear_fwd(samp, C, kvw, kvb, P, 2 * C, C, out=qkv, out_col=C, ldy=3 * C)
        o_p, lse = window_attn_fwd(qkv, C, heads, B, pdims, ws)
        x1 = _proj_residual_fwd(x, o_p, pw, pb, s1, dims, pdims, padded)
        y, mlp_saved = _mlp_fwd(x1, n2w, n2b, f1w, f1b, f2w, f2b, s2, dims, mlp_img)
        ctx.save_for_backward(x, xa_p, xn_p, mean1, rstd1, h16, pos, samp, qkv, o_p, lse, x1, *mlp_saved, n1w, qw, kvw,
                              pw, cw, lnw, lnb, w3, n2w, f1w, f2w, *([s1] if s1 is not None else []),
                              *([s2] if s2 is not None else []))
        ctx.meta = (dims, pdims, ws, padded, heads, s1 is not None, s2 is not None)
        ctx.biases = (n1b, qb, kvb, pb, cb, n2b, f1b, f2b)
        ctx.mlp_img = mlp_img
        ctx.cwp = cwp       # conv_offset.0.weight itself: its gradient is written in the parameter's layout (cw, cwk: layout copies)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        dims, pdims, ws, padded, heads, has1, has2 = ctx.meta
        n1b, qb, kvb, pb, cb, n2b, f1b, f2b = ctx.biases
        sv = list(ctx.saved_tensors)
        x, xa_p, xn_p, mean1, rstd1, h16, pos, samp, qkv, o_p, lse, x1 = sv[:12]
        mlp_saved = sv[12:17]
        n1w, qw, kvw, pw, cw, lnw, lnb, w3, n2w, f1w, f2w = sv[17:28]
        rest = sv[28:]
        s1 = rest.pop(0) if has1 else None
        s2 = rest.pop(0) if has2 else None
        B, D, H, W = dims
        Dp, Hp, Wp = pdims
        C = x.shape[-1]
        P = B * Dp * Hp * Wp
        HC = cw.shape[-1]
        dy = dy.contiguous()
        with zero_arena(12 * C * C + 27 * 2 * C * HC + 128 * C + 8192 + P * C, x), side_branch(defer=True) as sb:
            dx1, dn2w, dn2b, df1w, df1b, df2w, df2b = _mlp_bwd(sb, dy, x1, mlp_saved, n2w, f1w, f2w, s2, dims, n2b, f1b, f2b,
                                                               ctx.mlp_img)
            do_p, dpw, dpb = _proj_residual_bwd(sb, dx1, o_p, pw, s1, dims, pdims, padded, pb)
            dqkv = window_attn_bwd(qkv, o_p, do_p, lse, C, heads, B, pdims, ws)
            # the q path's data gradient only meets the offset branch again at the conv's backward-data (which accumulates
            # into dxn_p): it runs on the side branch, ahead of the q / kv weight gradients
            dxn_p = _empty((P, C), x)
            sb.hold(dqkv, dxn_p)
            (sb.run if _Q_SIDE else _inline)(linear_bwd_data, dqkv, 3 * C, qw, P, C, C, out=dxn_p, lddx=C)
            dxn_ready = sb.mark() if _Q_SIDE else None
            dqw, dqb = linear_bwd_weight_side(sb, dqkv, 3 * C, xn_p, C, P, C, C, wp=qw, bp=qb)
            dkvw, dkvb = linear_bwd_weight_side(sb, dqkv, 3 * C, samp, C, P, 2 * C, C, wp=kvw, bp=kvb, dy_col=C)
            dsamp = linear_bwd_data(dqkv, 3 * C, kvw, P, 2 * C, C, dy_col=C)
            dxa_p = _zeros((B, Dp, Hp, Wp, C), x)
            dpos = _empty((P, 3), x)
            N.call("mic_deform_sample_bwd", N.ptr(dsamp), N.ptr(xa_p), N.ptr(pos), N.ptr(dxa_p), N.ptr(dpos), B, Dp, Hp, Wp, Dp,
                   Hp, Wp, C)
            dh16 = _empty((P, HC), x)
            dlnw = _acc(lnw) if _acc(lnw) is not None else _zeros((HC,), x)
            dlnb = _acc(lnb) if _acc(lnb) is not None else _zeros((HC,), x)
            dw3 = _zeros((3, HC), x)
            N.call("mic_offset_head_bwd", N.ptr(dpos), N.ptr(h16), N.ptr(lnw), N.ptr(lnb), N.ptr(w3), N.ptr(dh16), N.ptr(dlnw),
                   N.ptr(dlnb), N.ptr(dw3), B, Dp, Hp, Wp, HC, LN_EPS)
            dlnw, dlnb = _gret(lnw, dlnw), _gret(lnb, dlnb)
            cwp = ctx.cwp
            dcw = _zeros(tuple(cw.shape), x) if cwp is None else (_acc(cwp) if _acc(cwp) is not None else _zeros(tuple(cwp.shape), x))
            dcb = _acc(cb) if _acc(cb) is not None else _zeros((HC,), x)
            sb.hold(dh16, xn_p, xa_p)
            sb.run(conv3_bwd_weight, dh16, xn_p, xa_p, dcw, dcb, B, (Dp, Hp, Wp), HC, False, cwp is not None)
            sb.wait(dxn_ready)                 # dq Wq is in dxn_p
            conv3_bwd_data(dh16, cw, dxn_p.view(B, Dp, Hp, Wp, C), True, dxa_p, True, B, (Dp, Hp, Wp), HC, False)
            dx, _, dn1w, dn1b = ln_bwd(dxn_p, x, None, n1w, mean1, rstd1, dx1, None, dims, pdims, beta=n1b)
        if padded:
            dxa = dxa_p[:, :D, :H, :W, :].contiguous()
        else:
            dxa = dxa_p
        dcw_old, dcw_new = (dcw, None) if cwp is None else (None, _gret(cwp, dcw))
        return (dx, dxa, None, None, None, None, dn1w, dn1b, dqw, dqb, dkvw, dkvb, dpw, dpb, dcw_old, None, _gret(cb, dcb), dlnw, dlnb,
                dw3, dn2w, dn2b, df1w, df1b, df2w, df2b, None, dcw_new)


# ----------------------------------------------------------------------------------------------------------
# fully fused blocks (2x2x2 windows, C = 48: the train config's stage 0): two tcgen05 kernels per direction
# ----------------------------------------------------------------------------------------------------------
_ATTN_P = ("n1w", "n1b", "qw", "qb", "kvw", "kvb", "pw", "pb")


def _grad_bufs(ps, like):
    gs = [_acc(p) if _acc(p) is not None else _zeros(tuple(p.shape), like) for p in ps]
    return gs, [_gret(p, g) for p, g in zip(ps, gs)]


class FusedSelfBlockFn(torch.autograd.Function):
    """TransformerBlock3D.forward (reference M:473-524) as two fused kernels: attention half (LN1, q|kv, 8-token window
    attention, proj, residual) and MLP half (LN2, fc1, GELU, fc2, residual).  Saves only x and x1; the backward kernels
    recompute everything else on chip and accumulate all 14 parameter gradients.

    args: x, s1, s2, heads, attention images, MLP images, then norm1.{w,b}, q.{w,b}, kv.{w,b}, proj.{w,b}, norm2.{w,b},
    fc1.{w,b}, fc2.{w,b}."""

    @staticmethod
    def forward(ctx, x, s1, s2, heads, aimg, mimg, n1w, n1b, qw, qb, kvw, kvb, pw, pb, n2w, n2b, f1w, f1b, f2w, f2b):
        N.check_cuda_f32(x, n1w, qw, kvw, pw, f1w, f2w)
        B, D, H, W, C = x.shape
        x1 = F_.attn_block_fwd(x, None, aimg, n1w, n1b, qb, kvb, pb, s1, heads, LN_EPS)
        y = F_.mlp_block_fwd(x1, mimg, n2w, n2b, f1b, f2b, s2, D * H * W, LN_EPS)
        ctx.save_for_backward(x, x1, *([s1] if s1 is not None else []), *([s2] if s2 is not None else []))
        ctx.meta = (heads, s1 is not None, s2 is not None)
        ctx.params = (n1w, n1b, qw, qb, kvw, kvb, pw, pb, n2w, n2b, f1w, f1b, f2w, f2b)
        ctx.imgs = (aimg, mimg)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        heads, has1, has2 = ctx.meta
        sv = list(ctx.saved_tensors)
        x, x1 = sv[:2]
        rest = sv[2:]
        s1 = rest.pop(0) if has1 else None
        s2 = rest.pop(0) if has2 else None
        n1w, n1b, qw, qb, kvw, kvb, pw, pb, n2w, n2b, f1w, f1b, f2w, f2b = ctx.params
        aimg, mimg = ctx.imgs
        B, D, H, W, C = x.shape
        dy = dy.contiguous()
        with zero_arena(12 * C * C + 64 * C + 4096, x):
            mg, mret = _grad_bufs((n2w, n2b, f1w, f1b, f2w, f2b), x)
            dx1 = F_.mlp_block_bwd(dy, x1, mimg, n2w, n2b, f1b, s2, D * H * W, LN_EPS, *mg)
            ag, aret = _grad_bufs((n1w, n1b, qw, qb, kvw, kvb, pw, pb), x)
            dx, _ = F_.attn_block_bwd(dx1, x, None, aimg, n1w, n1b, qb, kvb, s1, heads, LN_EPS, *ag)
        return (dx, None, None, None, None, None, *aret, *mret)


class FusedCrossBlockFn(torch.autograd.Function):
    """CrossTransformerBlock3D.forward (reference M:339-426) with the attention and MLP halves as fused kernels.  The
    offset branch (LN(x) -> conv_offset on cat[LN(x), xa] -> offset head -> deformable resampling of xa) keeps its own
    kernels and produces the k/v source ``samp`` of the fused attention kernel.

    args: x, xa, s1, s2, heads, attention images, MLP images, then the parameters in CrossBlockFn's order."""

    @staticmethod
    def forward(ctx, x, xa, s1, s2, heads, aimg, mimg, n1w, n1b, qw, qb, kvw, kvb, pw, pb, cw, cwk, cb, lnw, lnb, w3, n2w,
                n2b, f1w, f1b, f2w, f2b, cwp=None):
        N.check_cuda_f32(x, xa, n1w, qw, kvw, pw, cw, w3, f1w, f2w)
        ctx.cwp = cwp
        B, D, H, W, C = x.shape
        dims = (B, D, H, W)
        P = B * D * H * W
        HC = cw.shape[-1]
        xn, mean1, rstd1 = ln_fwd(x, None, n1w, n1b, dims)
        h16 = _empty((P, HC), x)
        conv3_fwd(xn, xa, cw, cwk, cb, h16, B, (D, H, W), HC, False)
        pos = _empty((P, 3), x)
        N.call("mic_offset_head_fwd", N.ptr(h16), N.ptr(lnw), N.ptr(lnb), N.ptr(w3), N.ptr(pos), B, D, H, W, HC, LN_EPS)
        samp = torch.empty_like(x)
        N.call("mic_deform_sample_fwd", N.ptr(xa), N.ptr(pos), N.ptr(samp), B, D, H, W, D, H, W, C)
        x1 = F_.attn_block_fwd(x, samp, aimg, n1w, n1b, qb, kvb, pb, s1, heads, LN_EPS)
        y = F_.mlp_block_fwd(x1, mimg, n2w, n2b, f1b, f2b, s2, D * H * W, LN_EPS)
        ctx.save_for_backward(x, xa, xn, mean1, rstd1, h16, pos, samp, x1, *([s1] if s1 is not None else []),
                              *([s2] if s2 is not None else []))
        ctx.meta = (heads, s1 is not None, s2 is not None)
        ctx.params = (n1w, n1b, qw, qb, kvw, kvb, pw, pb, cw, cb, lnw, lnb, w3, n2w, n2b, f1w, f1b, f2w, f2b)
        ctx.imgs = (aimg, mimg)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        heads, has1, has2 = ctx.meta
        sv = list(ctx.saved_tensors)
        x, xa, xn, mean1, rstd1, h16, pos, samp, x1 = sv[:9]
        rest = sv[9:]
        s1 = rest.pop(0) if has1 else None
        s2 = rest.pop(0) if has2 else None
        n1w, n1b, qw, qb, kvw, kvb, pw, pb, cw, cb, lnw, lnb, w3, n2w, n2b, f1w, f1b, f2w, f2b = ctx.params
        aimg, mimg = ctx.imgs
        B, D, H, W, C = x.shape
        dims = (B, D, H, W)
        P = B * D * H * W
        HC = cw.shape[-1]
        dy = dy.contiguous()
        with zero_arena(12 * C * C + 27 * 2 * C * HC + 64 * C + 8192 + P * C, x), side_branch(defer=True) as sb:
            mg, mret = _grad_bufs((n2w, n2b, f1w, f1b, f2w, f2b), x)
            dx1 = F_.mlp_block_bwd(dy, x1, mimg, n2w, n2b, f1b, s2, D * H * W, LN_EPS, *mg)
            # attention half: dxq = dx1 + LN1'(dq Wq) (the q path of norm1), dsamp = [dk|dv] Wkv; norm1's gradients through the
            # q path are accumulated here, those through the offset branch by ln_bwd below (LayerNorm' is linear in its input)
            ag, aret = _grad_bufs((n1w, n1b, qw, qb, kvw, kvb, pw, pb), x)
            dxq, dsamp = F_.attn_block_bwd(dx1, x, samp, aimg, n1w, n1b, qb, kvb, s1, heads, LN_EPS, *ag)
            dxa = _zeros((B, D, H, W, C), x)
            dpos = _empty((P, 3), x)
            N.call("mic_deform_sample_bwd", N.ptr(dsamp), N.ptr(xa), N.ptr(pos), N.ptr(dxa), N.ptr(dpos), B, D, H, W, D, H, W, C)
            dh16 = _empty((P, HC), x)
            dlnw = _acc(lnw) if _acc(lnw) is not None else _zeros((HC,), x)
            dlnb = _acc(lnb) if _acc(lnb) is not None else _zeros((HC,), x)
            dw3 = _zeros((3, HC), x)
            N.call("mic_offset_head_bwd", N.ptr(dpos), N.ptr(h16), N.ptr(lnw), N.ptr(lnb), N.ptr(w3), N.ptr(dh16), N.ptr(dlnw),
                   N.ptr(dlnb), N.ptr(dw3), B, D, H, W, HC, LN_EPS)
            cwp = ctx.cwp
            dcw = _zeros(tuple(cw.shape), x) if cwp is None else (_acc(cwp) if _acc(cwp) is not None else _zeros(tuple(cwp.shape), x))
            dcb = _acc(cb) if _acc(cb) is not None else _zeros((HC,), x)
            sb.hold(dh16, xn, xa)
            sb.run(conv3_bwd_weight, dh16, xn, xa, dcw, dcb, B, (D, H, W), HC, False, cwp is not None)
            dxn = _empty((B, D, H, W, C), x)
            conv3_bwd_data(dh16, cw, dxn, False, dxa, True, B, (D, H, W), HC, False)
            # second use of norm1's backward: the gradient that reached LN(x) through the offset conv
            dx, _, dn1w2, dn1b2 = ln_bwd(dxn, x, None, n1w, mean1, rstd1, dxq, None, dims, beta=n1b)
            if _acc(n1w) is None:                    # no arena: add the two contributions of norm1's parameters
                aret[0] = aret[0] + dn1w2
                aret[1] = aret[1] + dn1b2
        dcw_old, dcw_new = (dcw, None) if cwp is None else (None, _gret(cwp, dcw))
        return (dx, dxa, None, None, None, None, None, *aret, dcw_old, None, _gret(cb, dcb), _gret(lnw, dlnw), _gret(lnb, dlnb),
                dw3, *mret, dcw_new)


# ----------------------------------------------------------------------------------------------------------
# non-block operators
# ----------------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last axis of a token grid, optionally on the channel concatenation [x0 | x1]
    (``norm`` M:1011-1012; cat + ``norm2`` M:1033-1034)."""

    @staticmethod
    def forward(ctx, x0, x1, w, b):
        N.check_cuda_f32(x0, x1, w, b)
        B = x0.shape[0]
        rows = x0.numel() // x0.shape[-1] // B
        dims = (B, 1, 1, rows)
        y, mean, rstd = ln_fwd(x0, x1, w, b, dims)
        ctx.save_for_backward(x0, x1 if x1 is not None else x0, w, mean, rstd)
        ctx.meta = (dims, x1 is not None)
        ctx.beta = b
        C = x0.shape[-1] + (x1.shape[-1] if x1 is not None else 0)
        return y.view(*x0.shape[:-1], C)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x0, x1, w, mean, rstd = ctx.saved_tensors
        dims, has1 = ctx.meta
        dx0, dx1, dw, db = ln_bwd(dy.contiguous(), x0, x1 if has1 else None, w, mean, rstd, None, None, dims, beta=ctx.beta)
        return dx0, dx1, dw, db


class SkipLinearFn(torch.autograd.Function):
    """concat_back_dim: Linear(2C -> C) on cat[up, skip] (reference M:1027-1030) as two accumulating GEMMs."""

    @staticmethod
    def forward(ctx, a, b, w, bias):
        N.check_cuda_f32(a, b, w, bias)
        Ca, Cb = a.shape[-1], b.shape[-1]
        Cout = w.shape[0]
        T = a.numel() // Ca
        y = linear_fwd(a, Ca, w, bias, T, Cout, Ca, ldw=Ca + Cb)
        linear_fwd(b, Cb, w, None, T, Cout, Cb, ldw=Ca + Cb, w_col=Ca, out=y, ldy=Cout, accumulate=True)
        ctx.save_for_backward(a, b, w)
        ctx.bias = bias
        return y.view(*a.shape[:-1], Cout)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        a, b, w = ctx.saved_tensors
        Ca, Cb = a.shape[-1], b.shape[-1]
        Cout = w.shape[0]
        T = a.numel() // Ca
        dy = dy.contiguous()
        da = linear_bwd_data(dy, Cout, w, T, Cout, Ca, ldw=Ca + Cb).view(a.shape)
        db_ = linear_bwd_data(dy, Cout, w, T, Cout, Cb, ldw=Ca + Cb, w_col=Ca).view(b.shape)
        dW = _acc(w) if _acc(w) is not None else torch.zeros_like(w)
        _, dbias = linear_bwd_weight(dy, Cout, a, Ca, T, Cout, Ca, dW=dW, lddw=Ca + Cb, db=_acc(ctx.bias))
        linear_bwd_weight(dy, Cout, b, Cb, T, Cout, Cb, dW=dW, lddw=Ca + Cb, dw_col=Ca, want_bias=False)
        return da, db_, _gret(w, dW), _gret(ctx.bias, dbias)


def block_permute(src: Tensor, dst: Tensor, B, Dq, Hq, Wq, k, C, grid_batch_stride, to_rows, src_off=0):
    N.call("mic_block_permute", src.data_ptr() + 4 * src_off if to_rows else src.data_ptr(),
           dst.data_ptr() if to_rows else dst.data_ptr() + 4 * src_off, B, Dq, Hq, Wq, k, C, grid_batch_stride, int(to_rows))


class PatchEmbedFn(torch.autograd.Function):
    """PatchEmbed3D (reference M:860-878, norm off): Conv3d(1->E,k4,s4) of channel ``ch`` of the (B,2,D,H,W) input,
    emitted channels-last (M:1001-1002).  4^3 gather -> (rows, 64) -> GEMM.  No input gradient."""

    @staticmethod
    def forward(ctx, vol, ch, w, b):
        N.check_cuda_f32(vol, w, b)
        B, Cin, D, H, W = vol.shape
        if D % 4 or H % 4 or W % 4:
            raise RuntimeError("PatchEmbedFn takes sizes divisible by 4: MicFormer._trunk zero-pads the volume first (M:864-869)")
        Dq, Hq, Wq = D // 4, H // 4, W // 4
        E = w.shape[0]
        rows = _empty((B * Dq * Hq * Wq, 64), vol)
        block_permute(vol, rows, B, Dq, Hq, Wq, 4, 1, Cin * D * H * W, True, src_off=ch * D * H * W)
        y = linear_fwd(rows, 64, w, b, rows.shape[0], E, 64)
        ctx.save_for_backward(rows)
        ctx.E = E
        ctx.wshape = tuple(w.shape)
        return y.view(B, Dq, Hq, Wq, E)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (rows,) = ctx.saved_tensors
        E = ctx.E
        dy = dy.contiguous()
        dW, db = linear_bwd_weight(dy, E, rows, 64, rows.shape[0], E, 64)
        return None, None, dW.view(ctx.wshape), db


class PatchMergeFn(torch.autograd.Function):
    """PatchMerging (reference M:542-561): Conv3d(C->2C,k2,s2) + LayerNorm(2C).  w2 is the conv weight permuted
    to (2C, kz,ky,kx, C) and flattened to (2C, 8C)."""

    @staticmethod
    def forward(ctx, x, w2, b, nw, nb):
        N.check_cuda_f32(x, w2, b, nw, nb)
        B, D, H, W, C = x.shape
        if D % 2 or H % 2 or W % 2:
            raise RuntimeError("PatchMergeFn takes even sizes: PatchMerging.forward zero-pads odd ones first (M:551-555)")
        Dq, Hq, Wq = D // 2, H // 2, W // 2
        R = B * Dq * Hq * Wq
        Co = w2.shape[0]
        rows = _empty((R, 8 * C), x)
        block_permute(x, rows, B, Dq, Hq, Wq, 2, C, D * H * W * C, True)
        z = linear_fwd(rows, 8 * C, w2, b, R, Co, 8 * C)
        y, mean, rstd = ln_fwd(z, None, nw, nb, (B, Dq, Hq, Wq))
        ctx.save_for_backward(rows, z, mean, rstd, w2, nw)
        ctx.meta = (B, Dq, Hq, Wq, C, Co)
        ctx.nb = nb
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        rows, z, mean, rstd, w2, nw = ctx.saved_tensors
        B, Dq, Hq, Wq, C, Co = ctx.meta
        R = B * Dq * Hq * Wq
        dz, _, dnw, dnb = ln_bwd(dy.contiguous(), z, None, nw, mean, rstd, None, None, (B, Dq, Hq, Wq), beta=ctx.nb)
        with side_branch() as sb:
            dW, db = linear_bwd_weight_side(sb, dz, Co, rows, 8 * C, R, Co, 8 * C)
            drows = linear_bwd_data(dz, Co, w2, R, Co, 8 * C)
            dx = _empty((B, 2 * Dq, 2 * Hq, 2 * Wq, C), dz)
            block_permute(drows, dx, B, Dq, Hq, Wq, 2, C, 8 * Dq * Hq * Wq * C, False)
        return dx, dW, db, dnw, dnb


class PatchExpandFn(torch.autograd.Function):
    """PatchExpand (reference M:571-579): ConvTranspose3d(C->C/2,k2,s2) + LayerNorm(C/2).  wk is the weight permuted
    to (C, kz,ky,kx, C/2) and flattened to (C, 8*C/2); b8 is the bias tiled 8x (one per sub-voxel)."""

    @staticmethod
    def forward(ctx, x, wk, b8, nw, nb):
        N.check_cuda_f32(x, wk, b8, nw, nb)
        B, D, H, W, C = x.shape
        Co = nw.shape[0]
        T = B * D * H * W
        rows = linear_fwd(x, C, wk, b8, T, 8 * Co, C, w_is_kn=True)
        z = _empty((B, 2 * D, 2 * H, 2 * W, Co), x)
        block_permute(rows, z, B, D, H, W, 2, Co, 8 * D * H * W * Co, False)
        y, mean, rstd = ln_fwd(z, None, nw, nb, (B, 2 * D, 2 * H, 2 * W))
        ctx.save_for_backward(x, z, mean, rstd, wk, nw)
        ctx.meta = (B, D, H, W, C, Co)
        ctx.nb = nb
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, z, mean, rstd, wk, nw = ctx.saved_tensors
        B, D, H, W, C, Co = ctx.meta
        T = B * D * H * W
        dz, _, dnw, dnb = ln_bwd(dy.contiguous(), z, None, nw, mean, rstd, None, None, (B, 2 * D, 2 * H, 2 * W), beta=ctx.nb)
        drows = _empty((T, 8 * Co), dz)
        block_permute(dz, drows, B, D, H, W, 2, Co, 8 * D * H * W * Co, True)
        with side_branch() as sb:
            dW, db8 = linear_bwd_weight_side(sb, drows, 8 * Co, x, C, T, 8 * Co, C, w_is_kn=True)
            dx = linear_bwd_data(drows, 8 * Co, wk, T, 8 * Co, C, w_is_kn=True).view(x.shape)
        return dx, dW, db8, dnw, dnb


class UnpatchFn(torch.autograd.Function):
    """The tail of ``MicFormer.forward`` on its own (reference M:1033-1037): cat[moving, fixed] -> norm2 ->
    ConvTranspose3d(2E->E/2,k4,s4), emitted channels-last (B, 4D, 4H, 4W, E/2); any E.  ``Head`` uses ``SegHeadFn``
    (this plus out_conv) instead.  wr: (2E, 64*E/2) permuted (kz,ky,kx,co); br64: bias tiled 64x."""

    @staticmethod
    def forward(ctx, xm, xf, n2w, n2b, wr, br64):
        N.check_cuda_f32(xm, xf, n2w, n2b, wr, br64)
        B, D, H, W, E = xm.shape
        T = B * D * H * W
        Ch = wr.shape[1] // 64
        xn, mean, rstd = ln_fwd(xm, xf, n2w, n2b, (B, D, H, W))
        rows = linear_fwd(xn, 2 * E, wr, br64, T, 64 * Ch, 2 * E, w_is_kn=True)
        y = _empty((B, 4 * D, 4 * H, 4 * W, Ch), xm)
        block_permute(rows, y, B, D, H, W, 4, Ch, 64 * D * H * W * Ch, False)
        ctx.save_for_backward(xm, xf, n2w, mean, rstd, xn, wr)
        ctx.meta = (B, D, H, W, E, Ch)
        ctx.n2b = n2b
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xm, xf, n2w, mean, rstd, xn, wr = ctx.saved_tensors
        B, D, H, W, E, Ch = ctx.meta
        T = B * D * H * W
        dy = dy.contiguous()
        drows = _empty((T, 64 * Ch), xm)
        block_permute(dy, drows, B, D, H, W, 4, Ch, 64 * D * H * W * Ch, True)
        dxn = linear_bwd_data(drows, 64 * Ch, wr, T, 64 * Ch, 2 * E, w_is_kn=True)
        dwr, dbr64 = linear_bwd_weight(drows, 64 * Ch, xn, 2 * E, T, 64 * Ch, 2 * E, w_is_kn=True)
        dxm, dxf, dn2w, dn2b = ln_bwd(dxn, xm, xf, n2w, mean, rstd, None, None, (B, D, H, W), beta=ctx.n2b)
        return dxm, dxf, dn2w, dn2b, dwr, dbr64


_TAIL_VIEW = _os.environ.get("MICFORMER_TAIL_VIEW", "1") != "0"


def _tail_view(D: int, H: int, W: int, Ch: int, T: int):
    """(ch, dc, hc, wc) when the tail GEMMs can address the fine grid directly (mic_linear_unpatch_view), else None."""
    if _TAIL_VIEW and N.get_gemm_mode() == 1 and W == 32 and H % 4 == 0 and Ch % 8 == 0 and T % 128 == 0:
        return (Ch, D, H, W)
    return None


class SegHeadFn(torch.autograd.Function):
    """Decoder tail (reference M:1033-1037 + Head M:1053): cat[moving, fixed] -> norm2 -> ConvTranspose3d(2E->E/2,k4,s4)
    -> Conv3d(E/2->num_classes,k3,p1), NCDHW logits.  wr: (2E, 64*E/2) permuted (kz,ky,kx,co); br64: bias tiled 64x;
    wo: out_conv weight permuted to (27, E/2, NC)."""

    @staticmethod
    def forward(ctx, xm, xf, n2w, n2b, wr, br64, wo, wok, bo):
        N.check_cuda_f32(xm, xf, n2w, n2b, wr, br64, wo, bo)
        B, D, H, W, E = xm.shape
        T = B * D * H * W
        Ch = wo.shape[1]
        NC = wo.shape[2]
        xn, mean, rstd = ln_fwd(xm, xf, n2w, n2b, (B, D, H, W))
        _dbg = _os.environ.get("MICFORMER_DEBUG_EXACT_TAIL", "")
        _mode = N.get_gemm_mode()
        if "gemm" in _dbg:
            N.set_gemm_mode(0)
        y24 = _empty((B, 4 * D, 4 * H, 4 * W, Ch), xm)
        view = _tail_view(D, H, W, Ch, T)
        if view is not None:
            # the GEMM's TMA stores scatter the ConvTranspose rows straight into the fine grid: no rows buffer, no permute pass
            linear_fwd(xn, 2 * E, wr, br64, T, 64 * Ch, 2 * E, w_is_kn=True, out=y24, ldy=64 * Ch, view=view)
        else:
            rows = linear_fwd(xn, 2 * E, wr, br64, T, 64 * Ch, 2 * E, w_is_kn=True)
            block_permute(rows, y24, B, D, H, W, 4, Ch, 64 * D * H * W * Ch, False)
            del rows
        N.set_gemm_mode(_mode)
        logits = _empty((B, NC, 4 * D, 4 * H, 4 * W), xm)
        if "conv" in _dbg:
            N.set_gemm_mode(0)
        conv3_fwd(y24, None, wo, wok, bo, logits, B, (4 * D, 4 * H, 4 * W), NC, True)
        N.set_gemm_mode(_mode)
        ctx.save_for_backward(xm, xf, n2w, mean, rstd, xn, wr, y24, wo)
        ctx.meta = (B, D, H, W, E, Ch, NC)
        ctx.pb = (n2b, bo)
        return logits

    @staticmethod
    @once_differentiable
    def backward(ctx, dlog):
        xm, xf, n2w, mean, rstd, xn, wr, y24, wo = ctx.saved_tensors
        B, D, H, W, E, Ch, NC = ctx.meta
        T = B * D * H * W
        dlog = dlog.contiguous()
        n2b, bo = ctx.pb
        with side_branch() as sb:
            # the two weight-gradient kernels of the tail (1.1 + 0.25 ms at 128^3) feed nothing downstream: side branch
            dwo = torch.zeros_like(wo)
            dbo = _acc(bo) if _acc(bo) is not None else _zeros((NC,), xm)
            sb.hold(dlog, y24)
            sb.run(conv3_bwd_weight, dlog, y24, None, dwo, dbo, B, (4 * D, 4 * H, 4 * W), NC, True)
            dy24 = torch.empty_like(y24)
            conv3_bwd_data(dlog, wo, dy24, False, None, False, B, (4 * D, 4 * H, 4 * W), NC, True)
            view = _tail_view(D, H, W, Ch, T)
            if view is not None:
                drows = dy24           # the GEMMs read the fine grid through 5-D tensor maps
            else:
                drows = _empty((T, 64 * Ch), xm)
                block_permute(dy24, drows, B, D, H, W, 4, Ch, 64 * D * H * W * Ch, True)
            del dy24
            dwr, dbr64 = linear_bwd_weight_side(sb, drows, 64 * Ch, xn, 2 * E, T, 64 * Ch, 2 * E, w_is_kn=True, view=view)
            dxn = linear_bwd_data(drows, 64 * Ch, wr, T, 64 * Ch, 2 * E, w_is_kn=True, view=view)
            dxm, dxf, dn2w, dn2b = ln_bwd(dxn, xm, xf, n2w, mean, rstd, None, None, (B, D, H, W), beta=n2b)
        return dxm, dxf, dn2w, dn2b, dwr, dbr64, dwo, None, _gret(bo, dbo)


class DiceBceLossFn(torch.autograd.Function):
    """MDiceLoss.forward (reference loss/dice.py:158-166) in one reduction pass + closed-form backward.
    ``target``: float32 one-hot (the reference contract) or uint8 / bool one-hot (read as bytes, no float copy).
    ``group`` (a torch.distributed process group or None): all-reduce the 4*C partial sums so that the Dice
    terms span the GLOBAL batch; None keeps the reference's per-process semantics.
    ``w_dice``, ``w_bce``: (0.7, 0.3) for MDiceLoss, (1, 0) for MDiceLoss_Val (loss/dice.py:216-221).

    With a process group the backward yields d(global loss)/d(local logits); data-parallel gradient averaging
    (GradSync, 1/world) must then be replaced by a SUM -- ``MDiceLoss.grad_reduce`` says which."""

    @staticmethod
    def forward(ctx, logits, target, group, world, w_dice=0.7, w_bce=0.3):
        N.check_cuda_f32(logits)
        if logits.shape != target.shape:
            raise RuntimeError(f"MDiceLoss: logits {tuple(logits.shape)} vs target {tuple(target.shape)}")
        u8 = target.dtype in (torch.uint8, torch.bool)
        if not u8 and target.dtype != torch.float32:
            raise RuntimeError(f"MDiceLoss: target dtype {target.dtype} (float32, uint8 or bool one-hot expected)")
        if not (target.is_cuda and target.is_contiguous()):
            raise RuntimeError("MDiceLoss: target must be a contiguous CUDA tensor")
        B, C = logits.shape[:2]
        S = logits[0, 0].numel()
        sums = torch.zeros(C * 4, device=logits.device, dtype=torch.float64)
        if u8:
            N.call("mic_dice_bce_partial_u8", N.ptr(logits), N.ptr(target), N.ptr(sums), B, C, S)
        else:
            N.call("mic_dice_bce_partial", N.ptr(logits), N.ptr(target), N.ptr(sums), B, C, S)
        n = float(B * S)
        if group is not None and world > 1:
            torch.distributed.all_reduce(sums, group=group)
            n *= world
        loss = _empty((), logits)
        coef = _empty((C * 3,), logits)
        N.call("mic_dice_bce_finalize_weighted", N.ptr(sums), N.ptr(loss), N.ptr(coef), C, n, float(w_dice), float(w_bce))
        ctx.save_for_backward(logits, target, coef)
        ctx.u8 = u8
        ctx.n = n
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        logits, target, coef = ctx.saved_tensors
        B, C = logits.shape[:2]
        S = logits[0, 0].numel()
        dl = torch.empty_like(logits)
        dloss = dloss.contiguous().float()
        if ctx.u8:
            N.call("mic_dice_bce_bwd_u8", N.ptr(logits), N.ptr(target), N.ptr(coef), N.ptr(dloss), N.ptr(dl), B, C, S)
        else:
            N.call("mic_dice_bce_bwd", N.ptr(logits), N.ptr(target), N.ptr(coef), N.ptr(dloss), N.ptr(dl), B, C, S, ctx.n)
        return dl, None, None, None, None, None


class DeformSampleFn(torch.autograd.Function):
    """SpatialTransformer.forward (reference models/STN.py:9-32) on channels-last tensors:
    src (B,D,H,W,C), pos (B,D,H,W,3) -> (B,D,H,W,C)."""

    @staticmethod
    def forward(ctx, src, pos):
        N.check_cuda_f32(src, pos)
        B, D, H, W, C = src.shape
        out = torch.empty_like(src)
        N.call("mic_deform_sample_fwd", N.ptr(src), N.ptr(pos), N.ptr(out), B, D, H, W, D, H, W, C)
        ctx.save_for_backward(src, pos)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        src, pos = ctx.saved_tensors
        B, D, H, W, C = src.shape
        dsrc = torch.zeros_like(src)
        dpos = torch.empty_like(pos)
        N.call("mic_deform_sample_bwd", N.ptr(dout.contiguous()), N.ptr(src), N.ptr(pos), N.ptr(dsrc), N.ptr(dpos), B, D, H,
               W, D, H, W, C)
        return dsrc, dpos
