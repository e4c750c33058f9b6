"""Drop-in ``nn.Module`` surface of the reference's ``MicFormer/models/MICFormer_self.py``.

Same class names, constructor signatures, attribute names, parameter registration ORDER (so
``torch.manual_seed(s); Head(...)`` draws the same initial weights) and ``state_dict`` keys (1626 tensors, no
buffers, for the train config) -- a reference checkpoint loads unchanged and vice-versa.  ``nn.Linear`` /
``nn.Conv3d`` / ``nn.LayerNorm`` objects are kept purely as parameter containers; their ``forward`` is never
called.  All arithmetic runs in the sm_100a kernels behind ``micformer_b200.ops`` (no CPU / torch-op fallback:
calling a module on a CPU tensor raises).

Reference line numbers (``M:``) refer to /root/reference/MicFormer/models/MICFormer_self.py.
"""
from __future__ import annotations

from functools import reduce
from operator import mul

import os
import torch
import torch.nn as nn

from .. import fused, ops
from .STN import SpatialTransformer, Re_SpatialTransformer  # noqa: F401  (re-exported like the reference, M:13)


class DropPath(nn.Module):
    """timm.models.layers.DropPath stand-in (the reference imports it at M:5).  Holds the rate; the per-sample
    Bernoulli(keep)/keep scale it draws is applied inside the fused residual epilogues."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def sample_scale(self, batch: int, device) -> "torch.Tensor | None":
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        mask = torch.empty(batch, device=device, dtype=torch.float32).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return mask

    def forward(self, x):  # kept for API compatibility; not on the fused path
        s = self.sample_scale(x.shape[0], x.device)
        return x if s is None else x * s.view(-1, *([1] * (x.ndim - 1)))

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


import os as _os

_TWO_STREAMS = _os.environ.get("MICFORMER_TWO_STREAMS", "1") != "0"
if _TWO_STREAMS and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
    # shared modules (patch_embed, samplers, norm) run on both branch streams by design
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
_SIDE = {}


def _side_stream(device):
    st = _SIDE.get(device)
    if st is None:
        # MICFORMER_STREAM_PRIO=1: the modality stream (like the caller's stream, see bench.py) outranks the auxiliary
        # weight-gradient streams of ops.side_branch, so pending CTAs of the dependent chain are placed first
        st = _SIDE[device] = torch.cuda.Stream(device=device, priority=-1 if os.environ.get("MICFORMER_STREAM_PRIO", "0") == "1" else 0)
    return st


def _drop_scale(mod, batch, device):
    return mod.sample_scale(batch, device) if isinstance(mod, DropPath) else None


def _block_scales(block, batch, device):
    """Per-sample DropPath scales of a block's two residual branches: pre-drawn in one shot by the enclosing
    MicFormer (``_predraw_drop_path``) or drawn here when the block is used stand-alone."""
    pre = block.__dict__.pop("_dp_scales", None)
    if pre is not None:
        return pre
    return _drop_scale(block.drop_path, batch, device), _drop_scale(block.drop_path, batch, device)


def _current(img):
    fused.ensure_current(img)
    return img


def _attn_images(attn):
    """bf16 weight images of a (Cross)WindowAttention3D's q / kv / proj for the fused attention kernels (cached)"""
    qw, kvw, pw = attn.q.weight, attn.kv.weight, attn.proj.weight
    if not qw.is_cuda:
        return None
    img = attn.__dict__.get("_mic_img")
    if img is None or not img.valid():
        img = attn.__dict__["_mic_img"] = fused.attn_images(qw.detach(), kvw.detach(), pw.detach())
    return img


# MICFORMER_CONV_LAYOUTS=1: conv_offset weights reach the kernels through persistent layout buffers rewritten once per forward
# (one launch) and their gradient is written in the parameter's own layout: 144 fewer torch copy kernels per step, but the
# strided gradient flush costs 5-8 us per cross block on the weight-gradient branch (21.45 vs 21.25 ms/step): off by default
_CONV_LAYOUTS = _os.environ.get("MICFORMER_CONV_LAYOUTS", "0") == "1"


def _conv_layouts(conv):
    """persistent [27][Cin][Co] / [27][Co][Cin] copies of a Conv3d(k=3) weight for the conv kernels (cached on the module)"""
    lay = conv.__dict__.get("_mic_lay")
    if lay is None or not lay.valid():
        lay = conv.__dict__["_mic_lay"] = fused.ConvLayouts(conv.weight.detach())
    return lay


_QKV_CAT = _os.environ.get("MICFORMER_QKV_CAT", "1") != "0"


def _qkv_cat(attn):
    """persistent q|kv weight / bias concatenation of a WindowAttention3D for the unfused self blocks (cached on the module)"""
    if not (_QKV_CAT and attn.q.weight.is_cuda):
        return None
    cat = attn.__dict__.get("_mic_cat")
    if cat is None or not cat.valid():
        with torch.no_grad():
            cat = attn.__dict__["_mic_cat"] = fused.QkvCat(attn.q.weight.detach(), None if attn.q.bias is None else attn.q.bias.detach(),
                                                           attn.kv.weight.detach(), None if attn.kv.bias is None else attn.kv.bias.detach())
    return cat


def _fused_block_images(block, attn, dims):
    """(attention images, MLP images) when BOTH halves of this block run as fused kernels for this geometry, else None"""
    C = block.dim
    if not fused.attn_supported(C, block.num_heads, block.window_size, dims):
        return None
    mimg = block.mlp.fused_images()
    if mimg is None or attn.q.bias is None:
        return None
    return _attn_images(attn), mimg


class Mlp(nn.Module):
    """M:16-34.  Parameter container; fused as LN -> fc1 -> GELU -> fc2 -> +residual inside the block ops."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if drop != 0.0:
            raise NotImplementedError("micformer_b200: dropout p>0 is not on the reference's path (p=0 everywhere)")
        if act_layer is not nn.GELU:
            raise NotImplementedError("micformer_b200: only exact-erf GELU is built")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def fused_images(self):
        """bf16 weight images for the fused tcgen05 MLP kernels (None when this size / mode is not fused)."""
        w1, w2 = self.fc1.weight, self.fc2.weight
        if not (w1.is_cuda and fused.mlp_supported(w1.shape[1], w1.shape[0])):
            return None
        img = self.__dict__.get("_mic_img")
        if img is None or not img.valid():
            img = self.__dict__["_mic_img"] = fused.mlp_images(w1.detach(), w2.detach())
        return img

    def forward(self, x):
        shape = x.shape
        C = shape[-1]
        T = x.numel() // C
        x = x.contiguous()
        Hd = self.fc1.weight.shape[0]
        return _MlpOnlyFn.apply(x.view(T, C), self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias).view(
            *shape[:-1], self.fc2.weight.shape[0])


class _MlpOnlyFn(torch.autograd.Function):
    """Standalone fc1 -> GELU -> fc2 (only used when ``Mlp`` is called directly; blocks fuse it)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        ops.N.check_cuda_f32(x, w1, b1, w2, b2)
        T, C = x.shape
        Hd, Co = w1.shape[0], w2.shape[0]
        hpre = torch.empty(T, Hd, device=x.device)
        h = ops.linear_fwd(x, C, w1, b1, T, Hd, C, act=True, pre=hpre)
        y = ops.linear_fwd(h, Hd, w2, b2, T, Co, Hd)
        ctx.save_for_backward(x, hpre, h, w1, w2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, hpre, h, w1, w2 = ctx.saved_tensors
        T, C = x.shape
        Hd, Co = w1.shape[0], w2.shape[0]
        dy = dy.contiguous()
        dh = ops.linear_bwd_data(dy, Co, w2, T, Co, Hd, gelu_pre=hpre)
        dw2, db2 = ops.linear_bwd_weight(dy, Co, h, Hd, T, Co, Hd)
        dx = ops.linear_bwd_data(dh, Hd, w1, T, Hd, C)
        dw1, db1 = ops.linear_bwd_weight(dh, Hd, x, C, T, Hd, C)
        return dx, dw1, db1, dw2, db2


def window_partition(x, window_size):
    """M:37-50 (kept for API parity; the kernels never materialise windows)."""
    B, D, H, W, C = x.shape
    x = x.view(B, D // window_size[0], window_size[0], H // window_size[1], window_size[1], W // window_size[2],
               window_size[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, reduce(mul, window_size), C)


def window_reverse(windows, window_size, B, D, H, W):
    """M:117-132."""
    x = windows.view(B, D // window_size[0], H // window_size[1], W // window_size[2], window_size[0], window_size[1],
                     window_size[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(B, D, H, W, -1)


def get_window_size(x_size, window_size, shift_size=None):
    """M:135-145."""
    use_window_size = list(window_size)
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            use_window_size[i] = x_size[i]
    return tuple(use_window_size)


class _WindowAttentionBase(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if qk_scale is not None or attn_drop != 0.0 or proj_drop != 0.0:
            raise NotImplementedError("micformer_b200: qk_scale / attention dropout are not on the reference's path")
        if not qkv_bias:
            raise NotImplementedError("micformer_b200: qkv_bias=False is not built (the model always passes True, M:913)")
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    def _params(self):
        return (self.q.weight, self.q.bias, self.kv.weight, self.kv.bias, self.proj.weight, self.proj.bias)


class CrossWindowAttention3D(_WindowAttentionBase):
    """M:148-203.  ``forward(x, xa)`` on pre-partitioned windows (B_, N, C) / (B_, M, C) with M == N."""

    def forward(self, x, xa):
        return _WindowedAttnFn.apply(x.contiguous(), xa.contiguous(), self.num_heads, *self._params())


class WindowAttention3D(_WindowAttentionBase):
    """M:206-261."""

    def forward(self, x):
        x = x.contiguous()
        return _WindowedAttnFn.apply(x, x, self.num_heads, *self._params())


class _WindowedAttnFn(torch.autograd.Function):
    """q/kv Linear -> softmax(q k^T * scale) v -> proj on already-partitioned windows (B_, N, C): each window is a
    (1,1,N) grid with window (1,1,N).  Used by the standalone attention modules and the kernel-isolation bench."""

    @staticmethod
    def forward(ctx, x, xa, heads, qw, qb, kvw, kvb, pw, pb):
        ops.N.check_cuda_f32(x, xa, qw, qb, kvw, kvb, pw, pb)
        B_, Ntok, C = x.shape
        if xa.shape != x.shape:
            raise NotImplementedError("micformer_b200: cross attention with M != N is not on the reference's path")
        P = B_ * Ntok
        qkv = torch.empty(P, 3 * C, device=x.device)
        ops.linear_fwd(x, C, qw, qb, P, C, C, out=qkv, out_col=0, ldy=3 * C)
        ops.linear_fwd(xa, C, kvw, kvb, P, 2 * C, C, out=qkv, out_col=C, ldy=3 * C)
        o, lse = ops.window_attn_fwd(qkv, C, heads, B_, (1, 1, Ntok), (1, 1, Ntok))
        y = ops.linear_fwd(o, C, pw, pb, P, C, C)
        ctx.save_for_backward(x, xa, qkv, o, lse, qw, kvw, pw)
        ctx.heads = heads
        return y.view(B_, Ntok, C)

    @staticmethod
    def backward(ctx, dy):
        x, xa, qkv, o, lse, qw, kvw, pw = ctx.saved_tensors
        heads = ctx.heads
        B_, Ntok, C = x.shape
        P = B_ * Ntok
        dy = dy.contiguous()
        do = ops.linear_bwd_data(dy, C, pw, P, C, C)
        dpw, dpb = ops.linear_bwd_weight(dy, C, o, C, P, C, C)
        dqkv = ops.window_attn_bwd(qkv, o, do, lse, C, heads, B_, (1, 1, Ntok), (1, 1, Ntok))
        dx = ops.linear_bwd_data(dqkv, 3 * C, qw, P, C, C).view(x.shape)
        dxa = ops.linear_bwd_data(dqkv, 3 * C, kvw, P, 2 * C, C, dy_col=C).view(x.shape)
        dqw, dqb = ops.linear_bwd_weight(dqkv, 3 * C, x, C, P, C, C)
        dkvw, dkvb = ops.linear_bwd_weight(dqkv, 3 * C, xa, C, P, 2 * C, C, dy_col=C)
        return dx, dxa, None, dqw, dqb, dkvw, dkvb, dpw, dpb


class LayerNormProxy(nn.Module):
    """M:263-273 (parameter container: LN over 16 channels is fused into the offset head kernel)."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.dim = dim


class CrossTransformerBlock3D(nn.Module):
    """M:277-426.  ``forward(x, xa)`` on (B, D, H, W, C)."""

    def __init__(self, dim, num_heads, window_size=(4, 4, 4), hidden_channels=16, kk=3, offset_range_factor=2,
                 mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, use_checkpoint=False):
        super().__init__()
        if kk != 3 or hidden_channels != 16 or offset_range_factor < 0 or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("micformer_b200: only kk=3, hidden_channels=16, offset_range_factor>=0, LayerNorm "
                                      "(the reference's only configuration) are built")
        self.dim = dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.mlp_ratio = mlp_ratio
        self.use_checkpoint = use_checkpoint      # accepted, ignored (dead path in the reference)
        self.hidden_channels = hidden_channels
        self.kk = kk
        self.offset_range_factor = offset_range_factor

        self.norm1 = norm_layer(dim)
        self.cross_attn = CrossWindowAttention3D(dim, window_size=self.window_size, num_heads=num_heads,
                                                 qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                                 proj_drop=drop)
        self.conv_offset = nn.Sequential(
            nn.Conv3d(dim * 2, self.hidden_channels, self.kk, 1, self.kk // 2),
            LayerNormProxy(self.hidden_channels),
            nn.GELU(),
            nn.Conv3d(self.hidden_channels, 3, 1, 1, 0, bias=False))
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.stn = SpatialTransformer()

    def forward(self, x, xa):
        s1, s2 = _block_scales(self, x.shape[0], x.device)
        a = self.cross_attn
        co = self.conv_offset
        imgs = _fused_block_images(self, a, x.shape[1:4]) if x.is_cuda else None
        w3 = co[3].weight.reshape(3, self.hidden_channels)
        if x.is_cuda and co[0].weight.is_contiguous() and _CONV_LAYOUTS:
            # the conv kernels read layout copies of conv_offset.0.weight that are rewritten once per model forward; the weight's
            # gradient comes back in the parameter's own layout (cwp)
            lay = _current(_conv_layouts(co[0]))
            cw, cwk, cwp = lay.tcio, lay.toci, co[0].weight
        else:
            cw = co[0].weight.permute(2, 3, 4, 1, 0).reshape(27, 2 * self.dim, self.hidden_channels).contiguous()
            cwk = co[0].weight.detach().permute(2, 3, 4, 0, 1).reshape(27, self.hidden_channels, 2 * self.dim).contiguous()
            cwp = None
        if imgs is not None:
            return ops.FusedCrossBlockFn.apply(
                x.contiguous(), xa.contiguous(), s1, s2, self.num_heads, _current(imgs[0]), _current(imgs[1]),
                self.norm1.weight, self.norm1.bias, a.q.weight, a.q.bias, a.kv.weight, a.kv.bias, a.proj.weight, a.proj.bias,
                cw, cwk, co[0].bias, co[1].norm.weight, co[1].norm.bias, w3, self.norm2.weight, self.norm2.bias,
                self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias, cwp)
        if ops.N.get_gemm_mode() != 1:      # only the tcgen05 forward reads the weight as [tap][out][in]
            cwk = None
        return ops.CrossBlockFn.apply(
            x.contiguous(), xa.contiguous(), s1, s2, self.num_heads, tuple(self.window_size),
            self.norm1.weight, self.norm1.bias, a.q.weight, a.q.bias, a.kv.weight, a.kv.bias, a.proj.weight,
            a.proj.bias, cw, cwk, co[0].bias, co[1].norm.weight, co[1].norm.bias, w3, self.norm2.weight, self.norm2.bias,
            self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias, _current(self.mlp.fused_images()), cwp)


class TransformerBlock3D(nn.Module):
    """M:430-524.  ``forward(x)`` on (B, D, H, W, C)."""

    def __init__(self, dim, num_heads, window_size=(4, 4, 4), hidden_channels=16, kk=3, offset_range_factor=2,
                 mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, use_checkpoint=False):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("micformer_b200: only nn.LayerNorm is built")
        self.dim = dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.mlp_ratio = mlp_ratio
        self.use_checkpoint = use_checkpoint
        self.hidden_channels = hidden_channels
        self.kk = kk
        self.offset_range_factor = offset_range_factor

        self.norm1 = norm_layer(dim)
        self.self_attn = WindowAttention3D(dim, window_size=self.window_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                           qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        s1, s2 = _block_scales(self, x.shape[0], x.device)
        a = self.self_attn
        imgs = _fused_block_images(self, a, x.shape[1:4]) if x.is_cuda else None
        if imgs is not None:
            return ops.FusedSelfBlockFn.apply(
                x.contiguous(), s1, s2, self.num_heads, _current(imgs[0]), _current(imgs[1]),
                self.norm1.weight, self.norm1.bias, a.q.weight, a.q.bias, a.kv.weight, a.kv.bias, a.proj.weight, a.proj.bias,
                self.norm2.weight, self.norm2.bias, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight,
                self.mlp.fc2.bias)
        return ops.SelfBlockFn.apply(
            x.contiguous(), s1, s2, self.num_heads, tuple(self.window_size),
            self.norm1.weight, self.norm1.bias, a.q.weight, a.q.bias, a.kv.weight, a.kv.bias, a.proj.weight,
            a.proj.bias, self.norm2.weight, self.norm2.bias, self.mlp.fc1.weight, self.mlp.fc1.bias,
            self.mlp.fc2.weight, self.mlp.fc2.bias, _current(self.mlp.fused_images()), _current(_qkv_cat(a)))


class PatchMerging(nn.Module):
    """M:527-561."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.down_conv = nn.Conv3d(dim, 2 * dim, (2, 2, 2), stride=2, padding=0)
        self.norm = norm_layer(2 * dim)

    def forward(self, x):
        w2 = self.down_conv.weight.permute(0, 2, 3, 4, 1).reshape(2 * self.dim, 8 * self.dim).contiguous()
        B, D, H, W, C = x.shape
        if (D % 2) or (H % 2) or (W % 2):       # odd sizes: one zero plane on the trailing side (M:551-555); torch plumbing
            x = torch.nn.functional.pad(x, (0, 0, 0, W % 2, 0, H % 2, 0, D % 2))
        return ops.PatchMergeFn.apply(x.contiguous(), w2, self.down_conv.bias, self.norm.weight, self.norm.bias)


class PatchExpand(nn.Module):
    """M:564-579."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.up_conv = nn.ConvTranspose3d(dim, dim // 2, (2, 2, 2), stride=2, padding=0)
        self.norm = norm_layer(dim // 2)

    def forward(self, x):
        Co = self.dim // 2
        wk = self.up_conv.weight.permute(0, 2, 3, 4, 1).reshape(self.dim, 8 * Co).contiguous()
        b8 = self.up_conv.bias.repeat(8)
        return ops.PatchExpandFn.apply(x.contiguous(), wk, b8, self.norm.weight, self.norm.bias)


class BasicLayer(nn.Module):
    """M:582-707: per depth two self blocks (one per stream) then the two cross blocks on the pre-update pair."""

    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.window_size = window_size
        self.depth = depth
        self.use_checkpoint = use_checkpoint

        def mk(cls):
            return nn.ModuleList([
                cls(dim=dim, num_heads=num_heads, window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                    qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                    drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path, norm_layer=norm_layer,
                    use_checkpoint=use_checkpoint) for i in range(depth)])

        self.blocks1 = mk(CrossTransformerBlock3D)
        self.blocks2 = mk(CrossTransformerBlock3D)
        self.self_blocks1 = mk(TransformerBlock3D)
        self.self_blocks2 = mk(TransformerBlock3D)
        self.downsample = downsample
        if self.downsample is not None:
            self.downsample = downsample(dim=dim, norm_layer=norm_layer)

    def forward(self, x, xa):
        if not (_TWO_STREAMS and x.is_cuda):
            for i in range(len(self.blocks1)):
                x, xa = self.self_blocks1[i](x), self.self_blocks2[i](xa)
                x, xa = self.blocks1[i](x, xa), self.blocks2[i](xa, x)
            if self.downsample is not None:
                return x, xa, self.downsample(x), self.downsample(xa)
            return x, xa, x, xa
        # The CT branch runs on the current stream, the MR branch on a side stream: the two self blocks, the two
        # cross blocks (both read the pre-update pair, M:700-701) and the two shared-sampler calls are independent,
        # and at batch 2 the deep stages are latency-bound (8..48 CTAs per kernel on 148 SMs).  Inside a captured
        # CUDA graph the two streams become parallel branches; autograd replays each backward on its forward stream.
        main = torch.cuda.current_stream()
        side = _side_stream(x.device)

        def on_side(fn, *args):
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out = fn(*args)
            for a in args:                 # inputs allocated on `main`, consumed on `side`
                a.record_stream(side)
            return out

        for i in range(len(self.blocks1)):
            xa_s = on_side(self.self_blocks2[i], xa)
            x_s = self.self_blocks1[i](x)
            main.wait_stream(side)
            xa_s.record_stream(main)
            xa = on_side(self.blocks2[i], xa_s, x_s)
            x = self.blocks1[i](x_s, xa_s)
            main.wait_stream(side)
            xa.record_stream(main)
        if self.downsample is None:
            return x, xa, x, xa
        xa_d = on_side(self.downsample, xa)
        x_d = self.downsample(x)
        main.wait_stream(side)
        xa_d.record_stream(main)
        return x, xa, x_d, xa_d


class PatchEmbed3D(nn.Module):
    """M:837-878."""

    def __init__(self, patch_size=(4, 4, 4), in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        if tuple(patch_size) != (4, 4, 4) or in_chans != 1 or norm_layer is not None:
            raise NotImplementedError("micformer_b200: PatchEmbed3D is built for patch 4^3, 1 input channel, no norm "
                                      "(what MicFormer instantiates, M:934-936 with patch_norm=False)")
        self.patch_size = patch_size
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = None

    def forward(self, x, channel: int = 0):
        """x (B, Cin, D, H, W) NCDHW; embeds channel ``channel`` -> (B, D/4, H/4, W/4, E) channels-last.
        (The reference returns NCDHW and permutes right after, M:1001-1002; the permute is folded in.)"""
        return ops.PatchEmbedFn.apply(x, channel, self.proj.weight, self.proj.bias)


class MicFormer(nn.Module):
    """M:881-1039."""

    def __init__(self, pretrained=None, pretrained2d=False, patch_size=(4, 4, 4), in_chans=1, embed_dim=64,
                 depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=(7, 7, 7), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=nn.LayerNorm,
                 patch_norm=False, frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if drop_rate != 0.0:
            raise NotImplementedError("micformer_b200: drop_rate>0 is not on the reference's path")
        self.pretrained = pretrained
        self.pretrained2d = pretrained2d
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.frozen_stages = frozen_stages
        self.window_size = window_size
        self.patch_size = patch_size

        self.patch_embed = PatchEmbed3D(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                        norm_layer=norm_layer if self.patch_norm else None)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]

        self.layers = nn.ModuleList()
        for i_layer in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i_layer), depth=depths[i_layer], num_heads=num_heads[i_layer],
                window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                attn_drop=attn_drop_rate, drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])],
                norm_layer=norm_layer, downsample=PatchMerging if i_layer < self.num_layers - 1 else None,
                use_checkpoint=use_checkpoint))

        self.up_layers = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for i_layer in reversed(range(self.num_layers)):
            concat_linear = nn.Linear(2 * int(embed_dim * 2 ** i_layer), int(embed_dim * 2 ** i_layer))
            up_layer = BasicLayer(
                dim=int(embed_dim * 2 ** i_layer), depth=depths[i_layer], num_heads=num_heads[i_layer],
                window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                attn_drop=attn_drop_rate, drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])],
                norm_layer=norm_layer, downsample=PatchExpand if i_layer > 0 else None, use_checkpoint=use_checkpoint)
            self.up_layers.append(up_layer)
            self.concat_back_dim.append(concat_linear)

        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.norm = norm_layer(self.num_features)
        self.norm2 = norm_layer(self.embed_dim * 2)
        self.reverse_patch_embedding = nn.ConvTranspose3d(2 * embed_dim, embed_dim // 2, (4, 4, 4), stride=4)

    # -- pieces -------------------------------------------------------------------------------------------
    def _predraw_drop_path(self, batch, device):
        """Training mode: draw the Bernoulli(keep)/keep scales of ALL blocks' two residual branches with three
        torch kernels (one rand, one compare, one divide) instead of two tiny kernels per branch."""
        blocks = [b for b in self.modules() if isinstance(b, (CrossTransformerBlock3D, TransformerBlock3D))
                  and isinstance(b.drop_path, DropPath) and b.drop_path.drop_prob > 0.0]
        if not blocks or not self.training or self.__dict__.get("_dp_external", False):
            return          # (_dp_external: a caller supplies every block's `_dp_scales` itself -- shared-mask parity runs)
        keep = getattr(self, "_dp_keep", None)
        if keep is None or keep.device != device or keep.numel() != len(blocks):
            keep = torch.tensor([1.0 - b.drop_path.drop_prob for b in blocks], device=device).view(-1, 1, 1)
            self.__dict__["_dp_keep"] = keep
        scales = (torch.rand(len(blocks), 2, batch, device=device) < keep).float() / keep
        for i, b in enumerate(blocks):
            b.__dict__["_dp_scales"] = (scales[i, 0], scales[i, 1])

    def _trunk(self, vol):
        """Everything up to (not including) cat -> norm2 -> reverse_patch_embedding; vol is (B, 2, D, H, W)."""
        self._predraw_drop_path(vol.shape[0], vol.device)
        D0, H0, W0 = vol.shape[2:]
        if (D0 % 4) or (H0 % 4) or (W0 % 4):    # PatchEmbed3D pads to a multiple of the patch size (M:864-869); torch plumbing
            vol = torch.nn.functional.pad(vol, (0, (4 - W0 % 4) % 4, 0, (4 - H0 % 4) % 4, 0, (4 - D0 % 4) % 4)).contiguous()
        # the optimizer changed the weights since the last forward: rebuild all fused-kernel weight images in one launch
        imgs = [m.fused_images() for m in self.modules() if isinstance(m, Mlp)]
        if fused.enabled() and vol.is_cuda:
            for m in self.modules():
                if isinstance(m, (CrossTransformerBlock3D, TransformerBlock3D)) and (m.dim, m.dim // m.num_heads) in fused.FUSED_ATTN \
                        and tuple(m.window_size) == (2, 2, 2):
                    imgs.append(_attn_images(m.cross_attn if isinstance(m, CrossTransformerBlock3D) else m.self_attn))
        convs = [_conv_layouts(m.conv_offset[0]) for m in self.modules()
                 if isinstance(m, CrossTransformerBlock3D) and m.conv_offset[0].weight.is_contiguous()] if (vol.is_cuda and _CONV_LAYOUTS) else []
        cats = []
        if vol.is_cuda and _QKV_CAT:
            for m in self.modules():
                if isinstance(m, TransformerBlock3D) and not (fused.enabled() and (m.dim, m.dim // m.num_heads) in fused.FUSED_ATTN
                                                              and tuple(m.window_size) == (2, 2, 2) and m.mlp.fused_images() is not None):
                    cats.append(_qkv_cat(m.self_attn))
        fused.model_refresh([i for i in imgs if i is not None], convs, [c for c in cats if c is not None])
        moving = self.patch_embed(vol, 0)
        fixed = self.patch_embed(vol, 1)
        feats_m, feats_f = [], []
        for layer in self.layers:
            mo, fo, moving, fixed = layer(moving, fixed)
            feats_m.append(mo)
            feats_f.append(fo)
        moving = ops.LayerNormFn.apply(moving, None, self.norm.weight, self.norm.bias)
        fixed = ops.LayerNormFn.apply(fixed, None, self.norm.weight, self.norm.bias)
        hook = self.__dict__.get("_bottleneck_grad_hook")
        if hook is not None and moving.requires_grad:
            # data parallelism (parallel.GradSync.enable_overlap): when the backward pass arrives here the decoder's gradients
            # are complete and their all-reduce can start under the encoder's backward
            moving.register_hook(hook)
            fixed.register_hook(hook)
        L = self.num_layers
        for inx, layer_up in enumerate(self.up_layers):
            if inx > 0:
                skip_m, skip_f = feats_m[L - 1 - inx], feats_f[L - 1 - inx]   # reference hard-codes 3 - inx (M:1018-1028)
                if moving.shape != skip_m.shape:
                    # odd sizes (volumes not divisible by 32): the up-sampled map is resized to the skip's grid, trilinear with
                    # align_corners=True (M:1018-1025).  Rare validation-time branch: torch's interpolate (plumbing), not a kernel
                    size = tuple(skip_m.shape[1:4])
                    rs = lambda t: torch.nn.functional.interpolate(t.permute(0, 4, 1, 2, 3), size=size, mode="trilinear",
                                                                   align_corners=True).permute(0, 2, 3, 4, 1).contiguous()
                    moving, fixed = rs(moving), rs(fixed)
                lin = self.concat_back_dim[inx]
                moving = ops.SkipLinearFn.apply(moving, skip_m, lin.weight, lin.bias)
                fixed = ops.SkipLinearFn.apply(fixed, skip_f, lin.weight, lin.bias)
            _, _, moving, fixed = layer_up(moving, fixed)
        return moving, fixed

    def _tail_params(self):
        E = self.embed_dim
        Ch = E // 2
        wr = self.reverse_patch_embedding.weight.permute(0, 2, 3, 4, 1).reshape(2 * E, 64 * Ch).contiguous()
        br64 = self.reverse_patch_embedding.bias.repeat(64)
        return wr, br64

    def forward(self, moving, fixed):
        """Reference signature: two (B,1,D,H,W) volumes -> (B, E/2, D, H, W) features (M:992-1039), any embed_dim.
        ``Head`` does not come through here: it fuses this tail with ``out_conv`` (``ops.SegHeadFn``)."""
        vol = torch.cat([moving, fixed], dim=1).contiguous()
        if vol.dtype != torch.float32:
            vol = vol.float()
        m, f = self._trunk(vol)
        wr, br64 = self._tail_params()
        y = ops.UnpatchFn.apply(m, f, self.norm2.weight, self.norm2.bias, wr, br64)      # channels-last
        return y.permute(0, 4, 1, 2, 3).contiguous()                                      # NCDHW like the reference


class Head(nn.Module):
    """M:1042-1055: the module ``train_mmwhs_noPad.py:92`` builds.  forward (B,2,D,H,W) -> (B,num_classes,D,H,W)."""

    def __init__(self, n_channels=1, embed_dim=96, num_classes=14, window_size=(2, 2, 2)):
        super().__init__()
        self.swin = MicFormer(window_size=window_size, in_chans=n_channels, embed_dim=embed_dim)
        self.out_conv = nn.Conv3d(embed_dim // 2, num_classes, 3, padding=1)

    def forward(self, x):
        if x.dim() != 5 or x.shape[1] != 2:
            raise ValueError(f"Head expects (B, 2, D, H, W) [CT, MR]; got {tuple(x.shape)}")   # torch.split unpack, M:1050
        if x.dtype != torch.float32:
            x = x.float()          # validation runs under autocast with fp16 inputs (utils.py:222-240)
        x = x.contiguous()
        m, f = self.swin._trunk(x)
        wr, br64 = self.swin._tail_params()
        Ch = self.swin.embed_dim // 2
        NC = self.out_conv.weight.shape[0]
        wo = self.out_conv.weight.permute(2, 3, 4, 1, 0).reshape(27, Ch, NC).contiguous()
        wok = None
        if ops.N.get_gemm_mode() == 1:
            wok = self.out_conv.weight.detach().permute(2, 3, 4, 0, 1).reshape(27, NC, Ch).contiguous()
        return ops.SegHeadFn.apply(m, f, self.swin.norm2.weight, self.swin.norm2.bias, wr, br64, wo, wok,
                                   self.out_conv.bias)
