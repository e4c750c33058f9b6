"""CPU oracle for the MicFormer dual-stream hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the reference algorithm in plain functional
torch-on-CPU code over a flat ``state_dict``.  It exists so that the CUDA path in
``micformer_b200`` can be checked on a machine where ``/root/reference`` does not exist.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package never imports it.

Pinning: the reference ships no tests / golden vectors for this path, so the
oracle is pinned against outputs of the *unmodified reference executed in the build container*:
``tests/golden/make_golden.py`` imports ``/root/reference/MicFormer`` (with a 10-line stub for the
one missing third-party symbol, ``timm.models.layers.DropPath``), runs it on seeded inputs and
commits the vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them here.

Every function cites the reference lines it restates (paths relative to ``/root/reference``;
``M:`` = ``MicFormer/models/MICFormer_self.py``, ``S:`` = ``MicFormer/models/STN.py``,
``L:`` = ``MicFormer/loss/dice.py``).  The arithmetic itself lives in third-party ATen
(torch 2.11.0 is the operative version; the reference pins none).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-5  # nn.LayerNorm default, M:308,321,540,569,987-988,267


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class Config:
    """Hyper-parameters of ``Head`` / ``MicFormer`` (M:903-921, M:1042-1046)."""
    embed_dim: int = 48
    num_classes: int = 8
    window_size: Tuple[int, int, int] = (2, 2, 2)
    depths: Tuple[int, ...] = (2, 2, 6, 2)
    num_heads: Tuple[int, ...] = (3, 6, 12, 24)
    mlp_ratio: float = 4.0
    drop_path_rate: float = 0.2
    in_chans: int = 1
    hidden_channels: int = 16  # CrossTransformerBlock3D default, M:295

    @property
    def num_layers(self) -> int:
        return len(self.depths)

    def drop_path_rates(self) -> List[float]:
        # M:941 -- torch.linspace(0, rate, sum(depths))
        return [x.item() for x in torch.linspace(0, self.drop_path_rate, sum(self.depths))]


TRAIN = Config()                                                    # train_mmwhs_noPad.py:92
W7 = Config(embed_dim=96, window_size=(7, 7, 7))                    # BASELINE.json configs[3] family
TINY = Config(embed_dim=24, depths=(2, 2, 1, 1), num_heads=(2, 2, 2, 2))  # 4-stage tiny


# --------------------------------------------------------------------------------------
# parameter shapes + deterministic synthetic weights (independent of module build order)
# --------------------------------------------------------------------------------------
def _block_shapes(prefix: str, C: int, cross: bool, hidden: int, mlp_ratio: float) -> Dict[str, Tuple[int, ...]]:
    attn = "cross_attn" if cross else "self_attn"
    Hd = int(C * mlp_ratio)
    s = {
        f"{prefix}.norm1.weight": (C,), f"{prefix}.norm1.bias": (C,),
        f"{prefix}.{attn}.q.weight": (C, C), f"{prefix}.{attn}.q.bias": (C,),
        f"{prefix}.{attn}.kv.weight": (2 * C, C), f"{prefix}.{attn}.kv.bias": (2 * C,),
        f"{prefix}.{attn}.proj.weight": (C, C), f"{prefix}.{attn}.proj.bias": (C,),
    }
    if cross:
        s.update({
            f"{prefix}.conv_offset.0.weight": (hidden, 2 * C, 3, 3, 3), f"{prefix}.conv_offset.0.bias": (hidden,),
            f"{prefix}.conv_offset.1.norm.weight": (hidden,), f"{prefix}.conv_offset.1.norm.bias": (hidden,),
            f"{prefix}.conv_offset.3.weight": (3, hidden, 1, 1, 1),
        })
    s.update({
        f"{prefix}.norm2.weight": (C,), f"{prefix}.norm2.bias": (C,),
        f"{prefix}.mlp.fc1.weight": (Hd, C), f"{prefix}.mlp.fc1.bias": (Hd,),
        f"{prefix}.mlp.fc2.weight": (C, Hd), f"{prefix}.mlp.fc2.bias": (C,),
    })
    return s


def param_shapes(cfg: Config) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape, in the reference's registration order (M:934-990, M:1045-1046)."""
    E, L = cfg.embed_dim, cfg.num_layers
    s: Dict[str, Tuple[int, ...]] = {
        "swin.patch_embed.proj.weight": (E, cfg.in_chans, 4, 4, 4), "swin.patch_embed.proj.bias": (E,)}

    def layer(prefix: str, C: int, depth: int, sampler: Optional[str]):
        for name, cross in (("blocks1", True), ("blocks2", True), ("self_blocks1", False), ("self_blocks2", False)):
            for j in range(depth):
                s.update(_block_shapes(f"{prefix}.{name}.{j}", C, cross, cfg.hidden_channels, cfg.mlp_ratio))
        if sampler == "down":   # PatchMerging M:539-540
            s[f"{prefix}.downsample.down_conv.weight"] = (2 * C, C, 2, 2, 2)
            s[f"{prefix}.downsample.down_conv.bias"] = (2 * C,)
            s[f"{prefix}.downsample.norm.weight"] = (2 * C,)
            s[f"{prefix}.downsample.norm.bias"] = (2 * C,)
        elif sampler == "up":   # PatchExpand M:568-569
            s[f"{prefix}.downsample.up_conv.weight"] = (C, C // 2, 2, 2, 2)
            s[f"{prefix}.downsample.up_conv.bias"] = (C // 2,)
            s[f"{prefix}.downsample.norm.weight"] = (C // 2,)
            s[f"{prefix}.downsample.norm.bias"] = (C // 2,)

    for i in range(L):
        layer(f"swin.layers.{i}", E * 2 ** i, cfg.depths[i], "down" if i < L - 1 else None)
    for k, i in enumerate(reversed(range(L))):
        layer(f"swin.up_layers.{k}", E * 2 ** i, cfg.depths[i], "up" if i > 0 else None)
    for k, i in enumerate(reversed(range(L))):
        C = E * 2 ** i
        s[f"swin.concat_back_dim.{k}.weight"] = (C, 2 * C)
        s[f"swin.concat_back_dim.{k}.bias"] = (C,)
    C = E * 2 ** (L - 1)
    s["swin.norm.weight"] = (C,); s["swin.norm.bias"] = (C,)
    s["swin.norm2.weight"] = (2 * E,); s["swin.norm2.bias"] = (2 * E,)
    s["swin.reverse_patch_embedding.weight"] = (2 * E, E // 2, 4, 4, 4)
    s["swin.reverse_patch_embedding.bias"] = (E // 2,)
    s["out_conv.weight"] = (cfg.num_classes, E // 2, 3, 3, 3)
    s["out_conv.bias"] = (cfg.num_classes,)
    return s


def _key_seed(key: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def synth_state_dict(cfg: Config, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic weights that do not depend on module construction order (one CPU generator per key).

    Magnitudes mimic torch's default init (uniform(+-1/sqrt(fan_in)); LayerNorm weight near 1) but are
    deliberately *not* trivial for LayerNorm so gamma/beta paths are exercised."""
    out: Dict[str, Tensor] = {}
    for key, shape in param_shapes(cfg).items():
        g = torch.Generator().manual_seed(_key_seed(key, seed))
        is_norm = ".norm" in key and len(shape) == 1
        if is_norm and key.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if len(shape) == 1:  # bias: use a fan_in guess that keeps it small but non-zero
                fan_in = max(shape[0], 16)
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[key] = t.to(dtype)
    return out


def synth_inputs(B: int, S: int, num_classes: int = 8, seed: int = 1, dtype=torch.float32):
    """N(0,1) dual-modality volume + random one-hot labels (float 0/1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 2, S, S, S, generator=g).to(dtype)
    lab = torch.randint(0, num_classes, (B, S, S, S), generator=g)
    onehot = F.one_hot(lab, num_classes).permute(0, 4, 1, 2, 3).contiguous().to(dtype)
    return x, onehot


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def mlp(x: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    """M:28-34 -- fc1 -> exact-erf GELU -> fc2 (dropout p=0)."""
    h = F.linear(x, p[f"{pre}.fc1.weight"], p[f"{pre}.fc1.bias"])
    h = F.gelu(h)  # nn.GELU() default approximate='none'
    return F.linear(h, p[f"{pre}.fc2.weight"], p[f"{pre}.fc2.bias"])


def get_window_size(x_size: Sequence[int], window_size: Sequence[int]) -> Tuple[int, ...]:
    """M:135-145 -- clamp the window to the feature size per axis."""
    return tuple(min(x, w) if x <= w else w for x, w in zip(x_size, window_size))


def window_partition(x: Tensor, ws: Sequence[int]) -> Tensor:
    """M:37-50 -- (B,D,H,W,C) -> (B*nW, wd*wh*ww, C)."""
    B, D, H, W, C = x.shape
    x = x.view(B, D // ws[0], ws[0], H // ws[1], ws[1], W // ws[2], ws[2], C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, ws[0] * ws[1] * ws[2], C)


def window_reverse(w: Tensor, ws: Sequence[int], B: int, D: int, H: int, W: int) -> Tensor:
    """M:117-132."""
    x = w.view(B, D // ws[0], H // ws[1], W // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(B, D, H, W, -1)


def window_attention(xq: Tensor, xkv: Tensor, p: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """M:179-203 (cross) and M:237-261 (self: xkv is xq).  No mask, no relative-position bias."""
    B_, N, C = xq.shape
    M = xkv.shape[1]
    d = C // heads
    scale = d ** -0.5
    q = F.linear(xq, p[f"{pre}.q.weight"], p[f"{pre}.q.bias"]).reshape(B_, N, heads, d).permute(0, 2, 1, 3)
    kv = F.linear(xkv, p[f"{pre}.kv.weight"], p[f"{pre}.kv.bias"]).reshape(B_, M, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = (q * scale) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(o, p[f"{pre}.proj.weight"], p[f"{pre}.proj.bias"])


def _pad_to_window(x: Tensor, ws: Sequence[int]) -> Tensor:
    """M:345-350 / M:479-483 -- zero-pad D,H,W (trailing side) to multiples of the window."""
    _, D, H, W, _ = x.shape
    pd = (ws[0] - D % ws[0]) % ws[0]
    pb = (ws[1] - H % ws[1]) % ws[1]
    pr = (ws[2] - W % ws[2]) % ws[2]
    if pd or pb or pr:
        x = F.pad(x, (0, 0, 0, pr, 0, pb, 0, pd))
    return x


def ref_points(Dk: int, Hk: int, Wk: int, dtype=torch.float32) -> Tensor:
    """M:326-337 -- NOTE the permuted normalisers: ch0 (z index) / H, ch1 (y) / W, ch2 (x) / D."""
    z = torch.linspace(0.5, Dk - 0.5, Dk, dtype=dtype)
    y = torch.linspace(0.5, Hk - 0.5, Hk, dtype=dtype)
    x = torch.linspace(0.5, Wk - 0.5, Wk, dtype=dtype)
    rz, ry, rx = torch.meshgrid(z, y, x, indexing="ij")
    ref = torch.stack((rz, ry, rx), -1)
    ref[..., 2].div_(Dk).mul_(2).sub_(1)
    ref[..., 1].div_(Wk).mul_(2).sub_(1)
    ref[..., 0].div_(Hk).mul_(2).sub_(1)
    return ref  # (Dk,Hk,Wk,3)


def offset_net(xn: Tensor, xa: Tensor, p: Dict[str, Tensor], pre: str) -> Tensor:
    """M:313-318, M:354-358 -- Conv3d(2C->16,k3,p1) -> LN(16) -> GELU -> Conv3d(16->3,k1,no bias).
    Inputs channels-last (B,Dp,Hp,Wp,C); returns offsets channels-last (B,Dp,Hp,Wp,3)."""
    cat = torch.cat([xn, xa], dim=-1).permute(0, 4, 1, 2, 3)
    h = F.conv3d(cat, p[f"{pre}.0.weight"], p[f"{pre}.0.bias"], padding=1)
    h = h.permute(0, 2, 3, 4, 1)
    h = layer_norm(h, p[f"{pre}.1.norm.weight"], p[f"{pre}.1.norm.bias"])
    h = F.gelu(h)
    return F.linear(h, p[f"{pre}.3.weight"].reshape(3, -1))  # 1x1x1 conv == per-voxel linear


def stn_sample(src: Tensor, pos: Tensor) -> Tensor:
    """S:9-32 via ATen grid_sample (same call as the reference).  src (B,C,D,H,W); pos (B,3,D,H,W).
    v_i = idx_i + pos_i ; g_i = 2*(v_i/(S_i-1) - .5) ; sample (x,y,z)=g[2,1,0], bilinear, zeros,
    align_corners=False."""
    shape = pos.shape[2:]
    grids = torch.meshgrid([torch.arange(0, s) for s in shape], indexing="ij")
    grid = torch.stack(grids).unsqueeze(0).to(device=pos.device, dtype=pos.dtype)
    new_locs = grid + pos
    comps = [2 * (new_locs[:, i] / (shape[i] - 1) - 0.5) for i in range(3)]
    g = torch.stack([comps[2], comps[1], comps[0]], dim=-1)
    return F.grid_sample(src, g, mode="bilinear", padding_mode="zeros", align_corners=False)


def stn_sample_closed_form(src_cl: Tensor, pos_cl: Tensor) -> Tensor:
    """Independent restatement of S:9-32 without grid_sample: trilinear gather at
    c_i = (idx_i + pos_i) * S_i/(S_i-1) - 0.5 with zero padding.  Channels-last in and out:
    src_cl (B,D,H,W,C), pos_cl (B,D,H,W,3) -> (B,D,H,W,C).  This is the formula the CUDA sampler
    implements; tests check it against ``stn_sample``."""
    B, D, H, W, C = src_cl.shape
    dev, dt = src_cl.device, src_cl.dtype
    sizes = (D, H, W)
    idx = torch.meshgrid([torch.arange(s, device=dev, dtype=dt) for s in sizes], indexing="ij")
    coord = []
    for i in range(3):
        v = idx[i].unsqueeze(0) + pos_cl[..., i]
        # same operation order as the reference: 2*(v/(S-1) - .5) then ((g+1)*S-1)/2
        g = 2 * (v / (sizes[i] - 1) - 0.5)
        coord.append(((g + 1) * sizes[i] - 1) / 2)
    f = [torch.floor(c) for c in coord]
    t = [c - fl for c, fl in zip(coord, f)]
    out = torch.zeros_like(src_cl)
    flat = src_cl.reshape(B, D * H * W, C)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                zi, yi, xi = f[0] + dz, f[1] + dy, f[2] + dx
                wgt = ((t[0] if dz else 1 - t[0]) * (t[1] if dy else 1 - t[1]) * (t[2] if dx else 1 - t[2]))
                ok = (zi >= 0) & (zi < D) & (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
                lin = (zi.clamp(0, D - 1) * H + yi.clamp(0, H - 1)) * W + xi.clamp(0, W - 1)
                val = torch.gather(flat, 1, lin.reshape(B, -1, 1).long().expand(-1, -1, C)).reshape(B, D, H, W, C)
                out = out + val * (wgt * ok.to(dt)).unsqueeze(-1)
    return out


def _drop_path(x: Tensor, rate: float, training: bool, gen: Optional[torch.Generator]) -> Tensor:
    """timm DropPath semantics (M:5,320,419,424): per-sample Bernoulli(keep)/keep; identity in eval."""
    if not training or rate == 0.0:
        return x
    keep = 1.0 - rate
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    mask = torch.empty(shape, dtype=x.dtype).bernoulli_(keep, generator=gen)     # drawn on the host (CPU generator)
    return x * mask.div_(keep).to(x.device)


def self_block(x: Tensor, p, pre: str, heads: int, window, dp=0.0, training=False, gen=None) -> Tensor:
    """TransformerBlock3D.forward M:473-524."""
    B, D, H, W, C = x.shape
    ws = get_window_size((D, H, W), window)
    h = _pad_to_window(layer_norm(x, p[f"{pre}.norm1.weight"], p[f"{pre}.norm1.bias"]), ws)
    _, Dp, Hp, Wp, _ = h.shape
    xw = window_partition(h, ws)
    aw = window_attention(xw, xw, p, f"{pre}.self_attn", heads)
    h = window_reverse(aw, ws, B, Dp, Hp, Wp)[:, :D, :H, :W, :]
    x = x + _drop_path(h, dp, training, gen)
    return x + _drop_path(mlp(layer_norm(x, p[f"{pre}.norm2.weight"], p[f"{pre}.norm2.bias"]), p, f"{pre}.mlp"),
                          dp, training, gen)


def cross_block_part1(x: Tensor, xa: Tensor, p, pre: str, heads: int, window, closed_form_stn=False) -> Tensor:
    """CrossTransformerBlock3D.forward_part1 M:339-401: only the query stream is normalised (M:343)."""
    B, D, H, W, C = x.shape
    ws = get_window_size((D, H, W), window)
    xn = _pad_to_window(layer_norm(x, p[f"{pre}.norm1.weight"], p[f"{pre}.norm1.bias"]), ws)
    xap = _pad_to_window(xa, ws)
    _, Dp, Hp, Wp, _ = xn.shape
    offsets = offset_net(xn, xap, p, f"{pre}.conv_offset")
    pos = offsets + ref_points(Dp, Hp, Wp, x.dtype).to(x.device).unsqueeze(0)   # offset_range_factor=2 >= 0, M:363-364
    if closed_form_stn:
        sampled = stn_sample_closed_form(xap, pos)
    else:
        sampled = stn_sample(xap.permute(0, 4, 1, 2, 3), pos.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
    aw = window_attention(window_partition(xn, ws), window_partition(sampled.contiguous(), ws), p,
                          f"{pre}.cross_attn", heads)
    return window_reverse(aw, ws, B, Dp, Hp, Wp)[:, :D, :H, :W, :]


def cross_block(x, xa, p, pre, heads, window, dp=0.0, training=False, gen=None, closed_form_stn=False) -> Tensor:
    """CrossTransformerBlock3D.forward M:406-426."""
    x = x + _drop_path(cross_block_part1(x, xa, p, pre, heads, window, closed_form_stn), dp, training, gen)
    return x + _drop_path(mlp(layer_norm(x, p[f"{pre}.norm2.weight"], p[f"{pre}.norm2.bias"]), p, f"{pre}.mlp"),
                          dp, training, gen)


def patch_merging(x: Tensor, p, pre: str) -> Tensor:
    """M:542-561, incl. the odd-size branch M:551-555: odd D/H/W are zero-padded by one on the trailing side."""
    B, D, H, W, C = x.shape
    xc = x.permute(0, 4, 1, 2, 3)
    if (D % 2) or (H % 2) or (W % 2):
        xc = F.pad(xc, (0, W % 2, 0, H % 2, 0, D % 2))
    y = F.conv3d(xc, p[f"{pre}.down_conv.weight"], p[f"{pre}.down_conv.bias"], stride=2)
    return layer_norm(y.permute(0, 2, 3, 4, 1), p[f"{pre}.norm.weight"], p[f"{pre}.norm.bias"])


def patch_expand(x: Tensor, p, pre: str) -> Tensor:
    """M:571-579."""
    y = F.conv_transpose3d(x.permute(0, 4, 1, 2, 3), p[f"{pre}.up_conv.weight"], p[f"{pre}.up_conv.bias"], stride=2)
    return layer_norm(y.permute(0, 2, 3, 4, 1), p[f"{pre}.norm.weight"], p[f"{pre}.norm.bias"])


def basic_layer(x, xa, p, pre, depth, heads, window, sampler, dprs, training=False, gen=None,
                closed_form_stn=False):
    """BasicLayer.forward M:689-707: self blocks per stream, then BOTH cross blocks on the pre-update
    pair (simultaneous tuple assignment, M:700-701); sampler module shared by the two streams."""
    for i in range(depth):
        x = self_block(x, p, f"{pre}.self_blocks1.{i}", heads, window, dprs[i], training, gen)
        xa = self_block(xa, p, f"{pre}.self_blocks2.{i}", heads, window, dprs[i], training, gen)
        x_new = cross_block(x, xa, p, f"{pre}.blocks1.{i}", heads, window, dprs[i], training, gen, closed_form_stn)
        xa_new = cross_block(xa, x, p, f"{pre}.blocks2.{i}", heads, window, dprs[i], training, gen, closed_form_stn)
        x, xa = x_new, xa_new
    if sampler == "down":
        return x, xa, patch_merging(x, p, f"{pre}.downsample"), patch_merging(xa, p, f"{pre}.downsample")
    if sampler == "up":
        return x, xa, patch_expand(x, p, f"{pre}.downsample"), patch_expand(xa, p, f"{pre}.downsample")
    return x, xa, x, xa


def micformer_forward(moving: Tensor, fixed: Tensor, p, cfg: Config, training=False, gen=None,
                      closed_form_stn=False) -> Tensor:
    """MicFormer.forward M:992-1039.  For ``len(depths) != 4`` the reference's hard-coded
    ``features[3 - inx]`` (M:1018-1028) is read as ``features[num_layers-1-inx]``."""
    L = cfg.num_layers
    dpr = cfg.drop_path_rates()
    w, b = p["swin.patch_embed.proj.weight"], p["swin.patch_embed.proj.bias"]
    D0, H0, W0 = moving.shape[2:]
    if (D0 % 4) or (H0 % 4) or (W0 % 4):          # PatchEmbed3D pads to a multiple of the patch size on the trailing side, M:864-869
        pads = (0, (4 - W0 % 4) % 4, 0, (4 - H0 % 4) % 4, 0, (4 - D0 % 4) % 4)
        moving, fixed = F.pad(moving, pads), F.pad(fixed, pads)
    moving = F.conv3d(moving, w, b, stride=4).permute(0, 2, 3, 4, 1).contiguous()   # M:871, M:1001
    fixed = F.conv3d(fixed, w, b, stride=4).permute(0, 2, 3, 4, 1).contiguous()
    feats_m, feats_f = [], []
    for i in range(L):
        C = cfg.embed_dim * 2 ** i
        d0 = sum(cfg.depths[:i])
        mo, fo, moving, fixed = basic_layer(moving, fixed, p, f"swin.layers.{i}", cfg.depths[i], cfg.num_heads[i],
                                            cfg.window_size, "down" if i < L - 1 else None,
                                            dpr[d0:d0 + cfg.depths[i]], training, gen, closed_form_stn)
        feats_m.append(mo); feats_f.append(fo)
    moving = layer_norm(moving, p["swin.norm.weight"], p["swin.norm.bias"])
    fixed = layer_norm(fixed, p["swin.norm.weight"], p["swin.norm.bias"])
    for k, i in enumerate(reversed(range(L))):
        if k > 0:
            skip_m, skip_f = feats_m[L - 1 - k], feats_f[L - 1 - k]
            if moving.shape != skip_m.shape:      # odd sizes: trilinear resize (align_corners=True) to the skip's grid, M:1018-1025
                size = tuple(skip_m.shape[1:4])
                moving = F.interpolate(moving.permute(0, 4, 1, 2, 3), size=size, mode="trilinear", align_corners=True).permute(0, 2, 3, 4, 1)
                fixed = F.interpolate(fixed.permute(0, 4, 1, 2, 3), size=size, mode="trilinear", align_corners=True).permute(0, 2, 3, 4, 1)
            wcb, bcb = p[f"swin.concat_back_dim.{k}.weight"], p[f"swin.concat_back_dim.{k}.bias"]
            moving = F.linear(torch.cat([moving, skip_m], -1), wcb, bcb)
            fixed = F.linear(torch.cat([fixed, skip_f], -1), wcb, bcb)
        d0 = sum(cfg.depths[:i])
        _, _, moving, fixed = basic_layer(moving, fixed, p, f"swin.up_layers.{k}", cfg.depths[i], cfg.num_heads[i],
                                          cfg.window_size, "up" if i > 0 else None,
                                          dpr[d0:d0 + cfg.depths[i]], training, gen, closed_form_stn)
    x = layer_norm(torch.cat([moving, fixed], -1), p["swin.norm2.weight"], p["swin.norm2.bias"])
    return F.conv_transpose3d(x.permute(0, 4, 1, 2, 3), p["swin.reverse_patch_embedding.weight"],
                              p["swin.reverse_patch_embedding.bias"], stride=4)


def head_forward(x: Tensor, p, cfg: Config, training=False, gen=None, closed_form_stn=False) -> Tensor:
    """Head.forward M:1049-1055: channel 0 = moving (CT), channel 1 = fixed (MR)."""
    moving, fixed = torch.split(x, 1, dim=1)
    y = micformer_forward(moving, fixed, p, cfg, training, gen, closed_form_stn)
    return F.conv3d(y, p["out_conv.weight"], p["out_conv.bias"], padding=1)


def mdice_loss(logits: Tensor, target: Tensor) -> Tensor:
    """MDiceLoss.forward L:158-166 + binary_dice L:130-151: per channel, sums over batch AND space;
    BCE is torch's BCELoss on sigmoid probabilities (log clamped at -100)."""
    C = target.shape[1]
    dice = logits.new_zeros(())
    ce = logits.new_zeros(())
    for i in range(C):
        pr = torch.sigmoid(logits[:, i])
        t = target[:, i]
        inter = (pr * t).sum()
        dice = dice + (1 - (2 * inter + 1.0) / (pr.pow(2).sum() + t.pow(2).sum() + 1.0))
        ce = ce + F.binary_cross_entropy(pr, t)
    return (0.7 * dice + 0.3 * ce) / C


def mdice_sums(logits: Tensor, target: Tensor) -> Tensor:
    """Per-channel partial sums [sum p*t, sum p^2, sum t^2, sum bce] -- the quantities the fused CUDA
    loss kernel reduces (and the ones a global-batch DDP loss would all-reduce)."""
    pr = torch.sigmoid(logits)
    dims = (0, 2, 3, 4)
    bce = F.binary_cross_entropy(pr, target, reduction="none")
    return torch.stack([(pr * target).sum(dims), pr.pow(2).sum(dims), target.pow(2).sum(dims), bce.sum(dims)], 1)


def train_step(x: Tensor, target: Tensor, p: Dict[str, Tensor], cfg: Config, training=False, gen=None):
    """One forward + MDiceLoss + backward over leaf copies of ``p``.  Returns (logits, loss, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    logits = head_forward(x, leaves, cfg, training, gen)
    loss = mdice_loss(logits, target)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items()}
    return logits.detach(), loss.detach(), grads
