"""Attention modules of the SkyEye path, B200-native (parameter layout = reference
/root/reference/skyeye/core/models/attention.py; compute = native launches).

* CombinedAttention  : CBAM gate (attention.py:11-130) -> skb_cbam_bf16.
* CrossLayerAttention: reference op in its closed form (the region loop resamples
  the same map R^2 times and nn.Softmax(dim=3) normalises over image rows, X18) with repair R4
  (K projected to the query width, X4).  Projections are tcgen05 1x1 GEMMs (k||v fused), the
  softmax-gate is skb_cla_core_bf16, the residual add is fused into the output projection.
* TransformerLayer   : pre-LN encoder layer (attention.py:244-309); QKV / out / FFN projections are
  tcgen05 GEMMs with fused bias, ReLU and residual; the N x N attention is a flash-style tcgen05
  kernel (skb_flash_attn_bf16) instead of nn.MultiheadAttention's materialised weights.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _lowering as L
from ...engine import ACT_NONE, ACT_RELU, PackedConv, Plan, View


class ChannelAttention(nn.Module):
    def __init__(self, channels, reduction_ratio=16):
        super().__init__()
        r = max(channels // reduction_ratio, 1)
        self.shared_mlp = nn.Sequential(nn.Linear(channels, r, bias=False), nn.ReLU(inplace=True), nn.Linear(r, channels, bias=False))


class SpatialAttention(nn.Module):
    def __init__(self, kernel_size=7):
        super().__init__()
        if kernel_size != 7:
            raise NotImplementedError("spatial attention is lowered for the 7x7 kernel on the path")
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)


class CombinedAttention(nn.Module):
    """x * sigmoid(MLP(avg) + MLP(max)) then * sigmoid(conv7x7([mean_c, max_c]))."""

    def __init__(self, channels, reduction_ratio=16):
        super().__init__()
        self.channel_attention = ChannelAttention(channels, reduction_ratio)
        self.spatial_attention = SpatialAttention()
        self.channels = channels

    def lower(self, plan: Plan, x: View, out: View = None, name="cbam") -> View:
        dev = plan.device
        w0 = self.channel_attention.shared_mlp[0].weight.detach().float().to(dev).contiguous()
        w1 = self.channel_attention.shared_mlp[2].weight.detach().float().to(dev).contiguous()
        w7 = self.spatial_attention.conv.weight.detach().float().to(dev).contiguous()
        if out is None:
            out = plan.buf(x.n, x.h, x.w, x.c)
        ws = plan.ws(L.E.N.lib().skb_cbam_workspace_bytes(x.n, x.h, x.w, x.c))
        plan.keep += [w0, w1, w7]
        plan.add(name, lambda s: L.E.cbam(x, w0, w1, w7, out, ws, s), "cbam", 0.0, 2.0 * x.n * x.h * x.w * x.c * 4, 4,
                 outs=[dict(view=out, label=L.ref(self))])
        return out

    def forward(self, x):
        return L.run_module(self, x)


class CrossLayerAttention(nn.Module):
    def __init__(self, query_channels, key_channels, value_channels=None, region_size=2, output_channels=None, heads=4):
        super().__init__()
        value_channels = key_channels if value_channels is None else value_channels
        output_channels = query_channels if output_channels is None else output_channels
        self.scale = 1.0 / math.sqrt(query_channels)  # attention.py:159 (not 1/sqrt(head_dim))
        self.heads, self.region_size = heads, region_size
        self.query_channels, self.key_channels, self.value_channels = query_channels, key_channels, value_channels
        self.query_projection = nn.Conv2d(query_channels, query_channels, 1)
        self.key_projection = nn.Conv2d(key_channels, query_channels, 1)  # R4
        self.value_projection = nn.Conv2d(value_channels, value_channels, 1)
        self.output_projection = nn.Conv2d(value_channels, output_channels, 1)

    def lower(self, plan: Plan, query: View, key: View, out: View = None, residual: View = None, name="cla") -> View:
        dev = plan.device
        cq, cv = self.query_channels, self.value_channels
        q = plan.buf(query.n, query.h, query.w, cq)
        plan.conv(name + ".q", query, PackedConv(self.query_projection.weight, self.query_projection.bias, dev), q, 1, ACT_NONE,
                  label=L.ref(self, ".q"))
        kv = plan.buf(key.n, key.h, key.w, cq + cv)
        pk = PackedConv(self.key_projection.weight, self.key_projection.bias, dev)
        pv = PackedConv(self.value_projection.weight, self.value_projection.bias, dev)
        plan.conv(name + ".k|v", key, PackedConv.concat([pk, pv]), kv, 1, ACT_NONE,
                  label=[(0, cq, L.ref(self, ".k")), (cq, cq + cv, L.ref(self, ".v"))])
        o = plan.buf(query.n, query.h, query.w, cv)
        ws = plan.ws(L.E.N.lib().skb_cla_workspace_bytes(query.n, query.h, query.w, self.heads))
        k, v = kv.slice(0, cq), kv.slice(cq, cq + cv)
        r2 = float(self.region_size * self.region_size)
        npx = query.n * query.h * query.w
        plan.add(name + ".core", lambda s: L.E.cla_core(q, k, v, o, self.heads, self.scale, r2, ws, s), "cla", 0.0,
                 2.0 * npx * (cq + cv) + 2.0 * (npx // 4) * (cq + cv), 3, outs=[dict(view=o, label=L.ref(self, ".core"))])
        if out is None:
            out = plan.buf(query.n, query.h, query.w, self.output_projection.out_channels)
        po = PackedConv(self.output_projection.weight, self.output_projection.bias, dev)
        return plan.conv(name + ".out", o, po, out, 1, ACT_NONE, residual, label=L.ref(self, ".out"))

    def forward(self, query, key):
        if not query.is_cuda:
            raise RuntimeError("skyeye (B200) modules run on CUDA only; there is no CPU fallback")
        plan = Plan(query.device)
        out = self.lower(plan, L.E.from_nchw(query.float()), L.E.from_nchw(key.float()))
        plan.run()
        return out.nchw().float()


class TransformerLayer(nn.Module):
    def __init__(self, dim, num_heads, feedforward_dim=None, dropout=0.1):
        super().__init__()
        feedforward_dim = dim * 4 if feedforward_dim is None else feedforward_dim
        self.self_attn = nn.MultiheadAttention(dim, num_heads, dropout=dropout)  # parameter container only
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.feedforward = nn.Sequential(nn.Linear(dim, feedforward_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
                                         nn.Linear(feedforward_dim, dim), nn.Dropout(dropout))
        self.dim, self.num_heads = dim, num_heads

    def lower(self, plan: Plan, x: View, out: View = None, name="tl") -> View:
        dev, C = plan.device, self.dim
        if C // self.num_heads != 64:
            raise NotImplementedError(f"flash attention kernel is built for head_dim 64 (got {C // self.num_heads})")
        f32 = lambda t: t.detach().float().to(dev).contiguous()
        g1, b1, g2, b2 = f32(self.norm1.weight), f32(self.norm1.bias), f32(self.norm2.weight), f32(self.norm2.bias)
        plan.keep += [g1, b1, g2, b2]
        n, h, w = x.n, x.h, x.w
        xn = plan.buf(n, h, w, C)
        plan.add(name + ".ln1", lambda s: L.E.layernorm(x, g1, b1, xn, self.norm1.eps, s), "layernorm", 0.0, 4.0 * n * h * w * C,
                 outs=[dict(view=xn, label=L.ref(self, ".ln1"))])
        qkv = plan.buf(n, h, w, 3 * C)
        plan.conv(name + ".qkv", xn, PackedConv(self.self_attn.in_proj_weight, self.self_attn.in_proj_bias, dev), qkv, 1, ACT_NONE,
                  label=L.ref(self, ".qkv"))
        o = plan.buf(n, h, w, C)
        scale = 1.0 / math.sqrt(C // self.num_heads)
        ntok = h * w
        plan.add(name + ".attn", lambda s: L.E.flash_attn(qkv, o, self.num_heads, scale, s), "attention",
                 4.0 * n * float(ntok) * ntok * C, 2.0 * n * ntok * 4 * C,  # 4*N^2*C per image
                 outs=[dict(view=o, label=L.ref(self, ".attn"))])
        t = plan.buf(n, h, w, C)
        plan.conv(name + ".proj", o, PackedConv(self.self_attn.out_proj.weight, self.self_attn.out_proj.bias, dev), t, 1, ACT_NONE, x,
                  label=L.ref(self, ".proj"))
        xn2 = plan.buf(n, h, w, C)
        plan.add(name + ".ln2", lambda s: L.E.layernorm(t, g2, b2, xn2, self.norm2.eps, s), "layernorm", 0.0, 4.0 * n * h * w * C,
                 outs=[dict(view=xn2, label=L.ref(self, ".ln2"))])
        ff0, ff3 = self.feedforward[0], self.feedforward[3]
        hid = plan.buf(n, h, w, ff0.out_features)
        plan.conv(name + ".ff0", xn2, PackedConv(ff0.weight, ff0.bias, dev), hid, 1, ACT_RELU, label=L.ref(self, ".ff0"))
        if out is None:
            out = plan.buf(n, h, w, C)
        return plan.conv(name + ".ff3", hid, PackedConv(ff3.weight, ff3.bias, dev), out, 1, ACT_NONE, t, label=L.ref(self, ".ff3"))

    def forward(self, x):
        return L.run_module(self, x)


class WindowedSelfAttention(nn.Module):
    """Window attention with relative position bias (reference attention.py:312-399; the class is defined but
    never instantiated there).  Same parameters and buffers (qkv, proj, relative_position_bias_table,
    relative_position_index); forward(x [B*nW, w*w, C], mask [nW, w*w, w*w] | None) -> [B*nW, w*w, C].
    qkv / proj are tcgen05 GEMMs, the per-window softmax(q k^T + bias + mask) v is skb_window_attn_bf16."""

    def __init__(self, dim, window_size, num_heads):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        ys, xs = torch.meshgrid(torch.arange(window_size), torch.arange(window_size), indexing="ij")
        coords = torch.stack((ys, xs)).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += window_size - 1
        rel[:, :, 1] += window_size - 1
        rel[:, :, 0] *= 2 * window_size - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def lower(self, plan: Plan, x: View, mask: torch.Tensor = None, out: View = None, name="wsa") -> View:
        """x: view [B*nW, 1, w*w, C]."""
        dev, C, n_tok = plan.device, self.dim, self.window_size ** 2
        assert x.h == 1 and x.w == n_tok and x.c == C, (x.h, x.w, x.c)
        bias = self.relative_position_bias_table.detach().float()[self.relative_position_index.view(-1)]
        bias = bias.view(n_tok, n_tok, self.num_heads).permute(2, 0, 1).contiguous().to(dev)
        m = None if mask is None else mask.detach().float().contiguous().to(dev)
        plan.keep += [bias, m]
        qkv = plan.buf(x.n, 1, n_tok, 3 * C)
        plan.conv(name + ".qkv", x, PackedConv(self.qkv.weight, self.qkv.bias, dev), qkv, 1, ACT_NONE)
        o = plan.buf(x.n, 1, n_tok, C)
        plan.add(name + ".attn", lambda s: L.E.window_attn(qkv, bias, m, o, self.num_heads, self.scale, s), "attention",
                 4.0 * x.n * n_tok * n_tok * C, 2.0 * x.n * n_tok * 4 * C)
        if out is None:
            out = plan.buf(x.n, 1, n_tok, C)
        return plan.conv(name + ".proj", o, PackedConv(self.proj.weight, self.proj.bias, dev), out, 1, ACT_NONE)

    def forward(self, x, mask=None):
        if not x.is_cuda:
            raise RuntimeError("skyeye (B200) modules run on CUDA only; there is no CPU fallback")
        plan = Plan(x.device)
        xin = L.E.View(x.float().to(torch.bfloat16).unsqueeze(1).contiguous())
        out = self.lower(plan, xin, mask)
        plan.run()
        return out.torch().squeeze(1).float()


def window_partition(x: torch.Tensor, window: int) -> torch.Tensor:
    """[B, H, W, C] -> [B * nW, window * window, C], windows row-major over the window grid (the input layout
    WindowedSelfAttention.forward documents, reference attention.py:358-365; the reference ships no partition helper)."""
    B, H, W, C = x.shape
    x = x.view(B, H // window, window, W // window, window, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, window * window, C)


def window_reverse(windows: torch.Tensor, window: int, H: int, W: int) -> torch.Tensor:
    """Inverse of ``window_partition``: [B * nW, window * window, C] -> [B, H, W, C]."""
    B = windows.shape[0] // ((H // window) * (W // window))
    x = windows.view(B, H // window, W // window, window, window, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


class WindowedTransformerLayer(nn.Module):
    """``head: windowed`` (NOT IN REFERENCE): TransformerLayer (attention.py:244-309) with its global
    nn.MultiheadAttention replaced by the reference's WindowedSelfAttention class (attention.py:312-399) over non-overlapping
    ``window_size`` x ``window_size`` windows -- O(N w^2) instead of O(N^2):
        t = t + reverse(attn(partition(norm1(t))));  t = t + feedforward(norm2(t)).
    Partition and reverse never run: qkv / proj are per-token GEMMs on the whole map and the attention kernel addresses
    the windows in place (skb_window_attn2d_bf16)."""

    def __init__(self, dim, num_heads, window_size=8, feedforward_dim=None, dropout=0.1):
        super().__init__()
        feedforward_dim = dim * 4 if feedforward_dim is None else feedforward_dim
        self.attn = WindowedSelfAttention(dim, window_size, num_heads)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.feedforward = nn.Sequential(nn.Linear(dim, feedforward_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
                                         nn.Linear(feedforward_dim, dim), nn.Dropout(dropout))
        self.dim, self.num_heads, self.window_size = dim, num_heads, window_size

    def lower(self, plan: Plan, x: View, out: View = None, name="wtl") -> View:
        dev, C, ws, a = plan.device, self.dim, self.window_size, self.attn
        n, h, w = x.n, x.h, x.w
        if h % ws or w % ws:
            raise ValueError(f"windowed head: the {h}x{w} level map is not a multiple of window_size {ws}")
        f32 = lambda t: t.detach().float().to(dev).contiguous()
        g1, b1, g2, b2 = f32(self.norm1.weight), f32(self.norm1.bias), f32(self.norm2.weight), f32(self.norm2.bias)
        n_tok = ws * ws
        bias = a.relative_position_bias_table.detach().float()[a.relative_position_index.view(-1)]
        bias = bias.view(n_tok, n_tok, self.num_heads).permute(2, 0, 1).contiguous().to(dev)
        plan.keep += [g1, b1, g2, b2, bias]
        xn = plan.buf(n, h, w, C)
        plan.add(name + ".ln1", lambda s: L.E.layernorm(x, g1, b1, xn, self.norm1.eps, s), "layernorm", 0.0, 4.0 * n * h * w * C,
                 outs=[dict(view=xn, label=L.ref(self, ".ln1"))])
        qkv = plan.buf(n, h, w, 3 * C)
        plan.conv(name + ".qkv", xn, PackedConv(a.qkv.weight, a.qkv.bias, dev), qkv, 1, ACT_NONE, label=L.ref(self, ".qkv"))
        o = plan.buf(n, h, w, C)
        plan.add(name + ".attn", lambda s: L.E.window_attn2d(qkv, bias, None, o, self.num_heads, ws, a.scale, s), "attention",
                 4.0 * n * h * w * n_tok * C, 2.0 * n * h * w * 4 * C, outs=[dict(view=o, label=L.ref(self, ".attn"))])
        t = plan.buf(n, h, w, C)
        plan.conv(name + ".proj", o, PackedConv(a.proj.weight, a.proj.bias, dev), t, 1, ACT_NONE, x, label=L.ref(self, ".proj"))
        xn2 = plan.buf(n, h, w, C)
        plan.add(name + ".ln2", lambda s: L.E.layernorm(t, g2, b2, xn2, self.norm2.eps, s), "layernorm", 0.0, 4.0 * n * h * w * C,
                 outs=[dict(view=xn2, label=L.ref(self, ".ln2"))])
        ff0, ff3 = self.feedforward[0], self.feedforward[3]
        hid = plan.buf(n, h, w, ff0.out_features)
        plan.conv(name + ".ff0", xn2, PackedConv(ff0.weight, ff0.bias, dev), hid, 1, ACT_RELU, label=L.ref(self, ".ff0"))
        if out is None:
            out = plan.buf(n, h, w, C)
        return plan.conv(name + ".ff3", hid, PackedConv(ff3.weight, ff3.bias, dev), out, 1, ACT_NONE, t, label=L.ref(self, ".ff3"))

    def forward(self, x):
        return L.run_module(self, x)

