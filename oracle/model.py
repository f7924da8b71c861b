"""CPU oracle (TEST INFRASTRUCTURE): fp32 PyTorch restatement of the SkyEye forward path.

Functional, stateless: every function takes a reference-keyed ``state_dict`` (the same key
names the reference's ``nn.Module`` tree produces, so weights interchange with the reference
through ``load_state_dict``) and follows the reference arithmetic line by line.  Citations
are into ``/root/reference``.  Pinned against the reference itself by
``tests/golden/make_golden.py`` (run in the build container) -> ``tests/golden/*.npz``.

Two arithmetic modes:
  * ``emu=None``   : pure fp32, the reference arithmetic (what the reference computes on CPU).
  * ``emu='bf16'`` : same dataflow, but BN is folded into the conv weights, weights/activations
    are rounded to bf16 at exactly the points where the B200 pipeline stores bf16 (fp32
    accumulation everywhere).  This is the "fp32-accumulate" comparison target: differences
    to the CUDA path are accumulation order and intrinsic error only.
"""
from __future__ import annotations

import math
import zlib
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# detector.py:39-43 default anchors (pixel units; multiplied by stride AGAIN in decode, X16)
DEFAULT_ANCHORS = [
    [[10, 13], [16, 30], [33, 23]],
    [[30, 61], [62, 45], [59, 119]],
    [[116, 90], [156, 198], [373, 326]],
]

# SURVEY.md §0.3 D1 (reference configs/models/*.yaml are empty, X6): variants through base_channels.
VARIANTS = {
    "skyeye_s": dict(base_channels=32, depth_multiple=0.33, nc=10, enhanced=False, head_dim=64),
    "skyeye_m": dict(base_channels=48, depth_multiple=0.67, nc=10, enhanced=False, head_dim=64),
    "skyeye_l": dict(base_channels=64, depth_multiple=1.0, nc=10, enhanced=True, head_dim=64),
    # small variants for fast CPU tests / committed golden vectors (same code paths)
    "skyeye_tiny": dict(base_channels=8, depth_multiple=0.33, nc=10, enhanced=False, head_dim=16),
    "skyeye_tiny_l": dict(base_channels=8, depth_multiple=0.33, nc=10, enhanced=True, head_dim=16),
    # smallest variant whose every channel count is a multiple of 32 (tcgen05 path test size)
    "skyeye_nano_l": dict(base_channels=32, depth_multiple=0.33, nc=10, enhanced=True, head_dim=64),
    # head: windowed (NOT IN REFERENCE): the global attention of the D4 heads replaced by the reference's
    # WindowedSelfAttention class (attention.py:312-399) over window_size x window_size windows
    "skyeye_nano_lw": dict(base_channels=32, depth_multiple=0.33, nc=10, enhanced=True, head_dim=64, head="windowed", window_size=8),
    "skyeye_lw": dict(base_channels=64, depth_multiple=1.0, nc=10, enhanced=True, head_dim=64, head="windowed", window_size=8),
}


def get_cfg(variant) -> dict:
    cfg = dict(VARIANTS[variant]) if isinstance(variant, str) else dict(variant)
    cfg.setdefault("nc", 10)
    cfg.setdefault("enhanced", False)
    cfg.setdefault("head_dim", 64)
    cfg.setdefault("anchors", None)
    cfg.setdefault("head", "transformer" if cfg["enhanced"] else "conv")  # D4: the enhanced variants carry transformer heads
    cfg.setdefault("window_size", 8)
    return cfg


def channels(cfg) -> Tuple[int, int, int, int, int]:
    b = cfg["base_channels"]  # backbone.py:38-42 with width_multiple == 1 (R3)
    return b, 2 * b, 4 * b, 8 * b, 16 * b


def depths(cfg) -> Tuple[int, int]:
    d = cfg["depth_multiple"]  # backbone.py:34-35
    return max(round(3 * d), 1), max(round(9 * d), 1)


# ----------------------------------------------------------------------------------------------
# state-dict spec (reference key names) and deterministic init
# ----------------------------------------------------------------------------------------------

def _convblock(spec, p, cin, cout, k, rv):
    spec.append((p + ".conv.weight", (cout, cin, k, k), ("conv", k * k * cout)))
    spec.append((p + ".bn.weight", (cout,), ("const", 1.0)))
    spec.append((p + ".bn.bias", (cout,), ("const", 0.0)))
    spec.append((p + ".bn.running_mean", (cout,), ("const", 0.0)))
    spec.append((p + ".bn.running_var", (cout,), ("const", rv)))
    spec.append((p + ".bn.num_batches_tracked", (), ("int", 1 if rv != 1.0 else 0)))


def _csp(spec, p, cin, cout, n, rv):
    h = int(cout * 0.5)  # blocks.py:109
    _convblock(spec, p + ".cv1", cin, h, 1, rv)
    _convblock(spec, p + ".cv2", cin, h, 1, rv)
    _convblock(spec, p + ".cv3", 2 * h, cout, 1, rv)
    for i in range(n):  # blocks.py:115-118, expansion 1.0
        _convblock(spec, f"{p}.bottlenecks.{i}.cv1", h, h, 1, rv)
        _convblock(spec, f"{p}.bottlenecks.{i}.cv2", h, h, 3, rv)


def state_spec(cfg) -> List[Tuple[str, tuple, tuple]]:
    """Ordered (key, shape, init) list; key names/order == reference ``state_dict()``."""
    cfg = get_cfg(cfg)
    c1, c2, c3, c4, c5 = channels(cfg)
    d3, d9 = depths(cfg)
    nc = cfg["nc"]
    no = nc + 5
    spec: list = []
    bb = "backbone.backbone."
    rv = 0.9  # X17: constructor runs the backbone once in train mode on zeros
    _convblock(spec, bb + "stage1.0.conv", 12, c1, 3, rv)       # FocusBlock (blocks.py:167)
    _convblock(spec, bb + "stage1.1", c1, c2, 3, rv)
    _csp(spec, bb + "stage1.2", c2, c2, d3, rv)
    _convblock(spec, bb + "stage2.0", c2, c3, 3, rv)
    _csp(spec, bb + "stage2.1", c3, c3, d9, rv)
    _convblock(spec, bb + "stage3.0", c3, c4, 3, rv)
    _csp(spec, bb + "stage3.1", c4, c4, d9, rv)
    r = max(c4 // 16, 1)  # attention.py:29
    spec.append((bb + "stage3.2.channel_attention.shared_mlp.0.weight", (r, c4), ("normal", 0.01)))
    spec.append((bb + "stage3.2.channel_attention.shared_mlp.2.weight", (c4, r), ("normal", 0.01)))
    spec.append((bb + "stage3.2.spatial_attention.conv.weight", (1, 2, 7, 7), ("conv", 49)))
    _convblock(spec, bb + "stage4.0", c4, c5, 3, rv)
    _csp(spec, bb + "stage4.1", c5, c5, d3, rv)
    _convblock(spec, bb + "stage4.2.cv1", c5, c5 // 2, 1, rv)    # SPPBlock (blocks.py:139-141)
    _convblock(spec, bb + "stage4.2.cv2", (c5 // 2) * 4, c5, 1, rv)
    # FeatureNeck (detector.py:170-188); in_channels = (c3, c4, c5) after R2
    _convblock(spec, "neck.lateral_conv5", c5, c4, 1, 1.0)
    _convblock(spec, "neck.lateral_conv4", c4, c3, 1, 1.0)
    _csp(spec, "neck.fpn_conv4", 2 * c4, c4, 3, 1.0)
    _csp(spec, "neck.fpn_conv3", 2 * c3, c3, 3, 1.0)
    _convblock(spec, "neck.downsample3", c3, c3, 3, 1.0)
    _convblock(spec, "neck.downsample4", c4, c4, 3, 1.0)
    _csp(spec, "neck.pan_conv4", c3 + c4, c4, 3, 1.0)
    _csp(spec, "neck.pan_conv5", c4 + c5, c5, 3, 1.0)
    for i, c in enumerate((c3, c4, c5)):  # detector.py:56-59
        spec.append((f"detection_head.detection_layers.{i}.weight", (3 * no, c, 1, 1), ("normal", 1.0 / math.sqrt(c))))
        spec.append((f"detection_head.detection_layers.{i}.bias", (3 * no,), ("normal", 0.5)))
    if cfg["enhanced"]:
        for name, cq, ck in (("cross_attention_p5_p4", c4, c5), ("cross_attention_p4_p3", c3, c4)):
            for proj, co, ci in (("query_projection", cq, cq), ("key_projection", cq, ck),  # R4
                                 ("value_projection", ck, ck), ("output_projection", cq, ck)):
                spec.append((f"{name}.{proj}.weight", (co, ci, 1, 1), ("normal", 1.0 / math.sqrt(ci))))
                spec.append((f"{name}.{proj}.bias", (co,), ("normal", 0.1)))
        for i, c in enumerate((c3, c4, c5) if cfg["head"] != "conv" else ()):  # D4: TransformerLayer per level (attention.py:244-309)
            p = f"head_transformers.{i}"
            if cfg["head"] == "windowed":  # WindowedSelfAttention parameters + buffer (attention.py:333-353) in registration order
                ws, nh = cfg["window_size"], max(c // cfg["head_dim"], 1)
                spec.append((p + ".attn.relative_position_bias_table", ((2 * ws - 1) ** 2, nh), ("normal", 0.02)))
                spec.append((p + ".attn.relative_position_index", (ws * ws, ws * ws), ("relidx", ws)))
                spec.append((p + ".attn.qkv.weight", (3 * c, c), ("normal", 1.0 / math.sqrt(c))))
                spec.append((p + ".attn.qkv.bias", (3 * c,), ("normal", 0.02)))
                spec.append((p + ".attn.proj.weight", (c, c), ("normal", 1.0 / math.sqrt(c))))
                spec.append((p + ".attn.proj.bias", (c,), ("normal", 0.02)))
            else:
                spec.append((p + ".self_attn.in_proj_weight", (3 * c, c), ("normal", 1.0 / math.sqrt(c))))
                spec.append((p + ".self_attn.in_proj_bias", (3 * c,), ("normal", 0.02)))
                spec.append((p + ".self_attn.out_proj.weight", (c, c), ("normal", 1.0 / math.sqrt(c))))
                spec.append((p + ".self_attn.out_proj.bias", (c,), ("normal", 0.02)))
            spec.append((p + ".norm1.weight", (c,), ("normal1", 0.1)))
            spec.append((p + ".norm1.bias", (c,), ("normal", 0.1)))
            spec.append((p + ".norm2.weight", (c,), ("normal1", 0.1)))
            spec.append((p + ".norm2.bias", (c,), ("normal", 0.1)))
            spec.append((p + ".feedforward.0.weight", (4 * c, c), ("normal", 1.0 / math.sqrt(c))))
            spec.append((p + ".feedforward.0.bias", (4 * c,), ("normal", 0.02)))
            spec.append((p + ".feedforward.3.weight", (c, 4 * c), ("normal", 1.0 / math.sqrt(4 * c))))
            spec.append((p + ".feedforward.3.bias", (c,), ("normal", 0.02)))
    return spec


def make_state_dict(cfg, seed: int = 0, trained_like: bool = True) -> SD:
    """Deterministic random weights with the reference's key names and shapes.

    Distribution follows ``_initialize_weights`` (detector.py:326-341): conv ~ N(0, sqrt(2/n)),
    Linear ~ N(0, .01); with ``trained_like`` BN affine/running stats and biases are perturbed
    (a freshly-initialised reference has gamma=1, beta=0, mean=0: that would leave the BN-fold
    path untested).  Each tensor draws from its own PCG64 stream keyed by (seed, crc32(key)),
    so the dict is identical on every machine and independent of construction order.
    """
    sd: SD = {}
    for key, shape, (kind, arg) in state_spec(cfg):
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))
        if kind == "int":
            sd[key] = torch.tensor(int(arg), dtype=torch.long)
            continue
        if kind == "relidx":  # registered buffer of WindowedSelfAttention (attention.py:340-352)
            sd[key] = relative_position_index(int(arg))
            continue
        if kind == "conv":
            a = rng.standard_normal(shape, dtype=np.float32) * math.sqrt(2.0 / arg)
        elif kind == "normal":
            a = rng.standard_normal(shape, dtype=np.float32) * arg
        elif kind == "normal1":
            a = 1.0 + rng.standard_normal(shape, dtype=np.float32) * arg
        elif kind == "const":
            a = np.full(shape, arg, dtype=np.float32)
            if trained_like:
                if key.endswith("bn.weight"):
                    a = (1.0 + 0.2 * rng.standard_normal(shape)).astype(np.float32)
                elif key.endswith("bn.bias") or key.endswith("running_mean"):
                    a = (0.1 * rng.standard_normal(shape)).astype(np.float32)
                elif key.endswith("running_var"):
                    a = (arg * np.exp(0.2 * rng.standard_normal(shape))).astype(np.float32)
        else:
            raise ValueError(kind)
        sd[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


# ----------------------------------------------------------------------------------------------
# arithmetic
# ----------------------------------------------------------------------------------------------
BN_EPS = 1e-5  # nn.BatchNorm2d default, blocks.py:32


def bf16_round(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class Ctx:
    """Arithmetic mode. ``q`` rounds a tensor at an HBM storage point of the B200 pipeline.

    ``taps`` (a dict) collects every stored intermediate as an NCHW fp32 tensor keyed by the
    reference state-dict path of the module that produced it (``"neck.fpn_conv4.cv3"``; sub-results
    of composite modules get a suffix: ``".mp9"``, ``".q"``, ``".core"``, ``".ln1"`` ...).  The
    teacher-forced parity test feeds these to the CUDA launches one by one.
    ``calib`` makes every ConvolutionBlock overwrite its BN running statistics in ``sd`` with the
    batch statistics of the tensor it sees (see ``calibrate_bn``)."""

    def __init__(self, emu: Optional[str] = None, taps: Optional[dict] = None, calib: bool = False):
        assert emu in (None, "bf16")
        self.emu = emu
        self.taps = taps
        self.calib = calib

    def q(self, t: Tensor) -> Tensor:
        return bf16_round(t) if self.emu == "bf16" else t

    def tap(self, name: str, t: Tensor) -> Tensor:
        if self.taps is not None:
            self.taps[name] = t
        return t


FP32 = Ctx(None)


def fold_bn(sd: SD, p: str) -> Tuple[Tensor, Tensor]:
    """w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps) (SURVEY.md §5 checkpoint row)."""
    w = sd[p + ".conv.weight"]
    s = sd[p + ".bn.weight"] / torch.sqrt(sd[p + ".bn.running_var"] + BN_EPS)
    return w * s.view(-1, 1, 1, 1), sd[p + ".bn.bias"] - sd[p + ".bn.running_mean"] * s


def conv_block(x: Tensor, sd: SD, p: str, stride: int = 1, ctx: Ctx = FP32,
               residual: Optional[Tensor] = None) -> Tensor:
    """ConvolutionBlock.forward (blocks.py:36-38): SiLU(BN(conv(x))), conv bias=False, pad k//2.
    ``residual`` is BottleneckBlock's shortcut, added after the activation (blocks.py:90)."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    if ctx.emu is None:
        y = F.conv2d(x, w, None, stride, k // 2)
        if ctx.calib:  # BN statistics of THIS input become the running statistics (calibrate_bn)
            sd[p + ".bn.running_mean"] = y.mean((0, 2, 3))
            sd[p + ".bn.running_var"] = y.var((0, 2, 3), unbiased=False).clamp_min(1e-6)
        y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"],
                         sd[p + ".bn.weight"], sd[p + ".bn.bias"], False, 0.0, BN_EPS)
        y = F.silu(y)
        return ctx.tap(p, y if residual is None else residual + y)
    wf, bf = fold_bn(sd, p)
    y = F.conv2d(x, bf16_round(wf), bf, stride, k // 2)
    y = F.silu(y)
    if residual is not None:
        y = residual + y
    return ctx.tap(p, ctx.q(y))


def bottleneck(x, sd, p, ctx=FP32):
    """BottleneckBlock.forward (blocks.py:88-90), shortcut always active inside CSP."""
    return conv_block(conv_block(x, sd, p + ".cv1", 1, ctx), sd, p + ".cv2", 1, ctx, residual=x)


def csp(x, sd, p, n, ctx=FP32):
    """CSPBlock.forward (blocks.py:119-123): cv3(cat(bottlenecks(cv1(x)), cv2(x)))."""
    y1 = conv_block(x, sd, p + ".cv1", 1, ctx)
    for i in range(n):
        y1 = bottleneck(y1, sd, f"{p}.bottlenecks.{i}", ctx)
    y2 = conv_block(x, sd, p + ".cv2", 1, ctx)
    return conv_block(torch.cat((y1, y2), 1), sd, p + ".cv3", 1, ctx)


def spp(x, sd, p, ctx=FP32):
    """SPPBlock.forward (blocks.py:146-149), kernel sizes (5, 9, 13), stride 1, pad k//2."""
    x = conv_block(x, sd, p + ".cv1", 1, ctx)
    pools = [ctx.tap(f"{p}.mp{k}", F.max_pool2d(x, k, 1, k // 2)) for k in (5, 9, 13)]
    return conv_block(torch.cat([x] + pools, 1), sd, p + ".cv2", 1, ctx)


def focus(x, sd, p, ctx=FP32):
    """FocusBlock.forward (blocks.py:170-182): channel = patch*3 + c, patches TL, BL, TR, BR."""
    patches = [x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]]
    return conv_block(ctx.q(torch.cat(patches, 1)), sd, p + ".conv", 1, ctx)


def cbam(x, sd, p, ctx=FP32):
    """CombinedAttention (attention.py:101-130) = ChannelAttention (:37-60) then SpatialAttention (:80-98)."""
    b, c = x.shape[:2]
    w0 = sd[p + ".channel_attention.shared_mlp.0.weight"]
    w1 = sd[p + ".channel_attention.shared_mlp.2.weight"]
    mlp = lambda v: F.linear(F.relu(F.linear(v, w0)), w1)
    att = torch.sigmoid(mlp(x.mean((2, 3))) + mlp(x.amax((2, 3)))).view(b, c, 1, 1)
    x = x * att  # kept in fp32 registers in the fused B200 kernel: no storage point here
    m = torch.cat([x.mean(1, keepdim=True), x.amax(1, keepdim=True)], 1)
    sa = torch.sigmoid(F.conv2d(m, sd[p + ".spatial_attention.conv.weight"], None, 1, 3))
    return ctx.tap(p, ctx.q(x * sa))


def backbone(x, sd, cfg, ctx=FP32) -> List[Tensor]:
    """Backbone.forward (backbone.py:82-99) -> [s2 (stride 8), s3 (16), s4 (32)]."""
    d3, d9 = depths(cfg)
    p = "backbone.backbone."
    x = focus(x, sd, p + "stage1.0", ctx)
    x = conv_block(x, sd, p + "stage1.1", 2, ctx)
    s1 = csp(x, sd, p + "stage1.2", d3, ctx)
    x = conv_block(s1, sd, p + "stage2.0", 2, ctx)
    s2 = csp(x, sd, p + "stage2.1", d9, ctx)
    x = conv_block(s2, sd, p + "stage3.0", 2, ctx)
    x = csp(x, sd, p + "stage3.1", d9, ctx)
    s3 = cbam(x, sd, p + "stage3.2", ctx)
    x = conv_block(s3, sd, p + "stage4.0", 2, ctx)
    x = csp(x, sd, p + "stage4.1", d3, ctx)
    s4 = spp(x, sd, p + "stage4.2", ctx)
    return [s2, s3, s4]


def neck(feats, sd, ctx=FP32) -> List[Tensor]:
    """FeatureNeck.forward (detector.py:197-231). Note p4_td comes from the RAW backbone P4 (:211)
    and the concat orders [upsampled, skip] (:215,219) / [downsampled, lateral] (:224,228)."""
    p3, p4, p5 = feats
    p5_td = conv_block(p5, sd, "neck.lateral_conv5", 1, ctx)
    p4_td = conv_block(p4, sd, "neck.lateral_conv4", 1, ctx)
    up5 = F.interpolate(p5_td, size=p4.shape[2:], mode="nearest")
    p4_proc = csp(torch.cat([up5, p4], 1), sd, "neck.fpn_conv4", 3, ctx)
    up4 = F.interpolate(p4_td, size=p3.shape[2:], mode="nearest")
    p3_proc = csp(torch.cat([up4, p3], 1), sd, "neck.fpn_conv3", 3, ctx)
    d3 = conv_block(p3_proc, sd, "neck.downsample3", 2, ctx)
    p4_out = csp(torch.cat([d3, p4_proc], 1), sd, "neck.pan_conv4", 3, ctx)
    d4 = conv_block(p4_out, sd, "neck.downsample4", 2, ctx)
    p5_out = csp(torch.cat([d4, p5], 1), sd, "neck.pan_conv5", 3, ctx)
    return [p3_proc, p4_out, p5_out]


def _conv1x1(x, sd, p, ctx):
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    if ctx.emu:
        w = bf16_round(w)
    return F.conv2d(x, w, b)


def cla(query, key, sd, p, heads=4, region=2, ctx=FP32) -> Tensor:
    """CrossLayerAttention.forward (attention.py:174-241) with R4, in the closed form of
    SURVEY.md §8 A10: the region loop ignores (i, j) (X18) so all R^2 patches are the same
    bilinear resample, and nn.Softmax(dim=3) on [B,h,R^2,H,W] normalises over IMAGE ROWS H:
        s[b,g,y,x] = scale * sum_{c in head g} Q*K ;  a = softmax_y(s) ;  O = R^2 * a * V.
    scale = 1/sqrt(query_channels) (attention.py:159)."""
    B, cq, H, W = query.shape
    q = ctx.tap(p + ".q", ctx.q(_conv1x1(query, sd, p + ".query_projection", ctx)))
    k = ctx.tap(p + ".k", ctx.q(_conv1x1(key, sd, p + ".key_projection", ctx)))
    v = ctx.tap(p + ".v", ctx.q(_conv1x1(key, sd, p + ".value_projection", ctx)))
    ku = F.interpolate(k, size=(H, W), mode="bilinear", align_corners=False)
    vu = F.interpolate(v, size=(H, W), mode="bilinear", align_corners=False)
    cv = vu.shape[1]
    s = (q.view(B, heads, cq // heads, H, W) * ku.view(B, heads, cq // heads, H, W)).sum(2)
    a = torch.softmax(s * (1.0 / math.sqrt(cq)), dim=2)  # over H
    o = (float(region * region) * a).unsqueeze(2) * vu.view(B, heads, cv // heads, H, W)
    o = ctx.tap(p + ".core", ctx.q(o.reshape(B, cv, H, W)))
    return _conv1x1(o, sd, p + ".output_projection", ctx)


def _attention_chunked(q, k, v, scale, chunk=2048):
    """softmax(q k^T * scale) v without materialising N x N (the reference does, attention.py:298
    need_weights=True; values are identical up to fp32 rounding)."""
    outs = []
    for i in range(0, q.shape[-2], chunk):
        a = torch.softmax((q[..., i:i + chunk, :] @ k.transpose(-1, -2)) * scale, dim=-1)
        outs.append(a @ v)
    return torch.cat(outs, dim=-2)


def transformer_layer(x, sd, p, heads, ctx=FP32) -> Tensor:
    """TransformerLayer.forward (attention.py:282-309), eval mode (dropout inert):
    pre-LN -> nn.MultiheadAttention (packed in_proj, scale 1/sqrt(head_dim)) -> residual ->
    pre-LN -> Linear(C,4C)-ReLU-Linear(4C,C) -> residual. Tokens = flatten(2) (row-major H, W)."""
    B, C, H, W = x.shape
    N = H * W
    hd = C // heads
    wr = (lambda t: bf16_round(t)) if ctx.emu else (lambda t: t)
    tap = lambda sfx, z: ctx.tap(p + sfx, z.transpose(1, 2).reshape(B, z.shape[-1], H, W)) is None or z  # tokens -> NCHW
    t = x.flatten(2).transpose(1, 2)  # [B, N, C]
    xn = tap(".ln1", ctx.q(F.layer_norm(t, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)))
    qkv = tap(".qkv", ctx.q(F.linear(xn, wr(sd[p + ".self_attn.in_proj_weight"]), sd[p + ".self_attn.in_proj_bias"])))
    q, k, v = (z.reshape(B, N, heads, hd).transpose(1, 2) for z in qkv.chunk(3, dim=-1))
    o = _attention_chunked(q, k, v, 1.0 / math.sqrt(hd))
    o = tap(".attn", ctx.q(o.transpose(1, 2).reshape(B, N, C)))
    t = tap(".proj", ctx.q(t + F.linear(o, wr(sd[p + ".self_attn.out_proj.weight"]), sd[p + ".self_attn.out_proj.bias"])))
    xn = tap(".ln2", ctx.q(F.layer_norm(t, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)))
    h = tap(".ff0", ctx.q(F.relu(F.linear(xn, wr(sd[p + ".feedforward.0.weight"]), sd[p + ".feedforward.0.bias"]))))
    t = tap(".ff3", ctx.q(t + F.linear(h, wr(sd[p + ".feedforward.3.weight"]), sd[p + ".feedforward.3.bias"])))
    return t.transpose(1, 2).reshape(B, C, H, W)


def head(feats, sd, nc, ctx=FP32) -> List[Tensor]:
    """DetectionHead.forward (detector.py:61-86): 1x1 conv(+bias) -> [B, na, H, W, no]."""
    no, na = nc + 5, 3
    outs = []
    for i, f in enumerate(feats):
        y = ctx.tap(f"detection_head.detection_layers.{i}", _conv1x1(f, sd, f"detection_head.detection_layers.{i}", ctx))
        b, _, h, w = y.shape
        outs.append(y.view(b, na, no, h, w).permute(0, 1, 3, 4, 2).contiguous())
    return outs


def decode(raws: Sequence[Tensor], input_hw, anchors=None) -> Tensor:
    """DetectionHead.process_detections (detector.py:88-145). stride = max(H/h, W/w) as a float
    (:107-109); grid order (x, y) (:115); anchors are multiplied by stride again (X16, :121)."""
    anchors = DEFAULT_ANCHORS if anchors is None else anchors
    out = []
    for i, raw in enumerate(raws):
        b, na, gh, gw, no = raw.shape
        stride = max(input_hw[0] / gh, input_hw[1] / gw)
        yv, xv = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
        grid = torch.stack((xv, yv), 2).view(1, 1, gh, gw, 2).float()
        ag = torch.tensor(anchors[i], dtype=torch.float32).view(1, na, 1, 1, 2) * stride
        y = torch.sigmoid(raw.float())
        xy = (y[..., 0:2] * 2 - 0.5 + grid) * torch.tensor(stride)
        wh = (y[..., 2:4] * 2) ** 2 * ag
        out.append(torch.cat((xy, wh, y[..., 4:]), -1).view(b, -1, no))
    return torch.cat(out, 1)


def relative_position_index(window: int) -> Tensor:
    """Pair-wise relative position index of a window (attention.py:340-350): [w*w, w*w] into the bias table."""
    ys, xs = torch.meshgrid(torch.arange(window), torch.arange(window), indexing="ij")
    coords = torch.stack((ys, xs)).flatten(1)                      # [2, w*w]
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window - 1
    rel[:, :, 1] += window - 1
    rel[:, :, 0] *= 2 * window - 1
    return rel.sum(-1)


def windowed_self_attention(x, sd, p, window, heads, mask=None, ctx=FP32) -> Tensor:
    """WindowedSelfAttention.forward (attention.py:358-399): x [B*nW, w*w, C] -> same shape.
    qkv Linear -> per head softmax(q*scale @ k^T + bias[rel_idx] (+ mask[window])) @ v -> proj Linear."""
    B_, N, C = x.shape
    hd = C // heads
    wq, wp = sd[p + ".qkv.weight"], sd[p + ".proj.weight"]
    if ctx.emu:
        wq, wp = bf16_round(wq), bf16_round(wp)
    qkv = ctx.q(F.linear(ctx.q(x), wq, sd[p + ".qkv.bias"])).reshape(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + ".relative_position_bias_table"][relative_position_index(window).view(-1)].view(N, N, heads)
    attn = attn + bias.permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    o = ctx.q((torch.softmax(attn, dim=-1) @ v).transpose(1, 2).reshape(B_, N, C))
    return ctx.q(F.linear(o, wp, sd[p + ".proj.bias"]))


def window_partition(t: Tensor, window: int) -> Tensor:
    """[B, H, W, C] -> [B * nW, window * window, C]: the input layout WindowedSelfAttention.forward documents
    (attention.py:358-365); windows row-major over the window grid, tokens row-major inside a window.  NOT IN REFERENCE
    (the class is never wired): the standard Swin partition."""
    B, H, W, C = t.shape
    t = t.view(B, H // window, window, W // window, window, C)
    return t.permute(0, 1, 3, 2, 4, 5).reshape(-1, window * window, C)


def window_reverse(wins: Tensor, window: int, H: int, W: int) -> Tensor:
    """Inverse of ``window_partition``: [B * nW, window * window, C] -> [B, H, W, C]."""
    B = wins.shape[0] // ((H // window) * (W // window))
    t = wins.view(B, H // window, W // window, window, window, -1)
    return t.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def windowed_transformer_layer(x, sd, p, heads, window, ctx=FP32) -> Tensor:
    """``head: windowed`` (NOT IN REFERENCE): TransformerLayer.forward (attention.py:282-309) with
    ``self_attn`` replaced by WindowedSelfAttention (attention.py:358-399) over non-overlapping windows:
    t + reverse(attn(partition(norm1(t)))), then t + feedforward(norm2(t)).  Stored points as in ``transformer_layer``."""
    B, C, H, W = x.shape
    hd = C // heads
    wr = (lambda t: bf16_round(t)) if ctx.emu else (lambda t: t)
    tap = lambda sfx, z: ctx.tap(p + sfx, z.transpose(1, 2).reshape(B, z.shape[-1], H, W)) is None or z  # tokens -> NCHW
    t = x.flatten(2).transpose(1, 2)  # [B, N, C]
    xn = tap(".ln1", ctx.q(F.layer_norm(t, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)))
    a = p + ".attn"
    qkv = tap(".qkv", ctx.q(F.linear(xn, wr(sd[a + ".qkv.weight"]), sd[a + ".qkv.bias"])))          # per token: order-free
    n_tok = window * window
    qw = window_partition(qkv.view(B, H, W, 3 * C), window)                                             # [B*nW, n_tok, 3C]
    qw = qw.reshape(-1, n_tok, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qw[0] * (hd ** -0.5), qw[1], qw[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[a + ".relative_position_bias_table"][relative_position_index(window).view(-1)].view(n_tok, n_tok, heads)
    attn = attn + bias.permute(2, 0, 1).unsqueeze(0)
    o = (torch.softmax(attn, dim=-1) @ v).transpose(1, 2).reshape(-1, n_tok, C)
    o = tap(".attn", ctx.q(window_reverse(o, window, H, W).reshape(B, H * W, C)))
    t = tap(".proj", ctx.q(t + F.linear(o, wr(sd[a + ".proj.weight"]), sd[a + ".proj.bias"])))
    xn = tap(".ln2", ctx.q(F.layer_norm(t, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)))
    h = tap(".ff0", ctx.q(F.relu(F.linear(xn, wr(sd[p + ".feedforward.0.weight"]), sd[p + ".feedforward.0.bias"]))))
    t = tap(".ff3", ctx.q(t + F.linear(h, wr(sd[p + ".feedforward.3.weight"]), sd[p + ".feedforward.3.bias"])))
    return t.transpose(1, 2).reshape(B, C, H, W)


def features(x, sd, cfg, ctx=FP32) -> List[Tensor]:
    """Everything before the detection head: [p3, p4, p5] level features."""
    cfg = get_cfg(cfg)
    p3, p4, p5 = neck(backbone(ctx.q(x), sd, cfg, ctx), sd, ctx)
    if cfg["enhanced"]:  # EnhancedSkyEyeDetector.forward (detector.py:485-491) + D4
        p4 = ctx.tap("cross_attention_p5_p4.out", ctx.q(cla(p4, p5, sd, "cross_attention_p5_p4", ctx=ctx) + p4))
        p3 = ctx.tap("cross_attention_p4_p3.out", ctx.q(cla(p3, p4, sd, "cross_attention_p4_p3", ctx=ctx) + p3))
        hd = cfg["head_dim"]
        lv = []
        for i, f in enumerate((p3, p4, p5)):
            nh = max(f.shape[1] // hd, 1)
            if cfg["head"] == "windowed":
                lv.append(windowed_transformer_layer(f, sd, f"head_transformers.{i}", nh, cfg["window_size"], ctx))
            elif cfg["head"] == "transformer":
                lv.append(transformer_layer(f, sd, f"head_transformers.{i}", nh, ctx))
            else:
                lv.append(f)
        p3, p4, p5 = lv
    return [p3, p4, p5]


@torch.no_grad()
def forward(x: Tensor, sd: SD, cfg, emu: Optional[str] = None, taps: Optional[dict] = None) -> Tuple[Tensor, List[Tensor]]:
    """SkyEyeDetector.forward in eval mode (detector.py:300-324): (detections, raw_outputs).
    ``taps``: dict filled with every stored intermediate (see ``Ctx``)."""
    cfg = get_cfg(cfg)
    ctx = Ctx(emu, taps)
    raws = head(features(x.float(), sd, cfg, ctx), sd, cfg["nc"], ctx)
    det = decode(raws, x.shape[2:], cfg.get("anchors"))
    ctx.tap("det", det)
    return det, raws


@torch.no_grad()
def calibrate_bn(sd: SD, cfg, x: Tensor) -> SD:
    """Returns a copy of ``sd`` whose BatchNorm running statistics are the batch statistics of the
    fp32 reference forward pass on ``x`` (what a few training steps with momentum 1 would leave
    behind; nn.BatchNorm2d train-mode arithmetic, blocks.py:32).  A reference model at random init
    keeps running_mean 0 / running_var 1 (0.9 in the backbone, X17), so its activations grow by
    ~sqrt(fan_in * 2 / n) per conv: through skyeye_l's 111 convs they reach ~1e5 and the network's
    output is numerically meaningless in ANY arithmetic (fp32 vs bf16 emulation differ by 50-65 %).
    Calibrated statistics keep every activation O(1) -- the state a trained checkpoint is in --
    without touching a single weight.  Deterministic: same (sd, x) -> same result."""
    cfg = get_cfg(cfg)
    out = dict(sd)
    ctx = Ctx(None, None, calib=True)
    # every BatchNorm sits in the backbone / neck, upstream of the per-level heads: the heads need not run (and the windowed
    # heads could not, on a calibration image whose level maps are not multiples of the window)
    features(x.float(), out, dict(cfg, head="conv"), ctx)
    return out


RESIDUAL_GAMMA = 0.25


def make_calibrated_state_dict(cfg, seed: int = 0, calib_batch: int = 8, calib_hw=(320, 320), calib_seed: int = 4321) -> SD:
    """The state-dict recipe of the benchmark and of the whole-network parity tests ("trained-like"
    random weights): ``make_state_dict`` + the BN scale of every residual branch (BottleneckBlock.cv2,
    blocks.py:82) multiplied by RESIDUAL_GAMMA (the small-gamma residual init trained ResNets start
    from; keeps x + f(x) from doubling the variance 36 times) + BN statistics calibrated on
    ``calib_batch`` seeded uniform-noise images.  No conv / attention / transformer weight is touched.
    Measured on skyeye_l 640x640 (CPU, this file): every activation O(1..15), logits |.| < 10, and the
    oracle's own bf16 emulation sits 1.7 / 5.3 / 5.5 % rms from its fp32 result (P3 / P4 / P5 logits) --
    a random deep network amplifies rounding noise ~1.5x per stage, which no choice of statistics removes."""
    cfg = get_cfg(cfg)
    sd = make_state_dict(cfg, seed)
    for k in sd:
        if ".bottlenecks." in k and k.endswith(".cv2.bn.weight"):
            sd[k] = sd[k] * RESIDUAL_GAMMA
    g = np.random.Generator(np.random.PCG64([calib_seed, seed]))
    x = torch.from_numpy(g.integers(0, 256, (calib_batch, 3, *calib_hw), dtype=np.uint8)).float() / 255.0
    # ONE host thread: the calibration pass is fp32 CPU arithmetic whose summation order follows the thread count, and
    # torchrun starts its ranks with OMP_NUM_THREADS=1 -- with the default thread count the weights (and so the detection
    # checksum of bench.py's tiled leg) differed in the last bits between a plain run and a torchrun rank on the same box
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        return calibrate_bn(sd, cfg, x)
    finally:
        torch.set_num_threads(nt)
