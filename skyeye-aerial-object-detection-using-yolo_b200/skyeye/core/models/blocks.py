"""Conv building blocks of the SkyEye backbone/neck, B200-native.

Each class keeps the reference's parameter layout (same attribute names => the same state-dict
keys as /root/reference/skyeye/core/models/blocks.py, so reference checkpoints load unchanged) but
owns no eager forward: ``lower(plan, x, out)`` appends fused native launches to an engine.Plan.
Fusions: BN folded into the weights, SiLU / residual / concat-slice write in the GEMM epilogue,
CSP cv1||cv2 as one GEMM, SPP's 9x9 and 13x13 pools as cascaded 5x5 pools (exact for max).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lowering as L
from ...engine import ACT_SILU, PackedConv, Plan, View


class ConvolutionBlock(nn.Module):
    """conv(bias=False, pad=k//2) -> BatchNorm -> SiLU   (reference blocks.py:10-41)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, kernel_size // 2, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.kernel_size, self.stride = kernel_size, stride
        self.in_channels, self.out_channels = in_channels, out_channels

    def packed(self, device, cin_pad=None) -> PackedConv:
        w, b = PackedConv.fold_bn(self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var,
                                  self.bn.eps)
        return PackedConv(w, b, device, cin_pad)

    def lower(self, plan: Plan, x: View, out: View = None, residual: View = None, upsample2x=False, name="conv") -> View:
        s = self.stride
        if out is None:
            up = 2 if upsample2x else 1
            out = plan.buf(x.n, x.h // s * up, x.w // s * up, self.out_channels)
        return plan.conv(name, x, self.packed(plan.device, x.c), out, s, ACT_SILU, residual, upsample2x, label=L.ref(self))

    def forward(self, x):
        return L.run_module(self, x)


class BottleneckBlock(nn.Module):
    """x + cv2_3x3(cv1_1x1(x))   (reference blocks.py:69-90; inside CSP expansion is 1.0)."""

    def __init__(self, in_channels, out_channels, shortcut=True, expansion=0.5):
        super().__init__()
        hidden = int(out_channels * expansion)
        self.cv1 = ConvolutionBlock(in_channels, hidden, 1, 1)
        self.cv2 = ConvolutionBlock(hidden, out_channels, 3, 1)
        self.use_shortcut = shortcut and in_channels == out_channels

    def lower(self, plan: Plan, x: View, out: View = None, name="bottleneck") -> View:
        t = self.cv1.lower(plan, x, name=name + ".cv1")
        out = x if out is None and self.use_shortcut else out  # in-place residual update of the CSP slice
        return self.cv2.lower(plan, t, out, residual=x if self.use_shortcut else None, name=name + ".cv2")

    def forward(self, x):
        return L.run_module(self, x)


class CSPBlock(nn.Module):
    """cv3(cat(bottlenecks(cv1(x)), cv2(x)))   (reference blocks.py:93-123)."""

    def __init__(self, in_channels, out_channels, num_blocks=1, shortcut=True, expansion=0.5):
        super().__init__()
        hidden = int(out_channels * expansion)
        self.cv1 = ConvolutionBlock(in_channels, hidden, 1, 1)
        self.cv2 = ConvolutionBlock(in_channels, hidden, 1, 1)
        self.cv3 = ConvolutionBlock(2 * hidden, out_channels, 1, 1)
        self.bottlenecks = nn.Sequential(*[BottleneckBlock(hidden, hidden, shortcut, 1.0) for _ in range(num_blocks)])
        self.hidden, self.out_channels = hidden, out_channels

    def lower(self, plan: Plan, x: View, out: View = None, name="csp") -> View:
        h = self.hidden
        cat = plan.buf(x.n, x.h, x.w, 2 * h)
        # cv1 and cv2 read the same input: one GEMM with N = 2*hidden writes both halves of the concat
        fused = PackedConv.concat([self.cv1.packed(plan.device, x.c), self.cv2.packed(plan.device, x.c)])
        plan.conv(name + ".cv1|cv2", x, fused, cat, 1, ACT_SILU, label=[(0, h, L.ref(self.cv1)), (h, 2 * h, L.ref(self.cv2))])
        y = cat.slice(0, h)
        for i, b in enumerate(self.bottlenecks):
            b.lower(plan, y, name=f"{name}.m{i}")
        return self.cv3.lower(plan, cat, out, name=name + ".cv3")

    def forward(self, x):
        return L.run_module(self, x)


class SPPBlock(nn.Module):
    """cv2(cat[x, mp5(x), mp9(x), mp13(x)]), x = cv1(in)   (reference blocks.py:126-149)."""

    def __init__(self, in_channels, out_channels, kernel_sizes=(5, 9, 13)):
        super().__init__()
        if tuple(kernel_sizes) != (5, 9, 13):
            raise NotImplementedError("B200 SPP lowers (5, 9, 13) as a 5x5 cascade; other sizes are not on the path")
        hidden = in_channels // 2
        self.cv1 = ConvolutionBlock(in_channels, hidden, 1, 1)
        self.cv2 = ConvolutionBlock(hidden * 4, out_channels, 1, 1)
        self.hidden = hidden

    def lower(self, plan: Plan, x: View, out: View = None, name="spp") -> View:
        h = self.hidden
        cat = plan.buf(x.n, x.h, x.w, 4 * h)
        self.cv1.lower(plan, x, cat.slice(0, h), name=name + ".cv1")
        sl = [cat.slice(i * h, (i + 1) * h) for i in range(4)]
        if 2 * x.h * x.w * 32 <= 110 * 1024 and h % 16 == 0:
            # one pass: the map is read once, the 5 / 9 / 13 pools (a 5x5 cascade in shared memory) go to their concat slices
            plan.add(f"{name}.mp", lambda s: L.E.spp_pools(sl[0], sl[1], sl[2], sl[3], s), "maxpool", 0.0, 2.0 * 4 * x.n * x.h * x.w * h,
                     outs=[dict(view=sl[i + 1], label=L.ref(self, f".mp{5 + 4 * i}")) for i in range(3)])
        else:
            for i in range(3):  # mp9 = mp5(mp5), mp13 = mp5(mp5(mp5)) with -inf padding: exact
                plan.add(f"{name}.mp{5 + 4 * i}", lambda s, a=sl[i], b=sl[i + 1]: L.E.maxpool5(a, b, s), "maxpool", 0.0, 4.0 * x.n * x.h * x.w * h,
                         outs=[dict(view=sl[i + 1], label=L.ref(self, f".mp{5 + 4 * i}"))])
        return self.cv2.lower(plan, cat, out, name=name + ".cv2")

    def forward(self, x):
        return L.run_module(self, x)


class FocusBlock(nn.Module):
    """space-to-depth (4 strided slices + cat) -> ConvolutionBlock   (reference blocks.py:152-182).
    The slicing, the NCHW->NHWC transpose and the fp32->bf16 cast are one kernel; the 12 focus
    channels are zero-padded to the GEMM's K granule (the padded weight columns are zero)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1):
        super().__init__()
        if in_channels != 3:
            raise NotImplementedError("FocusBlock is lowered for RGB input")
        self.conv = ConvolutionBlock(in_channels * 4, out_channels, kernel_size, stride)

    def lower_image(self, plan: Plan, img_holder, n, h, w, name="focus") -> View:
        c = self.conv
        if c.kernel_size != 3 or c.stride != 1:
            raise NotImplementedError("FocusBlock is lowered for the k3 s1 conv on the path (backbone.py:47)")
        wf, bf = PackedConv.fold_bn(c.conv.weight, c.bn.weight, c.bn.bias, c.bn.running_mean, c.bn.running_var, c.bn.eps)
        pw = L.E.PackedFocusConv(wf, bf, plan.device)
        out = plan.buf(n, h // 2, w // 2, c.out_channels)
        ws = plan.ws(L.E.N.lib().skb_focus_conv_workspace_bytes(n, h, w))
        plan.keep.append(pw)
        ho, wo = h // 2, w // 2
        flops = 2.0 * n * ho * wo * c.out_channels * 9 * 12
        nbytes = 3.0 * n * h * w + 2.0 * (2.0 * n * ho * (wo + 4) * 16) + 2.0 * n * ho * wo * c.out_channels  # uint8 image
        if plan.tile_src is not None:  # tiles of larger frames read in place through a (frame, y0, x0) table
            table = plan.tile_src
            plan.add(name, lambda s: L.E.focus_conv_tiles(img_holder[0], table, (h, w), pw, out, ws, ACT_SILU, s), "conv", flops, nbytes, 2,
                     outs=[dict(view=out, label=L.ref(c))])
            return out
        plan.add(name, lambda s: L.E.focus_conv(img_holder[0], pw, out, ws, ACT_SILU, s), "conv", flops, nbytes, 2,
                 outs=[dict(view=out, label=L.ref(c))])
        return out
