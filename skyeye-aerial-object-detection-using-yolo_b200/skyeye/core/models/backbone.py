"""CSPDarknet-style backbone -> [P3 (stride 8), P4 (16), P5 (32)]  (reference
/root/reference/skyeye/core/models/backbone.py:12-159), lowered to native launches."""
from __future__ import annotations

import torch
import torch.nn as nn

from .attention import CombinedAttention
from .blocks import ConvolutionBlock, CSPBlock, FocusBlock, SPPBlock
from ...engine import Plan, View


class Backbone(nn.Module):
    def __init__(self, base_channels=64, depth_multiple=1.0, width_multiple=1.0):
        super().__init__()
        if width_multiple != 1.0:
            # the reference's neck only type-checks at width_multiple == 1
            raise ValueError("express width through base_channels; width_multiple must be 1.0")
        d = lambda n: max(round(n * depth_multiple), 1)
        c1, c2, c3, c4, c5 = (base_channels * m for m in (1, 2, 4, 8, 16))
        self.stage1 = nn.Sequential(FocusBlock(3, c1, kernel_size=3), ConvolutionBlock(c1, c2, 3, stride=2), CSPBlock(c2, c2, num_blocks=d(3)))
        self.stage2 = nn.Sequential(ConvolutionBlock(c2, c3, 3, stride=2), CSPBlock(c3, c3, num_blocks=d(9)))
        self.stage3 = nn.Sequential(ConvolutionBlock(c3, c4, 3, stride=2), CSPBlock(c4, c4, num_blocks=d(9)), CombinedAttention(c4))
        self.stage4 = nn.Sequential(ConvolutionBlock(c4, c5, 3, stride=2), CSPBlock(c5, c5, num_blocks=d(3)), SPPBlock(c5, c5))
        self.out_channels = [c3, c4, c5]

    def lower(self, plan: Plan, img_holder, n, h, w, outs=(None, None, None)):
        """outs: destination views for P3/P4/P5 (channel slices of the neck's concat buffers)."""
        x = self.stage1[0].lower_image(plan, img_holder, n, h, w, "b.s1.focus")
        x = self.stage1[1].lower(plan, x, name="b.s1.down")
        x = self.stage1[2].lower(plan, x, name="b.s1.csp")
        x = self.stage2[0].lower(plan, x, name="b.s2.down")
        p3 = self.stage2[1].lower(plan, x, outs[0], name="b.s2.csp")
        x = self.stage3[0].lower(plan, p3, name="b.s3.down")
        x = self.stage3[1].lower(plan, x, name="b.s3.csp")
        p4 = self.stage3[2].lower(plan, x, outs[1], name="b.s3.cbam")
        x = self.stage4[0].lower(plan, p4, name="b.s4.down")
        x = self.stage4[1].lower(plan, x, name="b.s4.csp")
        p5 = self.stage4[2].lower(plan, x, outs[2], name="b.s4.spp")
        return [p3, p4, p5]


class CSPDarknet(Backbone):
    pass


class SkyEyeBackbone(nn.Module):
    def __init__(self, base_channels=64, depth_multiple=1.0, width_multiple=1.0):
        super().__init__()
        self.backbone = CSPDarknet(base_channels, depth_multiple, width_multiple)
        self.channels = list(self.backbone.out_channels)  # true channels of [s2, s3, s4] (repair R2 of X2)

    def lower(self, plan, img_holder, n, h, w, outs=(None, None, None)):
        return self.backbone.lower(plan, img_holder, n, h, w, outs)
