"""SkyEye detector assembly, B200-native: backbone -> PAN/FPN neck -> (cross-layer attention ->
transformer heads) -> 1x1 detection heads -> anchor decode.

Mirrors the interface of /root/reference/skyeye/core/models/detector.py (class and attribute names,
state-dict keys, ``model(tensor) -> (detections, raw_outputs)`` in eval mode, detector.py:300-324)
and the README API (``SkyEyeDetector(weights=...)``, ``model(image) -> Results``; README.md:39-54).
The forward pass is a cached launch plan of native kernels per input shape; nothing runs in eager
PyTorch and there is no CPU fallback.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import yaml

from ... import engine as E
from .. import _lowering as L
from ...engine import ACT_NONE, PackedConv, Plan, View
from .attention import CrossLayerAttention, TransformerLayer, WindowedTransformerLayer
from .backbone import SkyEyeBackbone
from .blocks import ConvolutionBlock, CSPBlock

CONFIG_DIR = Path(__file__).resolve().parents[3] / "configs" / "models"

# detector.py:39-43
DEFAULT_ANCHORS = [
    [[10, 13], [16, 30], [33, 23]],
    [[30, 61], [62, 45], [59, 119]],
    [[116, 90], [156, 198], [373, 326]],
]


class DetectionHead(nn.Module):
    """Per-level 1x1 conv (+bias) -> [B, na, H, W, no] and the anchor decode (detector.py:18-145)."""

    def __init__(self, num_classes=80, anchors=None, channels=None):
        super().__init__()
        self.num_classes = num_classes
        self.num_outputs = num_classes + 5
        self.anchors = anchors if anchors is not None else DEFAULT_ANCHORS
        self.num_anchors = len(self.anchors[0])
        self.num_layers = len(self.anchors)
        channels = [256, 512, 1024] if channels is None else channels
        self.detection_layers = nn.ModuleList(nn.Conv2d(ch, self.num_anchors * self.num_outputs, 1) for ch in channels)

    def lower(self, plan: Plan, feats: Sequence[View], input_hw) -> tuple:
        na, no = self.num_anchors, self.num_outputs
        cpad = (na * no + 7) // 8 * 8
        raws = []
        for i, (f, layer) in enumerate(zip(feats, self.detection_layers)):
            r = plan.buf(f.n, f.h, f.w, cpad, torch.float32)
            plan.conv(f"head.{i}", f, PackedConv(layer.weight, layer.bias, plan.device), r, 1, ACT_NONE, label=L.ref(layer))
            raws.append(r)
        rows = sum(na * r.h * r.w for r in raws)
        det = torch.empty((raws[0].n, rows, no), dtype=torch.float32, device=plan.device)
        raw_out = [torch.empty((r.n, na, r.h, r.w, no), dtype=torch.float32, device=plan.device) for r in raws]
        plan.keep += [det, raw_out]
        plan.add("decode", lambda s: E.decode(raws, na, no, self.anchors, input_hw, det, raw_out, s), "decode", 0.0,
                 4.0 * raws[0].n * rows * no * 3,
                 outs=[dict(view=det, label="det" if L.ref(self) else None, layout="flat")] +
                      [dict(view=t, label=L.ref(layer), layout="raw", na=na) for t, layer in zip(raw_out, self.detection_layers)])
        return det, raw_out


class FeatureNeck(nn.Module):
    """Top-down + bottom-up fusion (detector.py:148-231). torch.cat / F.interpolate never run: every
    producer writes straight into its channel slice of the consumer's concat buffer and the nearest
    2x upsample is replicated stores in the lateral conv's epilogue."""

    def __init__(self, in_channels, width_multiple=1.0):
        super().__init__()
        if width_multiple != 1.0:
            raise ValueError("express width through base_channels; width_multiple must be 1.0")
        c3, c4, c5 = in_channels
        self.lateral_conv5 = ConvolutionBlock(c5, c4, 1, 1)
        self.lateral_conv4 = ConvolutionBlock(c4, c3, 1, 1)
        self.fpn_conv4 = CSPBlock(c4 * 2, c4, num_blocks=3)
        self.fpn_conv3 = CSPBlock(c3 * 2, c3, num_blocks=3)
        self.downsample3 = ConvolutionBlock(c3, c3, 3, 2)
        self.downsample4 = ConvolutionBlock(c4, c4, 3, 2)
        self.pan_conv4 = CSPBlock(c3 + c4, c4, num_blocks=3)
        self.pan_conv5 = CSPBlock(c4 + c5, c5, num_blocks=3)
        self.in_channels = [c3, c4, c5]
        self.out_channels = [c3, c4, c5]

    def alloc(self, plan: Plan, n, h, w):
        """Concat buffers; returns them plus the slices the backbone must write P3/P4/P5 into."""
        c3, c4, c5 = self.in_channels
        bufs = dict(p4_merged=plan.buf(n, h // 16, w // 16, 2 * c4),   # [up(lateral5(p5)), p4]      detector.py:215
                    p3_merged=plan.buf(n, h // 8, w // 8, 2 * c3),     # [up(lateral4(p4)), p3]      detector.py:219
                    p4_cat=plan.buf(n, h // 16, w // 16, c3 + c4),     # [down(p3_proc), p4_proc]    detector.py:224
                    p5_cat=plan.buf(n, h // 32, w // 32, c4 + c5))     # [down(p4_out), p5]          detector.py:228
        outs = (bufs["p3_merged"].slice(c3, 2 * c3), bufs["p4_merged"].slice(c4, 2 * c4), bufs["p5_cat"].slice(c4, c4 + c5))
        return bufs, outs

    def lower(self, plan: Plan, feats: Sequence[View], bufs) -> List[View]:
        c3, c4, c5 = self.in_channels
        p3, p4, p5 = feats
        self.lateral_conv5.lower(plan, p5, bufs["p4_merged"].slice(0, c4), upsample2x=True, name="n.lat5+up")
        self.lateral_conv4.lower(plan, p4, bufs["p3_merged"].slice(0, c3), upsample2x=True, name="n.lat4+up")  # raw backbone P4 (:211)
        p4_proc = self.fpn_conv4.lower(plan, bufs["p4_merged"], bufs["p4_cat"].slice(c3, c3 + c4), name="n.fpn4")
        p3_proc = self.fpn_conv3.lower(plan, bufs["p3_merged"], name="n.fpn3")
        self.downsample3.lower(plan, p3_proc, bufs["p4_cat"].slice(0, c3), name="n.down3")
        p4_out = self.pan_conv4.lower(plan, bufs["p4_cat"], name="n.pan4")
        self.downsample4.lower(plan, p4_out, bufs["p5_cat"].slice(0, c4), name="n.down4")
        p5_out = self.pan_conv5.lower(plan, bufs["p5_cat"], name="n.pan5")
        return [p3_proc, p4_out, p5_out]


def _load_cfg(cfg) -> dict:
    if isinstance(cfg, dict):
        return dict(cfg)
    p = Path(cfg)
    if not p.exists():
        cand = CONFIG_DIR / p.name
        if not cand.suffix:
            cand = cand.with_suffix(".yaml")
        p = cand
    with open(p, errors="ignore") as f:
        d = yaml.safe_load(f)
    if not isinstance(d, dict):
        raise ValueError(f"model config {p} is empty")
    return d


class Results:
    """Per-image detections of a README-style call ``model(image)`` (README.md:46-53).  ``pred[i]`` holds rows
    ``[x1, y1, x2, y2, conf, cls]`` in ORIGINAL-image pixels (rows of the reference wrapper's 7-column form
    ``[cx, cy, w, h, obj, cls_prob, cls_id]`` are accepted too)."""

    def __init__(self, pred: List[torch.Tensor], images=None, names=None, files=None):
        self.pred, self.images, self.names = pred, images, names or []
        self.files = files or [f"image{i}.jpg" for i in range(len(pred))]

    def __len__(self):
        return len(self.pred)

    @staticmethod
    def _xyxy_conf_cls(p: torch.Tensor):
        p = p.detach().float().cpu()
        if p.shape[1] > 6:  # reference wrapper rows (quirk X8): centre form, class id in column 6
            xy, wh = p[:, 0:2], p[:, 2:4]
            return torch.cat((xy - wh / 2, xy + wh / 2), 1), p[:, 4], p[:, 6]
        return p[:, :4], p[:, 4], p[:, 5]

    def render(self) -> list:
        """Annotated copies (BGR uint8) of the input images: box + ``name conf`` tag per detection."""
        from ...utils.visualization import ImageAnnotator, colors
        out = []
        for im, p in zip(self.images or [], self.pred):
            ann = ImageAnnotator(im.copy())
            box, conf, cls = self._xyxy_conf_cls(p)
            for b, c, k in zip(box.tolist(), conf.tolist(), cls.tolist()):
                name = self.names[int(k)] if int(k) < len(self.names) else str(int(k))
                ann.box_label(b, f"{name} {c:.2f}", color=colors(int(k)))
            out.append(ann.result())
        return out

    def save(self, save_dir="outputs/"):
        """Writes ``<name>.jpg`` (annotated image) and ``<name>.txt`` (rows ``cls x1 y1 x2 y2 conf`` in pixels) per image."""
        import cv2
        os.makedirs(save_dir, exist_ok=True)
        drawn = self.render()
        for i, (f, p) in enumerate(zip(self.files, self.pred)):
            stem = Path(f).stem
            box, conf, cls = self._xyxy_conf_cls(p)
            with open(os.path.join(save_dir, stem + ".txt"), "w") as fh:
                for b, c, k in zip(box.tolist(), conf.tolist(), cls.tolist()):
                    fh.write(("%d " + "%g " * 5).rstrip() % (int(k), *b, c) + "\n")
            if i < len(drawn):
                cv2.imwrite(os.path.join(save_dir, stem + ".jpg"), drawn[i])
        return save_dir

    def show(self):
        """Prints a one-line summary per image and, when a display is available, opens the annotated images."""
        for f, p in zip(self.files, self.pred):
            _, _, cls = self._xyxy_conf_cls(p)
            ids, cnt = (torch.unique(cls.long(), return_counts=True) if p.shape[0] else (torch.zeros(0), torch.zeros(0)))
            parts = [f"{int(n)} {self.names[int(k)] if int(k) < len(self.names) else int(k)}" for k, n in zip(ids.tolist(), cnt.tolist())]
            print(f"{f}: {p.shape[0]} detections" + (" (" + ", ".join(parts) + ")" if parts else ""))
        if os.environ.get("DISPLAY") and self.images:
            try:
                import cv2
                for f, im in zip(self.files, self.render()):
                    cv2.imshow(str(f), im)
                cv2.waitKey(1)
            except Exception:
                pass
        return self


class SkyEyeDetector(nn.Module):
    def __init__(self, cfg="skyeye_s.yaml", channels=3, num_classes=None, anchors=None, weights=None, device=None):
        super().__init__()
        if weights is not None and cfg == "skyeye_s.yaml":  # README: SkyEyeDetector(weights='weights/skyeye_l.pt')
            stem = Path(str(weights)).stem
            if (CONFIG_DIR / f"{stem}.yaml").exists():
                cfg = f"{stem}.yaml"
        self.cfg = _load_cfg(cfg)
        if num_classes and num_classes != self.cfg.get("nc"):
            self.cfg["nc"] = num_classes
        if anchors:
            self.cfg["anchors"] = anchors
        if channels != 3:
            raise NotImplementedError("the B200 path is lowered for 3-channel images")
        self.backbone = SkyEyeBackbone(self.cfg.get("base_channels", 64), self.cfg.get("depth_multiple", 1.0),
                                       self.cfg.get("width_multiple", 1.0))
        self.neck = FeatureNeck(self.backbone.channels, self.cfg.get("width_multiple", 1.0))
        self.detection_head = DetectionHead(self.cfg["nc"], self.cfg.get("anchors"), self.neck.out_channels)
        self._initialize_weights()
        self.stride = torch.tensor([8, 16, 32])
        self.names = [str(i) for i in range(self.cfg["nc"])]
        self._plans = {}
        self._img = [None]
        # The forward plan is captured into a CUDA graph per input shape and replayed, and the returned tensors are the
        # plan's own output buffers: they are valid until the next forward call with the same input shape.  Set
        # reuse_output_buffers = False to get private copies (two device copies per call), use_cuda_graph = False to launch
        # kernel by kernel.
        self.reuse_output_buffers = True
        self.use_cuda_graph = True
        self.eval()
        if weights is not None:
            self.load_from_pretrained(weights)
        if device is not None:
            self.to(device)

    # -- weights ---------------------------------------------------------------------------------
    def _initialize_weights(self):
        """Random init with the reference's distributions (detector.py:326-341, with the bias guard R1)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, (2.0 / n) ** 0.5)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0.0, 0.01)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def load_from_pretrained(self, weights_path):
        """Checkpoint formats of detector.py:353-359: {'model': nn.Module}, {'state_dict': ...} or a bare
        state dict; keys filtered by name + shape, strict=False."""
        ckpt = torch.load(weights_path, map_location="cpu", weights_only=False)
        if isinstance(ckpt, dict) and "model" in ckpt and hasattr(ckpt["model"], "state_dict"):
            sd = ckpt["model"].float().state_dict()
        else:
            sd = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
        own = self.state_dict()
        ok = {k: v for k, v in sd.items() if k in own and v.shape == own[k].shape}
        self.load_state_dict(ok, strict=False)
        print(f"Loaded {len(ok)}/{len(own)} layers from {weights_path}")
        # strict=False hides parameters the checkpoint did not provide (they keep their random init): say which
        self.uninitialized_keys = [k for k in own if k not in ok and not k.endswith("num_batches_tracked")]
        self.unexpected_keys = [k for k in sd if k not in ok]
        if self.uninitialized_keys:
            import warnings
            mods = sorted({k.rsplit(".", 1)[0] for k in self.uninitialized_keys})
            warnings.warn(f"{weights_path}: {len(self.uninitialized_keys)} parameters of {type(self).__name__} were NOT in the checkpoint "
                          f"(or had another shape) and keep their random initialisation: {mods[:6]}{' ...' if len(mods) > 6 else ''}; "
                          f"{len(self.unexpected_keys)} checkpoint entries were not used", stacklevel=2)
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._plans.clear()  # packed weights are baked into the plans
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # -- lowering --------------------------------------------------------------------------------
    def _lower_features(self, plan: Plan, n, h, w) -> List[View]:
        bufs, outs = self.neck.alloc(plan, n, h, w)
        feats = self.backbone.lower(plan, self._img, n, h, w, outs)
        return self.neck.lower(plan, feats, bufs)

    def _build_plan(self, n, h, w, device, tile_table=None) -> Plan:
        if h % 32 or w % 32:
            raise ValueError(f"input H, W must be multiples of 32 (got {h}x{w}); letterbox first")
        plan = Plan(device)
        plan.tile_src = tile_table
        for name, m in self.named_modules():  # reference state-dict paths label the plan's outputs (Plan.run_teacher_forced)
            m._ref = name
        feats = self._lower_features(plan, n, h, w)
        plan.det, plan.raw_out = self.detection_head.lower(plan, feats, (h, w))
        plan.feats = feats
        return plan

    def plan_for(self, x: torch.Tensor) -> Plan:
        key = (tuple(x.shape), x.dtype, x.device.index)  # uint8 and fp32 images take different first kernels / graph inputs
        plan = self._plans.get(key)
        if plan is None:
            n, _, h, w = x.shape
            plan = self._build_plan(n, h, w, x.device)
            self._plans[key] = plan
        return plan

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, augment=False):
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            return self.predict(x)
        if self.training:
            raise NotImplementedError("training is outside the B200 forward-path scope (SURVEY.md §2)")
        if not x.is_cuda:
            raise RuntimeError("SkyEyeDetector (B200) needs a CUDA tensor; there is no CPU fallback")
        # fp32 images in [0,1] as in the reference; uint8 images are scaled by 1/255 inside the first kernel
        xin = x if (x.dtype in (torch.float32, torch.uint8) and x.is_contiguous()) else x.float().contiguous()
        with torch.cuda.device(xin.device):  # native launches use the CURRENT device's stream and per-device function attributes
            plan = self.plan_for(xin)
            if self.use_cuda_graph:
                if plan.graph is None:
                    plan.static_in = xin.clone()
                    self._img[0] = plan.static_in
                    plan.capture()
                plan.static_in.copy_(xin)
                plan.replay()
            else:
                self._img[0] = xin
                plan.run()
        if self.reuse_output_buffers:
            return plan.det, plan.raw_out
        return plan.det.clone(), [r.clone() for r in plan.raw_out]

    @torch.no_grad()
    def forward_tiles(self, frames: torch.Tensor, tiles: torch.Tensor, tile_hw=(1280, 1280)):
        """forward() on windows of larger frames WITHOUT materialising the tile batch (tiled inference of 4K drone frames,
        SURVEY.md D8): frames uint8 / fp32 CUDA [F,3,FH,FW], tiles int32 [n,3] rows (frame, y0, x0).  Image i of the returned
        (detections, raw_outputs) is the tile_hw window at tiles[i]; coordinates are tile-local.  The plan (and its CUDA
        graph) is bound to the frame buffer's address and owns a device copy of the table, refreshed on every call."""
        if not (frames.is_cuda and frames.dim() == 4 and frames.is_contiguous() and frames.dtype in (torch.uint8, torch.float32)):
            raise RuntimeError("forward_tiles needs a contiguous uint8 / fp32 CUDA frame tensor [F,3,H,W]")
        n, (th, tw) = int(tiles.shape[0]), tile_hw
        key = ("tiles", n, th, tw, tuple(frames.shape), frames.dtype, frames.device.index, frames.data_ptr())
        with torch.cuda.device(frames.device):
            plan = self._plans.get(key)
            if plan is None:
                table = torch.zeros((n, 3), dtype=torch.int32, device=frames.device)
                plan = self._build_plan(n, th, tw, frames.device, tile_table=table)
                plan.table = table
                self._plans[key] = plan
            plan.table.copy_(tiles.to(torch.int32), non_blocking=True)
            self._img[0] = frames
            if self.use_cuda_graph:
                if plan.graph is None:
                    plan.capture()
                plan.replay()
            else:
                plan.run()
        if self.reuse_output_buffers:
            return plan.det, plan.raw_out
        return plan.det.clone(), [r.clone() for r in plan.raw_out]

    @torch.no_grad()
    def predict(self, source, img_size=640, conf_thres=0.25, iou_thres=0.45, max_det=300) -> Results:
        """README-style call on image path(s) / HWC BGR uint8 array(s): letterbox -> forward -> NMS -> boxes mapped back to
        the original image.  Rows are ``[x1, y1, x2, y2, conf, cls]`` (what the reference wrapper's docstring promises,
        metrics.py:383; its actual 7-column centre-form rows cannot be drawn)."""
        from ...utils.general import letterbox_geometry, load_images_gpu, scale_boxes
        from ...utils.metrics import non_max_suppression
        dev = next(self.parameters()).device
        batch, files, origs = load_images_gpu(source, img_size, dev)  # letterbox + BGR->RGB + HWC->CHW on the GPU, uint8
        det, _ = self.forward(batch)
        rows = non_max_suppression(det, conf_thres, iou_thres, max_detections=max_det, compat="fixed")
        out = []
        for r, im in zip(rows, origs):
            r = r.clone()
            _, _, _, _, top, left, ratio = letterbox_geometry(im.shape[0], im.shape[1], img_size)
            scale_boxes(batch.shape[2:], r[:, :4], im.shape[:2], ((ratio, ratio), (left, top)))
            out.append(r)
        return Results(out, origs, self.names, files)


class EnhancedSkyEyeDetector(SkyEyeDetector):
    """+ cross-layer attention between neck levels (detector.py:436-501, with repair R4) and, for the
    skyeye_l variant, a TransformerLayer per level in front of the detection head."""

    def __init__(self, cfg="skyeye_l.yaml", channels=3, num_classes=None, anchors=None, weights=None, device=None):
        super().__init__(cfg, channels, num_classes, anchors, None, None)
        c3, c4, c5 = self.neck.out_channels
        self.cross_attention_p5_p4 = CrossLayerAttention(c4, c5, region_size=2, heads=4)
        self.cross_attention_p4_p3 = CrossLayerAttention(c3, c4, region_size=2, heads=4)
        # default = the reference architecture (bare 1x1 heads, detector.py:436-501); the transformer heads of the skyeye_l
        # variant are selected by "head: transformer" in the model config
        # "head: windowed"  swaps the global attention for the reference's WindowedSelfAttention class
        head, hd = self.cfg.get("head", "conv"), self.cfg.get("head_dim", 64)
        if head == "transformer":
            self.head_transformers = nn.ModuleList(TransformerLayer(c, max(c // hd, 1)) for c in (c3, c4, c5))
        elif head == "windowed":
            ws = self.cfg.get("window_size", 8)
            self.head_transformers = nn.ModuleList(WindowedTransformerLayer(c, max(c // hd, 1), ws) for c in (c3, c4, c5))
        elif head == "conv":
            self.head_transformers = None
        else:
            raise ValueError(f"head: {head!r} (expected conv | transformer | windowed)")
        self._initialize_weights()
        self.eval()
        if weights is not None:
            self.load_from_pretrained(weights)
        if device is not None:
            self.to(device)

    def _lower_features(self, plan: Plan, n, h, w) -> List[View]:
        p3, p4, p5 = super()._lower_features(plan, n, h, w)
        p4e = self.cross_attention_p5_p4.lower(plan, p4, p5, residual=p4, name="cla54")   # CLA(p4, p5) + p4   :488
        p3e = self.cross_attention_p4_p3.lower(plan, p3, p4e, residual=p3, name="cla43")  # CLA(p3, p4e) + p3  :489
        lv = [p3e, p4e, p5]
        if self.head_transformers is not None:
            lv = [t.lower(plan, f, name=f"tl{i}") for i, (t, f) in enumerate(zip(self.head_transformers, lv))]
        return lv


def parse_model(model_cfg, in_channels=3) -> dict:
    cfg = _load_cfg(model_cfg)
    return {"base_channels": cfg.get("base_channels", 64), "depth_multiple": cfg.get("depth_multiple", 1.0),
            "width_multiple": cfg.get("width_multiple", 1.0), "nc": cfg.get("nc", 80), "in_channels": in_channels,
            "anchors": cfg.get("anchors"), "detector": cfg.get("detector", "base"), "head": cfg.get("head", "conv"),
            "head_dim": cfg.get("head_dim", 64), "window_size": cfg.get("window_size", 8)}


def construct_model(model_cfg, in_channels=3, num_classes=None, anchors=None):
    cfg = parse_model(model_cfg, in_channels)
    if num_classes is not None:
        cfg["nc"] = num_classes
    if anchors is not None:
        cfg["anchors"] = anchors
    cls = EnhancedSkyEyeDetector if cfg.get("detector") == "enhanced" else SkyEyeDetector
    return cls(cfg, in_channels)


def load_model(weights=None, cfg=None, device="cuda"):
    """The loader the reference CLIs import but never define (validate.py:21,185)."""
    if cfg is None:
        stem = Path(str(weights)).stem if weights else "skyeye_s"
        cfg = f"{stem}.yaml" if (CONFIG_DIR / f"{stem}.yaml").exists() else "skyeye_s.yaml"
    model = construct_model(cfg)
    if weights:
        model.load_from_pretrained(weights)
    return model.to(device).eval()
