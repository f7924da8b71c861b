"""Host utilities around the path: image loading + letterbox for the README-style call
(reference skyeye/core/data/augmentation.py:442-496 letterbox; detect.py:131-135 BGR->RGB, /255)."""
from __future__ import annotations

from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np
import torch


def letterbox(img: np.ndarray, new_shape=640, color=(114, 114, 114), stride=32) -> Tuple[np.ndarray, float, Tuple[int, int]]:
    """Resize keeping aspect ratio, pad to a stride-multiple rectangle with ``color``."""
    import cv2
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    h, w = img.shape[:2]
    r = min(new_shape[0] / h, new_shape[1] / w)
    nh, nw = int(round(h * r)), int(round(w * r))
    if (nh, nw) != (h, w):
        img = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    ph, pw = (-nh) % stride, (-nw) % stride
    top, left = ph // 2, pw // 2
    img = cv2.copyMakeBorder(img, top, ph - top, left, pw - left, cv2.BORDER_CONSTANT, value=color)
    return img, r, (left, top)


def load_images(source, img_size=640):
    """path | ndarray(HWC, BGR uint8) | list of those -> (float tensor [B,3,H,W] in [0,1], names, originals)."""
    import cv2
    items = source if isinstance(source, (list, tuple)) else [source]
    ims, names, origs = [], [], []
    for i, it in enumerate(items):
        if isinstance(it, (str, Path)):
            im = cv2.imread(str(it))
            if im is None:
                raise FileNotFoundError(it)
            names.append(str(it))
        else:
            im = np.asarray(it)
            names.append(f"image{i}.jpg")
        origs.append(im)
        lb, _, _ = letterbox(im, img_size)
        ims.append(lb)
    H = max(i.shape[0] for i in ims)
    W = max(i.shape[1] for i in ims)
    batch = np.full((len(ims), H, W, 3), 114, dtype=np.uint8)
    for b, im in enumerate(ims):
        batch[b, : im.shape[0], : im.shape[1]] = im
    t = torch.from_numpy(np.ascontiguousarray(batch[..., ::-1].transpose(0, 3, 1, 2))).float() / 255.0
    return t, names, origs


def xywh2xyxy(x):
    """[cx, cy, w, h] -> [x1, y1, x2, y2] (imported but never defined by the reference, validate.py:21-24)."""
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    y[..., 0] = x[..., 0] - x[..., 2] / 2
    y[..., 1] = x[..., 1] - x[..., 3] / 2
    y[..., 2] = x[..., 0] + x[..., 2] / 2
    y[..., 3] = x[..., 1] + x[..., 3] / 2
    return y


def xyxy2xywh(x):
    """[x1, y1, x2, y2] -> [cx, cy, w, h]."""
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None):
    """Map xyxy boxes from the letterboxed network input (img1_shape = (h, w)) back to the original image
    (img0_shape), in place, and clip them (call sites validate.py:274,279; undefined in the reference, X11).
    ratio_pad = ((gain_h, gain_w), (pad_w, pad_h)) as produced by the dataset's letterbox (dataset.py shapes)."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = ((img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2)
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    boxes[..., [0, 2]] -= pad[0]
    boxes[..., [1, 3]] -= pad[1]
    boxes[..., :4] /= gain
    if isinstance(boxes, torch.Tensor):
        boxes[..., 0].clamp_(0, img0_shape[1]); boxes[..., 1].clamp_(0, img0_shape[0])
        boxes[..., 2].clamp_(0, img0_shape[1]); boxes[..., 3].clamp_(0, img0_shape[0])
    else:
        boxes[..., [0, 2]] = boxes[..., [0, 2]].clip(0, img0_shape[1])
        boxes[..., [1, 3]] = boxes[..., [1, 3]].clip(0, img0_shape[0])
    return boxes


def check_img_size(img_size, stride=32, s=None):
    """Round the image size up to a stride multiple (general.py:248-268; validate.py:188 calls it with s=)."""
    stride = int(s if s is not None else stride)
    import math
    new = max(int(math.ceil(img_size / stride) * stride), stride)
    return new


def letterbox_geometry(h: int, w: int, new_shape=640, stride: int = 32):
    """Sizes of `letterbox` (augmentation.py:442-496 semantics as used by this package's host version above):
    returns (H, W, new_h, new_w, top, left, ratio) of the padded output and the resized image inside it."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    nh, nw = int(round(h * r)), int(round(w * r))
    ph, pw = (-nh) % stride, (-nw) % stride
    return nh + ph, nw + pw, nh, nw, ph // 2, pw // 2, r


def letterbox_gpu(img, new_shape=640, color: int = 114, stride: int = 32, out: "torch.Tensor" = None):
    """letterbox + BGR->RGB + HWC->CHW on the GPU in one kernel (skb_letterbox_u8).  img: uint8 [h, w, 3] BGR, numpy or
    torch (host or device).  Returns (uint8 CUDA tensor [3, H, W] RGB, ratio, (left, top)) -- ready for model(x[None])."""
    from .. import _native as N
    t = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
    assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, "expected a uint8 HWC BGR image"
    t = t.contiguous().cuda(non_blocking=True)
    h, w = int(t.shape[0]), int(t.shape[1])
    H, W, nh, nw, top, left, r = letterbox_geometry(h, w, new_shape, stride)
    if out is None:
        out = torch.empty((3, H, W), dtype=torch.uint8, device=t.device)
    assert out.shape == (3, H, W) and out.is_contiguous() and out.is_cuda
    N.check(N.lib().skb_letterbox_u8(t.data_ptr(), h, w, 3 * w, out.data_ptr(), H, W, nh, nw, top, left, int(color),
                                     torch.cuda.current_stream().cuda_stream), "skb_letterbox_u8")
    return out, r, (left, top)


def load_images_gpu(source, img_size=640, device="cuda"):
    """Like load_images, with the letterbox / colour / layout conversion on the GPU: returns (uint8 CUDA tensor
    [B,3,H,W] RGB -- the model scales by 1/255 in its first kernel --, names, originals).  Images whose letterboxed
    sizes differ are padded (114) to the largest, as the host version does."""
    import cv2
    items = source if isinstance(source, (list, tuple)) else [source]
    names, origs = [], []
    for i, it in enumerate(items):
        if isinstance(it, (str, Path)):
            im = cv2.imread(str(it))
            if im is None:
                raise FileNotFoundError(it)
            names.append(str(it))
        else:
            im = np.asarray(it)
            names.append(f"image{i}.jpg")
        origs.append(im)
    geo = [letterbox_geometry(im.shape[0], im.shape[1], img_size) for im in origs]
    H, W = max(g[0] for g in geo), max(g[1] for g in geo)
    batch = torch.full((len(origs), 3, H, W), 114, dtype=torch.uint8, device=device)
    with torch.cuda.device(batch.device):
        for b, (im, g) in enumerate(zip(origs, geo)):
            if (g[0], g[1]) == (H, W):
                letterbox_gpu(im, img_size, out=batch[b])
            else:  # smaller letterbox: top-left aligned inside the batch frame like the host version
                t, _, _ = letterbox_gpu(im, img_size)
                batch[b, :, : g[0], : g[1]] = t
    return batch, names, origs
