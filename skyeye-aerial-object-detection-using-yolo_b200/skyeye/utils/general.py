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
