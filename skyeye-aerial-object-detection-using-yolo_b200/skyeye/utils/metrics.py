"""``non_max_suppression`` with the reference signature (skyeye/utils/metrics.py:361-369), running as
B200 kernels.  The mAP bookkeeping in the rest of the reference's metrics.py is outside the hot
path (SURVEY.md §2, row N2 of §8f)."""
from __future__ import annotations

from typing import List

import torch

from .nms import batched_nms_padded


def non_max_suppression(prediction, conf_threshold=0.25, iou_threshold=0.45, classes=None, agnostic=False, multi_label=False,
                        max_detections=300, compat="reference") -> List[torch.Tensor]:
    """Returns one tensor per image.  compat="reference" reproduces the reference rows bit for bit,
    quirks included: (n,7) rows [cx,cy,w,h,obj,cls_prob,cls_id] for nc>1, (n,6) for
    nc==1, zeros((0,6)) for images without candidates.  compat="fixed" returns what the reference
    docstring promises: (n,6) [x1,y1,x2,y2,conf,cls] with class-aware NMS on corner boxes."""
    rows, counts = batched_nms_padded(prediction, conf_threshold, iou_threshold, classes, agnostic, multi_label, max_detections, compat)
    nc = prediction.shape[2] - 5
    ncol = 6 if (compat == "fixed" or nc <= 1) else 7
    cnt = counts.tolist()  # the single device->host sync of the wrapper
    out = []
    for b, c in enumerate(cnt):
        out.append(rows[b, :c, :ncol] if c > 0 else torch.zeros((0, 6), device=prediction.device))
    return out
