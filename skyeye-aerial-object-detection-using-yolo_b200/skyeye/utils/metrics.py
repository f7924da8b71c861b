"""``non_max_suppression`` with the reference signature (skyeye/utils/metrics.py:361-369), running as
B200 kernels.  The mAP bookkeeping in the rest of the reference's metrics.py is outside the hot
path (SURVEY.md §2, row N2 of §8f)."""
from __future__ import annotations

from pathlib import Path
from typing import List

import numpy as np
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


def box_iou(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """IoU matrix (N, M) of xyxy boxes (N, 4) x (M, 4) -- the documented contract of metrics.py:17-44.
    (The reference body indexes box1[0..3] as if box1 were 4 x N and only works for N == 4; this is the
    (N, 4) semantics its docstring and its caller validate.py:93 need.)  Plain torch: host-side mAP
    bookkeeping after the path, not a hot kernel."""
    a1, a2 = box1[:, None, :2], box1[:, None, 2:4]
    b1, b2 = box2[None, :, :2], box2[None, :, 2:4]
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp(0).prod(2)
    area1 = (box1[:, 2] - box1[:, 0]) * (box1[:, 3] - box1[:, 1])
    area2 = (box2[:, 2] - box2[:, 0]) * (box2[:, 3] - box2[:, 1])
    return inter / (area1[:, None] + area2[None, :] - inter + eps)


def compute_ap(recall, precision):
    """Area under the precision envelope over the points where recall changes (metrics.py:124-149).
    Returns (ap, mpre, mrec)."""
    mrec = np.concatenate(([0.0], np.asarray(recall, dtype=np.float64), [1.0]))
    mpre = np.concatenate(([0.0], np.asarray(precision, dtype=np.float64), [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]  # right-to-left running maximum
    idx = np.nonzero(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1])), mpre, mrec


def ap_per_class(tp, conf, pred_cls, target_cls, plot=False, save_dir=Path(""), names=(), eps=1e-16):
    """Per-class precision / recall / AP / F1 (metrics.py:152-225).  tp [n, n_iou] bool, conf [n],
    pred_cls [n], target_cls [m].  Returns (p, r, ap [classes, n_iou], f1, classes) with p, r, f1 taken at
    the confidence that maximises mean F1 over a 1000-point grid."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes = np.unique(target_cls)
    grid = np.linspace(0, 1, 1000)
    ap = np.zeros((classes.shape[0], tp.shape[1]))
    prec = np.zeros((classes.shape[0], 1000))
    rec = np.zeros((classes.shape[0], 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        n_gt, n_pred = int((target_cls == c).sum()), int(sel.sum())
        if n_gt == 0 or n_pred == 0:
            continue
        tpc = tp[sel].cumsum(0)
        fpc = (1 - tp[sel]).cumsum(0)
        recall_curve = tpc / (n_gt + eps)
        precision_curve = tpc / (tpc + fpc)
        rec[ci] = np.interp(-grid, -conf[sel], recall_curve[:, 0])
        prec[ci] = np.interp(-grid, -conf[sel], precision_curve[:, 0])
        for j in range(tp.shape[1]):
            ap[ci, j] = compute_ap(recall_curve[:, j], precision_curve[:, j])[0]
    f1 = 2 * prec * rec / (prec + rec + eps)
    best = int(f1.mean(0).argmax())
    return prec[:, best], rec[:, best], ap, f1[:, best], classes.astype(int)
