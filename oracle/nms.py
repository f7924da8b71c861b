"""CPU oracle (TEST INFRASTRUCTURE): NMS core (C, ``nms_ref.c``) + numpy restatement of the
reference wrapper ``non_max_suppression`` (skyeye/utils/metrics.py:361-457), quirks X8 verbatim.

Pinned: C core vs ``torchvision.ops.nms`` (CPU) and wrapper vs the reference function executed
in the build container (fixtures in ``tests/golden/nms_*.npz`` made by ``make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = os.path.join(_BUILD, "libnms_ref.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc the C restatement (no FMA contraction so fp32 results equal torchvision's CPU op)."""
    src = os.path.join(_HERE, "nms_ref.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", _LIB])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.skyeye_oracle_nms.restype = ctypes.c_int64
        _lib.skyeye_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_float, ctypes.c_void_p]
    return _lib


def nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision.ops.nms semantics (call site metrics.py:442): kept original indices, int64,
    descending score, stable on ties."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    n = scores.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    nk = _load().skyeye_oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, float(thr), keep.ctypes.data)
    return keep[:nk].copy()


def non_max_suppression(prediction: np.ndarray, conf_threshold: float = 0.25, iou_threshold: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False,
                        multi_label: bool = False, max_detections: int = 300,
                        compat: str = "reference") -> List[np.ndarray]:
    """Restatement of metrics.py:361-457 on numpy fp32 ``prediction [B, N, 5+nc]``.

    compat="reference" (default) keeps the reference's quirks (X8): boxes are fed to NMS as if
    (cx,cy,w,h) were corners (:438-439), the NMS score is objectness only (:439), rows have 7
    columns [cx,cy,w,h,obj,cls_prob,cls_id] for nc>1 (:413,417) so the class offset uses the class
    PROBABILITY column 5 (:438) and the ``classes`` filter compares column 5 (:426); images with
    no candidates return zeros((0,6)) (:399).  compat="fixed" is what the docstring promises
    (:383): xywh->xyxy, conf = obj*cls, offset = cls_id*4096, rows [x1,y1,x2,y2,conf,cls].
    """
    assert compat in ("reference", "fixed")
    pred = np.asarray(prediction, dtype=np.float32)
    nc = pred.shape[2] - 5
    max_wh, max_nms = np.float32(4096), 30000  # :393
    multi_label = bool(multi_label) and nc > 1  # :396
    thr = np.float32(conf_threshold)
    out: List[np.ndarray] = []
    for x in pred:  # :400
        empty = np.zeros((0, 6), dtype=np.float32)
        x = x[x[:, 4] > thr]  # :391,402
        if not x.shape[0]:
            out.append(empty)
            continue
        if compat == "fixed":
            cls = x[:, 5:] * x[:, 4:5] if nc > 0 else x[:, 4:5]
            xy, wh = x[:, 0:2], x[:, 2:4]
            box = np.concatenate((xy - wh / np.float32(2), xy + wh / np.float32(2)), 1).astype(np.float32)
            if nc > 1 and multi_label:
                i, j = np.nonzero(cls > thr)
                x = np.concatenate((box[i], cls[i, j, None], j[:, None].astype(np.float32)), 1)
            else:
                j = cls.argmax(1) if nc > 0 else np.zeros(x.shape[0], dtype=np.int64)
                conf = cls[np.arange(cls.shape[0]), j if nc > 0 else 0]
                x = np.concatenate((box, conf[:, None], j[:, None].astype(np.float32)), 1)[conf > thr]
            if classes is not None:
                x = x[np.isin(x[:, 5], np.asarray(classes, dtype=np.float32))]
            if not x.shape[0]:
                out.append(empty)
                continue
            if x.shape[0] > max_nms:
                x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
            c = x[:, 5:6] * (np.float32(0) if agnostic else max_wh)
            keep = nms(x[:, :4] + c, x[:, 4], iou_threshold)[:max_detections]
            out.append(x[keep])
            continue
        if nc > 1:
            if multi_label:  # :407-410
                i, j = np.nonzero(x[:, 5:] > thr)
                x = np.concatenate((x[i, :5], x[i, j + 5, None], j[:, None].astype(np.float32)), 1)
            else:  # :411-414
                j = x[:, 5:].argmax(1)
                conf = x[np.arange(x.shape[0]), 5 + j]
                x = np.concatenate((x[:, :5], conf[:, None], j[:, None].astype(np.float32)), 1)[conf > thr]
        else:  # :415-419
            x = np.concatenate((x[:, :4], x[:, 4:5], np.zeros_like(x[:, 4:5])), 1)
        if classes is not None:  # :422-423 (compares column 5)
            x = x[np.isin(x[:, 5], np.asarray(classes, dtype=np.float32))]
        n = x.shape[0]
        if not n:  # :426-428
            out.append(empty)
            continue
        if n > max_nms:  # :431-432 (reference argsort is non-stable; tie-free inputs only)
            x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        c = x[:, 5:6] * (np.float32(0) if agnostic else max_wh)  # :435
        boxes, scores = (x[:, :4] + c).astype(np.float32), x[:, 4]  # :436
        keep = nms(boxes, scores, iou_threshold)[:max_detections]  # :439-441
        out.append(x[keep])  # :452
    return out
