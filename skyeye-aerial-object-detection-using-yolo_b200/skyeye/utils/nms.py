"""NMS entry points bound to libskyeye_b200.so.

``nms`` has torchvision.ops.nms semantics (the op the reference calls, skyeye/utils/metrics.py:442)
and is bit-exact with the CPU op; ``batched_nms_padded`` is the whole reference wrapper
(metrics.py:361-457) as sync-free kernels returning padded rows + counts.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from .. import _native as N

_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index if hasattr(device, "index") else 0)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """boxes [n,4] xyxy fp32, scores [n] fp32 (CUDA) -> kept indices int64, descending score."""
    if not boxes.is_cuda:
        raise RuntimeError("skyeye.utils.nms runs on CUDA only; there is no CPU fallback")
    boxes = boxes.float().contiguous()
    scores = scores.float().contiguous()
    n = boxes.shape[0]
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=boxes.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    ws = _workspace(N.lib().skb_nms_workspace_bytes(n), boxes.device)
    with torch.cuda.device(boxes.device):
        N.check(N.lib().skb_nms_f32(boxes.data_ptr(), scores.data_ptr(), n, float(iou_threshold), keep.data_ptr(), cnt.data_ptr(),
                                    ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "skb_nms_f32")
    return keep[: int(cnt.item())]


def batched_nms_padded(prediction: torch.Tensor, conf_threshold=0.25, iou_threshold=0.45, classes: Optional[Sequence[int]] = None,
                       agnostic=False, multi_label=False, max_detections=300, compat="reference",
                       out: Optional[torch.Tensor] = None, out_count: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """prediction [B,N,5+nc] fp32 CUDA -> (rows [B,max_det,7] fp32, counts [B] int32), no host sync."""
    if not prediction.is_cuda:
        raise RuntimeError("skyeye NMS runs on CUDA only; there is no CPU fallback")
    assert compat in ("reference", "fixed")
    pred = prediction.float().contiguous()
    B, Nb, no = pred.shape
    nc = no - 5
    if out is None:
        out = torch.zeros((B, max_detections, 7), dtype=torch.float32, device=pred.device)
    if out_count is None:
        out_count = torch.zeros(B, dtype=torch.int32, device=pred.device)
    ml = 1 if (multi_label and nc > 1) else 0
    ws = _workspace(N.lib().skb_nms_batched_workspace_bytes(B, Nb, nc, ml), pred.device)
    cls_arr, ncls = None, 0
    if classes is not None:
        ncls = len(classes)
        cls_arr = (ctypes.c_int32 * ncls)(*[int(c) for c in classes])
    with torch.cuda.device(pred.device):
        N.check(N.lib().skb_nms_batched_f32(pred.data_ptr(), B, Nb, nc, float(conf_threshold), float(iou_threshold), cls_arr, ncls,
                                            1 if agnostic else 0, ml, int(max_detections), 0 if compat == "reference" else 1,
                                            out.data_ptr(), out_count.data_ptr(), ws.data_ptr(), ws.numel(),
                                            torch.cuda.current_stream().cuda_stream), "skb_nms_batched_f32")
    return out, out_count
