"""Tiled inference of large drone frames and the cross-tile NMS merge (BASELINE config 4).

[NOT IN REFERENCE: the reference has no tile/slice code; SURVEY.md D8 fixes the scheme]
3840x2160 frames are cut into a 4x2 grid of overlapping 1280x1280 tiles, origins
x0 in {0, 853, 1707, 2560}, y0 in {0, 880}.  Tiles shard across ranks round-robin (tile_id % world)
with NO collective in the forward pass; each rank runs forward + per-tile NMS on its tiles, shifts
the kept boxes into frame coordinates, and one all_gather of fixed-size padded detections
([tiles, max_det, 7] fp32 + int32 counts, 8.4 KB per tile) feeds the per-frame merge NMS, which every
rank runs redundantly.  Concatenation order is (tile_id, per-tile keep order), so the result does
not depend on the world size.

The detector and NMS are injected callables: on the GPU they are SkyEyeDetector.forward and
skyeye.utils.nms.batched_nms_padded; the CPU gloo tests pass stand-ins.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

TILE = 1280


def tile_origins(height: int = 2160, width: int = 3840, tile: int = TILE, nx: Optional[int] = None, ny: Optional[int] = None) -> List[Tuple[int, int]]:
    """Row-major list of (y0, x0). For 3840x2160 -> the D8 grid (x0 0/853/1707/2560, y0 0/880)."""
    if height < tile or width < tile:
        raise ValueError(f"frame {height}x{width} smaller than the {tile} tile")
    nx = nx or max(1, -(-(width - tile) // int(tile * 0.69)) + 1) if width > tile else 1
    ny = ny or max(1, -(-(height - tile) // int(tile * 0.69)) + 1) if height > tile else 1
    xs = [round(i * (width - tile) / (nx - 1)) for i in range(nx)] if nx > 1 else [0]
    ys = [round(i * (height - tile) / (ny - 1)) for i in range(ny)] if ny > 1 else [0]
    return [(y, x) for y in ys for x in xs]


def local_tile_ids(n_tiles_total: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_tiles_total, world))


def slice_tiles(frames: torch.Tensor, origins: Sequence[Tuple[int, int]], tile_ids: Sequence[int], tile: int = TILE) -> torch.Tensor:
    """frames [F,3,H,W] -> [len(tile_ids),3,tile,tile]; global tile id = frame * len(origins) + k."""
    T = len(origins)
    out = torch.empty((len(tile_ids), frames.shape[1], tile, tile), dtype=frames.dtype, device=frames.device)
    for i, t in enumerate(tile_ids):
        f, k = divmod(t, T)
        y0, x0 = origins[k]
        out[i].copy_(frames[f, :, y0:y0 + tile, x0:x0 + tile])
    return out


def shift_rows(rows: torch.Tensor, origins: Sequence[Tuple[int, int]], tile_ids: Sequence[int]) -> torch.Tensor:
    """Per-tile NMS rows [n_local, max_det, 7] ([cx,cy,w,h,obj,cls_prob,cls_id]) -> frame coordinates."""
    T = len(origins)
    off = torch.tensor([[origins[t % T][1], origins[t % T][0]] for t in tile_ids], dtype=rows.dtype, device=rows.device)
    out = rows.clone()
    out[:, :, 0:2] += off[:, None, :]
    return out


def gather_tiles(rows: torch.Tensor, counts: torch.Tensor, n_tiles_total: int, rank: int, world: int):
    """all_gather of padded per-tile detections; returns tensors ordered by GLOBAL tile id."""
    import torch.distributed as dist
    n_local_max = -(-n_tiles_total // world)
    md = rows.shape[1]
    pad_rows = torch.zeros((n_local_max, md, 7), dtype=rows.dtype, device=rows.device)
    pad_cnt = torch.zeros(n_local_max, dtype=torch.int32, device=rows.device)
    pad_rows[: rows.shape[0]] = rows
    pad_cnt[: counts.shape[0]] = counts
    if world > 1:
        all_rows = torch.empty((world * n_local_max, md, 7), dtype=rows.dtype, device=rows.device)
        all_cnt = torch.empty(world * n_local_max, dtype=torch.int32, device=rows.device)
        dist.all_gather_into_tensor(all_rows, pad_rows)  # rank-major concatenation (NCCL over NVLink; gloo in CPU tests)
        dist.all_gather_into_tensor(all_cnt, pad_cnt)
        all_rows = all_rows.view(world, n_local_max, md, 7)
        all_cnt = all_cnt.view(world, n_local_max)
    else:
        all_rows, all_cnt = pad_rows[None], pad_cnt[None]
    # rank r holds tiles r, r+world, ... at local index i  ->  global id = i*world + r
    ids = torch.arange(n_tiles_total, device=rows.device)
    g_rows = all_rows[ids % world, ids // world]
    g_cnt = all_cnt[ids % world, ids // world]
    return g_rows, g_cnt


def merge_prediction(g_rows: torch.Tensor, g_cnt: torch.Tensor, n_frames: int, tiles_per_frame: int, nc: int) -> torch.Tensor:
    """Re-express gathered rows as a prediction tensor [F, T*max_det, 5+nc] so the SAME NMS wrapper
    (metrics.py:361-457 semantics) performs the cross-tile merge: columns 0..4 are copied, the class
    probability goes to column 5+cls_id (best-class selection recovers (cls_prob, cls_id) exactly),
    padded rows keep objectness 0 and are dropped by the confidence filter."""
    md = g_rows.shape[1]
    valid = (torch.arange(md, device=g_rows.device)[None, :] < g_cnt[:, None]).to(g_rows.dtype)
    rows = g_rows * valid[:, :, None]
    pred = torch.zeros((rows.shape[0], md, 5 + nc), dtype=rows.dtype, device=rows.device)
    pred[:, :, :5] = rows[:, :, :5]
    if nc > 1:
        pred.scatter_(2, (rows[:, :, 6].long().clamp(0, nc - 1) + 5)[:, :, None], rows[:, :, 5:6])
    elif nc == 1:
        pred[:, :, 5] = 1.0
    return pred.view(n_frames, tiles_per_frame * md, 5 + nc)


def tiled_detect(frames: torch.Tensor, detect: Callable[[torch.Tensor], torch.Tensor], nms_padded: Callable[..., Tuple[torch.Tensor, torch.Tensor]],
                 nc: int, rank: int = 0, world: int = 1, conf: float = 0.25, iou: float = 0.45, max_det: int = 300,
                 origins: Optional[Sequence[Tuple[int, int]]] = None, tile: int = TILE, max_batch: int = 16):
    """Full config-4 step. Returns (rows [F, max_det, 7], counts [F]) in frame coordinates."""
    F = frames.shape[0]
    origins = origins or tile_origins(frames.shape[2], frames.shape[3], tile)
    T = len(origins)
    ids = local_tile_ids(F * T, rank, world)
    rows_l, cnt_l = [], []
    for i in range(0, len(ids), max_batch):
        chunk = ids[i:i + max_batch]
        det = detect(slice_tiles(frames, origins, chunk, tile))
        r, c = nms_padded(det, conf, iou, max_detections=max_det)
        rows_l.append(shift_rows(r, origins, chunk))
        cnt_l.append(c.clone())
    rows = torch.cat(rows_l) if rows_l else torch.zeros((0, max_det, 7), device=frames.device)
    cnt = torch.cat(cnt_l) if cnt_l else torch.zeros(0, dtype=torch.int32, device=frames.device)
    g_rows, g_cnt = gather_tiles(rows, cnt, F * T, rank, world)
    pred = merge_prediction(g_rows, g_cnt, F, T, nc)
    return nms_padded(pred, conf, iou, max_detections=max_det)
