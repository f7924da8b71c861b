"""Tiled inference of large drone frames and the cross-tile NMS merge (BASELINE config 4).

[NOT IN REFERENCE: the reference has no tile/slice code; SURVEY.md D8 fixes the scheme]
3840x2160 frames are cut into a 4x2 grid of overlapping 1280x1280 tiles, origins
x0 in {0, 853, 1707, 2560}, y0 in {0, 880}.  Tiles shard across ranks round-robin (tile_id % world)
with NO collective in the forward pass; each rank runs forward + per-tile NMS on its tiles, shifts
the kept boxes into frame coordinates, and one all_gather of fixed-size padded detections feeds the
per-frame merge NMS, which every rank runs redundantly.  Concatenation order is (tile_id, per-tile
keep order), so the result does not depend on the world size.

Row semantics (``compat``): the merge can only suppress the duplicates that the 31 % tile overlap
produces if boxes are corner boxes, so the default is ``"fixed"`` -- rows ``[x1,y1,x2,y2,conf,cls]``,
class-aware NMS on corners, what the reference wrapper's docstring promises (metrics.py:383).
``"reference"`` keeps the wrapper's actual arithmetic (quirk X8: ``(cx,cy,w,h)`` read as corners,
rows ``[cx,cy,w,h,obj,cls_prob,cls_id]``); in frame coordinates nothing overlaps under that reading,
so the merge degenerates to top-max_det by objectness -- kept for parity experiments only.

Two implementations of the same step:
  * ``tiled_detect``  : generic, injected detector / NMS callables, eager torch glue (used by the CPU
    gloo tests with stand-ins and as the checker of the native path);
  * ``TiledDetector`` : the B200 path.  Tiles are read in place out of the resident frames through a
    (frame, y0, x0) table consumed by the first kernel (no tile batch in memory), the per-tile NMS
    kernel adds the tile origin and writes zero-padded rows + count straight into the all_gather
    send buffer, ONE all_gather moves rows and counts together, a kernel builds the merge
    prediction, and gather + merge of step t run on a second stream under the forward of step t+1.
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


def shift_rows(rows: torch.Tensor, origins: Sequence[Tuple[int, int]], tile_ids: Sequence[int], compat: str = "fixed") -> torch.Tensor:
    """Per-tile NMS rows [n_local, max_det, 7] -> frame coordinates: (x0, y0) is added to columns 0,1 of reference rows
    ([cx,cy,w,h,...]) and to columns 0..3 of fixed rows ([x1,y1,x2,y2,...])."""
    T = len(origins)
    off = torch.tensor([[origins[t % T][1], origins[t % T][0]] for t in tile_ids], dtype=rows.dtype, device=rows.device)
    out = rows.clone()
    out[:, :, 0:2] += off[:, None, :]
    if compat == "fixed":
        out[:, :, 2:4] += off[:, None, :]
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


def merge_prediction(g_rows: torch.Tensor, g_cnt: torch.Tensor, n_frames: int, tiles_per_frame: int, nc: int, compat: str = "reference") -> torch.Tensor:
    """Re-express gathered rows as a prediction tensor [F, T*max_det, 5+nc] so the SAME NMS wrapper
    (metrics.py:361-457 semantics) performs the cross-tile merge.  Reference rows: columns 0..4 are copied, the class
    probability goes to column 5+cls_id (best-class selection recovers (cls_prob, cls_id) exactly).  Fixed rows
    [x1,y1,x2,y2,conf,cls]: centre form, objectness = conf, class probability 1 (conf * 1 is exact).  Padded rows keep
    objectness 0 and are dropped by the confidence filter.  (skb_tile_merge_pred_f32 is the same arithmetic as a kernel.)"""
    md = g_rows.shape[1]
    valid = (torch.arange(md, device=g_rows.device)[None, :] < g_cnt[:, None]).to(g_rows.dtype)
    rows = g_rows * valid[:, :, None]
    pred = torch.zeros((rows.shape[0], md, 5 + nc), dtype=rows.dtype, device=rows.device)
    if compat == "fixed":
        pred[:, :, 0] = (rows[:, :, 0] + rows[:, :, 2]) * 0.5
        pred[:, :, 1] = (rows[:, :, 1] + rows[:, :, 3]) * 0.5
        pred[:, :, 2] = rows[:, :, 2] - rows[:, :, 0]
        pred[:, :, 3] = rows[:, :, 3] - rows[:, :, 1]
        pred[:, :, 4] = rows[:, :, 4]
        pred.scatter_(2, (rows[:, :, 5].long().clamp(0, nc - 1) + 5)[:, :, None], valid[:, :, None])
    else:
        pred[:, :, :5] = rows[:, :, :5]
        if nc > 1:
            pred.scatter_(2, (rows[:, :, 6].long().clamp(0, nc - 1) + 5)[:, :, None], rows[:, :, 5:6])
        elif nc == 1:
            pred[:, :, 5] = 1.0
    return pred.view(n_frames, tiles_per_frame * md, 5 + nc)


def tiled_detect(frames: torch.Tensor, detect: Callable[[torch.Tensor], torch.Tensor], nms_padded: Callable[..., Tuple[torch.Tensor, torch.Tensor]],
                 nc: int, rank: int = 0, world: int = 1, conf: float = 0.25, iou: float = 0.45, max_det: int = 300,
                 origins: Optional[Sequence[Tuple[int, int]]] = None, tile: int = TILE, max_batch: int = 16, compat: str = "fixed"):
    """Full config-4 step with eager glue. ``nms_padded(pred, conf, iou, max_detections=, compat=)`` -> (rows, counts).
    Returns (rows [F, max_det, 7], counts [F]) in frame coordinates."""
    F = frames.shape[0]
    origins = origins or tile_origins(frames.shape[2], frames.shape[3], tile)
    T = len(origins)
    ids = local_tile_ids(F * T, rank, world)
    rows_l, cnt_l = [], []
    for i in range(0, len(ids), max_batch):
        chunk = ids[i:i + max_batch]
        det = detect(slice_tiles(frames, origins, chunk, tile))
        r, c = nms_padded(det, conf, iou, max_detections=max_det, compat=compat)
        rows_l.append(shift_rows(r, origins, chunk, compat))
        cnt_l.append(c.clone())
    rows = torch.cat(rows_l) if rows_l else torch.zeros((0, max_det, 7), device=frames.device)
    cnt = torch.cat(cnt_l) if cnt_l else torch.zeros(0, dtype=torch.int32, device=frames.device)
    g_rows, g_cnt = gather_tiles(rows, cnt, F * T, rank, world)
    pred = merge_prediction(g_rows, g_cnt, F, T, nc, compat)
    return nms_padded(pred, conf, iou, max_detections=max_det, compat=compat)


class TiledDetector:
    """The B200 config-4 step (see the module docstring).  ``step(frames)`` enqueues forward + per-tile NMS on the current
    stream and gather + merge on a side stream and returns the slot holding the result; ``result(slot)`` waits for it.
    Consecutive ``step`` calls pipeline: the merge of step t overlaps the forward of step t+1 (two result slots)."""

    def __init__(self, model, n_frames: int, frame_hw: Tuple[int, int], rank: int = 0, world: int = 1, conf: float = 0.25,
                 iou: float = 0.45, max_det: int = 300, compat: str = "fixed", tile: int = TILE, max_batch: int = 16,
                 origins: Optional[Sequence[Tuple[int, int]]] = None, device=None, overlap: bool = True):
        from .. import _native as N
        assert compat in ("fixed", "reference")
        self.N, self.model = N, model
        self.dev = torch.device(device) if device is not None else next(model.parameters()).device
        self.F, self.rank, self.world = n_frames, rank, world
        self.conf, self.iou, self.max_det, self.compat, self.tile, self.max_batch = conf, iou, max_det, 1 if compat == "fixed" else 0, tile, max_batch
        self.nc = int(model.cfg["nc"])
        self.origins = list(origins) if origins is not None else tile_origins(frame_hw[0], frame_hw[1], tile)
        self.T = len(self.origins)
        self.n_tiles = n_frames * self.T
        ids = local_tile_ids(self.n_tiles, rank, world)
        self.n_local, self.n_local_max = len(ids), -(-self.n_tiles // world)
        dev = self.dev
        tab = [[t // self.T, self.origins[t % self.T][0], self.origins[t % self.T][1]] for t in ids]
        self.table = torch.tensor(tab, dtype=torch.int32).reshape(-1, 3).to(dev)              # (frame, y0, x0) per local tile
        self.xy = torch.tensor([[r[2], r[1]] for r in tab], dtype=torch.int32).reshape(-1, 2).to(dev)  # (x0, y0)
        R = max_det + 1
        self.send = [torch.zeros((self.n_local_max, R, 7), dtype=torch.float32, device=dev) for _ in range(2)]
        self.gath = [torch.zeros((world, self.n_local_max, R, 7), dtype=torch.float32, device=dev) if world > 1 else None for _ in range(2)]
        self.tile_cnt = torch.zeros(self.n_local_max, dtype=torch.int32, device=dev)
        self.pred = torch.zeros((n_frames, self.T * max_det, 5 + self.nc), dtype=torch.float32, device=dev)
        self.rows = [torch.zeros((n_frames, max_det, 7), dtype=torch.float32, device=dev) for _ in range(2)]
        self.cnt = [torch.zeros(n_frames, dtype=torch.int32, device=dev) for _ in range(2)]
        self.ws_tile, self.ws_merge = None, torch.empty(int(N.lib().skb_nms_batched_workspace_bytes(n_frames, self.T * max_det, self.nc, 0)) + 256,
                                                        dtype=torch.uint8, device=dev)
        self.side = torch.cuda.Stream(device=dev) if overlap else None
        self.ev_tiles = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.ev_merge = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2)]
        self.steps = 0

    def _tile_nms(self, det: torch.Tensor, i0: int, n: int, slot: int, stream: int):
        N = self.N
        B, Nb, no = det.shape
        need = int(N.lib().skb_nms_batched_workspace_bytes(B, Nb, no - 5, 0))
        if self.ws_tile is None or self.ws_tile.numel() < need:
            self.ws_tile = torch.empty(need + 256, dtype=torch.uint8, device=self.dev)
        N.check(N.lib().skb_nms_batched_tiles_f32(det.data_ptr(), B, Nb, no - 5, self.conf, self.iou, 0, 0, self.max_det, self.compat,
                                                  self.xy[i0:i0 + n].data_ptr(), self.send[slot][i0:i0 + n].data_ptr(),
                                                  self.tile_cnt[i0:i0 + n].data_ptr(), self.ws_tile.data_ptr(), self.ws_tile.numel(), stream),
                "skb_nms_batched_tiles_f32")

    def _merge(self, slot: int, stream: int):
        N = self.N
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.gath[slot], self.send[slot])  # ONE collective: rows and counts travel together
            src = self.gath[slot]
        else:
            src = self.send[slot]
        N.check(N.lib().skb_tile_merge_pred_f32(src.data_ptr(), self.world, self.n_local_max, self.F, self.T, self.max_det, self.nc, self.compat,
                                                self.pred.data_ptr(), stream), "skb_tile_merge_pred_f32")
        N.check(N.lib().skb_nms_batched_f32(self.pred.data_ptr(), self.F, self.T * self.max_det, self.nc, self.conf, self.iou, None, 0, 0, 0,
                                            self.max_det, self.compat, self.rows[slot].data_ptr(), self.cnt[slot].data_ptr(),
                                            self.ws_merge.data_ptr(), self.ws_merge.numel(), stream), "skb_nms_batched_f32")

    @torch.no_grad()
    def step(self, frames: torch.Tensor) -> int:
        slot = self.steps & 1
        with torch.cuda.device(self.dev):
            main = torch.cuda.current_stream()
            if self.steps >= 2:
                main.wait_event(self.ev_done[slot])  # the merge that last read send[slot] has finished
            for i0 in range(0, self.n_local, self.max_batch):
                n = min(self.max_batch, self.n_local - i0)
                det, _ = self.model.forward_tiles(frames, self.table[i0:i0 + n], (self.tile, self.tile))
                self._tile_nms(det, i0, n, slot, main.cuda_stream)
            self.ev_tiles[slot].record(main)
            st = self.side if self.side is not None else main
            with torch.cuda.stream(st):
                st.wait_event(self.ev_tiles[slot])
                self.ev_merge[slot][0].record(st)
                self._merge(slot, st.cuda_stream)
                self.ev_merge[slot][1].record(st)
                self.ev_done[slot].record(st)
        self.steps += 1
        return slot

    def result(self, slot: int):
        """(rows [F, max_det, 7], counts [F]) of the step that returned ``slot``; waits for its merge."""
        self.ev_done[slot].synchronize()
        return self.rows[slot], self.cnt[slot]

    def merge_ms(self, slot: int) -> float:
        a, b = self.ev_merge[slot]
        b.synchronize()
        return a.elapsed_time(b)

    def __call__(self, frames: torch.Tensor):
        return self.result(self.step(frames))
