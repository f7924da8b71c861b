"""Host logic of tiled multi-GPU inference (BASELINE config 4), CPU only: tile geometry, round-robin
sharding, gloo all_gather (world_size 2) and the cross-tile merge must give the SAME detections for
every world size.  The detector is a deterministic stand-in and NMS is the CPU oracle (the CUDA
kernels are exercised by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nms as onms
from skyeye.utils import tiling


def test_d8_tile_grid_for_4k_frames():
    o = tiling.tile_origins(2160, 3840)
    assert len(o) == 8
    assert sorted({x for _, x in o}) == [0, 853, 1707, 2560]
    assert sorted({y for y, _ in o}) == [0, 880]
    assert all(y + 1280 <= 2160 and x + 1280 <= 3840 for y, x in o)
    assert tiling.tile_origins(1280, 1280) == [(0, 0)]


def test_round_robin_covers_every_tile_once():
    for world in (1, 2, 4, 8):
        got = sorted(t for r in range(world) for t in tiling.local_tile_ids(24, r, world))
        assert got == list(range(24))


def _fake_detect(tiles):
    """Deterministic per-tile stand-in (independent of batch composition): boxes derived from the
    tile's own pixels -> [n, 50, 15]."""
    n = tiles.shape[0]
    det = torch.zeros((n, 50, 15))
    for i in range(n):
        g = torch.Generator().manual_seed(int(tiles[i, :, ::64, ::64].double().sum().item() * 1e6) % (2 ** 31))
        det[i, :, 0:2] = torch.rand((50, 2), generator=g) * 1280
        det[i, :, 2:4] = torch.rand((50, 2), generator=g) * 200 + 20
        det[i, :, 4] = torch.rand(50, generator=g)
        det[i, :, 5:] = torch.rand((50, 10), generator=g)
    return det


def _oracle_nms_padded(pred, conf, iou, max_detections=300, compat="reference"):
    out = onms.non_max_suppression(pred.numpy(), conf, iou, max_detections=max_detections, compat=compat)
    rows = torch.zeros((pred.shape[0], max_detections, 7))
    cnt = torch.zeros(pred.shape[0], dtype=torch.int32)
    for b, o in enumerate(out):
        rows[b, : o.shape[0], : o.shape[1]] = torch.from_numpy(o)
        cnt[b] = o.shape[0]
    return rows, cnt


def _frames():
    g = torch.Generator().manual_seed(5)
    return torch.rand((2, 3, 2160, 3840), generator=g)


def _run(rank, world, port, q):
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for compat in ("fixed", "reference"):
        rows, cnt = tiling.tiled_detect(_frames(), _fake_detect, _oracle_nms_padded, nc=10, rank=rank, world=world, max_batch=3, compat=compat)
        out += [rows.numpy(), cnt.numpy()]
    q.put((rank, *out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_tiled_detect_is_world_size_invariant_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    _run(0, 1, 0, q)
    _, *one = q.get()
    assert one[1].sum() > 0 and one[3].sum() > 0
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    for _, *two in res:
        for a, b in zip(one, two):
            assert np.array_equal(a, b)  # identical detections on every rank, independent of world size, in both row semantics


def test_object_straddling_two_tiles_is_merged_into_one_box():
    """One object in the overlap band of two horizontally adjacent tiles is detected by both (in tile-local coordinates);
    the cross-tile merge must return ONE box in frame coordinates (compat="fixed": corner boxes).  Under the reference
    wrapper's reading of (cx, cy, w, h) as corners nothing overlaps in frame coordinates and both copies survive."""
    origins = [(0, 0), (0, 853)]
    frames = torch.zeros((1, 3, 1280, 2133))
    cx, cy, w, h = 1000.0, 600.0, 80.0, 60.0          # frame coordinates: inside both tiles (853 <= x < 1280)

    def detect(tiles):
        det = torch.zeros((tiles.shape[0], 4, 15))
        for i, (y0, x0) in enumerate(origins[: tiles.shape[0]]):
            det[i, 0, :5] = torch.tensor([cx - x0, cy - y0, w, h, 0.9 - 0.05 * i])
            det[i, 0, 5 + 3] = 0.8
        return det

    rows, cnt = tiling.tiled_detect(frames, detect, _oracle_nms_padded, nc=10, origins=origins, compat="fixed")
    assert int(cnt[0]) == 1
    assert torch.allclose(rows[0, 0, :6], torch.tensor([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, 0.9 * 0.8, 3.0]), atol=1e-4)
    rows_r, cnt_r = tiling.tiled_detect(frames, detect, _oracle_nms_padded, nc=10, origins=origins, compat="reference")
    assert int(cnt_r[0]) == 2   # the quirk: duplicates are not merged


def test_merge_prediction_roundtrip_recovers_rows():
    rows = torch.zeros((4, 6, 7))
    rows[..., 0:2] = torch.rand(4, 6, 2) * 100
    rows[..., 2:4] = torch.rand(4, 6, 2) * 10 + 1
    rows[..., 4] = torch.rand(4, 6) * 0.5 + 0.4
    rows[..., 5] = torch.rand(4, 6) * 0.5 + 0.4
    rows[..., 6] = torch.randint(0, 10, (4, 6)).float()
    cnt = torch.tensor([6, 3, 0, 1], dtype=torch.int32)
    pred = tiling.merge_prediction(rows, cnt, n_frames=2, tiles_per_frame=2, nc=10)
    assert pred.shape == (2, 12, 15)
    assert float(pred[0, 9:, 4].abs().sum()) == 0  # padded rows of tile 1 carry objectness 0
    out = onms.non_max_suppression(pred.numpy(), 0.25, 1.0, max_detections=300)  # iou thr 1.0: nothing suppressed
    got = out[0][np.argsort(-out[0][:, 4], kind="stable")]
    valid = torch.cat([rows[0, :6], rows[1, :3]]).numpy()
    want = valid[np.argsort(-valid[:, 4], kind="stable")]
    assert np.array_equal(got, want)
