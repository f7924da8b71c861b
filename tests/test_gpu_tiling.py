"""BASELINE config 4 on the GPU: tiles read in place out of resident frames, per-tile NMS writing shifted rows into the gather
buffer, merge-prediction kernel, per-frame merge NMS -- against (a) the same step with eager glue and (b) the CPU oracle NMS
on the same per-tile detections.  [NOT IN REFERENCE: SURVEY.md D8 / §8e; nearest call site validate.py:234-256]"""
import numpy as np
import pytest
import torch

import cases
from oracle import model as om
from oracle import nms as onms

pytestmark = pytest.mark.gpu

TILE = 256
ORIGINS = [(0, 0), (0, 131), (0, 259), (0, 384), (127, 0), (128, 131), (127, 259), (128, 384)]  # odd origins: unaligned loads


def _model(variant="skyeye_s"):
    from skyeye.core.detector import construct_model
    cfg = om.get_cfg(variant)
    sd = om.make_state_dict(cfg, 0)
    m = construct_model(f"{variant}.yaml")
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _frames(n=2):
    return torch.from_numpy(cases.rng("tiling", "frames").integers(0, 256, (n, 3, 384, 640), dtype=np.uint8)).cuda()


def test_forward_tiles_equals_forward_on_materialised_slices():
    from skyeye.utils import tiling
    m = _model()
    frames = _frames()
    ids = list(range(2 * len(ORIGINS)))
    table = torch.tensor([[t // 8, ORIGINS[t % 8][0], ORIGINS[t % 8][1]] for t in ids], dtype=torch.int32).cuda()
    det_t, raw_t = m.forward_tiles(frames, table, (TILE, TILE))
    det_t, raw_t = det_t.clone(), [r.clone() for r in raw_t]
    det_s, raw_s = m(tiling.slice_tiles(frames, ORIGINS, ids, TILE))
    assert torch.equal(det_t, det_s) and all(torch.equal(a, b) for a, b in zip(raw_t, raw_s))
    ff = frames.float() / 255.0   # fp32 frames take the same path
    det_f, _ = m.forward_tiles(ff, table, (TILE, TILE))
    assert torch.equal(det_f, det_s)


@pytest.mark.parametrize("compat", ["fixed", "reference"])
def test_native_tiled_step_equals_eager_glue_and_the_oracle_nms(compat):
    from skyeye.utils import tiling
    from skyeye.utils.nms import batched_nms_padded
    m = _model()
    frames = _frames()
    conf, iou, md = 0.3, 0.5, 40
    td = tiling.TiledDetector(m, 2, (384, 640), conf=conf, iou=iou, max_det=md, compat=compat, tile=TILE, max_batch=5, origins=ORIGINS)
    rows, cnt = td(frames)
    rows, cnt = rows.clone(), cnt.clone()
    rows2, cnt2 = td(frames)                      # second step (other slot, pipelined path)
    assert torch.equal(rows, rows2) and torch.equal(cnt, cnt2)
    assert int(cnt.sum()) > 10
    # (a) eager glue around the same kernels
    e_rows, e_cnt = tiling.tiled_detect(frames, lambda t: m(t)[0], batched_nms_padded, nc=10, conf=conf, iou=iou, max_det=md,
                                        origins=ORIGINS, tile=TILE, max_batch=5, compat=compat)
    assert torch.equal(cnt, e_cnt)
    for f in range(2):
        assert torch.equal(rows[f, : int(cnt[f])], e_rows[f, : int(cnt[f])])
    # (b) CPU oracle NMS (per tile and merge) on the GPU's own per-tile detections
    def o_nms(pred, c, i, max_detections=300, compat="reference"):
        out = onms.non_max_suppression(pred.cpu().numpy(), c, i, max_detections=max_detections, compat=compat)
        r = torch.zeros((pred.shape[0], max_detections, 7))
        n = torch.zeros(pred.shape[0], dtype=torch.int32)
        for b, o in enumerate(out):
            r[b, : o.shape[0], : o.shape[1]] = torch.from_numpy(o)
            n[b] = o.shape[0]
        return r, n
    o_rows, o_cnt = tiling.tiled_detect(frames, lambda t: m(t)[0].cpu(), o_nms, nc=10, conf=conf, iou=iou, max_det=md, origins=ORIGINS,
                                        tile=TILE, max_batch=5, compat=compat)
    assert torch.equal(cnt.cpu(), o_cnt)
    for f in range(2):
        assert torch.equal(rows[f, : int(cnt[f])].cpu(), o_rows[f, : int(cnt[f])])
    assert float(rows[0, int(cnt[0]):].abs().sum()) == 0.0   # padding rows are zero


def test_straddling_object_is_one_box_after_the_gpu_merge():
    from skyeye.utils import tiling
    from skyeye.utils.nms import batched_nms_padded
    origins = [(0, 0), (0, 853)]
    frames = torch.zeros((1, 3, 1280, 2133), device="cuda")
    cx, cy, w, h = 1000.0, 600.0, 80.0, 60.0

    def detect(tiles):
        det = torch.zeros((tiles.shape[0], 4, 15), device="cuda")
        for i, (y0, x0) in enumerate(origins[: tiles.shape[0]]):
            det[i, 0, :5] = torch.tensor([cx - x0, cy - y0, w, h, 0.9 - 0.05 * i])
            det[i, 0, 5 + 3] = 0.8
        return det

    rows, cnt = tiling.tiled_detect(frames, detect, batched_nms_padded, nc=10, origins=origins, compat="fixed")
    assert int(cnt[0]) == 1
    assert torch.allclose(rows[0, 0, :6].cpu(), torch.tensor([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, 0.9 * 0.8, 3.0]), atol=1e-4)
