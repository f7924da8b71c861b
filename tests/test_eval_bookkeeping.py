"""Host-side evaluation bookkeeping around the path: box utilities, IoU matching and
AP computation of skyeye.cli.validate / skyeye.utils.metrics, checked against hand-computed cases and --
in the build container -- against the reference's own ``compute_ap`` / ``ap_per_class`` / ``process_batch``."""
import numpy as np
import pytest
import torch

import cases


def _mods():
    from skyeye.cli import validate as V
    from skyeye.utils import general as G
    from skyeye.utils import metrics as M
    return V, G, M


def test_box_conversions_and_scale_boxes():
    _, G, _ = _mods()
    b = torch.tensor([[50.0, 40.0, 20.0, 10.0]])
    xyxy = G.xywh2xyxy(b)
    assert xyxy.tolist() == [[40.0, 35.0, 60.0, 45.0]]
    assert torch.equal(G.xyxy2xywh(xyxy), b)
    # 100x200 original letterboxed into 128x256 (gain 1.28, no padding) -> boxes scale back by 1/1.28 and clip
    boxes = torch.tensor([[12.8, 25.6, 128.0, 64.0], [-5.0, 0.0, 300.0, 200.0]])
    out = G.scale_boxes((128, 256), boxes.clone(), (100, 200))
    assert torch.allclose(out[0], torch.tensor([10.0, 20.0, 100.0, 50.0]), atol=1e-4)
    assert out[1].tolist() == [0.0, 0.0, 200.0, 100.0]
    # with explicit ratio/pad as produced by the loader
    out2 = G.scale_boxes((128, 256), torch.tensor([[20.0, 30.0, 60.0, 70.0]]), (50, 100), ((2.0, 2.0), (28.0, 14.0)))
    assert torch.allclose(out2[0], torch.tensor([-4.0, 8.0, 16.0, 28.0]).clamp(min=0), atol=1e-5)
    assert G.check_img_size(641, s=32) == 672 and G.check_img_size(640, 32) == 640


def test_box_iou_matrix():
    _, _, M = _mods()
    a = torch.tensor([[0.0, 0.0, 10.0, 10.0], [5.0, 5.0, 15.0, 15.0], [20.0, 20.0, 30.0, 30.0]])
    iou = M.box_iou(a, a[:2])
    assert iou.shape == (3, 2)
    assert abs(float(iou[0, 0]) - 1.0) < 1e-6 and abs(float(iou[0, 1]) - 25.0 / 175.0) < 1e-6 and float(iou[2, 0]) == 0.0


def test_compute_ap_hand_cases():
    _, _, M = _mods()
    ap, mpre, mrec = M.compute_ap(np.array([0.5, 1.0]), np.array([1.0, 0.5]))
    assert abs(ap - (0.5 * 1.0 + 0.5 * 0.5)) < 1e-12 and mrec[0] == 0.0 and mrec[-1] == 1.0
    assert abs(M.compute_ap(np.array([1.0]), np.array([1.0]))[0] - 1.0) < 1e-12
    assert M.compute_ap(np.array([0.0]), np.array([0.0]))[0] == 0.0


def _toy_stats(seed, n=400, m=150, nc=4, niou=10):
    rng = cases.rng("ap", seed)
    conf = rng.random(n).astype(np.float64)
    pred_cls = rng.integers(0, nc, n).astype(np.float64)
    target_cls = rng.integers(0, nc, m).astype(np.float64)
    base = rng.random(n) < 0.5
    tp = np.stack([base & (rng.random(n) < 1.0 - 0.08 * j) for j in range(niou)], 1)
    return tp, conf, pred_cls, target_cls


def test_ap_per_class_properties():
    _, _, M = _mods()
    tp, conf, pcls, tcls = _toy_stats(0)
    p, r, ap, f1, cls = M.ap_per_class(tp, conf, pcls, tcls)
    assert ap.shape == (len(cls), 10) and p.shape == r.shape == f1.shape == (len(cls),)
    assert (ap >= 0).all() and (p >= 0).all() and (p <= 1 + 1e-9).all() and (r >= 0).all()  # (toy tp can exceed n_gt)
    # a perfect detector has AP 1 in every class
    tcls2 = np.array([0.0, 0.0, 1.0])
    tp2 = np.ones((3, 10), bool)
    p2, r2, ap2, _, _ = M.ap_per_class(tp2, np.array([0.9, 0.8, 0.7]), np.array([0.0, 0.0, 1.0]), tcls2)
    assert np.allclose(ap2, 1.0, atol=1e-9)


def test_process_batch_matches_greedy_one_to_one():
    V, _, _ = _mods()
    iouv = torch.linspace(0.5, 0.95, 10)
    labels = torch.tensor([[1.0, 0, 0, 10, 10], [2.0, 20, 20, 30, 30]])
    det = torch.tensor([[0.0, 0, 10, 10, 0.9, 1.0],      # exact match of label 0
                        [1.0, 1, 10, 10, 0.8, 1.0],      # second detection of the same label: not counted
                        [20.0, 20, 30, 29, 0.7, 2.0],    # IoU 0.9 with label 1
                        [20.0, 20, 30, 30, 0.6, 3.0]])   # right box, wrong class
    c = V.process_batch(det, labels, iouv)
    assert c[0].all() and not c[1].any() and not c[3].any()
    assert c[2, :9].all() and not bool(c[2, 9])         # 0.9 >= thresholds up to 0.90, < 0.95
    assert V.process_batch(det[:0], labels, iouv).shape == (0, 10)


@pytest.mark.reference
def test_bookkeeping_matches_live_reference():
    """compute_ap / ap_per_class restated here vs the reference's functions (metrics.py:124-225) on seeded stats."""
    from oracle import ref_loader
    _, _, M = _mods()
    ref = ref_loader.load()
    for seed in range(3):
        tp, conf, pcls, tcls = _toy_stats(seed)
        got = M.ap_per_class(tp, conf, pcls, tcls)
        exp = ref.metrics.ap_per_class(tp, conf, pcls, tcls)
        for a, b in zip(got, exp):
            assert np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=1e-12)
    rec, prec = np.sort(cases.rng("r").random(50)), np.sort(cases.rng("p").random(50))[::-1]
    assert abs(M.compute_ap(rec, prec)[0] - ref.metrics.compute_ap(rec, prec)[0]) < 1e-12


def test_folder_loader_letterboxes_and_maps_labels(tmp_path):
    cv2 = pytest.importorskip("cv2")
    V, _, _ = _mods()
    (tmp_path / "images" / "val").mkdir(parents=True)
    (tmp_path / "labels" / "val").mkdir(parents=True)
    rng = cases.rng("folder")
    for i, (h, w) in enumerate([(100, 200), (120, 120), (64, 96)]):
        cv2.imwrite(str(tmp_path / "images" / "val" / f"{i}.png"), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        with open(tmp_path / "labels" / "val" / f"{i}.txt", "w") as f:
            f.write("3 0.5 0.5 0.2 0.4\n1 0.25 0.75 0.1 0.1\n")
    dl = V.FolderLoader(tmp_path / "images" / "val", img_size=128, batch_size=2)
    batches = list(dl)
    assert len(dl) == 2 and len(batches) == 2
    img, targets, paths, shapes = batches[0]
    assert img.dtype == torch.uint8 and img.shape[0] == 2 and img.shape[1] == 3 and img.shape[2] % 32 == 0 and img.shape[3] % 32 == 0
    assert targets.shape == (4, 6) and set(targets[:, 0].tolist()) == {0.0, 1.0}
    # Mixed aspect ratios in one batch: image 0 is 100x200 -> gain 0.64 -> 64x128, image 1 is 120x120 -> 128x128, so
    # the batch frame is 128x128 and image 0 sits in its top half.  Targets are normalised by the BATCH frame (validate() scales
    # them back by it): image 0's centre label lands at pixel (64, 32) of the frame, not at the frame centre.
    H, W = img.shape[2], img.shape[3]
    assert (H, W) == (128, 128)
    t0 = targets[targets[:, 0] == 0][0]
    assert abs(float(t0[2]) * W - 64.0) < 1e-3 and abs(float(t0[3]) * H - 32.0) < 1e-3 and shapes[0][0] == (100, 200)
    assert abs(float(t0[4]) * W - 0.2 * 128) < 1e-3 and abs(float(t0[5]) * H - 0.4 * 64) < 1e-3
    t1 = targets[targets[:, 0] == 1][0]   # image 1 fills the frame: its centre label stays at the centre
    assert abs(float(t1[2]) - 0.5) < 1e-3 and abs(float(t1[3]) - 0.5) < 1e-3
