"""NMS parity through the C ABI: keep indices / output rows must be BIT-EXACT against the CPU oracle
(oracle/nms_ref.c + oracle/nms.py) and against the committed reference outputs (tests/golden)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import nms as onms

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", cases.NMS_CORE_CASES)
def test_core_keep_indices_bit_exact(name, golden_dir):
    from skyeye.utils.nms import nms
    boxes, scores, thr = cases.nms_core_case(name)
    keep = nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr).cpu().numpy()
    assert keep.dtype == np.int64
    assert np.array_equal(keep, onms.nms(boxes, scores, thr))
    assert np.array_equal(keep, np.load(os.path.join(golden_dir, f"nms_core_{name}.npz"))["keep"])


def test_core_empty_and_30000():
    from skyeye.utils.nms import nms
    assert nms(torch.zeros((0, 4), device="cuda"), torch.zeros((0,), device="cuda"), 0.5).numel() == 0
    g = cases.rng("nms30000")
    n = 30000
    c = g.random((n, 2), dtype=np.float32) * 1280
    wh = np.exp(g.uniform(np.log(4), np.log(64), (n, 2))).astype(np.float32)
    b = np.concatenate((c - wh / 2, c + wh / 2), 1).astype(np.float32)
    s = g.permutation(np.linspace(0.002, 0.999, n)).astype(np.float32)
    keep = nms(torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda(), 0.6).cpu().numpy()
    assert np.array_equal(keep, onms.nms(b, s, 0.6))


@pytest.mark.parametrize("name", cases.NMS_WRAPPER_CASES)
def test_wrapper_rows_bit_exact_vs_reference_golden(name, golden_dir):
    from skyeye.utils.metrics import non_max_suppression
    pred, kw = cases.nms_wrapper_case(name)
    out = non_max_suppression(torch.from_numpy(pred).cuda(), **kw)
    g = np.load(os.path.join(golden_dir, f"nms_wrap_{name}.npz"))
    oracle = onms.non_max_suppression(pred, **kw)
    for i, o in enumerate(out):
        o = o.cpu().numpy()
        assert o.shape == g[f"img{i}"].shape, (i, o.shape, g[f"img{i}"].shape)
        assert np.array_equal(o, g[f"img{i}"])
        assert np.array_equal(o, oracle[i])


@pytest.mark.parametrize("kw", [dict(), dict(multi_label=True, conf_threshold=0.3), dict(agnostic=True)])
def test_wrapper_fixed_mode_matches_oracle(kw):
    from skyeye.utils.metrics import non_max_suppression
    pred, _ = cases.nms_wrapper_case("nc10_best")
    out = non_max_suppression(torch.from_numpy(pred).cuda(), compat="fixed", **kw)
    ref = onms.non_max_suppression(pred, compat="fixed", **kw)
    for a, b in zip(out, ref):
        assert np.array_equal(a.cpu().numpy(), b)


def test_wrapper_stress_config5_sample():
    """BASELINE config 5 shape class (50k candidates/image, 10 classes), 4 images checked bit-exact."""
    from skyeye.utils.metrics import non_max_suppression
    g = cases.rng("stress")
    B, N = 4, 50000
    p = np.empty((B, N, 15), dtype=np.float32)
    p[..., 0:2] = g.random((B, N, 2), dtype=np.float32) * 1280
    p[..., 2:4] = np.exp(g.uniform(np.log(4), np.log(64), (B, N, 2))).astype(np.float32)
    for b in range(B):
        p[b, :, 4] = g.permutation(np.linspace(0.002, 0.999, N)).astype(np.float32)
    p[..., 5:] = g.random((B, N, 10), dtype=np.float32)
    for kw in (dict(conf_threshold=0.001, iou_threshold=0.6), dict(conf_threshold=0.25, iou_threshold=0.45)):
        out = non_max_suppression(torch.from_numpy(p).cuda(), **kw)
        ref = onms.non_max_suppression(p, **kw)
        for a, b in zip(out, ref):
            assert np.array_equal(a.cpu().numpy(), b)


def test_wrapper_dense_clusters_many_chunks_bit_exact():
    """Heavily overlapping candidates: hundreds of 512-candidate chunks are walked before max_det boxes are kept, so the
    kept-list kernel's phase A (against earlier chunks), its pairwise bit masks and the greedy resolution are all exercised
    with real suppression (the stress case above keeps almost every candidate of its first chunk)."""
    from skyeye.utils.metrics import non_max_suppression
    g = cases.rng("dense-clusters")
    B, N, nc = 3, 20000, 4
    p = np.empty((B, N, 5 + nc), dtype=np.float32)
    centers = g.random((B, 60, 2), dtype=np.float32) * 1200 + 40
    which = g.integers(0, 60, (B, N))
    for b in range(B):
        p[b, :, 0:2] = centers[b, which[b]] + g.normal(0, 3.0, (N, 2)).astype(np.float32)
    p[..., 2:4] = (30 + g.random((B, N, 2), dtype=np.float32) * 10)
    for b in range(B):
        p[b, :, 4] = g.permutation(np.linspace(0.3, 0.999, N)).astype(np.float32)
    p[..., 5:] = g.random((B, N, nc), dtype=np.float32) * 0.5 + 0.5
    for kw in (dict(conf_threshold=0.25, iou_threshold=0.45), dict(conf_threshold=0.25, iou_threshold=0.3, agnostic=True),
               dict(conf_threshold=0.25, iou_threshold=0.6, max_detections=1000)):
        # compat="fixed": corner boxes (the reference-compat rows treat (cx, cy, w, h) as corners, where nothing overlaps)
        out = non_max_suppression(torch.from_numpy(p).cuda(), compat="fixed", **kw)
        ref = onms.non_max_suppression(p, compat="fixed", **kw)
        for a, r in zip(out, ref):
            assert a.shape[0] == r.shape[0] and a.shape[0] > 30
            assert np.array_equal(a.cpu().numpy(), r)
