"""Pins oracle/nms_ref.c + oracle/nms.py: against torchvision.ops.nms (the third-party op the
reference calls, metrics.py:442) and against the reference wrapper's committed outputs."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import nms as onms


@pytest.mark.parametrize("name", cases.NMS_CORE_CASES)
def test_core_matches_golden_keep_indices(name, golden_dir):
    boxes, scores, thr = cases.nms_core_case(name)
    keep = onms.nms(boxes, scores, thr)
    g = np.load(os.path.join(golden_dir, f"nms_core_{name}.npz"))["keep"]
    assert keep.dtype == np.int64
    assert np.array_equal(keep, g)  # bit-exact


@pytest.mark.parametrize("name", cases.NMS_CORE_CASES)
def test_core_matches_torchvision_live(name):
    tv = pytest.importorskip("torchvision")
    boxes, scores, thr = cases.nms_core_case(name)
    ref = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
    assert np.array_equal(onms.nms(boxes, scores, thr), ref)


def test_core_empty():
    assert onms.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).shape == (0,)


@pytest.mark.parametrize("name", cases.NMS_WRAPPER_CASES)
def test_wrapper_matches_reference_golden(name, golden_dir):
    pred, kw = cases.nms_wrapper_case(name)
    out = onms.non_max_suppression(pred, **kw)
    g = np.load(os.path.join(golden_dir, f"nms_wrap_{name}.npz"))
    assert len(out) == pred.shape[0]
    for i, o in enumerate(out):
        assert o.shape == g[f"img{i}"].shape, (i, o.shape, g[f"img{i}"].shape)
        assert np.array_equal(o, g[f"img{i}"])  # bit-exact rows


def test_wrapper_fixed_mode_is_class_aware_on_corners():
    pred, _ = cases.nms_wrapper_case("nc10_best")
    out = onms.non_max_suppression(pred[:1], 0.25, 0.45, compat="fixed")[0]
    assert out.shape[1] == 6 and (out[:, 2] >= out[:, 0]).all() and (out[:, 3] >= out[:, 1]).all()
    assert np.all(np.diff(out[:, 4]) <= 0)
