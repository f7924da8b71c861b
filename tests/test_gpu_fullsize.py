"""Parity at BASELINE.json's FULL sizes through size-independent properties (the CPU oracle cannot run these sizes in
seconds: one 1280x1280 skyeye_l image takes it ~3 s and the P3 attention matrix alone is 10.5 GB in the reference).

  * attention at N = 25600 (config 3, level P3): softmax rows sum to one (V = 1 gives exactly 1), linearity in V,
    permutation invariance over the key/value order, agreement of a 128-query slab with a direct fp64 evaluation;
  * the whole skyeye_l network at 1280x1280: an image's logits do not depend on its batch position or on the batch size
    (bit for bit: no kernel on the path reduces across images), two runs are bit-identical;
  * the NMS wrapper at config 5 (50 000 candidates, 10 classes): kept rows are sorted by score, pairwise IoU within a class
    is <= the threshold, the class/score filter holds, and NMS is idempotent on its own output.
"""
import math

import numpy as np
import pytest
import torch

import cases
from gpu_util import bf16r, randn, rel_err

pytestmark = pytest.mark.gpu


def _attn(qkv_f32, b, h, w, heads):
    from skyeye import engine as E
    C = heads * 64
    qv = E.View(qkv_f32.view(b, h, w, 3 * C).to(torch.bfloat16).cuda().contiguous())
    o = E.new_buffer(b, h, w, C)
    o.t.zero_()
    E.flash_attn(qv, o, heads, 1.0 / 8.0)
    torch.cuda.synchronize()
    return o.torch().reshape(b, h * w, C).float().cpu()


def test_attention_full_p3_rows_sum_to_one_and_is_linear_in_v():
    b, h, w, heads = 1, 160, 160, 4  # N = 25600 tokens, the P3 level of skyeye_l at 1280x1280
    C, N = heads * 64, h * w
    qkv = bf16r(randn(("full", "p3"), (b, N, 3 * C)))
    ones = qkv.clone()
    ones[..., 2 * C:] = 1.0
    out1 = _attn(ones, b, h, w, heads)
    assert torch.equal(out1, torch.ones_like(out1)), "softmax weights do not sum to one"  # numerator and denominator are the same MMA sums
    v1 = qkv[..., 2 * C:].clone()
    v2 = bf16r(randn(("full", "p3", "v2"), (b, N, C)))
    a, c = 0.5, 2.0  # exact in bf16
    mix = qkv.clone()
    mix[..., 2 * C:] = bf16r(a * v1 + c * v2)
    o1 = _attn(qkv, b, h, w, heads)
    q2 = qkv.clone()
    q2[..., 2 * C:] = v2
    o2 = _attn(q2, b, h, w, heads)
    om_ = _attn(mix, b, h, w, heads)
    # bf16 rounding of the mixed V and of the three outputs: <= 1e-2 of the output scale
    assert rel_err(om_, a * o1 + c * o2) < 1e-2


def test_attention_full_p3_key_permutation_invariance_and_fp64_slab():
    b, h, w, heads = 1, 160, 160, 4
    C, N = heads * 64, h * w
    qkv = bf16r(randn(("full", "p3", "perm"), (b, N, 3 * C)))
    o = _attn(qkv, b, h, w, heads)
    perm = torch.from_numpy(cases.rng("full", "perm").permutation(N))
    shuf = qkv.clone()
    shuf[:, :, C:] = qkv[:, perm, C:]  # same key/value SET in another order: different tiles, different summation order
    o_p = _attn(shuf, b, h, w, heads)
    assert rel_err(o_p, o) < 1e-2
    # 128 queries of head 1 against all 25600 keys in fp64
    rows = slice(3 * 128, 4 * 128)
    hd = slice(64, 128)
    q = qkv[0, rows, hd].double()
    k = qkv[0, :, C + 64:C + 128].double()
    v = qkv[0, :, 2 * C + 64:2 * C + 128].double()
    ref = torch.softmax(q @ k.T / 8.0, dim=-1) @ v
    assert rel_err(o[0, rows, hd], ref.float()) < 1e-2


def _build_l():
    from oracle import model as om
    from skyeye.core.detector import construct_model
    cfg = om.get_cfg("skyeye_l")
    sd = om.make_calibrated_state_dict(cfg, 0)   # the benchmark's weights (activations O(1): attention / CLA see real logits)
    m = construct_model("skyeye_l.yaml")
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def test_skyeye_l_1280_is_batch_invariant_and_deterministic():
    m = _build_l()
    x = torch.from_numpy(cases.rng("full", "img").integers(0, 256, (3, 3, 1280, 1280), dtype=np.uint8)).cuda()
    det3, raw3 = m(x)
    det3, raw3 = det3.clone(), [r.clone() for r in raw3]
    det3b, raw3b = m(x)
    assert torch.equal(det3, det3b) and all(torch.equal(a, b) for a, b in zip(raw3, raw3b)), "two runs differ"
    assert det3.shape == (3, 100800, 15) and bool(torch.isfinite(det3).all())
    for i in (2, 0):
        det1, raw1 = m(x[i:i + 1])
        for lvl, (a, b) in enumerate(zip(raw1, raw3)):
            assert torch.equal(a[0], b[i]), f"image {i}, level {lvl}: logits depend on the batch"
        assert torch.equal(det1[0], det3[i])


def _iou_xyxy(a, b):
    iw = (np.minimum(a[2], b[:, 2]) - np.maximum(a[0], b[:, 0])).clip(0)
    ih = (np.minimum(a[3], b[:, 3]) - np.maximum(a[1], b[:, 1])).clip(0)
    inter = iw * ih
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter)


def test_nms_config5_full_batch_properties():
    """pred [64, 50000, 15] through the wrapper in compat="fixed" (the documented contract: rows [x1, y1, x2, y2, conf, cls]
    with conf = obj * cls_prob, class-aware NMS on corner boxes).  The reference-compat rows are checked bit for bit against
    the oracle on a sample in test_gpu_nms.py; here the WHOLE batch is checked through properties."""
    from skyeye.utils.metrics import non_max_suppression
    g = cases.rng("full", "stress")
    B, N = 64, 50000
    p = np.empty((B, N, 15), dtype=np.float32)
    p[..., 0:2] = g.random((B, N, 2), dtype=np.float32) * 1280
    p[..., 2:4] = np.exp(g.uniform(np.log(4), np.log(64), (B, N, 2))).astype(np.float32)
    lin = np.linspace(0.002, 0.999, N).astype(np.float32)
    for b in range(B):
        p[b, :, 4] = g.permutation(lin)
    p[..., 5:] = g.random((B, N, 10), dtype=np.float32)
    conf, iou, max_det = 0.25, 0.45, 300
    out = non_max_suppression(torch.from_numpy(p).cuda(), conf, iou, max_detections=max_det, compat="fixed")
    assert len(out) == B
    for b in range(B):
        r = out[b].cpu().numpy()
        assert r.shape == (max_det, 6)
        assert np.all(r[:-1, 4] >= r[1:, 4]), "kept rows are not in descending score order"
        assert np.all(r[:, 4] > conf)
        assert np.all(r[:, 2] > r[:, 0]) and np.all(r[:, 3] > r[:, 1])
        if b % 8:
            continue
        # within a class no kept pair overlaps more than the threshold.  The kernel (like the reference) evaluates the IoU
        # on boxes shifted by cls * 4096 in fp32, which moves a coordinate by up to 4e-3: hence the tolerance.
        for i in range(0, max_det, 5):
            same = r[:, 5] == r[i, 5]
            same[i] = False
            if same.any():
                assert float(_iou_xyxy(r[i], r[same]).max()) <= iou + 2e-2
    # every kept row is one of the image's candidates with its best class and conf = obj * cls_prob
    b = 9
    r = out[b].cpu().numpy()
    cx = {np.float32(row[0] - row[2] / np.float32(2)).item(): j for j, row in enumerate(p[b])}
    for row in r[::13]:
        j = cx[float(row[0])]
        assert int(row[5]) == int((p[b, j, 5:] * p[b, j, 4]).argmax())
        assert row[4] == np.float32(p[b, j, 5 + int(row[5])] * p[b, j, 4])
    # idempotence: the kept rows fed back as predictions (objectness 1, one-hot class score = conf) are all kept, in order
    rt = out[5]
    again = torch.zeros((1, max_det, 15), device=rt.device)
    again[0, :, 0] = (rt[:, 0] + rt[:, 2]) / 2
    again[0, :, 1] = (rt[:, 1] + rt[:, 3]) / 2
    again[0, :, 2] = rt[:, 2] - rt[:, 0]
    again[0, :, 3] = rt[:, 3] - rt[:, 1]
    again[0, :, 4] = 1.0
    again[0, torch.arange(max_det), 5 + rt[:, 5].long()] = rt[:, 4]
    out2 = non_max_suppression(again, conf, iou, max_detections=max_det, compat="fixed")[0]
    assert out2.shape == rt.shape and torch.equal(out2[:, 4:], rt[:, 4:])
    assert float((out2[:, :4] - rt[:, :4]).abs().max()) < 1e-3  # xyxy -> xywh -> xyxy round trip in fp32
