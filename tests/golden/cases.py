"""Seeded input generators shared by make_golden.py (reference side) and the tests (oracle / CUDA
side). Everything derives from numpy PCG64 streams so inputs are identical on every machine."""
import math
import zlib

import numpy as np
import torch


def rng(*key):
    return np.random.Generator(np.random.PCG64([zlib.crc32(repr(key).encode())]))


def image(shape, seed=1234):
    return torch.from_numpy(rng("image", shape, seed).random(shape, dtype=np.float32))


# name -> (variant, weight seed, input shape)
MODEL_CASES = {
    "tiny": ("skyeye_tiny", 0, (2, 3, 64, 96)),
    "tiny_l": ("skyeye_tiny_l", 0, (2, 3, 64, 96)),
}

# name -> (Cq, Ck, Hq, Wq, seed)    key map is (Hq/2, Wq/2)
CLA_CASES = {"c32": (32, 64, 12, 10, 3), "c64": (64, 128, 8, 16, 4)}
# name -> (C, heads, H, W, seed)
TL_CASES = {"c32h2": (32, 2, 6, 7, 5), "c128h2": (128, 2, 9, 8, 6)}


# name -> (dim, window, heads, n_windows_total, n_mask_windows (0 = no mask), seed)
WSA_CASES = {"c64w4h2": (64, 4, 2, 6, 0, 7), "c128w8h2": (128, 8, 2, 4, 0, 8), "c64w8h1_mask": (64, 8, 1, 6, 3, 9)}


def _normal(key, shape, std):
    return torch.from_numpy((rng(*key).standard_normal(shape, dtype=np.float32) * std).astype(np.float32))


def cla_state(cq, ck, seed):
    sd = {}
    for proj, co, ci in (("query_projection", cq, cq), ("key_projection", cq, ck),
                         ("value_projection", ck, ck), ("output_projection", cq, ck)):
        sd[f"cla.{proj}.weight"] = _normal(("claw", proj, seed), (co, ci, 1, 1), 1.0 / math.sqrt(ci))
        sd[f"cla.{proj}.bias"] = _normal(("clab", proj, seed), (co,), 0.1)
    return sd


def cla_inputs(cq, ck, hq, wq, seed):
    return (_normal(("claq", seed), (2, cq, hq, wq), 1.0), _normal(("clak", seed), (2, ck, hq // 2, wq // 2), 1.0))


def tl_state(c, seed):
    sd = {}
    for key, shape, std in (("self_attn.in_proj_weight", (3 * c, c), 1 / math.sqrt(c)),
                            ("self_attn.in_proj_bias", (3 * c,), 0.02),
                            ("self_attn.out_proj.weight", (c, c), 1 / math.sqrt(c)),
                            ("self_attn.out_proj.bias", (c,), 0.02),
                            ("norm1.bias", (c,), 0.1), ("norm2.bias", (c,), 0.1),
                            ("feedforward.0.weight", (4 * c, c), 1 / math.sqrt(c)),
                            ("feedforward.0.bias", (4 * c,), 0.02),
                            ("feedforward.3.weight", (c, 4 * c), 1 / math.sqrt(4 * c)),
                            ("feedforward.3.bias", (c,), 0.02)):
        sd["tl." + key] = _normal(("tl", key, seed), shape, std)
    sd["tl.norm1.weight"] = 1.0 + _normal(("tl", "n1w", seed), (c,), 0.1)
    sd["tl.norm2.weight"] = 1.0 + _normal(("tl", "n2w", seed), (c,), 0.1)
    return sd


def wsa_state(dim, window, heads, seed):
    """WindowedSelfAttention parameters (attention.py:333-357), reference key names under 'wsa.'."""
    sd = {"wsa.qkv.weight": _normal(("wsaw", seed), (3 * dim, dim), 1 / math.sqrt(dim)),
          "wsa.qkv.bias": _normal(("wsab", seed), (3 * dim,), 0.02),
          "wsa.proj.weight": _normal(("wsapw", seed), (dim, dim), 1 / math.sqrt(dim)),
          "wsa.proj.bias": _normal(("wsapb", seed), (dim,), 0.02),
          "wsa.relative_position_bias_table": _normal(("wsat", seed), ((2 * window - 1) ** 2, heads), 0.5)}
    return sd


def wsa_inputs(dim, window, n_win, n_mask, seed):
    x = _normal(("wsax", seed), (n_win, window * window, dim), 1.0)
    mask = None
    if n_mask:
        m = rng("wsam", seed).random((n_mask, window * window, window * window)) < 0.2
        mask = torch.from_numpy(np.where(m, -100.0, 0.0).astype(np.float32))  # Swin-style additive mask
    return x, mask


def tl_input(c, h, w, seed):
    return _normal(("tlx", seed), (2, c, h, w), 1.0)


# ------------------------------------------------------------------------------------------------
# NMS cases
# ------------------------------------------------------------------------------------------------
NMS_CORE_CASES = ["random2000", "ties5000", "zero_area", "dense_small", "single", "all_same"]


def nms_core_case(name):
    """-> boxes [n,4] xyxy fp32, scores [n] fp32, iou threshold."""
    g = rng("nms_core", name)
    if name == "random2000":
        n = 2000
        xy = g.random((n, 2), dtype=np.float32) * 400
        wh = g.random((n, 2), dtype=np.float32) * 80 + 1
        return np.concatenate((xy, xy + wh), 1).astype(np.float32), g.random(n, dtype=np.float32), 0.45
    if name == "ties5000":  # quantised scores (many ties) + exact duplicate boxes
        n = 5000
        xy = np.floor(g.random((n, 2)) * 300).astype(np.float32)
        wh = np.floor(g.random((n, 2)) * 40 + 2).astype(np.float32)
        b = np.concatenate((xy, xy + wh), 1).astype(np.float32)
        b[n // 2:] = b[: n - n // 2]  # duplicates
        s = (np.floor(g.random(n) * 64) / 64).astype(np.float32)
        return b, s, 0.5
    if name == "zero_area":  # 0/0 -> NaN never suppresses: keep [0,1,2]
        return (np.array([[1, 1, 1, 1], [1, 1, 1, 1], [0, 0, 2, 2]], dtype=np.float32),
                np.array([0.9, 0.8, 0.7], dtype=np.float32), 0.5)
    if name == "dense_small":  # VisDrone-like: many small boxes with class offsets
        n = 8000
        c = g.random((n, 2), dtype=np.float32) * 1280
        wh = np.exp(g.uniform(np.log(4), np.log(64), (n, 2))).astype(np.float32)
        off = (g.integers(0, 10, (n, 1)) * 4096).astype(np.float32)
        b = np.concatenate((c - wh / 2 + off, c + wh / 2 + off), 1).astype(np.float32)
        s = g.permutation(np.linspace(0.002, 0.999, n)).astype(np.float32)
        return b, s, 0.6
    if name == "single":
        return np.array([[0, 0, 10, 10]], dtype=np.float32), np.array([0.5], dtype=np.float32), 0.5
    if name == "all_same":
        n = 300
        return (np.tile(np.array([[5, 5, 25, 30]], dtype=np.float32), (n, 1)),
                g.permutation(np.linspace(0.1, 0.9, n)).astype(np.float32), 0.5)
    raise KeyError(name)


NMS_WRAPPER_CASES = ["nc10_best", "nc10_multi", "nc1", "nc10_agnostic", "nc10_classes", "cap30000",
                     "validate_multi", "empty", "maxdet1000"]


def _pred(key, B, N, nc, span=640.0, unique=True):
    g = rng("nms_wrap", key)
    p = np.empty((B, N, 5 + nc), dtype=np.float32)
    p[..., 0:2] = g.random((B, N, 2), dtype=np.float32) * span
    p[..., 2:4] = np.exp(g.uniform(np.log(4), np.log(96), (B, N, 2))).astype(np.float32)
    for b in range(B):  # unique objectness -> tie-free ordering (SURVEY §7 hard parts)
        p[b, :, 4] = g.permutation(np.linspace(0.002, 0.999, N)).astype(np.float32)
    p[..., 5:] = g.random((B, N, nc), dtype=np.float32)
    return p


def nms_wrapper_case(name):
    """-> prediction [B,N,5+nc] fp32 and kwargs for non_max_suppression (metrics.py:361-369)."""
    if name == "nc10_best":
        return _pred(name, 3, 3000, 10), dict(conf_threshold=0.25, iou_threshold=0.45)
    if name == "nc10_multi":
        return _pred(name, 2, 1500, 10), dict(conf_threshold=0.3, iou_threshold=0.45, multi_label=True)
    if name == "nc1":
        return _pred(name, 2, 2500, 1), dict(conf_threshold=0.25, iou_threshold=0.45)
    if name == "nc10_agnostic":
        return _pred(name, 2, 2500, 10), dict(conf_threshold=0.25, iou_threshold=0.45, agnostic=True)
    if name == "nc10_classes":  # quirk X8c: the filter compares the class-PROBABILITY column with ids
        p = _pred(name, 2, 500, 10)
        p[:, ::7, 5:] = 0.0
        p[:, ::7, 8] = 1.0  # best-class prob exactly 1.0 -> passes classes=[1]
        return p, dict(conf_threshold=0.1, iou_threshold=0.45, classes=[1])
    if name == "cap30000":
        return _pred(name, 1, 36000, 10, span=1280.0), dict(conf_threshold=0.001, iou_threshold=0.6)
    if name == "validate_multi":
        # validate.py:255 settings. 2900*10 rows stay below the 30000 cap on purpose: multi_label rows
        # of one box share the same objectness, and the cap's argsort is NON-stable (metrics.py:432),
        # so the reference's row order above the cap is unspecified for ties (SURVEY §7 hard parts).
        return _pred(name, 1, 2900, 10, span=1280.0), dict(conf_threshold=0.001, iou_threshold=0.6, multi_label=True)
    if name == "empty":
        p = _pred(name, 2, 300, 10)
        p[0, :, 4] *= 0.1  # image 0 has no candidate above 0.25
        return p, dict(conf_threshold=0.25, iou_threshold=0.45)
    if name == "maxdet1000":
        return _pred(name, 1, 6000, 10, span=2000.0), dict(conf_threshold=0.05, iou_threshold=0.45, max_detections=1000)
    raise KeyError(name)
