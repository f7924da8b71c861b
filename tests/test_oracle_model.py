"""Pins the CPU oracle (oracle/model.py) to the reference: committed golden outputs were produced
by the reference's own modules (tests/golden/make_golden.py); reference-marked tests re-run the
reference live when /root/reference is present."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import model as om

TOL = 2e-5  # fp32 re-association only (relative to per-tensor max-abs)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("name", list(cases.MODEL_CASES))
def test_model_matches_reference_golden(name, golden_dir):
    variant, seed, shape = cases.MODEL_CASES[name]
    cfg = om.get_cfg(variant)
    sd = om.make_state_dict(cfg, seed)
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    det, raws = om.forward(cases.image(shape), sd, cfg)
    for i, r in enumerate(raws):
        assert r.shape == g[f"raw{i}"].shape
        assert _rel(r, torch.from_numpy(g[f"raw{i}"])) < TOL
    assert det.shape == g["det"].shape
    assert _rel(det, torch.from_numpy(g["det"])) < TOL


@pytest.mark.parametrize("name", list(cases.CLA_CASES))
def test_cla_closed_form_matches_reference_golden(name, golden_dir):
    cq, ck, hq, wq, seed = cases.CLA_CASES[name]
    sd = cases.cla_state(cq, ck, seed)
    q, k = cases.cla_inputs(cq, ck, hq, wq, seed)
    y = om.cla(q, k, sd, "cla")
    g = torch.from_numpy(np.load(os.path.join(golden_dir, f"cla_{name}.npz"))["y"])
    assert _rel(y, g) < TOL


@pytest.mark.parametrize("name", list(cases.TL_CASES))
def test_transformer_layer_matches_reference_golden(name, golden_dir):
    c, heads, h, w, seed = cases.TL_CASES[name]
    y = om.transformer_layer(cases.tl_input(c, h, w, seed), cases.tl_state(c, seed), "tl", heads)
    g = torch.from_numpy(np.load(os.path.join(golden_dir, f"tl_{name}.npz"))["y"])
    assert _rel(y, g) < TOL


def test_state_spec_shapes_and_determinism():
    cfg = om.get_cfg("skyeye_s")
    sd = om.make_state_dict(cfg, 0)
    n = sum(v.numel() for k, v in sd.items() if "running" not in k and "num_batches" not in k)
    assert n == 9_310_391 or abs(n - 9.31e6) < 0.02e6  # 9.31 M params
    sd2 = om.make_state_dict(cfg, 0)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert om.forward(torch.rand(1, 3, 64, 64), sd, cfg)[0].shape == (1, 3 * (64 + 16 + 4), 15)


def test_bf16_emulation_close_to_fp32():
    cfg = om.get_cfg("skyeye_tiny")
    sd = om.make_state_dict(cfg, 0)
    x = cases.image((1, 3, 64, 64))
    _, r32 = om.forward(x, sd, cfg)
    _, r16 = om.forward(x, sd, cfg, emu="bf16")
    for a, b in zip(r32, r16):
        assert _rel(b, a) < 0.05


@pytest.mark.reference
@pytest.mark.parametrize("variant", ["skyeye_tiny", "skyeye_tiny_l"])
def test_live_reference_matches_oracle(variant):
    from oracle import ref_loader
    cfg = om.get_cfg(variant)
    sd = om.make_state_dict(cfg, 7)
    m = ref_loader.build_reference_model(cfg)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    x = cases.image((1, 3, 96, 64), seed=9)
    with torch.no_grad():
        det, raws = m(x)
    d2, r2 = om.forward(x, sd, cfg)
    assert _rel(d2, det) < TOL
    for a, b in zip(r2, raws):
        assert _rel(a, b) < TOL


@pytest.mark.parametrize("name", list(cases.WSA_CASES))
def test_windowed_self_attention_matches_reference_golden(name, golden_dir):
    """oracle.windowed_self_attention vs the reference class executed in the build container (attention.py:312-399)."""
    dim, window, heads, n_win, n_mask, seed = cases.WSA_CASES[name]
    x, mask = cases.wsa_inputs(dim, window, n_win, n_mask, seed)
    y = om.windowed_self_attention(x, cases.wsa_state(dim, window, heads, seed), "wsa", window, heads, mask)
    g = torch.from_numpy(np.load(os.path.join(golden_dir, f"wsa_{name}.npz"))["y"])
    assert y.shape == g.shape
    assert float((y - g).abs().max()) <= 2e-5 * float(g.abs().max())
