"""``head: windowed`` (NOT IN REFERENCE beyond the WindowedSelfAttention class itself, attention.py:312-399).

CPU: the oracle's window partition / reverse are inverses, the windowed layer's attention equals the reference class's
restatement (``om.windowed_self_attention``, pinned to the reference by tests/golden/wsa_*.npz) applied to explicitly
partitioned windows, and equals a GLOBAL attention with a block mask + the relative-position bias (an independent statement
of what "window attention" means); the model's state dict has the reference class's parameter names.
GPU (-m gpu): the native layer (partition / reverse folded into skb_window_attn2d_bf16's addressing) and the whole
skyeye_nano_lw network against the oracle, teacher-forced per launch."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
from gpu_util import bf16r, randn
from oracle import model as om


def _layer_sd(C, heads, ws, seed=0):
    cfg = dict(base_channels=C // 4, depth_multiple=0.33, nc=10, enhanced=True, head_dim=C // heads, head="windowed", window_size=ws)
    sd = om.make_state_dict(cfg, seed)
    return {k[len("head_transformers.0."):]: v for k, v in sd.items() if k.startswith("head_transformers.0.")}, cfg


def test_partition_reverse_roundtrip_and_layout():
    x = torch.arange(2 * 8 * 12 * 3, dtype=torch.float32).view(2, 8, 12, 3)
    w = om.window_partition(x, 4)
    assert w.shape == (2 * 2 * 3, 16, 3)
    assert torch.equal(om.window_reverse(w, 4, 8, 12), x)
    # window 4 of image 0 = window-grid cell (1, 1): pixels rows 4..7, cols 4..7; token 5 = (row 1, col 1) of it
    assert torch.equal(w[4, 5], x[0, 5, 5])
    from skyeye.core.models.attention import window_partition, window_reverse
    assert torch.equal(window_partition(x, 4), w) and torch.equal(window_reverse(w, 4, 8, 12), x)


def test_windowed_layer_attention_equals_the_reference_class_on_partitioned_windows_and_a_block_masked_global_attention():
    C, heads, ws, B, H, W = 32, 2, 4, 2, 8, 12
    sd, _ = _layer_sd(C, heads, ws)
    sd = {"l." + k: v for k, v in sd.items()}
    x = randn(("wtl", C), (B, C, H, W))
    taps = {}
    y = om.windowed_transformer_layer(x, sd, "l", heads, ws, om.Ctx(None, taps))
    assert y.shape == x.shape
    # (1) the attention sub-block == the reference class restatement on explicitly partitioned windows
    t = x.flatten(2).transpose(1, 2)
    xn = F.layer_norm(t, (C,), sd["l.norm1.weight"], sd["l.norm1.bias"], 1e-5)
    wins = om.window_partition(xn.view(B, H, W, C), ws)
    a = om.windowed_self_attention(wins, sd, "l.attn", ws, heads)               # includes the proj Linear
    t1 = t + om.window_reverse(a, ws, H, W).reshape(B, H * W, C)
    assert torch.allclose(taps["l.proj"].flatten(2).transpose(1, 2), t1, atol=1e-5)
    # (2) == global attention over all H*W tokens with -inf outside the token's window and the relative-position bias inside
    hd = C // heads
    qkv = F.linear(xn, sd["l.attn.qkv.weight"], sd["l.attn.qkv.bias"]).view(B, H * W, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    same = ((ys[:, None] // ws) == (ys[None, :] // ws)) & ((xs[:, None] // ws) == (xs[None, :] // ws))
    rel = ((ys[:, None] - ys[None, :] + ws - 1) * (2 * ws - 1) + (xs[:, None] - xs[None, :] + ws - 1)).clamp(0, (2 * ws - 1) ** 2 - 1)
    bias = sd["l.attn.relative_position_bias_table"][rel.view(-1)].view(H * W, H * W, heads).permute(2, 0, 1)
    logits = (q @ k.transpose(-2, -1)) + bias.unsqueeze(0)
    logits = logits.masked_fill(~same.view(1, 1, H * W, H * W), float("-inf"))
    o = (torch.softmax(logits, -1) @ v).transpose(1, 2).reshape(B, H * W, C)
    assert torch.allclose(taps["l.attn"].flatten(2).transpose(1, 2), o, atol=1e-5)


def test_state_dict_names_follow_the_reference_class_and_the_model_loads_them():
    from skyeye.core.detector import construct_model
    cfg = om.get_cfg("skyeye_nano_lw")
    sd = om.make_state_dict(cfg, 0)
    m = construct_model("skyeye_nano_lw.yaml")
    assert set(m.state_dict()) == set(sd)
    m.load_state_dict(sd, strict=True)
    for k in ("attn.qkv.weight", "attn.proj.bias", "attn.relative_position_bias_table", "attn.relative_position_index"):
        assert f"head_transformers.1.{k}" in sd       # WindowedSelfAttention's own names (attention.py:333-352)
    assert torch.equal(sd["head_transformers.0.attn.relative_position_index"], m.head_transformers[0].attn.relative_position_index)


def test_plan_labels_of_the_windowed_variant_cover_the_oracle_taps():
    from skyeye.core.detector import construct_model
    from skyeye.engine import Plan, View
    cfg = om.get_cfg("skyeye_nano_lw")
    m = construct_model("skyeye_nano_lw.yaml")
    plan = m._build_plan(1, 256, 256, torch.device("meta"))
    taps = {}
    om.forward(torch.rand(1, 3, 256, 256), om.make_state_dict(cfg, 0), cfg, emu="bf16", taps=taps)
    labels = [o["label"] for outs in plan.outs for o in outs]
    assert all(plan.outs) and sorted(set(labels)) == sorted(taps)
    for outs in plan.outs:
        for o in outs:
            got = o["view"].torch() if isinstance(o["view"], View) else o["view"]
            assert tuple(got.shape) == tuple(Plan._expected(o, taps).shape), o["label"]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("C,heads,ws,B,H,W", [(128, 2, 8, 2, 16, 24), (64, 1, 4, 1, 8, 8), (256, 4, 8, 1, 40, 40), (64, 4, 8, 1, 8, 16)])
def test_native_windowed_layer_matches_oracle(C, heads, ws, B, H, W):
    """Layer parity incl. partition / reverse by addressing: head_dim 64 / 16, 64- and 16-token windows, non-square maps."""
    from skyeye.core.models.attention import WindowedTransformerLayer
    sd, _ = _layer_sd(C, heads, ws)
    layer = WindowedTransformerLayer(C, heads, ws)
    layer.load_state_dict(sd, strict=True)
    layer = layer.cuda().eval()
    x = bf16r(randn(("wtl-gpu", C, H, W), (B, C, H, W)))
    ref = om.windowed_transformer_layer(x, {"l." + k: v for k, v in sd.items()}, "l", heads, ws, om.Ctx("bf16"))
    got = layer(x.cuda()).cpu()
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-2   # bf16 storage at the same points as the oracle's emulation


@pytest.mark.gpu
def test_skyeye_nano_lw_teacher_forced_and_whole_model():
    from skyeye.core.detector import construct_model
    from skyeye.utils.metrics import non_max_suppression
    from oracle import nms as onms
    cfg = om.get_cfg("skyeye_nano_lw")
    sd = om.make_state_dict(cfg, 0)
    m = construct_model("skyeye_nano_lw.yaml")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x = cases.image((1, 3, 256, 256))
    taps = {}
    d_ref, r_ref = om.forward(x, sd, cfg, emu="bf16", taps=taps)
    xc = x.cuda()
    plan = m.plan_for(xc)
    m._img[0] = xc
    rows = plan.run_teacher_forced(taps)
    bad = []
    for r in rows:
        rel = r["max_err"] / max(r["ref_max"], 1e-20)
        beyond = r["beyond_ulp"] / max(r["ref_max"], 1e-20)
        ok = rel <= 1e-3 if r["store"] == "f32" else (beyond <= 1e-3 and rel <= 8e-3)
        if not ok:
            bad.append((r["step"], r["label"], rel, beyond))
    assert not bad, bad[:8]
    assert any(r["step"].endswith(".attn") and r["kind"] == "attention" for r in rows)
    det, raws = m(xc)
    for a, b in zip(raws, r_ref):
        assert float((a.cpu() - b).abs().max() / b.abs().max()) < 1.2e-1
    out = non_max_suppression(det, 0.25, 0.45)
    ref = onms.non_max_suppression(det.cpu().numpy(), 0.25, 0.45)
    assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(out, ref))
