"""Parity of the HBM-bound kernels and the flash-attention kernel (through the C ABI) against the
CPU oracle (oracle/model.py) on identical seeded inputs.  Inputs are bf16-representable; outputs
are stored in bf16 by these kernels, so the stated bound is one bf16 rounding (2^-8) on top of the
1e-3 fp32-accumulate tolerance: <= 6e-3 relative to max |reference| (decode is fp32: <= 1e-5)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
from gpu_util import bf16r, randn, rel_err
from oracle import model as om

pytestmark = pytest.mark.gpu
TOL = 6e-3


def test_focus_space_to_depth_bit_exact():
    from skyeye import engine as E
    img = cases.image((2, 3, 32, 48)).cuda()
    y = E.new_buffer(2, 16, 24, 32)
    y.t.fill_(1.0)
    E.focus(img, y)
    torch.cuda.synchronize()
    x = img.cpu()
    ref = torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)  # blocks.py:174-181
    got = y.nchw().float().cpu()
    assert torch.equal(got[:, :12], bf16r(ref))
    assert bool((got[:, 12:] == 0).all())


@pytest.mark.parametrize("shape,cout,u8", [((2, 3, 32, 48), 64, False), ((1, 3, 64, 40), 32, True), ((3, 3, 24, 264), 64, True),
                                           ((1, 3, 16, 16), 128, False)])
def test_focus_block_fused_matches_oracle(shape, cout, u8):
    """Whole FocusBlock (blocks.py:170-182) through skb_focus_conv_bf16: padded space-to-depth + the
    3x3 conv over sliding 128-byte windows.  Reference = conv2d over the concatenated strided slices
    with the same bf16-rounded operands (fp32 accumulate); includes the left/right/top/bottom borders
    and ragged tiles (264 / 2 = 132 pixels = one full 128-pixel tile + 4)."""
    from skyeye import engine as E
    n, _, h, w = shape
    if u8:
        img8 = torch.from_numpy(cases.rng("focus_u8", shape).integers(0, 256, shape, dtype=np.uint8))
        x = img8.float() / 255.0
        dev_img = img8.cuda()
    else:
        x = cases.image(shape)
        dev_img = x.cuda()
    wt = bf16r(randn(("focus_w", cout), (cout, 12, 3, 3), (2.0 / 108) ** 0.5))
    b = randn(("focus_b", cout), (cout,), 0.1)
    pw = E.PackedFocusConv(wt, b)
    y = E.new_buffer(n, h // 2, w // 2, cout)
    y.t.fill_(7.0)
    ws = E.workspace(E.N.lib().skb_focus_conv_workspace_bytes(n, h, w))
    E.focus_conv(dev_img, pw, y, ws)
    torch.cuda.synchronize()
    s2d = bf16r(torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1))
    ref = F.silu(F.conv2d(s2d, wt, b, 1, 1))
    assert rel_err(y.nchw(), ref) < TOL


@pytest.mark.parametrize("shape", [(2, 64, 20, 20), (1, 256, 7, 9), (2, 32, 40, 40)])
def test_maxpool5_cascade_equals_spp_pools_bit_exact(shape):
    from skyeye import engine as E
    x = bf16r(randn(("mp", shape), shape))
    n, c, h, w = shape
    cat = torch.zeros((n, h, w, 4 * c), dtype=torch.bfloat16, device="cuda")
    cat[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    for i in range(3):
        E.maxpool5(E.View(cat, i * c, c), E.View(cat, (i + 1) * c, c))
    torch.cuda.synchronize()
    for i, k in enumerate((5, 9, 13)):  # blocks.py:143-149
        ref = F.max_pool2d(x, k, 1, k // 2)
        got = cat[..., (i + 1) * c:(i + 2) * c].permute(0, 3, 1, 2).float().cpu()
        assert torch.equal(got, ref), k


@pytest.mark.parametrize("shape", [(2, 64, 20, 20), (1, 256, 7, 9), (2, 32, 40, 40), (1, 512, 40, 40), (2, 16, 1, 3)])
def test_fused_spp_pools_bit_exact(shape):
    """skb_spp_pools_bf16: the 5 / 9 / 13 pools of blocks.py:143-149 in one pass, written to the concat slices behind the input."""
    from skyeye import engine as E
    x = bf16r(randn(("spp", shape), shape))
    n, c, h, w = shape
    cat = torch.zeros((n, h, w, 4 * c), dtype=torch.bfloat16, device="cuda")
    cat[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    E.spp_pools(*(E.View(cat, i * c, c) for i in range(4)))
    torch.cuda.synchronize()
    assert torch.equal(cat[..., :c].permute(0, 3, 1, 2).float().cpu(), x)   # the input slice is untouched
    for i, k in enumerate((5, 9, 13)):
        ref = F.max_pool2d(x, k, 1, k // 2)
        got = cat[..., (i + 1) * c:(i + 2) * c].permute(0, 3, 1, 2).float().cpu()
        assert torch.equal(got, ref), k


@pytest.mark.parametrize("shape", [(2, 256, 20, 20), (1, 512, 40, 24), (3, 64, 9, 11)])
def test_cbam_matches_oracle(shape):
    from skyeye import engine as E
    n, c, h, w = shape
    r = max(c // 16, 1)
    x = bf16r(randn(("cbx", shape), shape))
    sd = {"p.channel_attention.shared_mlp.0.weight": randn(("cb0", shape), (r, c), 0.1),
          "p.channel_attention.shared_mlp.2.weight": randn(("cb1", shape), (c, r), 0.1),
          "p.spatial_attention.conv.weight": randn(("cb7", shape), (1, 2, 7, 7), 0.2)}
    ref = om.cbam(x, sd, "p")
    xv = E.from_nchw(x.cuda())
    y = E.new_buffer(n, h, w, c)
    ws = E.workspace(E.N.lib().skb_cbam_workspace_bytes(n, h, w, c))
    E.cbam(xv, sd["p.channel_attention.shared_mlp.0.weight"].cuda(), sd["p.channel_attention.shared_mlp.2.weight"].cuda(),
           sd["p.spatial_attention.conv.weight"].cuda().contiguous(), y, ws)
    torch.cuda.synchronize()
    assert rel_err(y.nchw(), ref) < TOL


@pytest.mark.parametrize("cq,ck,hq,wq", [(64, 128, 12, 10), (256, 512, 16, 16), (512, 1024, 8, 6), (32, 64, 10, 12)])
def test_cla_core_matches_oracle_closed_form(cq, ck, hq, wq):
    """softmax over image rows, R^2 * a * bilinear(V) (attention.py:208-235)."""
    from skyeye import engine as E
    heads, n = 4, 2
    q = bf16r(randn(("clq", cq, hq), (n, cq, hq, wq)))
    k = bf16r(randn(("clk", cq, hq), (n, cq, hq // 2, wq // 2)))
    v = bf16r(randn(("clv", ck, hq), (n, ck, hq // 2, wq // 2)))
    ku = F.interpolate(k, size=(hq, wq), mode="bilinear", align_corners=False)
    vu = F.interpolate(v, size=(hq, wq), mode="bilinear", align_corners=False)
    s = (q.view(n, heads, cq // heads, hq, wq) * ku.view(n, heads, cq // heads, hq, wq)).sum(2) / math.sqrt(cq)
    a = torch.softmax(s, dim=2)
    ref = ((4.0 * a).unsqueeze(2) * vu.view(n, heads, ck // heads, hq, wq)).reshape(n, ck, hq, wq)
    o = E.new_buffer(n, hq, wq, ck)
    ws = E.workspace(E.N.lib().skb_cla_workspace_bytes(n, hq, wq, heads))
    E.cla_core(E.from_nchw(q.cuda()), E.from_nchw(k.cuda()), E.from_nchw(v.cuda()), o, heads, 1.0 / math.sqrt(cq), 4.0, ws)
    torch.cuda.synchronize()
    assert rel_err(o.nchw(), ref) < TOL


@pytest.mark.parametrize("c", [64, 256, 512, 1024])
def test_layernorm_matches_oracle(c):
    from skyeye import engine as E
    x = bf16r(randn(("lnx", c), (2, c, 5, 7), 2.0) + 0.5)
    g = 1.0 + randn(("lng", c), (c,), 0.1)
    b = randn(("lnb", c), (c,), 0.1)
    ref = F.layer_norm(x.permute(0, 2, 3, 1), (c,), g, b, 1e-5).permute(0, 3, 1, 2)  # attention.py:297
    y = E.new_buffer(2, 5, 7, c)
    E.layernorm(E.from_nchw(x.cuda()), g.cuda(), b.cuda(), y)
    torch.cuda.synchronize()
    assert rel_err(y.nchw(), ref) < TOL


def _attn_ref(qkv, heads):
    B, N, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    q, k, v = (z.reshape(B, N, heads, hd).transpose(1, 2) for z in qkv.chunk(3, dim=-1))
    o = om._attention_chunked(q, k, v, 1.0 / math.sqrt(hd))
    return o.transpose(1, 2).reshape(B, N, C)


@pytest.mark.parametrize("b,h,w,heads", [(1, 16, 16, 1), (2, 16, 24, 2), (1, 4, 4, 4), (2, 10, 10, 1), (1, 40, 40, 4), (1, 64, 80, 2),
                                           (1, 41, 100, 2), (1, 72, 72, 1)])
def test_flash_attention_matches_oracle(b, h, w, heads):
    """N = h*w tokens incl. ragged query and key tiles (16, 100, 384, 1600, 5120); N >= 4096 runs two query tiles per CTA
    (4100: the last CTA's second tile is entirely out of range; 5184 = 20.25 double tiles)."""
    from skyeye import engine as E
    C = heads * 64
    qkv = bf16r(randn(("att", b, h, w, heads), (b, h * w, 3 * C)))
    ref = _attn_ref(qkv, heads)
    qv = E.View(qkv.view(b, h, w, 3 * C).to(torch.bfloat16).cuda().contiguous())
    o = E.new_buffer(b, h, w, C)
    o.t.zero_()
    E.flash_attn(qv, o, heads, 1.0 / 8.0)
    torch.cuda.synchronize()
    got = o.torch().reshape(b, h * w, C).float().cpu()
    assert rel_err(got, ref) < TOL


def test_flash_attention_peaky_logits_exercise_lazy_rescale():
    """Large-magnitude, position-dependent logits force reference-max growth (O rescale path)."""
    from skyeye import engine as E
    b, h, w, heads = 1, 32, 32, 1
    C = 64
    qkv = randn(("attp",), (b, h * w, 3 * C))
    ramp = torch.linspace(0.2, 6.0, h * w).view(1, -1, 1)
    qkv[..., C:2 * C] *= ramp  # keys grow along the sequence -> running max keeps increasing
    qkv = bf16r(qkv)
    ref = _attn_ref(qkv, heads)
    o = E.new_buffer(b, h, w, C)
    E.flash_attn(E.View(qkv.view(b, h, w, 3 * C).to(torch.bfloat16).cuda().contiguous()), o, heads, 1.0 / 8.0)
    torch.cuda.synchronize()
    assert rel_err(o.torch().reshape(b, h * w, C), ref) < 2e-2


@pytest.mark.parametrize("h,w,heads", [(40, 40, 2), (80, 80, 2)])  # N = 1600 (one query tile per CTA) / 6400 (two)
def test_flash_attention_late_rescale_in_one_warp_does_not_race_the_epilogue(h, w, heads):
    """Regression (found by the teacher-forced test on real skyeye_l activations): rows 32..63 of every query tile see
    their largest logits only in the LAST three key tiles (growing by > 8 in the log2 domain each time), so that one
    softmax warp takes the O-rescale path late while the other three warps run two tiles ahead into the epilogue.
    The epilogue used to wait on a per-tile barrier by parity and fell through two phases early."""
    from skyeye import engine as E
    b, C, N = 1, heads * 64, h * w
    qkv = randn(("att_late", h, w), (b, N, 3 * C), 0.5)
    q, k = qkv[..., :C], qkv[..., C:2 * C]
    rows = (torch.arange(N) % 128 >= 32) & (torch.arange(N) % 128 < 64)
    for hd in range(heads):
        q[:, :, hd * 64] = 0.0
        q[:, rows, hd * 64] = 4.0
        k[:, :, hd * 64] = 0.0
        for t, c in ((3, 16.0), (2, 32.0), (1, 48.0)):   # logits 8, 16, 24 (natural) in the last three 64-key tiles
            k[:, N - 64 * t:N - 64 * (t - 1), hd * 64] = c
    qkv = bf16r(qkv)
    ref = _attn_ref(qkv, heads)
    qv = E.View(qkv.view(b, h, w, 3 * C).to(torch.bfloat16).cuda().contiguous())
    for _ in range(5):
        o = E.new_buffer(b, h, w, C)
        o.t.zero_()
        E.flash_attn(qv, o, heads, 1.0 / 8.0)
        torch.cuda.synchronize()
        assert rel_err(o.torch().reshape(b, N, C), ref) < TOL


@pytest.mark.parametrize("hw", [(64, 96), (160, 128)])
def test_decode_matches_oracle(hw):
    """process_detections (detector.py:88-145) incl. the [B,na,h,w,no] relayout of raw outputs."""
    from skyeye import engine as E
    H, W = hw
    B, na, no = 2, 3, 15
    raws_ref, views = [], []
    for i, s in enumerate((8, 16, 32)):
        r = randn(("dec", hw, i), (B, na, H // s, W // s, no), 2.0)
        raws_ref.append(r)
        buf = torch.zeros((B, H // s, W // s, 48), dtype=torch.float32, device="cuda")
        buf[..., :45] = r.permute(0, 2, 3, 1, 4).reshape(B, H // s, W // s, 45).cuda()
        views.append(E.View(buf))
    ref = om.decode(raws_ref, (H, W))
    det = torch.zeros(ref.shape, dtype=torch.float32, device="cuda")
    raw_out = [torch.zeros(r.shape, dtype=torch.float32, device="cuda") for r in raws_ref]
    E.decode(views, na, no, om.DEFAULT_ANCHORS, (H, W), det, raw_out)
    torch.cuda.synchronize()
    assert rel_err(det, ref) < 1e-5
    assert float((det.cpu() - ref).abs().max() / ref.abs().clamp_min(1.0).max()) < 1e-5
    for a, b_ in zip(raw_out, raws_ref):
        assert torch.equal(a.cpu(), b_)


@pytest.mark.parametrize("name", list(cases.WSA_CASES) + ["c256w8h4"])
def test_windowed_self_attention_matches_oracle(name):
    """WindowedSelfAttention (attention.py:312-399) through the native lowering vs the CPU oracle on identical
    bf16-representable inputs; includes a 64-token window with head_dim 64 and a masked case."""
    from skyeye.core.models.attention import WindowedSelfAttention
    dim, window, heads, n_win, n_mask, seed = cases.WSA_CASES.get(name, (256, 8, 4, 5, 0, 11))
    sd = {k: (bf16r(v) if v.dim() == 2 and "table" not in k else v) for k, v in cases.wsa_state(dim, window, heads, seed).items()}
    x, mask = cases.wsa_inputs(dim, window, n_win, n_mask, seed)
    x = bf16r(x)
    mod = WindowedSelfAttention(dim, window, heads)
    mod.load_state_dict({k[len("wsa."):]: v for k, v in sd.items()}, strict=False)
    mod = mod.cuda().eval()
    got = mod(x.cuda(), None if mask is None else mask.cuda())
    torch.cuda.synchronize()
    ref = om.windowed_self_attention(x, sd, "wsa", window, heads, mask, om.Ctx("bf16"))
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL


@pytest.mark.parametrize("B,H,W,C,heads", [(2, 80, 80, 512, 8), (1, 40, 24, 256, 4), (3, 8, 8, 64, 1), (1, 16, 8, 1024, 16)])
def test_window_attn2d_tensor_core_path_matches_fp32_reference(B, H, W, C, heads):
    """skb_window_attn2d_bf16 on the shape class the detector uses (8 x 8 windows, head_dim 64, no mask: the persistent tcgen05
    kernel, pairs of windows per tile, several tiles per CTA, odd window counts) vs softmax(q k^T * scale + bias) v
    (attention.py:372-395) in fp32 on the same bf16 qkv, windows partitioned / reversed explicitly."""
    from skyeye import engine as E
    from skyeye.core.models.attention import window_partition, window_reverse
    qkv = bf16r(randn(("wa2d", B, H, W, C), (B, H, W, 3 * C)))
    bias = randn(("wa2d-bias", heads), (heads, 64, 64))
    scale = 64 ** -0.5
    o = E.new_buffer(B, H, W, C)
    o.t.fill_(7.0)
    E.window_attn2d(E.View(qkv.to(torch.bfloat16).cuda()), bias.cuda(), None, o, heads, 8, scale)
    torch.cuda.synchronize()
    wins = window_partition(qkv, 8).view(-1, 64, 3, heads, 64).permute(2, 0, 3, 1, 4)      # [3, nW, heads, 64, 64]
    att = torch.softmax(wins[0] @ wins[1].transpose(-2, -1) * scale + bias.unsqueeze(0), dim=-1) @ wins[2]
    ref = window_reverse(att.transpose(1, 2).reshape(-1, 64, C), 8, H, W)
    assert rel_err(o.torch().float().cpu(), ref) < TOL   # NHWC both


@pytest.mark.parametrize("hw,new", [((100, 200), 128), ((720, 1280), 640), ((333, 517), 256), ((64, 96), 96), ((96, 128), 128),
                                    ((50, 70), 160)])
def test_letterbox_gpu_matches_cv2_pipeline(hw, new):
    """skb_letterbox_u8 vs the host pipeline it replaces: cv2.resize(INTER_LINEAR) + copyMakeBorder(114) (letterbox,
    augmentation.py:442-496) + BGR->RGB + HWC->CHW (detect.py:131-132).  Bit-exact when down-scaling or copying; when
    up-scaling OpenCV's dispatched resize differs from its own fixed-point reference by 1 LSB on a few samples."""
    cv2 = pytest.importorskip("cv2")
    from skyeye.utils.general import letterbox, letterbox_gpu
    img = cases.rng("lb", hw, new).integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    ref_hwc, r, (left, top) = letterbox(img, new)
    ref = np.ascontiguousarray(ref_hwc[..., ::-1].transpose(2, 0, 1))
    got, r2, (left2, top2) = letterbox_gpu(img, new)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.shape == ref.shape and (left2, top2) == (left, top) and abs(r - r2) < 1e-12
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    if r <= 1.0:
        assert int(diff.max()) == 0
    else:
        assert int(diff.max()) <= 1 and float((diff > 0).mean()) < 5e-3


@pytest.mark.parametrize("n,m,ncls,seed", [(300, 40, 10, 0), (1000, 500, 3, 1), (7, 1, 1, 2), (64, 200, 80, 3)])
def test_match_detections_kernel_equals_the_reference_bookkeeping(n, m, ncls, seed):
    """process_batch (validate.py:71-108) on the device vs its torch / numpy sequence on the host (which tests/test_eval_bookkeeping.py
    checks against the live reference): the correct-prediction matrices must be identical."""
    from skyeye.cli import validate as V
    g = cases.rng("match", n, m, seed)
    def boxes(k):
        xy = g.random((k, 2), dtype=np.float32) * 600
        wh = g.random((k, 2), dtype=np.float32) * 80 + 8
        return np.concatenate((xy, xy + wh), 1)
    lab_boxes = boxes(m)
    labels = np.concatenate((g.integers(0, ncls, (m, 1)).astype(np.float32), lab_boxes), 1)
    # detections: jittered copies of labels (so many pairs overlap strongly) + random boxes, random classes for a third of them
    src = g.integers(0, m, n)
    det_boxes = lab_boxes[src] + g.normal(0, 4.0, (n, 4)).astype(np.float32)
    det_boxes[:, 2:] = np.maximum(det_boxes[:, 2:], det_boxes[:, :2] + 1)
    cls = labels[src, 0].copy()
    flip = g.random(n) < 0.33
    cls[flip] = g.integers(0, ncls, int(flip.sum())).astype(np.float32)
    det = np.concatenate((det_boxes, g.random((n, 1), dtype=np.float32), cls[:, None]), 1).astype(np.float32)
    iouv = torch.linspace(0.5, 0.95, 10)
    ref = V.process_batch(torch.from_numpy(det), torch.from_numpy(labels), iouv)
    got = V.process_batch(torch.from_numpy(det).cuda(), torch.from_numpy(labels).cuda(), iouv.cuda()).cpu()
    assert got.shape == ref.shape and got.dtype == torch.bool
    assert torch.equal(got, ref), int((got != ref).sum())
    assert int(ref.sum()) > 0
    assert V.process_batch(torch.from_numpy(det[:0]).cuda(), torch.from_numpy(labels).cuda(), iouv.cuda()).shape == (0, 10)
    assert not V.process_batch(torch.from_numpy(det).cuda(), torch.from_numpy(labels[:0]).cuda(), iouv.cuda()).any()
