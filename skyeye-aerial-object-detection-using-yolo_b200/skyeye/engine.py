"""Host-side execution engine: NHWC views over torch-owned buffers, weight preparation (BN folding,
bf16 packing into the implicit-GEMM layout) and a static launch plan per input shape.

PyTorch is plumbing here (device memory, streams); every op appended to a plan is one call into
libskyeye_b200.so.  A plan owns all intermediate buffers, takes no host synchronisation while it
runs and can therefore be captured into a CUDA graph (``Plan.capture``).
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, List, Optional, Sequence

import torch

from . import _native as N
from ._native import ACT_NONE, ACT_RELU, ACT_SILU, SKB_BF16, SKB_F32, skb_view

BN_EPS = 1e-5


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class View:
    """Channel slice [c0, c0+c) of an NHWC torch buffer [n, h, w, pitch]."""
    __slots__ = ("t", "n", "h", "w", "c", "c0", "pitch", "dtype", "_s")

    def __init__(self, t: torch.Tensor, c0: int = 0, c: Optional[int] = None):
        assert t.dim() == 4 and t.is_contiguous()
        self.t = t
        self.n, self.h, self.w, self.pitch = t.shape
        self.c0 = c0
        self.c = self.pitch - c0 if c is None else c
        assert 0 <= c0 and self.c0 + self.c <= self.pitch
        self.dtype = SKB_BF16 if t.dtype == torch.bfloat16 else SKB_F32
        assert t.dtype in (torch.bfloat16, torch.float32)
        self._s = skb_view(t.data_ptr() + c0 * t.element_size(), self.n, self.h, self.w, self.c, self.pitch, self.dtype)

    def slice(self, c0: int, c1: int) -> "View":
        return View(self.t, self.c0 + c0, c1 - c0)

    @property
    def ref(self):
        return ctypes.byref(self._s)

    def torch(self) -> torch.Tensor:
        """[n, h, w, c] torch view (no copy)."""
        return self.t[..., self.c0:self.c0 + self.c]

    def nchw(self) -> torch.Tensor:
        return self.torch().permute(0, 3, 1, 2)


def new_buffer(n, h, w, c, dtype=torch.bfloat16, device="cuda") -> View:
    return View(torch.empty((n, h, w, c), dtype=dtype, device=device))


def from_nchw(x: torch.Tensor, dtype=torch.bfloat16) -> View:
    """Test helper: NCHW tensor -> NHWC buffer view (host-side plumbing, not on the hot path)."""
    return View(x.permute(0, 2, 3, 1).contiguous().to(dtype))


# ----------------------------------------------------------------------------------------------
# weight preparation
# ----------------------------------------------------------------------------------------------
def round_cout(c: int) -> int:
    return 32 if c <= 32 else (c + 63) // 64 * 64


class PackedConv:
    """bf16 weights [cout_pad, k*k*cin] (K order = (ky, kx, cin): matches the kernel's tap-major
    K loop) + fp32 bias [cout_pad]; rows >= cout are zero."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], device="cuda", cin_pad: Optional[int] = None):
        w = weight.detach().float()
        if w.dim() == 2:
            w = w[:, :, None, None]
        cout, cin, kh, kw = w.shape
        assert kh == kw
        cin_p = cin if cin_pad is None else cin_pad
        self.cout, self.cin, self.k = cout, cin_p, kh
        self.cin_real = cin
        self.cout_pad = round_cout(cout)
        wp = torch.zeros((self.cout_pad, kh, kw, cin_p), dtype=torch.float32)
        wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1).cpu()
        self.w = wp.reshape(self.cout_pad, kh * kw * cin_p).to(device=device, dtype=torch.bfloat16).contiguous()
        b = torch.zeros(self.cout_pad, dtype=torch.float32)
        if bias is not None:
            b[:cout] = bias.detach().float().cpu()
        self.b = b.to(device)

    @staticmethod
    def fold_bn(conv_w, bn_w, bn_b, bn_mean, bn_var, eps=BN_EPS):
        """w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps) (eval-mode BN, blocks.py:32,38)."""
        s = bn_w.detach().float() / torch.sqrt(bn_var.detach().float() + eps)
        return conv_w.detach().float() * s.view(-1, 1, 1, 1), bn_b.detach().float() - bn_mean.detach().float() * s

    @staticmethod
    def concat(parts: Sequence["PackedConv"]) -> "PackedConv":
        """Stack several convs that read the same input along Cout (CSP cv1||cv2, CLA k||v)."""
        p0 = parts[0]
        assert all(p.cin == p0.cin and p.k == p0.k for p in parts), "fused convs must share input channels and kernel size"
        # only the REAL output rows of every part are stacked (padding rows are re-created at the end), so the
        # parts may have any channel count (skyeye_m: 48 + 48 -> 96 rows padded to 128)
        out = PackedConv.__new__(PackedConv)
        out.cin, out.k = p0.cin, p0.k
        out.cin_real = p0.cin_real
        out.cout = sum(p.cout for p in parts)
        w = torch.cat([p.w[:p.cout] for p in parts], 0)
        b = torch.cat([p.b[:p.cout] for p in parts], 0)
        out.cout_pad = round_cout(out.cout)
        if out.cout_pad != out.cout:
            w = torch.cat([w, torch.zeros((out.cout_pad - out.cout, w.shape[1]), dtype=w.dtype, device=w.device)], 0)
            b = torch.cat([b, torch.zeros(out.cout_pad - out.cout, dtype=b.dtype, device=b.device)], 0)
        out.w, out.b = w.contiguous(), b.contiguous()
        return out


class PackedFocusConv:
    """FocusBlock conv weights [cout, 12, 3, 3] in the row-tap layout of skb_focus_conv_bf16:
    bf16 [cout_pad][ky 3][kx slot 4][channel 16]; slot 3 and channels 12..15 are zero."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], device="cuda"):
        w = weight.detach().float().cpu()
        cout, cin, kh, kw = w.shape
        assert (cin, kh, kw) == (12, 3, 3), w.shape
        self.cout, self.cout_pad = cout, round_cout(cout)
        wp = torch.zeros((self.cout_pad, 3, 4, 16), dtype=torch.float32)
        wp[:cout, :, :3, :12] = w.permute(0, 2, 3, 1)
        self.w = wp.reshape(self.cout_pad, 192).to(device=device, dtype=torch.bfloat16).contiguous()
        b = torch.zeros(self.cout_pad, dtype=torch.float32)
        if bias is not None:
            b[:cout] = bias.detach().float().cpu()
        self.b = b.to(device)


def focus_conv(img: torch.Tensor, pw: PackedFocusConv, y: View, ws: torch.Tensor, act: int = ACT_SILU, stream=None) -> None:
    """FocusBlock.forward on an fp32 (already /255) or uint8 NCHW image: one native call."""
    assert img.dtype in (torch.float32, torch.uint8) and img.is_contiguous() and img.dim() == 4 and img.shape[1] == 3
    n, _, h, w = img.shape
    N.check(N.lib().skb_focus_conv_bf16(img.data_ptr(), SKB_F32 if img.dtype == torch.float32 else N.SKB_U8, n, h, w,
                                        pw.w.data_ptr(), pw.b.data_ptr(), y.ref, pw.cout_pad, act, ws.data_ptr(), ws.numel(),
                                        _stream_ptr() if stream is None else stream), "skb_focus_conv_bf16")


def focus_conv_tiles(frames: torch.Tensor, tiles: torch.Tensor, tile_hw, pw: PackedFocusConv, y: View, ws: torch.Tensor,
                     act: int = ACT_SILU, stream=None) -> None:
    """FocusBlock.forward on windows of larger frames read in place: frames [F,3,FH,FW] (uint8 / fp32), tiles int32 CUDA
    [n,3] rows (frame, y0, x0); image i of the batch is the tile_hw window at tiles[i] (SURVEY.md D8 tiling)."""
    assert frames.dtype in (torch.float32, torch.uint8) and frames.is_contiguous() and frames.dim() == 4 and frames.shape[1] == 3
    assert tiles.dtype == torch.int32 and tiles.is_cuda and tiles.is_contiguous() and tiles.dim() == 2 and tiles.shape[1] == 3
    n = tiles.shape[0]
    N.check(N.lib().skb_focus_conv_tiles_bf16(frames.data_ptr(), SKB_F32 if frames.dtype == torch.float32 else N.SKB_U8,
                                              frames.shape[2], frames.shape[3], tiles.data_ptr(), n, int(tile_hw[0]), int(tile_hw[1]),
                                              pw.w.data_ptr(), pw.b.data_ptr(), y.ref, pw.cout_pad, act, ws.data_ptr(), ws.numel(),
                                              _stream_ptr() if stream is None else stream), "skb_focus_conv_tiles_bf16")


# ----------------------------------------------------------------------------------------------
# eager op wrappers (one native call each)
# ----------------------------------------------------------------------------------------------
def conv2d(x: View, pw: PackedConv, y: View, stride: int = 1, act: int = ACT_SILU, residual: Optional[View] = None,
           upsample2x: bool = False, stream: Optional[int] = None) -> None:
    assert x.c == pw.cin, (x.c, pw.cin)
    rc = N.lib().skb_conv2d_bf16(x.ref, pw.w.data_ptr(), pw.b.data_ptr(), residual.ref if residual is not None else None,
                                 y.ref, pw.cout_pad, pw.k, stride, act, 1 if upsample2x else 0,
                                 _stream_ptr() if stream is None else stream)
    N.check(rc, "skb_conv2d_bf16")


def focus(img: torch.Tensor, y: View, stream=None) -> None:
    assert img.dtype in (torch.float32, torch.uint8) and img.is_contiguous() and img.dim() == 4 and img.shape[1] == 3
    n, _, h, w = img.shape
    fn = N.lib().skb_focus_nchw_f32 if img.dtype == torch.float32 else N.lib().skb_focus_nchw_u8
    N.check(fn(img.data_ptr(), n, h, w, y.ref, _stream_ptr() if stream is None else stream), "skb_focus_nchw")


def maxpool5(x: View, y: View, stream=None) -> None:
    N.check(N.lib().skb_maxpool5_bf16(x.ref, y.ref, _stream_ptr() if stream is None else stream), "skb_maxpool5_bf16")


def spp_pools(x: View, y5: View, y9: View, y13: View, stream=None) -> None:
    N.check(N.lib().skb_spp_pools_bf16(x.ref, y5.ref, y9.ref, y13.ref, _stream_ptr() if stream is None else stream), "skb_spp_pools_bf16")


def cbam(x: View, w0: torch.Tensor, w1: torch.Tensor, w7: torch.Tensor, y: View, ws: torch.Tensor, stream=None) -> None:
    N.check(N.lib().skb_cbam_bf16(x.ref, w0.data_ptr(), w1.data_ptr(), w0.shape[0], w7.data_ptr(), y.ref, ws.data_ptr(),
                                  ws.numel(), _stream_ptr() if stream is None else stream), "skb_cbam_bf16")


def cla_core(q: View, k: View, v: View, o: View, heads: int, scale: float, r2: float, ws: torch.Tensor, stream=None) -> None:
    N.check(N.lib().skb_cla_core_bf16(q.ref, k.ref, v.ref, o.ref, heads, scale, r2, ws.data_ptr(), ws.numel(),
                                      _stream_ptr() if stream is None else stream), "skb_cla_core_bf16")


def layernorm(x: View, gamma: torch.Tensor, beta: torch.Tensor, y: View, eps: float = 1e-5, stream=None) -> None:
    N.check(N.lib().skb_layernorm_bf16(x.ref, gamma.data_ptr(), beta.data_ptr(), eps, y.ref,
                                       _stream_ptr() if stream is None else stream), "skb_layernorm_bf16")


def flash_attn(qkv: View, o: View, heads: int, scale: float, stream=None) -> None:
    N.check(N.lib().skb_flash_attn_bf16(qkv.ref, o.ref, heads, scale, _stream_ptr() if stream is None else stream),
            "skb_flash_attn_bf16")


def window_attn(qkv: View, bias: torch.Tensor, mask: Optional[torch.Tensor], o: View, heads: int, scale: float, stream=None) -> None:
    N.check(N.lib().skb_window_attn_bf16(qkv.ref, bias.data_ptr(), mask.data_ptr() if mask is not None else None,
                                         mask.shape[0] if mask is not None else 0, o.ref, heads, scale,
                                         _stream_ptr() if stream is None else stream), "skb_window_attn_bf16")


def window_attn2d(qkv: View, bias: torch.Tensor, mask: Optional[torch.Tensor], o: View, heads: int, window: int, scale: float,
                  stream=None) -> None:
    """Window attention on unpartitioned feature maps [B,H,W,.]: partition / reverse are the kernel's addressing."""
    N.check(N.lib().skb_window_attn2d_bf16(qkv.ref, bias.data_ptr(), mask.data_ptr() if mask is not None else None,
                                           mask.shape[0] if mask is not None else 0, o.ref, heads, window, scale,
                                           _stream_ptr() if stream is None else stream), "skb_window_attn2d_bf16")


def decode(raws: Sequence[View], na: int, no: int, anchors, in_hw, det: torch.Tensor,
           raw_out: Optional[Sequence[torch.Tensor]] = None, stream=None) -> None:
    L = len(raws)
    arr = (skb_view * L)(*[r._s for r in raws])
    flat = [float(v) for lvl in anchors for a in lvl for v in a]
    anc = (ctypes.c_float * len(flat))(*flat)
    ro = None
    if raw_out is not None:
        ro = (ctypes.c_void_p * L)(*[t.data_ptr() for t in raw_out])
    N.check(N.lib().skb_decode_f32(arr, L, na, no, anc, int(in_hw[0]), int(in_hw[1]), det.data_ptr(), ro,
                                   _stream_ptr() if stream is None else stream), "skb_decode_f32")


def workspace(nbytes: int, device="cuda") -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------------------------
# launch plan
# ----------------------------------------------------------------------------------------------
class Plan:
    """Ordered native launches over statically allocated buffers for one input shape."""

    def __init__(self, device="cuda"):
        self.device = device
        self.steps: List[Callable[[int], None]] = []
        self.names: List[str] = []
        self.meta: List[dict] = []  # per step: kind, algorithmic flops / bytes, kernel launches
        self.outs: List[list] = []  # per step: what it writes, labelled with the reference module path (teacher forcing)
        self.keep = []  # buffers / packed weights kept alive
        self.graph = None
        self.n_launch_calls = 0
        self.tile_src = None  # (frames holder [tensor], tile table int32 CUDA [n,3]) when the images are windows of larger frames

    def buf(self, n, h, w, c, dtype=torch.bfloat16) -> View:
        v = new_buffer(n, h, w, c, dtype, self.device)
        self.keep.append(v)
        return v

    def ws(self, nbytes: int) -> torch.Tensor:
        t = workspace(nbytes, self.device)
        self.keep.append(t)
        return t

    def add(self, name: str, fn: Callable[[int], None], kind: str = "other", flops: float = 0.0, bytes: float = 0.0,
            launches: int = 1, outs: Optional[list] = None) -> None:
        """``outs``: [dict(view=View | Tensor, label=<reference state-dict path of the producing module>,
        layout='nhwc' | 'raw' | 'flat', up2=bool, na=int)] -- consumed by ``run_teacher_forced``."""
        self.names.append(name)
        self.steps.append(fn)
        self.meta.append(dict(kind=kind, flops=float(flops), bytes=float(bytes), launches=int(launches)))
        self.outs.append([o for o in (outs or []) if o.get("label")])

    # -- op builders ---------------------------------------------------------------------------
    def conv(self, name, x: View, pw: PackedConv, y: View, stride=1, act=ACT_SILU, residual=None, upsample2x=False, label=None):
        """``label``: reference module path of the output, or [(c0, c1, path)] for a fused GEMM writing several
        modules' outputs into channel slices of ``y``."""
        self.keep.append(pw)
        if isinstance(label, str):
            outs = [dict(view=y.slice(0, min(pw.cout, y.c)), label=label, up2=upsample2x)]
        else:
            outs = [dict(view=y.slice(c0, c1), label=lb, up2=upsample2x) for c0, c1, lb in (label or [])]
        ho, wo = x.h // stride, x.w // stride
        m = x.n * ho * wo
        flops = 2.0 * m * pw.cout * pw.k * pw.k * pw.cin_real  # algorithmic 2*M*N*K, unpadded
        esz = 4 if y.dtype == SKB_F32 else 2
        nbytes = 2.0 * x.n * x.h * x.w * x.c + 2.0 * pw.w.numel() + esz * m * y.c * (4 if upsample2x else 1) + \
            (2.0 * m * y.c if residual is not None else 0.0)
        self.add(name, lambda s: conv2d(x, pw, y, stride, act, residual, upsample2x, s), "conv", flops, nbytes, 1, outs)
        return y

    def run(self, stream: Optional[int] = None) -> None:
        s = _stream_ptr() if stream is None else stream
        for fn in self.steps:
            fn(s)

    @property
    def launches(self) -> int:
        return sum(m["launches"] for m in self.meta)

    def run_timed(self):
        """Per-step CUDA-event timing on the launching stream: list of (name, ms)."""
        evs = []
        s = _stream_ptr()
        for fn in self.steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(s)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return [(n, a.elapsed_time(b)) for n, (a, b) in zip(self.names, evs)]

    @staticmethod
    def _expected(o: dict, taps: dict) -> torch.Tensor:
        t = taps[o["label"]]
        layout = o.get("layout", "nhwc")
        if layout == "flat":
            return t
        if layout == "raw":  # DetectionHead.forward's view + permute (detector.py:81-82)
            b, c, h, w = t.shape
            na = o["na"]
            return t.view(b, na, c // na, h, w).permute(0, 1, 3, 4, 2)
        if o.get("up2"):  # nearest 2x upsample fused into the producer's stores (detector.py:214,218)
            t = t.repeat_interleave(2, 2).repeat_interleave(2, 3)
        return t.permute(0, 2, 3, 1)

    def run_teacher_forced(self, taps: dict) -> List[dict]:
        """Verification mode: run the launches one by one; after each, compare what it wrote with ``taps[label]``
        (NCHW fp32 host tensors keyed by reference module path, produced by an external reference run on the same
        input and weights) and then OVERWRITE the output with the reference values, so every launch reads exactly
        the reference's input for it.  Per-launch errors therefore do not chain.  Returns one row per output."""
        rows = []
        s = _stream_ptr()
        for name, fn, outs, meta in zip(self.names, self.steps, self.outs, self.meta):
            fn(s)
            torch.cuda.synchronize()
            for o in outs:
                got = o["view"].torch() if isinstance(o["view"], View) else o["view"]
                exp = self._expected(o, taps).to(got.device, torch.float32)
                assert tuple(got.shape) == tuple(exp.shape), (name, o["label"], tuple(got.shape), tuple(exp.shape))
                d = (got.float() - exp).abs()
                if got.dtype == torch.bfloat16:
                    # what is left after allowing the stored value to land on the neighbouring bf16 number (a 1e-7 difference
                    # in the fp32 result flips the rounding of a value that sits on a rounding boundary)
                    ulp = torch.exp2(torch.floor(torch.log2(exp.abs().clamp_min(1e-30))) - 7.0)
                    beyond = float((d - ulp).clamp_min(0).max())
                else:
                    beyond = float(d.max())
                rows.append(dict(step=name, label=o["label"], kind=meta["kind"], store="f32" if got.dtype == torch.float32 else "bf16",
                                 max_err=float(d.max()), beyond_ulp=beyond, rms_err=float(d.pow(2).mean().sqrt()),
                                 ref_max=float(exp.abs().max()), ref_rms=float(exp.pow(2).mean().sqrt()), numel=exp.numel()))
                got.copy_(exp)
        return rows

    def capture(self) -> None:
        """Capture the whole plan into a CUDA graph (tensor maps are baked in as kernel params)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run()
        self.graph = g

    def replay(self) -> None:
        if self.graph is None:
            self.run()
        else:
            self.graph.replay()
