"""Per-kernel micro-benchmark (GPU): the conv shape classes of SURVEY.md §8(a) for skyeye_l@1280 B16 and
the three attention levels, timed with CUDA events on the launching stream (inputs >> L2 or L2
flushed between iterations).  --once runs each selected shape exactly once (for ncu)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")]
from skyeye import engine as E  # noqa: E402

# name: (B, H, W, Cin, Cout, k, stride, residual)
CONV = {
    "c3x3_focus_16_64_640": (16, 640, 640, 16, 64, 3, 1, False),
    "c3x3_128_160": (16, 160, 160, 128, 128, 3, 1, True),
    "c3x3_256_80": (16, 80, 80, 256, 256, 3, 1, True),
    "c3x3_512_40": (16, 40, 40, 512, 512, 3, 1, True),
    "c3x3_64_320": (16, 320, 320, 64, 64, 3, 1, True),
    "c3x3_64_320_nores": (16, 320, 320, 64, 64, 3, 1, False),
    "c3x3s2_64_128": (16, 640, 640, 64, 128, 3, 2, False),
    "c3x3s2_128_256": (16, 320, 320, 128, 256, 3, 2, False),
    "c1x1_128_128_320": (16, 320, 320, 128, 128, 1, 1, False),
    "c1x1_64_64_320": (16, 320, 320, 64, 64, 1, 1, False),
    "c1x1_128_128_160": (16, 160, 160, 128, 128, 1, 1, False),
    "c1x1_256_256_160": (16, 160, 160, 256, 256, 1, 1, False),
    "c1x1_256_256_80": (16, 80, 80, 256, 256, 1, 1, False),
    "c1x1_512_512_80": (16, 80, 80, 512, 512, 1, 1, False),
    "c1x1_1024_512_40": (16, 40, 40, 1024, 512, 1, 1, False),
    "c1x1_2048_1024_40": (16, 40, 40, 2048, 1024, 1, 1, False),
    "qkv_256_768_160": (16, 160, 160, 256, 768, 1, 1, False),
    "ff0_256_1024_160": (16, 160, 160, 256, 1024, 1, 1, False),
    "ff3_1024_256_160": (16, 160, 160, 1024, 256, 1, 1, True),
}
ATTN = {"attn_p3": (2, 160, 160, 4), "attn_p4": (16, 80, 80, 8), "attn_p5": (16, 40, 40, 16)}


def flush_l2(buf):
    buf.zero_()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    sel = [s for s in a.only.split(",") if s]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = {}
    for name, (B, H, W, ci, co, k, s, r) in CONV.items():
        if sel and name not in sel:
            continue
        x = E.View(torch.randn((B, H, W, ci), device="cuda").to(torch.bfloat16))
        wt = torch.randn((co, ci, k, k)) * (2.0 / (ci * k * k)) ** 0.5
        pw = E.PackedConv(wt, torch.zeros(co))
        y = E.new_buffer(B, H // s, W // s, co)
        y.t.zero_()
        rv = y if r else None
        fl = 2.0 * B * (H // s) * (W // s) * co * ci * k * k
        by = 2.0 * (x.t.numel() + y.t.numel() * (2 if r else 1) + pw.w.numel())
        if a.once:
            E.conv2d(x, pw, y, s, 1, rv)
            torch.cuda.synchronize()
            continue
        for _ in range(2):
            E.conv2d(x, pw, y, s, 1, rv)
        ts = []
        for _ in range(a.iters):
            flush_l2(flush)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            E.conv2d(x, pw, y, s, 1, rv)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        res[name] = dict(ms=ms, tflops=fl / ms / 1e9, gbs=by / ms / 1e6)
        print(f"{name:22s} {ms:8.3f} ms {fl / ms / 1e9:8.1f} TF/s {by / ms / 1e6:8.0f} GB/s", flush=True)
    for name, (B, H, W, heads) in ATTN.items():
        if sel and name not in sel:
            continue
        C = heads * 64
        qkv = E.View(torch.randn((B, H, W, 3 * C), device="cuda").to(torch.bfloat16))
        o = E.new_buffer(B, H, W, C)
        N = H * W
        fl = 4.0 * B * N * N * C
        if a.once:
            E.flash_attn(qkv, o, heads, 0.125)
            torch.cuda.synchronize()
            continue
        for _ in range(2):
            E.flash_attn(qkv, o, heads, 0.125)
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            E.flash_attn(qkv, o, heads, 0.125)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        res[name] = dict(ms=ms, tflops=fl / ms / 1e9)
        print(f"{name:22s} {ms:8.3f} ms {fl / ms / 1e9:8.1f} TF/s", flush=True)
    # head: windowed: the per-window attention core on the unpartitioned qkv map, window 8, head_dim 64
    for name, (B, H, W, heads) in {"wattn_p3": (16, 160, 160, 4), "wattn_p4": (16, 80, 80, 8), "wattn_p5": (16, 40, 40, 16)}.items():
        if sel and name not in sel:
            continue
        C = heads * 64
        qkv = E.View(torch.randn((B, H, W, 3 * C), device="cuda").to(torch.bfloat16))
        o = E.new_buffer(B, H, W, C)
        bias = torch.randn((heads, 64, 64), device="cuda") * 0.02
        for _ in range(2):
            E.window_attn2d(qkv, bias, None, o, heads, 8, 0.125)
        ts = []
        for _ in range(a.iters):
            flush_l2(flush)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            E.window_attn2d(qkv, bias, None, o, heads, 8, 0.125)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        by = 2.0 * (qkv.t.numel() + o.t.numel())
        res[name] = dict(ms=ms, tflops=4.0 * B * H * W * 64 * C / ms / 1e9, gbs=by / ms / 1e6)
        print(f"{name:22s} {ms:8.3f} ms {res[name]['tflops']:8.1f} TF/s {by / ms / 1e6:8.0f} GB/s", flush=True)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
