"""Window-attention core (skb_window_attn2d_bf16) on the three levels of skyeye_lw at 1280^2 B16, CUDA events on the launching
stream, L2 flushed between iterations.  SKB_WATT_TC=0 selects the CUDA-core kernel.  Algorithmic bytes = qkv read once + o
written once; flops = 4 * 64 * 64 * 64 per (window, head)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")]
from skyeye import engine as E  # noqa: E402

LEVELS = {"p3": (16, 160, 160, 256, 4), "p4": (16, 80, 80, 512, 8), "p5": (16, 40, 40, 1024, 16)}


def main():
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for name, (B, H, W, C, heads) in LEVELS.items():
        qkv = E.View(torch.randn((B, H, W, 3 * C), device="cuda").to(torch.bfloat16))
        bias = torch.randn((heads, 64, 64), device="cuda")
        o = E.new_buffer(B, H, W, C)
        for _ in range(3):
            E.window_attn2d(qkv, bias, None, o, heads, 8, 0.125)
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            E.window_attn2d(qkv, bias, None, o, heads, 8, 0.125)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        byt = B * H * W * 4 * C * 2
        fl = B * (H // 8) * (W // 8) * heads * 4 * 64 * 64 * 64
        print(f"window_attn {name}: {ms:.4f} ms  {byt / ms / 1e6:8.1f} GB/s  {fl / ms / 1e9:7.1f} TF/s  (TC={os.environ.get('SKB_WATT_TC', '1')})")


if __name__ == "__main__":
    main()
