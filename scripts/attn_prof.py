"""Cycle-level phase profile of flash_attn_kernel (run with SKB_ATT_PROF=1): where each role warp spends an iteration."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")]
from skyeye import _native as N  # noqa: E402
from skyeye import engine as E  # noqa: E402

NAMES = ["softmax wait S", "softmax TMEM load", "softmax exp/max/pack", "softmax TMEM store+arrive", "softmax iters",
         "MMA wait P", "MMA issue PV", "MMA wait K/V", "MMA issue QK", "MMA iters", "TMA wait empty", "TMA iters",
         "hop P-arrive -> MMA awake", "hop S-commit -> softmax awake"]


def main():
    B, H, W, heads = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (2, 160, 160, 4)))
    C = heads * 64
    qkv = E.View(torch.randn((B, H, W, 3 * C), device="cuda").to(torch.bfloat16))
    o = E.new_buffer(B, H, W, C)
    L = N.lib()
    buf = (ctypes.c_ulonglong * 16)()
    E.flash_attn(qkv, o, heads, 0.125)
    torch.cuda.synchronize()
    L.skb_debug_attn_prof(ctypes.addressof(buf), 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    E.flash_attn(qkv, o, heads, 0.125)
    e1.record()
    torch.cuda.synchronize()
    L.skb_debug_attn_prof(ctypes.addressof(buf), 1)
    v = list(buf)
    print(f"B={B} N={H * W} heads={heads}: {e0.elapsed_time(e1):.3f} ms  (instrumented variant, SKB_ATT_PROF=1)")
    its = {0: v[4], 1: v[4], 2: v[4], 3: v[4], 5: v[9], 6: v[9], 7: v[9], 8: v[9], 10: v[11], 12: v[9], 13: v[4]}
    for i, n in enumerate(NAMES):
        if i in its and its[i]:
            print(f"  {n:34s} {v[i] / its[i]:9.1f} cycles / iteration")


if __name__ == "__main__":
    main()
