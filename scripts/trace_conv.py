"""Device-side timeline of CTA 0 of one conv launch (debug): prints per-role event deltas."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200"), os.path.join(ROOT, "scripts")]
from skyeye import engine as E
from skyeye import _native as N
from bench_layers import CONV

name = sys.argv[1] if len(sys.argv) > 1 else "c1x1_128_128_320"
B, H, W, ci, co, k, s, r = CONV[name]
x = E.View(torch.randn((B, H, W, ci), device="cuda").to(torch.bfloat16))
pw = E.PackedConv(torch.randn((co, ci, k, k)) * 0.05, torch.zeros(co))
y = E.new_buffer(B, H // s, W // s, co); y.t.zero_()
rv = y if r else None
for _ in range(2):
    E.conv2d(x, pw, y, s, 1, rv)
tr = torch.zeros(3 * 8192, dtype=torch.int64, device="cuda")
N.lib().skb_debug_conv_trace(tr.data_ptr())
E.conv2d(x, pw, y, s, 1, rv)
torch.cuda.synchronize()
N.lib().skb_debug_conv_trace(None)
t = tr.cpu().view(3, 4096, 2)
t0 = min(int(t[r_, 0, 1]) for r_ in range(3) if int(t[r_, 0, 1]) > 0)
for role, nm in enumerate(("producer", "mma", "epilogue")):
    ev = [(int(a), int(b) - t0) for a, b in t[role].tolist() if b > 0]
    print(f"--- {nm}: {len(ev)} events; first 70:")
    prev = 0
    out = []
    for e, c in ev[:70]:
        out.append(f"{e}@{c}(+{c - prev})")
        prev = c
    print(" ".join(out))
    if len(ev) > 400:
        print("    ... steady state events 300..340:")
        prev = ev[299][1]
        out = []
        for e, c in ev[300:340]:
            out.append(f"{e}@{c}(+{c - prev})")
            prev = c
        print(" ".join(out))
