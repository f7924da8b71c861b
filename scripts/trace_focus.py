"""Device-side timeline of CTA 0 of the fused Focus conv (debug): per-role event deltas, plus its launch times."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")]
from skyeye import engine as E
from skyeye import _native as N

B, H, W, co = 16, 1280, 1280, 64
img = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device="cuda")
pw = E.PackedFocusConv(torch.randn((co, 12, 3, 3)) * 0.1, torch.zeros(co))
y = E.new_buffer(B, H // 2, W // 2, co)
ws = torch.empty(int(N.lib().skb_focus_conv_workspace_bytes(B, H, W)) + 256, dtype=torch.uint8, device="cuda")
for _ in range(2):
    E.focus_conv(img, pw, y, ws)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); E.focus_conv(img, pw, y, ws); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("focus_conv (pad + conv) ms:", sorted(ts))
tr = torch.zeros(3 * 8192, dtype=torch.int64, device="cuda")
N.lib().skb_debug_conv_trace(tr.data_ptr())
E.focus_conv(img, pw, y, ws)
torch.cuda.synchronize()
N.lib().skb_debug_conv_trace(None)
t = tr.cpu().view(3, 4096, 2)
t0 = min(int(t[r_, 0, 1]) for r_ in range(3) if int(t[r_, 0, 1]) > 0)
for role, nm in enumerate(("producer", "mma", "epilogue")):
    ev = [(int(a), int(b) - t0) for a, b in t[role].tolist() if b > 0]
    print(f"--- {nm}: {len(ev)} events; steady state events 200..240:")
    if len(ev) > 240:
        prev = ev[199][1]
        out = []
        for e, c in ev[200:260]:
            out.append(f"{e}@{c}(+{c - prev})")
            prev = c
        print(" ".join(out))
    print("  total span", ev[-1][1] - ev[0][1] if ev else 0, "events", len(ev))
