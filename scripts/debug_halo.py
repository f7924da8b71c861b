"""Diagnostic for the halo-tile conv: single-tap weights, error per tap and error map."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")]
from skyeye import engine as E

torch.manual_seed(0)
n, h, w, ci, co = 1, 16, 16, 64, 64
x = torch.randn(n, ci, h, w).to(torch.bfloat16).float()
xv = E.from_nchw(x.cuda())
for tap in range(9):
    wt = torch.zeros(co, ci, 3, 3)
    wt[:, :, tap // 3, tap % 3] = (torch.randn(co, ci) * 0.1).to(torch.bfloat16).float()
    pw = E.PackedConv(wt, torch.zeros(co))
    y = E.new_buffer(n, h, w, co)
    y.t.zero_()
    E.conv2d(xv, pw, y, 1, 0)
    torch.cuda.synchronize()
    ref = F.conv2d(x, wt, None, 1, 1)
    got = y.nchw().float().cpu()
    err = (got - ref).abs().amax(dim=1)[0]  # [h, w]
    print(f"tap {tap} (dy={tap//3}, dx={tap%3}): max err {float(err.max()):.4f} ref max {float(ref.abs().max()):.3f}")
    if float(err.max()) > 0.05:
        bad = (err > 0.05).int()
        print("  bad pixel map (rows = y):")
        for r in range(h):
            print("   ", "".join(str(int(v)) for v in bad[r]))
