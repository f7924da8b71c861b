"""Sweep of the generic conv kernel's (BN, NCTA) choice over the layer shapes of skyeye_l @1280 B16, each timed
alone with CUDA events and an L2 flush: prints the default pick (cost model in skb_conv2d_bf16) beside the best forced pick
(SKB_CONV_FORCE="BN,NCTA").  Used to re-fit the cost model after the warp-uniform issue change."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")]
from skyeye import engine as E  # noqa: E402

# (count in the network, Cin, Cout, k, stride, Hin, residual)
SHAPES = [
    (3, 64, 64, 1, 1, 320, False), (2, 128, 64, 1, 1, 320, False), (1, 128, 128, 1, 1, 320, False), (1, 64, 128, 3, 2, 640, False),
    (12, 128, 128, 1, 1, 160, False), (2, 256, 128, 1, 1, 160, False), (2, 256, 256, 1, 1, 160, False), (2, 512, 128, 1, 1, 160, False),
    (1, 128, 256, 3, 2, 320, False), (15, 256, 256, 1, 1, 80, False), (3, 512, 256, 1, 1, 80, False), (3, 512, 512, 1, 1, 80, False),
    (2, 768, 256, 1, 1, 80, False), (2, 1024, 256, 1, 1, 80, False), (1, 256, 512, 3, 2, 160, False), (1, 256, 256, 3, 2, 160, False),
    (6, 512, 512, 1, 1, 40, False), (4, 1024, 512, 1, 1, 40, False), (2, 1024, 1024, 1, 1, 40, False), (2, 1536, 512, 1, 1, 40, False),
    (1, 2048, 1024, 1, 1, 40, False), (1, 512, 1024, 3, 2, 80, False), (6, 512, 512, 3, 1, 40, True), (1, 512, 512, 3, 2, 80, False),
    (1, 256, 768, 1, 1, 160, False), (1, 256, 1024, 1, 1, 160, False), (1, 1024, 256, 1, 1, 160, True), (1, 256, 256, 1, 1, 160, True),
    (1, 512, 1536, 1, 1, 80, False), (1, 512, 2048, 1, 1, 80, False), (1, 2048, 512, 1, 1, 80, True), (1, 512, 512, 1, 1, 80, True),
]


def time_conv(x, pw, y, s, rv, flush, iters=5):
    for _ in range(2):
        E.conv2d(x, pw, y, s, 1, rv)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        E.conv2d(x, pw, y, s, 1, rv)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    B = 16
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    tot_def = tot_best = 0.0
    out = []
    for cnt, ci, co, k, s, hin, res in SHAPES:
        x = E.View(torch.randn((B, hin, hin, ci), device="cuda").to(torch.bfloat16))
        pw = E.PackedConv(torch.randn((co, ci, k, k)) * (2.0 / (ci * k * k)) ** 0.5, torch.zeros(co))
        y = E.new_buffer(B, hin // s, hin // s, co)
        y.t.zero_()
        rv = y if res else None
        os.environ.pop("SKB_CONV_FORCE", None)
        t_def = time_conv(x, pw, y, s, rv, flush)
        best = (t_def, "default")
        row = {}
        for bn in (256, 128, 64, 32):
            if co % bn:
                continue
            for nc in (1, 2):
                if nc == 2 and bn < 64:
                    continue
                os.environ["SKB_CONV_FORCE"] = f"{bn},{nc}"
                t = time_conv(x, pw, y, s, rv, flush)
                row[f"{bn},{nc}"] = round(t, 4)
                if t < best[0]:
                    best = (t, f"{bn},{nc}")
        os.environ.pop("SKB_CONV_FORCE", None)
        tot_def += cnt * t_def
        tot_best += cnt * best[0]
        fl = 2.0 * B * (hin // s) ** 2 * co * ci * k * k
        print(f"{cnt:2d}x {ci:4d}->{co:4d} k{k} s{s} @{hin // s:3d} default {t_def:.4f} ms ({fl / t_def / 1e9:6.0f} TF/s)  best {best[1]:8s} {best[0]:.4f} ms  "
              + " ".join(f"{k2}:{v:.4f}" for k2, v in row.items()), flush=True)
        out.append(dict(cnt=cnt, cin=ci, cout=co, k=k, s=s, hout=hin // s, default_ms=t_def, best=best[1], best_ms=best[0], sweep=row))
    print(f"sum over the network: default {tot_def:.3f} ms, best-per-layer {tot_best:.3f} ms")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_conv.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
