"""Idle time BETWEEN the kernels of one step (CUDA-graph replay of the forward plan + NMS), from CUPTI start / end timestamps
(torch.profiler): span of a step, sum of kernel durations, and the gaps start[i+1] - end[i] grouped by the kernel that
precedes the gap (negative = the next kernel's prologue overlapped under programmatic dependent launch).
usage: python scripts/step_gaps.py [variant] [size] [batch]"""
import os
import sys
from collections import defaultdict

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200"), os.path.join(ROOT, "tests", "golden")]
from oracle import model as om  # noqa: E402  (weight recipe only)
from skyeye.core.detector import construct_model  # noqa: E402
from skyeye.utils.nms import batched_nms_padded  # noqa: E402


def short(n):
    return n.split("(")[0].replace("void ", "").replace("skb::", "")[:60]


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "skyeye_l"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    cfg = om.get_cfg(variant)
    m = construct_model(f"{variant}.yaml")
    m.load_state_dict(om.make_calibrated_state_dict(cfg, 0), strict=True)
    m = m.cuda().eval()
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(1234)).integers(0, 256, (batch, 3, size, size), dtype=np.uint8)).cuda()
    for _ in range(4):
        det, _ = m(x)
        batched_nms_padded(det, 0.25, 0.45)
    torch.cuda.synchronize()
    reps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            det, _ = m(x)
            batched_nms_padded(det, 0.25, 0.45)
        torch.cuda.synchronize()
    evs = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in ev.name.lower() and "memset" not in ev.name.lower():
            evs.append((ev.time_range.start, ev.time_range.end, short(ev.name)))
    evs.sort()
    span = evs[-1][1] - evs[0][0]
    ksum = sum(e - s for s, e, _ in evs)
    gaps = defaultdict(lambda: [0.0, 0, 0.0])
    pos = neg = 0.0
    for (s0, e0, n0), (s1, e1, n1) in zip(evs, evs[1:]):
        g = s1 - e0
        if g > 1000:   # the boundary between two profiled steps (host launch of the next replay)
            continue
        key = n0 + " -> " + n1
        gaps[key][0] += g
        gaps[key][1] += 1
        pos += max(g, 0.0)
        neg += min(g, 0.0)
    print(f"{variant} {size} B{batch}: {len(evs) // reps} kernels/step, span {span / reps / 1e3:.3f} ms/step (incl. replay boundaries), "
          f"kernel time {ksum / reps / 1e3:.3f} ms, idle gaps {pos / reps / 1e3:.3f} ms, overlap {neg / reps / 1e3:.3f} ms")
    for k, (us, n, _) in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{us / reps:9.2f} us/step {n // reps:4d} x {us / max(n, 1):7.2f} us  {k}")


if __name__ == "__main__":
    main()
