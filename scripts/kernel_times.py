"""Per-KERNEL device times of one forward plan + NMS (torch.profiler / CUPTI; no ncu needed): the plan's steps that consist of
several launches (Focus = pad + conv, CLA core = score + colstat + apply, CBAM = 4 kernels, NMS = filter + sort + kept-list)
are split into their kernels.  usage: python scripts/kernel_times.py [variant] [size] [batch]"""
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


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "skyeye_l"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    cfg = om.get_cfg(variant)
    m = construct_model(f"{variant}.yaml")
    m.load_state_dict(om.make_calibrated_state_dict(cfg, 0), strict=True)
    m = m.cuda().eval()
    m.use_cuda_graph = False
    x = torch.from_numpy(np.random.Generator(np.random.PCG64(1234)).integers(0, 256, (batch, 3, size, size), dtype=np.uint8)).cuda()
    for _ in range(3):
        det, _ = m(x)
        batched_nms_padded(det, 0.25, 0.45)
    torch.cuda.synchronize()
    reps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            det, _ = m(x)
            batched_nms_padded(det, 0.25, 0.45)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.split("(")[0].replace("void ", "").replace("skb::", "")
            agg[name][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"{variant} {size}x{size} B{batch}: {tot / reps / 1e3:.3f} ms of kernel time per step")
    for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{us / reps / 1e3:9.4f} ms {100 * us / tot:5.1f} %  {n // reps:4d} launches  {k[:110]}")


if __name__ == "__main__":
    main()
