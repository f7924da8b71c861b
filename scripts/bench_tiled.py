"""BASELINE config 4: 4K drone frames (3840x2160) sliced into overlapping 1280x1280 tiles, tile-sharded across
the ranks of one box with no collective in the forward pass, NCCL all_gather of the padded per-tile detections,
cross-tile merge NMS.  Launch: python scripts/bench_tiled.py  (1 GPU)  or
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_tiled.py
Prints one JSON line with frames/s, tiles/s and a checksum of the merged detections (identical for every N)."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")]
from skyeye.core.detector import construct_model  # noqa: E402
from skyeye.utils import tiling  # noqa: E402
from skyeye.utils.nms import batched_nms_padded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--variant", default="skyeye_l")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = construct_model(f"{a.variant}.yaml").to(dev).eval()
    model.reuse_output_buffers = True
    nc = model.cfg["nc"]
    rng = np.random.Generator(np.random.PCG64(1234))
    frames = torch.from_numpy(rng.integers(0, 256, (a.frames, 3, 2160, 3840), dtype=np.uint8)).to(dev)
    detect = lambda tiles: model(tiles)[0]

    def step():
        return tiling.tiled_detect(frames, detect, batched_nms_padded, nc, rank, world)

    rows, cnt = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        rows, cnt = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    valid = (torch.arange(rows.shape[1], device=dev)[None, :] < cnt[:, None]).float()[:, :, None]
    digest = hashlib.sha256((rows * valid).cpu().numpy().tobytes() + cnt.cpu().numpy().tobytes()).hexdigest()[:16]
    if rank == 0:
        tiles = a.frames * len(tiling.tile_origins())
        print(json.dumps({"config": "4K frames -> 1280^2 tiles, tile-sharded, NCCL gather + merge NMS", "variant": a.variant,
                          "n_gpus": world, "frames": a.frames, "tiles": tiles, "ms_per_step": ms, "frames_per_s": a.frames / ms * 1e3,
                          "tiles_per_s": tiles / ms * 1e3, "detections": int(cnt.sum()), "sha256_16": digest}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
