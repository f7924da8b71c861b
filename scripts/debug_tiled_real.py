"""Real multi-GPU check of config 4 (torchrun, world >= 2): every rank also computes all tiles locally (world 1) and compares
the gathered per-tile rows and the merged result with it."""
import os, sys, hashlib
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200"), os.path.join(ROOT, "tests", "golden")]
from oracle import model as om
from skyeye.core.detector import construct_model
from skyeye.utils import tiling

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = om.get_cfg("skyeye_l")
m = construct_model("skyeye_l.yaml")
m.load_state_dict(om.make_calibrated_state_dict(cfg, 0), strict=True)
m = m.to(dev).eval()
NF, FH, FW = 16, 2160, 3840
frames = torch.from_numpy(np.random.Generator(np.random.PCG64(4242)).integers(0, 256, (NF, 3, FH, FW), dtype=np.uint8)).to(dev)
def sha(r, c):
    return hashlib.sha256(r.cpu().numpy().tobytes() + c.cpu().numpy().tobytes()).hexdigest()[:16]
td1 = tiling.TiledDetector(m, NF, (FH, FW), rank=0, world=1, compat="fixed", overlap=False)
r1, c1 = td1(frames)
r1, c1, send1 = r1.clone(), c1.clone(), td1.send[0].clone()
tdw = tiling.TiledDetector(m, NF, (FH, FW), rank=rank, world=world, compat="fixed", overlap=True)
for k in range(3):
    rw, cw = tdw(frames)
torch.cuda.synchronize()
slot = (tdw.steps - 1) & 1
g = tdw.gath[slot]
bad = [t for t in range(NF * tdw.T) if not torch.equal(send1[t], g[t % world, t // world])]
print(f"rank {rank}: world-1 sha {sha(r1, c1)}  world-{world} sha {sha(rw, cw)}  merged equal {torch.equal(rw, r1)}  gathered tiles differing {len(bad)} {bad[:8]}", flush=True)
if bad:
    t = bad[0]
    a, b = send1[t], g[t % world, t // world]
    d = (a != b).nonzero()
    print(f"rank {rank}: tile {t} (owner rank {t % world}) first diffs {d[:5].tolist()} a {a[d[0][0]].tolist()} b {b[d[0][0]].tolist()}", flush=True)
dist.destroy_process_group()
