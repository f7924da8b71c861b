"""Single-GPU emulation of config 4 at world 2 / 8 against world 1: the merged detections must be identical."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200"), os.path.join(ROOT, "tests", "golden")]
from oracle import model as om
from skyeye.core.detector import construct_model
from skyeye.utils import tiling
import torch.distributed as dist

variant = "skyeye_l"
cfg = om.get_cfg(variant)
m = construct_model(f"{variant}.yaml")
m.load_state_dict(om.make_calibrated_state_dict(cfg, 0), strict=True)
m = m.cuda().eval()
NF, FH, FW = 16, 2160, 3840
frames = torch.from_numpy(np.random.Generator(np.random.PCG64(4242)).integers(0, 256, (NF, 3, FH, FW), dtype=np.uint8)).cuda()

td1 = tiling.TiledDetector(m, NF, (FH, FW), rank=0, world=1, compat="fixed", overlap=False)
r1, c1 = td1(frames)
r1, c1 = r1.clone(), c1.clone()
send1 = td1.send[0].clone()
for W in (2, 8):
    tds = [tiling.TiledDetector(m, NF, (FH, FW), rank=r, world=W, compat="fixed", overlap=False) for r in range(W)]
    # forward + per-tile NMS of every emulated rank (the merge is run by hand below)
    for td in tds:
        main = torch.cuda.current_stream()
        for i0 in range(0, td.n_local, td.max_batch):
            n = min(td.max_batch, td.n_local - i0)
            det, _ = m.forward_tiles(frames, td.table[i0:i0 + n], (td.tile, td.tile))
            td._tile_nms(det, i0, n, 0, main.cuda_stream)
    torch.cuda.synchronize()
    gathered = torch.stack([td.send[0] for td in tds])          # what all_gather_into_tensor delivers: rank-major
    # compare the per-tile rows with world 1 (global tile g = local index g // W of rank g % W)
    bad = 0
    for g in range(NF * tds[0].T):
        a, b = send1[g], gathered[g % W, g // W]
        if not torch.equal(a, b):
            bad += 1
            if bad <= 3:
                d = (a != b).nonzero()
                print(f"  world {W}: tile {g} rows differ at {d[:4].tolist()} count {a[300,0].item()} vs {b[300,0].item()}")
    td = tds[0]
    td.gath[0].copy_(gathered)
    N = td.N
    st = torch.cuda.current_stream().cuda_stream
    N.check(N.lib().skb_tile_merge_pred_f32(td.gath[0].data_ptr(), W, td.n_local_max, td.F, td.T, td.max_det, td.nc, td.compat, td.pred.data_ptr(), st), "merge")
    N.check(N.lib().skb_nms_batched_f32(td.pred.data_ptr(), td.F, td.T * td.max_det, td.nc, td.conf, td.iou, None, 0, 0, 0, td.max_det, td.compat,
                                        td.rows[0].data_ptr(), td.cnt[0].data_ptr(), td.ws_merge.data_ptr(), td.ws_merge.numel(), st), "nms")
    torch.cuda.synchronize()
    print(f"world {W}: per-tile rows differing {bad} / {NF * td.T}; merged rows equal {torch.equal(td.rows[0], r1)}, counts equal {torch.equal(td.cnt[0], c1)}")
    if not torch.equal(td.rows[0], r1):
        d = (td.rows[0] != r1).nonzero()
        print("   first merged differences:", d[:6].tolist())

import hashlib
def sha(r, c):
    return hashlib.sha256(r.cpu().numpy().tobytes() + c.cpu().numpy().tobytes()).hexdigest()[:16]
print("world 1, no overlap, one call:", sha(r1, c1))
td = tiling.TiledDetector(m, NF, (FH, FW), rank=0, world=1, compat="fixed", overlap=True)
for k in range(5):
    slot = td.step(frames)
    torch.cuda.current_stream().wait_event(td.ev_done[slot])
    torch.cuda.synchronize()
    r, c = td.result(slot)
    print("world 1, overlap, step", k, "slot", slot, sha(r, c), "equal to one-call:", torch.equal(r, r1))
# back-to-back steps without host sync (the bench loop)
td = tiling.TiledDetector(m, NF, (FH, FW), rank=0, world=1, compat="fixed", overlap=True)
for k in range(4):
    slot = td.step(frames)
torch.cuda.synchronize()
for sl in (0, 1):
    r, c = td.result(sl)
    print("pipelined, slot", sl, sha(r, c), torch.equal(r, r1), "diff rows:", int((r != r1).any(dim=2).sum()))
