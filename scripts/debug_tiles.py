"""Does a tile's result depend on which other tiles share its batch?  (config 4 must not depend on the world size)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200"), os.path.join(ROOT, "tests", "golden")]
from oracle import model as om
from skyeye.core.detector import construct_model
from skyeye.utils.tiling import tile_origins

variant = sys.argv[1] if len(sys.argv) > 1 else "skyeye_l"
cfg = om.get_cfg(variant)
m = construct_model(f"{variant}.yaml")
m.load_state_dict(om.make_calibrated_state_dict(cfg, 0), strict=True)
m = m.cuda().eval()
NF, FH, FW = 16, 2160, 3840
frames = torch.from_numpy(np.random.Generator(np.random.PCG64(4242)).integers(0, 256, (NF, 3, FH, FW), dtype=np.uint8)).cuda()
org = tile_origins(FH, FW, 1280)
T = len(org)
def table(ids):
    return torch.tensor([[t // T, org[t % T][0], org[t % T][1]] for t in ids], dtype=torch.int32).cuda()
for use_graph in (True, False):
    m.use_cuda_graph = use_graph
    m._plans.clear()
    a_ids = list(range(16))
    b_ids = [0] + list(range(8, 128, 8))
    da, ra = m.forward_tiles(frames, table(a_ids)); da = da.clone(); ra = [r.clone() for r in ra]
    db, rb = m.forward_tiles(frames, table(b_ids)); db = db.clone(); rb = [r.clone() for r in rb]
    da2, _ = m.forward_tiles(frames, table(a_ids)); da2 = da2.clone()
    torch.cuda.synchronize()
    print("graph", use_graph, "tile 0 (slot 0 in both): det equal", torch.equal(da[0], db[0]), "max diff", float((da[0] - db[0]).abs().max()),
          "| tile 8 (slot 8 vs slot 1):", torch.equal(da[8], db[1]), float((da[8] - db[1]).abs().max()),
          "| repeat A equal", torch.equal(da, da2))
    for lv in range(3):
        print("   level", lv, "raw tile0 max diff", float((ra[lv][0] - rb[lv][0]).abs().max()), "tile8", float((ra[lv][8] - rb[lv][1]).abs().max()))
