"""Stage-by-stage numerics probe (diagnostic, GPU): where does the CUDA path diverge from the
bf16-emulating oracle?  Prints max-rel and rms-rel error of each stage output."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200"), os.path.join(ROOT, "tests", "golden")]
import cases  # noqa: E402
from oracle import model as om  # noqa: E402
from skyeye import engine as E  # noqa: E402
from skyeye.core.detector import construct_model  # noqa: E402


def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b)
    return float(d.abs().max() / b.abs().max()), float(d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt())


def main(variant="skyeye_s", shape=(2, 3, 128, 160)):
    cfg = om.get_cfg(variant)
    sd = om.make_state_dict(cfg, 0)
    m = construct_model(f"{variant}.yaml")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = cases.image(shape)
    ctx = om.Ctx("bf16")
    f32 = om.Ctx(None)
    bb = m.backbone.backbone
    plan = E.Plan("cuda")
    img = [x.cuda().contiguous()]
    n, _, h, w = shape
    taps = {}
    v = bb.stage1[0].lower_image(plan, img, n, h, w); taps["focus.conv"] = v
    v = bb.stage1[1].lower(plan, v); taps["s1.down"] = v
    v = bb.stage1[2].lower(plan, v); taps["s1.csp"] = v
    v = bb.stage2[0].lower(plan, v); taps["s2.down"] = v
    v = bb.stage2[1].lower(plan, v); taps["s2.csp"] = v
    v = bb.stage3[0].lower(plan, v); taps["s3.down"] = v
    v = bb.stage3[1].lower(plan, v); taps["s3.csp"] = v
    v = bb.stage3[2].lower(plan, v); taps["s3.cbam"] = v
    v = bb.stage4[0].lower(plan, v); taps["s4.down"] = v
    v = bb.stage4[1].lower(plan, v); taps["s4.csp"] = v
    v = bb.stage4[2].lower(plan, v); taps["s4.spp"] = v
    plan.run()
    torch.cuda.synchronize()
    d3, d9 = om.depths(cfg)
    p = "backbone.backbone."
    for name, c in (("emu", ctx), ("fp32", f32)):
        o = {}
        t = om.focus(c.q(x), sd, p + "stage1.0", c); o["focus.conv"] = t
        t = om.conv_block(t, sd, p + "stage1.1", 2, c); o["s1.down"] = t
        t = om.csp(t, sd, p + "stage1.2", d3, c); o["s1.csp"] = t
        t = om.conv_block(t, sd, p + "stage2.0", 2, c); o["s2.down"] = t
        t = om.csp(t, sd, p + "stage2.1", d9, c); o["s2.csp"] = t
        t = om.conv_block(t, sd, p + "stage3.0", 2, c); o["s3.down"] = t
        t = om.csp(t, sd, p + "stage3.1", d9, c); o["s3.csp"] = t
        t = om.cbam(t, sd, p + "stage3.2", c); o["s3.cbam"] = t
        t = om.conv_block(t, sd, p + "stage4.0", 2, c); o["s4.down"] = t
        t = om.csp(t, sd, p + "stage4.1", d3, c); o["s4.csp"] = t
        t = om.spp(t, sd, p + "stage4.2", c); o["s4.spp"] = t
        for k in taps:
            mx, rms = err(taps[k].nchw(), o[k])
            print(f"{variant} gpu vs {name:5s} {k:12s} max {mx:.3e} rms {rms:.3e}")
    # single-layer exactness: feed the ORACLE's emu input of a layer to the GPU layer
    t_in = om.focus(ctx.q(x), sd, p + "stage1.0", ctx)
    ref = om.conv_block(t_in, sd, p + "stage1.1", 2, ctx)
    plan2 = E.Plan("cuda")
    out = bb.stage1[1].lower(plan2, E.from_nchw(t_in.cuda()))
    plan2.run(); torch.cuda.synchronize()
    got = out.nchw().float().cpu()
    print("single layer s1.down on oracle input: max/rms", err(got, ref), "mismatching bf16 values:",
          int((got != ref).sum()), "of", ref.numel())
    y32 = E.new_buffer(out.n, out.h, out.w, out.c, torch.float32)
    E.conv2d(E.from_nchw(t_in.cuda()), bb.stage1[1].packed("cuda"), y32, 2, 1)
    torch.cuda.synchronize()
    wf, bf = om.fold_bn(sd, p + "stage1.1")
    r32 = torch.nn.functional.silu(torch.nn.functional.conv2d(t_in, om.bf16_round(wf), bf, 2, 1))
    print("single layer fp32 store: max/rms", err(y32.nchw(), r32))


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["skyeye_s"]))
