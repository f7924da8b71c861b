#!/usr/bin/env python
"""bench.py -- images/sec of SkyEye's batched detector forward path on B200.

Workload (BASELINE.json configs[2], the config the metric is quoted on): skyeye_l (CLA neck +
transformer heads) at 1280x1280, batch 16 per GPU, bf16 operands / fp32 accumulate.  A step is one
pass of the hot path over one batch of synthetic images: forward -> anchor decode -> NMS
(conf .25, iou .45, max_det 300).  Images shard by rank with no data-path collective (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's B200 path
  python bench.py --impl reference [...]                     # the CPU oracle port of the reference path

Prints ONE JSON line (see the driver contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "images/sec at 1280^2 (skyeye_l, bf16)"
VARIANT, H, W, BATCH = "skyeye_l", 1280, 1280, 16
CONF, IOU, MAX_DET = 0.25, 0.45, 300


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default=VARIANT)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--size", type=int, default=H)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the batch-1 latency probe (keeps an ncu launch list to the timed steps)")
    ap.add_argument("--no-graph", action="store_true", help="launch the forward plan kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-tiled", action="store_true", help="skip the config-4 leg (16 4K frames = 128 tiles, tile-sharded, gather + merge)")
    ap.add_argument("--tiled-steps", type=int, default=0, help="timed steps of the config-4 leg (default: 4 at N=1, --steps otherwise)")
    ap.add_argument("--no-variants", action="store_true", help="skip the informational skyeye_lw (windowed heads) leg")
    ap.add_argument("--profile-json", default=None, help="write the per-kernel table here")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    except Exception:
        return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, source="fallback")


def synthetic_images(batch, size, seed):
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, (batch, 3, size, size), dtype=np.uint8)


# ------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
WEIGHTS = ("random 'trained-like' weights, seed 0 (oracle.model.make_calibrated_state_dict: reference init distributions, "
           "BN statistics calibrated on seeded noise images, residual-branch BN scale 0.25); the SAME state dict in both arms")


def shared_state_dict(variant):
    """The one state dict both arms load (weight recipe = test infrastructure under oracle/; building the weights is
    not part of any timed region)."""
    from oracle import model as om
    cfg = om.get_cfg(variant)
    return om.make_calibrated_state_dict(cfg, 0), cfg


def workload_config(variant, size, batch):
    """Identical in both arms (the driver compares the two lines' configs)."""
    return {"workload": f"{variant} {size}x{size} batch {batch}/GPU: forward (CSP backbone, PAN neck, CLA, transformer heads) + decode + "
                        f"NMS(conf {CONF}, iou {IOU}, max_det {MAX_DET})",
            "weights": WEIGHTS, "input": "synthetic uint8 images, PCG64 seed 1234 + rank",
            "sharding": f"images by rank, {batch} per GPU, no data-path collective"}


def cpu_reference_kind():
    """"reference" when the reference's own files are reachable (the build container's /root/reference, or the copy build()
    staged under the git-ignored oracle/_ref/ for the GPU box), else "port" (oracle/model.py + oracle/nms.py)."""
    if os.environ.get("SKYEYE_CPU_ARM") == "port":
        return "port"
    try:
        from oracle import ref_loader
        return "reference" if ref_loader.available() else "port"
    except Exception:
        return "port"


def cpu_oracle_run(variant, size, n_images, steps, warmup, seed=1234, sd=None, cfg=None, kind=None):
    """Times the reference path on this box's host cores: kind "reference" = the reference's OWN modules (detector.py,
    backbone.py, blocks.py, attention.py incl. nn.MultiheadAttention materialising the N x N weights, metrics.py's
    non_max_suppression over torchvision.ops.nms) with the enumerated repairs R1-R4 of SURVEY.md §0.2; kind "port" = the
    restatement under oracle/ (same arithmetic, attention evaluated in chunks).
    Returns (images/s, cores, seconds/step, (det, raws, nms_rows) of the last pass, kind)."""
    import torch
    from oracle import model as om
    from oracle import nms as onms
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if sd is None:
        sd, cfg = shared_state_dict(variant)
    kind = kind or cpu_reference_kind()
    x = torch.from_numpy(synthetic_images(n_images, size, seed)).float() / 255.0
    ref_model = ref_nms = None
    if kind == "reference":
        from oracle import ref_loader
        ref_model = ref_loader.build_reference_model(cfg)
        missing = ref_model.load_state_dict(sd, strict=True)
        ref_nms = ref_loader.load().metrics.non_max_suppression
    else:
        onms.build()
    ts = []
    last = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if ref_model is not None:
            with torch.no_grad():
                det, raws = ref_model(x)
                rows = [r.numpy() for r in ref_nms(det, CONF, IOU, max_detections=MAX_DET)]
        else:
            det, raws = om.forward(x, sd, cfg)
            rows = onms.non_max_suppression(det.numpy(), CONF, IOU, max_detections=MAX_DET)
        dt = time.perf_counter() - t0
        last = (det, raws, rows)
        if i >= warmup:
            ts.append(dt)
    total = sum(ts)
    return n_images * len(ts) / total, cores, total / len(ts), last, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img = 1  # bounded sample: one image of the 16-image batch per step
    value, cores, sec, _, kind = cpu_oracle_run(args.variant, args.size, n_img, args.steps, args.warmup)
    sample = f"{n_img} image of the {args.batch}-image batch per step, {args.variant} {args.size}x{args.size} fp32, forward+decode+NMS"
    arm = ("the reference's own PyTorch modules (oracle/_ref: detector/backbone/blocks/attention.py + metrics.non_max_suppression over "
           "torchvision.ops.nms, repairs R1-R4), fp32, eager, all host cores" if kind == "reference"
           else "CPU oracle port of the reference path (fp32, PyTorch eager + C NMS)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.variant, args.size, args.batch),
        "run": {"arm": arm, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [q.strip() for q in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        # samples under load = power draw at least half of the maximum seen (the sampler starts during warm-up)
        pmax = max(power) if power else 0.0
        loaded = sorted(c for c, w in zip(sm, power) if w >= 0.5 * pmax) or sorted(sm)
        med = loaded[len(loaded) // 2] if loaded else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(loaded),
                "power_w_max": pmax if power else None}


def run_b200(args):
    import torch
    import torch.distributed as dist

    from skyeye import _native
    from skyeye.core.detector import construct_model
    from skyeye.utils.nms import batched_nms_padded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (b200 arm) needs a CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _native.check(_native.lib().skb_device_check(), "skb_device_check")

    sd, ocfg = shared_state_dict(args.variant)       # the same weights the CPU arm runs (built outside every timed region)
    model = construct_model(f"{args.variant}.yaml")
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    model.reuse_output_buffers = True
    model.use_cuda_graph = not args.no_graph  # the forward plan (no host syncs) is captured once per input shape and replayed
    B, S = args.batch, args.size

    host = torch.from_numpy(synthetic_images(B, S, 1234 + rank)).pin_memory()
    x_dev = host.to(dev)
    nms_out = torch.zeros((B, MAX_DET, 7), dtype=torch.float32, device=dev)
    nms_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    out_host = torch.empty((B, MAX_DET, 7), dtype=torch.float32).pin_memory()
    cnt_host = torch.empty(B, dtype=torch.int32).pin_memory()

    def step_resident():
        det, _ = model(x_dev)
        batched_nms_padded(det, CONF, IOU, max_detections=MAX_DET, out=nms_out, out_count=nms_cnt)

    # End-to-end step: every step copies ITS input batch from pinned host memory and reads ITS result back.
    # Like any input pipeline, the H2D copy of step i+1 is enqueued on a copy stream before step i's result is
    # awaited, so it overlaps step i's kernels (two device-side input buffers).
    copy_stream = torch.cuda.Stream()
    xbuf = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def enqueue_h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])    # the forward pass that last read this buffer has finished
            xbuf[slot].copy_(host, non_blocking=True)  # H2D of one step's uint8 images from pinned memory
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = state["i"] & 1
        if not state["primed"]:
            for sl in (0, 1):
                consumed[sl].record()
            enqueue_h2d(cur)
            state["primed"] = True
        main = torch.cuda.current_stream()
        main.wait_event(ready[cur])
        det, _ = model(xbuf[cur])                     # public API call (validate.py:245)
        consumed[cur].record(main)
        enqueue_h2d(cur ^ 1)                          # next step's input, overlapping this step's kernels
        batched_nms_padded(det, CONF, IOU, max_detections=MAX_DET, out=nms_out, out_count=nms_cnt)  # (validate.py:255)
        out_host.copy_(nms_out, non_blocking=True)    # D2H of the step's result
        cnt_host.copy_(nms_cnt, non_blocking=True)
        main.synchronize()
        state["i"] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    warm = max(args.warmup, 3)
    step_resident()  # builds the plan (weight packing, buffers, graph capture): not part of any sampled region
    torch.cuda.synchronize()
    # nvidia-smi needs a few hundred ms to deliver its first sample and the timed region can be shorter than that,
    # so the sampler also covers the warm-up steps (same kernels, same load); idle samples are filtered in stop().
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:  # keep the GPU under the bench load until the sampler is running
            step_resident()
        torch.cuda.synchronize()
    for _ in range(warm):
        step_resident()
    torch.cuda.synchronize()
    plan = model.plan_for(x_dev)

    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # per-kernel shares: CUDA events around every launch of K more steps (same stream, same inputs)
    agg = {}
    per_step = []
    for _ in range(min(args.steps, 5)):
        model._img[0] = x_dev
        rows = plan.run_timed()
        per_step.append(rows)
    for rows in per_step:
        for (name, ms), meta in zip(rows, plan.meta):
            a = agg.setdefault(meta["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0, calls=0))
            a["ms"] += ms
            a["flops"] += meta["flops"]
            a["bytes"] += meta["bytes"]
            a["launches"] += meta["launches"]
            a["calls"] += 1
    nrep = len(per_step)
    # NMS timing (not part of the plan)
    det = plan.det
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nrep):
        batched_nms_padded(det, CONF, IOU, max_detections=MAX_DET, out=nms_out, out_count=nms_cnt)
    e1.record()
    torch.cuda.synchronize()
    NMS_LAUNCHES = 2  # compacting filter + the per-image order-and-keep kernel (radix select + bitonic sort + greedy pass in one CTA)
    agg["nms"] = dict(ms=e0.elapsed_time(e1), flops=0.0, bytes=4.0 * det.numel() * nrep, launches=NMS_LAUNCHES * nrep, calls=nrep)
    pk = peaks()
    table = {}
    tot_ms = sum(a["ms"] for a in agg.values())
    for k, a in agg.items():
        ms1 = a["ms"] / nrep
        table[k] = dict(ms_per_step=ms1, share=a["ms"] / tot_ms, launches_per_step=a["launches"] // nrep,
                        tflops=(a["flops"] / nrep) / (ms1 * 1e-3) / 1e12 if ms1 > 0 else 0.0,
                        gbs=(a["bytes"] / nrep) / (ms1 * 1e-3) / 1e9 if ms1 > 0 else 0.0)
    dom = max(("conv", "attention"), key=lambda k: table.get(k, {"ms_per_step": 0})["ms_per_step"])
    d = agg[dom]
    achieved = (d["flops"] / d["calls"]) / ((d["ms"] / d["calls"]) * 1e-3) / 1e12
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if dom == "attention" and args.variant == VARIANT and S == H and B == BATCH:
            traffic = tj["flash_attn_kernel"]["dram_bytes_per_launch_avg"]
    except Exception:
        traffic = None
    roofline = {"bound": "tensor", "kernel": "flash_attn_kernel" if dom == "attention" else "conv_gemm_kernel",
                "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"],
                "traffic": traffic, "peak_source": f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": d["launches"] // nrep, "avg_launch_ms": d["ms"] / d["calls"],
                "share_of_step": table[dom]["share"],
                "algorithmic_flops_per_launch": d["flops"] / d["calls"], "algorithmic_bytes_per_launch": d["bytes"] / d["calls"],
                "note": "attention at head_dim 64: the MUFU pipe (16 ex2/clk/SM measured, 256 flop per exponential) caps it at "
                        "~1.12 PFLOP/s at 1.85 GHz (16 of 64 exponentials run as polynomials on the FMA pipes); the measured bound is each "
                        "softmax warp's dependency chain through the MUFU phase (XU 70 %, tensor pipe 52 % busy), and the board sits on its "
                        "1000 W cap (SM clock ~1.63 GHz), so cycles saved come back as clock, not as time "
                        "(profiles/r2b_uniform_issue.md, profiles/r2k_ncu_attention.md)" if dom == "attention" else ""}

    # end-to-end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # p50 batch-1 latency (second half of BASELINE.json's metric): one resident image through model(x) + NMS
    latency = None
    if not args.no_latency:   # at N > 1 every rank probes its own GPU (replicas); the line reports the slowest rank's percentiles
        x1 = x_dev[:1].contiguous()
        o1 = torch.zeros((1, MAX_DET, 7), dtype=torch.float32, device=dev)
        c1 = torch.zeros(1, dtype=torch.int32, device=dev)

        def step_b1():
            d1, _ = model(x1)
            batched_nms_padded(d1, CONF, IOU, max_detections=MAX_DET, out=o1, out_count=c1)

        for _ in range(3):
            step_b1()
        torch.cuda.synchronize()
        lat = []
        for _ in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_b1()
            b.record()
            torch.cuda.synchronize()
            lat.append(a.elapsed_time(b))
        lat.sort()
        pcts = [lat[len(lat) // 2], lat[int(len(lat) * 0.9)], lat[0]]
        if world > 1:
            t = torch.tensor(pcts, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pcts = [float(v) for v in t.tolist()]
        latency = {"p50_ms": pcts[0], "p90_ms": pcts[1], "min_ms": pcts[2], "samples": len(lat),
                   "scope": f"{args.variant} {S}x{S} batch 1: model(x) + NMS, image resident, CUDA events"
                            + (f"; max over the {world} ranks (one replica per GPU)" if world > 1 else "")}

    # BASELINE config 4 (north_star: "scales >= 7x from 1 to 8 GPUs on tiled 4K frames"): 16 synthetic 3840x2160 frames = 128 tiles of
    # 1280^2, tiles round-robin over ranks, forward + per-tile NMS with no collective, ONE all_gather of padded rows + counts,
    # per-frame merge NMS on every rank (skyeye/utils/tiling.TiledDetector).  Strong scaling: the 16 frames are fixed.
    tiled = None
    if not args.no_tiled and args.variant == VARIANT and S == H:
        import hashlib
        from skyeye.utils.tiling import TiledDetector
        NF, FH, FW = 16, 2160, 3840
        import numpy as np
        frames = torch.from_numpy(np.random.Generator(np.random.PCG64(4242)).integers(0, 256, (NF, 3, FH, FW), dtype=np.uint8)).to(dev)  # same on every rank
        td = TiledDetector(model, NF, (FH, FW), rank=rank, world=world, conf=CONF, iou=IOU, max_det=MAX_DET, compat="fixed")
        for _ in range(2):
            td(frames)
        torch.cuda.synchronize()
        tsteps = args.tiled_steps or (4 if world == 1 else max(args.steps, 4))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slot = 0
        for _ in range(tsteps):
            slot = td.step(frames)          # merge of step t (side stream) overlaps the forward of step t+1
        torch.cuda.current_stream().wait_event(td.ev_done[slot])
        if tsteps > 1:
            torch.cuda.current_stream().wait_event(td.ev_done[slot ^ 1])
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_t], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_t = float(t.item())
        rows_t, cnt_t = td.result(slot)
        sha = hashlib.sha256(rows_t.cpu().numpy().tobytes() + cnt_t.cpu().numpy().tobytes()).hexdigest()[:16]
        merge_ms = sum(td.merge_ms(sl) for sl in (0, 1)) / 2
        tiles_local = td.n_local
        tiled = {"frames_per_s": NF * tsteps / (ms_t / 1e3), "tiles_per_s": td.n_tiles * tsteps / (ms_t / 1e3), "ms_per_step": ms_t / tsteps,
                 "steps": tsteps, "frames_per_step": NF, "tiles_per_step": td.n_tiles, "tiles_per_rank": tiles_local,
                 "ms_gather_merge": merge_ms, "sha256_16": sha, "detections": int(cnt_t.sum()), "scaling": "strong",
                 "untiled_ms_for_the_same_tiles": ms_step * tiles_local / B,
                 "config": f"{NF} frames {FW}x{FH} -> {td.T} tiles of 1280^2 per frame (SURVEY D8), tile_id % world, per-tile NMS(conf {CONF}, iou {IOU}, "
                           f"max_det {MAX_DET}, rows [x1,y1,x2,y2,conf,cls]), one all_gather [{td.n_local_max},{MAX_DET + 1},7] fp32 per rank, merge NMS on "
                           "every rank; gather + merge of step t on a second stream under the forward of step t+1; frames resident"}
        del td, frames

    # Informational (beside -- not instead of -- the headline): the same network with `head: windowed`
    # (WindowedSelfAttention over 8x8 windows, O(N w^2) instead of O(N^2)); same backbone / neck weights, its own head weights.
    windowed = None
    if world == 1 and not args.no_variants and args.variant == VARIANT and S == H:
        sdw, _ = shared_state_dict("skyeye_lw")
        mw = construct_model("skyeye_lw.yaml")
        mw.load_state_dict(sdw, strict=True)
        mw = mw.to(dev).eval()

        def step_w():
            dw, _ = mw(x_dev)
            batched_nms_padded(dw, CONF, IOU, max_detections=MAX_DET, out=nms_out, out_count=nms_cnt)

        for _ in range(3):
            step_w()
        wsteps = min(args.steps, 5)
        ms_w = timed(step_w, wsteps) / wsteps
        pw = mw.plan_for(x_dev)
        mw._img[0] = x_dev
        rows_w = pw.run_timed()
        att_w = sum(ms for (nm, ms), meta in zip(rows_w, pw.meta) if meta["kind"] == "attention")
        att_g = table["attention"]["ms_per_step"]
        windowed = {"variant": "skyeye_lw (head: windowed, window_size 8)", "images_per_s": B / (ms_w / 1e3), "ms_per_step": ms_w,
                    "attention_ms_per_step": att_w, "global_attention_ms_per_step": att_g,
                    "note": "window attention core = skb_window_attn2d_bf16 (persistent tcgen05 kernel: pairs of 64-token windows per "
                            "M128 tile, HBM-bound); detections differ from skyeye_l by construction (a different head)"}
        del mw, pw

    # Informational: SURVEY.md §8(d) config 2 -- skyeye_s (plain CSP / PAN detector, conv heads) at 640^2, batch 32, same protocol
    # (inputs resident, CUDA-graph replay, forward + decode + NMS); plain seeded init (the variant is well conditioned as it is).
    config2 = None
    if world == 1 and not args.no_variants and args.variant == VARIANT and S == H:
        from oracle import model as om2
        cfg_s = om2.get_cfg("skyeye_s")
        ms_model = construct_model("skyeye_s.yaml")
        ms_model.load_state_dict(om2.make_state_dict(cfg_s, 0), strict=True)
        ms_model = ms_model.to(dev).eval()
        ms_model.reuse_output_buffers = True
        xs = torch.from_numpy(synthetic_images(32, 640, 99)).to(dev)
        o2 = torch.zeros((32, MAX_DET, 7), dtype=torch.float32, device=dev)
        c2 = torch.zeros(32, dtype=torch.int32, device=dev)

        def step_s():
            ds, _ = ms_model(xs)
            batched_nms_padded(ds, CONF, IOU, max_detections=MAX_DET, out=o2, out_count=c2)

        for _ in range(5):
            step_s()
        ssteps = 30
        ms_s = timed(step_s, ssteps) / ssteps
        ps = ms_model.plan_for(xs)
        ms_model._img[0] = xs
        rows_s = ps.run_timed()
        fl_s = sum(m["flops"] for m in ps.meta)
        by_s = sum(m["bytes"] for m in ps.meta)
        config2 = {"variant": "skyeye_s 640x640 batch 32 (SURVEY §8d config 2)", "images_per_s": 32 / (ms_s / 1e3), "ms_per_step": ms_s,
                   "kernel_ms_per_step": sum(ms for nm, ms in rows_s), "launches": ps.launches + NMS_LAUNCHES,
                   "algorithmic_tflops": fl_s / 1e12, "algorithmic_gbytes": by_s / 1e9,
                   "tflops": fl_s / (ms_s * 1e9), "gbs": by_s / (ms_s * 1e6),
                   "note": "HBM-bound network (most layers below the 206 flop/B ridge): gbs counts each launch's unfused bf16 input + weights + output"}
        del ms_model, ps, xs

    launches_per_step = plan.launches + NMS_LAUNCHES
    from skyeye.engine import View
    act_gb = sum(v.t.numel() * v.t.element_size() for v in plan.keep if isinstance(v, View)) / 1e9

    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # GPU result for image 0 of the batch (state left by the last resident step), then the CPU leg on the same image + weights
        step_resident()
        torch.cuda.synchronize()
        g_raws = [r[0].float().cpu() for r in plan.raw_out]
        g_det0 = plan.det[:1].float().cpu()
        g_rows = nms_out[0, :int(nms_cnt[0])].cpu().numpy()
        v, cores, sec, (o_det, o_raws, o_rows), ckind = cpu_oracle_run(args.variant, S, 1, 1, 1, seed=1234 + rank, sd=sd, cfg=ocfg)
        o_raws = [r.detach() for r in o_raws]
        o_det = o_det.detach()
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": ckind,
               "sample": f"1 image of the {B}-image batch (1 warm-up + 1 timed pass, {sec:.1f} s), {args.variant} {S}x{S} fp32 "
                         f"{'reference modules' if ckind == 'reference' else 'oracle port'} forward+decode+NMS"}
        import numpy as np
        from oracle import nms as onms
        lv = []
        for a, b in zip(g_raws, o_raws):
            b = b[0]
            lv.append({"max_rel": float((a - b).abs().max() / b.abs().max()), "rms": float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())})
        ref_rows = onms.non_max_suppression(g_det0.numpy(), CONF, IOU, max_detections=MAX_DET)[0]   # oracle NMS on the GPU's own detections
        dec = ((g_det0[..., 4:] > 0.5) == (o_det[..., 4:] > 0.5)).float().mean()
        parity = {"vs": f"fp32 CPU {'reference modules' if ckind == 'reference' else 'oracle port'}, image 0 of the batch, same state dict", "max_rel": max(l["max_rel"] for l in lv),
                  "rms": max(l["rms"] for l in lv), "per_level": lv,
                  "nms_rows_equal": bool(g_rows.shape == ref_rows.shape and np.array_equal(g_rows, ref_rows)),
                  "nms_rows": int(g_rows.shape[0]), "sigmoid_decisions_agree": float(dec),
                  "note": "max_rel / rms of the raw head logits relative to per-level max |logit| / rms logit (bf16 activations vs fp32 reference "
                          "arithmetic; tests/test_gpu_teacher_forced.py holds the per-launch 1e-3 gate); nms_rows_equal = GPU NMS rows bit-equal to "
                          "the oracle NMS on identical detections"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": workload_config(args.variant, S, B),
            "run": {"arm": "B200 native path (BN folded, bf16 operands, fp32 accumulate)",
                    "launch": "CUDA graph replay of the forward plan + NMS launches" if model.use_cuda_graph else "per-kernel launches",
                    "l2": f"inputs+activations per step {act_gb:.1f} GB >> 126 MB L2 (no flush needed)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(host.numel()),
                    "d2h_bytes_per_step": int(out_host.numel() * 4 + cnt_host.numel() * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "tiled4k": tiled,
            "windowed_head": windowed,
            "config2_skyeye_s": config2,
            "latency_b1": latency,
            "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in table.items()},
        }
        print(json.dumps(line))
        if args.profile_json:
            with open(args.profile_json, "w") as f:
                json.dump({"line": line, "per_launch": [(n, round(ms, 4), m) for (n, ms), m in zip(per_step[-1], plan.meta)]}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
