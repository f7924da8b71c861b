"""``python -m skyeye.cli.validate`` -- the evaluation entry point around the B200 forward path.

Keeps the interface of the reference's ``skyeye/cli/validate.py`` (``validate(...)`` signature and return
tuple ``(mp, mr, map50, map, *val_loss)``, docstring :169-171; CLI flags of README.md:69) -- the reference
file itself is truncated mid-statement (:337) and imports names that exist nowhere, so
the flow below is restated from what it spells out (:176-326): uint8 batch -> device -> model ->
``non_max_suppression(out, conf_thres, iou_thres, multi_label=True, agnostic=single_cls)`` -> per-image
IoU matching at 10 thresholds -> ``ap_per_class`` -> P / R / mAP@.5 / mAP@.5:.95 and the 3-bucket
pre-process / inference / NMS timing line.

Differences that are forced, and stated:
  * the model consumes the uint8 batch directly (the ``/255`` of :237-238 is fused into the first kernel);
    ``half`` is accepted and ignored (the path computes in bf16 with fp32 accumulation);
  * NMS rows: the reference wrapper returns ``[cx,cy,w,h,obj,cls_prob,cls_id]`` (quirk X8) while this
    function's bookkeeping (:268-288) needs ``[x1,y1,x2,y2,conf,cls]``; ``nms_compat="fixed"`` (default)
    selects the rows the reference docstring promises, ``"reference"`` reproduces the quirk verbatim;
  * the reference's dataloader package cannot be imported (X14): pass any iterable of
    ``(uint8 img [B,3,H,W], targets [n,6] = (image, class, cx, cy, w, h normalised), paths, shapes)``
    batches (the collate format of dataset.py:349-365), or let ``FolderLoader`` read ``images/`` +
    ``labels/*.txt`` folders named by the data YAML.
"""
from __future__ import annotations

import argparse
import json
import logging
import time
from pathlib import Path

import numpy as np
import torch
import yaml

from ..core.models.detector import load_model
from ..utils.general import check_img_size, letterbox, scale_boxes, xywh2xyxy, xyxy2xywh
from ..utils.metrics import ap_per_class, box_iou, non_max_suppression

LOGGER = logging.getLogger("skyeye")
IMG_EXT = {".jpg", ".jpeg", ".png", ".bmp"}


def time_sync() -> float:
    """cuda.synchronize() + time.time() (torch_utils.py:109-118)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.time()


def save_one_txt(pred, save_conf, shape, file):
    """One ``cls x y w h [conf]`` line per detection, xywh normalised by the image size, ``%g`` (validate.py:31-46)."""
    gn = torch.tensor(shape)[[1, 0, 1, 0]].float()
    Path(file).parent.mkdir(parents=True, exist_ok=True)
    with open(file, "a") as f:
        for *xyxy, conf, cls in pred[:, :6].tolist():
            xywh = (xyxy2xywh(torch.tensor(xyxy).view(1, 4)) / gn).view(-1).tolist()
            line = (cls, *xywh, conf) if save_conf else (cls, *xywh)
            f.write(("%g " * len(line)).rstrip() % line + "\n")


def save_one_json(pred, jdict, path, class_map):
    """COCO-style records: bbox = top-left xywh rounded to 3 dp, score to 5 dp (validate.py:49-68)."""
    path = Path(path)
    image_id = int(path.stem) if path.stem.isnumeric() else path.stem
    box = xyxy2xywh(pred[:, :4])
    box[:, :2] -= box[:, 2:] / 2
    for p, b in zip(pred.tolist(), box.tolist()):
        cid = int(p[5])
        if class_map is not None:  # validate.py:63: class_map[int(p[5])] (dict or list; identity when the id is not mapped)
            try:
                cid = class_map[cid]
            except (KeyError, IndexError):
                pass
        jdict.append({"image_id": image_id, "category_id": cid, "bbox": [round(x, 3) for x in b], "score": round(p[4], 5)})


def process_batch(detections, labels, iouv):
    """Correct-prediction matrix [n_det, n_iou] (validate.py:71-108): a detection is correct at threshold t if
    it is the best-IoU unmatched detection of a same-class label with IoU >= t; one label per detection and
    one detection per label, highest IoU first."""
    if detections.is_cuda:  # the validation loop: one kernel, no host round trip (skb_match_detections_f32, csrc/eval.cu)
        from .. import _native as N
        n, m, T = detections.shape[0], labels.shape[0], iouv.shape[0]
        out = torch.zeros((n, T), dtype=torch.uint8, device=detections.device)
        if n == 0 or m == 0:
            return out.bool()
        det = detections[:, :6].float().contiguous()
        lab = labels[:, :5].float().contiguous().to(det.device)
        thr = iouv.float().contiguous().to(det.device)
        ws = torch.empty(int(N.lib().skb_match_workspace_bytes(n, m)) + 256, dtype=torch.uint8, device=det.device)
        with torch.cuda.device(det.device):
            N.check(N.lib().skb_match_detections_f32(lab.data_ptr(), m, det.data_ptr(), n, thr.data_ptr(), T, out.data_ptr(), ws.data_ptr(),
                                                     ws.numel(), torch.cuda.current_stream().cuda_stream), "skb_match_detections_f32")
        return out.bool()
    # host tensors (unit tests of the bookkeeping against the live reference): the reference's own sequence of torch / numpy steps
    correct = torch.zeros(detections.shape[0], iouv.shape[0], dtype=torch.bool, device=iouv.device)
    if detections.shape[0] == 0 or labels.shape[0] == 0:
        return correct
    iou = box_iou(labels[:, 1:], detections[:, :4])
    same = labels[:, 0:1] == detections[:, 5]
    for i in range(iouv.shape[0]):
        li, di = torch.where((iou >= iouv[i]) & same)
        if li.numel() == 0:
            continue
        m = torch.stack((li.float(), di.float(), iou[li, di]), 1).cpu().numpy()
        if m.shape[0] > 1:
            m = m[m[:, 2].argsort()[::-1]]
            m = m[np.unique(m[:, 1], return_index=True)[1]]
            m = m[m[:, 2].argsort()[::-1]]
            m = m[np.unique(m[:, 0], return_index=True)[1]]
        correct[torch.from_numpy(m[:, 1].astype(np.int64)).to(iouv.device), i] = True
    return correct


class FolderLoader:
    """Minimal stand-in for ``create_dataloader(..., rect=True, pad=0.5)`` (validate.py:213-222): walks an image
    folder, letterboxes every image to ``img_size`` (gray 114, stride multiple), reads YOLO ``labels/<stem>.txt``
    (``cls cx cy w h`` normalised) and yields collated batches in the reference's format."""

    def __init__(self, image_dir, img_size=640, batch_size=32, stride=32):
        self.files = sorted(p for p in Path(image_dir).rglob("*") if p.suffix.lower() in IMG_EXT)
        if not self.files:
            raise FileNotFoundError(f"no images under {image_dir}")
        self.img_size, self.batch_size, self.stride = img_size, batch_size, stride

    def __len__(self):
        return (len(self.files) + self.batch_size - 1) // self.batch_size

    @staticmethod
    def label_path(img_path: Path) -> Path:
        parts = list(img_path.parts)
        if "images" in parts:
            parts[len(parts) - 1 - parts[::-1].index("images")] = "labels"
        return Path(*parts).with_suffix(".txt")

    def __iter__(self):
        import cv2
        for i in range(0, len(self.files), self.batch_size):
            ims, tg, paths, shapes = [], [], [], []
            for k, f in enumerate(self.files[i:i + self.batch_size]):
                im0 = cv2.imread(str(f))
                if im0 is None:
                    raise FileNotFoundError(f)
                h0, w0 = im0.shape[:2]
                im, r, (pw, ph) = letterbox(im0, self.img_size, stride=self.stride)
                ims.append(im)
                lp = self.label_path(f)
                if lp.exists():
                    lab = np.loadtxt(lp, ndmin=2, dtype=np.float32).reshape(-1, 5)
                    if lab.size:
                        # normalised original-image xywh -> PIXELS of the letterboxed image (which sits at the top-left of the
                        # batch frame); normalised by the batch frame below, once its size is known
                        cx, cy = lab[:, 1] * w0 * r + pw, lab[:, 2] * h0 * r + ph
                        bw, bh = lab[:, 3] * w0 * r, lab[:, 4] * h0 * r
                        tg.append(np.stack([np.full(len(lab), k, np.float32), lab[:, 0], cx, cy, bw, bh], 1))
                paths.append(str(f))
                shapes.append(((h0, w0), ((r, r), (pw, ph))))
            H, W = max(a.shape[0] for a in ims), max(a.shape[1] for a in ims)
            batch = np.full((len(ims), H, W, 3), 114, np.uint8)
            for k, a in enumerate(ims):
                batch[k, :a.shape[0], :a.shape[1]] = a
            img = torch.from_numpy(np.ascontiguousarray(batch[..., ::-1].transpose(0, 3, 1, 2)))  # BGR -> RGB, NCHW
            targets = torch.from_numpy(np.concatenate(tg, 0)) if tg else torch.zeros((0, 6))
            # images of different aspect ratios letterbox to different sizes; validate() rescales targets by the BATCH frame
            # (validate.py:252), so that is what they are normalised by
            targets[:, 2:] /= torch.tensor([W, H, W, H], dtype=targets.dtype)
            yield img, targets, paths, shapes


@torch.no_grad()
def validate(data, weights=None, batch_size=32, img_size=640, conf_thres=0.001, iou_thres=0.6, task="val", device="",
             workers=8, single_cls=False, augment=False, verbose=False, save_txt=False, save_hybrid=False, save_conf=False,
             save_json=False, project="runs/val", name="exp", exist_ok=False, half=True, model=None, dataloader=None,
             save_dir=Path(""), plots=False, compute_loss=None, nms_compat="fixed", max_det=300):
    """Returns ``(mp, mr, map50, map, *val_loss)``; see the module docstring for the flow."""
    training = model is not None
    if isinstance(data, (str, Path)):
        with open(data, errors="ignore") as f:
            data_dict = yaml.safe_load(f)
    else:
        data_dict = dict(data)
    if training:
        dev = next(model.parameters()).device
    else:
        dev = torch.device(device if device and device != "cpu" and not str(device).isdigit() else f"cuda:{int(device) if str(device).isdigit() else 0}")
        save_dir = Path(project) / name
        save_dir.mkdir(parents=True, exist_ok=True)
        model = load_model(weights, data_dict.get("model_cfg"), device=dev)
    if dev.type != "cuda":
        raise RuntimeError("skyeye.cli.validate runs the B200 path; there is no CPU fallback")
    stride = int(model.stride.max())
    img_size = check_img_size(img_size, s=stride)
    model.eval()
    nc = 1 if single_cls else int(data_dict["nc"])
    iouv = torch.linspace(0.5, 0.95, 10, device=dev)
    niou = iouv.numel()
    if dataloader is None:
        root = Path(data_dict.get("path", "."))
        dataloader = FolderLoader(root / data_dict[task if task in ("train", "val", "test") else "val"], img_size, batch_size, stride)
        model(torch.zeros(1, 3, img_size, img_size, dtype=torch.uint8, device=dev))  # warm-up (validate.py:208-209)
    names = dict(enumerate(getattr(model, "names", [str(i) for i in range(nc)])))
    seen, dt = 0, [0.0, 0.0, 0.0]
    mp = mr = map50 = map_ = 0.0
    loss = torch.zeros(3, device=dev)
    jdict, stats, ap_class = [], [], []
    p = r = ap = ap50 = np.zeros(0)
    shape_seen = None
    for img, targets, paths, shapes in dataloader:
        t1 = time_sync()
        img = img.to(dev, non_blocking=True)
        if img.dtype != torch.uint8:
            img = img.float()
            if img.max() > 1.5:
                img = img / 255.0
        img = img.contiguous()
        targets = targets.to(dev).float()
        nb, _, height, width = img.shape
        shape_seen = (nb, 3, height, width)
        t2 = time_sync()
        dt[0] += t2 - t1
        out, train_out = model(img)                                   # validate.py:245
        t3 = time_sync()
        dt[1] += t3 - t2
        if compute_loss:
            loss += compute_loss([x.float() for x in train_out], targets)[1]
        targets[:, 2:] *= torch.tensor([width, height, width, height], device=dev, dtype=torch.float32)
        out = non_max_suppression(out, conf_thres, iou_thres, multi_label=True, agnostic=single_cls,
                                  max_detections=max_det, compat=nms_compat)  # validate.py:255
        dt[2] += time_sync() - t3
        for si, pred in enumerate(out):
            labels = targets[targets[:, 0] == si, 1:]
            nl = labels.shape[0]
            tcls = labels[:, 0].tolist() if nl else []
            shape0 = shapes[si][0]
            seen += 1
            if pred.shape[0] == 0:
                if nl:
                    stats.append((np.zeros((0, niou), bool), np.zeros(0), np.zeros(0), np.asarray(tcls)))
                continue
            pred = pred.clone()
            if single_cls:
                pred[:, 5] = 0
            predn = pred.clone()
            scale_boxes((height, width), predn[:, :4], shape0, shapes[si][1])
            if nl:
                tbox = xywh2xyxy(labels[:, 1:5])
                scale_boxes((height, width), tbox, shape0, shapes[si][1])
                correct = process_batch(predn, torch.cat((labels[:, 0:1], tbox), 1), iouv)
            else:
                correct = torch.zeros(pred.shape[0], niou, dtype=torch.bool)
            stats.append((correct.cpu().numpy(), pred[:, 4].cpu().numpy(), pred[:, 5].cpu().numpy(), np.asarray(tcls)))
            if save_txt:
                save_one_txt(predn.cpu(), save_conf, shape0, Path(save_dir) / "labels" / (Path(paths[si]).stem + ".txt"))
            if save_json:
                save_one_json(predn.cpu(), jdict, paths[si], list(range(max(nc, 1))))
    nt = np.zeros(1)
    if stats:
        tp, conf, pcls, tcls_all = (np.concatenate(x, 0) for x in zip(*stats))
        if tp.shape[0] and tp.any():
            p, r, ap, f1, ap_class = ap_per_class(tp, conf, pcls, tcls_all, names=names)
            ap50, ap = ap[:, 0], ap.mean(1)
            mp, mr, map50, map_ = float(p.mean()), float(r.mean()), float(ap50.mean()), float(ap.mean())
        nt = np.bincount(tcls_all.astype(np.int64), minlength=nc)
    pf = "%20s" + "%11i" * 2 + "%11.3g" * 4
    LOGGER.info(("%20s" + "%11s" * 6) % ("Class", "Images", "Labels", "P", "R", "mAP@.5", "mAP@.5:.95"))
    LOGGER.info(pf % ("all", seen, int(nt.sum()), mp, mr, map50, map_))
    if (verbose or (nc < 50 and not training)) and nc > 1 and len(ap_class):
        for i, c in enumerate(ap_class):
            LOGGER.info(pf % (names.get(int(c), str(c)), seen, int(nt[c]), p[i], r[i], ap50[i], ap[i]))
    if seen and not training:
        t = tuple(x / seen * 1e3 for x in dt)
        LOGGER.info(f"Speed: %.1fms pre-process, %.1fms inference, %.1fms NMS per image at shape {shape_seen}" % t)
    if save_json and jdict:
        stem = Path(weights).stem if weights else "model"
        with open(Path(save_dir) / f"{stem}_predictions.json", "w") as f:
            json.dump(jdict, f)
    n_batches = max(len(dataloader), 1) if hasattr(dataloader, "__len__") else 1
    return (mp, mr, map50, map_, *(loss.cpu() / n_batches).tolist())


def parse_opt(argv=None):
    ap = argparse.ArgumentParser(prog="skyeye.cli.validate")
    ap.add_argument("--data", type=str, default="configs/data/drone.yaml", help="dataset yaml")
    ap.add_argument("--weights", type=str, default=None, help="model weights (.pt); random init when omitted")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--img-size", "--imgsz", "--img", type=int, default=640)
    ap.add_argument("--conf-thres", type=float, default=0.001)
    ap.add_argument("--iou-thres", type=float, default=0.6)
    ap.add_argument("--task", default="val")
    ap.add_argument("--device", default="0")
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--single-cls", action="store_true")
    ap.add_argument("--augment", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--save-txt", action="store_true")
    ap.add_argument("--save-hybrid", action="store_true")
    ap.add_argument("--save-conf", action="store_true")
    ap.add_argument("--save-json", action="store_true")
    ap.add_argument("--project", default="runs/val")
    ap.add_argument("--name", default="exp")
    ap.add_argument("--exist-ok", action="store_true")
    ap.add_argument("--half", action="store_true")
    ap.add_argument("--nms-compat", default="fixed", choices=["fixed", "reference"])
    return ap.parse_args(argv)


def main(argv=None):
    logging.basicConfig(format="%(message)s", level=logging.INFO)
    opt = parse_opt(argv)
    res = validate(**vars(opt))
    print(json.dumps({"P": res[0], "R": res[1], "mAP@.5": res[2], "mAP@.5:.95": res[3]}))
    return res


if __name__ == "__main__":
    main()
