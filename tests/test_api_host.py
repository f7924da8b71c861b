"""Host side of the drop-in API (no GPU): checkpoint formats, on-disk writers, CLI argument surface, the folder loader and
the Results drawing.  Reference: skyeye/core/models/detector.py:343-371, skyeye/cli/validate.py:31-68,112-171,
README.md:39-54,69."""
import json
import warnings

import numpy as np
import pytest
import torch

from oracle import model as om


def _model(variant="skyeye_tiny_l"):
    from skyeye.core.models.detector import EnhancedSkyEyeDetector, SkyEyeDetector
    cfg = om.get_cfg(variant)
    d = dict(nc=cfg["nc"], base_channels=cfg["base_channels"], depth_multiple=cfg["depth_multiple"], width_multiple=1.0,
             head="conv", head_dim=cfg["head_dim"])
    return (EnhancedSkyEyeDetector if cfg["enhanced"] else SkyEyeDetector)(d), cfg


def _sd_without_transformer(cfg, seed=3):
    return {k: v for k, v in om.make_state_dict(cfg, seed).items() if not k.startswith("head_transformers.")}


@pytest.mark.parametrize("fmt", ["model", "state_dict", "bare"])
def test_load_from_pretrained_accepts_the_three_reference_checkpoint_formats(tmp_path, fmt):
    m, cfg = _model()
    sd = _sd_without_transformer(cfg)
    if fmt == "model":        # {'model': nn.Module}  (detector.py:356-357)
        src, _ = _model()
        src.load_state_dict(sd, strict=True)
        ckpt = {"model": src, "epoch": 3}
    elif fmt == "state_dict":  # {'state_dict': ...}   (detector.py:358-359)
        ckpt = {"state_dict": sd}
    else:                      # bare state dict
        ckpt = sd
    path = tmp_path / "w.pt"
    torch.save(ckpt, path)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_from_pretrained(str(path))
    after = m.state_dict()
    assert all(torch.equal(after[k], sd[k].to(after[k].dtype)) for k in sd), "checkpoint tensors were not loaded"
    assert any(not torch.equal(before[k], after[k]) for k in sd)
    assert m.uninitialized_keys == [] and m.unexpected_keys == []


def test_load_from_pretrained_filters_by_name_and_shape_and_reports_what_stays_random(tmp_path, capsys):
    m, cfg = _model()
    sd = _sd_without_transformer(cfg)
    good = dict(sd)
    k_shape = "neck.lateral_conv5.conv.weight"
    good[k_shape] = torch.zeros(3, 3)                       # wrong shape -> skipped (detector.py:362-364)
    good["not.a.parameter"] = torch.zeros(1)                # unknown name -> skipped
    del good["detection_head.detection_layers.0.bias"]      # missing -> keeps its init
    path = tmp_path / "w.pt"
    torch.save({"state_dict": good}, path)
    ref_w = m.state_dict()[k_shape].clone()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m.load_from_pretrained(str(path))
    out = capsys.readouterr().out
    n_own = len(m.state_dict())
    assert f"Loaded {len(sd) - 2}/{n_own} layers" in out    # the reference's log line (detector.py:369)
    assert torch.equal(m.state_dict()[k_shape], ref_w)
    assert set(m.uninitialized_keys) == {k_shape, "detection_head.detection_layers.0.bias"}
    assert set(m.unexpected_keys) == {k_shape, "not.a.parameter"}
    assert any("keep their random initialisation" in str(x.message) for x in w)


def test_reference_architecture_is_the_default_head_and_transformer_heads_are_opt_in():
    from skyeye.core.models.detector import EnhancedSkyEyeDetector, parse_model
    base = dict(nc=10, base_channels=8, depth_multiple=0.33, width_multiple=1.0, head_dim=16)
    assert EnhancedSkyEyeDetector(dict(base)).head_transformers is None          # reference: bare 1x1 heads (detector.py:436-501)
    assert EnhancedSkyEyeDetector(dict(base, head="transformer")).head_transformers is not None
    assert parse_model(dict(base))["head"] == "conv" and parse_model("skyeye_l.yaml")["head"] == "transformer"


def test_weights_kwarg_selects_the_variant_config_by_file_name(tmp_path):
    from skyeye.core.detector import SkyEyeDetector   # README module path
    cfg = om.get_cfg("skyeye_s")
    path = tmp_path / "skyeye_s.pt"
    sd = om.make_state_dict(cfg, 1)
    torch.save({"state_dict": sd}, path)
    m = SkyEyeDetector(weights=str(path))             # README.md:44
    assert m.cfg["base_channels"] == 32 and m.uninitialized_keys == []
    assert torch.equal(m.state_dict()["neck.lateral_conv4.conv.weight"], sd["neck.lateral_conv4.conv.weight"])


def test_save_one_txt_format(tmp_path):
    from skyeye.cli.validate import save_one_txt
    pred = torch.tensor([[10.0, 20.0, 110.0, 220.0, 0.87654321, 3.0], [0.0, 0.0, 64.0, 48.0, 0.5, 0.0]])
    f = tmp_path / "labels" / "img.txt"
    save_one_txt(pred, True, (480, 640), f)            # shape = (h, w); gain = whwh (validate.py:41)
    save_one_txt(pred[:1], False, (480, 640), f)       # appends (validate.py:45)
    lines = f.read_text().splitlines()
    # cls cx/w cy/h bw/w bh/h conf with %g (validate.py:42-46)
    assert lines[0] == "3 0.09375 0.25 0.15625 0.416667 0.876543"
    assert lines[1] == "0 0.05 0.05 0.1 0.1 0.5"
    assert lines[2] == "3 0.09375 0.25 0.15625 0.416667"


def test_save_one_json_format():
    from pathlib import Path
    from skyeye.cli.validate import save_one_json
    pred = torch.tensor([[10.0, 20.0, 110.5556, 220.0, 0.876543219, 3.0]])
    jd = []
    save_one_json(pred, jd, Path("/data/000042.jpg"), {3: 7})
    save_one_json(pred, jd, Path("/data/frame_a.jpg"), list(range(10)))
    assert jd[0] == {"image_id": 42, "category_id": 7, "bbox": [10.0, 20.0, 100.556, 200.0], "score": 0.87654}
    assert jd[1]["image_id"] == "frame_a" and jd[1]["category_id"] == 3
    json.dumps(jd)


def test_cli_flags_match_the_readme_invocation():
    from skyeye.cli.validate import parse_opt
    o = parse_opt(["--weights", "weights/skyeye_l.pt", "--data", "configs/data/drone.yaml", "--img-size", "640"])  # README.md:69
    assert (o.weights, o.data, o.img_size) == ("weights/skyeye_l.pt", "configs/data/drone.yaml", 640)
    assert (o.conf_thres, o.iou_thres, o.batch_size, o.task) == (0.001, 0.6, 32, "val")   # validate.py:112-121 defaults
    o = parse_opt(["--save-txt", "--save-conf", "--save-json", "--single-cls", "--batch-size", "4"])
    assert o.save_txt and o.save_conf and o.save_json and o.single_cls and o.batch_size == 4


def test_folder_loader_places_labels_in_the_batch_frame_for_mixed_aspect_ratios(tmp_path):
    """Two images of different aspect ratio letterbox to different sizes; the smaller one sits top-left in the batch frame.
    validate() multiplies targets by the BATCH width / height, so every label must come back at its true pixel position."""
    cv2 = pytest.importorskip("cv2")
    from skyeye.cli.validate import FolderLoader
    (tmp_path / "images").mkdir()
    (tmp_path / "labels").mkdir()
    specs = {"a": (200, 400), "b": (400, 400)}          # (h, w): 'a' letterboxes to 64x128, 'b' to 128x128 at img_size 128
    for name, (h, w) in specs.items():
        cv2.imwrite(str(tmp_path / "images" / f"{name}.png"), np.full((h, w, 3), 200, np.uint8))
        (tmp_path / "labels" / f"{name}.txt").write_text("2 0.5 0.5 0.25 0.5\n")
    img, targets, paths, shapes = next(iter(FolderLoader(tmp_path / "images", img_size=128, batch_size=2)))
    assert tuple(img.shape) == (2, 3, 128, 128) and img.dtype == torch.uint8
    H, W = img.shape[2:]
    px = targets[:, 2:] * torch.tensor([W, H, W, H])
    for row, p in zip(targets, px):
        h0, w0 = specs[["a", "b"][int(row[0])]]
        r = min(128 / h0, 128 / w0)
        (ratio, _), (pw, ph) = shapes[int(row[0])][1]
        assert ratio == pytest.approx(r)
        exp = torch.tensor([0.5 * w0 * r + pw, 0.5 * h0 * r + ph, 0.25 * w0 * r, 0.5 * h0 * r])
        assert torch.allclose(p, exp, atol=1e-4), (p, exp)
    assert int(targets[0, 1]) == 2


def test_results_render_and_save_draw_boxes(tmp_path):
    pytest.importorskip("cv2")
    from skyeye.core.models.detector import Results
    im = np.zeros((120, 160, 3), np.uint8)
    pred = torch.tensor([[20.0, 30.0, 100.0, 90.0, 0.9, 1.0]])
    res = Results([pred], [im], ["car", "bus"], ["frame7.jpg"])
    drawn = res.render()[0]
    assert drawn.shape == im.shape and drawn.any() and not im.any()      # drawn on a copy
    assert drawn[30, 60].any() and not drawn[60, 60].any()               # box edge painted, interior untouched
    out = res.save(tmp_path / "o")
    assert (tmp_path / "o" / "frame7.jpg").exists()
    assert (tmp_path / "o" / "frame7.txt").read_text().strip() == "1 20 30 100 90 0.9"
    # 7-column rows of the reference wrapper (centre form, class id in column 6) are drawn as the same box
    ref_rows = torch.tensor([[60.0, 60.0, 80.0, 60.0, 0.9, 0.8, 1.0]])
    assert np.array_equal(Results([ref_rows], [im], ["car", "bus"]).render()[0], drawn)
    res.show()
