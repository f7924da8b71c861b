"""Host logic of the teacher-forced parity check (no GPU): every launch of a detector plan labels what it writes
with the reference module path, and those labels are exactly the intermediates the oracle records (``taps``),
with matching shapes.  The plan is built on the meta device (no kernels run)."""
import pytest
import torch

from oracle import model as om


@pytest.mark.parametrize("variant", ["skyeye_s", "skyeye_l"])
def test_plan_output_labels_cover_the_oracle_taps(variant):
    from skyeye.core.detector import construct_model
    from skyeye.engine import Plan, View
    cfg = om.get_cfg(variant)
    m = construct_model(f"{variant}.yaml")
    plan = m._build_plan(1, 64, 96, torch.device("meta"))
    taps = {}
    om.forward(torch.rand(1, 3, 64, 96), om.make_state_dict(cfg, 0), cfg, emu="bf16", taps=taps)
    labels = [o["label"] for outs in plan.outs for o in outs]
    assert all(plan.outs), [n for n, o in zip(plan.names, plan.outs) if not o]   # no unlabelled launch
    assert sorted(set(labels)) == sorted(taps)                                   # same set of intermediates
    for outs in plan.outs:
        for o in outs:
            got = o["view"].torch() if isinstance(o["view"], View) else o["view"]
            assert tuple(got.shape) == tuple(Plan._expected(o, taps).shape), o["label"]


def test_calibrated_state_dict_is_deterministic_and_keeps_activations_bounded():
    cfg = om.get_cfg("skyeye_nano_l")
    a = om.make_calibrated_state_dict(cfg, 0, calib_batch=8, calib_hw=(128, 128))
    b = om.make_calibrated_state_dict(cfg, 0, calib_batch=8, calib_hw=(128, 128))
    assert all(torch.equal(a[k], b[k]) for k in a)
    raw = om.make_state_dict(cfg, 0)
    changed = {k for k in a if not torch.equal(a[k], raw[k])}
    # only BN statistics and the residual-branch BN scale differ from the plain recipe: no conv / attention weight
    assert changed and all(k.endswith(("running_mean", "running_var")) or (".bottlenecks." in k and k.endswith(".cv2.bn.weight"))
                           for k in changed), sorted(changed)[:5]
    taps = {}
    om.forward(torch.rand(1, 3, 128, 128, generator=torch.Generator().manual_seed(0)), a, cfg, taps=taps)
    assert max(float(v.abs().max()) for k, v in taps.items() if k != "det") < 100.0
