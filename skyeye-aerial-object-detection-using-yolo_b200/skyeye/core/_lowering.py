"""Helpers shared by the module classes: run a single module eagerly through a throw-away plan
(used by per-block parity tests; the detector caches one plan per input shape instead)."""
from __future__ import annotations

import torch

from .. import engine as E
from ..engine import Plan, View


def run_module(mod, x: torch.Tensor) -> torch.Tensor:
    """NCHW float tensor in -> NCHW fp32 tensor out through the module's native lowering."""
    if not x.is_cuda:
        raise RuntimeError("skyeye (B200) modules run on CUDA only; there is no CPU fallback")
    plan = Plan(x.device)
    xin = E.from_nchw(x.float())
    out = mod.lower(plan, xin)
    plan.run()
    return out.nchw().float()


def ref(mod, suffix: str = ""):
    """Reference state-dict path of ``mod`` (set by SkyEyeDetector._build_plan from named_modules) + suffix; None when
    the module is lowered stand-alone.  Labels the plan's outputs for Plan.run_teacher_forced."""
    r = getattr(mod, "_ref", None)
    return None if r is None else r + suffix
