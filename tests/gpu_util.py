"""Shared helpers of the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np
import torch

import cases


def bf16r(t):
    return t.to(torch.bfloat16).float()


def randn(key, shape, std=1.0):
    return torch.from_numpy((cases.rng(*key).standard_normal(shape, dtype=np.float32) * std).astype(np.float32))


def rel_err(a, b):
    """max |a - b| / max |b|  (the tolerance convention of SURVEY.md §8d)."""
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
