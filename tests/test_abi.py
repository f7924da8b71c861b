"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads without a GPU and exports
exactly the symbols include/skyeye_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "skyeye_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def native():
    from skyeye import _native
    if _native.needs_build():
        if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            pytest.skip("library not built and nvcc absent")
        _native.build()
    return _native


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("skb_conv2d_bf16", "skb_flash_attn_bf16", "skb_decode_f32", "skb_nms_f32", "skb_nms_batched_f32", "skb_last_error"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol(native):
    L = ctypes.CDLL(native.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/skyeye_b200.h but not exported"
    assert set(native.EXPORTED_SYMBOLS) == set(_declared_symbols())


def test_version_and_error_string_without_gpu(native):
    L = native.lib()
    assert L.skb_version() == 100
    assert isinstance(native.last_error(), str)


def test_workspace_queries_are_pure_host_functions(native):
    L = native.lib()
    assert L.skb_nms_workspace_bytes(30000) > 30000 * 469 * 8
    assert L.skb_nms_batched_workspace_bytes(64, 50000, 10, 0) >= 8 * 64 * 50000      # one 64-bit key per candidate box (no sort double buffer)
    assert L.skb_cbam_workspace_bytes(16, 80, 80, 512) > 0
    assert L.skb_cla_workspace_bytes(16, 160, 160, 4) >= 16 * 4 * 160 * 160 * 4


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "skyeye-aerial-object-detection-using-yolo_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_no_gpu_means_loud_failure_not_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from skyeye.utils.nms import nms
    with pytest.raises(RuntimeError):
        nms(torch.zeros(4, 4), torch.zeros(4), 0.5)
    from skyeye.core.detector import construct_model
    m = construct_model("skyeye_s.yaml")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))
