"""ctypes binding of libskyeye_b200.so (C ABI declared in include/skyeye_b200.h).

The shared library is built in-tree by ``build()`` (nvcc, sm_100a) and loaded lazily.  Compute
entry points raise RuntimeError(skb_last_error()) on failure -- a missing library or a non-sm_100
device is a hard error, never a fallback.
"""
from __future__ import annotations

import ctypes
import glob
import os
import shutil
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
INCLUDE = os.path.join(_ROOT, "include")
LIB_PATH = os.path.join(_HERE, "libskyeye_b200.so")

SKB_BF16, SKB_F32, SKB_U8 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_RELU = 0, 1, 2
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


class skb_view(Structure):
    _fields_ = [("ptr", c_void_p), ("n", c_int32), ("h", c_int32), ("w", c_int32), ("c", c_int32),
                ("pitch", c_int32), ("dtype", c_int32)]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into skyeye/libskyeye_b200.so (cross-compiles on a
    CPU-only box).  Objects are built in parallel, then linked with a static cudart; libcuda is NOT
    linked (the tensor-map encoder is resolved through cudaGetDriverEntryPoint at run time)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and libskyeye_b200.so is missing or stale")
    objdir = os.path.join(_PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = list(NVCC_FLAGS)
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *flags, "-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode(errors='replace')}")
        objs.append(obj)
    tmp = LIB_PATH + ".tmp"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *objs, "-cudart", "static"])
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None

_SIGNATURES = {
    "skb_version": (c_int32, []),
    "skb_last_error": (c_char_p, []),
    "skb_device_check": (c_int32, []),
    "skb_conv2d_bf16": (c_int32, [POINTER(skb_view), c_void_p, c_void_p, POINTER(skb_view), POINTER(skb_view),
                                  c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "skb_debug_conv_trace": (c_int32, [c_void_p]),
    "skb_focus_nchw_f32": (c_int32, [c_void_p, c_int32, c_int32, c_int32, POINTER(skb_view), c_void_p]),
    "skb_focus_nchw_u8": (c_int32, [c_void_p, c_int32, c_int32, c_int32, POINTER(skb_view), c_void_p]),
    "skb_focus_conv_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "skb_focus_conv_bf16": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, POINTER(skb_view),
                                      c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "skb_focus_conv_tiles_bf16": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                            POINTER(skb_view), c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "skb_letterbox_u8": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                   c_int32, c_void_p]),
    "skb_maxpool5_bf16": (c_int32, [POINTER(skb_view), POINTER(skb_view), c_void_p]),
    "skb_spp_pools_bf16": (c_int32, [POINTER(skb_view), POINTER(skb_view), POINTER(skb_view), POINTER(skb_view), c_void_p]),
    "skb_cbam_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "skb_cbam_bf16": (c_int32, [POINTER(skb_view), c_void_p, c_void_p, c_int32, c_void_p, POINTER(skb_view),
                                c_void_p, c_size_t, c_void_p]),
    "skb_cla_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "skb_cla_core_bf16": (c_int32, [POINTER(skb_view), POINTER(skb_view), POINTER(skb_view), POINTER(skb_view),
                                    c_int32, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "skb_layernorm_bf16": (c_int32, [POINTER(skb_view), c_void_p, c_void_p, c_float, POINTER(skb_view), c_void_p]),
    "skb_flash_attn_bf16": (c_int32, [POINTER(skb_view), POINTER(skb_view), c_int32, c_float, c_void_p]),
    "skb_debug_attn_prof": (c_int32, [c_void_p, c_int32]),
    "skb_window_attn_bf16": (c_int32, [POINTER(skb_view), c_void_p, c_void_p, c_int32, POINTER(skb_view), c_int32, c_float, c_void_p]),
    "skb_window_attn2d_bf16": (c_int32, [POINTER(skb_view), c_void_p, c_void_p, c_int32, POINTER(skb_view), c_int32, c_int32, c_float,
                                         c_void_p]),
    "skb_decode_f32": (c_int32, [POINTER(skb_view), c_int32, c_int32, c_int32, POINTER(c_float), c_int32, c_int32,
                                 c_void_p, POINTER(c_void_p), c_void_p]),
    "skb_nms_workspace_bytes": (c_size_t, [c_int32]),
    "skb_nms_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "skb_nms_batched_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "skb_nms_batched_f32": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_float, c_float, POINTER(c_int32), c_int32,
                                      c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "skb_debug_nms_pair_counter": (c_int32, [c_void_p]),
    "skb_match_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "skb_match_detections_f32": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "skb_nms_batched_tiles_f32": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_int32, c_int32, c_int32, c_int32,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "skb_tile_merge_pred_f32": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded CDLL with argtypes set. Raises if the library cannot be loaded (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(needs nvcc). The B200 path has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().skb_last_error() or b"").decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libskyeye_b200 {what} failed (code {rc}): {last_error()}")
