"""skyeye (B200-native): drop-in for the SkyEye batched detector forward path
(SkyEyeDetector.__call__ -> backbone -> neck -> CLA -> transformer heads -> decode -> NMS).

Host code is Python/PyTorch (tensors, streams, torch.distributed); all compute on the path runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/skyeye_b200.h (libskyeye_b200.so,
bound with ctypes in skyeye/_native.py).  There is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
