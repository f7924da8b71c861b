"""Box ops and NMS of the SkyEye path (B200-native)."""
from .metrics import non_max_suppression  # noqa: F401
from .nms import nms  # noqa: F401
