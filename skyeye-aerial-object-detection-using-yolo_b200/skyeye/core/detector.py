"""README module path (README.md:41: ``from skyeye.core.detector import SkyEyeDetector``)."""
from .models.detector import (EnhancedSkyEyeDetector, Results, SkyEyeDetector, construct_model,  # noqa: F401
                              load_model)
