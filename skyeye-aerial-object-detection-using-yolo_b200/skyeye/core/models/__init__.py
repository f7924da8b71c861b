from .blocks import BottleneckBlock, ConvolutionBlock, CSPBlock, FocusBlock, SPPBlock  # noqa: F401
from .attention import CombinedAttention, CrossLayerAttention, TransformerLayer  # noqa: F401
from .backbone import SkyEyeBackbone  # noqa: F401
from .detector import (DetectionHead, EnhancedSkyEyeDetector, FeatureNeck, Results, SkyEyeDetector,  # noqa: F401
                       construct_model, load_model, parse_model)
