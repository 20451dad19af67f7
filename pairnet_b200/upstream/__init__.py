"""Upstream plumbing (backbone + pixel decoder): PyTorch/cuDNN on device, outside the hand-written hot path."""
from .backbone import ResNet  # noqa: F401
from .pixel_decoder import MSDeformAttnPixelDecoder  # noqa: F401
from .swin import SwinTransformer  # noqa: F401
