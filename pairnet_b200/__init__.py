"""pairnet_b200 -- B200-native (sm_100a) implementation of Pair-Net's relation-head hot path.

Host side (this package): the reference's module-registry surface (``CrossHead2``, ``PSGTr``, config
loader) in Python.  Device side: ``csrc/`` hand-written CUDA behind the C-ABI of
``include/pairnet_b200.h``, loaded with ctypes (``_native``).  No CPU fallback for the hot path."""
from . import bricks, losses, upstream  # noqa: F401  (registers the modules)
from .detector import PSGTr, GraphedForward  # noqa: F401
from .head import ConvTiny, CrossHead2, creat_cnn  # noqa: F401
from .registry import Config, build_detector, build_head  # noqa: F401

__version__ = "0.1.0"
