"""reference path ``pairnet/models/frameworks/cnn_factory.py`` -> weights container + CUDA ConvTiny."""
from pairnet_b200.head import ConvTiny, creat_cnn  # noqa: F401
