from .cnn_factory import ConvTiny, creat_cnn  # noqa: F401
from .psgtr import PSGTr  # noqa: F401
