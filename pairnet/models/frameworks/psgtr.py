"""reference path ``pairnet/models/frameworks/psgtr.py`` -> B200-native ``PSGTr``."""
from pairnet_b200.detector import PSGTr  # noqa: F401
