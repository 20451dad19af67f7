"""reference path ``pairnet/models/relation_heads/pairnet_head.py`` -> B200-native ``CrossHead2``."""
from pairnet_b200.head import CrossHead2  # noqa: F401
