"""Import-path mirror of the reference package so that ``custom_imports`` in
``configs/mask2former/pairnet.py:214-225`` and user code (``from pairnet.models... import CrossHead2``)
resolve to the B200-native implementation in ``pairnet_b200``."""
