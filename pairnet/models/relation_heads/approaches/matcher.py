"""Import target of ``custom_imports`` (reference ``approaches/matcher.py``): the Hungarian triplet matcher the config
names (``configs/mask2former/pairnet.py:191-196``) lives in ``pairnet_b200/training.py``."""
from pairnet_b200.training import IdMatcher  # noqa: F401
