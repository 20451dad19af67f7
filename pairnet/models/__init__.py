from .frameworks import *  # noqa: F401,F403
from .losses import *  # noqa: F401,F403
from .relation_heads import *  # noqa: F401,F403
