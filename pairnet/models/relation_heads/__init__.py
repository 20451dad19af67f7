from .pairnet_head import CrossHead2  # noqa: F401
