"""reference path ``pairnet/models/losses/seg_losses.py`` (only ``BCEWithLogitsLoss`` is used by Pair-Net)."""
from pairnet_b200.losses import BCEWithLogitsLoss  # noqa: F401
