from .seg_losses import BCEWithLogitsLoss  # noqa: F401
