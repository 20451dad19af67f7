"""Loss modules named by the Pair-Net configs (``configs/mask2former/pairnet.py:153-190``).  They are built
by ``CrossHead2.__init__`` so the reference config constructs unchanged, but the training path
(``pairnet_head.py:419-718``) is SURVEY §8f rank 2 and not part of this round's hot path: CE / BCE are
plain torch; Seesaw / Dice hold their hyper-parameters and raise when called."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .registry import LOSSES


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    def __init__(self, use_sigmoid=False, use_mask=False, reduction="mean", class_weight=None, ignore_index=None,
                 loss_weight=1.0, avg_non_ignore=False):
        super().__init__()
        self.use_sigmoid, self.reduction, self.loss_weight = use_sigmoid, reduction, loss_weight
        self.class_weight = class_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, **kwargs):
        cw = None if self.class_weight is None else cls_score.new_tensor(self.class_weight)
        if self.use_sigmoid:
            loss = F.binary_cross_entropy_with_logits(cls_score, label.float(), reduction="none")
        else:
            loss = F.cross_entropy(cls_score, label, weight=cw, reduction="none")
        if weight is not None:
            loss = loss * weight
        loss = loss.sum() / avg_factor if avg_factor is not None else (loss.mean() if self.reduction == "mean" else loss.sum())
        return self.loss_weight * loss


@LOSSES.register_module()
class BCEWithLogitsLoss(nn.Module):
    """reference ``pairnet/models/losses/seg_losses.py:153-166`` (pos_weight passed per call)."""

    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, pos_weight=None):
        pw = None if pos_weight is None else torch.as_tensor(pos_weight, device=pred.device, dtype=pred.dtype)
        return self.loss_weight * F.binary_cross_entropy_with_logits(pred, target, pos_weight=pw, reduction=self.reduction)


class _ConfigOnlyLoss(nn.Module):
    use_sigmoid = False

    def __init__(self, **cfg):
        super().__init__()
        self.cfg = cfg
        self.use_sigmoid = cfg.get("use_sigmoid", False)
        self.loss_weight = cfg.get("loss_weight", 1.0)

    def forward(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__}: training losses are SURVEY §8f rank 2 (not built yet)")


@LOSSES.register_module()
class SeesawLoss(_ConfigOnlyLoss):
    pass


@LOSSES.register_module()
class DiceLoss(_ConfigOnlyLoss):
    pass


@LOSSES.register_module()
class FocalLoss(_ConfigOnlyLoss):
    pass
