"""Loss modules named by the Pair-Net configs (``configs/mask2former/pairnet.py:153-190``), built by
``CrossHead2.__init__`` so the reference config constructs unchanged and a reference checkpoint (which carries
``rel_cls_loss.cum_samples``) loads with ``strict=True``.  Plain torch on the device: the losses are a few [B*K, C]
element-wise passes, SURVEY 8f rank 2 -- not a hot kernel."""
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


def _reduce(loss, weight, reduction, avg_factor):
    """mmdet ``weight_reduce_loss``."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == "mean" else (loss.sum() if reduction == "sum" else loss)
    if reduction == "mean":
        return loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)
    if reduction == "none":
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


@LOSSES.register_module()
class SeesawLoss(nn.Module):
    """Seesaw loss (Wang et al., CVPR 2021) as mmdet 2.25.1 ``SeesawLoss`` implements it -- the ``rel_cls_loss`` of
    ``configs/mask2former/pairnet.py:153-158``, called at ``pairnet_head.py:538-545`` on ``[r_cls_scores | 2 dummy
    objectness logits]`` and read through ``["loss_cls_classes"]``.

    Carries mmdet's persistent buffer ``cum_samples`` [num_classes + 1] (running per-class sample counts), so a
    reference checkpoint's ``bbox_head.rel_cls_loss.cum_samples`` loads with ``strict=True``.  mmdet builds DDP with
    ``broadcast_buffers=False``: the buffer is per rank and diverges across ranks in the reference too."""

    def __init__(self, use_sigmoid=False, p=0.8, q=2.0, num_classes=1203, eps=1e-2, reduction="mean", loss_weight=1.0,
                 return_dict=True):
        super().__init__()
        assert not use_sigmoid
        self.use_sigmoid = False
        self.p, self.q, self.num_classes, self.eps = p, q, num_classes, eps
        self.reduction, self.loss_weight, self.return_dict = reduction, loss_weight, return_dict
        self.register_buffer("cum_samples", torch.zeros(num_classes + 1, dtype=torch.float))

    def _seesaw_ce(self, cls_score, labels, label_weights, cum_samples, reduction, avg_factor):
        C = self.num_classes
        onehot = F.one_hot(labels, C)
        w = cls_score.new_ones(onehot.size())
        if self.p > 0:  # mitigation factor
            ratio = cum_samples[None, :].clamp(min=1) / cum_samples[:, None].clamp(min=1)
            idx = (ratio < 1.0).float()
            sw = ratio.pow(self.p) * idx + (1 - idx)
            w = w * sw[labels.long(), :]
        if self.q > 0:  # compensation factor
            scores = F.softmax(cls_score.detach(), dim=1)
            self_scores = scores[torch.arange(0, len(scores), device=scores.device).long(), labels.long()]
            sm = scores / self_scores[:, None].clamp(min=self.eps)
            idx = (sm > 1.0).float()
            w = w * (sm.pow(self.q) * idx + (1 - idx))
        cls_score = cls_score + (w.log() * (1 - onehot))
        loss = F.cross_entropy(cls_score, labels, weight=None, reduction="none")
        return _reduce(loss, label_weights.float() if label_weights is not None else None, reduction, avg_factor)

    def forward(self, cls_score, labels, label_weights=None, avg_factor=None, reduction_override=None):
        reduction = reduction_override if reduction_override else self.reduction
        assert cls_score.size(-1) == self.num_classes + 2
        pos = labels < self.num_classes
        obj_labels = (labels == self.num_classes).long()
        for u in labels.unique():  # accumulate the samples of each category
            self.cum_samples[u] += (labels == u.item()).sum()
        label_weights = label_weights.float() if label_weights is not None else labels.new_ones(labels.size(), dtype=torch.float)
        cls_c, cls_o = cls_score[..., :-2], cls_score[..., -2:]
        if pos.sum() > 0:
            loss_c = self.loss_weight * self._seesaw_ce(cls_c[pos], labels[pos], label_weights[pos],
                                                        self.cum_samples[:self.num_classes], reduction, avg_factor)
        else:
            loss_c = cls_c[pos].sum()
        loss_o = self.loss_weight * _reduce(F.cross_entropy(cls_o, obj_labels, reduction="none"), label_weights, reduction,
                                            avg_factor)
        if self.return_dict:
            return dict(loss_cls_objectness=loss_o, loss_cls_classes=loss_c)
        return loss_c + loss_o


@LOSSES.register_module()
class DiceLoss(nn.Module):
    """mmdet ``DiceLoss`` (``loss_dice`` of the config; built by the reference head, unused by its ``loss()``)."""

    def __init__(self, use_sigmoid=True, activate=True, reduction="mean", naive_dice=False, loss_weight=1.0, eps=1e-3):
        super().__init__()
        self.use_sigmoid, self.activate, self.reduction = use_sigmoid, activate, reduction
        self.naive_dice, self.loss_weight, self.eps = naive_dice, loss_weight, eps

    def forward(self, pred, target, weight=None, reduction_override=None, avg_factor=None):
        reduction = reduction_override if reduction_override else self.reduction
        if self.activate:
            assert self.use_sigmoid
            pred = pred.sigmoid()
        inp, tgt = pred.flatten(1), target.flatten(1).float()
        a = torch.sum(inp * tgt, 1)
        if self.naive_dice:
            d = (2 * a + self.eps) / (torch.sum(inp, 1) + torch.sum(tgt, 1) + self.eps)
        else:
            d = 2 * a / ((torch.sum(inp * inp, 1) + self.eps) + (torch.sum(tgt * tgt, 1) + self.eps))
        return self.loss_weight * _reduce(1 - d, weight, reduction, avg_factor)


@LOSSES.register_module()
class FocalLoss(nn.Module):
    """mmdet sigmoid ``FocalLoss`` (python path), for configs that name it."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0, activated=False):
        super().__init__()
        assert use_sigmoid
        self.use_sigmoid, self.gamma, self.alpha = True, gamma, alpha
        self.reduction, self.loss_weight, self.activated = reduction, loss_weight, activated

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        reduction = reduction_override if reduction_override else self.reduction
        if target.dim() == 1:
            target = F.one_hot(target, num_classes=pred.size(1) + 1)[:, :pred.size(1)]
        target = target.type_as(pred)
        p = pred if self.activated else pred.sigmoid()
        pt = (1 - p) * target + p * (1 - target)
        fw = (self.alpha * target + (1 - self.alpha) * (1 - target)) * pt.pow(self.gamma)
        bce = F.binary_cross_entropy(p, target, reduction="none") if self.activated else \
            F.binary_cross_entropy_with_logits(pred, target, reduction="none")
        loss = bce * fw
        if weight is not None and weight.dim() != loss.dim():
            weight = weight.view(-1, 1)
        return self.loss_weight * _reduce(loss, weight, reduction, avg_factor)
