"""Pin ``oracle/train.py`` to the REFERENCE's own training targets + losses (TEST INFRASTRUCTURE, dev container only).

    python -m oracle.pin_train            # check + mint tests/golden/train_ref_*.npz
    python -m oracle.pin_train --check

Executed from /root/reference by path, nothing copied: ``CrossHead2.loss / loss_single / get_targets /
_get_target_single`` (pairnet_head.py:419-718), ``MaskHungarianAssigner`` + ``CrossEntropyLossCost`` + ``DiceCost`` +
``MaskPseudoSampler`` (panoptic_heads/mask_hungarian_assigner.py), ``IdMatcher`` (relation_heads/approaches/matcher.py)
and ``BCEWithLogitsLoss`` (losses/seg_losses.py:153-166), over the constructor shims of ``oracle/pin_reference.py``
plus the ones below.  What the shims supply here (absent mmdet / mmcv, restated -> stays "parity unpinned"):
``ClassificationCost``, ``CrossEntropyLoss``, ``SeesawLoss``, ``AssignResult`` / ``SamplingResult`` containers,
``multi_apply``, ``point_sample``.  The oracle must agree BIT FOR BIT with the reference's four losses (same torch RNG
stream for the sampled points) on the head outputs of two fixture cases."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pin_reference as pr
from . import train as ot

TRAIN_CASES = [  # (tag, B, (H4,W4), seed of inputs, seed of GT / sample points, gt mask scale)
    ("b2_32x48", 2, (32, 48), 21, 301, 4),
    ("b3_24x40", 3, (24, 40), 25, 302, 2),
]

TRAIN_CFG = dict(
    id_assigner=dict(type="IdMatcher", sub_id_cost=dict(type="ClassificationCost", weight=1.0),
                     obj_id_cost=dict(type="ClassificationCost", weight=1.0),
                     r_cls_cost=dict(type="ClassificationCost", weight=0.0)),
    num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75,
    mask_assigner=dict(type="MaskHungarianAssigner", cls_cost=dict(type="ClassificationCost", weight=2.0),
                       mask_cost=dict(type="CrossEntropyLossCost", weight=5.0, use_sigmoid=True),
                       dice_cost=dict(type="DiceCost", weight=5.0, pred_act=True, eps=1.0)),
    sampler=dict(type="MaskPseudoSampler"))   # configs/mask2former/pairnet.py:190-207


def install_train_shims():
    pr.install_shims()
    M = pr._module
    assigners, samplers, costs, losses = pr._Registry(), pr._Registry(), pr._Registry(), pr._Registry()

    class AssignResult:
        def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
            self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels

    class BaseAssigner:
        pass

    class BaseSampler:
        pass

    class SamplingResult:
        pass

    def build_from(reg):
        def build(cfg, **kw):
            cfg = dict(cfg)
            return reg.d[cfg.pop("type")](**cfg)
        return build

    @costs.register_module()
    class ClassificationCost:  # mmdet 2.25.1 (not on disk): restated
        def __init__(self, weight=1.0):
            self.weight = weight

        def __call__(self, cls_pred, gt_labels):
            return -cls_pred.softmax(-1)[:, gt_labels] * self.weight

    def multi_apply(func, *args, **kwargs):
        from functools import partial
        pfunc = partial(func, **kwargs) if kwargs else func
        return tuple(map(list, zip(*map(pfunc, *args))))

    class CrossEntropyLoss(nn.Module):  # mmdet (not on disk): softmax CE, class_weight, mean reduction, loss_weight
        def __init__(self, use_sigmoid=False, reduction="mean", class_weight=None, loss_weight=1.0, **kw):
            super().__init__()
            self.use_sigmoid, self.class_weight, self.loss_weight = use_sigmoid, class_weight, loss_weight

        def forward(self, cls_score, label):
            cw = None if self.class_weight is None else cls_score.new_tensor(self.class_weight)
            return self.loss_weight * F.cross_entropy(cls_score, label, weight=cw, reduction="none").mean()

    class SeesawLoss(nn.Module):  # mmdet (not on disk): restated in oracle.train.OSeesawLoss
        def __init__(self, num_classes=1203, loss_weight=1.0, return_dict=True, **kw):
            super().__init__()
            self.use_sigmoid = False
            self.impl = ot.OSeesawLoss(num_classes, loss_weight=loss_weight)
            self.register_buffer("cum_samples", self.impl.cum_samples)

        def forward(self, cls_score, labels):
            self.impl.cum_samples = self.cum_samples
            return self.impl(cls_score, labels)

    M("mmcv.ops", point_sample=ot.point_sample)
    M("mmdet.core", build_assigner=build_from(assigners), build_sampler=build_from(samplers), multi_apply=multi_apply,
      AssignResult=AssignResult, BaseAssigner=BaseAssigner, bbox_cxcywh_to_xyxy=None)
    M("mmdet.core.bbox")
    M("mmdet.core.bbox.assigners")
    M("mmdet.core.bbox.assigners.assign_result", AssignResult=AssignResult)
    M("mmdet.core.bbox.assigners.base_assigner", BaseAssigner=BaseAssigner)
    M("mmdet.core.bbox.builder", BBOX_ASSIGNERS=assigners, BBOX_SAMPLERS=samplers)
    M("mmdet.core.bbox.iou_calculators", bbox_overlaps=None)
    M("mmdet.core.bbox.match_costs", build_match_cost=build_from(costs))
    M("mmdet.core.bbox.match_costs.builder", build_match_cost=build_from(costs))
    M("mmdet.core.bbox.match_costs.match_cost", MATCH_COST=costs)
    M("mmdet.core.bbox.transforms", bbox_cxcywh_to_xyxy=None, bbox_xyxy_to_cxcywh=None)
    M("mmdet.core.bbox.samplers")
    M("mmdet.core.bbox.samplers.base_sampler", BaseSampler=BaseSampler)
    M("mmdet.core.bbox.samplers.sampling_result", SamplingResult=SamplingResult)
    M("mmdet.models.losses")
    M("mmdet.models.losses.utils", weighted_loss=lambda f: f)
    # the reference's own in-repo pieces
    pr._load("ref_mask_hungarian_assigner", f"{pr.REF}/pairnet/models/panoptic_heads/mask_hungarian_assigner.py")
    pr._load("ref_matcher", f"{pr.REF}/pairnet/models/relation_heads/approaches/matcher.py")
    builder = sys.modules["mmdet.models.builder"]
    builder.LOSSES = losses
    seg = pr._load("ref_seg_losses", f"{pr.REF}/pairnet/models/losses/seg_losses.py")

    def build_loss(cfg):
        cfg = dict(cfg)
        t = cfg.pop("type")
        if t == "BCEWithLogitsLoss":
            m = seg.BCEWithLogitsLoss(**cfg)     # the reference's own class
            m.use_sigmoid = True
            return m
        if t == "SeesawLoss":
            return SeesawLoss(**cfg)
        if t == "CrossEntropyLoss":
            return CrossEntropyLoss(**cfg)
        m = nn.Module()                           # DiceLoss: built, never called by loss()
        m.use_sigmoid = cfg.get("use_sigmoid", False)
        return m
    builder.build_loss = build_loss


def build_reference_train_head(N=100, R=100):
    install_train_shims()
    fb = pr._load("ref_facebook_detr", f"{pr.REF}/pairnet/models/relation_heads/facebook_detr.py")
    tr = sys.modules["mmcv.cnn.bricks.transformer"]

    class RefLayer(fb.BaseTransformerLayer2):
        def forward(self, *a, **kw):
            return super().forward(*a, **kw)[0]
    tr.ATTENTION_CLS, tr.LAYER_CLS = fb.MultiheadAttention2, RefLayer
    ph = pr._load("ref_pairnet_head_train", f"{pr.REF}/pairnet/models/relation_heads/pairnet_head.py")
    cfg = pr.reference_head_cfg(N, R)
    for k in ("transformer_decoder", "relation_decoder"):
        cfg[k]["transformerlayers"].pop("feedforward_channels")
    cfg.update(
        train_cfg=pr.AttrDict(TRAIN_CFG),
        rel_cls_loss=dict(type="SeesawLoss", num_classes=56, return_dict=True, loss_weight=2.0),
        subobj_cls_loss=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=4.0, reduction="mean",
                             class_weight=[1.0] * 134),
        importance_match_loss=dict(type="BCEWithLogitsLoss", reduction="mean", loss_weight=5.0))
    return ph.CrossHead2(**cfg).eval()


def case_inputs(B, hw4, seed, gt_seed, scale):
    """Head outputs of the fixture case (oracle forward; bit-equal to the reference forward by pin_reference) + GT."""
    from .head import HeadHyper, OCrossHead2
    from .make_golden import small_head_inputs
    from .weights import fixture_state_dict
    oracle = OCrossHead2(HeadHyper(with_pixel_decoder=False)).eval()
    oracle.load_state_dict(fixture_state_dict(oracle, 10086))
    mf, mems = small_head_inputs(B, hw4, seed)
    with torch.no_grad():
        cls, msk = oracle.forward_from_memories(mf, mems)
    rels, labels, masks = ot.synthetic_gt(B, (hw4[0] * scale, hw4[1] * scale), gt_seed)
    return cls, msk, rels, labels, masks


def main(check_only=False):
    assert os.path.isdir(pr.REF), "the reference tree is only present in the dev container"
    head = build_reference_train_head()
    for tag, B, hw4, seed, gt_seed, scale in TRAIN_CASES:
        cls, msk, rels, labels, masks = case_inputs(B, hw4, seed, gt_seed, scale)
        head.rel_cls_loss.cum_samples.zero_()
        torch.manual_seed(gt_seed)
        with torch.no_grad():
            ref = head.loss(cls, msk, rels, None, labels, masks, [dict()] * B)
        torch.manual_seed(gt_seed)
        with torch.no_grad():
            got, tg = ot.loss(cls, msk, rels, labels, masks, return_targets=True)
        for k in ref:
            assert torch.equal(ref[k], got[k]), f"{tag}: oracle {k} {float(got[k])!r} != reference {float(ref[k])!r}"
        print(f"pinned {tag}: reference CrossHead2.loss == oracle, bit-equal: " +
              ", ".join(f"{k}={float(v):.6f}" for k, v in ref.items()) +
              f"; matched triplets {int((tg['r_label_weights'] > 0).sum())}, gt_importance positives "
              f"{int((tg['gt_importance'] > 0).sum())}")
        if not check_only:
            np.savez_compressed(
                os.path.join(pr.GOLDEN, f"train_ref_{tag}.npz"),
                **{k: v.numpy() for k, v in ref.items()}, **{k: v.numpy() for k, v in tg.items()},
                meta=np.array([B, hw4[0], hw4[1], seed, gt_seed, scale]))
    return 0


if __name__ == "__main__":
    sys.exit(main(check_only="--check" in sys.argv))
