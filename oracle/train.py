"""TEST INFRASTRUCTURE -- CPU restatement of the reference's training targets and losses.

Follows ``/root/reference/pairnet/models/relation_heads/pairnet_head.py``:
``loss`` :419-480, ``loss_single`` :482-564, ``get_targets`` :566-611, ``_get_target_single`` :613-718, and the
in-repo assigners / costs it builds from ``configs/mask2former/pairnet.py:190-207``:
``MaskHungarianAssigner.assign`` (``panoptic_heads/mask_hungarian_assigner.py:50-132``), ``CrossEntropyLossCost``
(:156-196), ``DiceCost`` (:220-253), ``MaskPseudoSampler.sample`` (:328-351), ``IdMatcher.assign``
(``relation_heads/approaches/matcher.py:207-274``), ``BCEWithLogitsLoss`` (``losses/seg_losses.py:153-166``).
Restated from the published upstream code because it is not on disk: mmdet 2.25.1 ``ClassificationCost``,
``CrossEntropyLoss`` (softmax, ``class_weight``, mean reduction), ``SeesawLoss``, mmcv ``point_sample``.

Pinned by ``oracle/pin_train.py``: the reference's own ``loss`` is executed from /root/reference over shims and must
agree bit for bit (same torch RNG stream for the 12 544 sample points per image); goldens ``tests/golden/train_ref_*.npz``.
Only ``tests/`` and ``bench.py``'s baseline legs may import this module."""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment


def point_sample(inp, points, align_corners=False):
    """mmcv.ops.point_sample: inp [N,C,H,W], points [N,P,2] in [0,1]x[0,1] (x, y) -> [N,C,P] (bilinear, zero padding)."""
    grid = 2.0 * points.unsqueeze(2) - 1.0
    return F.grid_sample(inp, grid, align_corners=align_corners).squeeze(3)


def classification_cost(cls_pred, gt_labels, weight):
    """mmdet ClassificationCost: -softmax(cls_pred)[:, gt_labels] * weight."""
    return -cls_pred.softmax(-1)[:, gt_labels] * weight


def mask_bce_cost(pred, gt, weight):
    """CrossEntropyLossCost(use_sigmoid=True), mask_hungarian_assigner.py:156-196."""
    pred = pred.flatten(1).float()
    gt = gt.flatten(1).float()
    n = pred.shape[1]
    pos = F.binary_cross_entropy_with_logits(pred, torch.ones_like(pred), reduction="none")
    neg = F.binary_cross_entropy_with_logits(pred, torch.zeros_like(pred), reduction="none")
    c = torch.einsum("nc,mc->nm", pos, gt) + torch.einsum("nc,mc->nm", neg, 1 - gt)
    return c / n * weight


def dice_cost(pred, gt, weight, eps=1.0):
    """DiceCost(pred_act=True, naive_dice=True), mask_hungarian_assigner.py:220-253."""
    pred = pred.sigmoid().flatten(1)
    gt = gt.flatten(1).float()
    num = 2 * torch.einsum("nc,mc->nm", pred, gt)
    den = pred.sum(-1)[:, None] + gt.sum(-1)[None, :]
    return (1 - (num + eps) / (den + eps)) * weight


def hungarian(cost):
    r, c = linear_sum_assignment(cost.detach().cpu())
    return torch.from_numpy(r), torch.from_numpy(c)


def mask_assign(cls_pred, mask_points_pred, gt_labels, gt_points_masks, w_cls=2.0, w_mask=5.0, w_dice=5.0):
    """MaskHungarianAssigner.assign -> gt_inds [N] (0 = background, k = 1-based gt index)."""
    num_gt, num_query = gt_labels.shape[0], mask_points_pred.shape[0]
    gt_inds = mask_points_pred.new_full((num_query,), -1, dtype=torch.long)
    if num_gt == 0 or num_query == 0:
        if num_gt == 0:
            gt_inds[:] = 0
        return gt_inds
    cost = (classification_cost(cls_pred, gt_labels, w_cls) + mask_bce_cost(mask_points_pred, gt_points_masks, w_mask)
            + dice_cost(mask_points_pred, gt_points_masks, w_dice))
    r, c = hungarian(cost)
    gt_inds[:] = 0
    gt_inds[r] = c + 1
    return gt_inds


def id_assign(sub_score, obj_score, rel_score, gt_sub_cls, gt_obj_cls, gt_rel_labels, w_sub=1.0, w_obj=1.0, w_rel=0.0):
    """IdMatcher.assign (matcher.py:207-274) -> gt_inds [R]."""
    num_gts, num_q = gt_rel_labels.shape[0], rel_score.shape[0]
    gt_inds = rel_score.new_full((num_q,), -1, dtype=torch.long)
    if num_gts == 0 or num_q == 0:
        if num_gts == 0:
            gt_inds[:] = 0
        return gt_inds
    cost = (classification_cost(sub_score, gt_sub_cls, w_sub) + classification_cost(obj_score, gt_obj_cls, w_obj)
            + classification_cost(rel_score, gt_rel_labels, w_rel))
    r, c = hungarian(cost)
    gt_inds[:] = 0
    gt_inds[r] = c + 1
    return gt_inds


def pseudo_sample(gt_inds):
    """MaskPseudoSampler.sample: positives and the gt each is assigned to."""
    pos = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze(-1).unique()
    return pos, gt_inds[pos] - 1


def get_target_single(subject_score, object_score, cls_score, mask_pred, r_cls_score, gt_rels, gt_labels, gt_masks,
                      num_obj_query, num_rel_query, num_points=12544):
    """pairnet_head.py:613-718."""
    num_gts = gt_labels.shape[0]
    point_coords = torch.rand((1, num_points, 2), device=cls_score.device)
    mask_points_pred = point_sample(mask_pred.unsqueeze(1), point_coords.repeat(num_obj_query, 1, 1)).squeeze(1)
    gt_points_masks = point_sample(gt_masks.unsqueeze(1).float(), point_coords.repeat(num_gts, 1, 1)).squeeze(1)
    od_gt_inds = mask_assign(cls_score, mask_points_pred, gt_labels, gt_points_masks)
    od_pos_inds, od_pos_gt = pseudo_sample(od_gt_inds)

    gt_label_assigned_query = torch.ones_like(gt_labels)
    gt_label_assigned_query[od_pos_gt] = od_pos_inds
    gt_rels = gt_rels.T.long()
    gt_rel_labels = gt_rels[2] - 1
    gt_sub_cls = gt_labels[gt_rels[0]]
    gt_obj_cls = gt_labels[gt_rels[1]]
    gt_sub_pos = gt_label_assigned_query[gt_rels[0]]
    gt_obj_pos = gt_label_assigned_query[gt_rels[1]]
    gt_importance = torch.zeros((num_obj_query, num_obj_query), device=gt_labels.device)
    gt_importance[gt_sub_pos[:], gt_obj_pos[:]] += 1

    tri_gt_inds = id_assign(subject_score, object_score, r_cls_score, gt_sub_cls, gt_obj_cls, gt_rel_labels)
    pos_inds, pos_gt = pseudo_sample(tri_gt_inds)
    gt_subject_ids = torch.full((num_rel_query,), -1, dtype=torch.long, device=gt_labels.device)
    gt_subject_ids[pos_inds] = gt_sub_cls[pos_gt]
    gt_object_ids = torch.full((num_rel_query,), -1, dtype=torch.long, device=gt_labels.device)
    gt_object_ids[pos_inds] = gt_obj_cls[pos_gt]
    r_labels = torch.full((num_rel_query,), -1, dtype=torch.long, device=gt_labels.device)
    r_labels[pos_inds] = gt_rel_labels[pos_gt]
    r_label_weights = gt_labels.new_zeros(num_rel_query)
    r_label_weights[pos_inds] = 1.0
    return r_labels, r_label_weights, gt_subject_ids, gt_object_ids, gt_importance


class OSeesawLoss:
    """mmdet 2.25.1 SeesawLoss (p = 0.8, q = 2.0, eps = 1e-2, mean reduction) with its running ``cum_samples``."""

    def __init__(self, num_classes=56, p=0.8, q=2.0, eps=1e-2, loss_weight=2.0):
        self.num_classes, self.p, self.q, self.eps, self.loss_weight = num_classes, p, q, eps, loss_weight
        self.cum_samples = torch.zeros(num_classes + 1)

    def _ce(self, cls_score, labels, label_weights, cum_samples):
        onehot = F.one_hot(labels, self.num_classes)
        w = cls_score.new_ones(onehot.size())
        if self.p > 0:
            ratio = cum_samples[None, :].clamp(min=1) / cum_samples[:, None].clamp(min=1)
            idx = (ratio < 1.0).float()
            sw = ratio.pow(self.p) * idx + (1 - idx)
            w = w * sw[labels.long(), :]
        if self.q > 0:
            scores = F.softmax(cls_score.detach(), dim=1)
            self_scores = scores[torch.arange(0, len(scores)).to(scores.device).long(), labels.long()]
            sm = scores / self_scores[:, None].clamp(min=self.eps)
            idx = (sm > 1.0).float()
            w = w * (sm.pow(self.q) * idx + (1 - idx))
        cls_score = cls_score + (w.log() * (1 - onehot))
        loss = F.cross_entropy(cls_score, labels, weight=None, reduction="none")
        return (loss * label_weights.float()).mean()

    def __call__(self, cls_score, labels):
        assert cls_score.size(-1) == self.num_classes + 2
        pos = labels < self.num_classes
        obj_labels = (labels == self.num_classes).long()
        self.cum_samples = self.cum_samples.to(labels.device)
        for u in labels.unique():
            self.cum_samples[u] += (labels == u.item()).sum()
        lw = labels.new_ones(labels.size(), dtype=torch.float)
        cls_c, cls_o = cls_score[..., :-2], cls_score[..., -2:]
        if pos.sum() > 0:
            loss_c = self.loss_weight * self._ce(cls_c[pos], labels[pos], lw[pos], self.cum_samples[:self.num_classes])
        else:
            loss_c = cls_c[pos].sum()
        loss_o = self.loss_weight * (F.cross_entropy(cls_o, obj_labels, reduction="none") * lw).mean()
        return dict(loss_cls_objectness=loss_o, loss_cls_classes=loss_c)


def loss(all_cls_scores, all_mask_preds, gt_rels_list, gt_labels_list, gt_masks_list, seesaw=None, num_points=12544,
         w_subobj=4.0, w_match=5.0, num_object_classes=133, return_targets=False):
    """pairnet_head.py:419-564 -> dict(loss_r_cls, loss_sub_cls, loss_obj_cls, loss_match)."""
    od_cls, mask_preds = all_cls_scores["cls"], all_mask_preds["mask"]
    importance, r_cls = all_cls_scores["importance"], all_cls_scores["rel"]
    sub_preds, obj_preds = all_cls_scores["sub"], all_cls_scores["obj"]
    B, N, R = od_cls.size(0), od_cls.size(1), r_cls.size(1)
    num_relations = r_cls.size(-1)
    seesaw = seesaw or OSeesawLoss(num_relations)
    tg = [get_target_single(sub_preds[i], obj_preds[i], od_cls[i], mask_preds[i], r_cls[i], gt_rels_list[i],
                            gt_labels_list[i], gt_masks_list[i], N, R, num_points) for i in range(B)]
    r_labels_list, r_w_list, gt_sub_list, gt_obj_list, gt_imp_list = map(list, zip(*tg))
    r_label_weights = torch.cat(r_w_list, 0)
    m = r_label_weights > 0
    cw = od_cls.new_tensor([1.0] * (num_object_classes + 1))
    gt_object_ids = torch.cat(gt_obj_list, 0)
    loss_obj = w_subobj * F.cross_entropy(obj_preds.flatten(0, 1)[m], gt_object_ids[m], weight=cw, reduction="none").mean()
    gt_subject_ids = torch.cat(gt_sub_list, 0)
    loss_sub = w_subobj * F.cross_entropy(sub_preds.flatten(0, 1)[m], gt_subject_ids[m], weight=cw, reduction="none").mean()
    r_labels = torch.cat(r_labels_list, 0)
    r_scores = r_cls.reshape(-1, num_relations)
    dummy = torch.zeros((int(m.sum()), 2)).to(r_scores.device)
    loss_r = seesaw(torch.cat([r_scores[m], dummy], dim=1), r_labels[m])["loss_cls_classes"]
    gt_importance = torch.stack(gt_imp_list, 0)
    pos_weight = torch.numel(gt_importance) / (gt_importance > 0).sum()
    loss_match = w_match * F.binary_cross_entropy_with_logits(importance, gt_importance, pos_weight=pos_weight,
                                                              reduction="mean")
    out = dict(loss_r_cls=loss_r, loss_sub_cls=loss_sub, loss_obj_cls=loss_obj, loss_match=loss_match)
    if return_targets:
        return out, dict(r_labels=r_labels, r_label_weights=r_label_weights, gt_subject_ids=gt_subject_ids,
                         gt_object_ids=gt_object_ids, gt_importance=gt_importance)
    return out


def synthetic_gt(B, hw, seed, num_gt=12, num_rel=10, num_object_classes=133, num_relations=56):
    """per image 12 random-rectangle masks on the mask-prediction grid, labels U{0..132},
    10 triplets [sub, obj, predicate] with sub != obj, predicate U{1..56}."""
    g = torch.Generator().manual_seed(seed)
    H, W = hw
    rels, labels, masks = [], [], []
    for _ in range(B):
        m = torch.zeros(num_gt, H, W, dtype=torch.bool)
        for k in range(num_gt):
            h = int(torch.randint(max(2, H // 8), max(3, H // 2), (1,), generator=g))
            w = int(torch.randint(max(2, W // 8), max(3, W // 2), (1,), generator=g))
            y0 = int(torch.randint(0, H - h + 1, (1,), generator=g))
            x0 = int(torch.randint(0, W - w + 1, (1,), generator=g))
            m[k, y0:y0 + h, x0:x0 + w] = True
        masks.append(m)
        labels.append(torch.randint(0, num_object_classes, (num_gt,), generator=g))
        s = torch.randint(0, num_gt, (num_rel,), generator=g)
        o = (s + torch.randint(1, num_gt, (num_rel,), generator=g)) % num_gt
        p = torch.randint(1, num_relations + 1, (num_rel,), generator=g)
        rels.append(torch.stack([s, o, p], 1))
    return rels, labels, masks
