"""Training targets and losses of ``CrossHead2`` on the device.

Mirrors the reference interface -- ``loss`` / ``loss_single`` / ``get_targets`` / ``_get_target_single``
(``pairnet/models/relation_heads/pairnet_head.py:419-718``), the assigners the config names
(``configs/mask2former/pairnet.py:190-207``): ``MaskHungarianAssigner`` + ``CrossEntropyLossCost`` + ``DiceCost`` +
``MaskPseudoSampler`` (``panoptic_heads/mask_hungarian_assigner.py:19-351``), ``IdMatcher``
(``relation_heads/approaches/matcher.py:207-274``), mmdet ``ClassificationCost``.

Everything runs on the tensors' device.  The two Hungarian assignments per image are solved on the host with scipy
exactly as the reference does (cost matrices 100 x ~12 and 100 x ~10 -> one small D2H each); the 12 544-point bilinear
sampling, the three cost matrices and the losses are batched device ops.  Not a hot kernel: the
step time is in the forward / backward of the head."""
import torch
import torch.nn.functional as F

from .registry import BBOX_ASSIGNERS, BBOX_SAMPLERS, MATCH_COST, build_assigner, build_sampler

try:
    from scipy.optimize import linear_sum_assignment
except ImportError:  # pragma: no cover
    linear_sum_assignment = None


def point_sample(inp, points, align_corners=False, **kwargs):
    """mmcv.ops.point_sample: inp [N,C,H,W], points [N,P,2] (x, y) in [0,1]^2 -> [N,C,P]."""
    add_dim = points.dim() == 3
    if add_dim:
        points = points.unsqueeze(2)
    out = F.grid_sample(inp, 2.0 * points - 1.0, align_corners=align_corners, **kwargs)
    return out.squeeze(3) if add_dim else out


class AssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels


class SamplingResult:
    """``MaskSamplingResult`` fields the head reads (mask_hungarian_assigner.py:261-285)."""

    def __init__(self, pos_inds, neg_inds, assign_result):
        self.pos_inds, self.neg_inds = pos_inds, neg_inds
        self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1
        self.pos_gt_labels = assign_result.labels[pos_inds] if assign_result.labels is not None else None


@MATCH_COST.register_module()
class ClassificationCost:
    def __init__(self, weight=1.0):
        self.weight = weight

    def __call__(self, cls_pred, gt_labels):
        return -cls_pred.softmax(-1)[:, gt_labels] * self.weight


@MATCH_COST.register_module()
class CrossEntropyLossCost:
    """mask_hungarian_assigner.py:136-196."""

    def __init__(self, weight=1.0, use_sigmoid=True):
        assert use_sigmoid, "use_sigmoid = False is not supported yet."
        self.weight, self.use_sigmoid = weight, use_sigmoid

    def __call__(self, cls_pred, gt_labels):
        cls_pred = cls_pred.flatten(1).float()
        gt_labels = gt_labels.flatten(1).float()
        n = cls_pred.shape[1]
        pos = F.binary_cross_entropy_with_logits(cls_pred, torch.ones_like(cls_pred), reduction="none")
        neg = F.binary_cross_entropy_with_logits(cls_pred, torch.zeros_like(cls_pred), reduction="none")
        cost = torch.einsum("nc,mc->nm", pos, gt_labels) + torch.einsum("nc,mc->nm", neg, 1 - gt_labels)
        return cost / n * self.weight


@MATCH_COST.register_module()
class DiceCost:
    """mask_hungarian_assigner.py:199-253."""

    def __init__(self, weight=1.0, pred_act=False, eps=1e-3, naive_dice=True):
        self.weight, self.pred_act, self.eps, self.naive_dice = weight, pred_act, eps, naive_dice

    def __call__(self, mask_preds, gt_masks):
        if self.pred_act:
            mask_preds = mask_preds.sigmoid()
        mask_preds = mask_preds.flatten(1)
        gt_masks = gt_masks.flatten(1).float()
        numerator = 2 * torch.einsum("nc,mc->nm", mask_preds, gt_masks)
        if self.naive_dice:
            denominator = mask_preds.sum(-1)[:, None] + gt_masks.sum(-1)[None, :]
        else:
            denominator = mask_preds.pow(2).sum(1)[:, None] + gt_masks.pow(2).sum(1)[None, :]
        return (1 - (numerator + self.eps) / (denominator + self.eps)) * self.weight


def _hungarian(cost, device):
    if linear_sum_assignment is None:
        raise ImportError('Please run "pip install scipy" to install scipy first.')
    r, c = linear_sum_assignment(cost.detach().cpu())
    return torch.from_numpy(r).to(device), torch.from_numpy(c).to(device)


@BBOX_ASSIGNERS.register_module()
class MaskHungarianAssigner:
    """mask_hungarian_assigner.py:19-132."""

    def __init__(self, cls_cost=dict(type="ClassificationCost", weight=1.0),
                 mask_cost=dict(type="FocalLossCost", weight=1.0, binary_input=True),
                 dice_cost=dict(type="DiceCost", weight=1.0)):
        self.cls_cost, self.mask_cost, self.dice_cost = (MATCH_COST.build(cls_cost), MATCH_COST.build(mask_cost),
                                                         MATCH_COST.build(dice_cost))

    def assign(self, cls_pred, mask_pred, gt_labels, gt_mask, img_meta, gt_bboxes_ignore=None, eps=1e-7):
        assert gt_bboxes_ignore is None, "Only case when gt_bboxes_ignore is None is supported."
        num_gt, num_query = gt_labels.shape[0], mask_pred.shape[0]
        assigned_gt_inds = mask_pred.new_full((num_query,), -1, dtype=torch.long)
        assigned_labels = mask_pred.new_full((num_query,), -1, dtype=torch.long)
        if num_gt == 0 or num_query == 0:
            if num_gt == 0:
                assigned_gt_inds[:] = 0
            return AssignResult(num_gt, assigned_gt_inds, None, labels=assigned_labels)
        cls_cost = self.cls_cost(cls_pred, gt_labels) if (self.cls_cost.weight != 0 and cls_pred is not None) else 0
        mask_cost = self.mask_cost(mask_pred, gt_mask) if self.mask_cost.weight != 0 else 0
        dice_cost = self.dice_cost(mask_pred, gt_mask) if self.dice_cost.weight != 0 else 0
        rows, cols = _hungarian(cls_cost + mask_cost + dice_cost, mask_pred.device)
        assigned_gt_inds[:] = 0
        assigned_gt_inds[rows] = cols + 1
        assigned_labels[rows] = gt_labels[cols]
        return AssignResult(num_gt, assigned_gt_inds, None, labels=assigned_labels)


@BBOX_ASSIGNERS.register_module()
class IdMatcher:
    """relation_heads/approaches/matcher.py:207-274."""

    def __init__(self, sub_id_cost=dict(type="ClassificationCost", weight=1.0),
                 obj_id_cost=dict(type="ClassificationCost", weight=1.0),
                 r_cls_cost=dict(type="ClassificationCost", weight=1.0)):
        self.sub_id_cost, self.obj_id_cost, self.r_cls_cost = (MATCH_COST.build(sub_id_cost), MATCH_COST.build(obj_id_cost),
                                                               MATCH_COST.build(r_cls_cost))

    def assign(self, sub_score, obj_score, rel_cls_score, gt_sub_cls, gt_obj_cls, gt_rel_labels, img_meta,
               gt_bboxes_ignore=None, eps=1e-7):
        assert gt_bboxes_ignore is None, "Only case when gt_bboxes_ignore is None is supported."
        num_gts, num_bboxes = gt_rel_labels.shape[0], rel_cls_score.shape[0]
        assigned_gt_inds = rel_cls_score.new_full((num_bboxes,), -1, dtype=torch.long)
        assigned_s_labels = rel_cls_score.new_full((num_bboxes,), -1, dtype=torch.long)
        if num_gts == 0 or num_bboxes == 0:
            if num_gts == 0:
                assigned_gt_inds[:] = 0
            return AssignResult(num_gts, assigned_gt_inds, None, labels=assigned_s_labels)
        cost = (self.sub_id_cost(sub_score, gt_sub_cls) + self.obj_id_cost(obj_score, gt_obj_cls)
                + self.r_cls_cost(rel_cls_score, gt_rel_labels))
        rows, cols = _hungarian(cost, rel_cls_score.device)
        assigned_gt_inds[:] = 0
        assigned_gt_inds[rows] = cols + 1
        assigned_s_labels[rows] = gt_sub_cls[cols]
        return AssignResult(num_gts, assigned_gt_inds, None, labels=assigned_s_labels)


@BBOX_SAMPLERS.register_module()
class MaskPseudoSampler:
    """mask_hungarian_assigner.py:313-351."""

    def __init__(self, **kwargs):
        pass

    def sample(self, assign_result, masks, gt_masks, **kwargs):
        pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).squeeze(-1).unique()
        neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).squeeze(-1).unique()
        return SamplingResult(pos_inds, neg_inds, assign_result)


def multi_apply(func, *args, **kwargs):
    """mmdet.core.multi_apply."""
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


class TrainMixin:
    """``loss`` / ``get_targets`` of the reference head; mixed into ``CrossHead2`` (needs ``num_queries``,
    ``num_obj_query``, ``num_rel_query``, ``num_relations`` and the three built losses)."""

    def _init_train_cfg(self, train_cfg):
        """pairnet_head.py:129-137."""
        if train_cfg:
            self.mask_assigner = build_assigner(train_cfg["mask_assigner"])
            self.sampler = build_sampler(train_cfg["sampler"], context=self)
            self.num_points = train_cfg.get("num_points", 12544)
            self.oversample_ratio = train_cfg.get("oversample_ratio", 3.0)
            self.importance_sample_ratio = train_cfg.get("importance_sample_ratio", 0.75)
            self.id_assigner = build_assigner(train_cfg["id_assigner"])

    def loss(self, all_cls_scores, all_mask_preds, gt_rels_list, gt_bboxes_list, gt_labels_list, gt_masks_list, img_metas,
             gt_bboxes_ignore=None):
        """pairnet_head.py:419-480."""
        assert gt_bboxes_ignore is None, "Only supports for gt_bboxes_ignore setting to None."
        r, s, o, m = multi_apply(self.loss_single, [all_cls_scores["sub"]], [all_cls_scores["obj"]],
                                 [all_cls_scores["importance"]], [all_cls_scores["cls"]], [all_mask_preds["mask"]],
                                 [all_cls_scores["rel"]], [gt_rels_list], [gt_labels_list], [gt_masks_list], [img_metas],
                                 [gt_bboxes_ignore])
        return dict(loss_r_cls=r[-1], loss_sub_cls=s[-1], loss_obj_cls=o[-1], loss_match=m[-1])

    def loss_single(self, sub_cls_preds, obj_cls_preds, importance, od_cls_scores, mask_preds, r_cls_scores, gt_rels_list,
                    gt_labels_list, gt_masks_list, img_metas, gt_bboxes_ignore_list=None):
        """pairnet_head.py:482-564."""
        num_imgs = od_cls_scores.size(0)
        (r_labels_list, r_label_weights_list, gt_subject_id_list, gt_object_id_list, gt_importance_list) = self.get_targets(
            [sub_cls_preds[i] for i in range(num_imgs)], [obj_cls_preds[i] for i in range(num_imgs)],
            [od_cls_scores[i] for i in range(num_imgs)], [mask_preds[i] for i in range(num_imgs)],
            [r_cls_scores[i] for i in range(num_imgs)], gt_rels_list, gt_labels_list, gt_masks_list, img_metas,
            gt_bboxes_ignore_list)
        r_label_weights = torch.cat(r_label_weights_list, 0)
        mask = r_label_weights > 0
        gt_object_ids = torch.cat(gt_object_id_list, 0)
        loss_obj_cls = self.subobj_cls_loss(obj_cls_preds.flatten(0, 1)[mask], gt_object_ids[mask])
        gt_subject_ids = torch.cat(gt_subject_id_list, 0)
        loss_sub_cls = self.subobj_cls_loss(sub_cls_preds.flatten(0, 1)[mask], gt_subject_ids[mask])
        r_labels = torch.cat(r_labels_list, 0)
        r_cls_scores = r_cls_scores.reshape(-1, self.num_relations)
        dummy_objectness = torch.zeros((int(mask.sum()), 2)).to(r_cls_scores.device)   # for seesaw loss
        r_loss_cls = self.rel_cls_loss(torch.cat([r_cls_scores[mask], dummy_objectness], dim=1),
                                       r_labels[mask])["loss_cls_classes"]
        gt_importance = torch.stack(gt_importance_list, 0)
        pos_weight = torch.numel(gt_importance) / (gt_importance > 0).sum()   # per rank, over the local batch
        loss_match = self.importance_match_loss(importance, gt_importance, pos_weight)
        return r_loss_cls, loss_sub_cls, loss_obj_cls, loss_match

    def get_targets(self, subject_scores_list, object_scores_list, cls_scores_list, mask_preds_list, r_cls_scores_list,
                    gt_rels_list, gt_labels_list, gt_masks_list, img_metas, gt_bboxes_ignore_list=None):
        """pairnet_head.py:566-611."""
        assert gt_bboxes_ignore_list is None, "Only supports for gt_bboxes_ignore setting to None."
        n = len(r_cls_scores_list)
        return multi_apply(self._get_target_single, subject_scores_list, object_scores_list, cls_scores_list,
                           mask_preds_list, r_cls_scores_list, gt_rels_list, gt_labels_list, gt_masks_list, img_metas,
                           [None] * n)

    def _get_target_single(self, subject_score, object_score, cls_score, mask_pred, r_cls_score, gt_rels, gt_labels,
                           gt_masks, img_metas, gt_bboxes_ignore=None):
        """pairnet_head.py:613-718."""
        num_gts = gt_labels.shape[0]
        point_coords = torch.rand((1, self.num_points, 2), device=cls_score.device)
        mask_points_pred = point_sample(mask_pred.unsqueeze(1), point_coords.repeat(self.num_queries, 1, 1)).squeeze(1)
        gt_points_masks = point_sample(gt_masks.unsqueeze(1).float(), point_coords.repeat(num_gts, 1, 1)).squeeze(1)
        assign_result = self.mask_assigner.assign(cls_score, mask_points_pred, gt_labels, gt_points_masks, img_metas)
        sampling_result = self.sampler.sample(assign_result, mask_pred, gt_masks)
        od_pos_inds = sampling_result.pos_inds
        # scene graph: gt object -> matched object query (unmatched ones keep query 1, as in the reference)
        gt_label_assigned_query = torch.ones_like(gt_labels)
        gt_label_assigned_query[sampling_result.pos_assigned_gt_inds] = od_pos_inds
        gt_rels = gt_rels.T.long()
        gt_rel_labels = gt_rels[2] - 1
        gt_sub_cls, gt_obj_cls = gt_labels[gt_rels[0]], gt_labels[gt_rels[1]]
        gt_sub_pos, gt_obj_pos = gt_label_assigned_query[gt_rels[0]], gt_label_assigned_query[gt_rels[1]]
        gt_importance = torch.zeros((self.num_obj_query, self.num_obj_query), device=gt_labels.device)
        gt_importance[gt_sub_pos[:], gt_obj_pos[:]] += 1
        triplet_assign_result = self.id_assigner.assign(subject_score, object_score, r_cls_score, gt_sub_cls, gt_obj_cls,
                                                        gt_rel_labels, img_metas, gt_bboxes_ignore)
        triplet_sampling_result = self.sampler.sample(triplet_assign_result, torch.ones_like(subject_score),
                                                      torch.ones_like(subject_score))
        pos_inds = triplet_sampling_result.pos_inds
        pos_gt = triplet_sampling_result.pos_assigned_gt_inds
        full = lambda: torch.full((self.num_rel_query,), -1, dtype=torch.long, device=gt_labels.device)
        gt_subject_ids, gt_object_ids, r_labels = full(), full(), full()
        gt_subject_ids[pos_inds] = gt_sub_cls[pos_gt]
        gt_object_ids[pos_inds] = gt_obj_cls[pos_gt]
        r_labels[pos_inds] = gt_rel_labels[pos_gt]
        r_label_weights = gt_labels.new_zeros(self.num_rel_query)
        r_label_weights[pos_inds] = 1.0
        return r_labels, r_label_weights, gt_subject_ids, gt_object_ids, gt_importance
