"""Result container and ``triplet2Result`` of the reference detector wrapper
(``pairnet/models/relation_heads/approaches/relation_util.py:20-98``, ``pairnet/models/frameworks/psgtr.py:15-71``):
the object ``PSGTr.simple_test`` hands to the evaluation code.  Host-side glue, same field names."""
import torch


class Result(object):
    """little container class for holding the detection result (same keyword surface as the reference's)."""

    _FIELDS = ("bboxes", "dists", "labels", "masks", "formatted_masks", "points", "rels", "key_rels", "relmaps",
               "refine_bboxes", "formatted_bboxes", "refine_scores", "refine_dists", "refine_labels", "target_labels",
               "rel_scores", "rel_dists", "triplet_scores", "ranking_scores", "rel_pair_idxes", "rel_labels",
               "target_rel_labels", "target_key_rel_labels", "saliency_maps", "attrs", "rel_cap_inputs",
               "rel_cap_targets", "rel_ipts", "tgt_rel_cap_inputs", "tgt_rel_cap_targets", "tgt_rel_ipts",
               "rel_cap_scores", "rel_cap_seqs", "rel_cap_sents", "rel_ipt_scores", "cap_inputs", "cap_targets",
               "cap_scores", "cap_scores_from_triplet", "alphas", "rel_distribution", "obj_distribution",
               "word_obj_distribution", "cap_seqs", "cap_sents", "img_shape", "scenes", "target_scenes", "add_losses",
               "head_spec_losses", "pan_results", "sub_pos", "obj_pos")

    def __init__(self, **kwargs):
        unknown = set(kwargs) - set(self._FIELDS)
        if unknown:
            raise TypeError(f"Result got unexpected fields {sorted(unknown)}")
        for f in self._FIELDS:
            setattr(self, f, kwargs.get(f))

    def is_none(self):
        return all(getattr(self, f) is None for f in self._FIELDS)

    # the reference makes the object iterable / indexable as a 1-element sequence
    def __len__(self):
        return 1

    def __getitem__(self, i):
        return self

    def __iter__(self):
        yield self


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t


def triplet2Result(triplets, use_mask, eval_mask_rels=False):
    """psgtr.py:15-71: the 8-tuple (6-tuple without masks) of ``CrossHead2.get_bboxes`` -> ``Result`` of numpy arrays."""
    if use_mask:
        bboxes, labels, rel_pairs, masks, pan_seg, r_scores, r_labels, r_dists = triplets
        pan_seg = _np(pan_seg)
        return Result(refine_bboxes=_np(bboxes), labels=_np(labels), formatted_masks=dict(pan_results=_np(pan_seg)),
                      rel_pair_idxes=_np(rel_pairs), rel_dists=_np(r_dists), rel_labels=_np(r_labels),
                      pan_results=_np(pan_seg), masks=_np(masks))
    bboxes, labels, rel_pairs, r_scores, r_labels, r_dists = triplets
    return Result(refine_bboxes=_np(bboxes), labels=_np(labels), formatted_masks=dict(pan_results=None),
                  rel_pair_idxes=_np(rel_pairs), rel_dists=_np(r_dists), rel_labels=_np(r_labels), pan_results=None)
