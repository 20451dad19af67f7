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


_PINNED_MIN_BYTES = 1 << 20


def _np(t):
    """Tensor -> numpy.  Large CUDA tensors (the [2K,H,W] boolean masks: 213 MB per 800x1333 image) are copied into a
    PINNED host tensor from PyTorch's caching host allocator -- a pageable ``.cpu()`` runs at a fraction of the PCIe rate
    (measured: 218 ms per bs = 2 step, almost all of it this copy).  The numpy array owns a reference to that tensor;
    its block returns to the allocator's cache when the caller drops the result."""
    if not isinstance(t, torch.Tensor):
        return t
    t = t.detach()
    if t.is_cuda and t.numel() * t.element_size() >= _PINNED_MIN_BYTES:
        try:
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        except RuntimeError:      # locked-memory limit of the host: the pageable copy is only slower
            return t.cpu().numpy()
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host.numpy()
    return t.cpu().numpy()


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
