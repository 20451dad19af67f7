"""CPU restatement of the reference's inference post-processing (TEST INFRASTRUCTURE).

Follows ``CrossHead2.get_bboxes`` / ``_get_bboxes_single`` (``pairnet/models/relation_heads/pairnet_head.py:759-924``) and
``triplet2Result`` (``pairnet/models/frameworks/psgtr.py:15-51``) op for op, in torch on the CPU.  Pinned to the reference's
own code: ``python -m oracle.postproc`` EXECUTES the reference's ``get_bboxes`` from ``/root/reference`` (through the
constructor shims of ``oracle.pin_reference``) on the synthetic head outputs below, requires bit-equal results and mints
``tests/golden/postproc_ref_*.npz``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.
"""
import os
import sys
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

INSTANCE_OFFSET = 1000  # mmdet.datasets.coco_panoptic.INSTANCE_OFFSET
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

POST_CASES = [  # (tag, B, N, K, (h4, w4), img_shape (H, W), scale_factor, seed[, confident])
    ("b2_24x40", 2, 100, 100, (24, 40), (96, 160), (1.0, 1.0), 31),
    ("b1_32x48_scaled", 1, 100, 100, (32, 48), (128, 192), (1.6, 1.6), 32),
    ("b1_20x28_n40", 1, 40, 24, (20, 28), (80, 111), (1.0, 1.0), 33),
    ("b1_16x24_nothing_kept", 1, 40, 24, (16, 24), (64, 96), (1.0, 1.0), 34, False),  # no query passes score > 0.5
]


def get_bboxes_single(all_masks, all_cls_score, s_cls_score, o_cls_score, r_cls_score, s_mask_pred, o_mask_pred,
                      img_shape, scale_factor, num_relations, num_rel_query):
    """pairnet_head.py:788-924 (``rescale`` is unused by the reference)."""
    mask_size = (round(img_shape[0] / scale_factor[1]), round(img_shape[1] / scale_factor[0]))  # :803-806
    s_logits = F.softmax(s_cls_score, dim=-1)[..., :-1]                                         # :811-812
    o_logits = F.softmax(o_cls_score, dim=-1)[..., :-1]
    s_labels = s_logits.argmax(-1) + 1                                                           # :814-815
    o_labels = o_logits.argmax(-1) + 1
    r_dists = F.softmax(r_cls_score, dim=-1).reshape(-1, num_relations)                          # :817-820
    r_dists = torch.cat([torch.zeros(num_rel_query, 1), r_dists], dim=-1)
    complete_labels = torch.cat((s_labels, o_labels), 0)                                         # :822
    all_logits = F.softmax(all_cls_score, dim=-1)[..., :-1]                                      # :823
    all_scores, all_labels = all_logits.max(-1)                                                  # :825
    up = lambda m: F.interpolate(m.unsqueeze(1), size=mask_size, mode="bilinear", align_corners=False).squeeze(1)
    all_masks = up(all_masks)                                                                    # :826-828
    s_mask = torch.sigmoid(up(s_mask_pred)) > 0.5                                                # :830-843
    o_mask = torch.sigmoid(up(o_mask_pred)) > 0.5
    masks = torch.cat((s_mask, o_mask), 0)                                                       # :844
    keep = (all_labels != s_logits.shape[-1] - 1) & (all_scores > 0.5)                           # :846-848
    all_labels, all_masks, all_scores = all_labels[keep], all_masks[keep], all_scores[keep]
    h, w = all_masks.shape[-2:]
    if all_labels.numel() == 0:                                                                  # :854-855
        pan_img = torch.ones(mask_size).to(torch.long)
    else:
        all_masks = all_masks.flatten(1)
        stuff_equiv_classes = defaultdict(list)                                                  # :858-861
        for k, label in enumerate(all_labels):
            if label.item() >= 80:
                stuff_equiv_classes[label.item()].append(k)

        def get_ids_area(all_masks, all_scores, all_labels, dedup=False):                       # :863-891
            m_id = all_masks.transpose(0, 1).softmax(-1)
            if m_id.shape[-1] == 0:
                m_id = torch.zeros((h, w), dtype=torch.long)
            else:
                m_id = m_id.argmax(-1).view(h, w)
            if dedup:
                for equiv in stuff_equiv_classes.values():
                    if len(equiv) > 1:
                        for eq_id in equiv:
                            m_id.masked_fill_(m_id.eq(eq_id), equiv[0])
            seg_img = (m_id * INSTANCE_OFFSET + all_labels[m_id]).view(h, w).to(torch.long)
            area = [int(m_id.eq(i).sum().item()) for i in range(len(all_scores))]
            return area, seg_img

        area, pan_img = get_ids_area(all_masks, all_scores, all_labels, dedup=True)              # :893
        while True:                                                                              # :896-908
            filtered_small = torch.as_tensor([a <= 4 for a in area], dtype=torch.bool)
            if filtered_small.any().item():
                all_scores, all_labels, all_masks = all_scores[~filtered_small], all_labels[~filtered_small], \
                    all_masks[~filtered_small]
                area, pan_img = get_ids_area(all_masks, all_scores, all_labels)
            else:
                break
    det_bboxes = torch.zeros((num_rel_query * 2, 5))                                             # :910-912
    r_scores, r_labels = torch.zeros(num_rel_query), torch.zeros(num_rel_query)                  # :914-915
    rel_pairs = torch.arange(len(det_bboxes), dtype=torch.int).reshape(2, -1).T                  # :916
    return det_bboxes, complete_labels, rel_pairs, masks, pan_img, r_scores, r_labels, r_dists


def get_bboxes(cls_scores, mask_preds, img_metas, num_relations, num_rel_query):
    """pairnet_head.py:759-786."""
    return [get_bboxes_single(mask_preds["mask"][i], cls_scores["cls"][i], cls_scores["sub"][i], cls_scores["obj"][i],
                              cls_scores["rel"][i], mask_preds["sub_seg"][i], mask_preds["obj_seg"][i],
                              img_metas[i]["img_shape"], img_metas[i]["scale_factor"], num_relations, num_rel_query)
            for i in range(len(img_metas))]


# ------------------------------------------------------------------------------------------------ synthetic outputs
def synthetic_head_outputs(B, N, K, hw4, seed, num_classes=133, num_relations=56, confident=True):
    """Head outputs with the structure post-processing branches on: confident thing / stuff / background queries,
    duplicated stuff classes (dedup path), smooth mask logits plus a few needle masks that win <= 4 pixels (small-area
    filter loop).  Deterministic (numpy PCG64)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h, w = hw4
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    cls = rng.standard_normal((B, N, num_classes + 1))
    for b in range(B if confident else 0):
        conf = rng.permutation(N)[: max(6, N // 3)]
        for j, q in enumerate(conf):
            c = [85, 85, 120, 3, 17, num_classes][j] if j < 6 else int(rng.integers(0, num_classes + 1))
            cls[b, q, c] += 14.0 if j != 4 else 5.0   # one mid-confidence query (score < 0.5 -> dropped)
    # smooth fields: low-res noise upsampled
    lo = rng.standard_normal((B, N, (h + 3) // 4, (w + 3) // 4)) * 4.0
    mask = F.interpolate(f32(lo), size=(h, w), mode="bicubic", align_corners=False).numpy().astype(np.float64)
    mask += rng.standard_normal((B, N, h, w)) * 0.3
    for b in range(B if confident else 0):  # needle masks: strongly negative except one pixel
        for q in rng.permutation(N)[:3]:
            mask[b, q] = -30.0
            mask[b, q, int(rng.integers(0, h)), int(rng.integers(0, w))] = 40.0
            cls[b, q, int(rng.integers(0, 80))] += 14.0
    sub_pos = np.stack([rng.integers(0, N, K) for _ in range(B)])
    obj_pos = np.stack([rng.integers(0, N, K) for _ in range(B)])
    cls_t, mask_t = f32(cls), f32(mask)
    sp, op = torch.from_numpy(sub_pos), torch.from_numpy(obj_pos)
    gat = lambda t, idx: torch.stack([t[b, idx[b]] for b in range(B)])
    cls_scores = dict(cls=cls_t, sub=gat(cls_t, sp), obj=gat(cls_t, op), rel=f32(rng.standard_normal((B, K, num_relations)) * 2),
                      importance=f32(rng.standard_normal((B, N, N))))
    mask_preds = dict(mask=mask_t, sub_seg=gat(mask_t, sp), obj_seg=gat(mask_t, op))
    return cls_scores, mask_preds, sp, op


def case_inputs(case):
    tag, B, N, K, hw4, img_shape, sf, seed = case[:8]
    cls_scores, mask_preds, sp, op = synthetic_head_outputs(B, N, K, hw4, seed, confident=case[8] if len(case) > 8 else True)
    H, W = img_shape
    metas = [dict(img_shape=(int(round(H * sf[1])), int(round(W * sf[0])), 3),
                  scale_factor=np.array([sf[0], sf[1], sf[0], sf[1]], dtype=np.float32)) for _ in range(B)]
    return cls_scores, mask_preds, metas, sp, op


def _pack(results):
    out = {}
    for i, (bb, labels, pairs, masks, pan, rs, rl, rd) in enumerate(results):
        out[f"labels{i}"] = labels.numpy()
        out[f"masks{i}"] = np.packbits(masks.numpy(), axis=-1)
        out[f"mask_shape{i}"] = np.array(masks.shape)
        out[f"pan{i}"] = pan.numpy().astype(np.int32)
        out[f"r_dists{i}"] = rd.numpy()
        out[f"rel_pairs{i}"] = pairs.numpy()
    return out


def main(check_only=False):
    from . import pin_reference as pr
    assert os.path.isdir(pr.REF), "the reference tree is only present in the dev container"
    heads = {}
    for case in POST_CASES:
        tag, B, N, K = case[:4]
        if (N, K) not in heads:
            heads[(N, K)] = pr.build_reference_head(N, K)
        ref = heads[(N, K)]
        cls_scores, mask_preds, metas, _, _ = case_inputs(case)
        with torch.no_grad():
            r = ref.get_bboxes(cls_scores, mask_preds, metas)          # the reference's own code
            o = get_bboxes(cls_scores, mask_preds, metas, ref.num_relations, ref.num_rel_query)
        for i, (a, b) in enumerate(zip(r, o)):
            for j, (x, y) in enumerate(zip(a, b)):
                assert x.shape == y.shape and x.dtype == y.dtype and torch.equal(x, y), (tag, i, j)
        kept = [int((p[4] // INSTANCE_OFFSET).unique().numel()) for p in r]
        print(f"pinned {tag}: reference get_bboxes == oracle on all 8 outputs x {B} image(s); segments per image {kept}")
        if not check_only:
            np.savez_compressed(os.path.join(GOLDEN, f"postproc_ref_{tag}.npz"), **_pack(r))
    return 0


if __name__ == "__main__":
    sys.exit(main(check_only="--check" in sys.argv))
