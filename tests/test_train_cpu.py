"""Training targets + losses on the CPU: the oracle restatement against the goldens minted from the
REFERENCE's own ``CrossHead2.loss`` (oracle/pin_train.py), the product's ``loss`` / assigners against the oracle, and the
differentiable head (torch_head.py) against the oracle forward + its static set of trainable parameters."""
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN, oracle_small_head, product_small_head


def _case(tag):
    from oracle.pin_train import TRAIN_CASES, case_inputs
    t = [c for c in TRAIN_CASES if c[0] == tag][0]
    return t, case_inputs(*t[1:])


@pytest.mark.parametrize("tag", ["b2_32x48", "b3_24x40"])
def test_oracle_loss_matches_reference_golden_and_product_matches_oracle(tag):
    from oracle import train as ot
    (_, B, hw4, seed, gt_seed, scale), (cls, msk, rels, labels, masks) = _case(tag)
    g = np.load(os.path.join(GOLDEN, f"train_ref_{tag}.npz"))
    torch.manual_seed(gt_seed)
    with torch.no_grad():
        got, tg = ot.loss(cls, msk, rels, labels, masks, return_targets=True)
    for k in ("loss_r_cls", "loss_sub_cls", "loss_obj_cls", "loss_match"):
        assert abs(float(got[k]) - float(g[k])) <= 2e-6 * abs(float(g[k])), k   # reference's own loss()
    for k in ("r_labels", "r_label_weights", "gt_subject_ids", "gt_object_ids", "gt_importance"):
        assert np.array_equal(tg[k].numpy(), g[k]), k                            # targets: exact
    # the product's loss / get_targets / assigners on the same tensors and RNG stream
    head = product_small_head(oracle_small_head(), device="cpu")
    assert head.train_cfg is not None and head.num_points == 12544
    head.rel_cls_loss.cum_samples.zero_()
    torch.manual_seed(gt_seed)
    with torch.no_grad():
        mine = head.loss(cls, msk, rels, None, labels, masks, [dict()] * B)
    for k in got:
        assert torch.equal(mine[k], got[k]), (k, float(mine[k]), float(got[k]))
    # the Seesaw running counts moved by exactly the matched triplets' labels
    assert float(head.rel_cls_loss.cum_samples.sum()) == float((tg["r_label_weights"] > 0).sum())


def test_importance_pos_weight_and_unmatched_default_query():
    """pos_weight = numel / #positives over the local batch (pairnet_head.py:553-554); GT objects the mask matcher left
    unmatched keep query index 1 (``ones_like``, :649)."""
    from pairnet_b200.training import MaskHungarianAssigner
    head = product_small_head(oracle_small_head(), device="cpu")
    g = torch.Generator().manual_seed(3)
    N, R = head.num_obj_query, head.num_rel_query
    cls = torch.randn(N, 134, generator=g)
    mask_pred = torch.randn(N, 16, 16, generator=g)
    gt_masks = torch.zeros(3, 32, 32, dtype=torch.bool)
    gt_masks[0, :8, :8] = gt_masks[1, 8:, 8:] = gt_masks[2, 16:, :16] = True
    gt_labels = torch.tensor([5, 9, 77])
    rels = torch.tensor([[0, 1, 3], [2, 0, 10]])
    torch.manual_seed(0)
    out = head._get_target_single(torch.randn(R, 134, generator=g), torch.randn(R, 134, generator=g), cls, mask_pred,
                                  torch.randn(R, 56, generator=g), rels, gt_labels, gt_masks, dict())
    r_labels, r_w, sid, oid, imp = out
    assert int(imp.sum()) == 2 and imp.shape == (N, N)
    assert int((r_w > 0).sum()) == 2 and set(r_labels[r_w > 0].tolist()) == {2, 9}
    assert set(sid[r_w > 0].tolist()) == {5, 77} and set(oid[r_w > 0].tolist()) == {9, 5}
    assert isinstance(head.mask_assigner, MaskHungarianAssigner)


def test_differentiable_head_matches_oracle_and_trainable_set_is_static():
    from oracle.make_golden import small_head_inputs
    from pairnet_b200 import torch_head as th
    o = oracle_small_head()
    head = product_small_head(o, device="cpu")
    mf, mems = small_head_inputs(1, (16, 24), 31)
    with torch.no_grad():
        ocls, omsk = o.forward_from_memories(mf, mems)
    q, cls_pred, mask_pred = th.masked_decoder(head, mf, mems)
    cls_scores, mask_preds, _ = th.relation_side(head, q, cls_pred, mask_pred)
    for k in ("cls", "rel", "importance", "sub", "obj"):
        assert torch.allclose(cls_scores[k], ocls[k], rtol=0, atol=2e-5 * float(ocls[k].abs().max())), k
    assert torch.allclose(mask_preds["mask"], omsk["mask"], rtol=0, atol=2e-5 * float(omsk["mask"].abs().max()))
    loss = cls_scores["rel"].square().mean() + cls_scores["importance"].square().mean()
    loss.backward()
    want = {n for n, _ in th.trainable_parameters(head, "head")}
    have = {n for n, p in head.named_parameters() if p.grad is not None and bool(p.grad.abs().sum() > 0)}
    assert want == have, (sorted(want - have)[:5], sorted(have - want)[:5])
    rel_only = {n for n, _ in th.trainable_parameters(head, "relation")}
    assert rel_only < want and all(not n.startswith("transformer_decoder") for n in rel_only)
