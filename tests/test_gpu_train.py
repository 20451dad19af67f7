"""Training step on the B200: differentiable head vs the CUDA inference path, a few optimizer steps."""
import pytest
import torch

from tests.util import oracle_small_head, product_small_head, rel_err

pytestmark = pytest.mark.gpu


def test_train_outputs_match_the_cuda_forward():
    """Both training scopes produce the tensors the CUDA library produces on the same weights and inputs (scope
    'relation' consumes the library's last-layer queries; scope 'head' re-evaluates the whole head with torch ops)."""
    from oracle.make_golden import small_head_inputs
    from pairnet_b200 import torch_head as th
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    head = product_small_head(oracle_small_head())
    mf, mems = small_head_inputs(2, (32, 48), 21)
    mf, mems = mf.cuda(), [m.cuda() for m in mems]
    taps = {}
    cls, msk = head.forward_from_memories(mf, mems, taps=taps)
    q = taps["query_out"].transpose(0, 1).contiguous()
    with torch.no_grad():
        c1, m1, _ = th.relation_side(head, q, cls["cls"], msk["mask"])
        qf, cp, mp = th.masked_decoder(head, mf, mems)
        c2, m2, _ = th.relation_side(head, qf, cp, mp)
    for got in (c1, c2):
        assert rel_err(got["importance"], cls["importance"]) < 2e-5
        assert rel_err(got["rel"], cls["rel"]) < 1e-3       # pair order can differ at near-ties of the top-k
    assert rel_err(qf.transpose(0, 1), taps["query_out"]) < 1e-4
    assert rel_err(c2["cls"], cls["cls"]) < 1e-4 and rel_err(m2["mask"], msk["mask"]) < 1e-4


@pytest.mark.parametrize("scope,amp", [("relation", None), ("head", None), ("head", torch.bfloat16)])
def test_training_steps_update_exactly_the_trainable_set(scope, amp):
    from tests.util import ROOT
    import os
    from pairnet_b200.registry import Config, build_detector
    from pairnet_b200.trainer import TrainStep, synthetic_targets
    from pairnet_b200 import torch_head as th
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "pairnet_r50_b200.py"))
    torch.manual_seed(10086)
    model = build_detector(cfg.model)
    model.init_weights()
    model = model.cuda()
    H, W = 256, 320
    imgs = torch.randn(2, 3, H, W, generator=torch.Generator().manual_seed(1)).cuda()
    metas = [dict(img_shape=(H, W, 3), batch_input_shape=(H, W))] * 2
    rels, labels, masks = synthetic_targets(2, (H, W), 5, "cuda")
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    ts = TrainStep(model, scope=scope, lr=1e-3, amp_dtype=amp)
    torch.manual_seed(7)
    hist = []
    for _ in range(6):
        losses = ts(imgs, metas, rels, labels, masks)
        assert set(losses) == {"loss_r_cls", "loss_sub_cls", "loss_obj_cls", "loss_match"}
        assert all(bool(torch.isfinite(v)) for v in losses.values())
        hist.append(float((losses["loss_match"] + losses["loss_r_cls"]).detach()))
    trainable = {"bbox_head." + n for n, _ in th.trainable_parameters(model.bbox_head, scope)}
    for n, p in model.named_parameters():
        changed = not torch.equal(p.detach(), before[n])
        assert changed == (n in trainable), (n, changed)
    assert hist[-1] < hist[0]                                  # the same batch six times: the trained losses go down
    assert float(model.bbox_head.rel_cls_loss.cum_samples.sum()) > 0
    ts.reducer.close()


def test_inference_after_training_uses_the_updated_weights():
    """The CUDA library caches TF32 splits of the weights (prepared blobs); an optimizer step updates the parameters in
    place, so the cache key carries the tensors' version counters."""
    head = product_small_head(oracle_small_head())
    from oracle.make_golden import small_head_inputs
    mf, mems = small_head_inputs(1, (16, 24), 31)
    mf, mems = mf.cuda(), [m.cuda() for m in mems]
    a, _ = head.forward_from_memories(mf, mems)
    a = {k: v.clone() for k, v in a.items()}
    with torch.no_grad():
        head.rel_cls_embed.weight.mul_(1.5)
        head.relation_decoder.layers[0].ffns[0].layers[1].weight.add_(0.01)
    b, _ = head.forward_from_memories(mf, mems)
    assert not torch.allclose(a["rel"], b["rel"])
    assert torch.equal(a["cls"], b["cls"])
