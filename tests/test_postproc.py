"""Inference post-processing (``CrossHead2.get_bboxes`` pairnet_head.py:759-924, ``triplet2Result``
psgtr.py:15-71).  ``postproc_ref_*.npz`` were minted by the REFERENCE's own ``get_bboxes`` executed from /root/reference
(``python -m oracle.postproc``).  CPU: the oracle restatement against those goldens.  GPU: the B200 path
(``pn_upsample_threshold`` / ``pn_panoptic_merge`` + host logic) against the goldens -- integer outputs exact wherever the
deciding logit margin is not a rounding-level tie."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import postproc as pp
from tests.util import GOLDEN


def _golden(tag):
    return np.load(os.path.join(GOLDEN, f"postproc_ref_{tag}.npz"))


def _unpack_masks(g, i):
    shape = tuple(g[f"mask_shape{i}"])
    return np.unpackbits(g[f"masks{i}"], axis=-1)[..., : shape[-1]].astype(bool)


@pytest.mark.parametrize("case", pp.POST_CASES, ids=[c[0] for c in pp.POST_CASES])
def test_oracle_postproc_matches_reference_golden(case):
    g = _golden(case[0])
    cls_scores, mask_preds, metas, _, _ = pp.case_inputs(case)
    res = pp.get_bboxes(cls_scores, mask_preds, metas, 56, case[3])
    for i, (bb, labels, pairs, masks, pan, rs, rl, rd) in enumerate(res):
        assert np.array_equal(labels.numpy(), g[f"labels{i}"])
        assert np.array_equal(masks.numpy(), _unpack_masks(g, i))
        assert np.array_equal(pan.numpy(), g[f"pan{i}"])
        assert np.array_equal(rd.numpy(), g[f"r_dists{i}"])
        assert np.array_equal(pairs.numpy(), g[f"rel_pairs{i}"])
        assert bb.shape == (2 * case[3], 5) and not bb.any()


def test_oracle_postproc_equals_live_reference_if_present():
    if not os.path.isdir("/root/reference/pairnet"):
        pytest.skip("reference tree not present on this box")
    import subprocess  # subprocess: the harness installs mmcv/mmdet constructor shims into sys.modules
    import sys
    from tests.util import ROOT
    r = subprocess.run([sys.executable, "-m", "oracle.postproc", "--check"], cwd=ROOT, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("pinned ") == len(pp.POST_CASES)


def test_result_container_and_triplet2result():
    from pairnet.models.relation_heads.approaches import Result
    from pairnet_b200.results import triplet2Result
    t = (torch.zeros(4, 5), torch.tensor([1, 2, 3, 4]), torch.arange(4, dtype=torch.int).reshape(2, -1).T,
         torch.ones(4, 3, 3, dtype=torch.bool), torch.ones(3, 3, dtype=torch.long), torch.zeros(2), torch.zeros(2),
         torch.rand(2, 57))
    r = triplet2Result(t, use_mask=True)
    assert isinstance(r, Result) and len(r) == 1 and r[0] is r and list(r) == [r]
    assert isinstance(r.labels, np.ndarray) and r.rel_pair_idxes.tolist() == [[0, 2], [1, 3]]
    assert r.formatted_masks["pan_results"] is r.pan_results and r.masks.dtype == bool
    assert Result().is_none() and not r.is_none()
    with pytest.raises(TypeError):
        Result(no_such_field=1)


def test_stuff_remap_host_logic_matches_reference_masked_fill_semantics():
    """CPU: the remap table handed to ``pn_panoptic_merge`` vs the reference's in-place ``masked_fill_`` loop."""
    from collections import defaultdict
    from pairnet_b200.head import CrossHead2
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(1, 40))
        labels = rng.integers(60, 100, n).tolist()
        m_id = torch.from_numpy(rng.integers(0, n, (6, 7)))
        equiv = defaultdict(list)
        for k, lab in enumerate(labels):
            if lab >= 80:
                equiv[lab].append(k)
        ref = m_id.clone()
        for e in equiv.values():
            if len(e) > 1:
                for eq_id in e:
                    ref.masked_fill_(ref.eq(eq_id), e[0])
        remap = torch.tensor(CrossHead2._stuff_remap(labels))
        assert torch.equal(remap[m_id], ref)
    assert CrossHead2._stuff_remap([3, 85, 85, 90, 3, 85]) == [0, 1, 1, 3, 4, 1]


def _product_head(N, K):
    from oracle.head import HeadHyper, OCrossHead2
    from oracle.weights import fixture_state_dict
    from pairnet_b200.registry import build_head
    from tests.util import product_head_cfg
    cfg = product_head_cfg()
    cfg.update(pixel_decoder=None, num_obj_query=N, num_rel_query=K)
    p = build_head(cfg)
    shell = OCrossHead2(HeadHyper(num_obj_query=N, num_rel_query=K, with_pixel_decoder=False))
    p.load_state_dict(fixture_state_dict(shell, 10086), strict=True)
    return p.cuda().eval()


@pytest.mark.gpu
@pytest.mark.parametrize("case", pp.POST_CASES, ids=[c[0] for c in pp.POST_CASES])
def test_gpu_postproc_matches_reference_golden(case):
    tag, B, N, K, hw4, img_shape, sf, seed = case[:8]
    g = _golden(tag)
    cls_scores, mask_preds, metas, _, _ = pp.case_inputs(case)
    head = _product_head(N, K)
    cu = lambda d: {k: v.cuda() for k, v in d.items()}
    res = head.get_bboxes(cu(cls_scores), cu(mask_preds), metas)
    assert len(res) == B
    for i, (bb, labels, pairs, masks, pan, rs, rl, rd) in enumerate(res):
        H, W = g[f"pan{i}"].shape
        assert np.array_equal(labels.cpu().numpy(), g[f"labels{i}"])
        assert np.array_equal(pairs.numpy(), g[f"rel_pairs{i}"]) and pairs.dtype == torch.int32
        assert float((rd.cpu() - torch.from_numpy(g[f"r_dists{i}"])).abs().max()) < 1e-6
        assert masks.dtype == torch.bool and masks.shape == (2 * K, H, W) and masks.is_cuda
        assert pan.dtype == torch.long and not pan.is_cuda and bb.shape == (2 * K, 5) and not bb.any()
        # thresholded masks: exact wherever the upsampled logit is not within rounding of 0
        up = lambda m: F.interpolate(m.double().unsqueeze(1), size=(H, W), mode="bilinear", align_corners=False).squeeze(1)
        logit = torch.cat((up(mask_preds["sub_seg"][i]), up(mask_preds["obj_seg"][i])), 0)
        ref_m, got_m = torch.from_numpy(_unpack_masks(g, i)), masks.cpu()
        sure = logit.abs() > 1e-4
        assert torch.equal(got_m[sure], ref_m[sure])
        assert (got_m != ref_m).float().mean() < 1e-5
        # panoptic map: the reference's segment ids / labels, exact up to rounding-level argmax ties
        diff = (pan.numpy() != g[f"pan{i}"]).mean()
        assert diff < 1e-4, diff
        assert set(np.unique(pan.numpy())) == set(np.unique(g[f"pan{i}"]))


@pytest.mark.gpu
def test_gpu_upsample_threshold_with_index_and_full_size():
    """index path (masks[r] = up(mask[idx[r]]) > 0) at the BASELINE size 200x334 -> 800x1333, against ATen on the GPU."""
    from pairnet_b200 import _native as nat
    lib = nat.load()
    N, R, h, w, H, W = 100, 40, 200, 334, 800, 1333
    g = torch.Generator().manual_seed(5)
    mask = (torch.randn(N, h // 8, w // 8, generator=g) * 3).cuda()
    mask = F.interpolate(mask[None], size=(h, w), mode="bicubic", align_corners=False)[0].contiguous()
    idx = torch.randint(0, N, (R,), generator=g).cuda()
    out = torch.empty((R, H, W), dtype=torch.bool, device="cuda")
    nat.check(lib.pn_upsample_threshold(mask.data_ptr(), idx.data_ptr(), out.data_ptr(), N, R, h, w, H, W,
                                        torch.cuda.current_stream().cuda_stream), "up")
    ref_logit = F.interpolate(mask[idx].double()[None], size=(H, W), mode="bilinear", align_corners=False)[0]
    sure = ref_logit.abs() > 1e-4
    assert torch.equal(out[sure], (ref_logit > 0)[sure])
    assert (out != (ref_logit > 0)).float().mean() < 1e-5


@pytest.mark.gpu
def test_gpu_simple_test_end_to_end_small_image():
    """PSGTr.simple_test: backbone -> pixel decoder -> head -> get_bboxes -> triplet2Result (the evaluation entry)."""
    from pairnet_b200.registry import Config, build_detector
    from pairnet_b200.results import Result
    from tests.util import ROOT
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "pairnet_r50_b200.py"))
    torch.manual_seed(10086)
    model = build_detector(cfg.model)
    model.init_weights()
    model = model.cuda().eval()
    img = torch.randn(2, 3, 256, 320, device="cuda")
    metas = [dict(img_shape=(256, 320, 3), scale_factor=np.ones(4, dtype=np.float32), batch_input_shape=(256, 320))] * 2
    with torch.no_grad():
        res = model.simple_test(img, metas)
    assert len(res) == 2 and all(isinstance(r, Result) for r in res)
    r = res[0]
    assert r.masks.shape == (200, 256, 320) and r.masks.dtype == bool
    assert r.pan_results.shape == (256, 320) and r.rel_dists.shape == (100, 57) and r.labels.shape == (200,)
    assert r.rel_pair_idxes.shape == (100, 2) and r.refine_bboxes.shape == (200, 5)
