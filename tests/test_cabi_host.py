"""CPU: the C-ABI library builds/loads and exports every symbol ``include/pairnet_b200.h`` declares
(no compute calls); the host-side registry/config surface mirrors the reference's."""
import ctypes
import os
import re

import pytest
import torch

from tests.util import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pairnet_b200.h")).read()
    return sorted(set(re.findall(r"PN_API[^;(]*?\b(pn_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from pairnet_b200 import _native
    lib = _native.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _native.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_native.SIGNATURES) == set(syms)
    assert lib.pn_version() == 100
    assert isinstance(lib.pn_last_error_string(), bytes)


def test_struct_layouts_match_header_sizes():
    """sizeof() of the ctypes mirrors vs a C translation unit compiled against the real header."""
    import subprocess
    import tempfile
    from pairnet_b200 import _native as n
    src = r'''
#include <stdio.h>
#include "pairnet_b200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(PnDecoderLayer), sizeof(PnM2FWeights),
 sizeof(PnM2FInputs), sizeof(PnM2FOutputs), sizeof(PnRelWeights), sizeof(PnHeadWeights), sizeof(PnHeadOutputs),
 sizeof(PnConvTiny));return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mine = [ctypes.sizeof(t) for t in (n.PnDecoderLayer, n.PnM2FWeights, n.PnM2FInputs, n.PnM2FOutputs,
                                       n.PnRelWeights, n.PnHeadWeights, n.PnHeadOutputs, n.PnConvTiny)]
    assert sizes == mine


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pairnet_b200 import _native, ops
    with pytest.raises(_native.NativeError):
        ops.conv_tiny(torch.zeros(1, 8, 8), None)


def test_product_never_imports_oracle():
    import subprocess
    import sys
    code = "import sys; import pairnet_b200, pairnet.models; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pairnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("the CPU oracle", "").replace("CPU oracle", ""), f


def test_registry_surface_and_state_dict_names():
    from pairnet_b200.registry import Config, build_detector
    import pairnet.models.relation_heads.pairnet_head as ph
    import pairnet.models.frameworks.psgtr as pg
    from pairnet_b200 import CrossHead2, PSGTr
    assert ph.CrossHead2 is CrossHead2 and pg.PSGTr is PSGTr
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "pairnet_r50_b200.py"))
    model = build_detector(cfg.model)
    keys = set(model.state_dict().keys())
    expect = [
        "bbox_head.relation_decoder.layers.5.attentions.1.attn.in_proj_weight",
        "bbox_head.relation_decoder.layers.0.ffns.0.layers.0.0.weight",
        "bbox_head.relation_decoder.layers.0.ffns.0.layers.1.bias",
        "bbox_head.relation_decoder.post_norm.weight",
        "bbox_head.transformer_decoder.layers.8.norms.2.bias",
        "bbox_head.transformer_decoder.post_norm.bias",
        "bbox_head.rel_query_embed.weight", "bbox_head.rel_query_embed2.weight", "bbox_head.rel_query_embed3.weight",
        "bbox_head.rel_query_feat.weight", "bbox_head.update_importance.conv_layers.1.0.weight",
        "bbox_head.query_embed.weight", "bbox_head.query_feat.weight", "bbox_head.level_embed.weight",
        "bbox_head.cls_embed.weight", "bbox_head.mask_embed.4.bias", "bbox_head.sub_query_update.0.weight",
        "bbox_head.obj_query_update.4.bias", "bbox_head.rel_cls_embed.weight",
        "bbox_head.pixel_decoder.input_convs.0.conv.weight", "bbox_head.pixel_decoder.input_convs.2.gn.bias",
        "bbox_head.pixel_decoder.encoder.layers.5.attentions.0.sampling_offsets.weight",
        "bbox_head.pixel_decoder.level_encoding.weight", "bbox_head.pixel_decoder.lateral_convs.0.conv.weight",
        "bbox_head.pixel_decoder.output_convs.0.gn.weight", "bbox_head.pixel_decoder.mask_feature.bias",
        "backbone.layer4.2.conv3.weight",
    ]
    for k in expect:
        assert k in keys, k
    sd = model.state_dict()
    assert tuple(sd["bbox_head.relation_decoder.layers.0.attentions.0.attn.in_proj_weight"].shape) == (768, 256)
    assert tuple(sd["bbox_head.rel_query_embed2.weight"].shape) == (200, 256)
    assert tuple(sd["bbox_head.update_importance.conv_layers.1.0.weight"].shape) == (64, 64, 7, 7)
    assert tuple(sd["bbox_head.cls_embed.weight"].shape) == (134, 256)
    assert tuple(sd["bbox_head.rel_cls_embed.weight"].shape) == (56, 256)
    # the oracle (independent restatement) agrees on every key and shape
    from oracle.head import OPSGTr
    osd = OPSGTr().state_dict()
    assert set(osd.keys()) == keys
    assert all(osd[k].shape == sd[k].shape for k in keys)
    # mmdet's SeesawLoss (rel_cls_loss) keeps a persistent buffer: every reference checkpoint carries it, so the documented
    # `model.load_state_dict(ckpt["state_dict"])` (strict=True) must find a home for it
    assert "bbox_head.rel_cls_loss.cum_samples" in keys and tuple(sd["bbox_head.rel_cls_loss.cum_samples"].shape) == (57,)
    assert not any(k.startswith("bbox_head.loss_") or "subobj_cls_loss" in k or "importance_match_loss" in k for k in keys)
    model.load_state_dict(osd, strict=True)


def test_unsupported_geometry_is_rejected_not_miscomputed():
    """What the kernels hard-code (8 heads x 32, post-norm op order, the sine encoding's constants) must raise."""
    import copy
    from pairnet_b200.registry import build_head
    from tests.util import product_head_cfg
    base = product_head_cfg()
    base["pixel_decoder"] = None
    build_head(copy.deepcopy(base))
    bad = copy.deepcopy(base)
    bad["relation_decoder"]["transformerlayers"]["attn_cfgs"]["num_heads"] = 4
    with pytest.raises(NotImplementedError):
        build_head(bad)
    bad = copy.deepcopy(base)
    bad["positional_encoding"] = dict(type="SinePositionalEncoding", num_feats=128, normalize=False)
    with pytest.raises(NotImplementedError):
        build_head(bad)
    bad = copy.deepcopy(base)
    bad["positional_encoding"] = dict(type="SinePositionalEncoding", num_feats=128, normalize=True, temperature=20)
    with pytest.raises(NotImplementedError):
        build_head(bad)


def test_reference_config_loads_unchanged_if_present():
    path = "/root/reference/configs/mask2former/pairnet.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this box")
    from pairnet_b200.registry import Config, build_detector
    cfg = Config.fromfile(path)  # runs custom_imports with allow_failed_imports=False
    assert cfg.model.bbox_head.type == "CrossHead2" and cfg.find_unused_parameters is True
    assert cfg.dist_params.backend == "nccl"  # inherited from _base_/custom_runtime.py
    model = build_detector(cfg.model)
    assert model.bbox_head.num_rel_query == 100 and model.num_classes == 133


@pytest.mark.parametrize("name,backbone", [("pairnet_60e.py", "ResNet"), ("pairnet_swinb.py", "SwinTransformer")])
def test_other_shipped_pairnet_configs_build_unchanged_if_present(name, backbone):
    """The reference's other CrossHead2 configs (`configs/mask2former/`) load and build unchanged.  `pairnet_balanced.py`
    only LOADS: it passes `class_weight=` to mmdet's FocalLoss, which that class does not accept either."""
    path = "/root/reference/configs/mask2former/" + name
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this box")
    from pairnet_b200.registry import Config, build_detector
    model = build_detector(Config.fromfile(path).model)
    assert type(model.bbox_head).__name__ == "CrossHead2" and type(model.backbone).__name__ == backbone
    bal = Config.fromfile("/root/reference/configs/mask2former/pairnet_balanced.py")   # custom_imports resolve
    assert bal.model.bbox_head.type == "CrossHead2" and bal.model.bbox_head.rel_cls_loss.type == "FocalLoss"


def test_constructor_asserts_mirror_reference():
    from pairnet_b200.registry import build_head
    from tests.util import product_head_cfg
    cfg = product_head_cfg()
    cfg["pixel_decoder"] = None
    cfg["positional_encoding"] = dict(type="SinePositionalEncoding", num_feats=64, normalize=True)
    with pytest.raises(AssertionError):
        build_head(cfg)  # pairnet_head.py:74-78


def test_detector_forward_dispatch_mirrors_mmdet_base_detector():
    """Host logic only (no kernels): ``PSGTr.forward`` routes like mmdet's ``BaseDetector.forward`` -- the call
    ``tools/test.py`` makes is ``model(return_loss=False, rescale=True, img=[tensor], img_metas=[[dict, ...]])``."""
    from pairnet_b200.detector import PSGTr
    calls = []

    class Stub(PSGTr):
        def __init__(self):  # no backbone / head: dispatch only
            torch.nn.Module.__init__(self)

        def simple_test(self, img, img_metas, rescale=False):
            calls.append(("simple_test", tuple(img.shape), len(img_metas), rescale, img_metas[0]["batch_input_shape"]))
            return ["r"] * len(img_metas)

        def forward_dummy(self, img):
            calls.append(("forward_dummy", tuple(img.shape)))
            return "d"

    m = Stub()
    img = torch.zeros(2, 3, 8, 12)
    metas = [dict(img_shape=(8, 12, 3)), dict(img_shape=(8, 12, 3))]
    assert m(return_loss=False, rescale=True, img=[img], img_metas=[metas]) == ["r", "r"]
    assert calls[-1] == ("simple_test", (2, 3, 8, 12), 2, True, (8, 12))
    assert m(img) == "d" and calls[-1] == ("forward_dummy", (2, 3, 8, 12))
    with pytest.raises(TypeError):
        m(img, metas, return_loss=False)            # not wrapped in augmentation lists
    with pytest.raises(NotImplementedError):
        m([img, img], [metas, metas], return_loss=False)   # aug_test
    class TrainStub(Stub):
        def forward_train(self, img, img_metas, **kw):
            calls.append(("forward_train", tuple(img.shape), sorted(kw)))
            return dict(loss_match=0.0)
    t = TrainStub()
    assert t(img, metas, return_loss=True, gt_rels=[], gt_labels=[], gt_masks=[]) == dict(loss_match=0.0)
    assert calls[-1] == ("forward_train", (2, 3, 8, 12), ["gt_labels", "gt_masks", "gt_rels"])   # training: SURVEY 8f-2
