"""Pin the oracle to the REFERENCE's own ``CrossHead2.forward`` (TEST INFRASTRUCTURE, dev container only).

    python -m oracle.pin_reference            # check + mint tests/golden/head_ref_*.npz
    python -m oracle.pin_reference --check    # check only (used by tests/test_oracle_golden.py)

The reference head (``/root/reference/pairnet/models/relation_heads/pairnet_head.py``) cannot be imported as
is: its module top imports mmcv-full 1.7.0 / mmdet 2.25.1, which are absent.  This script installs *constructor
shims* for exactly the names those imports bind, then loads and EXECUTES the reference's own source files by
path -- nothing is copied into this repository:

* ``pairnet_head.py``   -> the real ``CrossHead2.__init__/_init_layers/forward_head/forward``  (:22-417)
* ``cnn_factory.py``    -> the real ``creat_cnn`` / ``ConvTiny``
* ``facebook_detr.py``  -> the real ``MultiheadAttention2.forward`` (:311-353) and ``BaseTransformerLayer2.forward``
                           (:378-432), the reference's in-repo restatement of mmcv's attention wrapper / layer
                           driver (they additionally return attention maps, dropped by a 3-line adapter below)

What the shims supply (and what therefore stays "parity unpinned", restated from the published upstream code):
the *constructors* of mmcv ``MultiheadAttention`` / ``BaseTransformerLayer`` / ``FFN`` (module containers with
mmcv's attribute and parameter names), mmdet ``DetrTransformerDecoder`` (``layers`` + ``post_norm``),
``SinePositionalEncoding.forward`` and the registries/loss builders that the forward never calls.  The pixel
decoder is a pass-through stub (it is upstream of the hot path): ``feats`` is ``(mask_features, memories)``.

The script then loads the oracle fixture weights into the reference head with ``strict=True`` (pins the
parameter names of SURVEY §8b), runs both on the same inputs and requires BIT-EQUAL outputs, and writes the
reference's outputs as ``tests/golden/head_ref_*.npz``.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

REF_CASES = [  # (tag, B, (H4,W4) of mask_feature, seed, N obj queries, R rel queries)
    ("b2_32x48", 2, (32, 48), 21, 100, 100),
    ("b1_40x56", 1, (40, 56), 22, 100, 100),
    ("b3_24x40_n48_r32", 3, (24, 40), 23, 48, 32),   # spread pair rows (20-25 distinct subjects of 32)
    ("b1_32x32_n200_r200", 1, (32, 32), 24, 200, 200),   # BASELINE config 4's query count
]


class AttrDict(dict):
    """mmcv ``ConfigDict`` stand-in: attribute access + ``.get``; nested dicts converted."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, AttrDict):
                self[k] = AttrDict(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __deepcopy__(self, memo):
        import copy
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _Registry:
    def __init__(self):
        self.d = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.d[name or cls.__name__] = cls
            return cls
        return deco


def install_shims():
    """Bind every mmcv / mmdet / pairnet name the two reference files import at module top."""
    from . import bricks as ob

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    class AnchorFreeHead(BaseModule):  # only its MRO position is used: super(AnchorFreeHead, self).__init__(init_cfg)
        pass

    # ---- constructor shims with mmcv 1.7.0 attribute names (forward comes from the reference where it has one)
    class MultiheadAttention(BaseModule):
        def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0,
                     dropout_layer=dict(type="Dropout", drop_prob=0.0), init_cfg=None, batch_first=False, **kw):
            super().__init__(init_cfg)
            self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
            self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)
            self.proj_drop = nn.Dropout(proj_drop)
            self.dropout_layer = nn.Dropout(dropout_layer.get("drop_prob", 0.0)) if dropout_layer else nn.Identity()

    class FFN(BaseModule):  # mmcv FFN (not on disk): restated
        def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=None, ffn_drop=0.0,
                     dropout_layer=None, add_identity=True, init_cfg=None, **kw):
            super().__init__(init_cfg)
            assert num_fcs == 2
            self.layers = nn.Sequential(
                nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
                nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))
            self.dropout_layer = nn.Identity()
            self.add_identity = add_identity

        def forward(self, x, identity=None):
            out = self.layers(x)
            if not self.add_identity:
                return self.dropout_layer(out)
            if identity is None:
                identity = x
            return identity + self.dropout_layer(out)

    tr = _module("mmcv.cnn.bricks.transformer")

    class BaseTransformerLayer(BaseModule):
        def __init__(self, attn_cfgs=None, ffn_cfgs=None, operation_order=None, norm_cfg=dict(type="LN"),
                     init_cfg=None, batch_first=False, **kw):
            super().__init__(init_cfg)
            self.batch_first = batch_first
            self.operation_order = tuple(operation_order)
            self.num_attn = operation_order.count("self_attn") + operation_order.count("cross_attn")
            self.pre_norm = operation_order[0] == "norm"
            if isinstance(attn_cfgs, dict):
                attn_cfgs = [dict(attn_cfgs) for _ in range(self.num_attn)]
            self.attentions = nn.ModuleList()
            for cfg in attn_cfgs:
                cfg = {k: v for k, v in cfg.items() if k != "type"}
                cfg.setdefault("batch_first", batch_first)
                self.attentions.append(tr.ATTENTION_CLS(**cfg))
            self.embed_dims = self.attentions[0].embed_dims
            n_ffn = operation_order.count("ffn")
            if isinstance(ffn_cfgs, dict):
                ffn_cfgs = [dict(ffn_cfgs) for _ in range(n_ffn)]
            self.ffns = nn.ModuleList()
            for cfg in ffn_cfgs:
                cfg = {k: v for k, v in cfg.items() if k != "type"}
                cfg.setdefault("embed_dims", self.embed_dims)
                self.ffns.append(FFN(**cfg))
            self.norms = nn.ModuleList([nn.LayerNorm(self.embed_dims) for _ in range(operation_order.count("norm"))])

    class SinePositionalEncoding(BaseModule):  # mmdet (not on disk): restated in oracle.bricks
        def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * np.pi, eps=1e-6, offset=0.0,
                     init_cfg=None):
            super().__init__(init_cfg)
            assert normalize
            self.kw = dict(num_feats=num_feats, temperature=temperature, scale=scale, eps=eps, offset=offset)

        def forward(self, mask):
            return ob.sine_positional_encoding(mask, **self.kw)

    class PassThroughPixelDecoder(BaseModule):
        def __init__(self, **kw):
            super().__init__()

        def init_weights(self):
            pass

        def forward(self, feats):
            return feats

    class _Loss(nn.Module):
        """loss builders are never called by the forward; mmdet's SeesawLoss contributes the persistent buffer
        ``cum_samples`` [num_classes + 1] to the module tree (and so to every reference checkpoint)."""

        def __init__(self, cfg):
            super().__init__()
            self.use_sigmoid = cfg.get("use_sigmoid", False)
            if cfg.get("type") == "SeesawLoss":
                self.register_buffer("cum_samples", torch.zeros(cfg.get("num_classes", 1203) + 1))

    def build_transformer_layer_sequence(cfg):
        """mmdet ``DetrTransformerDecoder`` container: ``layers`` (deep copies of one layer cfg), ``post_norm``,
        ``embed_dims``; its own forward is never called by the head."""
        seq = BaseModule()
        lcfg = {k: v for k, v in cfg["transformerlayers"].items() if k != "type"}
        seq.layers = nn.ModuleList([tr.LAYER_CLS(**lcfg) for _ in range(cfg["num_layers"])])
        seq.embed_dims = seq.layers[0].embed_dims
        seq.post_norm = nn.LayerNorm(seq.embed_dims)
        return seq

    tr.__dict__.update(MultiheadAttention=MultiheadAttention, BaseTransformerLayer=BaseTransformerLayer, FFN=FFN,
                       build_positional_encoding=lambda cfg: SinePositionalEncoding(
                           **{k: v for k, v in cfg.items() if k != "type"}),
                       build_transformer_layer_sequence=build_transformer_layer_sequence)
    mmcv = _module("mmcv")
    cnn = _module("mmcv.cnn", Conv2d=nn.Conv2d, Linear=nn.Linear, caffe2_xavier_init=lambda *a, **k: None,
                  build_plugin_layer=lambda cfg: ("pixel_decoder", PassThroughPixelDecoder()))
    bricks = _module("mmcv.cnn.bricks", transformer=tr)
    _module("mmcv.cnn.bricks.registry", ATTENTION=_Registry(), TRANSFORMER_LAYER=_Registry())
    cnn.bricks = bricks
    mmcv.cnn = cnn
    _module("mmcv.ops", point_sample=None)
    _module("mmcv.runner", ModuleList=nn.ModuleList, force_fp32=lambda **kw: (lambda f: f), BaseModule=BaseModule)
    _module("mmdet")
    _module("mmdet.core", build_assigner=lambda c: None, build_sampler=lambda c, **k: None, multi_apply=None)
    _module("mmdet.datasets")
    _module("mmdet.datasets.coco_panoptic", INSTANCE_OFFSET=1000)
    _module("mmdet.models")
    _module("mmdet.models.builder", HEADS=_Registry(), build_loss=lambda cfg: _Loss(cfg))
    _module("mmdet.models.dense_heads", AnchorFreeHead=AnchorFreeHead)
    # the reference's own torch-only leaf module, under the dotted path pairnet_head.py imports it by
    for pkg in ("pairnet", "pairnet.models", "pairnet.models.frameworks"):
        _module(pkg)
    _load("pairnet.models.frameworks.cnn_factory", f"{REF}/pairnet/models/frameworks/cnn_factory.py")
    _module("pairnet.models.frameworks.unet", UNet=None)
    return tr


def load_reference_head_class():
    """-> the reference's ``CrossHead2`` class, wired to the reference's own layer / attention forwards."""
    tr = install_shims()
    fb = _load("ref_facebook_detr", f"{REF}/pairnet/models/relation_heads/facebook_detr.py")

    class RefAttention(fb.MultiheadAttention2):       # reference forward (:311-353) returns (out, attn_map)
        pass

    class RefLayer(fb.BaseTransformerLayer2):         # reference forward (:378-432) returns (query, map, map)
        def forward(self, *a, **kw):
            return super().forward(*a, **kw)[0]

    tr.ATTENTION_CLS, tr.LAYER_CLS = RefAttention, RefLayer
    ph = _load("ref_pairnet_head", f"{REF}/pairnet/models/relation_heads/pairnet_head.py")
    return ph.CrossHead2


def reference_head_cfg(num_obj_query=100, num_rel_query=100):
    """The ``bbox_head`` dict of ``configs/mask2former/pairnet.py:20-170`` (hyper-parameters the forward uses)."""
    def layer(ffn_drop):
        return dict(
            type="DetrTransformerDecoderLayer",
            attn_cfgs=dict(type="MultiheadAttention", embed_dims=256, num_heads=8, attn_drop=0.0, proj_drop=0.0,
                           dropout_layer=None, batch_first=False),
            ffn_cfgs=dict(embed_dims=256, feedforward_channels=2048, num_fcs=2, act_cfg=dict(type="ReLU", inplace=True),
                          ffn_drop=ffn_drop, dropout_layer=None, add_identity=True),
            feedforward_channels=2048,
            operation_order=("cross_attn", "norm", "self_attn", "norm", "ffn", "norm"))
    return AttrDict(
        num_classes=133, num_relations=56, num_obj_query=num_obj_query, num_rel_query=num_rel_query,
        mapper="conv_tiny", in_channels=[256, 512, 1024, 2048], feat_channels=256, out_channels=256,
        num_transformer_feat_level=3, embed_dims=256, use_mask=True,
        pixel_decoder=dict(type="MSDeformAttnPixelDecoder", encoder=dict(transformerlayers=dict(attn_cfgs=dict(num_levels=3)))),
        transformer_decoder=dict(type="DetrTransformerDecoder", return_intermediate=True, num_layers=9,
                                 transformerlayers=layer(0.0), init_cfg=None),
        relation_decoder=dict(type="DetrTransformerDecoder", return_intermediate=True, num_layers=6,
                              transformerlayers=layer(0.1), init_cfg=None),
        positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True),
        loss_cls=dict(type="CrossEntropyLoss", use_sigmoid=False, class_weight=[1.0] * 133 + [0.1]),
        loss_mask=dict(type="CrossEntropyLoss", use_sigmoid=True), loss_dice=dict(type="DiceLoss"),
        rel_cls_loss=dict(type="SeesawLoss", num_classes=56), subobj_cls_loss=dict(type="CrossEntropyLoss"),
        importance_match_loss=dict(type="BCEWithLogitsLoss"), train_cfg=None, test_cfg=dict(max_per_img=100))


def build_reference_head(num_obj_query=100, num_rel_query=100):
    cls = load_reference_head_class()
    cfg = reference_head_cfg(num_obj_query, num_rel_query)
    for k in ("transformer_decoder", "relation_decoder"):   # the layer builder drops keys BaseTransformerLayer lacks
        cfg[k]["transformerlayers"].pop("feedforward_channels")
    return cls(**cfg).eval()


def run_case(ref_head, B, hw4, seed):
    from .head import HeadHyper, OCrossHead2
    from .make_golden import small_head_inputs
    from .weights import fixture_state_dict
    N, R = ref_head.num_obj_query, ref_head.num_rel_query
    oracle = OCrossHead2(HeadHyper(with_pixel_decoder=False, num_obj_query=N, num_rel_query=R)).eval()
    sd = fixture_state_dict(oracle, 10086)
    oracle.load_state_dict(sd)
    missing = ref_head.load_state_dict(sd, strict=True)   # parameter-name pin
    assert not missing.missing_keys and not missing.unexpected_keys
    mf, mems = small_head_inputs(B, hw4, seed)
    with torch.no_grad():
        rcls, rmsk = ref_head((mf, mems), [dict()] * B)
        tr = {}
        ocls, omsk = oracle.forward_from_memories(mf, mems, trace=tr)
    for k in rcls:
        assert torch.equal(rcls[k], ocls[k]), f"oracle != reference forward on all_cls_scores[{k}]"
    for k in rmsk:
        assert torch.equal(rmsk[k], omsk[k]), f"oracle != reference forward on all_mask_preds[{k}]"
    return rcls, rmsk, tr


def main(check_only=False):
    assert os.path.isdir(REF), "the reference tree is only present in the dev container"
    heads = {}
    for tag, B, hw4, seed, N, R in REF_CASES:
        if (N, R) not in heads:
            heads[(N, R)] = build_reference_head(N, R)
        rcls, rmsk, tr = run_case(heads[(N, R)], B, hw4, seed)
        print(f"pinned {tag}: reference CrossHead2.forward == oracle, bit-equal on "
              f"{sorted(rcls) + sorted(rmsk)}; sub_pos[0,:5]={tr['sub_pos'][0, :5].tolist()}")
        if check_only:
            continue
        os.makedirs(GOLDEN, exist_ok=True)
        np.savez_compressed(
            os.path.join(GOLDEN, f"head_ref_{tag}.npz"),
            cls=rcls["cls"].numpy(), rel=rcls["rel"].numpy(), importance=rcls["importance"].numpy(),
            sub=rcls["sub"].numpy(), obj=rcls["obj"].numpy(),
            mask_sub4=rmsk["mask"][:, :, ::4, ::4].numpy(),
            sub_seg_sub4=rmsk["sub_seg"][:, :, ::4, ::4].numpy(), obj_seg_sub4=rmsk["obj_seg"][:, :, ::4, ::4].numpy(),
            sub_pos=tr["sub_pos"].numpy(), obj_pos=tr["obj_pos"].numpy(),
            meta=np.array([B, hw4[0], hw4[1], seed, N, R]))
    return 0


if __name__ == "__main__":
    sys.exit(main(check_only="--check" in sys.argv))
