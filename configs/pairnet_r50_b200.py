# Pair-Net R50 / Mask2Former, 100 object queries, 100 relation queries -- the model of BASELINE.json's
# configs[0..2].  Hyper-parameters equal the reference's `configs/mask2former/pairnet.py` (which this
# framework also loads unchanged); this file only keeps what the forward needs and builds the nested
# dicts with helpers instead of spelling them out.
num_object_classes = 133
num_relation_classes = 56
EMBED = 256


def _mha():
    return dict(type="MultiheadAttention", embed_dims=EMBED, num_heads=8, attn_drop=0.0, proj_drop=0.0,
                dropout_layer=None, batch_first=False)


def _decoder(num_layers, ffn_drop, return_intermediate):
    return dict(
        type="DetrTransformerDecoder", return_intermediate=return_intermediate, num_layers=num_layers,
        transformerlayers=dict(
            type="BaseTransformerLayer", attn_cfgs=_mha(),
            ffn_cfgs=dict(embed_dims=EMBED, feedforward_channels=2048, num_fcs=2, act_cfg=dict(type="ReLU", inplace=True),
                          ffn_drop=ffn_drop, dropout_layer=None, add_identity=True),
            operation_order=("cross_attn", "norm", "self_attn", "norm", "ffn", "norm")))


_sine = dict(type="SinePositionalEncoding", num_feats=EMBED // 2, normalize=True)

_pixel_decoder = dict(
    type="MSDeformAttnPixelDecoder", num_outs=3, norm_cfg=dict(type="GN", num_groups=32), act_cfg=dict(type="ReLU"),
    encoder=dict(type="DetrTransformerEncoder", num_layers=6,
                 transformerlayers=dict(type="BaseTransformerLayer",
                                        attn_cfgs=dict(type="MultiScaleDeformableAttention", embed_dims=EMBED, num_heads=8,
                                                       num_levels=3, num_points=4, im2col_step=64, dropout=0.0,
                                                       batch_first=False, norm_cfg=None, init_cfg=None),
                                        ffn_cfgs=dict(type="FFN", embed_dims=EMBED, feedforward_channels=1024, num_fcs=2,
                                                      ffn_drop=0.0, act_cfg=dict(type="ReLU", inplace=True)),
                                        operation_order=("self_attn", "norm", "ffn", "norm")),
                 init_cfg=None),
    positional_encoding=_sine, init_cfg=None)

model = dict(
    type="PSGTr",
    backbone=dict(type="ResNet", depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                  norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="pytorch"),
    bbox_head=dict(
        type="CrossHead2", num_classes=num_object_classes, num_relations=num_relation_classes,
        num_obj_query=100, num_rel_query=100, mapper="conv_tiny", in_channels=[256, 512, 1024, 2048],
        feat_channels=EMBED, out_channels=EMBED, num_transformer_feat_level=3, embed_dims=EMBED,
        enforce_decoder_input_project=False, pixel_decoder=_pixel_decoder,
        transformer_decoder=_decoder(9, 0.0, False), relation_decoder=_decoder(6, 0.1, True),
        positional_encoding=_sine,
        rel_cls_loss=dict(type="SeesawLoss", num_classes=num_relation_classes, return_dict=True, loss_weight=2.0),
        subobj_cls_loss=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=4.0, reduction="mean",
                             class_weight=[1.0] * (num_object_classes + 1)),
        importance_match_loss=dict(type="BCEWithLogitsLoss", reduction="mean", loss_weight=5.0),
        loss_cls=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=2.0, reduction="mean",
                      class_weight=[1.0] * num_object_classes + [0.1]),
        loss_mask=dict(type="CrossEntropyLoss", use_sigmoid=True, reduction="mean", loss_weight=5.0),
        loss_dice=dict(type="DiceLoss", use_sigmoid=True, activate=True, reduction="mean", naive_dice=True, eps=1.0,
                       loss_weight=5.0)),
    # configs/mask2former/pairnet.py:190-207
    train_cfg=dict(
        id_assigner=dict(type="IdMatcher", sub_id_cost=dict(type="ClassificationCost", weight=1.0),
                         obj_id_cost=dict(type="ClassificationCost", weight=1.0),
                         r_cls_cost=dict(type="ClassificationCost", weight=0.0)),
        num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75,
        mask_assigner=dict(type="MaskHungarianAssigner", cls_cost=dict(type="ClassificationCost", weight=2.0),
                           mask_cost=dict(type="CrossEntropyLossCost", weight=5.0, use_sigmoid=True),
                           dice_cost=dict(type="DiceCost", weight=5.0, pred_act=True, eps=1.0)),
        sampler=dict(type="MaskPseudoSampler")),
    test_cfg=dict(max_per_img=100),
)

custom_imports = dict(imports=["pairnet.models.frameworks.psgtr"], allow_failed_imports=False)
