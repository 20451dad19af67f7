"""Parameter containers that reproduce the module tree (and therefore the ``state_dict`` key
names) of the mmcv-full 1.7.0 / mmdet 2.25.1 bricks ``CrossHead2`` is configured with, so that a
reference checkpoint loads with ``strict=True``.  They hold weights only: the
arithmetic of the hot path runs in the CUDA library (``pairnet_b200/csrc``); calling ``forward``
on these containers raises.

Config surface mirrored: ``configs/mask2former/pairnet.py:72-142`` (decoder / relation decoder /
positional encoding dicts)."""
import torch.nn as nn

from .registry import POSITIONAL_ENCODING, TRANSFORMER_LAYER_SEQUENCE

SUPPORTED_ORDER = ("cross_attn", "norm", "self_attn", "norm", "ffn", "norm")


class _WeightsOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} only stores weights; the hot path runs in libpairnet_b200.so "
                           "(call CrossHead2.forward)")


class MultiheadAttention(_WeightsOnly):
    """mmcv ``MultiheadAttention``: parameters live under ``.attn`` (an ``nn.MultiheadAttention``)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0, dropout_layer=None, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__()
        if batch_first:
            raise NotImplementedError("batch_first=True attention is not used by any Pair-Net config")
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)


class FFN(_WeightsOnly):
    """mmcv ``FFN`` (num_fcs=2): ``layers = Seq(Seq(Linear, act, Dropout), Linear, Dropout)``."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=None, ffn_drop=0.0,
                 dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        if num_fcs != 2 or not add_identity:
            raise NotImplementedError("only the 2-layer residual FFN of the Pair-Net configs is supported")
        if act_cfg is not None and act_cfg.get("type", "ReLU") != "ReLU":
            raise NotImplementedError("FFN activation must be ReLU")
        self.embed_dims, self.feedforward_channels = embed_dims, feedforward_channels
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims),
            nn.Dropout(ffn_drop),
        )


class BaseTransformerLayer(_WeightsOnly):
    def __init__(self, attn_cfgs=None, ffn_cfgs=None, operation_order=None, norm_cfg=None, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__()
        if tuple(operation_order) != SUPPORTED_ORDER:
            raise NotImplementedError(f"operation_order {operation_order} is not the Pair-Net decoder order "
                                      f"{SUPPORTED_ORDER}")
        if norm_cfg is not None and norm_cfg.get("type", "LN") != "LN":
            raise NotImplementedError("only LayerNorm decoder layers are supported (RMSNorm is a VG-config variant)")
        num_attn = 2
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [dict(attn_cfgs) for _ in range(num_attn)]
        self.attentions = nn.ModuleList()
        for cfg in attn_cfgs:
            cfg = dict(cfg)
            typ = cfg.pop("type", "MultiheadAttention")
            if typ != "MultiheadAttention":
                raise NotImplementedError(f"attention type {typ}")
            self.attentions.append(MultiheadAttention(**cfg))
        self.embed_dims = self.attentions[0].embed_dims
        ffn = dict(ffn_cfgs or {})
        ffn.pop("type", None)
        ffn.setdefault("embed_dims", self.embed_dims)
        self.ffns = nn.ModuleList([FFN(**ffn)])
        self.norms = nn.ModuleList([nn.LayerNorm(self.embed_dims) for _ in range(3)])
        self.operation_order = tuple(operation_order)
        self.pre_norm = False


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetrTransformerDecoder(_WeightsOnly):
    """mmdet ``DetrTransformerDecoder``: ``layers.{i}`` + ``post_norm``."""

    def __init__(self, transformerlayers=None, num_layers=None, post_norm_cfg=dict(type="LN"),
                 return_intermediate=False, init_cfg=None, **kwargs):
        super().__init__()
        if isinstance(transformerlayers, dict):
            transformerlayers = [dict(transformerlayers) for _ in range(num_layers)]
        self.layers = nn.ModuleList()
        for cfg in transformerlayers:
            cfg = dict(cfg)
            typ = cfg.pop("type", "BaseTransformerLayer")
            if typ != "BaseTransformerLayer":
                raise NotImplementedError(f"transformer layer type {typ}")
            self.layers.append(BaseTransformerLayer(**cfg))
        self.num_layers = num_layers
        self.embed_dims = self.layers[0].embed_dims
        self.return_intermediate = return_intermediate
        self.post_norm = nn.LayerNorm(self.embed_dims) if post_norm_cfg is not None else None


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding(nn.Module):
    """mmdet ``SinePositionalEncoding`` (no parameters).  The head evaluates it with the
    ``pn_sine_posenc`` kernel; this holder only carries the hyper-parameters."""

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * 3.141592653589793, eps=1e-6,
                 offset=0.0, init_cfg=None):
        super().__init__()
        self.num_feats, self.temperature, self.normalize = num_feats, temperature, normalize
        self.scale, self.eps, self.offset = scale, eps, offset
