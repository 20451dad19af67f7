"""Torch-only restatements of the mmcv-full 1.7.0 / mmdet 2.25.1 bricks that
``CrossHead2`` instantiates from its config.  TEST INFRASTRUCTURE (see
``oracle/__init__.py``).

Sources followed (reference tree, read-only):
* mmcv ``MultiheadAttention.forward``  -> ``pairnet/models/relation_heads/facebook_detr.py:311-353``
* mmcv ``BaseTransformerLayer.forward`` -> ``facebook_detr.py:378-432``
* hyper-parameters -> ``configs/mask2former/pairnet.py:20-142``
* module/parameter names -> SURVEY.md §8b (mmcv 1.7.0 naming)
mmcv ``FFN``, mmdet ``SinePositionalEncoding``, ``DetrTransformerDecoder`` and the
pixel decoder are not on disk; they are restated from the published upstream
algorithm ("parity unpinned").
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class OMultiheadAttention(nn.Module):
    """mmcv ``MultiheadAttention`` wrapper (batch_first=False)."""

    def __init__(self, embed_dims=256, num_heads=8):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, dropout=0.0)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_pos=None, attn_mask=None, **kwargs):
        # facebook_detr.py:322-353.  NB: ``value_pos`` lands in **kwargs and is ignored.
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        # mmcv calls ``self.attn(...)[0]`` with torch's default need_weights=True: the explicit
        # baddbmm -> softmax -> bmm path, not scaled_dot_product_attention.  Keep that arithmetic.
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        need_weights=True)[0]
        return identity + out  # proj_drop / dropout_layer are identities (p = 0)


class OFFN(nn.Module):
    """mmcv ``FFN``: layers = Seq(Seq(Linear, ReLU, Dropout), Linear, Dropout); out = x + layers(x)."""

    def __init__(self, embed_dims=256, feedforward_channels=2048):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                          nn.Dropout(0.0)),
            nn.Linear(feedforward_channels, embed_dims),
            nn.Dropout(0.0),
        )

    def forward(self, x, identity=None):
        if identity is None:
            identity = x
        return identity + self.layers(x)


class ODecoderLayer(nn.Module):
    """mmcv ``BaseTransformerLayer`` with operation_order
    (cross_attn, norm, self_attn, norm, ffn, norm)  -- configs/mask2former/pairnet.py:96-103."""

    def __init__(self, embed_dims=256, num_heads=8, feedforward_channels=2048):
        super().__init__()
        self.attentions = nn.ModuleList([OMultiheadAttention(embed_dims, num_heads),
                                         OMultiheadAttention(embed_dims, num_heads)])
        self.ffns = nn.ModuleList([OFFN(embed_dims, feedforward_channels)])
        self.norms = nn.ModuleList([nn.LayerNorm(embed_dims) for _ in range(3)])

    def forward(self, query, key, value, query_pos=None, key_pos=None, attn_masks=None, **kwargs):
        # facebook_detr.py:378-432 (pre_norm False -> identity argument is None everywhere)
        if attn_masks is None:
            attn_masks = [None, None]
        query = self.attentions[0](query, key, value, None, query_pos=query_pos, key_pos=key_pos,
                                   attn_mask=attn_masks[0], **kwargs)
        query = self.norms[0](query)
        query = self.attentions[1](query, query, query, None, query_pos=query_pos,
                                   key_pos=query_pos, attn_mask=attn_masks[1], **kwargs)
        query = self.norms[1](query)
        query = self.ffns[0](query, None)
        query = self.norms[2](query)
        return query


class ODetrTransformerDecoder(nn.Module):
    """mmdet ``DetrTransformerDecoder``: ``layers`` + ``post_norm`` (its own forward is bypassed
    by the head, pairnet_head.py:297,366)."""

    def __init__(self, num_layers, embed_dims=256, num_heads=8, feedforward_channels=2048):
        super().__init__()
        self.embed_dims = embed_dims
        self.layers = nn.ModuleList([ODecoderLayer(embed_dims, num_heads, feedforward_channels)
                                     for _ in range(num_layers)])
        self.post_norm = nn.LayerNorm(embed_dims)


def sine_positional_encoding(mask, num_feats=128, temperature=10000, scale=2 * math.pi,
                             eps=1e-6, offset=0.0, dtype=torch.float32):
    """mmdet ``SinePositionalEncoding(normalize=True)``; mask [B,H,W] bool -> [B,2*num_feats,H,W]."""
    mask = mask.to(torch.int)
    not_mask = 1 - mask
    y_embed = not_mask.cumsum(1, dtype=dtype)
    x_embed = not_mask.cumsum(2, dtype=dtype)
    y_embed = (y_embed + offset) / (y_embed[:, -1:, :] + eps) * scale
    x_embed = (x_embed + offset) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=dtype, device=mask.device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    B, H, W = mask.size()
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, W, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------
# Upstream of the hot path: mmdet MSDeformAttnPixelDecoder, torch-only.
# --------------------------------------------------------------------------------------
def ms_deform_attn_torch(value, spatial_shapes, sampling_locations, attention_weights):
    """mmcv ``multi_scale_deformable_attn_pytorch`` (the CPU path of the mmcv op)."""
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, _, num_levels, num_points, _ = sampling_locations.shape
    value_list = value.split([h * w for h, w in spatial_shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    out = None
    for lvl, (h, w) in enumerate(spatial_shapes):
        value_l = value_list[lvl].flatten(2).transpose(1, 2).reshape(bs * num_heads, embed_dims, h, w)
        grid_l = sampling_grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled = F.grid_sample(value_l, grid_l, mode="bilinear", padding_mode="zeros",
                                align_corners=False)  # [bs*heads, c, nq, points]
        w_l = attention_weights[:, :, :, lvl].transpose(1, 2).reshape(bs * num_heads, 1, num_queries, num_points)
        contrib = (sampled * w_l).sum(-1)
        out = contrib if out is None else out + contrib
    return out.view(bs, num_heads * embed_dims, num_queries).transpose(1, 2).contiguous()


class OMSDeformAttn(nn.Module):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=3, num_points=4):
        super().__init__()
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.0)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2)
        grid = grid.repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.view(-1))
        nn.init.constant_(self.attention_weights.weight, 0.0)
        nn.init.constant_(self.attention_weights.bias, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.0)

    def forward(self, query, query_pos, reference_points, spatial_shapes):
        """query/query_pos: [nq, bs, c] (batch_first False); reference_points [bs,nq,levels,2]."""
        identity = query
        value = query
        query = query + query_pos
        query = query.permute(1, 0, 2)
        value = value.permute(1, 0, 2)
        bs, nq, _ = query.shape
        value = self.value_proj(value).view(bs, nq, self.num_heads, -1)
        offs = self.sampling_offsets(query).view(bs, nq, self.num_heads, self.num_levels, self.num_points, 2)
        attw = self.attention_weights(query).view(bs, nq, self.num_heads, self.num_levels * self.num_points)
        attw = attw.softmax(-1).view(bs, nq, self.num_heads, self.num_levels, self.num_points)
        normalizer = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=query.dtype, device=query.device)
        loc = reference_points[:, :, None, :, None, :] + offs / normalizer[None, None, None, :, None, :]
        out = ms_deform_attn_torch(value, spatial_shapes, loc, attw)
        out = self.output_proj(out).permute(1, 0, 2)
        return out + identity


class OEncoderLayer(nn.Module):
    """BaseTransformerLayer, operation_order (self_attn, norm, ffn, norm) -- pairnet.py:40-65."""

    def __init__(self, embed_dims=256, feedforward_channels=1024):
        super().__init__()
        self.attentions = nn.ModuleList([OMSDeformAttn(embed_dims)])
        self.ffns = nn.ModuleList([OFFN(embed_dims, feedforward_channels)])
        self.norms = nn.ModuleList([nn.LayerNorm(embed_dims), nn.LayerNorm(embed_dims)])

    def forward(self, query, query_pos, reference_points, spatial_shapes):
        query = self.attentions[0](query, query_pos, reference_points, spatial_shapes)
        query = self.norms[0](query)
        query = self.ffns[0](query)
        return self.norms[1](query)


class OEncoder(nn.Module):
    def __init__(self, num_layers=6, embed_dims=256, feedforward_channels=1024):
        super().__init__()
        self.layers = nn.ModuleList([OEncoderLayer(embed_dims, feedforward_channels) for _ in range(num_layers)])


class OConvModule(nn.Module):
    """mmcv ``ConvModule`` (conv -> GN -> optional ReLU); parameter names ``conv`` / ``gn``."""

    def __init__(self, cin, cout, k, padding=0, bias=False, act=False, groups=32):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding, bias=bias)
        self.gn = nn.GroupNorm(groups, cout)
        self.act = act

    def forward(self, x):
        x = self.gn(self.conv(x))
        return F.relu(x) if self.act else x


class OMSDeformAttnPixelDecoder(nn.Module):
    """mmdet 2.25.1 ``MSDeformAttnPixelDecoder`` (cfg pairnet.py:33-71): returns
    (mask_feature [B,256,H/4,W/4], [mem_s32, mem_s16, mem_s8])."""

    def __init__(self, in_channels=(256, 512, 1024, 2048), strides=(4, 8, 16, 32), feat_channels=256,
                 out_channels=256, num_outs=3, num_encoder_levels=3, num_encoder_layers=6):
        super().__init__()
        self.strides = list(strides)
        self.num_input_levels = len(in_channels)
        self.num_encoder_levels = num_encoder_levels
        self.num_outs = num_outs
        self.input_convs = nn.ModuleList()
        for i in range(self.num_input_levels - 1, self.num_input_levels - num_encoder_levels - 1, -1):
            self.input_convs.append(OConvModule(in_channels[i], feat_channels, 1, bias=True))
        self.encoder = OEncoder(num_encoder_layers, feat_channels)
        self.level_encoding = nn.Embedding(num_encoder_levels, feat_channels)
        self.lateral_convs = nn.ModuleList()
        self.output_convs = nn.ModuleList()
        for i in range(self.num_input_levels - num_encoder_levels - 1, -1, -1):
            self.lateral_convs.append(OConvModule(in_channels[i], feat_channels, 1, bias=False))
            self.output_convs.append(OConvModule(feat_channels, feat_channels, 3, padding=1, bias=False, act=True))
        self.mask_feature = nn.Conv2d(feat_channels, out_channels, 1)

    def forward(self, feats):
        bs = feats[0].shape[0]
        enc_in, enc_pos, shapes, refs = [], [], [], []
        for i in range(self.num_encoder_levels):
            level_idx = self.num_input_levels - i - 1
            feat = feats[level_idx]
            proj = self.input_convs[i](feat)
            h, w = feat.shape[-2:]
            pos = sine_positional_encoding(torch.zeros((bs, h, w), dtype=torch.bool, device=feat.device),
                                           dtype=feat.dtype)
            lvl_pos = self.level_encoding.weight[i].view(1, -1, 1, 1) + pos
            # MlvlPointGenerator.single_level_grid_priors(offset=0.5), normalised by (w,h)*stride
            ys = (torch.arange(h, dtype=feat.dtype, device=feat.device) + 0.5) * self.strides[level_idx]
            xs = (torch.arange(w, dtype=feat.dtype, device=feat.device) + 0.5) * self.strides[level_idx]
            yy, xx = torch.meshgrid(ys, xs, indexing="ij")
            ref = torch.stack([xx.reshape(-1), yy.reshape(-1)], -1)
            ref = ref / (torch.tensor([[w, h]], dtype=feat.dtype, device=feat.device) * self.strides[level_idx])
            enc_in.append(proj.flatten(2).permute(2, 0, 1))
            enc_pos.append(lvl_pos.flatten(2).permute(2, 0, 1))
            shapes.append((h, w))
            refs.append(ref)
        query = torch.cat(enc_in, 0)
        query_pos = torch.cat(enc_pos, 0)
        reference_points = torch.cat(refs, 0)[None, :, None].repeat(bs, 1, self.num_encoder_levels, 1)
        for layer in self.encoder.layers:
            query = layer(query, query_pos, reference_points, shapes)
        memory = query.permute(1, 2, 0)
        outs = list(torch.split(memory, [h * w for h, w in shapes], dim=-1))
        outs = [x.reshape(bs, -1, shapes[i][0], shapes[i][1]) for i, x in enumerate(outs)]
        for i in range(self.num_input_levels - self.num_encoder_levels - 1, -1, -1):
            cur = self.lateral_convs[i](feats[i])
            y = cur + F.interpolate(outs[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
            outs.append(self.output_convs[i](y))
        return self.mask_feature(outs[-1]), outs[: self.num_outs]
