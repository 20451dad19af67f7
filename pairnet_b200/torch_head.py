"""Differentiable evaluation of ``CrossHead2`` for the TRAINING step, PyTorch ops on the device.

The inference forward is the hand-written CUDA library (``head.py`` -> ``pn_head_forward``); it has no backward kernels.
Training needs gradients, so this module evaluates the same graph -- ``CrossHead2.forward``,
``pairnet/models/relation_heads/pairnet_head.py:216-417``, with the mmcv layer semantics of ``facebook_detr.py:311-353``
(attention wrapper) and ``:378-432`` (layer driver) -- with autograd-capable torch ops over the head's OWN parameters
(the weight containers of ``bricks.py``).  Two scopes:

* ``relation``: the Pair-Net-specific parameters (sub/obj MLPs, ConvTiny, Relation Fusion decoder, relation embeddings
  and classifier) train; the Mask2Former decoder runs through the CUDA library without gradients and hands over the
  last layer's queries.  This is the fine-tuning recipe on a frozen Mask2Former.
* ``head``: everything after the pixel decoder is differentiable (what the reference trains, minus the pixel decoder and
  backbone, which stay on the no-grad CUDA / cuDNN path here).

Backward kernels for these stages are future work; until then this is the library (PyTorch) path of training, stated as
such in DESIGN.md.  Dropout follows the modules' ``training`` flag (the relation decoder's FFNs carry ``ffn_drop=0.1``)."""
import math

import torch
import torch.nn.functional as F


def _attention(attn, query, key, value, query_pos, key_pos, attn_mask):
    """mmcv ``MultiheadAttention.forward`` (restated in the reference at facebook_detr.py:311-353): q += query_pos,
    k += key_pos, no position on the value, out = identity + attention (dropouts are 0 in every Pair-Net config)."""
    q = query if query_pos is None else query + query_pos
    k = key if key_pos is None else key + key_pos
    out = F.multi_head_attention_forward(
        q, k, value, attn.embed_dim, attn.num_heads, attn.in_proj_weight, attn.in_proj_bias, None, None, False, 0.0,
        attn.out_proj.weight, attn.out_proj.bias, training=False, attn_mask=attn_mask, need_weights=False)[0]
    return query + out


def decoder_layer(layer, query, key, value, query_pos, key_pos, attn_mask=None):
    """mmcv ``BaseTransformerLayer.forward`` for (cross_attn, norm, self_attn, norm, ffn, norm) (facebook_detr.py:378-432)."""
    q = layer.norms[0](_attention(layer.attentions[0].attn, query, key, value, query_pos, key_pos, attn_mask))
    q = layer.norms[1](_attention(layer.attentions[1].attn, q, q, q, query_pos, query_pos, None))
    ffn = layer.ffns[0].layers
    h = ffn[0][2](F.relu(ffn[0][0](q)))          # Linear, ReLU, Dropout(ffn_drop)
    q = q + ffn[2](ffn[1](h))                    # Linear, Dropout(ffn_drop); add_identity
    return layer.norms[2](q)


def sine_positional_encoding(h, w, batch, num_feats, device, dtype, temperature=10000, scale=2 * math.pi, eps=1e-6):
    """mmdet ``SinePositionalEncoding(normalize=True)`` of an all-zero mask -> [h*w, batch, 2*num_feats]."""
    not_mask = torch.ones((batch, h, w), device=device, dtype=dtype)
    y = not_mask.cumsum(1)
    x = not_mask.cumsum(2)
    y = y / (y[:, -1:, :] + eps) * scale
    x = x / (x[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=dtype, device=device)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).view(batch, h, w, -1)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).view(batch, h, w, -1)
    return torch.cat((py, px), dim=3).flatten(1, 2).transpose(0, 1)


def forward_head(head, decoder_out, mask_feature, size):
    """pairnet_head.py:216-258."""
    decoder_out = head.transformer_decoder.post_norm(decoder_out).transpose(0, 1)
    cls_pred = head.cls_embed(decoder_out)
    mask_pred = torch.einsum("bqc,bchw->bqhw", head.mask_embed(decoder_out), mask_feature)
    attn_mask = F.interpolate(mask_pred, size, mode="bilinear", align_corners=False)
    attn_mask = attn_mask.flatten(2).unsqueeze(1).repeat((1, head.n_heads, 1, 1)).flatten(0, 1)
    return cls_pred, mask_pred, (attn_mask.sigmoid() < 0.5).detach()


def masked_decoder(head, mask_features, memorys):
    """pairnet_head.py:262-320 -> (query_feat [N,B,256] of the last layer, cls_pred, mask_pred)."""
    B = mask_features.shape[0]
    L = head.num_transformer_feat_level
    inputs, poss = [], []
    for i in range(L):
        m = memorys[i]
        inputs.append(m.flatten(2).permute(2, 0, 1) + head.level_embed.weight[i].view(1, 1, -1))
        poss.append(sine_positional_encoding(m.shape[2], m.shape[3], B, head.embed_dims // 2, m.device, m.dtype))
    query_feat = head.query_feat.weight.unsqueeze(1).repeat((1, B, 1))
    query_embed = head.query_embed.weight.unsqueeze(1).repeat((1, B, 1))
    cls_pred, mask_pred, attn_mask = forward_head(head, query_feat, mask_features, memorys[0].shape[-2:])
    for i, layer in enumerate(head.transformer_decoder.layers):
        l = i % L
        attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
        query_feat = decoder_layer(layer, query_feat, inputs[l], inputs[l], query_embed, poss[l], attn_mask)
        cls_pred, mask_pred, attn_mask = forward_head(head, query_feat, mask_features, memorys[(i + 1) % L].shape[-2:])
    return query_feat, cls_pred, mask_pred


def relation_side(head, query_feat, cls_pred, mask_pred=None):
    """pairnet_head.py:322-417 from the last layer's queries [N,B,256]: PPN -> top-k -> Relation Fusion -> outputs."""
    N, B, d = query_feat.shape
    R = head.num_rel_query
    sub_embed = F.normalize(head.sub_query_update(query_feat).transpose(0, 1), p=2, dim=-1, eps=1e-12)
    obj_embed = F.normalize(head.obj_query_update(query_feat).transpose(0, 1), p=2, dim=-1, eps=1e-12)
    importance = torch.matmul(sub_embed, obj_embed.transpose(1, 2)).unsqueeze(1)
    for seq in head.update_importance.conv_layers:           # ConvTiny (cnn_factory.py:49-53)
        conv = seq[0]
        importance = F.conv2d(importance, conv.weight, conv.bias, padding=conv.padding)
        if len(seq) > 1:
            importance = F.relu(importance)
    importance = importance.squeeze(1)
    idx = torch.topk(importance.flatten(-2, -1), k=R).indices
    sub_pos = torch.div(idx, N, rounding_mode="trunc")
    obj_pos = torch.remainder(idx, N)
    gat = lambda pos: torch.gather(query_feat, 0, pos.unsqueeze(-1).repeat(1, 1, d).transpose(0, 1))
    pair_feat = torch.cat([gat(sub_pos), gat(obj_pos)], dim=0)
    rel = head.rel_query_feat.weight.unsqueeze(1).repeat((1, B, 1))
    e1 = head.rel_query_embed.weight.unsqueeze(1).repeat((1, B, 1))
    e2 = head.rel_query_embed2.weight.unsqueeze(1).repeat((1, B, 1))
    for layer in head.relation_decoder.layers:
        rel = decoder_layer(layer, rel, pair_feat, pair_feat, e1, e2)
    rel_preds = head.rel_cls_embed(rel.transpose(0, 1))
    cls_d = cls_pred.detach()
    sub_cls = torch.gather(cls_d, 1, sub_pos.unsqueeze(-1).expand(-1, -1, cls_d.shape[-1]))
    obj_cls = torch.gather(cls_d, 1, obj_pos.unsqueeze(-1).expand(-1, -1, cls_d.shape[-1]))
    all_cls_scores = dict(sub=sub_cls, obj=obj_cls, cls=cls_pred, rel=rel_preds, importance=importance)
    all_mask_preds = dict(mask=mask_pred)
    return all_cls_scores, all_mask_preds, (sub_pos, obj_pos)


RELATION_PARAMS = ("sub_query_update", "obj_query_update", "update_importance", "relation_decoder.layers",
                   "rel_query_feat", "rel_query_embed.", "rel_query_embed2", "rel_cls_embed")


def trainable_parameters(head, scope):
    """Parameters that receive gradients in ``scope`` (static set: no ``find_unused_parameters``).  Never trained by the
    reference's loss either: ``rel_query_embed3`` (its value_pos is swallowed), ``relation_decoder.post_norm`` (never
    applied), and under ``relation`` scope everything upstream of the last decoder layer's output."""
    out = []
    for n, p in head.named_parameters():
        if n.startswith("pixel_decoder") or n.startswith("rel_query_embed3") or n.startswith("relation_decoder.post_norm"):
            continue
        if scope == "relation" and not any(n.startswith(k) for k in RELATION_PARAMS):
            continue
        if n.startswith("cls_embed") or n.startswith("mask_embed") or n.startswith("transformer_decoder.post_norm"):
            continue  # reach the loss through detached tensors only (attention masks, sub/obj logits, matching costs)
        out.append((n, p))
    return out
