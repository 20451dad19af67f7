"""Torch-only CPU restatement of ``CrossHead2`` (reference
``pairnet/models/relation_heads/pairnet_head.py:22-417``) and ``PSGTr.forward_dummy``
(``pairnet/models/frameworks/psgtr.py:92-110``).  TEST INFRASTRUCTURE (see
``oracle/__init__.py``); "parity unpinned" except ConvTiny.

The forward below deliberately keeps the reference's op order (full-resolution
``mask_pred`` + ``F.interpolate`` on every one of the 10 head calls, MLPs over all 9
stacked layers, expanded-index ``torch.gather``), so that it is both the checker and
the "reference CPU path" timed by ``bench.py --impl reference``.
"""
import math
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F

from .bricks import (ODetrTransformerDecoder, OMSDeformAttnPixelDecoder,
                     sine_positional_encoding)


class OConvTiny(nn.Module):
    """Matrix-Learner filter; restates ``pairnet/models/frameworks/cnn_factory.py:6-53``
    (3 x Conv2d k7 p3, 1->64->64->1, ReLU between).  PINNED against the reference module by
    ``oracle/make_golden.py`` / ``tests/test_oracle_golden.py``."""

    def __init__(self, mid_channels=64, kernel_size=7):
        super().__init__()
        self.conv_layers = nn.ModuleList([
            nn.Sequential(nn.Conv2d(1, mid_channels, kernel_size, padding=3), nn.ReLU(inplace=True)),
            nn.Sequential(nn.Conv2d(mid_channels, mid_channels, kernel_size, padding=3), nn.ReLU(inplace=True)),
            nn.Sequential(nn.Conv2d(mid_channels, 1, kernel_size, padding=3)),
        ])

    def forward(self, x):
        x = x.unsqueeze(1)
        for layer in self.conv_layers:
            x = layer(x)
        return x.squeeze(1)


def _mlp3(d_in, d_mid, d_out):
    return nn.Sequential(nn.Linear(d_in, d_mid), nn.ReLU(inplace=True), nn.Linear(d_mid, d_mid),
                         nn.ReLU(inplace=True), nn.Linear(d_mid, d_out))


@dataclass
class HeadHyper:
    """Hyper-parameters of ``configs/mask2former/pairnet.py:20-142``."""
    num_classes: int = 133
    num_relations: int = 56
    num_obj_query: int = 100
    num_rel_query: int = 100
    in_channels: tuple = (256, 512, 1024, 2048)
    feat_channels: int = 256
    num_heads: int = 8
    num_levels: int = 3
    dec_layers: int = 9
    rel_layers: int = 6
    ffn_channels: int = 2048
    with_pixel_decoder: bool = True


class OCrossHead2(nn.Module):
    """Same parameter tree / names as the reference head."""

    def __init__(self, hp: HeadHyper = None):
        super().__init__()
        hp = hp or HeadHyper()
        self.hp = hp
        d = hp.feat_channels
        # construction order follows pairnet_head.py:62-176
        self.relation_decoder = ODetrTransformerDecoder(hp.rel_layers, d, hp.num_heads, hp.ffn_channels)
        self.rel_query_embed = nn.Embedding(hp.num_rel_query, d)
        self.rel_query_embed2 = nn.Embedding(hp.num_rel_query * 2, d)
        self.rel_query_embed3 = nn.Embedding(hp.num_rel_query * 2, d)
        self.rel_query_feat = nn.Embedding(hp.num_rel_query, d)
        self.update_importance = OConvTiny()
        if hp.with_pixel_decoder:
            self.pixel_decoder = OMSDeformAttnPixelDecoder(hp.in_channels, feat_channels=d, out_channels=d)
        self.transformer_decoder = ODetrTransformerDecoder(hp.dec_layers, d, hp.num_heads, hp.ffn_channels)
        self.decoder_input_projs = nn.ModuleList([nn.Identity() for _ in range(hp.num_levels)])
        self.query_embed = nn.Embedding(hp.num_obj_query, d)
        self.query_feat = nn.Embedding(hp.num_obj_query, d)
        self.level_embed = nn.Embedding(hp.num_levels, d)
        self.cls_embed = nn.Linear(d, hp.num_classes + 1)
        self.mask_embed = _mlp3(d, d, d)
        self.sub_query_update = _mlp3(d, d, d)
        self.obj_query_update = _mlp3(d, d, d)
        self.rel_cls_embed = nn.Linear(d, hp.num_relations)
        # mmdet 2.25.1 SeesawLoss (rel_cls_loss, configs/mask2former/pairnet.py:153-158) registers the persistent buffer
        # ``cum_samples`` [num_classes + 1]: part of every reference checkpoint.  Registered LAST so that the synthetic
        # fixture weights (drawn in state_dict order, oracle/weights.py) of all other tensors are unchanged.
        self.rel_cls_loss = nn.Module()
        self.rel_cls_loss.register_buffer("cum_samples", torch.zeros(hp.num_relations + 1))
        self.n_heads = hp.num_heads
        self.num_obj_query = hp.num_obj_query
        self.num_rel_query = hp.num_rel_query
        self.embed_dims = d

    def init_weights(self):
        """pairnet_head.py:178-193 (xavier_normal_ on every >=2-D decoder parameter)."""
        for p in self.transformer_decoder.parameters():
            if p.dim() > 1:
                nn.init.xavier_normal_(p)
        for p in self.relation_decoder.parameters():
            if p.dim() > 1:
                nn.init.xavier_normal_(p)

    # pairnet_head.py:216-258
    def forward_head(self, decoder_out, mask_feature, attn_mask_target_size):
        decoder_out = self.transformer_decoder.post_norm(decoder_out)
        decoder_out = decoder_out.transpose(0, 1)
        cls_pred = self.cls_embed(decoder_out)
        mask_embed = self.mask_embed(decoder_out)
        mask_pred = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_feature)
        attn_mask = F.interpolate(mask_pred, attn_mask_target_size, mode="bilinear", align_corners=False)
        attn_mask = attn_mask.flatten(2).unsqueeze(1).repeat((1, self.n_heads, 1, 1)).flatten(0, 1)
        attn_mask = attn_mask.sigmoid() < 0.5
        return cls_pred, mask_pred, attn_mask.detach()

    # pairnet_head.py:260-417 with the pixel decoder factored out
    def forward_from_memories(self, mask_features, multi_scale_memorys, trace=None):
        batch_size = mask_features.shape[0]
        L = self.hp.num_levels
        decoder_inputs, decoder_pos = [], []
        for i in range(L):
            x = self.decoder_input_projs[i](multi_scale_memorys[i])
            x = x.flatten(2).permute(2, 0, 1)
            x = x + self.level_embed.weight[i].view(1, 1, -1)
            mask = x.new_zeros((batch_size,) + multi_scale_memorys[i].shape[-2:], dtype=torch.bool)
            pos = sine_positional_encoding(mask, self.embed_dims // 2, dtype=x.dtype)
            decoder_inputs.append(x)
            decoder_pos.append(pos.flatten(2).permute(2, 0, 1))
        query_feat = self.query_feat.weight.unsqueeze(1).repeat((1, batch_size, 1))
        query_embed = self.query_embed.weight.unsqueeze(1).repeat((1, batch_size, 1))
        query_feat_list = []
        cls_pred, mask_pred, attn_mask = self.forward_head(
            query_feat, mask_features, multi_scale_memorys[0].shape[-2:])
        for i, layer in enumerate(self.transformer_decoder.layers):
            level_idx = i % L
            attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
            if trace is not None:
                trace.setdefault("attn_mask", []).append(attn_mask.view(batch_size, self.n_heads, *attn_mask.shape[1:])[:, 0].clone())
            query_feat = layer(query=query_feat, key=decoder_inputs[level_idx], value=decoder_inputs[level_idx],
                               query_pos=query_embed, key_pos=decoder_pos[level_idx],
                               attn_masks=[attn_mask, None])
            cls_pred, mask_pred, attn_mask = self.forward_head(
                query_feat, mask_features, multi_scale_memorys[(i + 1) % L].shape[-2:])
            query_feat_list.append(query_feat)
            if trace is not None:
                trace.setdefault("query_feat", []).append(query_feat.clone())

        query_feats = torch.stack(query_feat_list)
        sub_embed = self.sub_query_update(query_feats)
        obj_embed = self.obj_query_update(query_feats)
        sub_embed = F.normalize(sub_embed[-1].transpose(0, 1), p=2, dim=-1, eps=1e-12)
        obj_embed = F.normalize(obj_embed[-1].transpose(0, 1), p=2, dim=-1, eps=1e-12)
        importance = torch.matmul(sub_embed, obj_embed.transpose(1, 2))
        if trace is not None:
            trace["sub_embed"], trace["obj_embed"], trace["importance_raw"] = sub_embed, obj_embed, importance
        importance = self.update_importance(importance)
        _, idx = torch.topk(importance.flatten(-2, -1), k=self.num_rel_query)
        sub_pos = torch.div(idx, self.num_obj_query, rounding_mode="trunc")
        obj_pos = torch.remainder(idx, self.num_obj_query)
        obj_query_feat = torch.gather(query_feat, 0, obj_pos.unsqueeze(-1).repeat(1, 1, self.embed_dims).transpose(0, 1))
        sub_query_feat = torch.gather(query_feat, 0, sub_pos.unsqueeze(-1).repeat(1, 1, self.embed_dims).transpose(0, 1))

        rel_query_feat = self.rel_query_feat.weight.unsqueeze(1).repeat((1, batch_size, 1))
        rel_query_embed = self.rel_query_embed.weight.unsqueeze(1).repeat((1, batch_size, 1))
        rel_query_embed2 = self.rel_query_embed2.weight.unsqueeze(1).repeat((1, batch_size, 1))
        rel_query_embed3 = self.rel_query_embed3.weight.unsqueeze(1).repeat((1, batch_size, 1))
        pair_feat = torch.cat([sub_query_feat, obj_query_feat], dim=0)
        for layer in self.relation_decoder.layers:
            rel_query_feat = layer(query=rel_query_feat, key=pair_feat, value=pair_feat,
                                   query_pos=rel_query_embed, key_pos=rel_query_embed2,
                                   value_pos=rel_query_embed3)
            if trace is not None:
                trace.setdefault("rel_feat", []).append(rel_query_feat.clone())
        rel_preds = self.rel_cls_embed(rel_query_feat.transpose(0, 1))

        sub_cls = torch.gather(cls_pred.clone().detach(), 1, sub_pos.unsqueeze(-1).expand(-1, -1, cls_pred.shape[-1]))
        obj_cls = torch.gather(cls_pred.clone().detach(), 1, obj_pos.unsqueeze(-1).expand(-1, -1, cls_pred.shape[-1]))
        sub_seg = torch.gather(mask_pred.clone().detach(), 1,
                               sub_pos[..., None, None].expand(-1, -1, mask_pred.shape[-2], mask_pred.shape[-1]))
        obj_seg = torch.gather(mask_pred.clone().detach(), 1,
                               obj_pos[..., None, None].expand(-1, -1, mask_pred.shape[-2], mask_pred.shape[-1]))
        if trace is not None:
            trace.update(sub_pos=sub_pos, obj_pos=obj_pos, pair_feat=pair_feat, topk_idx=idx)
        all_cls_scores = dict(sub=sub_cls, obj=obj_cls, cls=cls_pred, rel=rel_preds, importance=importance)
        all_mask_preds = dict(mask=mask_pred, sub_seg=sub_seg, obj_seg=obj_seg)
        return all_cls_scores, all_mask_preds

    def forward(self, feats, img_metas=None, trace=None):
        mask_features, memorys = self.pixel_decoder(feats)
        return self.forward_from_memories(mask_features, memorys, trace=trace)


class OPSGTr(nn.Module):
    """``PSGTr`` (psgtr.py:73-110) with a torchvision ResNet-50 standing in for mmdet ``ResNet``
    (same architecture / parameter names; style='pytorch', frozen BN in eval)."""

    def __init__(self, hp: HeadHyper = None):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        del r.fc, r.avgpool
        self.backbone = r
        self.bbox_head = OCrossHead2(hp)

    def extract_feat(self, img):
        b = self.backbone
        x = b.maxpool(b.relu(b.bn1(b.conv1(img))))
        outs = []
        for layer in (b.layer1, b.layer2, b.layer3, b.layer4):
            x = layer(x)
            outs.append(x)
        return tuple(outs)

    def forward_dummy(self, img):
        return self.bbox_head(self.extract_feat(img), [dict()] * img.shape[0])


def stable_topk(values, k):
    """Deterministic top-k contract used for tie tests: descending value, ties by ascending
    flat index.  values: 1-D numpy / tensor.  Returns int64 indices."""
    import numpy as np
    v = np.asarray(values)
    order = np.lexsort((np.arange(v.size), -v.astype(np.float64)))
    return order[:k].astype(np.int64)
