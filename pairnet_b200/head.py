"""``CrossHead2`` -- the Pair-Net relation head, B200-native.

Same constructor kwargs, parameter names, ``forward(feats, img_metas)`` contract and output dicts as the
reference head (``pairnet/models/relation_heads/pairnet_head.py:22-417``), but the forward
is ONE call into ``libpairnet_b200.so`` (``pn_head_forward``): Mask2Former masked-attention decoder ->
Pair Proposal Network (sub/obj MLP, L2 norm, N x N pair matrix, ConvTiny, top-k, pair gather) ->
Relation Fusion decoder -> output gathers.  There is no PyTorch/CPU fallback for that path.
"""
import copy
import ctypes as C

import torch
import torch.nn as nn

from . import _native as nat
from .registry import (HEADS, ConfigDict, build_loss, build_plugin_layer, build_positional_encoding,
                       build_transformer_layer_sequence, to_config)
from .training import TrainMixin


class ConvTiny(nn.Module):
    """Weights of the Matrix-Learner filter (reference ``cnn_factory.py:6-53``; names
    ``conv_layers.{0,1,2}.0``).  Evaluated by ``pn_conv_tiny``."""

    def __init__(self, in_channels=1, out_channels=1, kernel_size=7, mid_channels=64, layers=3):
        super().__init__()
        if (in_channels, out_channels, kernel_size, layers) != (1, 1, 7, 3):
            raise NotImplementedError("only the shipped ConvTiny geometry (1->mid->mid->1, k=7) is supported")
        self.mid_channels = mid_channels
        self.conv_layers = nn.ModuleList([
            nn.Sequential(nn.Conv2d(1, mid_channels, 7, padding=3), nn.ReLU(inplace=True)),
            nn.Sequential(nn.Conv2d(mid_channels, mid_channels, 7, padding=3), nn.ReLU(inplace=True)),
            nn.Sequential(nn.Conv2d(mid_channels, 1, 7, padding=3)),
        ])

    def forward(self, x):
        from . import ops
        return ops.conv_tiny(x, self)


def creat_cnn(name):
    """reference ``cnn_factory.creat_cnn`` (sic)."""
    if name == "conv_tiny":
        return ConvTiny()
    raise NotImplementedError(f"mapper '{name}': only 'conv_tiny' is used by the shipped Pair-Net configs "
                              "(conv_small hard-codes N=100, conv_base is a U-Net ablation)")


def _mlp3(d):
    return nn.Sequential(nn.Linear(d, d), nn.ReLU(inplace=True), nn.Linear(d, d), nn.ReLU(inplace=True), nn.Linear(d, d))


def _ptr(t):
    return t.data_ptr()


INSTANCE_OFFSET = 1000  # mmdet.datasets.coco_panoptic.INSTANCE_OFFSET (pairnet_head.py:16)


def attach_native(module):
    """INTEGRATION.md Option B: give ANY module that carries the reference ``CrossHead2``'s attribute names (the
    reference class itself inside an mmdet environment) the B200 forward.  Binds the native plumbing of this file's
    ``CrossHead2`` to ``module`` and creates the four cache slots; afterwards
    ``module.forward_from_memories(mask_features, multi_scale_memorys)`` is the hot path."""
    import types
    for name in ("_hot_params", "_mlp", "_layer", "native_weights", "_pos_table", "forward_from_memories"):
        setattr(module, name, types.MethodType(getattr(CrossHead2, name), module))
    module._lin, module._norm = CrossHead2._lin, CrossHead2._norm      # static helpers
    module._wkey, module._wstruct, module._pos_cache, module._ws = None, None, {}, None
    return module


@HEADS.register_module()
class CrossHead2(TrainMixin, nn.Module):
    def __init__(self, num_classes, in_channels, num_relations, num_obj_query=100, num_rel_query=100,
                 mapper="conv_tiny", use_mask=True, pixel_decoder=None, transformer_decoder=None, feat_channels=256,
                 out_channels=256, num_transformer_feat_level=3, embed_dims=256, relation_decoder=None,
                 enforce_decoder_input_project=False, n_heads=8,
                 positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True),
                 rel_cls_loss=None, subobj_cls_loss=None, importance_match_loss=None, loss_cls=None, loss_mask=None,
                 loss_dice=None, train_cfg=None, test_cfg=dict(max_per_img=100), init_cfg=None, **kwargs):
        super().__init__()
        transformer_decoder = to_config(transformer_decoder)
        relation_decoder = to_config(relation_decoder)
        pixel_decoder = to_config(pixel_decoder)
        positional_encoding = to_config(positional_encoding)
        if embed_dims != nat.EMBED_DIMS or feat_channels != nat.EMBED_DIMS or out_channels != nat.EMBED_DIMS:
            raise NotImplementedError("the CUDA library is compiled for embed_dims = feat_channels = out_channels = 256")
        if enforce_decoder_input_project:
            raise NotImplementedError("enforce_decoder_input_project=True is not used by the Pair-Net configs")
        self.num_classes = num_classes
        self.num_rel_query = num_rel_query
        self.num_relations = num_relations
        self.use_mask = use_mask
        # construction order follows the reference (pairnet_head.py:62-176)
        self.relation_decoder = build_transformer_layer_sequence(relation_decoder)
        self.rel_query_embed = nn.Embedding(num_rel_query, feat_channels)
        self.rel_query_embed2 = nn.Embedding(num_rel_query * 2, feat_channels)
        self.rel_query_embed3 = nn.Embedding(num_rel_query * 2, feat_channels)  # dead in the reference forward
        self.rel_query_feat = nn.Embedding(num_rel_query, feat_channels)
        self.update_importance = creat_cnn(mapper)
        self.n_heads = n_heads
        self.embed_dims = embed_dims
        assert "num_feats" in positional_encoding
        assert positional_encoding["num_feats"] * 2 == embed_dims, (
            f"embed_dims should be exactly 2 times of num_feats. Found {embed_dims} and {positional_encoding['num_feats']}.")
        self.num_queries = num_obj_query
        self.num_transformer_feat_level = num_transformer_feat_level
        self.num_heads = transformer_decoder.transformerlayers.attn_cfgs.num_heads
        if self.num_heads != 8 or n_heads != 8:
            raise NotImplementedError("the CUDA library is compiled for 8 heads x 32")
        # everything else the kernels hard-code is rejected here rather than silently mis-computed
        for name, dec in (("transformer_decoder", transformer_decoder), ("relation_decoder", relation_decoder)):
            tl = dec.transformerlayers
            if tl.attn_cfgs.num_heads != 8 or tl.attn_cfgs.embed_dims != nat.EMBED_DIMS:
                raise NotImplementedError(f"{name}: the CUDA library is compiled for 8 heads x 32 (embed_dims 256)")
            if tuple(tl.operation_order) != ("cross_attn", "norm", "self_attn", "norm", "ffn", "norm"):
                raise NotImplementedError(f"{name}: only operation_order (cross_attn, norm, self_attn, norm, ffn, norm)")
            if tl.ffn_cfgs.get("act_cfg", dict(type="ReLU")).get("type") != "ReLU":
                raise NotImplementedError(f"{name}: only ReLU FFNs")
        pe = positional_encoding
        if (pe.get("type") != "SinePositionalEncoding" or not pe.get("normalize", False)
                or pe.get("temperature", 10000) != 10000 or abs(pe.get("scale", 2 * 3.141592653589793) - 2 * 3.141592653589793) > 1e-9
                or pe.get("offset", 0.0) != 0.0 or pe.get("eps", 1e-6) != 1e-6):
            raise NotImplementedError("positional_encoding: the CUDA kernel implements SinePositionalEncoding(normalize=True, "
                                      "temperature=10000, scale=2*pi, offset=0, eps=1e-6) only")
        self.with_pixel_decoder = pixel_decoder is not None
        if self.with_pixel_decoder:
            assert pixel_decoder.encoder.transformerlayers.attn_cfgs.num_levels == num_transformer_feat_level
            pd = copy.deepcopy(pixel_decoder)
            pd.update(in_channels=in_channels, feat_channels=feat_channels, out_channels=out_channels)
            self.pixel_decoder = build_plugin_layer(pd)[1]
        self.transformer_decoder = build_transformer_layer_sequence(transformer_decoder)
        self.decoder_embed_dims = self.transformer_decoder.embed_dims
        assert self.decoder_embed_dims == feat_channels
        self.decoder_input_projs = nn.ModuleList([nn.Identity() for _ in range(num_transformer_feat_level)])
        self.decoder_positional_encoding = build_positional_encoding(positional_encoding)
        self.query_embed = nn.Embedding(num_obj_query, feat_channels)
        self.query_feat = nn.Embedding(num_obj_query, feat_channels)
        self.level_embed = nn.Embedding(num_transformer_feat_level, feat_channels)
        self.cls_embed = nn.Linear(feat_channels, num_classes + 1)
        self.mask_embed = _mlp3(feat_channels)
        self.test_cfg = test_cfg
        self.train_cfg = train_cfg
        self._init_train_cfg(train_cfg)   # assigners / sampler / num_points (pairnet_head.py:129-137)
        self.train_scope = "head"         # which parameters receive gradients in forward_train (torch_head.py)
        self.num_obj_query = num_obj_query
        self.in_channels = in_channels
        self.loss_cfgs = dict(loss_cls=loss_cls, loss_mask=loss_mask, loss_dice=loss_dice, rel_cls_loss=rel_cls_loss,
                              subobj_cls_loss=subobj_cls_loss, importance_match_loss=importance_match_loss)
        for name, cfg in self.loss_cfgs.items():
            setattr(self, name, build_loss(cfg) if cfg is not None else None)
        self.class_weight = loss_cls.get("class_weight") if loss_cls else None
        self.cls_out_channels = num_classes if (loss_cls and loss_cls.get("use_sigmoid")) else num_classes + 1
        self.sub_query_update = _mlp3(embed_dims)
        self.obj_query_update = _mlp3(embed_dims)
        self.rel_cls_embed = nn.Linear(embed_dims, num_relations)
        # native-side caches
        self._wkey = None
        self._wstruct = None
        self._pos_cache = {}
        self._ws = None

    # ------------------------------------------------------------------ reference API
    def init_weights(self):
        """pairnet_head.py:178-193."""
        if self.with_pixel_decoder:
            self.pixel_decoder.init_weights()
        for p in self.transformer_decoder.parameters():
            if p.dim() > 1:
                nn.init.xavier_normal_(p)
        for p in self.relation_decoder.parameters():
            if p.dim() > 1:
                nn.init.xavier_normal_(p)

    def forward(self, feats, img_metas=None):
        """feats: 4 backbone maps [B,C_i,H_i,W_i] -> (all_cls_scores, all_mask_preds)  (pairnet_head.py:260-417)."""
        mask_features, memorys = self.pixel_decoder(feats)
        return self.forward_from_memories(mask_features, memorys)

    def train_outputs(self, feats, img_metas=None, scope=None):
        """Head outputs WITH autograd history for the training step.  The pixel decoder (and the backbone before it) run
        on the no-grad CUDA / cuDNN path; ``scope="relation"`` additionally runs the Mask2Former decoder through the CUDA
        library (no gradients) and differentiates only the Pair-Net side; ``scope="head"`` differentiates everything
        after the pixel decoder (PyTorch ops on the device, ``torch_head.py``)."""
        from . import torch_head as th
        scope = scope or self.train_scope
        # the frozen, no-grad part runs on the CUDA library (fp32 tensors): an enclosing autocast (bf16 training) must not
        # re-type its PyTorch plumbing
        with torch.no_grad(), torch.autocast(feats[0].device.type, enabled=False):
            mask_features, memorys = self.pixel_decoder(feats)
        if scope == "relation":
            taps = {}
            with torch.autocast(feats[0].device.type, enabled=False):
                cls_scores, mask_preds = self.forward_from_memories(mask_features, memorys, taps=taps,
                                                                    materialize_seg=False)
            query_feat = taps["query_out"].transpose(0, 1).contiguous()      # [N,B,256], last decoder layer
            return th.relation_side(self, query_feat, cls_scores["cls"], mask_preds["mask"])[:2]
        if scope != "head":
            raise ValueError(f"train scope {scope!r}: 'relation' or 'head'")
        mask_features = mask_features.float().contiguous()
        query_feat, cls_pred, mask_pred = th.masked_decoder(self, mask_features, [m.float().contiguous() for m in memorys])
        return th.relation_side(self, query_feat, cls_pred, mask_pred)[:2]

    def forward_train(self, x, img_metas, gt_rels, gt_bboxes, gt_labels=None, gt_masks=None, gt_bboxes_ignore=None,
                      proposal_cfg=None, **kwargs):
        """pairnet_head.py:720-757 -> dict(loss_r_cls, loss_sub_cls, loss_obj_cls, loss_match)."""
        assert proposal_cfg is None, '"proposal_cfg" must be None'
        if self.train_cfg is None:
            raise RuntimeError("CrossHead2 was built without train_cfg (assigners / sampler): cannot train")
        outs = self.train_outputs(x, img_metas)
        if gt_labels is None:
            loss_inputs = outs + (gt_rels, gt_bboxes, gt_masks, img_metas)
        else:
            loss_inputs = outs + (gt_rels, gt_bboxes, gt_labels, gt_masks, img_metas)
        return self.loss(*loss_inputs, gt_bboxes_ignore=gt_bboxes_ignore)

    # ------------------------------------------------------------------ inference post-processing (pairnet_head.py:759-930)
    def simple_test_bboxes(self, feats, img_metas, rescale=False):
        """pairnet_head.py:926-930."""
        outs = self.forward(feats, img_metas)
        return self.get_bboxes(*outs, img_metas, rescale=rescale)

    simple_test = simple_test_bboxes

    def get_bboxes(self, cls_scores, mask_preds, img_metas, rescale=False):
        """pairnet_head.py:759-786: per image -> (det_bboxes[2K,5], labels[2K], rel_pairs[K,2] int32, masks[2K,H,W] bool,
        pan_img[H,W] long (CPU), r_scores[K], r_labels[K], r_dists[K,num_relations+1])."""
        return [self._get_bboxes_single(mask_preds["mask"][i], cls_scores["cls"][i], cls_scores["sub"][i],
                                        cls_scores["obj"][i], cls_scores["rel"][i], mask_preds["sub_seg"][i],
                                        mask_preds["obj_seg"][i], img_metas[i]["img_shape"],
                                        img_metas[i]["scale_factor"], rescale)
                for i in range(len(img_metas))]

    @staticmethod
    def _stuff_remap(labels, first_stuff_label=80):
        """pairnet_head.py:858-861 + :877-882: kept segments of the same stuff class (label >= 80) are merged into the
        first of them.  Returns ``remap`` with ``remap[k]`` = position (in the kept list) segment ``k`` is painted as."""
        first, remap = {}, []
        for k, lab in enumerate(labels):
            remap.append(first.setdefault(lab, k) if lab >= first_stuff_label else k)
        return remap

    @torch.no_grad()
    def _get_bboxes_single(self, all_masks, all_cls_score, s_cls_score, o_cls_score, r_cls_score, s_mask_pred,
                           o_mask_pred, img_shape, scale_factor, rescale=False):
        """pairnet_head.py:788-924.  The class / relation softmaxes are [K,134]-sized torch plumbing; the three
        full-image upsamples, the thresholds, the panoptic argmax and the per-segment areas run in
        ``pn_upsample_threshold`` / ``pn_panoptic_merge`` straight from the quarter-resolution logits."""
        from torch.nn import functional as F
        lib = nat.load()
        dev = all_masks.device
        if dev.type != "cuda":
            raise nat.NativeError("CrossHead2.get_bboxes needs CUDA tensors on a B200; there is no CPU fallback")
        stream = torch.cuda.current_stream(dev).cuda_stream
        assert len(s_cls_score) == len(o_cls_score) == len(r_cls_score)
        H, W = round(img_shape[0] / float(scale_factor[1])), round(img_shape[1] / float(scale_factor[0]))
        s_logits = F.softmax(s_cls_score, dim=-1)[..., :-1]
        o_logits = F.softmax(o_cls_score, dim=-1)[..., :-1]
        s_labels, o_labels = s_logits.argmax(-1) + 1, o_logits.argmax(-1) + 1
        r_dists = F.softmax(r_cls_score, dim=-1).reshape(-1, self.num_relations)
        r_dists = torch.cat([torch.zeros(self.num_rel_query, 1, device=dev), r_dists], dim=-1)
        complete_labels = torch.cat((s_labels, o_labels), 0)
        all_scores, all_labels = F.softmax(all_cls_score, dim=-1)[..., :-1].max(-1)
        all_masks = all_masks.float().contiguous()
        N, h, w = all_masks.shape
        K = s_cls_score.shape[0]
        masks = torch.empty((2 * K, H, W), dtype=torch.bool, device=dev)
        for half, src in enumerate((s_mask_pred, o_mask_pred)):
            src = src.float().contiguous()
            assert src.shape == (K, h, w)
            nat.check(lib.pn_upsample_threshold(src.data_ptr(), None, masks[half * K:].data_ptr(), K, K, h, w, H, W,
                                                stream), "pn_upsample_threshold")
        keep = (all_labels != s_logits.shape[-1] - 1) & (all_scores > 0.5)
        keep_idx = keep.nonzero().flatten().to(torch.int32)
        labels_k = all_labels[keep].contiguous()
        if keep_idx.numel() == 0:
            pan_img = torch.ones((H, W)).to(torch.long)
        else:
            pan = torch.empty((H, W), dtype=torch.int64, device=dev)

            def merge(keep_idx, labels_k, remap):
                n = keep_idx.numel()
                if n == 0:  # the reference indexes an empty label list here
                    raise IndexError("every kept mask was filtered as small (reference: all_labels[m_id] on an empty list)")
                remap_t = torch.tensor(remap, dtype=torch.int32, device=dev)
                area = torch.empty(n, dtype=torch.int32, device=dev)
                nat.check(lib.pn_panoptic_merge(all_masks.data_ptr(), keep_idx.data_ptr(), remap_t.data_ptr(),
                                                labels_k.data_ptr(), n, h, w, H, W, INSTANCE_OFFSET, pan.data_ptr(),
                                                area.data_ptr(), stream), "pn_panoptic_merge")
                return area.tolist()  # one small D2H per pass (the reference does one .item() per mask)

            area = merge(keep_idx, labels_k, self._stuff_remap(labels_k.tolist()))
            while True:  # drop segments of <= 4 pixels and re-run the argmax without them (:896-908)
                small = torch.tensor([a <= 4 for a in area], dtype=torch.bool, device=dev)
                if not bool(small.any()):
                    break
                keep_idx, labels_k = keep_idx[~small].contiguous(), labels_k[~small].contiguous()
                area = merge(keep_idx, labels_k, list(range(keep_idx.numel())))
            pan_img = pan.cpu()
        det_bboxes = torch.zeros((self.num_rel_query * 2, 5), device=dev)
        r_scores = torch.zeros(self.num_rel_query, device=dev)
        r_labels = torch.zeros(self.num_rel_query, device=dev)
        rel_pairs = torch.arange(len(det_bboxes), dtype=torch.int).reshape(2, -1).T
        return det_bboxes, complete_labels, rel_pairs, masks, pan_img, r_scores, r_labels, r_dists

    # ------------------------------------------------------------------ native plumbing
    def _hot_params(self):
        ps = [self.query_feat.weight, self.query_embed.weight, self.level_embed.weight, self.cls_embed.weight,
              self.rel_query_feat.weight, self.rel_query_embed.weight, self.rel_query_embed2.weight]
        for m in (self.transformer_decoder, self.relation_decoder, self.mask_embed, self.sub_query_update,
                  self.obj_query_update, self.update_importance, self.rel_cls_embed, self.cls_embed):
            ps.extend(m.parameters())
        return ps

    @staticmethod
    def _lin(dst, mod):
        dst.w = _ptr(mod.weight)
        dst.b = _ptr(mod.bias) if mod.bias is not None else None

    @staticmethod
    def _norm(dst, mod):
        dst.gamma, dst.beta = _ptr(mod.weight), _ptr(mod.bias)

    def _mlp(self, dst, seq):
        for i, j in enumerate((0, 2, 4)):
            self._lin(dst.l[i], seq[j])

    def _layer(self, dst, layer):
        for d, a in ((dst.cross_attn, layer.attentions[0].attn), (dst.self_attn, layer.attentions[1].attn)):
            d.in_proj_w, d.in_proj_b = _ptr(a.in_proj_weight), _ptr(a.in_proj_bias)
            d.out_proj_w, d.out_proj_b = _ptr(a.out_proj.weight), _ptr(a.out_proj.bias)
        self._lin(dst.ffn1, layer.ffns[0].layers[0][0])
        self._lin(dst.ffn2, layer.ffns[0].layers[1])
        for i in range(3):
            self._norm(dst.norm[i], layer.norms[i])

    def native_weights(self):
        """``PnHeadWeights`` over the current parameter storage (rebuilt when any pointer changes)."""
        params = self._hot_params()
        key = tuple((p.data_ptr(), p._version) for p in params)   # in-place updates (optimizer steps) invalidate the prepared splits
        if key == self._wkey:
            return self._wstruct
        for p in params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise nat.NativeError("CrossHead2 hot-path parameters must be contiguous fp32 CUDA tensors "
                                      "(move the head to a B200 with .cuda(); there is no CPU path)")
        w = nat.PnHeadWeights()
        m = w.m2f
        td = self.transformer_decoder
        m.num_queries, m.num_layers, m.num_levels = self.num_queries, len(td.layers), self.num_transformer_feat_level
        m.ffn_dims, m.num_cls = td.layers[0].ffns[0].feedforward_channels, self.num_classes + 1
        if len(td.layers) > nat.PN_MAX_LAYERS or len(self.relation_decoder.layers) > nat.PN_MAX_LAYERS:
            raise NotImplementedError("too many decoder layers for PN_MAX_LAYERS")
        m.query_feat, m.query_embed, m.level_embed = _ptr(self.query_feat.weight), _ptr(self.query_embed.weight), _ptr(self.level_embed.weight)
        self._norm(m.post_norm, td.post_norm)
        self._lin(m.cls_embed, self.cls_embed)
        self._mlp(m.mask_embed, self.mask_embed)
        for i, layer in enumerate(td.layers):
            self._layer(m.layers[i], layer)
        self._mlp(w.sub_query_update, self.sub_query_update)
        self._mlp(w.obj_query_update, self.obj_query_update)
        cv = w.update_importance
        cv.mid_channels = self.update_importance.conv_layers[0][0].out_channels
        for i in range(3):
            conv = self.update_importance.conv_layers[i][0]
            cv.w[i], cv.b[i] = _ptr(conv.weight), _ptr(conv.bias)
        r = w.rel
        rd = self.relation_decoder
        r.num_rel_queries, r.num_layers = self.num_rel_query, len(rd.layers)
        r.ffn_dims, r.num_rel_cls = rd.layers[0].ffns[0].feedforward_channels, self.num_relations
        r.rel_query_feat, r.rel_query_embed = _ptr(self.rel_query_feat.weight), _ptr(self.rel_query_embed.weight)
        r.rel_query_embed2 = _ptr(self.rel_query_embed2.weight)
        self._lin(r.rel_cls_embed, self.rel_cls_embed)
        for i, layer in enumerate(rd.layers):
            self._layer(r.layers[i], layer)
        # static operands of the tensor-core kernels (TF32 hi/lo weight splits): built once per weight version
        lib = nat.load()
        dev = params[0].device
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._prepared = []
        for sub, size_fn, prep_fn in ((m, lib.pn_m2f_prepared_bytes, lib.pn_m2f_prepare),
                                      (r, lib.pn_rel_prepared_bytes, lib.pn_rel_prepare)):
            need = size_fn(C.byref(sub))
            if need == 0:
                raise nat.NativeError("prepared-weights size query failed: " + lib.pn_last_error_string().decode())
            buf = torch.empty(need, dtype=torch.uint8, device=dev)
            nat.check(prep_fn(C.byref(sub), buf.data_ptr(), need, stream), "pn_*_prepare")
            sub.prepared = buf.data_ptr()
            self._prepared.append(buf)
        self._wkey, self._wstruct = key, w
        return w

    def _pos_table(self, h, w, device, stream):
        key = (h, w, str(device))
        if key not in self._pos_cache:
            t = torch.empty((h * w, nat.EMBED_DIMS), dtype=torch.float32, device=device)
            nat.check(nat.load().pn_sine_posenc(t.data_ptr(), h, w, stream), "pn_sine_posenc")
            self._pos_cache[key] = t
        return self._pos_cache[key]

    @torch.no_grad()
    def forward_from_memories(self, mask_features, memorys, taps=None, materialize_seg=True):
        """The hot path proper: everything in ``CrossHead2.forward`` after the pixel decoder.

        taps: optional dict; when given, per-stage intermediates are written into it (parity tests)."""
        lib = nat.load()
        dev = mask_features.device
        if dev.type != "cuda":
            raise nat.NativeError("CrossHead2.forward needs CUDA tensors on a B200; there is no CPU fallback")
        stream = torch.cuda.current_stream(dev).cuda_stream
        mask_features = mask_features.float()
        memorys = [m.float() for m in memorys]
        B, Cc, H4, W4 = mask_features.shape
        # a channels_last mask_features map is consumed in place as token-major [B,H4*W4,256] by the tensor-core mask path
        mf_tokens = (not mask_features.is_contiguous() and mask_features.stride() == (H4 * W4 * Cc, 1, W4 * Cc, Cc)
                     and mask_features.data_ptr() % 16 == 0 and lib.pn_get_option(nat.PN_OPT_TENSOR_CORES) != 0
                     and lib.pn_get_option(nat.PN_OPT_MASK_TC) != 0)
        if not mf_tokens:
            mask_features = mask_features.contiguous()
        assert Cc == nat.EMBED_DIMS and len(memorys) == self.num_transformer_feat_level
        N, R, K = self.num_queries, self.num_rel_query, self.num_rel_query
        ncls, nrel = self.num_classes + 1, self.num_relations
        w = self.native_weights()
        inp = nat.PnM2FInputs()
        inp.B, inp.H4, inp.W4, inp.mask_features = B, H4, W4, mask_features.data_ptr()
        inp.mask_features_token_major = int(mf_tokens)
        keep = []
        for l, m in enumerate(memorys):
            # the pixel decoder hands out NCHW *views* of its token-major encoder output ([B,nq,256] sliced per
            # level): consume them in place instead of materialising an NCHW copy that level prep would transpose back
            hl, wl = m.shape[2], m.shape[3]
            if (not m.is_contiguous() and m.stride(1) == 1 and m.stride(3) == Cc and m.stride(2) == wl * Cc
                    and m.stride(0) >= hl * wl * Cc and m.stride(0) % 4 == 0 and m.data_ptr() % 16 == 0):
                inp.memory_token_major[l], inp.memory_batch_stride[l] = 1, m.stride(0)
            else:
                m = m.contiguous()
            keep.append(m)
            inp.h[l], inp.w[l], inp.memory[l] = hl, wl, m.data_ptr()
            pt = self._pos_table(m.shape[2], m.shape[3], dev, stream)
            inp.pos[l] = pt.data_ptr()
            keep.append(pt)
        f32 = dict(dtype=torch.float32, device=dev)
        o = dict(cls=torch.empty((B, N, ncls), **f32), mask=torch.empty((B, N, H4, W4), **f32),
                 importance=torch.empty((B, N, N), **f32), rel=torch.empty((B, R, nrel), **f32),
                 sub_pos=torch.empty((B, K), dtype=torch.int64, device=dev),
                 obj_pos=torch.empty((B, K), dtype=torch.int64, device=dev),
                 sub=torch.empty((B, K, ncls), **f32), obj=torch.empty((B, K, ncls), **f32))
        if materialize_seg:
            o["sub_seg"] = torch.empty((B, K, H4, W4), **f32)
            o["obj_seg"] = torch.empty((B, K, H4, W4), **f32)
        out = nat.PnHeadOutputs()
        for k, t in o.items():
            setattr(out, k, t.data_ptr())
        if taps is not None:
            nl = len(self.transformer_decoder.layers)
            words = max((m.shape[2] * m.shape[3] + 63) // 64 * 2 for m in memorys)
            t = dict(query_out=torch.empty((B, N, 256), **f32), importance_raw=torch.empty((B, N, N), **f32),
                     pair_feat=torch.empty((B, 2 * K, 256), **f32), rel_feat=torch.empty((B, R, 256), **f32),
                     query_trace=torch.empty((nl, B, N, 256), **f32),
                     mask_trace=torch.zeros((nl, B, N, words), dtype=torch.int32, device=dev))
            for k, v in t.items():
                setattr(out, k, v.data_ptr())
            out.trace_words = words
            taps.update(t)
        need = lib.pn_head_workspace_bytes(C.byref(w), C.byref(inp))
        if need == 0:
            raise nat.NativeError("pn_head_workspace_bytes: " + lib.pn_last_error_string().decode())
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        nat.check(lib.pn_head_forward(C.byref(w), C.byref(inp), C.byref(out), self._ws.data_ptr(), self._ws.numel(),
                                      stream), "pn_head_forward")
        self.last_launch_count = lib.pn_last_launch_count()
        all_cls_scores = dict(sub=o["sub"], obj=o["obj"], cls=o["cls"], rel=o["rel"], importance=o["importance"])
        all_mask_preds = dict(mask=o["mask"])
        if materialize_seg:
            all_mask_preds.update(sub_seg=o["sub_seg"], obj_seg=o["obj_seg"])
        if taps is not None:
            taps.update(sub_pos=o["sub_pos"], obj_pos=o["obj_pos"])
        self.last_pairs = (o["sub_pos"], o["obj_pos"])
        return all_cls_scores, all_mask_preds
