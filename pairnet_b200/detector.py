"""``PSGTr`` detector wrapper -- the only caller of the hot path (reference
``pairnet/models/frameworks/psgtr.py:73-156``): backbone -> ``bbox_head``.  Same class name,
constructor and method signatures; backbone/pixel decoder are device-side PyTorch plumbing, the
head is the CUDA library."""
import warnings

import torch
import torch.nn as nn

from .registry import DETECTORS, build_backbone, build_head, to_config


@DETECTORS.register_module()
class PSGTr(nn.Module):
    def __init__(self, backbone, bbox_head, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None, neck=None):
        super().__init__()
        if neck is not None:
            raise NotImplementedError("PSGTr configs of Pair-Net have no neck")
        self.backbone = build_backbone(to_config(backbone))
        bbox_head = to_config(bbox_head)
        bbox_head.update(train_cfg=train_cfg, test_cfg=test_cfg)  # mmdet SingleStageDetector.__init__
        self.bbox_head = build_head(bbox_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.num_classes = self.bbox_head.num_classes
        # reduced-precision upstream for the bf16-class configs (BASELINE configs 3-4): torch.bfloat16 runs the backbone
        # under autocast; its feature maps are handed to the head as fp32 (the CUDA library's activation type)
        self.backbone_autocast = None

    def init_weights(self):
        self.backbone.init_weights()
        self.bbox_head.init_weights()

    def extract_feat(self, img):
        if self.backbone_autocast is None:
            return self.backbone(img)
        with torch.autocast(img.device.type, dtype=self.backbone_autocast):
            feats = self.backbone(img)
        return tuple(f.float() for f in feats)

    def forward_dummy(self, img):
        """psgtr.py:92-110."""
        batch_size, _, height, width = img.shape
        dummy_img_metas = [dict(batch_input_shape=(height, width), img_shape=(height, width, 3))
                           for _ in range(batch_size)]
        x = self.extract_feat(img)
        return self.bbox_head(x, dummy_img_metas)

    def forward_train(self, img, img_metas, gt_rels=None, gt_bboxes=None, gt_labels=None, gt_masks=None,
                      gt_bboxes_ignore=None):
        """psgtr.py:112-146.  ``gt_masks``: per image a ``[G,h,w]`` tensor (or an object with mmdet ``BitmapMasks``'
        ``to_ndarray()``); padded to the batch shape and resized to half resolution with nearest sampling."""
        import torch.nn.functional as F
        # the backbone (and the pixel decoder inside the head) stay on the no-grad path, outside any enclosing autocast
        with torch.no_grad(), torch.autocast(img.device.type, enabled=False):
            x = self.extract_feat(img)
        if self.bbox_head.use_mask:
            assert gt_masks is not None
            _, _, H, W = img.shape
            new_gt_masks = []
            for each in gt_masks:
                mask = each if torch.is_tensor(each) else torch.as_tensor(each.to_ndarray())
                mask = mask.to(x[0].device)
                _, h, w = mask.shape
                mask = F.interpolate(F.pad(mask, (0, W - w, 0, H - h)).unsqueeze(1).float(), size=(H // 2, W // 2),
                                     mode="nearest").squeeze(1).to(mask.dtype)
                new_gt_masks.append(mask)
            gt_masks = new_gt_masks
        return self.bbox_head.forward_train(x, img_metas, gt_rels, gt_bboxes, gt_labels, gt_masks, gt_bboxes_ignore)

    def simple_test(self, img, img_metas, rescale=False):
        """psgtr.py:148-156: head post-processing -> one ``Result`` per image."""
        from .results import triplet2Result
        feat = self.extract_feat(img)
        results_list = self.bbox_head.simple_test(feat, img_metas, rescale=rescale)
        return [triplet2Result(triplets, self.bbox_head.use_mask) for triplets in results_list]

    def forward_test(self, imgs, img_metas, **kwargs):
        """mmdet ``BaseDetector.forward_test``: ``imgs`` / ``img_metas`` are lists over test-time augmentations (what
        ``single_gpu_test`` passes: ``model(return_loss=False, rescale=True, **data)``, reference ``tools/test.py:250-267``)."""
        for var, name in ((imgs, "imgs"), (img_metas, "img_metas")):
            if not isinstance(var, (list, tuple)):
                raise TypeError(f"{name} must be a list, but got {type(var)}")
        if len(imgs) != len(img_metas):
            raise ValueError(f"num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})")
        for img, metas in zip(imgs, img_metas):
            for meta in metas:
                meta.setdefault("batch_input_shape", tuple(img.shape[-2:]))
        if len(imgs) == 1:
            return self.simple_test(imgs[0], img_metas[0], **kwargs)
        raise NotImplementedError("test-time augmentation (aug_test) is not provided by the reference detector either")

    def forward(self, img, img_metas=None, return_loss=True, **kwargs):
        """mmdet ``BaseDetector.forward``: ``return_loss=True`` -> ``forward_train``; otherwise ``forward_test`` on the
        augmentation lists.  A bare tensor with no metas is the ``forward_dummy`` call (FLOPs tools, ``bench.py``)."""
        if img_metas is None and torch.is_tensor(img):
            return self.forward_dummy(img)
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)


class GraphedForward:
    """Capture ``model.forward_dummy`` on a static input into one CUDA graph and replay it.

    The whole query -> pair -> relation forward (plus the PyTorch upstream) then costs a single graph
    launch per step; outputs are the static tensors of the captured run."""

    def __init__(self, fn, example, warmup=3):
        self.static_in = example.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                fn(self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)

    def __call__(self, x=None):
        if x is not None and x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
