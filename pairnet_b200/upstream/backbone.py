"""UPSTREAM of the hot path (out of scope for hand-written kernels): the mmdet
``ResNet`` backbone named by ``configs/mask2former/pairnet.py:9-19``, provided by torchvision's
ResNet (same architecture and parameter names for ``style='pytorch'``), frozen BN in eval."""
import torch.nn as nn

from ..registry import BACKBONES


@BACKBONES.register_module()
class ResNet(nn.Module):
    def __init__(self, depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=-1, norm_cfg=None,
                 norm_eval=True, style="pytorch", init_cfg=None, **kwargs):
        super().__init__()
        import torchvision
        ctor = {50: torchvision.models.resnet50, 101: torchvision.models.resnet101}.get(depth)
        if ctor is None:
            raise NotImplementedError(f"ResNet depth {depth}")
        if style != "pytorch":
            raise NotImplementedError("only style='pytorch' ResNets are provided")
        net = ctor(weights=None)
        for name in ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4"):
            setattr(self, name, getattr(net, name))
        self.out_indices = tuple(out_indices)
        self.frozen_stages = frozen_stages
        self.norm_eval = norm_eval
        if norm_cfg is not None and not norm_cfg.get("requires_grad", True):
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    for p in m.parameters():
                        p.requires_grad_(False)
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            for m in (self.conv1, self.bn1):
                for p in m.parameters():
                    p.requires_grad_(False)
        for i in range(1, self.frozen_stages + 1):
            for p in getattr(self, f"layer{i}").parameters():
                p.requires_grad_(False)

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    def init_weights(self):
        pass

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        outs = []
        for i in range(4):
            x = getattr(self, f"layer{i + 1}")(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)
