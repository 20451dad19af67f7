"""UPSTREAM of the hot path (out of scope for hand-written kernels): the mmdet
``ResNet`` backbone named by ``configs/mask2former/pairnet.py:9-19``, provided by torchvision's
ResNet (same architecture and parameter names for ``style='pytorch'``), frozen BN in eval."""
import copy

import torch
import torch.nn as nn
from torch.nn.utils.fusion import fuse_conv_bn_eval

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

    # ---- inference fast path: frozen BN folded into the convolutions, channels_last activations -------------
    def _param_state(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _fused(self):
        """A BN-folded, channels_last copy of the stem + stages (rebuilt whenever a weight changes).  The
        original modules (and their checkpoint-compatible names) stay untouched."""
        state = self._param_state()
        cache = self.__dict__.get("_fused_cache")
        if cache is not None and cache[0] == state:
            return cache[1]
        stem = fuse_conv_bn_eval(copy.deepcopy(self.conv1).eval(), copy.deepcopy(self.bn1).eval())
        stages = []
        for i in range(1, 5):
            layer = copy.deepcopy(getattr(self, f"layer{i}")).eval()
            for blk in layer:
                for c, b in (("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3")):
                    setattr(blk, c, fuse_conv_bn_eval(getattr(blk, c), getattr(blk, b)))
                    setattr(blk, b, nn.Identity())
                if blk.downsample is not None:
                    blk.downsample = nn.Sequential(fuse_conv_bn_eval(blk.downsample[0], blk.downsample[1]))
                    # bias of the fused (conv3 + projection shortcut) epilogue, see _forward_fused_epilogues
                    blk.shortcut_bias = (blk.conv3.bias + blk.downsample[0].bias).detach()
            stages.append(layer)
        fused = nn.ModuleList([stem] + stages).to(memory_format=torch.channels_last)
        for p in fused.parameters():
            p.requires_grad_(False)
        self.__dict__["_fused_cache"] = (state, fused)
        return fused

    fuse_epilogues = True  # bias / residual add / ReLU inside the cuDNN convolution epilogue (one kernel per conv)

    @staticmethod
    def _conv_act(x, conv, relu=True, residual=None):
        """conv (+ folded-BN bias) (+ residual) (+ ReLU) as ONE cuDNN fused-epilogue call."""
        if residual is not None:
            return torch.cudnn_convolution_add_relu(x, conv.weight, residual, 1.0, conv.bias, conv.stride, conv.padding,
                                                    conv.dilation, conv.groups)
        if relu:
            return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation,
                                                conv.groups)
        return conv(x)

    def _maxpool(self, x):
        """stem max-pool: ``pn_maxpool3x3s2_nhwc`` on channels_last maps (ATen's NHWC kernel is 4x off the HBM bound)."""
        mp = self.maxpool
        geom = (mp.kernel_size, mp.stride, mp.padding, mp.dilation, mp.ceil_mode)
        B, C, H, W = x.shape
        if (geom not in ((3, 2, 1, 1, False), ((3, 3), (2, 2), (1, 1), (1, 1), False)) or C % 4 or x.dtype != torch.float32
                or not x.is_contiguous(memory_format=torch.channels_last) or x.data_ptr() % 16):
            return mp(x)
        from .. import _native as nat
        y = torch.empty((B, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1), dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last)
        nat.check(nat.load().pn_maxpool3x3s2_nhwc(x.data_ptr(), y.data_ptr(), B, H, W, C,
                                                  torch.cuda.current_stream(x.device).cuda_stream), "pn_maxpool3x3s2_nhwc")
        return y

    def _forward_fused_epilogues(self, f, x):
        x = self._maxpool(self._conv_act(x, f[0]))
        outs = []
        for i in range(4):
            for blk in f[i + 1]:
                y = self._conv_act(x, blk.conv1)
                y = self._conv_act(y, blk.conv2)
                if blk.downsample is None:
                    x = self._conv_act(y, blk.conv3, residual=x)
                else:
                    # relu(conv3(y) + b3 + conv_ds(x) + b_ds): the projection shortcut runs bias-free and its (folded-BN)
                    # bias rides on conv3's fused epilogue instead of a separate elementwise pass over the map
                    ds = blk.downsample[0]
                    identity = torch.nn.functional.conv2d(x, ds.weight, None, ds.stride, ds.padding, ds.dilation, ds.groups)
                    c3 = blk.conv3
                    x = torch.cudnn_convolution_add_relu(y, c3.weight, identity, 1.0, blk.shortcut_bias, c3.stride,
                                                         c3.padding, c3.dilation, c3.groups)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)

    def forward(self, x):
        if not self.training and not torch.is_grad_enabled() and x.is_cuda and self.norm_eval:
            f = self._fused()
            x = x.contiguous(memory_format=torch.channels_last)
            if self.fuse_epilogues:
                return self._forward_fused_epilogues(f, x)
            x = self.maxpool(torch.relu_(f[0](x)))
            outs = []
            for i in range(4):
                x = f[i + 1](x)
                if i in self.out_indices:
                    outs.append(x)
            return tuple(outs)
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        outs = []
        for i in range(4):
            x = getattr(self, f"layer{i + 1}")(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)
