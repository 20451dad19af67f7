"""Swin backbone (upstream plumbing for BASELINE config 4, `configs/mask2former/pairnet_swinb.py:203-228`): the
reference's Swin-B config builds unchanged, the state-dict surface follows mmdet 2.25.1, and the shifted-window
attention equals a dense masked-attention formulation written independently.  mmdet is absent: values are not pinned
to the reference (stated in the module header)."""
import math
import os

import pytest
import torch

from pairnet_b200.upstream.swin import SwinBlock, SwinTransformer, PatchMerging

REF_CFG = "/root/reference/configs/mask2former/pairnet_swinb.py"


def test_swin_state_dict_surface_and_shapes():
    net = SwinTransformer(embed_dims=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8), window_size=4, frozen_stages=1)
    sd = net.state_dict()
    for k, shape in {
        "patch_embed.projection.weight": (32, 3, 4, 4), "patch_embed.norm.weight": (32,),
        "stages.0.blocks.1.attn.w_msa.relative_position_bias_table": (49, 1),
        "stages.0.blocks.1.attn.w_msa.relative_position_index": (16, 16),
        "stages.1.blocks.0.attn.w_msa.qkv.weight": (192, 64), "stages.1.blocks.0.attn.w_msa.proj.bias": (64,),
        "stages.2.blocks.0.ffn.layers.0.0.weight": (512, 128), "stages.2.blocks.0.ffn.layers.1.weight": (128, 512),
        "stages.0.downsample.norm.weight": (128,), "stages.0.downsample.reduction.weight": (64, 128),
        "norm3.weight": (256,),
    }.items():
        assert tuple(sd[k].shape) == shape, k
    assert not any(k.startswith("stages.3.downsample") for k in sd)
    assert not net.patch_embed.projection.weight.requires_grad and not net.stages[0].blocks[0].norm1.weight.requires_grad
    assert net.stages[1].blocks[0].norm1.weight.requires_grad
    with torch.no_grad():
        outs = net(torch.randn(2, 3, 50, 70))  # not a multiple of the patch / window size: padded
    assert [tuple(o.shape) for o in outs] == [(2, 32, 13, 18), (2, 64, 7, 9), (2, 128, 4, 5), (2, 256, 2, 3)]
    # original Swin index table: (dy + Wh - 1) * (2 Ww - 1) + (dx + Ww - 1)
    idx = net.stages[0].blocks[0].attn.w_msa.relative_position_index
    p = lambda i: (i // 4, i % 4)
    for i in (0, 5, 15):
        for j in (0, 6, 15):
            assert int(idx[i, j]) == (p(i)[0] - p(j)[0] + 3) * 7 + (p(i)[1] - p(j)[1] + 3)


@pytest.mark.parametrize("shift,H,W", [(False, 8, 8), (True, 8, 8), (True, 7, 10), (False, 5, 9)])
def test_shifted_window_block_equals_dense_masked_attention(shift, H, W):
    torch.manual_seed(3)
    ws, C, heads = 4, 16, 2
    blk = SwinBlock(C, heads, 32, ws, shift, True, None).double()
    with torch.no_grad():
        blk.attn.w_msa.relative_position_bias_table.normal_(0, 0.5)
    x = torch.randn(2, H * W, C, dtype=torch.float64)
    with torch.no_grad():
        got = blk(x, (H, W))
    # dense formulation over the padded, cyclically shifted grid: token i may attend token j iff both fall in the same
    # window AND in the same pre-shift region (the 3 x 3 partition of the rolled image)
    ss = ws // 2 if shift else 0
    Hp, Wp = math.ceil(H / ws) * ws, math.ceil(W / ws) * ws
    with torch.no_grad():
        h = blk.norm1(x).view(2, H, W, C)
        h = torch.nn.functional.pad(h, (0, 0, 0, Wp - W, 0, Hp - H))
        h = torch.roll(h, (-ss, -ss), (1, 2)).reshape(2, Hp * Wp, C)
        qkv = blk.attn.w_msa.qkv(h).view(2, Hp * Wp, 3, heads, C // heads)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        ys, xs = torch.meshgrid(torch.arange(Hp), torch.arange(Wp), indexing="ij")
        ys, xs = ys.flatten(), xs.flatten()
        def region(c, n):
            if ss == 0:
                return torch.zeros_like(c)
            return (c >= n - ws).long() + (c >= n - ss).long()
        same_win = (ys[:, None] // ws == ys[None] // ws) & (xs[:, None] // ws == xs[None] // ws)
        same_reg = (region(ys, Hp)[:, None] == region(ys, Hp)[None]) & (region(xs, Wp)[:, None] == region(xs, Wp)[None])
        dy, dx = ys[:, None] - ys[None], xs[:, None] - xs[None]
        table = blk.attn.w_msa.relative_position_bias_table
        bias = table[((dy + ws - 1) * (2 * ws - 1) + (dx + ws - 1)).clamp(0, table.shape[0] - 1)]  # [L,L,heads]
        logits = torch.einsum("blhd,bmhd->bhlm", q, k) * (C // heads) ** -0.5 + bias.permute(2, 0, 1)[None]
        logits = logits + torch.where(same_reg, 0.0, -100.0)[None, None]       # the reference's additive -100 mask
        logits = logits.masked_fill(~same_win[None, None], float("-inf"))
        att = torch.einsum("bhlm,bmhd->blhd", logits.softmax(-1), v).reshape(2, Hp * Wp, C)
        att = blk.attn.w_msa.proj(att).view(2, Hp, Wp, C)
        att = torch.roll(att, (ss, ss), (1, 2))[:, :H, :W].reshape(2, H * W, C)
        y = x + att
        want = y + blk.ffn.layers(blk.norm2(y))
    assert float((got - want).abs().max()) < 1e-10


def test_patch_merging_follows_unfold_channel_order():
    torch.manual_seed(0)
    pm = PatchMerging(6, 12).double()
    x = torch.randn(2, 5 * 7, 6, dtype=torch.float64)
    with torch.no_grad():
        got, hw = pm(x, (5, 7))
        img = x.view(2, 5, 7, 6).permute(0, 3, 1, 2)
        img = torch.nn.functional.pad(img, (0, 1, 0, 1))
        unf = torch.nn.Unfold(kernel_size=2, stride=2)(img).transpose(1, 2)   # what mmdet's PatchMerging samples
        want = pm.reduction(pm.norm(unf))
    assert hw == (3, 4) and float((got - want).abs().max()) < 1e-12


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present (GPU box)")
def test_reference_swinb_config_builds_unchanged():
    from pairnet_b200.registry import Config, build_detector
    det = build_detector(Config.fromfile(REF_CFG).model)
    assert type(det.backbone).__name__ == "SwinTransformer"
    assert det.backbone.num_features == [128, 256, 512, 1024]
    assert len(det.backbone.stages[2].blocks) == 18
    assert det.bbox_head.pixel_decoder.input_convs[0].conv.in_channels == 1024
    assert sum(1 for k in det.state_dict() if k.startswith("backbone.")) == 357
