"""UPSTREAM of the hot path (BASELINE config 4): the mmdet 2.25.1 ``SwinTransformer`` backbone named by
``configs/mask2former/pairnet_swinb.py:203-228`` as device-side PyTorch plumbing (cuBLAS linears,
``scaled_dot_product_attention``), like the ResNet of ``backbone.py``.  Module / parameter / buffer names follow mmdet
(``patch_embed.projection``, ``stages.N.blocks.M.attn.w_msa.{qkv,proj,relative_position_bias_table,
relative_position_index}``, ``ffn.layers.0.0`` / ``ffn.layers.1``, ``stages.N.downsample.{norm,reduction}``, ``normN``) so
that a converted Swin checkpoint loads with ``strict=True``.  mmdet is not installed here: PARITY UNPINNED (restated
from the published architecture; shapes and the state-dict surface are tested, values are not pinned)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..registry import BACKBONES


class WindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window_size, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.window_size = (window_size, window_size)
        head_dim = embed_dims // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        Wh = Ww = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * Wh - 1) * (2 * Ww - 1), num_heads))
        # mmdet WindowMSA: index = coords + coords^T, flipped along dim 1 (equals the original Swin index table)
        seq1 = torch.arange(0, (2 * Ww - 1) * Wh, 2 * Ww - 1)
        seq2 = torch.arange(0, Ww, 1)
        coords = (seq1[:, None] + seq2[None, :]).reshape(1, -1)
        self.register_buffer("relative_position_index", (coords + coords.T).flip(1).contiguous())
        self.qkv = nn.Linear(embed_dims, embed_dims * 3, bias=qkv_bias)
        self.proj = nn.Linear(embed_dims, embed_dims)

    def bias(self):
        n = self.window_size[0] * self.window_size[1]
        b = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(n, n, -1)
        return b.permute(2, 0, 1).contiguous()  # [heads, n, n]

    def forward(self, x, mask=None):
        """x [nW*B, n, C]; mask [nW, n, n] (0 / -100) or None."""
        Bn, n, C = x.shape
        qkv = self.qkv(x).reshape(Bn, n, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        bias = self.bias().unsqueeze(0)  # [1, heads, n, n]
        if mask is not None:
            nW = mask.shape[0]
            bias = (bias.unsqueeze(0) + mask[None, :, None]).expand(Bn // nW, nW, self.num_heads, n, n).reshape(
                Bn, self.num_heads, n, n)
        x = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype), scale=self.scale)
        return self.proj(x.transpose(1, 2).reshape(Bn, n, C))


class ShiftWindowMSA(nn.Module):
    def __init__(self, embed_dims, num_heads, window_size, shift_size, qkv_bias, qk_scale):
        super().__init__()
        self.window_size, self.shift_size = window_size, shift_size
        self.w_msa = WindowMSA(embed_dims, num_heads, window_size, qkv_bias, qk_scale)
        self._mask_cache = {}

    def _shift_mask(self, Hp, Wp, device):
        key = (Hp, Wp, str(device))
        if key not in self._mask_cache:
            ws, ss = self.window_size, self.shift_size
            img = torch.zeros((1, Hp, Wp, 1), device=device)
            cnt = 0
            for h in (slice(0, -ws), slice(-ws, -ss), slice(-ss, None)):
                for w in (slice(0, -ws), slice(-ws, -ss), slice(-ss, None)):
                    img[:, h, w, :] = cnt
                    cnt += 1
            mw = self._windows(img).reshape(-1, ws * ws)
            m = mw.unsqueeze(1) - mw.unsqueeze(2)
            self._mask_cache[key] = m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)
        return self._mask_cache[key]

    def _windows(self, x):
        B, H, W, C = x.shape
        ws = self.window_size
        x = x.view(B, H // ws, ws, W // ws, ws, C)
        return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)

    def forward(self, query, hw_shape):
        B, L, C = query.shape
        H, W = hw_shape
        ws, ss = self.window_size, self.shift_size
        x = query.view(B, H, W, C)
        pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
        x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
        Hp, Wp = x.shape[1], x.shape[2]
        mask = None
        if ss > 0:
            x = torch.roll(x, shifts=(-ss, -ss), dims=(1, 2))
            mask = self._shift_mask(Hp, Wp, x.device)
        win = self._windows(x).reshape(-1, ws * ws, C)
        out = self.w_msa(win, mask).view(B, Hp // ws, Wp // ws, ws, ws, C)
        x = out.permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
        if ss > 0:
            x = torch.roll(x, shifts=(ss, ss), dims=(1, 2))
        return x[:, :H, :W, :].reshape(B, H * W, C)


class _FFN(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(d, ff), nn.GELU(), nn.Dropout(0.0)), nn.Linear(ff, d),
                                    nn.Dropout(0.0))

    def forward(self, x, identity):
        return identity + self.layers(x)


class SwinBlock(nn.Module):
    def __init__(self, embed_dims, num_heads, ff, window_size, shift, qkv_bias, qk_scale):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dims)
        self.attn = ShiftWindowMSA(embed_dims, num_heads, window_size, window_size // 2 if shift else 0, qkv_bias, qk_scale)
        self.norm2 = nn.LayerNorm(embed_dims)
        self.ffn = _FFN(embed_dims, ff)

    def forward(self, x, hw_shape):  # drop_path is the identity at inference / with frozen stages
        x = x + self.attn(self.norm1(x), hw_shape)
        return self.ffn(self.norm2(x), identity=x)


class PatchMerging(nn.Module):
    """mmdet PatchMerging: nn.Unfold(2, stride 2) channel order (c-major, then kernel row, then kernel column)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.norm = nn.LayerNorm(4 * cin)
        self.reduction = nn.Linear(4 * cin, cout, bias=False)

    def forward(self, x, hw_shape):
        B, L, C = x.shape
        H, W = hw_shape
        x = x.view(B, H, W, C)
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        Ho, Wo = x.shape[1] // 2, x.shape[2] // 2
        x = x.view(B, Ho, 2, Wo, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(B, Ho * Wo, 4 * C)
        return self.reduction(self.norm(x)), (Ho, Wo)


class SwinBlockSequence(nn.Module):
    def __init__(self, embed_dims, num_heads, ff, depth, window_size, qkv_bias, qk_scale, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([SwinBlock(embed_dims, num_heads, ff, window_size, i % 2 == 1, qkv_bias, qk_scale)
                                     for i in range(depth)])
        self.downsample = downsample

    def forward(self, x, hw_shape):
        for blk in self.blocks:
            x = blk(x, hw_shape)
        if self.downsample is not None:
            xd, hw_down = self.downsample(x, hw_shape)
            return xd, hw_down, x, hw_shape
        return x, hw_shape, x, hw_shape


class _PatchEmbed(nn.Module):
    def __init__(self, in_channels, embed_dims, patch_size, patch_norm):
        super().__init__()
        self.patch_size = patch_size
        self.projection = nn.Conv2d(in_channels, embed_dims, patch_size, patch_size)
        self.norm = nn.LayerNorm(embed_dims) if patch_norm else None

    def forward(self, x):
        p = self.patch_size
        x = F.pad(x, (0, (p - x.shape[-1] % p) % p, 0, (p - x.shape[-2] % p) % p))  # mmdet AdaptivePadding('corner')
        x = self.projection(x)
        hw = (x.shape[2], x.shape[3])
        x = x.flatten(2).transpose(1, 2)
        return (self.norm(x) if self.norm is not None else x), hw


@BACKBONES.register_module()
class SwinTransformer(nn.Module):
    def __init__(self, pretrain_img_size=224, in_channels=3, embed_dims=96, patch_size=4, window_size=7, mlp_ratio=4,
                 depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), strides=(4, 2, 2, 2), out_indices=(0, 1, 2, 3),
                 qkv_bias=True, qk_scale=None, patch_norm=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1,
                 use_abs_pos_embed=False, act_cfg=None, norm_cfg=None, with_cp=False, pretrained=None,
                 convert_weights=False, frozen_stages=-1, init_cfg=None):
        super().__init__()
        if use_abs_pos_embed:
            raise NotImplementedError("absolute position embedding (not used by the Pair-Net configs)")
        self.out_indices, self.frozen_stages = tuple(out_indices), frozen_stages
        self.patch_embed = _PatchEmbed(in_channels, embed_dims, patch_size, patch_norm)
        self.drop_after_pos = nn.Dropout(drop_rate)
        self.stages = nn.ModuleList()
        c = embed_dims
        self.num_features = []
        for i, depth in enumerate(depths):
            last = i == len(depths) - 1
            down = None if last else PatchMerging(c, 2 * c)
            self.stages.append(SwinBlockSequence(c, num_heads[i], int(mlp_ratio * c), depth, window_size, qkv_bias,
                                                 qk_scale, down))
            self.num_features.append(c)
            if not last:
                c *= 2
        for i in self.out_indices:
            self.add_module(f"norm{i}", nn.LayerNorm(self.num_features[i]))
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            for p in self.patch_embed.parameters():
                p.requires_grad_(False)
        for i in range(1, self.frozen_stages + 1):
            mods = [self.stages[i - 1]] + ([getattr(self, f"norm{i - 1}")] if (i - 1) in self.out_indices else [])
            for m in mods:
                for p in m.parameters():
                    p.requires_grad_(False)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, WindowMSA):
                nn.init.trunc_normal_(m.relative_position_bias_table, std=0.02)

    def forward(self, x):
        """[B,3,H,W] -> tuple of [B,C_i,H_i,W_i] feature maps (channels_last strides: the token layout, viewed NCHW)."""
        x, hw = self.patch_embed(x)
        x = self.drop_after_pos(x)
        outs = []
        for i, stage in enumerate(self.stages):
            x, hw, out, out_hw = stage(x, hw)
            if i in self.out_indices:
                out = getattr(self, f"norm{i}")(out)
                outs.append(out.view(-1, out_hw[0], out_hw[1], self.num_features[i]).permute(0, 3, 1, 2))
        return tuple(outs)
