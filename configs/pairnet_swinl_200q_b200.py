# BASELINE config 4: Pair-Net with a Swin-L backbone, 200 object / 200 relation queries, 1024x1024 inputs.
# Not a shipped reference config: the reference ships Swin-B / 100 queries (configs/mask2former/pairnet_swinb.py:203-237);
# this is that file's model with SURVEY 8d's substitutions (embed_dims 192, heads 6/12/24/48, in_channels 192..1536,
# 200 queries) on top of the R50 head settings -- "extrapolated config".
_base_ = ["./pairnet_r50_b200.py"]

model = dict(
    backbone=dict(_delete_=True, type="SwinTransformer", embed_dims=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48],
                  window_size=12, mlp_ratio=4, qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                  drop_path_rate=0.3, patch_norm=True, out_indices=(0, 1, 2, 3), convert_weights=True, frozen_stages=3,
                  pretrain_img_size=384),
    bbox_head=dict(num_obj_query=200, num_rel_query=200, in_channels=[192, 384, 768, 1536], strides=[4, 8, 16, 32]),
)
