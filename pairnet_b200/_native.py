"""ctypes binding of ``libpairnet_b200.so`` (C-ABI declared in ``include/pairnet_b200.h``).

There is NO fallback: if the library is missing and cannot be built, or a kernel launch fails,
this module raises.  Nothing here touches the CPU oracle."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpairnet_b200.so")

PN_MAX_LAYERS = 16
PN_MAX_LEVELS = 4
EMBED_DIMS = 256

c_float_p = C.POINTER(C.c_float)
c_void_p = C.c_void_p


class PnLinear(C.Structure):
    _fields_ = [("w", c_void_p), ("b", c_void_p)]


class PnNorm(C.Structure):
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p)]


class PnMHA(C.Structure):
    _fields_ = [("in_proj_w", c_void_p), ("in_proj_b", c_void_p), ("out_proj_w", c_void_p), ("out_proj_b", c_void_p)]


class PnDecoderLayer(C.Structure):
    _fields_ = [("cross_attn", PnMHA), ("self_attn", PnMHA), ("ffn1", PnLinear), ("ffn2", PnLinear),
                ("norm", PnNorm * 3)]


class PnMlp3(C.Structure):
    _fields_ = [("l", PnLinear * 3)]


class PnConvTiny(C.Structure):
    _fields_ = [("w", c_void_p * 3), ("b", c_void_p * 3), ("mid_channels", C.c_int)]


class PnM2FWeights(C.Structure):
    _fields_ = [("num_queries", C.c_int), ("num_layers", C.c_int), ("num_levels", C.c_int), ("ffn_dims", C.c_int),
                ("num_cls", C.c_int), ("query_feat", c_void_p), ("query_embed", c_void_p), ("level_embed", c_void_p),
                ("post_norm", PnNorm), ("cls_embed", PnLinear), ("mask_embed", PnMlp3),
                ("layers", PnDecoderLayer * PN_MAX_LAYERS), ("prepared", c_void_p)]


class PnM2FInputs(C.Structure):
    _fields_ = [("B", C.c_int), ("H4", C.c_int), ("W4", C.c_int), ("mask_features", c_void_p),
                ("h", C.c_int * PN_MAX_LEVELS), ("w", C.c_int * PN_MAX_LEVELS),
                ("memory", c_void_p * PN_MAX_LEVELS), ("pos", c_void_p * PN_MAX_LEVELS),
                ("memory_token_major", C.c_int * PN_MAX_LEVELS), ("memory_batch_stride", C.c_longlong * PN_MAX_LEVELS),
                ("mask_features_token_major", C.c_int)]


class PnM2FOutputs(C.Structure):
    _fields_ = [("query_out", c_void_p), ("cls_pred", c_void_p), ("mask_pred", c_void_p), ("query_trace", c_void_p),
                ("mask_trace", c_void_p), ("trace_words", C.c_int)]


class PnRelWeights(C.Structure):
    _fields_ = [("num_rel_queries", C.c_int), ("num_layers", C.c_int), ("ffn_dims", C.c_int), ("num_rel_cls", C.c_int),
                ("rel_query_feat", c_void_p), ("rel_query_embed", c_void_p), ("rel_query_embed2", c_void_p),
                ("rel_cls_embed", PnLinear), ("layers", PnDecoderLayer * PN_MAX_LAYERS), ("prepared", c_void_p)]


class PnHeadWeights(C.Structure):
    _fields_ = [("m2f", PnM2FWeights), ("sub_query_update", PnMlp3), ("obj_query_update", PnMlp3),
                ("update_importance", PnConvTiny), ("rel", PnRelWeights)]


class PnHeadOutputs(C.Structure):
    _fields_ = [("cls", c_void_p), ("mask", c_void_p), ("importance", c_void_p), ("rel", c_void_p),
                ("sub_pos", c_void_p), ("obj_pos", c_void_p), ("sub", c_void_p), ("obj", c_void_p),
                ("sub_seg", c_void_p), ("obj_seg", c_void_p), ("query_out", c_void_p), ("importance_raw", c_void_p),
                ("pair_feat", c_void_p), ("rel_feat", c_void_p), ("query_trace", c_void_p), ("mask_trace", c_void_p),
                ("trace_words", C.c_int)]


class PnMsdaEncoderLayer(C.Structure):
    _fields_ = [("sampling_offsets", PnLinear), ("attention_weights", PnLinear), ("value_proj", PnLinear),
                ("output_proj", PnLinear), ("ffn1", PnLinear), ("ffn2", PnLinear), ("norm", PnNorm * 2)]


class PnMsdaEncoderWeights(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("num_levels", C.c_int), ("num_points", C.c_int), ("ffn_dims", C.c_int),
                ("layers", PnMsdaEncoderLayer * PN_MAX_LAYERS), ("prepared", c_void_p)]


i32, i64, sz, vp = C.c_int, C.c_longlong, C.c_size_t, c_void_p
P = C.POINTER

# name -> (restype, argtypes); every symbol declared in include/pairnet_b200.h
PN_OPT_TENSOR_CORES, PN_OPT_OVERLAP, PN_OPT_SKINNY, PN_OPT_MASK_TC, PN_OPT_FUSED_CHAIN = 0, 3, 8, 9, 10  # include/pairnet_b200.h
PN_OPT_TOPK_RADIX, PN_OPT_PPN_TC, PN_OPT_PPN_FUSED_TOPK, PN_OPT_CONV_TC = 6, 7, 11, 12
PN_OPT_UMMA_TMA_STORE, PN_OPT_SINGLE_PASS, PN_OPT_NVTX, PN_OPT_PDL, PN_OPT_ENC_BF16X3, PN_OPT_PPN_EPI2, PN_OPT_PPN_HALF_KB, PN_OPT_PPN_SPECULATE = 13, 14, 15, 16, 17, 18, 19, 20

SIGNATURES = {
    "pn_version": (i32, []),
    "pn_last_error_string": (C.c_char_p, []),
    "pn_last_launch_count": (i32, []),
    "pn_debug_chain_timing": (i32, [vp, i32]),
    "pn_set_option": (i32, [i32, i32]),
    "pn_get_option": (i32, [i32]),
    "pn_device_info": (i32, [P(i32), P(i32), P(i32)]),
    "pn_sine_posenc": (i32, [vp, i32, i32, vp]),
    "pn_level_prep": (i32, [vp, vp, vp, vp, vp, i32, i32, vp]),
    "pn_level_prep_tokens": (i32, [vp, i64, vp, vp, vp, vp, i32, i32, vp]),
    "pn_mask_feature_resize": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "pn_attn_mask_bits": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "pn_mask_pred": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "pn_mask_feature_resize_tokens": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "pn_nchw_to_tokens": (i32, [vp, vp, i32, i32, vp]),
    "pn_mask_tc_workspace_bytes": (sz, [i32, i32]),
    "pn_attn_mask_bits_tc": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]),
    "pn_mask_pred_tc": (i32, [vp, vp, vp, i32, i32, i32, vp, sz, vp]),
    "pn_linear": (i32, [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "pn_linear_tc_workspace_bytes": (sz, [i32, i32, i32]),
    "pn_linear_tc": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp, sz, vp]),
    "pn_split_tf32": (i32, [vp, vp, vp, sz, vp]),
    "pn_linear_tc_rawa": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "pn_split_bf16": (i32, [vp, vp, vp, sz, vp]),
    "pn_linear_tc_bf16x3": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "pn_linear_tc_presplit": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "pn_add_layernorm": (i32, [vp, vp, vp, vp, vp, i32, vp]),
    "pn_mha_workspace_bytes": (sz, [i32, i32, i32]),
    "pn_mha_core": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, i32, i32, i32, vp, sz, vp]),
    "pn_mha_core_tc_workspace_bytes": (sz, [i32, i32, i32]),
    "pn_mha_core_tc": (i32, [vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, vp, sz, vp]),
    "pn_m2f_decoder_workspace_bytes": (sz, [P(PnM2FWeights), P(PnM2FInputs)]),
    "pn_m2f_decoder_forward": (i32, [P(PnM2FWeights), P(PnM2FInputs), P(PnM2FOutputs), vp, sz, vp]),
    "pn_ppn_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "pn_ppn_forward": (i32, [vp, vp, P(PnMlp3), P(PnMlp3), P(PnConvTiny), vp, vp, vp, vp, vp, vp, i32, i32, i32,
                             vp, sz, vp]),
    "pn_ppn_pair_topk_bf16_workspace_bytes": (sz, [i32]),
    "pn_ppn_pair_topk_bf16": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]),
    "pn_conv_tiny": (i32, [vp, P(PnConvTiny), vp, i32, i32, vp, sz, vp]),
    "pn_topk_pairs": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "pn_rel_prepared_bytes": (sz, [P(PnRelWeights)]),
    "pn_rel_prepare": (i32, [P(PnRelWeights), vp, sz, vp]),
    "pn_m2f_prepared_bytes": (sz, [P(PnM2FWeights)]),
    "pn_m2f_prepare": (i32, [P(PnM2FWeights), vp, sz, vp]),
    "pn_relation_fusion_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "pn_relation_fusion_forward": (i32, [P(PnRelWeights), vp, vp, vp, i32, i32, vp, sz, vp]),
    "pn_gather_rows": (i32, [vp, vp, vp, i32, i32, i32, i64, vp]),
    "pn_upsample_threshold": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "pn_panoptic_merge": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, C.c_longlong, vp, vp, vp]),
    "pn_msda_encoder_prepared_bytes": (sz, [P(PnMsdaEncoderWeights)]),
    "pn_msda_encoder_prepare": (i32, [P(PnMsdaEncoderWeights), vp, sz, vp]),
    "pn_msda_encoder_workspace_bytes": (sz, [i32, i32, i32, i32, i32]),
    "pn_msda_encoder_forward": (i32, [P(PnMsdaEncoderWeights), vp, vp, P(i32), P(i32), vp, i32, vp, sz, vp]),
    "pn_group_norm_workspace_bytes": (sz, [i32, i32, i32]),
    "pn_group_norm": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, C.c_float, vp, sz, vp]),
    "pn_maxpool3x3s2_nhwc": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "pn_gn_upsample_add": (i32, [vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, i32, i32, C.c_float, vp, sz, vp]),
    "pn_conv1x1_nhwc_to_nchw_workspace_bytes": (sz, [i32]),
    "pn_conv1x1_nhwc_to_nchw": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, sz, vp]),
    "pn_msda_sample": (i32, [vp, vp, vp, P(i32), P(i32), i32, i32, i32, vp]),
    "pn_head_workspace_bytes": (sz, [P(PnHeadWeights), P(PnM2FInputs)]),
    "pn_head_forward": (i32, [P(PnHeadWeights), P(PnM2FInputs), P(PnHeadOutputs), vp, sz, vp]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load (building first if needed) the CUDA library.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise NativeError(f"{LIB_PATH} is missing; run `python -m pairnet_b200.build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.pn_version() < 100:
        raise NativeError("libpairnet_b200.so is older than this Python package")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pn_last_error_string()
        raise NativeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
