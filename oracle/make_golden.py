"""Mint the golden fixtures under ``tests/golden/`` (TEST INFRASTRUCTURE).

Run in the dev container (``python -m oracle.make_golden``).  Two families:

1. ``convtiny_ref_*.npz`` -- produced by the REFERENCE's own module
   ``/root/reference/pairnet/models/frameworks/cnn_factory.py`` (imported by file path; it is
   torch-only).  These pin the oracle's ``OConvTiny`` to the reference.
2. ``head_small_*.npz`` -- produced by this oracle (the reference head cannot be imported:
   mmcv/mmdet absent).  They are regression anchors for the oracle and the fixtures the GPU
   parity tests compare against; "parity unpinned" applies to them.

Weights/inputs come from ``oracle.weights`` (numpy PCG64) so fixtures only need to store outputs.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

from .head import HeadHyper, OConvTiny, OCrossHead2
from .weights import fixture_state_dict, numpy_state_dict, numpy_tensor

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_CNN = "/root/reference/pairnet/models/frameworks/cnn_factory.py"

CONV_CASES = [  # (tag, mid_channels, B, N, seed)
    ("m64_n100", 64, 2, 100, 11),
    ("m64_n37", 64, 1, 37, 12),
    ("m16_n24", 16, 3, 24, 13),
]

HEAD_CASES = [  # (tag, B, (H4,W4) of mask_feature, seed)
    ("b2_32x48", 2, (32, 48), 21),
    ("b1_40x56", 1, (40, 56), 22),
]


def small_head_inputs(B, hw4, seed):
    H, W = hw4
    mf = numpy_tensor((B, 256, H, W), seed, 0.5)
    mems = [numpy_tensor((B, 256, H // 8, W // 8), seed + 1), numpy_tensor((B, 256, H // 4, W // 4), seed + 2),
            numpy_tensor((B, 256, H // 2, W // 2), seed + 3)]
    return mf, mems


def build_small_head(seed=10086, dtype=torch.float32):
    head = OCrossHead2(HeadHyper(with_pixel_decoder=False))
    head.load_state_dict(fixture_state_dict(head, seed))
    return head.to(dtype).eval()


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    spec = importlib.util.spec_from_file_location("ref_cnn_factory", REF_CNN)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for tag, mid, B, N, seed in CONV_CASES:
        m = ref.ConvTiny(mid_channels=mid).eval()
        m.load_state_dict(numpy_state_dict(m, seed))
        x = torch.tanh(numpy_tensor((B, N, N), seed + 100))
        with torch.no_grad():
            y = m(x)
            mine = OConvTiny(mid_channels=mid).eval()
            mine.load_state_dict(m.state_dict())
            assert torch.equal(mine(x), y), "oracle ConvTiny != reference ConvTiny"
        np.savez_compressed(os.path.join(GOLDEN, f"convtiny_ref_{tag}.npz"), out=y.numpy(),
                            meta=np.array([mid, B, N, seed]))
        print("convtiny", tag, tuple(y.shape), float(y.abs().mean()))

    head = build_small_head()
    for tag, B, hw4, seed in HEAD_CASES:
        mf, mems = small_head_inputs(B, hw4, seed)
        tr = {}
        with torch.no_grad():
            cls, msk = head.forward_from_memories(mf, mems, trace=tr)
        np.savez_compressed(
            os.path.join(GOLDEN, f"head_small_{tag}.npz"),
            cls=cls["cls"].numpy(), rel=cls["rel"].numpy(), importance=cls["importance"].numpy(),
            sub=cls["sub"].numpy(), obj=cls["obj"].numpy(),
            sub_pos=tr["sub_pos"].numpy(), obj_pos=tr["obj_pos"].numpy(),
            importance_raw=tr["importance_raw"].numpy(),
            query_last=tr["query_feat"][-1].numpy(),
            rel_last=tr["rel_feat"][-1].numpy(),
            mask_sub4=msk["mask"][:, :, ::4, ::4].numpy(),
            mask_frac=np.array([float(a.float().mean()) for a in tr["attn_mask"]]),
            meta=np.array([B, hw4[0], hw4[1], seed]))
        print("head", tag, float(cls["rel"].abs().mean()), tr["sub_pos"][0, :5].tolist())


if __name__ == "__main__":
    sys.exit(main())
