"""CPU oracle for the Pair-Net relation-head hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the timed CPU baseline), never as the thing shipped.  The
product package ``pairnet_b200`` must not import from here and fails loudly
when its CUDA library is missing.

What it is: a torch-only (fp32 / fp64, CPU) restatement of
``CrossHead2.forward`` (reference ``pairnet/models/relation_heads/
pairnet_head.py:216-417``) plus the un-vendored mmcv-full 1.7.0 / mmdet 2.25.1
bricks that forward instantiates (restated from the in-repo copies at
``pairnet/models/relation_heads/facebook_detr.py:289-432`` and from the
published upstream algorithms).

PARITY PINNING STATUS
* ``ConvTiny`` (Matrix-Learner filter) is PINNED: the
  reference's own ``pairnet/models/frameworks/cnn_factory.py`` imports here, and
  ``oracle/make_golden.py`` ran it to mint ``tests/golden/convtiny_*.npz``.
* Every other row is UNPINNED by the reference ("parity unpinned"): the
  reference ships no tests / golden vectors and cannot be imported in this
  container (mmcv, mmdet, detectron2, panopticapi absent, no network).  Those
  rows are anchored on the same torch primitives the reference calls
  (``nn.MultiheadAttention``, ``nn.LayerNorm``, ``F.interpolate``,
  ``torch.topk``, ``torch.gather``) at the cited call sites.
"""
