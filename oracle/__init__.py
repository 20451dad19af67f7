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
* ``CrossHead2.forward`` / ``forward_head`` (the op order, the top-k / gather / concat
  index logic, the parameter names) is PINNED to the reference's own source: ``oracle/pin_reference.py``
  executes ``/root/reference/pairnet/models/relation_heads/pairnet_head.py`` itself (plus the reference's
  ``cnn_factory.py`` and the ``MultiheadAttention2`` / ``BaseTransformerLayer2`` forwards of
  ``facebook_detr.py``) over constructor shims for the absent mmcv/mmdet, loads the oracle's weights with
  ``strict=True`` and requires BIT-EQUAL outputs; its outputs are committed as ``tests/golden/head_ref_*.npz``.
* ``ConvTiny`` (row 6) is PINNED on its own as well (``oracle/make_golden.py`` -> ``convtiny_ref_*.npz``).
* Still "parity unpinned" (restated from the published upstream algorithm, not on disk): the constructors /
  defaults of mmcv ``MultiheadAttention`` / ``BaseTransformerLayer``, mmcv ``FFN.forward``, mmdet
  ``SinePositionalEncoding.forward``, and everything upstream of the hot path (pixel decoder, backbone).
"""
