"""Import target of ``custom_imports`` in ``configs/mask2former/pairnet_balanced.py:402-414``.  The reference module
defines the dual-decoder transformer of the PSGFormer baseline (out of scope, not on the CrossHead2 path); the
Pair-Net configs only import it for its registry side effects, so an empty module keeps them loading unchanged."""
