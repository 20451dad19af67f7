"""Import target of ``custom_imports`` in ``configs/mask2former/pairnet_balanced.py:402-414``.  The reference module
defines the PSGFormer baseline head (out of scope); no Pair-Net config instantiates it, so an empty module
keeps the config loading unchanged."""
