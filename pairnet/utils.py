"""Import target of ``custom_imports`` (reference ``pairnet/utils.py`` holds visualisation / NMS helpers
that are not on the CrossHead2 path)."""
