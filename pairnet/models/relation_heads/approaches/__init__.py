"""reference path ``pairnet/models/relation_heads/approaches`` -> the ``Result`` container of the detector wrapper."""
from pairnet_b200.results import Result  # noqa: F401
