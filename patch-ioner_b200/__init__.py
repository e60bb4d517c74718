"""patchioner_b200 -- B200-native (sm_100a) drop-in for Patch-ioner's patch -> region -> caption hot path.

``from patchioner_b200 import Patchioner`` mirrors ``from patchioner import Patchioner``
(Patch-ioner/src/__init__.py:1).  All compute is in ``libpio_sm100.so`` (C ABI: include/pio.h).
"""
from ._lib import PioError, lib  # noqa: F401
from .model import Patchioner  # noqa: F401

__all__ = ["Patchioner", "PioError", "lib"]
