"""patchioner_b200 -- B200-native (sm_100a) drop-in for Patch-ioner's patch -> region -> caption hot path.

``from patchioner_b200 import Patchioner`` mirrors ``from patchioner import Patchioner``
(Patch-ioner/src/__init__.py:1).  All compute is in ``libpio_sm100.so`` (C ABI: include/pio.h).
"""
from ._lib import PioError, lib  # noqa: F401
from .model import Patchioner  # noqa: F401


class AutoModel:
    """``transformers.AutoModel.from_pretrained(MODEL_ID, trust_remote_code=True)`` resolves, for the Patch-ioner checkpoints, to
    remote code that wraps the same class (reference README.md:44-52).  Here: a local config (dict / YAML path) instead of a
    hub id -- there is no network."""

    @staticmethod
    def from_pretrained(config, device="cuda", trust_remote_code=True, **overrides):
        return Patchioner.from_config(config, device=device, **overrides)


__all__ = ["Patchioner", "AutoModel", "PioError", "lib"]
