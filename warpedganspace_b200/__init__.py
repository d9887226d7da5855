"""warpedganspace_b200 — B200-native (sm_100a) implementation of the WarpedGANSpace training-step
hot path behind the reference's SupportSets / Reconstructor / generator-wrapper API."""
from . import _lib                                   # noqa: F401
from .support_sets import SupportSets                # noqa: F401

__all__ = ['SupportSets']
