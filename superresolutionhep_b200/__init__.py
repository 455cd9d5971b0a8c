"""B200-native (sm_100a) implementation of the SuperResolutionHEP sampling hot path."""
from .config import SrDims  # noqa: F401

__all__ = ["SrDims"]
