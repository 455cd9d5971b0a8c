"""B200-native (sm_100a) implementation of the SuperResolutionHEP sampling hot path."""
from .config import SrDims  # noqa: F401
from .flow_model import FlowModel, PackedEvents  # noqa: F401
from .lightning import SupResLightning  # noqa: F401

__all__ = ["SrDims", "FlowModel", "PackedEvents", "SupResLightning"]
