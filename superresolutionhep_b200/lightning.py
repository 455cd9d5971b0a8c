"""Drop-in for the inference-time surface of the reference's ``lightning.py:SupResLightning``
(lines 31-38): same constructor, ``.net`` attribute holding the flow model, and therefore
the same ``net.``-prefixed ``state_dict`` keys that ``inference.py:75-83`` loads with
``load_state_dict(checkpoint['state_dict'])`` before ``.eval()`` / ``.cuda()``.

``pytorch_lightning`` is optional: when importable the class derives from
``LightningModule`` (so Lightning checkpoints' hooks resolve), otherwise from ``nn.Module``.
Training / validation steps are out of scope.
"""
from __future__ import annotations

from torch import nn

from .flow_model import FlowModel

try:                                                    # pragma: no cover - not installed in this image
    from pytorch_lightning import LightningModule as _Base
except Exception:                                       # noqa: BLE001
    _Base = nn.Module


class SupResLightning(_Base):
    def __init__(self, config_mv, config_t, comet_logger=None, precision=None):
        super().__init__()
        self.config_mv = config_mv
        self.config_t = config_t
        self.net = FlowModel(self.config_mv["flow_model"], precision=precision)
        self.comet_logger = comet_logger

    def set_comet_logger(self, comet_logger):
        self.comet_logger = comet_logger

    def forward(self, batch, noisy_input, time_step):
        return self.net(batch, noisy_input, time_step)
