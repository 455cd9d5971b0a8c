"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box; callers must check
``available()``.  ``models/flow_model.py`` imports ``torchdiffeq`` and ``torchcfm`` at module
top (lines 11-12); neither is installable here, so two stub modules are injected before the
import.  The stubs are never executed on the paths we use:
``odeint`` is replaced by oracle/odeint.py and the flow matcher is training-only.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "flow_model.py"))


def _inject_stubs() -> None:
    if "torchdiffeq" not in sys.modules:
        td = types.ModuleType("torchdiffeq")
        td.odeint = None
        sys.modules["torchdiffeq"] = td
    if "torchcfm.conditional_flow_matching" not in sys.modules:
        cfm = types.ModuleType("torchcfm.conditional_flow_matching")
        cfm.TargetConditionalFlowMatcher = lambda sigma=0.0: types.SimpleNamespace(sigma=sigma)
        sys.modules.setdefault("torchcfm", types.ModuleType("torchcfm"))
        sys.modules["torchcfm.conditional_flow_matching"] = cfm


def import_reference():
    """Returns (FlowModel, SAPF) classes of the reference."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    _inject_stubs()
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)          # appended, never first (its lightning.py shadows PyPI's)
    from models.flow_model import FlowModel        # noqa: E402
    from pflow.models.model_pf import SAPF         # noqa: E402
    return FlowModel, SAPF


def build_reference_flow_model(flow_cfg: dict, state_dict: dict):
    """Reference FlowModel on CPU with ``state_dict`` loaded (cfg is deep-copied: the
    reference constructor mutates it)."""
    import copy
    FlowModel, _ = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        m = FlowModel(copy.deepcopy(flow_cfg))
    missing = m.load_state_dict(state_dict, strict=True)
    m.eval()
    return m
