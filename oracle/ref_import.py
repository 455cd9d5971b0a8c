"""Import the UNMODIFIED reference modules: from /root/reference in the build container, from the byte-for-byte copy
``oracle/_ref`` (made by oracle/make_ref.py, git-ignored, shipped with the snapshot) on the GPU box.

TEST INFRASTRUCTURE.  Callers must check ``available()``.  ``models/flow_model.py`` imports ``torchdiffeq`` and ``torchcfm`` at module
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

MOUNT = "/root/reference"
COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _has(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "models", "flow_model.py"))


REF_ROOT = MOUNT if _has(MOUNT) or not _has(COPY) else COPY


def available() -> bool:
    return _has(REF_ROOT)


def source() -> str:
    """'mount' (/root/reference), 'copy' (oracle/_ref) or 'absent'."""
    return "absent" if not available() else ("mount" if REF_ROOT == MOUNT else "copy")


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
        raise RuntimeError("neither /root/reference nor oracle/_ref is present on this machine")
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
