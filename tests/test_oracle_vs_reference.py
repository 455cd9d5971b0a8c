"""CPU, build container only: the oracles against the reference's OWN modules imported unmodified from
/root/reference (oracle/ref_import.py).  Skipped where the reference tree is absent (the GPU box); the
committed golden vectors carry the same comparison there (tests/test_oracle_golden.py, test_pflow_oracle.py)."""
import copy

import numpy as np
import pytest
import torch

from oracle import pflow_oracle, ref_import, sr_oracle
from superresolutionhep_b200.config import SrDims
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_noise, synthetic_pflow_events, synthetic_state_dict

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


@pytest.mark.parametrize("kind", ["single_e", "multipart"])
def test_sr_oracle_equals_reference_forward(kind):
    cfg = flow_config(kind)
    sd = synthetic_state_dict(SrDims.from_config(cfg), seed=3)
    ref = ref_import.build_reference_flow_model(cfg, sd)
    batch = synthetic_events(kind, 3, seed=5, counts=np.array([16, 48, 32]))
    x = synthetic_noise(batch, seed=6)
    t = torch.tensor([0.1, 0.5, 0.9])
    with torch.no_grad():
        v_ref = ref(batch, x, t)
        v = sr_oracle.flow_forward(sd, sr_oracle.derive_dims(cfg), batch, x, t)
    mask = batch["q_mask"]
    torch.testing.assert_close(v[mask], v_ref[mask], rtol=1e-5, atol=1e-6)


def test_pflow_oracle_equals_reference_sapf():
    import sys
    sys.path.insert(0, "tests/golden")
    from tests.golden.make_golden_pflow import reference_sapf
    m, sd, cfg = reference_sapf()
    batch = synthetic_pflow_events(5, seed=99, counts=np.array([33, 16, 250, 64, 7]))
    with torch.no_grad():
        lo_r, kin_r, inc_r = m(batch)
        lo, kin, inc, _ = pflow_oracle.sapf_forward(sd, cfg["pf_model"], cfg["var_transform"], batch)
    torch.testing.assert_close(lo, lo_r, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(kin, kin_r, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(inc, inc_r, rtol=1e-5, atol=1e-6)
