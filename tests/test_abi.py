"""CPU: the C-ABI library builds, loads and exports every symbol include/srhep.h declares;
host-side logic that needs no device (weight layout, argument validation, loud failure
without a GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

from superresolutionhep_b200 import FlowModel, _lib, build
from superresolutionhep_b200.config import SrDims
from superresolutionhep_b200.default_configs import flow_config
from superresolutionhep_b200.synthetic import synthetic_events, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared_functions(header="srhep.h", prefix="srhep_"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/srhep.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert b"sm_100a" in lib.srhep_version()


def test_every_declared_pflow_symbol_is_exported(lib):
    names = _declared_functions("pflow.h", "pflow_")
    assert len(names) == 6
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pflow.h but not exported"
    assert set(names) == set(_lib.PFLOW_EXPORTS)


def test_every_declared_post_symbol_is_exported(lib):
    names = _declared_functions("srhep_post.h", "srpost_")
    assert set(names) == set(_lib.POST_EXPORTS)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/srhep_post.h but not exported"


def test_pflow_weight_layout_matches_reference_checkpoint(lib, golden_dir):
    """The drop-in SAPF has the reference's state_dict keys, order and shapes (real pf_hr checkpoint)."""
    from superresolutionhep_b200.default_configs import pflow_config, pflow_var_transform
    from superresolutionhep_b200.pflow import PflowLightning, SAPF
    g = torch.load(os.path.join(golden_dir, "pflow_pf_hr.pt"))
    m = SAPF(pflow_config(), inference=True)
    assert list(m.state_dict().keys()) == list(g["state_dict"].keys())
    assert m.load_state_dict(g["state_dict"], strict=True).missing_keys == []
    assert lib.pflow_weight_count(C.byref(m.dims)) == 333537
    lm = PflowLightning({"pf_model": g["pf_model"], "var_transform": g["var_transform"]}, {}, inference=True)
    lm.load_state_dict({"net." + k: v for k, v in g["state_dict"].items()}, strict=True)        # Lightning checkpoint keys
    with pytest.raises(RuntimeError, match="no CPU path"):
        from superresolutionhep_b200.synthetic import synthetic_pflow_events
        lm.net(synthetic_pflow_events(1, counts=[16]))
    bad = pflow_config(); bad["kinematics_predictor"]["use_attn_kinematics"] = False
    with pytest.raises(ValueError):
        SAPF(bad)


@pytest.mark.parametrize("kind", ["single_e", "multipart"])
def test_weight_layout_matches_state_dict(lib, kind):
    d = SrDims.from_config(flow_config(kind))
    n = sum(int(torch.tensor(s).prod()) for s in d.param_shapes().values())
    assert n == 4179439
    dc = d.to_c()
    assert lib.srhep_weight_count(C.byref(dc)) == n + d.freq_dim // 2
    m = FlowModel(flow_config(kind))
    sd = synthetic_state_dict(d, 0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert m.load_state_dict(sd, strict=True).missing_keys == []


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu(lib):
    d = SrDims.from_config(flow_config("single_e"))
    dc = d.to_c()
    n = lib.srhep_weight_count(C.byref(dc))
    blob = torch.zeros(n)
    h = C.c_void_p()
    rc = lib.srhep_create(0, C.byref(dc), blob.data_ptr(), n, 0, C.byref(h))
    assert rc != 0 and not h
    assert b"no CUDA device" in lib.srhep_last_error(None)
    rc = lib.srhep_create(0, C.byref(dc), blob.data_ptr(), n - 1, 0, C.byref(h))
    assert rc == -1 and b"weight blob" in lib.srhep_last_error(None)
    m = FlowModel(flow_config("single_e"))
    batch = synthetic_events("single_e", 1, counts=[8])
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(batch, torch.zeros(1, 8, 1), torch.zeros(1))


def test_unsupported_configs_are_rejected():
    cfg = flow_config("single_e")
    cfg["transformer"]["type"] = "GPT-2+Normformer"
    with pytest.raises(ValueError, match="DiT"):
        FlowModel(cfg)
    cfg = flow_config("single_e")
    cfg["v_t_pred"]["hidden_layers"] = [128, 64]
    with pytest.raises(ValueError):
        FlowModel(cfg)
