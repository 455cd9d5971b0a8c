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


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "srhep.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srhep_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/srhep.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert b"sm_100a" in lib.srhep_version()


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
