"""GPU parity of the pflow forward (BASELINE.json configs[4]) through the C ABI (include/pflow.h) via the
SAPF drop-in: against the golden vectors minted from the reference's own SAPF with the REAL pf_hr
checkpoint, and against the CPU oracle on further ragged inputs.

Tolerance: the path is fp32 end to end -> ``rtol 1e-4`` (north_star) with ``atol 1e-4 * max|ref|``
(logits, kinematics) / ``1e-5`` (incidence weights, which live in [0, 1]); the predicted cardinality
(argmax) and the particle mask are bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import pflow_oracle
from superresolutionhep_b200.pflow import PflowLightning
from superresolutionhep_b200.synthetic import synthetic_pflow_events

pytestmark = pytest.mark.gpu


def close(got, ref, rtol, atol, what):
    torch.testing.assert_close(got.detach().float().cpu(), ref.float(), rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "pflow_pf_hr.pt"))


@pytest.fixture(scope="module")
def model(golden):
    lm = PflowLightning({"pf_model": golden["pf_model"], "var_transform": golden["var_transform"]}, {}, inference=True)
    lm.load_state_dict({"net." + k: v for k, v in golden["state_dict"].items()}, strict=True)
    return lm.eval().cuda()


def to_dev(b):
    return {k: v.cuda() for k, v in b.items()}


def check(model, batch, ref_logits, ref_kin, ref_inc, ref_npred):
    logits, kin, inc = model.net(to_dev(batch))
    assert logits.shape == ref_logits.shape and kin.shape == ref_kin.shape and inc.shape == ref_inc.shape
    close(logits, ref_logits, 1e-4, 1e-4 * float(ref_logits.abs().max()), "n_pred_logits")
    assert torch.equal(logits.argmax(-1).cpu(), ref_npred), "cardinality argmax must be bit-exact"
    assert torch.equal(model.net.last_n_pred.cpu().long(), ref_npred)
    close(inc, ref_inc, 1e-4, 1e-5, "inc_weights")
    close(kin, ref_kin, 1e-4, 1e-4 * float(ref_kin.abs().max()), "kin_pred")
    assert float(inc.cpu()[~batch["cell_mask"].unsqueeze(1).expand_as(inc)].abs().sum()) == 0.0     # padded cells stay 0


@pytest.mark.parametrize("case", ["ragged", "sample"])
def test_pflow_matches_reference_golden(model, golden, case):
    c = golden["cases"][case]
    counts = np.array(c["counts"]) if c["counts_given"] else None
    batch = synthetic_pflow_events(len(c["counts"]), seed=c["seed"], counts=counts)
    check(model, batch, c["logits"], c["kin_pred"], c["inc_weights"], c["n_pred"])


@pytest.mark.parametrize("counts", [[1], [1938, 16, 17, 128, 127, 129], [640] * 5])
def test_pflow_matches_oracle(model, golden, counts):
    batch = synthetic_pflow_events(len(counts), seed=77, counts=np.array(counts))
    with torch.no_grad():
        lo, kin, inc, pm = pflow_oracle.sapf_forward(golden["state_dict"], golden["pf_model"], golden["var_transform"], batch)
    check(model, batch, lo, kin, inc, lo.argmax(-1))


def test_pflow_event_order_and_padding_invariance(model):
    """Events are independent: permuting them permutes the outputs; extra padding changes nothing."""
    counts = np.array([48, 304, 16, 200])
    a = synthetic_pflow_events(4, seed=3, counts=counts)
    b = synthetic_pflow_events(4, seed=3, counts=counts, pad_to=512)
    la, ka, ia = model.net(to_dev(a))
    lb, kb, ib = model.net(to_dev(b))
    assert torch.equal(la, lb) and torch.equal(ka, kb) and torch.equal(ia, ib[:, :, : ia.shape[2]])
    perm = torch.tensor([2, 0, 3, 1])
    p = {k: v[perm] for k, v in a.items()}
    lp, kp, ip = model.net(to_dev(p))
    assert torch.equal(lp.cpu(), la.cpu()[perm]) and torch.equal(kp.cpu(), ka.cpu()[perm]) and torch.equal(ip.cpu(), ia.cpu()[perm])
