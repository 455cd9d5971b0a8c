"""CPU: the pflow oracle (oracle/pflow_oracle.py) against the golden vectors minted from the reference's
own SAPF with the real pf_hr checkpoint (tests/golden/make_golden_pflow.py)."""
import os

import numpy as np
import torch

from oracle import pflow_oracle
from superresolutionhep_b200.synthetic import synthetic_pflow_events


def golden_batch(case):
    counts = np.array(case["counts"]) if case["counts_given"] else None
    return synthetic_pflow_events(len(case["counts"]), seed=case["seed"], counts=counts)


def test_pflow_oracle_matches_reference_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "pflow_pf_hr.pt"))
    assert sum(v.numel() for v in g["state_dict"].values()) == 333537
    for name, case in g["cases"].items():
        batch = golden_batch(case)
        assert batch["cell_mask"].sum(1).tolist() == case["counts"]
        with torch.no_grad():
            logits, kin, inc, part_mask = pflow_oracle.sapf_forward(g["state_dict"], g["pf_model"], g["var_transform"], batch)
        torch.testing.assert_close(logits, case["logits"], rtol=1e-5, atol=1e-5, msg=name)
        torch.testing.assert_close(kin, case["kin_pred"], rtol=1e-5, atol=1e-5, msg=name)
        torch.testing.assert_close(inc, case["inc_weights"], rtol=1e-5, atol=1e-6, msg=name)
        assert torch.equal(logits.argmax(-1), case["n_pred"])
        assert torch.equal(part_mask, torch.arange(4).unsqueeze(0) < case["n_pred"].unsqueeze(1))


def test_pflow_padding_invariance(golden_dir):
    """Same events padded to twice the length: real-cell outputs unchanged (masks do their job)."""
    g = torch.load(os.path.join(golden_dir, "pflow_pf_hr.pt"))
    counts = np.array([16, 80, 33])
    a = synthetic_pflow_events(3, seed=5, counts=counts)
    b = synthetic_pflow_events(3, seed=5, counts=counts, pad_to=200)
    with torch.no_grad():
        la, ka, ia, _ = pflow_oracle.sapf_forward(g["state_dict"], g["pf_model"], g["var_transform"], a)
        lb, kb, ib, _ = pflow_oracle.sapf_forward(g["state_dict"], g["pf_model"], g["var_transform"], b)
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ka, kb, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ia, ib[:, :, :80], rtol=1e-5, atol=1e-6)
