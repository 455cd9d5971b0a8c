"""Mint the pflow golden vectors from the UNMODIFIED reference ``SAPF`` with the REAL ``pf_hr``
checkpoint (run in the build container; needs /root/reference).

    python tests/golden/make_golden_pflow.py

Output (committed): tests/golden/pflow_pf_hr.pt = the checkpoint's ``state_dict`` (``net.`` prefix
stripped, 333 537 fp32 parameters -- data, not source), the seeds/counts of the synthetic cells and
the reference outputs ``(n_pred_logits, kin_pred, inc_weights)`` of ``SAPF(config, inference=True)``.
"""
import contextlib
import copy
import io
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import                                           # noqa: E402
from superresolutionhep_b200.synthetic import synthetic_pflow_events    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CKPT = "/root/reference/saved_checkpoints/pf_hr/epoch=98-val_loss_to_optimize_on=0.3318.ckpt"
CFG = "/root/reference/saved_checkpoints/pf_hr/config_mv.yml"


def reference_sapf():
    _, SAPF = ref_import.import_reference()
    sys.path.append(ref_import.REF_ROOT)
    from utility.transformation import VarTransformation
    with open(CFG) as fp:
        cfg = yaml.safe_load(fp)
    sd = {k[len("net."):]: v for k, v in torch.load(CKPT, weights_only=True, map_location="cpu")["state_dict"].items()}
    with contextlib.redirect_stdout(io.StringIO()):
        m = SAPF(copy.deepcopy(cfg["pf_model"]), inference=True)
    m.load_state_dict(sd, strict=True)
    m.kinematics_predictor.kin_net.set_trans_dicts({k: VarTransformation(v) for k, v in cfg["var_transform"].items()})
    return m.eval(), sd, cfg


def main():
    m, sd, cfg = reference_sapf()
    cases = {"ragged": dict(seed=11, counts=[16, 48, 304, 1, 640, 128, 129, 33]), "sample": dict(seed=12, counts=None, n=24)}
    out = {"state_dict": {k: v.clone() for k, v in sd.items()}, "pf_model": cfg["pf_model"], "var_transform": cfg["var_transform"], "cases": {}}
    for name, c in cases.items():
        batch = synthetic_pflow_events(c.get("n", 0), seed=c["seed"], counts=None if c["counts"] is None else np.array(c["counts"]))
        with torch.no_grad():
            logits, kin, inc = m(batch)
        out["cases"][name] = dict(seed=c["seed"], counts=batch["cell_mask"].sum(1).tolist(), counts_given=c["counts"] is not None, logits=logits, kin_pred=kin, inc_weights=inc,
                                  n_pred=torch.argmax(logits, -1))
        print(name, "n_pred", torch.argmax(logits, -1).tolist())
    torch.save(out, os.path.join(OUT, "pflow_pf_hr.pt"))


if __name__ == "__main__":
    main()
