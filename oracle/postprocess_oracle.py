"""CPU restatement of the post-processing the reference does around its models (TEST INFRASTRUCTURE).

* ``target_inverse``: utility/target_transformation.py:17-33 + utility/transformation.py:50-58 (``inverse``).
* ``fill_high_tree``: the per-event loop of ``Inference.fill_the_dicts2write`` (inference.py:163-287) restricted to the
  branches that depend on the sampler output, written as the reference writes it (event by event, mask indexing).
* ``pflow_cells_from_sr``: pflow/dataset_pf.py:81-92 (energy cut) + :136-147 (derived inputs) + :201-259 (padding).

Pinned by tests/test_postprocess.py against the reference's own ``TargetTransformation`` / ``VarTransformation``
classes imported from /root/reference (build container) -- the loops themselves live in scripts that cannot be
imported here (inference.py and pflow/dataset_pf.py need uproot / pytorch_lightning), so they are restated.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from .pflow_oracle import var_forward


def target_inverse(cfg: dict, nn_out: torch.Tensor, proxy_raw: torch.Tensor) -> torch.Tensor:
    x = nn_out
    if cfg.get("scale_mode") == "standard":                                  # transformation.py:55-58
        x = x * cfg["std"] + cfg["mean"]
    ratio = 1 / (1 + torch.exp(-x))                                          # target_transformation.py:19-24
    ratio = (ratio - cfg["alpha"]) / (1 - 2 * cfg["alpha"])
    return ratio * proxy_raw * cfg["f"]


def fill_high_tree(cfg: dict, batch: Dict[str, torch.Tensor], pred_comp_list: List[torch.Tensor], ts_to_store: Sequence[float],
                   ts_to_store_idx: Sequence[int]) -> Dict[str, List[np.ndarray]]:
    """``pred_comp_list``: one ``(n_steps, B, Nmax, 1)`` tensor per ensemble member (inference.py:145-152)."""
    pred_avg = torch.stack(pred_comp_list, dim=0).mean(dim=0)
    out: Dict[str, List[np.ndarray]] = {}
    add = lambda k, v: out.setdefault(k, []).append(v.squeeze(-1).detach().cpu().numpy())
    for bs_i in range(batch["q_mask"].shape[0]):
        m = batch["q_mask"][bs_i]
        proxy = batch["e_proxy_raw"][bs_i][m]
        e_avg = target_inverse(cfg, pred_avg[-1, bs_i][m], proxy)
        tmp = {"e_pred_avg_raw": e_avg * 1e3, "e_pred_raw": e_avg * 1e3, "raw_nn_pred": pred_avg[-1, bs_i][m]}
        for t, ts_i in zip(ts_to_store, ts_to_store_idx):
            e_t = target_inverse(cfg, pred_avg[ts_i, bs_i][m], proxy)
            tmp[f"e_pred_avg_raw_{t:.2f}"] = e_t * 1e3
            tmp[f"raw_nn_pred_{t:.2f}"] = pred_avg[ts_i, bs_i][m]
            tmp[f"e_pred_raw_{t:.2f}"] = e_t * 1e3
        if len(pred_comp_list) > 1:                                           # inference.py:233-276
            tmp["e_pred_raw"] = torch.zeros_like(e_avg)
            for t in ts_to_store:
                tmp[f"e_pred_raw_{t:.2f}"] = torch.zeros_like(e_avg)
            for pred_comp in pred_comp_list:
                tmp["e_pred_raw"] += target_inverse(cfg, pred_comp[-1, bs_i][m], proxy) * 1e3
                for t, ts_i in zip(ts_to_store, ts_to_store_idx):
                    tmp[f"e_pred_raw_{t:.2f}"] += target_inverse(cfg, pred_comp[ts_i, bs_i][m], proxy) * 1e3
            tmp["e_pred_raw"] /= len(pred_comp_list)
            for t in ts_to_store:
                tmp[f"e_pred_raw_{t:.2f}"] /= len(pred_comp_list)
        for k, v in tmp.items():
            add(k, v)
    return out


def pflow_cells_from_sr(e_pred: List[np.ndarray], eta_raw: List[np.ndarray], phi: List[np.ndarray], layer: List[np.ndarray], var_transform: dict,
                        energy_threshold: float = 1.0) -> Dict[str, torch.Tensor]:
    """Per-event arrays (as read back from ``High_Tree``) -> padded pflow batch."""
    cells = []
    for e, et, ph, la in zip(e_pred, eta_raw, phi, layer):
        m = e > energy_threshold                                             # dataset_pf.py:82
        e, et, ph, la = (torch.from_numpy(np.asarray(a)[m]) for a in (e, et, ph, la))
        cells.append(dict(e_raw=e, eta_raw=et, phi=ph, layer=la, cosphi=torch.cos(ph), sinphi=torch.sin(ph),
                          e=var_forward(var_transform["e"], e), eta=var_forward(var_transform["eta"], et)))       # dataset_pf.py:136-147
    B = len(cells)
    nmax = max(max((len(c["e"]) for c in cells), default=1), 1)
    batch = {"cell_mask": torch.zeros(B, nmax, dtype=torch.bool)}
    for k in ("e", "eta", "cosphi", "sinphi", "phi", "e_raw", "eta_raw"):
        batch["cell_" + k] = torch.zeros(B, nmax)
    batch["cell_layer"] = torch.zeros(B, nmax, dtype=torch.int32)
    for i, c in enumerate(cells):
        n = len(c["e"])
        batch["cell_mask"][i, :n] = True
        for k in ("e", "eta", "cosphi", "sinphi", "phi", "e_raw", "eta_raw"):
            batch["cell_" + k][i, :n] = c[k].float()
        batch["cell_layer"][i, :n] = c["layer"].int()
    return batch
