"""Seeded synthetic checkpoints and events (no dataset / checkpoint can be downloaded).

* ``synthetic_state_dict``: the two SR checkpoints are missing from the reference tree
  (/root/reference/.MISSING_LARGE_BLOBS) and a freshly constructed reference model outputs
  v == 0 (adaLN and the last Linear are zero-initialised, models/flow_model.py:139-154), so
  tests and benchmarks use a deterministic random state_dict with the reference's key names
  and shapes, drawn tensor-by-tensor from one seeded CPU generator.
* ``synthetic_events``: the padded ``collate_graphs`` dict (dataset.py:341-349) filled with
  the distributions of SURVEY.md §8(d).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .config import SrDims


def synthetic_state_dict(dims: SrDims, seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in dims.param_shapes().items():
        if name.endswith(".weight") and len(shape) == 2 and "emb_table" not in name:
            fan_out, fan_in = shape
            a = math.sqrt(6.0 / (fan_in + fan_out))                 # xavier_uniform bound
            if "adaLN_modulation" in name:
                a *= 0.5                                            # keep modulation moderate
            w = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif "emb_table" in name:
            w = torch.randn(shape, generator=g) * 0.5
        elif name.endswith(".weight"):                              # LayerNorm gain
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                                       # biases (Linear and LayerNorm)
            w = 0.05 * torch.randn(shape, generator=g)
        sd[prefix + name] = w.float()
    return sd


def cell_counts(kind: str, n_events: int, rng: np.random.Generator) -> np.ndarray:
    """HR cells per event.  single_e: N(254.18, 37.54) clipped to [124, 804], multiple of 4;
    multipart: Gamma(k=4.06, theta=159.9) clipped to [16, 3280], multiple of 16 (mean 648,
    sigma 322 -- notebooks/data_inspection/*_cardinality.ipynb legends)."""
    if kind == "single_e":
        n = np.clip(rng.normal(254.18, 37.54, n_events), 124, 804)
        return (4 * np.round(n / 4)).astype(np.int64)
    if kind == "multipart":
        n = np.clip(rng.gamma(4.06, 159.9, n_events), 16, 3280)
        return (16 * np.round(n / 16)).astype(np.int64)
    if kind == "pflow":
        k = (316.07 / 223.81) ** 2
        n = np.clip(rng.gamma(k, 316.07 / k, n_events), 16, 1938)
        return (16 * np.round(n / 16)).astype(np.int64)
    raise ValueError(kind)


def synthetic_events(kind: str, n_events: int, seed: int = 1234,
                     counts: Optional[np.ndarray] = None, pad_to: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Padded batch dict with the keys ``FlowModel.forward`` reads (flow_model.py:187-189)."""
    rng = np.random.default_rng(seed)
    res2 = 4 if kind == "single_e" else 16
    n = cell_counts(kind, n_events, rng) if counts is None else np.asarray(counts, dtype=np.int64)
    n_events = len(n)
    nmax = int(max(int(n.max()) if n_events else 1, 1))
    if pad_to is not None:
        nmax = max(nmax, pad_to)
    eta = np.zeros((n_events, nmax, 1), np.float32)
    cosphi = np.zeros_like(eta); sinphi = np.zeros_like(eta); e_proxy = np.zeros_like(eta)
    layer = np.zeros((n_events, nmax, 1), np.int32)
    q_mask = np.zeros((n_events, nmax), bool)
    for i, ni in enumerate(n):
        ni = int(ni)
        c_eta, c_phi = rng.uniform(-2.5, 2.5), rng.uniform(-np.pi, np.pi)
        eta[i, :ni, 0] = (c_eta + rng.uniform(-0.2, 0.2, ni)) / 2.988
        phi = c_phi + rng.uniform(-0.2, 0.2, ni)
        cosphi[i, :ni, 0] = np.cos(phi); sinphi[i, :ni, 0] = np.sin(phi)
        layer[i, :ni, 0] = rng.integers(0, 3, ni)
        n_lr = (ni + res2 - 1) // res2
        e_proxy[i, :ni, 0] = np.repeat(rng.normal(0, 1, n_lr), res2)[:ni]
        q_mask[i, :ni] = True
    return {
        "eta": torch.from_numpy(eta), "cosphi": torch.from_numpy(cosphi),
        "sinphi": torch.from_numpy(sinphi), "e_proxy": torch.from_numpy(e_proxy),
        "layer": torch.from_numpy(layer), "q_mask": torch.from_numpy(q_mask),
        "edge_mask": None,
    }


def synthetic_noise(batch: Dict[str, torch.Tensor], seed: int = 0) -> torch.Tensor:
    """x0 drawn on CPU with the shape of ``e_proxy`` (padded slots included, as the
    reference's ``randn_like(e_proxy)`` does, flow_model.py:319) so CPU and GPU runs share it."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch["e_proxy"].shape, generator=g)


def synthetic_pflow_events(n_events: int, seed: int = 4321, counts: Optional[np.ndarray] = None,
                           pad_to: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Padded batch dict with the keys ``SAPF.forward`` reads (pflow/dataset_pf.py:246-259):
    cells of the SR output above 1 MeV: ``e_raw ~ 1 + Exp(50)`` MeV,
    ``cell_e = (sqrt(e_raw) - 7.35) / 15.65``, ``cell_eta = eta_raw / 2.988``."""
    rng = np.random.default_rng(seed)
    n = cell_counts("pflow", n_events, rng) if counts is None else np.asarray(counts, dtype=np.int64)
    n_events = len(n)
    nmax = int(max(int(n.max()) if n_events else 1, 1))
    if pad_to is not None:
        nmax = max(nmax, pad_to)
    z = lambda: np.zeros((n_events, nmax), np.float32)
    e, eta, phi, cosphi, sinphi, e_raw, eta_raw = z(), z(), z(), z(), z(), z(), z()
    layer = np.zeros((n_events, nmax), np.int32)
    mask = np.zeros((n_events, nmax), bool)
    for i, ni in enumerate(n):
        ni = int(ni)
        n_blob = int(rng.integers(1, 5))                                   # 1-4 particles' worth of energy blobs
        c_eta, c_phi = rng.uniform(-2.3, 2.3, n_blob), rng.uniform(-np.pi, np.pi, n_blob)
        which = rng.integers(0, n_blob, ni)
        er = (1.0 + rng.exponential(50.0, ni)).astype(np.float32)
        etar = (c_eta[which] + rng.normal(0, 0.08, ni)).astype(np.float32)
        ph = (c_phi[which] + rng.normal(0, 0.08, ni)).astype(np.float32)
        e_raw[i, :ni] = er; eta_raw[i, :ni] = etar; phi[i, :ni] = ph
        e[i, :ni] = (np.sqrt(er) - 7.35) / 15.65
        eta[i, :ni] = etar / 2.988
        cosphi[i, :ni] = np.cos(ph); sinphi[i, :ni] = np.sin(ph)
        layer[i, :ni] = rng.integers(0, 3, ni)
        mask[i, :ni] = True
    t = torch.from_numpy
    return {"cell_e": t(e), "cell_eta": t(eta), "cell_phi": t(phi), "cell_cosphi": t(cosphi), "cell_sinphi": t(sinphi),
            "cell_layer": t(layer), "cell_mask": t(mask), "cell_e_raw": t(e_raw), "cell_eta_raw": t(eta_raw)}
