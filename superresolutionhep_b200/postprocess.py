"""The two steps the reference runs in per-event Python loops around the models, on the device
(C ABI in include/srhep_post.h):

* ``ensemble_sample``: ``Inference.run_pred``'s ensemble loop (inference.py:145-152) as ONE batch of
  ``n_ensemble x B`` events -- the conditioning is shared, only the noise differs -- followed by the
  ensemble mean and ``TargetTransformation.inverse`` of every stored grid point in one kernel instead of the
  per-event loop of ``fill_the_dicts2write`` (inference.py:163-287).  Returns packed cells plus a helper that
  splits them into the reference's per-event ``High_Tree`` branches (same names).
* ``sr_to_pflow``: the SR -> pflow hand-off without the ROOT round trip (inference.py:291-310 ->
  pflow/dataset_pf.py:81-92,136-147): threshold the predicted energies, compact per event, derive the scaled
  inputs, return a ``collate_fn``-style batch (pflow/dataset_pf.py:246-259) ready for ``SAPF.forward``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .pflow import _var_transform_c


SR_KEYS = ("eta", "cosphi", "sinphi", "e_proxy", "layer", "q_mask")          # the collate keys FlowModel.forward reads


def stored_steps(n_steps: int, n_steps_to_store: int):
    """inference.py:54-69: grid points of ``linspace(0, 1, n_steps)`` nearest to ``linspace(0, 1, n_store + 1)``
    (last one dropped) -> (times, indices)."""
    used = np.linspace(0, 1, n_steps)
    ts, idx = [], []
    for t in np.linspace(0, 1, n_steps_to_store + 1):
        i = int(np.argmin(np.abs(used - t)))
        ts.append(float(used[i])); idx.append(i)
    return ts[:-1], idx[:-1]


def _target_c(cfg) -> _lib.SrpostTargetTransformC:
    get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    if get("transformation") != "logit_ratio":
        raise ValueError("target_transform.transformation must be 'logit_ratio' (utility/target_transformation.py:10,21)")
    mode = get("scale_mode")
    if mode not in (None, "standard"):
        raise ValueError("target_transform.scale_mode must be 'standard' or null")
    t = _lib.SrpostTargetTransformC()
    t.standard = 1 if mode == "standard" else 0
    t.mean, t.std = float(get("mean") or 0.0), float(get("std") or 1.0)
    t.alpha, t.f = float(get("alpha")), float(get("f"))
    return t


def ensemble_unscale(samples: torch.Tensor, proxy_raw: torch.Tensor, target_cfg, unit: float = 1e3):
    """``samples`` (E, S, T) packed network outputs, ``proxy_raw`` (T) -> (nn_avg, e_pred_avg_raw, e_pred_raw), each (S, T)."""
    if samples.device.type != "cuda":
        raise RuntimeError("postprocess.ensemble_unscale runs on CUDA only -- there is no CPU path")
    lib = _lib.load()
    samples = samples.float().contiguous()
    E, S, T = samples.shape
    proxy_raw = proxy_raw.to(samples.device).float().reshape(-1).contiguous()
    assert proxy_raw.numel() == T
    outs = [torch.empty(S, T, dtype=torch.float32, device=samples.device) for _ in range(3)]
    tt = _target_c(target_cfg)
    stream = torch.cuda.current_stream(samples.device).cuda_stream
    with torch.cuda.device(samples.device):
        rc = lib.srpost_ensemble_unscale(samples.data_ptr(), E, S, T, proxy_raw.data_ptr(), C.byref(tt), unit,
                                         outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), stream)
    _lib.check_post(lib, rc, "srpost_ensemble_unscale")
    return tuple(outs)


@torch.no_grad()
def ensemble_sample(model, batch: Dict[str, torch.Tensor], target_cfg, n_ensemble: int = 1, n_steps: int = 25, n_steps_to_store: int = 0,
                    method: str = "euler", x0: Optional[torch.Tensor] = None) -> Dict[str, object]:
    """``n_ensemble`` samples of every event of ``batch`` in one pass, ensemble-averaged and unscaled on the device.

    ``batch`` must carry ``e_proxy_raw`` besides the keys ``FlowModel.forward`` reads.  ``x0`` (optional,
    (E, B, Nmax, 1)) fixes the noise.  Returns packed tensors keyed like the reference's ``High_Tree`` branches:
    ``raw_nn_pred``, ``e_pred_avg_raw``, ``e_pred_raw`` (final state) and ``*_{t:.2f}`` for the stored grid points,
    plus ``counts`` (cells per event)."""
    dev = next(model.parameters()).device
    mask = batch["q_mask"].to(dev).bool()
    B, N = mask.shape
    E = int(n_ensemble)
    if x0 is None:
        x0 = torch.randn(E, B, N, 1, device=dev)
    x0 = x0.to(dev).reshape(E, B, N, 1)
    # only what FlowModel.forward reads travels (flow_model.py:187-189): collate_graphs also emits edge_mask (B x Nmax^2
    # bytes, never used) and the low_* / target / raw tensors, which must not be tiled E times
    cond = {k: batch[k].to(dev) for k in SR_KEYS}
    cond["q_mask"] = mask
    if method == "dopri5":
        # the reference draws the members in E separate odeint calls (inference.py:146-149): an adaptive solver shares ONE
        # step sequence and error norm across its batch, so the members stay separate calls here too (same binding reused)
        cond["edge_mask"] = None
        xs = torch.stack([model.generate_samples(cond, n_steps=n_steps, method=method, ret_seq=True, x0=x0[i]) for i in range(E)], 1)
        xs = xs.reshape(n_steps, E * B, N, 1)
    else:
        rep = {k: v.repeat(E, *([1] * (v.dim() - 1))) for k, v in cond.items()}
        rep["edge_mask"] = None
        xs = model.generate_samples(rep, n_steps=n_steps, method=method, ret_seq=True, x0=x0.reshape(E * B, N, 1))     # (n_steps, E*B, N, 1)
    ts, idx = stored_steps(n_steps, n_steps_to_store)
    keep = idx + [n_steps - 1]
    sel = xs[keep][..., 0].reshape(len(keep), E, B, N)[..., mask]                       # (S, E, T)
    samples = sel.permute(1, 0, 2).contiguous()                                          # (E, S, T)
    proxy_raw = batch["e_proxy_raw"].to(dev).reshape(B, N)[mask]
    nn_avg, e_avg, e_raw = ensemble_unscale(samples, proxy_raw, target_cfg)
    out: Dict[str, object] = {"counts": mask.sum(1).cpu().numpy(), "raw_nn_pred": nn_avg[-1], "e_pred_avg_raw": e_avg[-1], "e_pred_raw": e_raw[-1],
                              "samples": samples, "stored_times": ts}
    for j, t in enumerate(ts):
        out[f"raw_nn_pred_{t:.2f}"] = nn_avg[j]; out[f"e_pred_avg_raw_{t:.2f}"] = e_avg[j]; out[f"e_pred_raw_{t:.2f}"] = e_raw[j]
    return out


def split_events(packed: torch.Tensor, counts: Sequence[int]) -> List[np.ndarray]:
    """Packed cells -> one numpy array per event (what ``fill_the_dicts2write`` appends per branch)."""
    return [a for a in np.split(packed.detach().cpu().numpy(), np.cumsum(np.asarray(counts))[:-1])]


@torch.no_grad()
def sr_to_pflow(e_pred_raw: torch.Tensor, eta_raw: torch.Tensor, phi: torch.Tensor, layer: torch.Tensor, counts: Sequence[int], var_transform: dict,
                energy_threshold: float = 1.0) -> Dict[str, torch.Tensor]:
    """Packed SR cells (predicted energy in MeV, ``eta_raw``, ``phi``, ``layer``; ``counts`` cells per event) -> padded
    pflow batch dict (pflow/dataset_pf.py:246-259 keys that ``SAPF.forward`` reads) of the cells above ``energy_threshold``."""
    dev = e_pred_raw.device
    if dev.type != "cuda":
        raise RuntimeError("postprocess.sr_to_pflow runs on CUDA only -- there is no CPU path")
    lib = _lib.load()
    counts = np.asarray(counts, dtype=np.int64)
    B, T = len(counts), int(counts.sum())
    cu_in = torch.zeros(B + 1, dtype=torch.int32)
    cu_in[1:] = torch.from_numpy(np.cumsum(counts)).int()
    cu_in = cu_in.to(dev)
    f = lambda v, dt=torch.float32: v.to(dev).reshape(-1).to(dt).contiguous()
    e_pred_raw, eta_raw, phi, layer = f(e_pred_raw), f(eta_raw), f(phi), f(layer, torch.int32)
    names = ("e", "eta", "cosphi", "sinphi", "phi", "e_raw", "eta_raw")
    bufs = {n: torch.empty(max(T, 1), dtype=torch.float32, device=dev) for n in names}
    bufs["layer"] = torch.empty(max(T, 1), dtype=torch.int32, device=dev)
    cu_out = torch.empty(B + 1, dtype=torch.int32, device=dev)
    o = _lib.SrpostPflowOut(*(bufs[n].data_ptr() for n in names), bufs["layer"].data_ptr())
    tr_e, tr_eta = _var_transform_c(var_transform["e"]), _var_transform_c(var_transform["eta"])
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.srpost_select_cells(e_pred_raw.data_ptr(), eta_raw.data_ptr(), phi.data_ptr(), layer.data_ptr(), cu_in.data_ptr(), B, float(energy_threshold),
                                     C.byref(tr_e), C.byref(tr_eta), C.byref(o), cu_out.data_ptr(), stream)
    _lib.check_post(lib, rc, "srpost_select_cells")
    cu = cu_out.cpu().long()
    n_out = (cu[1:] - cu[:-1])
    nmax = max(int(n_out.max()) if B else 1, 1)
    mask = torch.arange(nmax, device=dev).unsqueeze(0) < n_out.to(dev).unsqueeze(1)
    batch = {"cell_mask": mask}
    tot = int(cu[-1])
    for n in names + ("layer",):
        padded = torch.zeros(B, nmax, dtype=bufs[n].dtype, device=dev)
        padded[mask] = bufs[n][:tot]
        batch["cell_" + n] = padded
    return batch
