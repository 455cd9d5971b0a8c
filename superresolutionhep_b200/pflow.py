"""Drop-in for the reference's particle-flow model on the inference path.

* ``SAPF(config_pf, inference=False)``: same constructor, same ``state_dict()`` keys and shapes as
  pflow/models/model_pf.py:SAPF (encoder.*, cardinality_predictor.*, kinematics_predictor.*), same
  ``forward(batch) -> (n_pred_logits, kin_pred, inc_weights)`` on the ``collate_fn`` dict of
  pflow/dataset_pf.py:246-259, and ``kinematics_predictor.kin_net.set_trans_dicts`` as called at
  pflow/lightning_pf.py:56-58.
* ``PflowLightning(config_mv, config_t, inference=True)``: the constructor ``inference_pf.py:76``
  uses, exposing ``.net`` (so Lightning checkpoints' ``net.``-prefixed keys load unchanged).

Nothing here computes the network in PyTorch: cells are packed (``cell_mask`` -> ``cu_seqlens``) and
handed to the C ABI of include/pflow.h.  Padded cells of ``inc_weights`` are 0, as in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib
from .flow_model import _Params, _register

try:                                                    # pragma: no cover - not installed in this image
    from pytorch_lightning import LightningModule as _Base
except Exception:                                       # noqa: BLE001
    _Base = nn.Module


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise ValueError(f"unsupported pf_model config for the sm_100a path: {msg}")


def _check_dense(name: str, c: dict) -> None:
    _require(list(c["hidden_layers"]) == [64], f"{name}: one hidden layer of 64 expected")
    _require(c.get("activation") == "LeakyReLU" and not c.get("final_activation"), f"{name}: LeakyReLU, no final activation expected")
    _require(c.get("norm_layer") == "LayerNorm" and not c.get("norm_final_layer"), f"{name}: LayerNorm on the hidden layer only expected")
    _require(not c.get("dropout") and not c.get("context_size"), f"{name}: no dropout, no context expected")


def pflow_dims(config_pf: dict) -> _lib.PflowDimsC:
    """``pf_model`` YAML block -> ``PflowDims`` (include/pflow.h)."""
    enc, card, kin = config_pf["encoder"], config_pf.get("cardinality_predictor"), config_pf.get("kinematics_predictor")
    _require(card is not None and kin is not None, "cardinality_predictor and kinematics_predictor blocks are required")
    _require(enc["transformer"].get("type", "DiT") == "DiT" and kin["transformer"].get("type", "DiT") == "DiT", "DiT transformers expected")
    _require(kin["init_particles"]["type"] == "embedding", "init_particles.type = 'embedding' expected")
    _require(bool(kin.get("use_attn_kinematics")), "use_attn_kinematics: true expected")
    _require(enc["transformer"]["context_size"] == config_pf["h_dim"] and kin["transformer"]["context_size"] == config_pf["h_dim"],
             "transformer context_size = h_dim expected")
    _check_dense("encoder.transformer.dense_config", enc["transformer"]["dense_config"])
    _check_dense("kinematics_predictor.transformer.dense_config", kin["transformer"]["dense_config"])
    _require(card.get("activation") == "LeakyReLU" and not card.get("final_activation") and card.get("norm_layer") == "LayerNorm"
             and not card.get("norm_final_layer") and not card.get("dropout"), "cardinality_predictor: LeakyReLU / LayerNorm Dense expected")
    hid = list(card["hidden_layers"])
    _require(len(hid) <= 4, "cardinality_predictor: at most 4 hidden layers")
    d = _lib.PflowDimsC()
    d.h_dim, d.heads = config_pf["h_dim"], enc["transformer"]["num_heads"]
    _require(kin["transformer"]["num_heads"] == d.heads, "same number of heads in both transformers expected")
    d.enc_layers, d.kin_layers = enc["transformer"]["num_transformer_layers"], kin["transformer"]["num_transformer_layers"]
    d.layer_emb_dim = enc["layer_emb_dim"]
    d.max_particles, d.part_emb_dim = config_pf["max_particles"], kin["init_particles"]["embedding_dim"]
    d.card_n_hidden = len(hid)
    for i, w in enumerate(hid):
        d.card_hidden[i] = w
    d.card_out = config_pf["max_particles"] + 1
    return d


def pflow_param_shapes(d: _lib.PflowDimsC) -> Dict[str, tuple]:
    """Reference ``SAPF.state_dict()`` names and shapes, in the reference's order."""
    H = d.h_dim
    out: Dict[str, tuple] = {}

    def lin(name, o, i):
        out[name + ".weight"] = (o, i); out[name + ".bias"] = (o,)

    def dit(prefix, n):
        for l in range(n):
            p = f"{prefix}.layers.{l}."
            for nm in ("q", "k", "v", "out"):
                lin(p + f"mha.linear_{nm}", H, H)
            lin(p + "dense.net.1", H, H); lin(p + "dense.net.3", H, H)
            for nm in ("norm1", "norm2"):
                out[p + nm + ".weight"] = (H,); out[p + nm + ".bias"] = (H,)
            lin(p + "adaLN_modulation.1", 6 * H, H)
        out[prefix + ".final_norm.weight"] = (H,); out[prefix + ".final_norm.bias"] = (H,)

    out["encoder.layer_emb_net.weight"] = (3, d.layer_emb_dim)
    lin("encoder.cell_init_net.0", H, 4 + d.layer_emb_dim); lin("encoder.cell_init_net.2", H, H)
    dit("encoder.transformer", d.enc_layers)
    win = H
    for i in range(d.card_n_hidden + 1):
        wout = d.card_hidden[i] if i < d.card_n_hidden else d.card_out
        idx = 1 + 3 * i if i < d.card_n_hidden else 3 * d.card_n_hidden      # [LN, Linear, act] per hidden layer, then the last Linear
        lin(f"cardinality_predictor.card_pred_net.net.{idx}", wout, win); win = wout
    out["kinematics_predictor.particle_emb_net.weight"] = (d.max_particles, d.part_emb_dim)
    lin("kinematics_predictor.particle_proj", H, d.part_emb_dim)
    dit("kinematics_predictor.transformer", d.kin_layers)
    lin("kinematics_predictor.kin_net.linear_q", H, H); lin("kinematics_predictor.kin_net.linear_k", H, H)
    return out


_TRANS = {None: 0, "pow(x,m)": 1, "pow(x,m)_signed": 2}
_SCALE = {None: 0, "min_max": 1, "standard": 2}


def _var_transform_c(cfg) -> _lib.PflowVarTransformC:
    """A ``VarTransformation`` object (utility/transformation.py) or its config dict -> C struct."""
    get = (lambda k, dflt=None: cfg.get(k, dflt)) if isinstance(cfg, dict) else (lambda k, dflt=None: getattr(cfg, k, dflt))
    t = _lib.PflowVarTransformC()
    t.trans, t.scale = _TRANS[get("transformation")], _SCALE[get("scale_mode")]
    f = lambda v: float(v) if v is not None else 0.0
    t.m, t.mean, t.std, t.min, t.max = f(get("m", 1.0)), f(get("mean")), f(get("std")) or 1.0, f(get("min")), f(get("max"))
    rng = get("range") or [0.0, 1.0]
    t.lo, t.hi = float(rng[0]), float(rng[1])
    return t


class _KinNet(_Params):
    def set_trans_dicts(self, trans_dicts):                 # kinematics_predictor.py:21-22
        self.trans_dicts = trans_dicts


class SAPF(nn.Module):
    def __init__(self, config_pf: dict, inference: bool = False):
        super().__init__()
        self.config_pf = config_pf
        self.inference = inference
        self.dims = pflow_dims(config_pf)
        for name, shape in pflow_param_shapes(self.dims).items():       # reference order (nn.Module keeps registration order)
            _register(self, name, shape)
        self.kinematics_predictor.kin_net.__class__ = _KinNet           # adds set_trans_dicts (kinematics_predictor.py:21-22)
        self._handle = None
        self._handle_key = None

    # ------------------------------------------------------------------ handle management
    def _device(self) -> torch.device:
        return next(self.parameters()).device

    def _trans(self):
        td = getattr(self.kinematics_predictor.kin_net, "trans_dicts", None)
        if td is None:
            raise RuntimeError("kin_net.set_trans_dicts(...) must be called before forward (pflow/lightning_pf.py:56-58)")
        return td

    def _ensure_handle(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("superresolutionhep_b200.pflow.SAPF runs on CUDA (sm_100a) only; call .cuda() first -- there is no CPU path")
        td = self._trans()
        key = (dev.index if dev.index is not None else torch.cuda.current_device(),
               tuple((p.data_ptr(), p._version) for p in self.parameters()), id(td))
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self.release()
        lib = _lib.load()
        blob = torch.cat([p.detach().float().cpu().reshape(-1) for p in self.state_dict().values()]).contiguous()
        need = lib.pflow_weight_count(C.byref(self.dims))
        if need != blob.numel():
            raise RuntimeError(f"weight blob mismatch: {blob.numel()} floats vs {need} expected")
        tr = (_lib.PflowVarTransformC * 3)(*[_var_transform_c(td[k]) for k in ("pt", "eta", "e")])
        h = C.c_void_p()
        rc = lib.pflow_create(key[0], C.byref(self.dims), blob.data_ptr(), blob.numel(), tr, C.byref(h))
        _lib.check_pflow(lib, None, rc, "pflow_create")
        self._handle, self._handle_key = h, key
        return h

    def release(self):
        if self._handle is not None:
            _lib.load().pflow_destroy(self._handle)
        self._handle = None
        self._handle_key = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(_lib.load().pflow_launch_count(self._handle)) if self._handle is not None else 0

    # ------------------------------------------------------------------ reference surface
    @torch.no_grad()
    def forward(self, batch):
        """pflow/models/model_pf.py:56-74 -> (n_pred_logits (B,5), kin_pred (B,P,4), inc_weights (B,P,Nmax))."""
        h = self._ensure_handle()
        lib = _lib.load()
        dev = self._device()
        mask = batch["cell_mask"].to(dev).bool()
        B, N = mask.shape
        P = self.dims.max_particles
        cu = torch.zeros(B + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(mask.sum(1, dtype=torch.int32).cpu(), 0, dtype=torch.int32)
        T = int(cu[-1])

        def take(key, dtype=torch.float32):
            return batch[key].to(dev).reshape(B, N)[mask].to(dtype).contiguous()

        cols = {k: take("cell_" + k) for k in ("e", "eta", "cosphi", "sinphi", "phi", "e_raw", "eta_raw")}
        layer = take("cell_layer", torch.int32)
        cells = _lib.PflowCells(*(cols[k].data_ptr() for k in ("e", "eta", "cosphi", "sinphi", "phi", "e_raw", "eta_raw")), layer.data_ptr())
        logits = torch.empty(B, self.dims.card_out, dtype=torch.float32, device=dev)
        n_pred = torch.empty(B, dtype=torch.int32, device=dev)
        kin = torch.empty(B, P, 4, dtype=torch.float32, device=dev)
        inc = torch.empty(P, max(T, 1), dtype=torch.float32, device=dev)
        pm = None
        if not self.inference:
            pm = batch["part_mask"].to(dev).to(torch.uint8).contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.pflow_forward(h, C.byref(cells), cu.data_ptr(), B, pm.data_ptr() if pm is not None else None,
                               logits.data_ptr(), n_pred.data_ptr(), kin.data_ptr(), inc.data_ptr(), stream)
        _lib.check_pflow(lib, h, rc, "pflow_forward")
        inc_w = torch.zeros(B, P, N, dtype=torch.float32, device=dev)
        if T:
            inc_w.transpose(0, 1)[:, mask] = inc[:, :T]
        self.last_n_pred = n_pred
        return logits, kin, inc_w


class PflowLightning(_Base):
    """Inference-time surface of pflow/lightning_pf.py:PflowLightning (lines 30-58)."""

    def __init__(self, config_mv, config_t, comet_logger=None, inference=False):
        super().__init__()
        self.config_mv = config_mv
        self.config_t = config_t
        self.comet_logger = comet_logger
        self.net = SAPF(self.config_mv["pf_model"], inference=inference)
        self.transform_dicts = dict(self.config_mv["var_transform"])
        if self.config_mv["pf_model"].get("kinematics_predictor", {}).get("use_attn_kinematics", False):
            self.net.kinematics_predictor.kin_net.set_trans_dicts(self.transform_dicts)

    def forward(self, batch):
        return self.net(batch)
