"""CPU restatement of the particle-flow network SAPF (TEST INFRASTRUCTURE, not product code).

Functional PyTorch (CPU, fp32 or fp64) restatement of ``SAPF.forward`` of the reference
(pflow/models/model_pf.py:56-74) written against a plain ``state_dict`` (reference key names,
without the Lightning ``net.`` prefix) and the YAML ``pf_model`` block.  It keeps the reference's
algorithm: padded ``(B, Nmax, .)`` tensors, materialised padding masks, ``masked_fill(-inf)``
softmax.  Every function cites the reference lines it follows (paths relative to /root/reference).

Pinned by: tests/test_oracle_vs_reference.py (reference ``SAPF`` imported in the build container,
real ``saved_checkpoints/pf_hr`` weights) and tests/golden/pflow_pf_hr.pt (golden vectors minted
from the reference with those weights).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-5


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(x, w=None, b=None):
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def masked_mean(x: Tensor, mask: Tensor) -> Tensor:
    """pflow/models/encoder.py:54-55, cardinality_predictor.py:18-19, kinematics_predictor.py:119-120."""
    f = mask.unsqueeze(-1)
    return torch.sum(x * f, dim=1) / torch.sum(f, dim=1)


def masked_softmax(x: Tensor, mask: Optional[Tensor], dim: int = -1) -> Tensor:
    """models/utils.py:23-34 (mask True = padded)."""
    if mask is not None:
        while mask.dim() < x.dim():
            mask = mask.unsqueeze(1)
        x = x.masked_fill(mask, -torch.inf)
    x = torch.softmax(x, dim=dim)
    if mask is not None:
        x = x.masked_fill(mask, 0)
    return x


def mha(sd, pre, heads, q, k, q_pad, kv_pad):
    """models/attention.py:135-221 without edge features; v = k (line 178)."""
    B, Lq, E = q.shape
    hd = E // heads
    mask = q_pad.unsqueeze(-1) | kv_pad.unsqueeze(-2)                        # models/utils.py:38-67
    qp = _lin(sd, pre + "linear_q", q).view(B, -1, heads, hd).transpose(1, 2)
    kp = _lin(sd, pre + "linear_k", k).view(B, -1, heads, hd).transpose(1, 2)
    vp = _lin(sd, pre + "linear_v", k).view(B, -1, heads, hd).transpose(1, 2)
    scores = torch.matmul(qp, kp.transpose(-2, -1)) / math.sqrt(hd)          # models/attention.py:250
    w = masked_softmax(scores, mask, dim=-1)
    out = torch.matmul(w, vp).transpose(1, 2).contiguous().view(B, -1, E)
    return _lin(sd, pre + "linear_out", out)


def dit_layer(sd, pre, heads, q, q_pad, context, k=None, kv_pad=None):
    """models/diffusion_transformer.py:30-53 with the pflow Dense: LN(no affine) -> Linear ->
    LeakyReLU -> Linear, no final activation (pflow/configs/model_and_var.yml:18-25)."""
    mod = _lin(sd, pre + "adaLN_modulation.1", F.silu(context))
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    modulate = lambda x, sh, sc: x * (1 + sc.unsqueeze(1)) + sh.unsqueeze(1)
    n1w, n1b = sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]
    if k is None:
        x = modulate(_ln(q, n1w, n1b), shift_msa, scale_msa)
        attn = mha(sd, pre + "mha.", heads, x, x, q_pad, q_pad)
    else:                                                                      # cross-attention: norm1 acts on the KEYS (line 43-45)
        kk = modulate(_ln(k, n1w, n1b), shift_msa, scale_msa)
        attn = mha(sd, pre + "mha.", heads, q, kk, q_pad, kv_pad)
    q = q + gate_msa.unsqueeze(1) * attn
    y = modulate(_ln(q, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"]), shift_mlp, scale_mlp)
    y = _lin(sd, pre + "dense.net.3", F.leaky_relu(_lin(sd, pre + "dense.net.1", _ln(y)), 0.01))
    return q + gate_mlp.unsqueeze(1) * y


def dit_encoder(sd, pre, n_layers, heads, q, q_pad, context, k=None, kv_pad=None):
    """models/diffusion_transformer.py:79-87."""
    for l in range(n_layers):
        q = dit_layer(sd, f"{pre}layers.{l}.", heads, q, q_pad, context, k, kv_pad)
    return _ln(q, sd[pre + "final_norm.weight"], sd[pre + "final_norm.bias"])


def var_forward(cfg: dict, x: Tensor) -> Tensor:
    """utility/transformation.py:20-31,39-48,62-65 (trans then scale)."""
    t = cfg.get("transformation")
    if t == "pow(x,m)":
        x = torch.pow(x, cfg["m"])
    elif t == "pow(x,m)_signed":
        x = ((x >= 0) * 2 - 1) * (abs(x) ** cfg["m"])
    s = cfg.get("scale_mode")
    if s == "min_max":
        lo, hi = cfg["range"]
        x = (x - cfg["min"]) / (cfg["max"] - cfg["min"]) * (hi - lo) + lo
    elif s == "standard":
        x = (x - cfg["mean"]) / cfg["std"]
    return x


def card_hidden_count(sd) -> int:
    n = 0
    while f"cardinality_predictor.card_pred_net.net.{1 + 3 * n}.weight" in sd:
        n += 1
    return n


def sapf_forward(sd: Dict[str, Tensor], cfg: dict, var_transform: dict, batch: Dict[str, Tensor],
                 inference: bool = True) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """pflow/models/model_pf.py:56-74 -> (n_pred_logits, kin_pred, inc_weights, part_mask)."""
    heads_e = cfg["encoder"]["transformer"]["num_heads"]
    heads_k = cfg["kinematics_predictor"]["transformer"]["num_heads"]
    cell_mask = batch["cell_mask"]
    # ---- Encoder.forward (pflow/models/encoder.py:38-58)
    layer_emb = sd["encoder.layer_emb_net.weight"][batch["cell_layer"].long()]
    feat0 = torch.cat([batch["cell_e"].unsqueeze(-1), batch["cell_eta"].unsqueeze(-1), batch["cell_cosphi"].unsqueeze(-1),
                       batch["cell_sinphi"].unsqueeze(-1), layer_emb], dim=-1)
    cell_feat = _lin(sd, "encoder.cell_init_net.2", F.leaky_relu(_lin(sd, "encoder.cell_init_net.0", feat0), 0.01))
    ctx = masked_mean(cell_feat, cell_mask)
    feat = dit_encoder(sd, "encoder.transformer.", cfg["encoder"]["transformer"]["num_transformer_layers"], heads_e,
                       cell_feat, ~cell_mask, ctx)
    # ---- CardinalityPredictor.forward (pflow/models/cardinality_predictor.py:17-22)
    g = masked_mean(feat, cell_mask)
    x = g
    nh = card_hidden_count(sd)                        # hidden Linears at net.{1,4,7,...}; the last one (no LN, no act) at net.{3*nh}
    for i in range(nh):
        x = F.leaky_relu(_lin(sd, f"cardinality_predictor.card_pred_net.net.{1 + 3 * i}", _ln(x)), 0.01)
    logits = _lin(sd, f"cardinality_predictor.card_pred_net.net.{3 * nh}", x)
    # ---- SAPF.forward inference branch (model_pf.py:65-67)
    n_max = cfg["max_particles"]
    if inference:
        n_pred = torch.argmax(logits, dim=-1)
        part_mask = torch.arange(n_max).unsqueeze(0) < n_pred.unsqueeze(1)
    else:
        part_mask = batch["part_mask"]
    # ---- KinematicsPredictor.forward (pflow/models/kinematics_predictor.py:99-135)
    B = feat.shape[0]
    pe = F.linear(sd["kinematics_predictor.particle_emb_net.weight"], sd["kinematics_predictor.particle_proj.weight"],
                  sd["kinematics_predictor.particle_proj.bias"]).unsqueeze(0).repeat(B, 1, 1)
    part = dit_encoder(sd, "kinematics_predictor.transformer.", cfg["kinematics_predictor"]["transformer"]["num_transformer_layers"],
                       heads_k, pe, ~part_mask, g, k=feat, kv_pad=~cell_mask)
    # ---- AttnKinematicNet.forward (kinematics_predictor.py:24-57)
    qp = _lin(sd, "kinematics_predictor.kin_net.linear_q", part)
    kp = _lin(sd, "kinematics_predictor.kin_net.linear_k", feat)
    mask = (~part_mask).unsqueeze(-1) | (~cell_mask).unsqueeze(-2)
    scores = torch.matmul(qp, kp.transpose(-2, -1)) / math.sqrt(feat.shape[-1])
    inc_w = masked_softmax(scores, mask, dim=1)                               # over particles
    e_inc = inc_w * batch["cell_e_raw"].unsqueeze(1)
    tot = e_inc.sum(dim=2, keepdim=True)
    inc = e_inc / (tot + (tot == 0))
    eta_raw = (inc * batch["cell_eta_raw"].unsqueeze(1)).sum(-1)
    phi = (inc * batch["cell_phi"].unsqueeze(1)).sum(-1)
    e_raw = e_inc.sum(-1)
    pt_raw = e_raw / torch.cosh(eta_raw)
    kin = torch.stack([var_forward(var_transform["pt"], pt_raw), var_forward(var_transform["eta"], eta_raw), phi,
                       var_forward(var_transform["e"], e_raw)], dim=-1)
    return logits, kin, inc_w, part_mask
