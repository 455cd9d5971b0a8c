"""CPU restatement of the SR velocity network (TEST INFRASTRUCTURE, not product code).

Functional PyTorch (CPU, fp32 or fp64) restatement of ``FlowModel.forward`` of the
reference, written against a plain ``state_dict`` (reference key names) and the YAML
``flow_model`` block.  It keeps the reference's *algorithm* -- padded ``(B, Nmax, .)``
tensors, materialised ``B x Nmax x Nmax`` padding masks, ``masked_fill(-inf)`` softmax,
context concatenation -- so that its CPU timing is representative of the reference's own
CPU path.  Every function cites the reference lines it follows (paths relative to
``/root/reference``).

Pinned by: tests/test_oracle_vs_reference.py (reference modules imported in the build
container) and tests/golden/sr_*.pt (golden vectors minted from the reference).
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-5           # nn.LayerNorm default, used everywhere (models/dense.py:62)
LEAKY_SLOPE = 0.01      # nn.LeakyReLU default (configs/single_e/model_and_var.yml:27)


# --------------------------------------------------------------------------------------
# Dense plan: which Sequential slots hold which op  (models/dense.py:49-78)
# --------------------------------------------------------------------------------------
def dense_plan(input_size: int, output_size: int, hidden_layers: List[int],
               activation: str = "ReLU", final_activation: Optional[str] = None,
               norm_layer: Optional[str] = None, norm_final_layer=False,
               dropout: float = 0.0, context_size: int = 0) -> List[Tuple]:
    """Slot-by-slot plan of ``Dense.net``; slot index == ``net.<i>`` in the state_dict.

    models/dense.py:49-78: for every layer ``[norm] [dropout] Linear [act]``; the norm is
    skipped on the final layer unless ``norm_final_layer`` is truthy; hidden layers use
    ``activation``, the final layer ``final_activation`` (or nothing)."""
    nodes = [input_size + context_size, *hidden_layers, output_size]
    plan, slot = [], 0
    n = len(nodes) - 1
    for i in range(n):
        final = i == n - 1
        if norm_layer and (norm_final_layer or not final):
            plan.append(("ln", slot, nodes[i])); slot += 1
        if dropout and (norm_final_layer or not final):
            plan.append(("dropout", slot, dropout)); slot += 1
        plan.append(("linear", slot, nodes[i], nodes[i + 1])); slot += 1
        if not final:
            plan.append(("act", slot, activation)); slot += 1
        elif final_activation:
            plan.append(("act", slot, final_activation)); slot += 1
    return plan


def _act(name: str, x: Tensor) -> Tensor:
    if name == "LeakyReLU":
        return F.leaky_relu(x, LEAKY_SLOPE)
    if name == "ReLU":
        return F.relu(x)
    if name == "SiLU":
        return F.silu(x)
    raise NotImplementedError(name)


def attach_context(x: Tensor, context: Tensor) -> Tensor:
    """models/utils.py:84-124 -- broadcast ``context`` over the set dims and concatenate."""
    while context.dim() < x.dim():
        context = context.unsqueeze(1)
    return torch.cat([x, context.expand(*x.shape[:-1], context.shape[-1])], dim=-1)


def dense_apply(sd: Dict[str, Tensor], prefix: str, plan: List[Tuple], x: Tensor,
                context: Optional[Tensor] = None, context_size: int = 0) -> Tensor:
    """models/dense.py:80-83 -- optional context concat, then the Sequential."""
    if context_size:
        x = attach_context(x, context)
    for op in plan:
        kind, slot = op[0], op[1]
        if kind == "ln":
            x = F.layer_norm(x, (op[2],), None, None, LN_EPS)       # elementwise_affine=False
        elif kind == "linear":
            x = F.linear(x, sd[f"{prefix}.net.{slot}.weight"], sd[f"{prefix}.net.{slot}.bias"])
        elif kind == "act":
            x = _act(op[2], x)
        # dropout: identity at inference
    return x


# --------------------------------------------------------------------------------------
# derived dimensions  (models/flow_model.py:29-110)
# --------------------------------------------------------------------------------------
def derive_dims(flow_cfg: dict) -> dict:
    """Reproduces the in-place config overwrites of ``FlowModel.__init__`` without mutating.

    models/flow_model.py:42 (context_size := time_embedding_size), :45/:49/:54/:62
    (embed nets get that context), :57-59 cond_emb_dim, :65 context_size_plus, :69-74
    feat_0 input/context, :101 v_t_input_dim, :108-109 v_t_pred input/context."""
    c = copy.deepcopy(flow_cfg)
    ctx = c["time_embedding_size"]
    c["etaphi_emb"]["context_size"] = ctx
    c["layer_emb"]["dense_config"]["context_size"] = ctx
    c["e_proxy_emb"]["context_size"] = ctx
    c["noisy_input_emb"]["context_size"] = ctx
    cond = (c["etaphi_emb"]["output_size"] + c["layer_emb"]["dense_config"]["output_size"]
            + c["e_proxy_emb"]["output_size"] + 1)
    ctx_plus = ctx + cond
    if c["feat_0_mlp"]["input_size"] == -1:
        c["feat_0_mlp"]["input_size"] = cond + c["noisy_input_emb"]["output_size"]
    c["feat_0_mlp"]["context_size"] = ctx_plus
    h = int(c["h_dim"])
    c["v_t_pred"]["input_size"] = h + cond
    c["v_t_pred"]["context_size"] = ctx_plus
    tr = c["transformer"]
    if tr["type"] != "DiT":
        raise NotImplementedError("only transformer.type == 'DiT' is on the hot path (SURVEY §2 #9)")
    return {
        "cfg": c, "h": h, "heads": tr["num_heads"], "layers": tr["num_transformer_layers"],
        "t_emb": ctx, "cond": cond, "ctx": ctx_plus, "v_in": h + cond,
        "final_modulation": bool(c.get("final_modulation", False)),
        "plans": {
            "etaphi_emb_net": dense_plan(**c["etaphi_emb"]),
            "layer_emb_net": dense_plan(**c["layer_emb"]["dense_config"]),
            "proxy_emb_net": dense_plan(**c["e_proxy_emb"]),
            "noisy_input_emb_net": dense_plan(**c["noisy_input_emb"]),
            "feat_0_mlp": dense_plan(**c["feat_0_mlp"]),
            "layer_dense": dense_plan(input_size=h, output_size=h, **tr["dense_config"]),
            "v_t_pred_net": dense_plan(**c["v_t_pred"]),
        },
    }


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def timestep_embed(sd: Dict[str, Tensor], t: Tensor, freq_dim: int = 256,
                   max_period: float = 10000.0, keep_dtype: bool = False) -> Tensor:
    """models/utils.py:144-166.  cos first, then sin; ``t`` is NOT rescaled; the reference
    force-casts to fp32 (line 157) -- ``keep_dtype=True`` lifts that for an fp64 oracle."""
    half = freq_dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    tt = t[:, None] if keep_dtype else t[:, None].float()
    if keep_dtype:
        freqs = freqs.to(t.dtype)
    args = tt * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    w0, b0 = sd["time_step_embedder.mlp.0.weight"], sd["time_step_embedder.mlp.0.bias"]
    w2, b2 = sd["time_step_embedder.mlp.2.weight"], sd["time_step_embedder.mlp.2.bias"]
    return F.linear(F.silu(F.linear(emb.to(w0.dtype), w0, b0)), w2, b2)


def modulate(x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    """models/diffusion_transformer.py:8-9."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def masked_attention(sd: Dict[str, Tensor], prefix: str, heads: int, q_in: Tensor,
                     k_in: Tensor, q_pad: Optional[Tensor], kv_pad: Optional[Tensor]) -> Tensor:
    """models/attention.py:135-221 + 238-265 + models/utils.py:23-67.

    Separate q/k/v projections, (B, heads, L, hd) split, scores / sqrt(hd), padding mask
    ``q_pad[:, :, None] | kv_pad[:, None, :]`` (True = padded) broadcast over heads,
    masked_fill(-inf) -> softmax -> masked_fill(0), weights @ V, merge heads, out proj."""
    B, Lq, E = q_in.shape
    hd = E // heads
    q = F.linear(q_in, sd[f"{prefix}.linear_q.weight"], sd[f"{prefix}.linear_q.bias"])
    k = F.linear(k_in, sd[f"{prefix}.linear_k.weight"], sd[f"{prefix}.linear_k.bias"])
    v = F.linear(k_in, sd[f"{prefix}.linear_v.weight"], sd[f"{prefix}.linear_v.bias"])
    q = q.view(B, -1, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    mask = None
    if q_pad is not None or kv_pad is not None:
        if q_pad is None:
            q_pad = torch.zeros(q_in.shape[:-1], dtype=torch.bool)
        if kv_pad is None:
            kv_pad = torch.zeros(k_in.shape[:-1], dtype=torch.bool)
        mask = (q_pad.unsqueeze(-1) | kv_pad.unsqueeze(-2)).unsqueeze(1)      # (B,1,Lq,Lk)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)
    if mask is not None:
        scores = scores.masked_fill(mask, -torch.inf)
    w = torch.softmax(scores, dim=-1)
    if mask is not None:
        w = w.masked_fill(mask, 0)            # all-padded query rows: NaN -> 0
    out = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, -1, E)
    return F.linear(out, sd[f"{prefix}.linear_out.weight"], sd[f"{prefix}.linear_out.bias"])


def dit_layer(sd: Dict[str, Tensor], prefix: str, heads: int, dense_pl: Optional[List[Tuple]],
              q: Tensor, q_pad: Optional[Tensor], context: Tensor,
              k: Optional[Tensor] = None, kv_pad: Optional[Tensor] = None) -> Tensor:
    """models/diffusion_transformer.py:30-53 (adaLN-Zero block; chunk order matters)."""
    E = q.shape[-1]
    mod = F.linear(F.silu(context), sd[f"{prefix}.adaLN_modulation.1.weight"],
                   sd[f"{prefix}.adaLN_modulation.1.bias"])
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
    n1w, n1b = sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"]
    if k is None:     # self-attention: keys = modulated LN1(q), kv mask = q mask (attention.py:178-181)
        a_in = modulate(F.layer_norm(q, (E,), n1w, n1b, LN_EPS), sh_a, sc_a)
        attn = masked_attention(sd, f"{prefix}.mha", heads, a_in, a_in, q_pad, q_pad)
    else:             # cross-attention: raw q, keys = modulated LN1(k)  (lines 42-45)
        k_in = modulate(F.layer_norm(k, (E,), n1w, n1b, LN_EPS), sh_a, sc_a)
        attn = masked_attention(sd, f"{prefix}.mha", heads, q, k_in, q_pad, kv_pad)
    q = q + g_a.unsqueeze(1) * attn
    if dense_pl is not None:
        n2w, n2b = sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"]
        m_in = modulate(F.layer_norm(q, (E,), n2w, n2b, LN_EPS), sh_m, sc_m)
        q = q + g_m.unsqueeze(1) * dense_apply(sd, f"{prefix}.dense", dense_pl, m_in)
    return q


def dit_encoder(sd: Dict[str, Tensor], prefix: str, n_layers: int, heads: int,
                dense_pl, q: Tensor, q_pad, context: Tensor,
                k: Optional[Tensor] = None, kv_pad=None, taps: Optional[dict] = None) -> Tensor:
    """models/diffusion_transformer.py:79-87 -- layers then the affine final LayerNorm."""
    for i in range(n_layers):
        q = dit_layer(sd, f"{prefix}.layers.{i}", heads, dense_pl, q, q_pad, context, k, kv_pad)
        if taps is not None:
            taps[f"layer_{i}"] = q
    E = q.shape[-1]
    return F.layer_norm(q, (E,), sd[f"{prefix}.final_norm.weight"], sd[f"{prefix}.final_norm.bias"], LN_EPS)


# --------------------------------------------------------------------------------------
# FlowModel.forward
# --------------------------------------------------------------------------------------
def flow_forward(sd: Dict[str, Tensor], dims: dict, batch: Dict[str, Tensor], x_t: Tensor,
                 t: Tensor, taps: Optional[dict] = None, keep_dtype: bool = False) -> Tensor:
    """models/flow_model.py:167-264.  ``batch`` is the ``collate_graphs`` dict
    (dataset.py:341-349): eta/cosphi/sinphi/e_proxy fp32 (B,Nmax,1), layer int (B,Nmax,1),
    q_mask bool (B,Nmax) True = real.  Returns v_t (B,Nmax,1)."""
    cfg, pl = dims["cfg"], dims["plans"]
    tctx = dims["t_emb"]
    time_emb = timestep_embed(sd, t, keep_dtype=keep_dtype)                                   # :173
    eta, cosphi, sinphi, layer = batch["eta"], batch["cosphi"], batch["sinphi"], batch["layer"]
    e_proxy, q_mask = batch["e_proxy"], batch["q_mask"]

    layer_emb = F.embedding(layer.squeeze(-1).long(), sd["layer_emb_table.weight"])           # :192
    layer_emb = dense_apply(sd, "layer_emb_net", pl["layer_emb_net"], layer_emb, time_emb, tctx)
    etaphi = dense_apply(sd, "etaphi_emb_net", pl["etaphi_emb_net"],
                         torch.cat([eta, cosphi, sinphi], dim=2), time_emb, tctx)             # :194
    proxy = dense_apply(sd, "proxy_emb_net", pl["proxy_emb_net"], e_proxy, time_emb, tctx)    # :195
    cond_feat = torch.cat([etaphi, layer_emb, proxy, e_proxy], dim=-1)                        # :207-209
    cond_global = (cond_feat * q_mask.unsqueeze(-1)).sum(1) / q_mask.sum(1, keepdim=True)     # :210-211
    noisy = dense_apply(sd, "noisy_input_emb_net", pl["noisy_input_emb_net"], x_t, time_emb, tctx)  # :215
    context = torch.cat([time_emb, cond_global], dim=-1)                                      # :222
    feat = dense_apply(sd, "feat_0_mlp", pl["feat_0_mlp"], torch.cat([cond_feat, noisy], dim=-1),
                       context, dims["ctx"])                                                  # :224-228
    if taps is not None:
        taps.update(time_emb=time_emb, cond_feat=cond_feat, context=context, feat_0=feat)
    feat = dit_encoder(sd, "transformer", dims["layers"], dims["heads"], pl["layer_dense"],
                       feat, ~q_mask, context, taps=taps)                                     # :234
    if taps is not None:
        taps["transformer_out"] = feat
    feat = torch.cat([feat, cond_feat], dim=-1)                                               # :241
    if dims["final_modulation"]:                                                              # :243-245
        mod = F.linear(F.silu(context), sd["v_t_adaLN_modulation.1.weight"], sd["v_t_adaLN_modulation.1.bias"])
        sh, sc = mod.chunk(2, dim=1)
        feat = modulate(F.layer_norm(feat, (dims["v_in"],), sd["norm_v_t.weight"], sd["norm_v_t.bias"], LN_EPS), sh, sc)
    v = dense_apply(sd, "v_t_pred_net", pl["v_t_pred_net"], feat, context, dims["ctx"])       # :258
    if taps is not None:
        taps["v_t"] = v
    return v


def make_velocity_fn(sd, dims, batch, keep_dtype: bool = False):
    """The lambda of models/flow_model.py:316-318: scalar t -> ``t * ones(B)``."""
    B = batch["e_proxy"].shape[0]
    dt = batch["e_proxy"].dtype

    def f(t, x):
        tb = (t * torch.ones(B)).to(dt if keep_dtype else torch.float32)
        return flow_forward(sd, dims, batch, x, tb, keep_dtype=keep_dtype)
    return f


def generate_samples(sd, dims, batch, x0: Tensor, n_steps: Optional[int] = None,
                     method: str = "dopri5", ret_seq: bool = False, record: Optional[list] = None):
    """models/flow_model.py:302-329 with the noise ``x0`` passed in explicitly (the reference
    draws ``torch.randn_like(e_proxy)`` at :319).  ``record`` collects (t, v) per evaluation."""
    from .odeint import odeint
    if n_steps is None:
        n_steps = dims["cfg"]["n_steps"]
    f = make_velocity_fn(sd, dims, batch)
    if record is not None:
        g = f

        def f(t, x):  # noqa: E306
            v = g(t, x)
            record.append((float(t), v.clone()))
            return v
    tgrid = torch.linspace(0, 1, n_steps)
    xs = odeint(f, x0, tgrid, method=method, atol=1e-4, rtol=1e-4)
    return xs if ret_seq else xs[-1]
