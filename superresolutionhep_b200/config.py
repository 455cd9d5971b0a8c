"""YAML -> dimensions for the SR velocity network and the pflow network.

Follows the dimension rules of the reference constructor (models/flow_model.py:29-110):
the YAML's ``context_size`` / ``input_size`` entries for the embedding nets, ``feat_0_mlp``
and ``v_t_pred`` are *overwritten* there; the same overwrite rules are applied here on a
copy (the reference mutates the dict it is given).

The CUDA path implements one architecture family (the one every shipped config uses);
``SrDims.from_config`` raises ``ValueError`` for anything else instead of silently computing
something different.
"""
from __future__ import annotations

import copy
import ctypes
from dataclasses import dataclass
from typing import Dict, List, Tuple


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise ValueError(f"unsupported flow_model config for the sm_100a path: {msg}")


def _check_embed(name: str, c: dict) -> None:
    _require(len(c["hidden_layers"]) == 1, f"{name}: exactly one hidden layer expected")
    _require(c.get("activation") == "LeakyReLU" and c.get("final_activation") == "LeakyReLU",
             f"{name}: LeakyReLU / LeakyReLU expected")
    _require(c.get("norm_layer") == "LayerNorm" and not c.get("norm_final_layer"),
             f"{name}: LayerNorm on hidden layers only expected")
    _require(not c.get("dropout"), f"{name}: dropout must be 0")


class SrDimsC(ctypes.Structure):
    """Mirror of ``SrhepDims`` in include/srhep.h (field order and types must match)."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "h_dim", "heads", "layers", "t_emb", "freq_dim",
        "etaphi_in", "etaphi_hid", "etaphi_out",
        "layer_emb_dim", "layer_hid", "layer_out",
        "proxy_hid", "proxy_out",
        "noisy_hid", "noisy_out",
        "mlp_hid",
        "head_h1", "head_h2", "head_h3", "head_final_ln",
        "cond", "ctx", "v_in")]


@dataclass
class SrDims:
    h_dim: int
    heads: int
    layers: int
    t_emb: int
    freq_dim: int
    etaphi_in: int
    etaphi_hid: int
    etaphi_out: int
    layer_emb_dim: int
    layer_hid: int
    layer_out: int
    proxy_hid: int
    proxy_out: int
    noisy_hid: int
    noisy_out: int
    mlp_hid: int
    head_h1: int
    head_h2: int
    head_h3: int
    head_final_ln: int
    cond: int = 0       # etaphi_out + layer_out + proxy_out + 1        (flow_model.py:57-59)
    ctx: int = 0        # t_emb + cond                                  (flow_model.py:65)
    v_in: int = 0       # h_dim + cond                                  (flow_model.py:101)
    n_steps: int = 10

    def __post_init__(self):
        self.cond = self.etaphi_out + self.layer_out + self.proxy_out + 1
        self.ctx = self.t_emb + self.cond
        self.v_in = self.h_dim + self.cond

    @property
    def mod_width(self) -> int:
        """Columns of the batched adaLN GEMM: 6*h per layer + 2*v_in for the head."""
        return 6 * self.h_dim * self.layers + 2 * self.v_in

    @classmethod
    def from_config(cls, flow_cfg: dict) -> "SrDims":
        c = copy.deepcopy(flow_cfg)
        tr = c["transformer"]
        _require(tr["type"] == "DiT", "transformer.type must be 'DiT' (SURVEY §2 #9)")
        _require(bool(c.get("final_modulation", False)), "final_modulation must be true")
        for name in ("etaphi_emb", "e_proxy_emb", "noisy_input_emb"):
            _check_embed(name, c[name])
        _check_embed("layer_emb.dense_config", c["layer_emb"]["dense_config"])
        _require(c["layer_emb"]["dense_config"]["input_size"] == c["layer_emb"]["emb_dim"],
                 "layer_emb dense input_size must equal emb_dim")
        _require(c["e_proxy_emb"]["input_size"] == 1 and c["noisy_input_emb"]["input_size"] == 1,
                 "e_proxy / noisy_input embeddings take one scalar per cell")
        f0 = c["feat_0_mlp"]
        _require(len(f0["hidden_layers"]) == 0 and f0.get("final_activation") == "LeakyReLU",
                 "feat_0_mlp must be a single Linear + LeakyReLU")
        _require(f0["output_size"] == int(c["h_dim"]), "feat_0_mlp output must be h_dim")
        if f0["input_size"] != -1:
            expect = (c["etaphi_emb"]["output_size"] + c["layer_emb"]["dense_config"]["output_size"]
                      + c["e_proxy_emb"]["output_size"] + 1 + c["noisy_input_emb"]["output_size"])
            _require(f0["input_size"] == expect, "feat_0_mlp.input_size must be -1 or cond+noisy")
        dc = tr["dense_config"]
        _require(len(dc["hidden_layers"]) == 1 and dc.get("activation") == "LeakyReLU"
                 and dc.get("final_activation") == "LeakyReLU" and dc.get("norm_layer") == "LayerNorm"
                 and not dc.get("norm_final_layer") and not dc.get("dropout")
                 and not dc.get("context_size"),
                 "transformer.dense_config: LN -> Linear -> LeakyReLU -> Linear -> LeakyReLU expected")
        vp = c["v_t_pred"]
        _require(len(vp["hidden_layers"]) == 3 and vp["output_size"] == 1
                 and vp.get("activation") == "LeakyReLU" and not vp.get("final_activation")
                 and vp.get("norm_layer") == "LayerNorm" and not vp.get("dropout"),
                 "v_t_pred: three LeakyReLU hidden layers with LayerNorm and a scalar output expected")
        return cls(
            h_dim=int(c["h_dim"]), heads=tr["num_heads"], layers=tr["num_transformer_layers"],
            t_emb=c["time_embedding_size"], freq_dim=256,
            etaphi_in=c["etaphi_emb"]["input_size"], etaphi_hid=c["etaphi_emb"]["hidden_layers"][0],
            etaphi_out=c["etaphi_emb"]["output_size"],
            layer_emb_dim=c["layer_emb"]["emb_dim"],
            layer_hid=c["layer_emb"]["dense_config"]["hidden_layers"][0],
            layer_out=c["layer_emb"]["dense_config"]["output_size"],
            proxy_hid=c["e_proxy_emb"]["hidden_layers"][0], proxy_out=c["e_proxy_emb"]["output_size"],
            noisy_hid=c["noisy_input_emb"]["hidden_layers"][0], noisy_out=c["noisy_input_emb"]["output_size"],
            mlp_hid=dc["hidden_layers"][0],
            head_h1=vp["hidden_layers"][0], head_h2=vp["hidden_layers"][1], head_h3=vp["hidden_layers"][2],
            head_final_ln=1 if vp.get("norm_final_layer") else 0,
            n_steps=int(c.get("n_steps", 10)),
        )

    def to_c(self) -> SrDimsC:
        s = SrDimsC()
        for name, _ in SrDimsC._fields_:
            setattr(s, name, int(getattr(self, name)))
        return s

    # ---------------------------------------------------------------- state_dict layout
    def head_slots(self) -> Tuple[int, int, int, int]:
        """Sequential slots of the four Linear layers of ``v_t_pred_net`` (dense.py:49-78):
        LN,Lin,Act | LN,Lin,Act | LN,Lin,Act | [LN],Lin  ->  1,4,7 and 10 (or 9 without the
        final LN, as in configs/multipart/model_and_var.yml:95)."""
        return (1, 4, 7, 10 if self.head_final_ln else 9)

    def param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """Every tensor of the reference ``FlowModel.state_dict()`` and its shape."""
        d = self
        s: Dict[str, Tuple[int, ...]] = {}

        def lin(prefix: str, out: int, inp: int):
            s[f"{prefix}.weight"] = (out, inp)
            s[f"{prefix}.bias"] = (out,)

        lin("time_step_embedder.mlp.0", d.t_emb, d.freq_dim)
        lin("time_step_embedder.mlp.2", d.t_emb, d.t_emb)
        lin("etaphi_emb_net.net.1", d.etaphi_hid, d.etaphi_in + d.t_emb)
        lin("etaphi_emb_net.net.3", d.etaphi_out, d.etaphi_hid)
        s["layer_emb_table.weight"] = (3, d.layer_emb_dim)
        lin("layer_emb_net.net.1", d.layer_hid, d.layer_emb_dim + d.t_emb)
        lin("layer_emb_net.net.3", d.layer_out, d.layer_hid)
        lin("proxy_emb_net.net.1", d.proxy_hid, 1 + d.t_emb)
        lin("proxy_emb_net.net.3", d.proxy_out, d.proxy_hid)
        lin("noisy_input_emb_net.net.1", d.noisy_hid, 1 + d.t_emb)
        lin("noisy_input_emb_net.net.3", d.noisy_out, d.noisy_hid)
        lin("feat_0_mlp.net.0", d.h_dim, d.cond + d.noisy_out + d.ctx)
        for i in range(d.layers):
            p = f"transformer.layers.{i}"
            for nm in ("q", "k", "v", "out"):
                lin(f"{p}.mha.linear_{nm}", d.h_dim, d.h_dim)
            lin(f"{p}.dense.net.1", d.mlp_hid, d.h_dim)
            lin(f"{p}.dense.net.3", d.h_dim, d.mlp_hid)
            s[f"{p}.norm1.weight"] = (d.h_dim,); s[f"{p}.norm1.bias"] = (d.h_dim,)
            s[f"{p}.norm2.weight"] = (d.h_dim,); s[f"{p}.norm2.bias"] = (d.h_dim,)
            lin(f"{p}.adaLN_modulation.1", 6 * d.h_dim, d.ctx)
        s["transformer.final_norm.weight"] = (d.h_dim,); s["transformer.final_norm.bias"] = (d.h_dim,)
        s["norm_v_t.weight"] = (d.v_in,); s["norm_v_t.bias"] = (d.v_in,)
        lin("v_t_adaLN_modulation.1", 2 * d.v_in, d.ctx)
        s1, s2, s3, s4 = d.head_slots()
        lin(f"v_t_pred_net.net.{s1}", d.head_h1, d.v_in + d.ctx)
        lin(f"v_t_pred_net.net.{s2}", d.head_h2, d.head_h1)
        lin(f"v_t_pred_net.net.{s3}", d.head_h3, d.head_h2)
        lin(f"v_t_pred_net.net.{s4}", 1, d.head_h3)
        return s

    def param_order(self) -> List[str]:
        """Canonical order of tensors in the packed fp32 weight blob handed to the C ABI."""
        return list(self.param_shapes().keys())
