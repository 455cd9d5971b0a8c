"""The shipped model configurations as Python dicts.

Users normally pass their own YAML (``configs/*/model_and_var.yml`` of the reference work
unchanged: ``yaml.safe_load`` them and hand ``['flow_model']`` to ``FlowModel``).  Tests,
``smoke()`` and ``bench.py`` run on a GPU box without the reference tree, so the two SR
architectures and the pflow architecture are restated here programmatically.  Values follow
configs/single_e/model_and_var.yml:8-96, configs/multipart/model_and_var.yml:4,95 and
saved_checkpoints/pf_hr/config_mv.yml.
"""
from __future__ import annotations

import copy


def _embed(inp: int, out: int, hidden: int = 64) -> dict:
    return dict(input_size=inp, output_size=out, hidden_layers=[hidden], activation="LeakyReLU",
                final_activation="LeakyReLU", norm_layer="LayerNorm", norm_final_layer=False, dropout=0.0)


def flow_config(kind: str = "single_e") -> dict:
    """``flow_model`` block.  single_e and multipart differ only in the head's final
    LayerNorm (``v_t_pred.norm_final_layer``: "LayerNorm" vs false)."""
    if kind not in ("single_e", "multipart"):
        raise ValueError(kind)
    cfg = dict(
        init_weights=dict(all_linear="xavier_uniform", layer_emb_table="normal",
                          time_step_embedder="normal", ln_modulation="zero", v_t_pred_linear="zero"),
        final_modulation=True, sigma_min=1.0e-5, n_steps=10, time_embedding_size=64, h_dim=256,
        etaphi_emb=_embed(3, 32),
        layer_emb=dict(emb_dim=5, dense_config=_embed(5, 32)),
        e_proxy_emb=_embed(1, 31),
        noisy_input_emb=_embed(1, 64),
        feat_0_mlp=dict(input_size=-1, output_size=256, hidden_layers=[], activation="LeakyReLU",
                        final_activation="LeakyReLU", norm_layer="LayerNorm", norm_final_layer=False,
                        dropout=0.0, context_size=10),
        transformer=dict(type="DiT", num_heads=4, num_transformer_layers=6,
                         dense_config=dict(hidden_layers=[256], activation="LeakyReLU",
                                           final_activation="LeakyReLU", norm_layer="LayerNorm",
                                           norm_final_layer=False, dropout=0.0)),
        v_t_pred=dict(input_size=256, output_size=1, hidden_layers=[128, 64, 32], activation="LeakyReLU",
                      final_activation=None, norm_layer="LayerNorm",
                      norm_final_layer="LayerNorm" if kind == "single_e" else False, dropout=0.0),
    )
    return cfg


def model_and_var_config(kind: str = "single_e") -> dict:
    """Whole ``model_and_var.yml`` equivalent (flow_model + the pieces the boundary reads)."""
    tt = {"single_e": dict(mean=-1.1424768, std=3.616942), "multipart": dict(mean=-3.5069792, std=2.5468976)}[kind]
    return dict(
        name="flow matching", graph_building="all2all", res_factor=2 if kind == "single_e" else 4,
        flow_model=flow_config(kind),
        target_transform=dict(transformation="logit_ratio", f=1.2, alpha=1.0e-6, scale_mode="standard", **tt),
    )


def _pf_dense() -> dict:
    return dict(hidden_layers=[64], activation="LeakyReLU", final_activation=None, norm_layer="LayerNorm",
                norm_final_layer=False, dropout=0.0)


def pflow_config() -> dict:
    """``pf_model`` block of saved_checkpoints/pf_hr/config_mv.yml (the shipped checkpoint)."""
    enc_dense = _pf_dense(); enc_dense["context_size"] = 0
    return dict(
        init_weights=dict(all_linear="xavier_uniform", layer_emb_table="normal", ln_modulation="zero"),
        h_dim=64, max_particles=4,
        encoder=dict(layer_emb_dim=4, transformer=dict(type="DiT", num_heads=4, num_transformer_layers=3,
                                                         dense_config=enc_dense, context_size=64)),
        cardinality_predictor=dict(input_size=64, output_size=None, hidden_layers=[128, 64, 32],
                                   activation="LeakyReLU", final_activation=None, norm_layer="LayerNorm",
                                   norm_final_layer=False, dropout=0.0),
        kinematics_predictor=dict(init_particles=dict(type="embedding", embedding_dim=4),
                                  transformer=dict(type="DiT", num_heads=4, num_transformer_layers=3,
                                                   dense_config=_pf_dense(), context_size=64),
                                  use_attn_kinematics=True),
    )


def pflow_var_transform() -> dict:
    """``var_transform`` of pflow/configs/model_and_var.yml:73-92."""
    return dict(
        eta=dict(transformation=None, scale_mode="min_max", mean=None, std=None, min=-2.988, max=2.988, range=[-1, 1]),
        e=dict(transformation="pow(x,m)", m=0.5, scale_mode="standard", mean=7.35, std=15.65, min=1.0, max=354.27, range=[-1, 1]),
        pt=dict(transformation="pow(x,m)", m=0.5, scale_mode="standard", mean=7.35, std=15.65, min=1.0, max=354.27, range=[-1, 1]),
    )


def clone(cfg: dict) -> dict:
    return copy.deepcopy(cfg)
